cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -3
python tools/gemm_perf.py attn_ 2>&1 | grep attn_ | tee gpurun_out/attn_perf.log
