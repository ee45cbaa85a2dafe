cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention" 2>&1 | tail -15 > gpurun_out/t_attn.log
cat gpurun_out/t_attn.log
python tools/gemm_perf.py attn_ 2>&1 | grep attn_ | tee gpurun_out/attn_perf.log
python tools/attn_timeline.py > gpurun_out/attn_tl3.log 2>&1
