cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "zoe or tail or gemm" 2>&1 | tail -3
python tools/zoe_tail_perf.py 2>&1 | tail -12 | tee gpurun_out/zoe_tail_now.log
