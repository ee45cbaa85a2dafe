cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -x -q -k "tok" 2>&1 | tail -5
python tools/micro_bench.py 2>&1 | head -4 | tee gpurun_out/micro_r2.log
