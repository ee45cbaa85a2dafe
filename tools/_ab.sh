cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or decode" 2>&1 | tail -3
python tools/decode_mega_check.py 2>&1 | grep "chain" | tee gpurun_out/decode_now.log
