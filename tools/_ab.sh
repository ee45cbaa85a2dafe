cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_lora_step.py tests/test_kernels_gpu.py -x -q -m gpu -k "tn or lora or grad" 2>&1 | tail -3
python tools/attn_bwd_perf.py 2>&1 | tail -5 | tee gpurun_out/tn_now.log
