cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_lora_step.py -x -q -m gpu -k "bwd or backward or lora or grad" 2>&1 | tail -3
python tools/attn_bwd_perf.py 2>&1 | head -3 | tee gpurun_out/attn_bwd_now.log
