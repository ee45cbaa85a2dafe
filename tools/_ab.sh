cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
python tools/gemm_perf.py attn_ 2>&1 | grep attn_ | tee gpurun_out/attn_perf.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_now.json 2> gpurun_out/bench_now.log
tail -c 1500 gpurun_out/bench_now.json
