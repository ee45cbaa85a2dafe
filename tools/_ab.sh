cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.log; tail -c 300 gpurun_out/bench_v4.json
python bench.py --workload latency_bs1 > gpurun_out/latency_v4.json 2> gpurun_out/latency_v4.log
python bench.py --workload lora_step --steps 5 --warmup 3 > gpurun_out/lora_v4.json 2> gpurun_out/lora_v4.log
python tools/attn_timeline.py > gpurun_out/attn_tl_final.log 2>&1
export SVLA_NO_GRAPHS=1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:^svla_gemm_tcgen05_kernel$" -s 289 -c 1 -f -o gpurun_out/ncu_r2b_gemm_gateup python bench.py --quick --steps 1 --warmup 1 > gpurun_out/ncu_r2b_gemm_gateup.log 2>&1
echo ncu rc=$?
