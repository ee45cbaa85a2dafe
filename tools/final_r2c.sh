#!/bin/bash
# Round-2 final pass (after the hi/lo decode chain + sliding-window predicate): GPU suite, smoke, the three bench lines, then --
# only after those plain commands exited -- the ncu launch list of one predict_action step and one `--set full` capture of the
# hi/lo gate/up skinny GEMM.  Outputs under gpurun_out/, copied into profiles/ by hand.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/r2c_tests.log 2>&1; echo "tests rc=$? $(tail -1 $OUT/r2c_tests.log)"
timeout 300 python __graft_entry__.py smoke > $OUT/r2c_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $OUT/r2c_smoke.log)"
timeout 600 python bench.py > $OUT/bench_r2_v10.json 2> $OUT/bench_r2_v10_breakdown.txt; echo "bench rc=$?"
timeout 300 python bench.py --workload latency_bs1 > $OUT/latency_bs1_r2_v10.json 2> $OUT/latency_bs1_r2_v10.log; echo "latency rc=$?"
timeout 400 python bench.py --workload lora_step > $OUT/lora_step_r2_v10.json 2> $OUT/lora_step_r2_v10_breakdown.txt; echo "lora rc=$?"
SVLA_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svla_ -c 6500 --csv --log-file $OUT/launches_r2c_predict.csv \
  python bench.py --quick --steps 1 --warmup 1 > $OUT/ncu_r2c_launch_predict.log 2>&1
echo "predict launch list rc=$?"
SVLA_NO_GRAPHS=1 timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^svla_gemm_skinny_kernel$' -s 2 -c 1 -f -o $OUT/ncu_r2c_skinny_hilo_gateup \
  python bench.py --quick --steps 1 --warmup 1 > $OUT/ncu_r2c_skinny_hilo_gateup.log 2>&1
echo "skinny capture rc=$? $(ls $OUT/ncu_r2c_skinny_hilo_gateup.ncu-rep 2>/dev/null)"
