cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --workload lora_step --steps 5 --warmup 3 > gpurun_out/lora_v9.json 2> gpurun_out/lora_v9.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/lora_v9.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])
PY
head -6 gpurun_out/lora_v9.log
python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v9.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['latency_bs1_ms_p50'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
