cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
python bench.py --steps 5 --warmup 3 --no-latency > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v6.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
PY
head -6 gpurun_out/bench_v6.log
python bench.py --workload lora_step --steps 5 --warmup 3 > gpurun_out/lora_v6.json 2> gpurun_out/lora_v6.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/lora_v6.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'])
PY
head -5 gpurun_out/lora_v6.log
