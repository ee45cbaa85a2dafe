cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_lora_step.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload lora_step --steps 5 --warmup 3 > gpurun_out/lora_v8.json 2> gpurun_out/lora_v8.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/lora_v8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])
PY
head -8 gpurun_out/lora_v8.log
