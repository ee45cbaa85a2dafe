#!/bin/bash
# Round-2 ncu evidence (run on the GPU box after the plain commands have exited 0 without ncu):
#   launch lists (gpu__time_duration.sum, --clock-control none) of one predict_action step and one LoRA step,
#   `--set full` captures of the dominant kernels of both.
OUT=gpurun_out
mkdir -p $OUT
python bench.py --quick --steps 1 --warmup 1 > $OUT/r2_quick.log 2>&1 || exit 1
python bench.py --workload lora_step --quick --steps 1 --warmup 1 > $OUT/r2_quick_lora.log 2>&1 || exit 1
SVLA_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svla_ -c 6000 --csv --log-file $OUT/launches_r2_predict.csv \
  python bench.py --quick --steps 1 --warmup 1 > $OUT/ncu_r2_launch_predict.log 2>&1
echo "predict launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svla_ -c 14000 --csv --log-file $OUT/launches_r2_lora.csv \
  python bench.py --workload lora_step --quick --steps 1 --warmup 1 > $OUT/ncu_r2_launch_lora.log 2>&1
echo "lora launch list rc=$?"
cap() {  # name regex skip command...
  local n=$1 rx=$2 sk=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $sk -c 1 -f -o $OUT/ncu_r2_$n "$@" > $OUT/ncu_r2_$n.log 2>&1
  echo "$n rc=$? $(ls $OUT/ncu_r2_$n.ncu-rep 2>/dev/null)"
}
export SVLA_NO_GRAPHS=1
cap gemm_gateup '^svla_gemm_tcgen05_kernel$' 288 python bench.py --quick --steps 1 --warmup 1
cap attn_siglip_tc '^svla_flash_attn_tc_kernel$' 2 python bench.py --quick --steps 1 --warmup 1
cap attn_beit_tc '^svla_flash_attn_tc_kernel$' 29 python bench.py --quick --steps 1 --warmup 1
cap skinny_gateup '^svla_gemm_skinny_kernel$' 3 python bench.py --quick --steps 1 --warmup 1
cap lora_attn_bwd_tc '^svla_attn_bwd_tc_kernel$' 0 python bench.py --workload lora_step --quick --steps 1 --warmup 1
cap lora_gemm_tn '^svla_gemm_tn_kernel$' 0 python bench.py --workload lora_step --quick --steps 1 --warmup 1
cap lora_rmsnorm_bwd '^svla_rmsnorm_bwd_kernel$' 1 python bench.py --workload lora_step --quick --steps 1 --warmup 1
cap lora_geglu_bwd '^svla_geglu_bwd_kernel$' 0 python bench.py --workload lora_step --quick --steps 1 --warmup 1
ls -la $OUT/*.ncu-rep | awk '{print $5, $9}'
