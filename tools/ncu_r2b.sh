#!/bin/bash
# Round-2 (second half) ncu evidence, run on the GPU box after the plain commands have exited 0 without ncu:
#   launch list (gpu__time_duration.sum, --clock-control none) of one predict_action step,
#   `--set full` captures of the dominant GEMM (roofline traffic), the three tcgen05 attention shapes and the skinny GEMM.
OUT=gpurun_out
mkdir -p $OUT
python bench.py --quick --steps 1 --warmup 1 > $OUT/r2b_quick.log 2>&1 || exit 1
SVLA_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svla_ -c 6000 --csv --log-file $OUT/launches_r2b_predict.csv \
  python bench.py --quick --steps 1 --warmup 1 > $OUT/ncu_r2b_launch_predict.log 2>&1
echo "predict launch list rc=$?"
cap() {  # name regex skip command...
  local n=$1 rx=$2 sk=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $sk -c 1 -f -o $OUT/ncu_r2b_$n "$@" > $OUT/ncu_r2b_$n.log 2>&1
  echo "$n rc=$? $(ls $OUT/ncu_r2b_$n.ncu-rep 2>/dev/null)"
}
export SVLA_NO_GRAPHS=1
cap gemm_gateup '^svla_gemm_tcgen05_kernel$' 289 python bench.py --quick --steps 1 --warmup 1
cap attn_siglip '^svla_flash_attn_tc_kernel$' 2 python bench.py --quick --steps 1 --warmup 1
cap attn_beit '^svla_flash_attn_tc_kernel$' 29 python bench.py --quick --steps 1 --warmup 1
cap attn_gemma '^svla_flash_attn_tc_kernel$' 57 python bench.py --quick --steps 1 --warmup 1
cap gemm_beit_fc1 '^svla_gemm_tcgen05_kernel$' 152 python bench.py --quick --steps 1 --warmup 1
cap conv_rowtile_n32 '^svla_conv3x3_rowtile_kernel$' 1 python bench.py --quick --steps 1 --warmup 1
cap skinny_gateup '^svla_gemm_skinny_kernel$' 3 python bench.py --quick --steps 1 --warmup 1
ls -la $OUT/ncu_r2b_*.ncu-rep | awk '{print $5, $9}'
