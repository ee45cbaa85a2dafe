#!/bin/bash
# One `ncu --set full` capture per hot kernel of one eager predict_action step (B=64, 4B-224).  Run on the GPU box:
#   bash tools/ncu_capture.sh <tag> [name ...]      -> gpurun_out/ncu_<tag>_<name>.ncu-rep   (no names = all)
# ncu matches the BASE function name (no template arguments); the launch-skip numbers index the launches of that base
# name inside the first (warm-up) step, read off profiles/launches_*.csv.
TAG=${1:-r1}; shift
OUT=gpurun_out
mkdir -p $OUT
declare -A RX SK
RX[gemm_gateup]='svla_gemm_tcgen05_kernel';           SK[gemm_gateup]=288     # M=17792 N=18432 K=2304 +GeGLU (<256,2>)
RX[gemm_beit_fc1]='svla_gemm_tcgen05_kernel';         SK[gemm_beit_fc1]=112   # M=36928 N=4096 K=1024 +bias+GELU (TMA-store epilogue)
RX[attn_beit]='svla_flash_attn_tc_kernel';            SK[attn_beit]=2         # d=64, S=577, rel-pos bias
RX[attn_gemma]='svla_flash_attn_tc_kernel';           SK[attn_gemma]=26       # d=256, S=278, soft-cap
RX[attn_siglip]='svla_flash_attn_kernel';             SK[attn_siglip]=2       # d=72 (mma.sync path)
RX[skinny_gateup]='svla_gemm_skinny_kernel';          SK[skinny_gateup]=3
RX[decode_attn]='svla_decode_attn_fused_kernel';      SK[decode_attn]=30
RX[rmsnorm_prefill]='svla_rmsnorm_residual_warp_kernel'; SK[rmsnorm_prefill]=3
RX[rope_prefill]='svla_rope_kv_vec_kernel';           SK[rope_prefill]=2
RX[depth_tail]='svla_zoe_depth_tail_fused_kernel';    SK[depth_tail]=0
RX[conv_rowtile]='svla_conv3x3_rowtile_kernel';       SK[conv_rowtile]=1      # relative-head conv2 128->32 @384^2 (the 2nd row-tile launch)
RX[bilinear]='svla_bilinear_nhwc_kernel';             SK[bilinear]=4
RX[layernorm]='svla_layernorm_warp_kernel';           SK[layernorm]=10
NAMES="$@"
[ -z "$NAMES" ] && NAMES="${!RX[@]}"
for n in $NAMES; do
  SVLA_NO_GRAPHS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^${RX[$n]}\$" -s ${SK[$n]} -c 1 -f \
    -o $OUT/ncu_${TAG}_$n python bench.py --quick --steps 1 --warmup 1 > $OUT/ncu_${TAG}_$n.log 2>&1
  echo "$n rc=$? $(ls $OUT/ncu_${TAG}_$n.ncu-rep 2>/dev/null)"
done
