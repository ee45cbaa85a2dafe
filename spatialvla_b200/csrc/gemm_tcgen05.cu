// G1: bf16 GEMM  D = epilogue(A[M,K] * W[N,K]^T)  on Blackwell 5th-gen tensor cores.
//
//   * tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16, fp32 accumulators in TMEM (double buffered:
//     2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * A and W tiles (128x64 / BNx64 bf16, K-major) staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through an
//     mbarrier ring of kStages stages; TMA zero-fills every M/N/K tail, so ragged shapes need no host padding;
//   * warp-specialised persistent CTA (192 threads): warp0 = TMA producer, warp1 = TMEM allocator + MMA issuer,
//     warps2-5 = epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> 128-bit global stores);
//   * implicit 3x3 convolution mode: the A tile of a tap is a 4-D TMA box {64 ch, 16 w, 8 h, 1 n} of the NHWC
//     activation shifted by the tap offset -- the TMA out-of-bounds zero fill *is* the conv padding, there is
//     no im2col buffer.  The K loop runs over 9 taps x (C/64) channel chunks.
//
// Reference ops replaced: see include/spatialvla_b200.h (svla_gemm).
#include <cuda.h>
#include <cstdlib>
#include <cudaTypedefs.h>
#include "tc_ptx.cuh"

namespace {
using namespace svla_ptx;

constexpr int kBM = 128;
constexpr int kBK = 64;             // 64 bf16 = 128 bytes = one SWIZZLE_128B atom row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;          // two epilogue warps per TMEM lane group, each owns half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kConvTileW = 16, kConvTileH = 8;

// CG = CTAs per MMA (tcgen05 cta_group): 1 = one CTA computes a 128 x BN tile; 2 = a CTA pair (cluster of 2 along M)
// computes 256 x BN with ONE tcgen05.mma.cta_group::2 per K step: each CTA stages its own 128 A rows and only
// BN/2 of the W rows, halving the L2 -> shared-memory traffic per FLOP (the 1-CTA kernel is L2-bandwidth bound).
template <int BN, int CG = 1> struct Cfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kStageBytes >= 49152) ? 4 : ((kStageBytes >= 32768) ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  // per-epilogue-warp staging tile: 32 rows x 128 bytes.  Legacy epilogue: fp32 transpose tile (32 rows x 16 cols, stride 20);
  // TMA-store epilogue: one SWIZZLE_128B box (32 rows x 64 bf16 or 32 rows x 32 fp32), so the tile bases are 1024-aligned.
  static constexpr int kStagingPerWarp = 4096;
  static constexpr int kStagingBytes = kEpiWarps * kStagingPerWarp;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kStagingBytes;
  static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB) exceeded");
};

struct EpiParams {
  const float* bias;
  const float* colscale;
  const __nv_bfloat16* res_bf16;
  const __nv_bfloat16* res2_bf16;
  const float* res_f32;
  long long res_mod;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  __nv_bfloat16* out_relu;
  long long m, n, ldo;
  float alpha, act_param;
  int act, flags;
  int tma_out;    // 0 = register/LDG/STG epilogue; 1 = bf16 tile via TMA store; 2 = fp32 tile via TMA store; 3 = fp32 TMA reduce-add
  // conv geometry
  int nb, h, wd, tiles_h, tiles_w;
  int rowtile;    // 1: conv m-tile = 128 consecutive pixels of ONE image row (svla_conv3x3_rowtile_kernel), tiles_w = ceil(wd/128)
};

// ------------------------------------------------------------------------------------------ fused epilogue
__device__ __forceinline__ float apply_act(float v, int act, float p) {
  switch (act) {
    case SVLA_ACT_GELU_TANH: return gelu_tanh_fast(v);
    case SVLA_ACT_GELU_ERF: return gelu_erf_f(v);
    case SVLA_ACT_RELU: return fmaxf(v, 0.f);
    case SVLA_ACT_SOFTCAP: return p * tanhf(v / p);
    case SVLA_ACT_SOFTPLUS: return softplus_f(v);
    default: return v;
  }
}

// Epilogue.  Each of the 8 epilogue warps owns 32 accumulator rows (its TMEM lane group) x half of the tile's columns.
// Per 32-column TMEM load:
//   phase 1 (row owner, lane = row): value = act(alpha*acc + bias) * colscale (or the GeGLU pair product).  bias and
//     colscale are read once per chunk (lane j loads column j) and broadcast with shuffles; the activation switch is
//     hoisted out of the element loop (a per-element switch inlined every libm body 32x -> I-cache-miss bound).
//   staging: 16 output columns at a time through a per-warp fp32 tile in shared memory (row stride 20 floats).
//   phase 2 (coalesced): 4 lanes cover one row's 16 columns, 8 rows per pass, 4 passes; all residual / accumulate
//     loads of the 4 passes are issued before the first use, then the stores follow.
// Row offsets / validity of the 4 passes do not depend on the column, so they are computed ONCE per tile (RowSet):
// the first versions recomputed 64-bit row math per pass per chunk and ran at ~800 instructions per chunk.
constexpr int kStageLd = 20;
constexpr int kPasses = 4;

struct RowSet {
  unsigned off[kPasses];      // grow * ldo
  unsigned off32[kPasses];    // (grow % res_mod) * ldo for the broadcast fp32 residual
  bool ok[kPasses];
};

__device__ __forceinline__ void epilogue_store16(const EpiParams& ep, const float* stage, int lane, const RowSet& rs,
                                                 long long col0, long long ncols) {
  const bool accum = (ep.flags & SVLA_GEMM_ACCUM_F32) != 0;
  const int cg = lane & 3;
  const long long col = col0 + 4 * cg;
  if (col >= ncols) return;
  const bool vec_ok = ((ep.ldo & 3) == 0) && (col + 4 <= ncols);
  float4 x[kPasses];
#pragma unroll
  for (int p = 0; p < kPasses; ++p)
    x[p] = *reinterpret_cast<const float4*>(stage + (p * 8 + (lane >> 2)) * kStageLd + 4 * cg);
  if (vec_ok) {
    if (ep.res_bf16) {
      uint2 t[kPasses];
#pragma unroll
      for (int p = 0; p < kPasses; ++p) t[p] = rs.ok[p] ? __ldg(reinterpret_cast<const uint2*>(ep.res_bf16 + rs.off[p] + col)) : make_uint2(0, 0);
#pragma unroll
      for (int p = 0; p < kPasses; ++p) {
        x[p].x += bf16_bits_to_float(t[p].x & 0xFFFFu); x[p].y += bf16_bits_to_float(t[p].x >> 16);
        x[p].z += bf16_bits_to_float(t[p].y & 0xFFFFu); x[p].w += bf16_bits_to_float(t[p].y >> 16);
      }
    }
    if (ep.res2_bf16) {
      uint2 t[kPasses];
#pragma unroll
      for (int p = 0; p < kPasses; ++p) t[p] = rs.ok[p] ? __ldg(reinterpret_cast<const uint2*>(ep.res2_bf16 + rs.off[p] + col)) : make_uint2(0, 0);
#pragma unroll
      for (int p = 0; p < kPasses; ++p) {
        x[p].x += bf16_bits_to_float(t[p].x & 0xFFFFu); x[p].y += bf16_bits_to_float(t[p].x >> 16);
        x[p].z += bf16_bits_to_float(t[p].y & 0xFFFFu); x[p].w += bf16_bits_to_float(t[p].y >> 16);
      }
    }
    if (ep.res_f32) {
      float4 t[kPasses];
#pragma unroll
      for (int p = 0; p < kPasses; ++p) t[p] = rs.ok[p] ? __ldg(reinterpret_cast<const float4*>(ep.res_f32 + rs.off32[p] + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kPasses; ++p) { x[p].x += t[p].x; x[p].y += t[p].y; x[p].z += t[p].z; x[p].w += t[p].w; }
    }
    if (ep.out_f32 && accum) {
      float4 t[kPasses];
#pragma unroll
      for (int p = 0; p < kPasses; ++p) t[p] = rs.ok[p] ? *reinterpret_cast<const float4*>(ep.out_f32 + rs.off[p] + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < kPasses; ++p) { x[p].x += t[p].x; x[p].y += t[p].y; x[p].z += t[p].z; x[p].w += t[p].w; }
    }
#pragma unroll
    for (int p = 0; p < kPasses; ++p) {
      if (!rs.ok[p]) continue;
      if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + rs.off[p] + col) = x[p];
      if (ep.out_bf16)
        *reinterpret_cast<uint2*>(ep.out_bf16 + rs.off[p] + col) = make_uint2(pack_bf16x2(x[p].x, x[p].y), pack_bf16x2(x[p].z, x[p].w));
      if (ep.out_relu)
        *reinterpret_cast<uint2*>(ep.out_relu + rs.off[p] + col) =
            make_uint2(pack_bf16x2(fmaxf(x[p].x, 0.f), fmaxf(x[p].y, 0.f)), pack_bf16x2(fmaxf(x[p].z, 0.f), fmaxf(x[p].w, 0.f)));
    }
  } else {
#pragma unroll
    for (int p = 0; p < kPasses; ++p) {
      if (!rs.ok[p]) continue;
      const float xv[4] = {x[p].x, x[p].y, x[p].z, x[p].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (col + e >= ncols) break;
        float y = xv[e];
        if (ep.res_bf16) y += __bfloat162float(ep.res_bf16[rs.off[p] + col + e]);
        if (ep.res2_bf16) y += __bfloat162float(ep.res2_bf16[rs.off[p] + col + e]);
        if (ep.res_f32) y += ep.res_f32[rs.off32[p] + col + e];
        if (ep.out_f32) {
          if (accum) y += ep.out_f32[rs.off[p] + col + e];
          ep.out_f32[rs.off[p] + col + e] = y;
        }
        if (ep.out_bf16) ep.out_bf16[rs.off[p] + col + e] = __float2bfloat16(y);
        if (ep.out_relu) ep.out_relu[rs.off[p] + col + e] = __float2bfloat16(fmaxf(y, 0.f));
      }
    }
  }
}

__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, const float (&acc)[32], float* stage, int lane,
                                               long long n0, const RowSet& rs) {
  const bool geglu = (ep.flags & SVLA_GEMM_GEGLU) != 0;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = acc[j] * ep.alpha;
  if (ep.bias) {
    const float b = (n0 + lane < ep.n) ? __ldg(ep.bias + n0 + lane) : 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, b, j);
  }
  if (!geglu) {
    switch (ep.act) {
      case SVLA_ACT_GELU_TANH:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
        break;
      case SVLA_ACT_GELU_ERF:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        break;
      case SVLA_ACT_RELU:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        break;
      case SVLA_ACT_SOFTCAP: {
        const float inv = 1.f / ep.act_param;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ep.act_param * tanhf(v[j] * inv);
        break;
      }
      case SVLA_ACT_SOFTPLUS:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = softplus_f(v[j]);
        break;
      default: break;
    }
  }
  if (ep.colscale) {
    const float c = (n0 + lane < ep.n) ? __ldg(ep.colscale + n0 + lane) : 1.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= __shfl_sync(0xffffffffu, c, j);
  }
  float4* srow = reinterpret_cast<float4*>(stage + lane * kStageLd);
  if (geglu) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      srow[j] = make_float4(gelu_tanh_fast(v[8 * j]) * v[8 * j + 1], gelu_tanh_fast(v[8 * j + 2]) * v[8 * j + 3],
                            gelu_tanh_fast(v[8 * j + 4]) * v[8 * j + 5], gelu_tanh_fast(v[8 * j + 6]) * v[8 * j + 7]);
    __syncwarp();
    epilogue_store16(ep, stage, lane, rs, n0 >> 1, ep.n >> 1);
    __syncwarp();
  } else {
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        srow[j] = make_float4(v[16 * hlf + 4 * j], v[16 * hlf + 4 * j + 1], v[16 * hlf + 4 * j + 2], v[16 * hlf + 4 * j + 3]);
      __syncwarp();
      epilogue_store16(ep, stage, lane, rs, n0 + 16 * hlf, ep.n);
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------ TMA-store epilogue
// Used when the GEMM has ONE output tensor and no residual operands (ep.tma_out != 0): the K ~ 1024 projections of SigLIP /
// BEiT are epilogue-bound with the register/LDG/STG path above (ncu: ~25 warp-instructions per element).  Here lane = row:
// value = act(alpha*acc + bias) * colscale for 32 consecutive columns, bias / colscale fetched as warp-uniform float4 loads
// (one L1 broadcast each, no shuffles), the row segment is written ONCE to a SWIZZLE_128B shared-memory box (conflict-free
// 16-byte stores) and a single elected lane hands the 32-row box to the TMA unit: cp.async.bulk.tensor store, or
// cp.reduce.async.bulk.tensor .add for `out_f32 +=` (the read-modify-write happens in L2, the SM never loads the old value).
// Rows >= M and columns >= N are clipped by the tensor map, so ragged tiles need no predicates.
// packed (f32x2) erf / tanh GELU of the TMA-store epilogue: gelu_erf_pair / gelu_tanh_pair in tc_ptx.cuh
__device__ __forceinline__ void epilogue_math32(const EpiParams& ep, float (&v)[32], long long n0) {
  if (ep.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
  }
  const bool full = n0 + 32 <= ep.n;                     // warp-uniform
  if (ep.bias) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        unpack_f32x2(add_f32x2(pack_f32x2(v[4 * j], v[4 * j + 1]), pack_f32x2(b.x, b.y)), v[4 * j], v[4 * j + 1]);
        unpack_f32x2(add_f32x2(pack_f32x2(v[4 * j + 2], v[4 * j + 3]), pack_f32x2(b.z, b.w)), v[4 * j + 2], v[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += (n0 + j < ep.n) ? __ldg(ep.bias + n0 + j) : 0.f;
    }
  }
  switch (ep.act) {
    case SVLA_ACT_GELU_TANH:
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu_tanh_pair(v[j], v[j + 1]);
      break;
    case SVLA_ACT_GELU_ERF:
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu_erf_pair(v[j], v[j + 1]);
      break;
    case SVLA_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case SVLA_ACT_SOFTCAP: {
      const float inv = 1.f / ep.act_param;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ep.act_param * tanhf(v[j] * inv);
      break;
    }
    case SVLA_ACT_SOFTPLUS:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = softplus_f(v[j]);
      break;
    default: break;
  }
  if (ep.colscale) {
    if (full) {
      const float4* c4 = reinterpret_cast<const float4*>(ep.colscale + n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 c = __ldg(c4 + j);
        v[4 * j] *= c.x; v[4 * j + 1] *= c.y; v[4 * j + 2] *= c.z; v[4 * j + 3] *= c.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= (n0 + j < ep.n) ? __ldg(ep.colscale + n0 + j) : 1.f;
    }
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// One warp drains its 32 accumulator rows x COLS columns [c_begin, c_begin + COLS) of the tile.  `tile` = this warp's 4 KB,
// 1024-byte aligned staging box; row0 = global output row of lane 0.
template <int COLS>
__device__ __forceinline__ void epilogue_tma(const EpiParams& ep, const CUtensorMap* tm_o, uint32_t taddr, uint8_t* tile, int lane,
                                             int c_begin, long long n_base, int row0) {
  const uint32_t row_s = smem_u32(tile) + static_cast<uint32_t>(lane) * 128u;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  // software pipeline: the TMEM load of chunk c + 1 is in flight during the arithmetic / staging of chunk c
  uint32_t r[32];
  if (n_base + c_begin < ep.n) tmem_ld32_issue(taddr + c_begin, r);
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_begin + COLS; c0 += 32) {
    const long long n0 = n_base + c0;
    if (n0 >= ep.n) break;                               // warp-uniform
    tmem_ld32_wait(r);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (c0 + 32 < c_begin + COLS && n0 + 32 < ep.n) tmem_ld32_issue(taddr + c0 + 32, r);
    epilogue_math32(ep, v, n0);
    if (ep.tma_out == 1) {
      // bf16: two 32-column chunks fill one 128-byte box row (64 columns)
      const uint32_t half = static_cast<uint32_t>((c0 - c_begin) >> 5) & 1u;
      if (half == 0) {
        if (lane == 0) bulk_wait_read_all();             // the previous box has left shared memory
        __syncwarp();
      }
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j)
        st_shared_v4(row_s + (((half * 4u + j) ^ sw) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      if (half == 1 || n0 + 32 >= ep.n) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tm_o, tile, static_cast<int>(n0) - 32 * static_cast<int>(half), row0);
          bulk_commit_group();
        }
      }
    } else {
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j)
        st_shared_v4(row_s + ((j ^ sw) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                     __float_as_uint(v[4 * j + 3]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (ep.tma_out == 3) tma_reduce_add_2d(tm_o, tile, static_cast<int>(n0), row0);
        else tma_store_2d(tm_o, tile, static_cast<int>(n0), row0);
        bulk_commit_group();
      }
    }
  }
}

// Row-tile conv variant of the bf16 TMA-store epilogue: the 32 accumulator rows of a warp are 32 consecutive pixels of ONE image
// row, stored through a 4-D map {channel, x, y, image} (pixels past the row end are clipped by the hardware).  The register /
// staging epilogue spent ~4000 clocks per 128 x 32 tile on row bookkeeping (64-bit divisions per pass) and strided 8-byte stores.
template <int COLS>
__device__ __forceinline__ void epilogue_tma_conv(const EpiParams& ep, const CUtensorMap* tm_o, uint32_t taddr, uint8_t* tile, int lane,
                                                  int c_begin, long long n_base, int x0, int y, int img) {
  const uint32_t row_s = smem_u32(tile) + static_cast<uint32_t>(lane) * 128u;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_begin + COLS; c0 += 32) {
    const long long n0 = n_base + c0;
    if (n0 >= ep.n) break;                               // warp-uniform
    uint32_t r[32];
    tmem_ld32(taddr + c0, r);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    epilogue_math32(ep, v, n0);
    const uint32_t half = static_cast<uint32_t>((c0 - c_begin) >> 5) & 1u;
    if (half == 0) {
      if (lane == 0) bulk_wait_read_all();               // the previous box has left shared memory
      __syncwarp();
    }
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j)
      st_shared_v4(row_s + (((half * 4u + j) ^ sw) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                   pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
    if (half == 1 || n0 + 32 >= ep.n) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(tm_o, tile, static_cast<int>(n0) - 32 * static_cast<int>(half), x0, y, img);
        bulk_commit_group();
      }
    }
  }
}

// Global output row of tile row r (and validity) in linear / conv mode.
__device__ __forceinline__ bool tile_row_to_global(const EpiParams& ep, bool conv, long long m_tile, int r, long long& grow) {
  if (!conv) {
    grow = m_tile * kBM + r;
    return grow < ep.m;
  }
  const int tiles_per_img = ep.tiles_h * ep.tiles_w;
  const int img = static_cast<int>(m_tile / tiles_per_img);
  const int rem = static_cast<int>(m_tile % tiles_per_img);
  if (ep.rowtile) {
    const int hh = rem / ep.tiles_w, ww = (rem % ep.tiles_w) * kBM + r;
    grow = (static_cast<long long>(img) * ep.h + hh) * ep.wd + ww;
    return ww < ep.wd && img < ep.nb;
  }
  const int hh = (rem / ep.tiles_w) * kConvTileH + r / kConvTileW;
  const int ww = (rem % ep.tiles_w) * kConvTileW + r % kConvTileW;
  grow = (static_cast<long long>(img) * ep.h + hh) * ep.wd + ww;
  return hh < ep.h && ww < ep.wd && img < ep.nb;
}

// Tile rasterisation: m-groups (1 or 2 m-tiles) are walked in super-rows of kRasterGroup groups; inside a super-row
// the n index is the slow one.  One wave of CTAs then touches ~16 A tiles x ~9 W tiles instead of every A tile of
// the problem, which keeps both operands L2-resident (ncu: 10x DRAM re-reads of A with the plain m-fastest order).
constexpr int kRasterGroup = 16;
__device__ __forceinline__ void raster_tile(long long tile, long long num_m_groups, long long num_n_tiles, long long& mg,
                                            long long& nt) {
  const long long per_super = static_cast<long long>(kRasterGroup) * num_n_tiles;
  const long long sr = tile / per_super;
  const long long rem = tile - sr * per_super;
  const long long m0 = sr * kRasterGroup;
  const long long rows = (num_m_groups - m0) < kRasterGroup ? (num_m_groups - m0) : kRasterGroup;
  nt = rem / rows;
  mg = m0 + (rem - nt * rows);
}

// ------------------------------------------------------------------------------------------ the kernel
template <int BN, int CG>
__global__ void __launch_bounds__(kThreads, 1)
svla_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_a2,
                         const __grid_constant__ CUtensorMap tm_b2, const EpiParams ep, const long long num_m_tiles, const long long num_n_tiles,
                         const int num_k_blocks, const int conv, const int c_chunks, const int kb_main) {
  // K extension (LoRA): k-blocks [kb_main, num_k_blocks) are loaded from a SECOND operand pair A2[M, K2] / W2[N, K2], so that
  // D = A W^T + A2 W2^T accumulates in one TMEM tile: y = x W^T + (s x A^T) B^T of an adapted Linear costs K2 / 64 extra k-blocks
  // of the base GEMM instead of a second pass over y (train/spatialvla_finetune.py:262-302; peft lora Linear.forward).
  using C = Cfg<BN, CG>;
  constexpr int BNL = BN / CG;       // W rows staged by this CTA
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned stage bases (identical offsets in both CTAs of a pair)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint8_t* staging = smem + C::kStages * C::kStageBytes;                // 1024-aligned (stage sizes are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + C::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (ep.tma_out) tma_prefetch_desc(&tm_o);
    if (kb_main < num_k_blocks) { tma_prefetch_desc(&tm_a2); tma_prefetch_desc(&tm_b2); }
#pragma unroll 1
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], kEpiWarps * CG);     // every epilogue warp of every CTA of the pair arrives on the leader's barrier
    mbar_init(&tmem_empty[1], kEpiWarps * CG);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::kTmemCols, CG>(tmem_ptr_smem);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile space: pairs of consecutive m-tiles (CG == 2) x n-tiles, m fastest so concurrent CTAs share the W tile in L2
  const long long num_m_groups = (num_m_tiles + CG - 1) / CG;
  const long long num_tiles = num_m_groups * num_n_tiles;
  const long long first = blockIdx.x / CG, step = gridDim.x / CG;

  if (warp == 0) {
    // ===================================================== TMA producer (one lane, every CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = first; tile < num_tiles; tile += step) {
        long long mg, n_tile;
        raster_tile(tile, num_m_groups, num_n_tiles, mg, n_tile);
        const long long m_tile = mg * CG + cta_rank;
        int img = 0, h0 = 0, w0 = 0;
        if (conv) {
          const int tiles_per_img = ep.tiles_h * ep.tiles_w;
          img = static_cast<int>(m_tile / tiles_per_img);
          const int rem = static_cast<int>(m_tile % tiles_per_img);
          h0 = (rem / ep.tiles_w) * kConvTileH;
          w0 = (rem % ep.tiles_w) * kConvTileW;
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem_a + stage * C::kABytes;
          uint8_t* sb = smem_b + stage * C::kBBytes;
          const bool ext = kb >= kb_main;                  // K-extension block: second operand pair
          const int kc = (ext ? kb - kb_main : kb) * kBK;
          if constexpr (CG == 1) {
            mbar_expect_tx(&full_bar[stage], C::kStageBytes);
            if (conv) {
              const int tap = kb / c_chunks, cc = kb - tap * c_chunks;
              tma_load_4d(sa, &tm_a, &full_bar[stage], cc * kBK, w0 + (tap % 3) - 1, h0 + (tap / 3) - 1, img);
            } else {
              tma_load_2d(sa, ext ? &tm_a2 : &tm_a, &full_bar[stage], kc, static_cast<int>(m_tile * kBM));
            }
            tma_load_2d(sb, ext ? &tm_b2 : &tm_b, &full_bar[stage], kc, static_cast<int>(n_tile * BN));
          } else {
            // the leader arms its barrier for the bytes of BOTH CTAs; the peer's TMA signals the leader's barrier
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
            if (conv) {
              const int tap = kb / c_chunks, cc = kb - tap * c_chunks;
              tma_load_4d_cg2(sa, &tm_a, &full_bar[stage], cc * kBK, w0 + (tap % 3) - 1, h0 + (tap / 3) - 1, img);
            } else {
              tma_load_2d_cg2(sa, ext ? &tm_a2 : &tm_a, &full_bar[stage], kc, static_cast<int>(m_tile * kBM));
            }
            tma_load_2d_cg2(sb, ext ? &tm_b2 : &tm_b, &full_bar[stage], kc, static_cast<int>(n_tile * BN + cta_rank * BNL));
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (the leader CTA's warp 1, converged; an elected lane issues)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM * CG, BN);
      const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(smem_a)), b_desc0 = make_kmajor_sw128_desc(smem_u32(smem_b));
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (long long tile = first; tile < num_tiles; tile += step, ++it) {
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        mbar_wait(&tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            // descriptors = constants + (stage offset >> 4) in the start-address field; +2 per 16 bf16 (32 bytes) along K
            const uint64_t da = a_desc0 + static_cast<uint64_t>((stage * C::kABytes) >> 4);
            const uint64_t db = b_desc0 + static_cast<uint64_t>((stage * C::kBBytes) >> 4);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              if constexpr (CG == 1)
                umma_bf16(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                          static_cast<uint32_t>((kb | k) != 0));
              else
                umma_bf16_cg2(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                              static_cast<uint32_t>((kb | k) != 0));
            }
            // frees the smem stage (in both CTAs for CG == 2) once these MMAs have read it
            if constexpr (CG == 1) umma_commit(&empty_bar[stage]); else umma_commit_cg2(&empty_bar[stage]);
            // accumulator complete -> epilogue warps (of both CTAs)
            if (kb == num_k_blocks - 1) {
              if constexpr (CG == 1) umma_commit(&tmem_full[as]); else umma_commit_cg2(&tmem_full[as]);
            }
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================================================== epilogue warps 2..9 (TMEM lane group = warp % 4)
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;                    // which half of the tile's columns this warp drains
    constexpr int kColsPerWarp = BN >= 64 ? BN / 2 : BN;
    uint32_t it = 0;
    uint8_t* tile_s = staging + (warp - 2) * C::kStagingPerWarp;
    float* stage = reinterpret_cast<float*>(tile_s);
    for (long long tile = first; tile < num_tiles; tile += step, ++it) {
      long long mg, n_tile;
      raster_tile(tile, num_m_groups, num_n_tiles, mg, n_tile);
      const long long m_tile = mg * CG + cta_rank;
      const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
      if constexpr (BN >= 128) {
        if (ep.tma_out) {                                  // warp-uniform, fixed for the launch
          mbar_wait(&tmem_full[as], aphase);
          tc_fence_after();
          epilogue_tma<BN / 2>(ep, &tm_o, taddr, tile_s, lane, chalf * (BN / 2), n_tile * BN, static_cast<int>(m_tile * kBM) + q * 32);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 1) mbar_arrive(&tmem_empty[as]); else mbar_arrive_remote(&tmem_empty[as], 0);
          }
          continue;
        }
      }
      RowSet rs;
#pragma unroll
      for (int p = 0; p < kPasses; ++p) {
        long long grow = 0;
        rs.ok[p] = m_tile < num_m_tiles && tile_row_to_global(ep, conv != 0, m_tile, q * 32 + p * 8 + (lane >> 2), grow);
        rs.off[p] = static_cast<unsigned>(grow * ep.ldo);
        rs.off32[p] = static_cast<unsigned>((ep.res_mod > 0 ? grow % ep.res_mod : grow) * ep.ldo);
      }
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (BN >= 64 || chalf == 0) {
#pragma unroll 1
        for (int c0 = chalf * kColsPerWarp; c0 < (chalf + 1) * kColsPerWarp; c0 += 32) {
          const long long n0 = n_tile * BN + c0;
          if (n0 >= ep.n) break;                 // warp-uniform
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
          epilogue_chunk(ep, acc, stage, lane, n0, rs);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tmem_empty[as]); else mbar_arrive_remote(&tmem_empty[as], 0);
      }
    }
    if (ep.tma_out && lane == 0) bulk_wait_read_all();     // the staging tiles must outlive the last TMA store's reads
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols, CG>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ 3x3 conv, row-tile variant
// The patch-tile conv above re-fetches the activation tile from L2 for each of the 9 taps; with few output channels (the
// relative head: 128 -> 32 at 384 x 384) the MMA needs more L2 -> shared-memory bytes per cycle than an SM can pull and the
// kernel sits at 14 % of the tensor peak.  Here an m-tile is 128 consecutive pixels of ONE image row, so the three horizontal
// taps of a kernel row read the SAME shared-memory buffer at row offsets 0 / 1 / 2: the A operand of tap dx is the K-major
// SWIZZLE_128B descriptor of the 130-pixel buffer advanced by dx * 128 bytes (measured on B200: tcgen05 derives the swizzle
// phase from the absolute shared-memory address, so a start that is not 1024-byte aligned is legal with base_offset = 0).
// One pipeline stage = (64-channel chunk, kernel row dy): a {64 ch, 130 w, 1 h} TMA box (zero-filled halo = padding) plus
// the three [BN x 64] weight tiles of that kernel row; 12 MMAs per stage.  Activation traffic: 3 instead of 9 loads per tile.
template <int BN> struct RowCfg {
  static constexpr int kARows = kBM + 2;
  static constexpr int kABytesRaw = kARows * kBK * 2;                       // 16 640 B transferred
  static constexpr int kABytes = (kABytesRaw + 1023) / 1024 * 1024;         // 17 408 B reserved (tile bases stay 1024-aligned)
  static constexpr int kBTile = BN * kBK * 2;
  static constexpr int kBBytes = 3 * kBTile;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kEpiWarps * 4096;
  static constexpr int kStages = (232448 - 1024 - 256 - kStagingBytes) / kStageBytes > 6 ? 6 : (232448 - 1024 - 256 - kStagingBytes) / kStageBytes;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kStagingBytes;
  static_assert(kStages >= 2 && kBTile % 1024 == 0, "row-tile conv: pipeline needs two stages and 1 KB aligned weight tiles");
};

// w_res != 0 (few output channels: all 9 x c_chunks weight tiles fit in shared memory, one n-tile): the weights are loaded ONCE
// per CTA and a pipeline stage is only the 130-pixel activation row (1 TMA instead of 4 per stage: with 12 sixteen-clock MMAs per
// stage the producer's issue rate bounded the 128 -> 32 conv at 384 x 384).  n_stages = ring depth chosen by the host.
constexpr int kRowMaxStages = 8;
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
svla_conv3x3_rowtile_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                            const __grid_constant__ CUtensorMap tm_o, const EpiParams ep,
                            const long long num_m_tiles, const long long num_n_tiles, const int c_chunks, const int cpad,
                            const int w_res, const int n_stages) {
  using C = RowCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem_al = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int w_res_bytes = w_res ? 9 * c_chunks * C::kBTile : 0;
  const int stage_bytes = w_res ? C::kABytes : C::kStageBytes;
  uint8_t* smem_w = smem_al;                                      // resident weight tiles [(dy * 3 + dx) * c_chunks + cc][BN x 64]
  uint8_t* smem = smem_al + w_res_bytes;                          // activation (+ weight) stages
  uint8_t* staging = smem + n_stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + C::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kRowMaxStages;
  uint64_t* tmem_full = bars + 2 * kRowMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_full = tmem_empty + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (ep.tma_out) tma_prefetch_desc(&tm_o);
#pragma unroll 1
    for (int s = 0; s < n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], kEpiWarps); mbar_init(&tmem_empty[1], kEpiWarps);
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::kTmemCols, 1>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const long long num_tiles = num_m_tiles * num_n_tiles;
  const int stages_per_tile = 3 * c_chunks;

  if (warp == 0) {
    if (lane == 0) {
      if (w_res) {
        mbar_expect_tx(w_full, w_res_bytes);
        for (int tap = 0; tap < 9; ++tap)
          for (int cc = 0; cc < c_chunks; ++cc)
            tma_load_2d(smem_w + (tap * c_chunks + cc) * C::kBTile, &tm_b, w_full, tap * cpad + cc * kBK, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long m_tile = tile / num_n_tiles, n_tile = tile - m_tile * num_n_tiles;
        const int tiles_per_img = ep.tiles_h * ep.tiles_w;
        const int img = static_cast<int>(m_tile / tiles_per_img), rem = static_cast<int>(m_tile % tiles_per_img);
        const int hrow = rem / ep.tiles_w, w0 = (rem % ep.tiles_w) * kBM;
        for (int it = 0; it < stages_per_tile; ++it) {
          const int cc = it / 3, dy = it - 3 * cc;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          if (w_res) {
            mbar_expect_tx(&full_bar[stage], C::kABytesRaw);
            tma_load_4d(sa, &tm_a, &full_bar[stage], cc * kBK, w0 - 1, hrow + dy - 1, img);
          } else {
            uint8_t* sb = sa + C::kABytes;
            mbar_expect_tx(&full_bar[stage], C::kABytesRaw + C::kBBytes);
            tma_load_4d(sa, &tm_a, &full_bar[stage], cc * kBK, w0 - 1, hrow + dy - 1, img);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              tma_load_2d(sb + dx * C::kBTile, &tm_b, &full_bar[stage], (dy * 3 + dx) * cpad + cc * kBK, static_cast<int>(n_tile * BN));
          }
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {                                                        // converged warp, an elected lane issues (see elect_one)
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN);
      const uint64_t desc0 = make_kmajor_sw128_desc(smem_u32(smem));
      const uint64_t wdesc0 = make_kmajor_sw128_desc(smem_u32(smem_w));
      if (w_res) mbar_wait(w_full, 0);
      int stage = 0;
      uint32_t phase = 0, it_tile = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it_tile) {
        const uint32_t as = it_tile & 1u, aphase = (it_tile >> 1) & 1u;
        mbar_wait(&tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int it = 0; it < stages_per_tile; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const int cc = it / 3, dy = it - 3 * cc;
            const uint64_t a_desc = desc0 + static_cast<uint64_t>((stage * stage_bytes) >> 4);
            const uint64_t b_desc = w_res ? wdesc0 + static_cast<uint64_t>(((dy * 3 * c_chunks + cc) * C::kBTile) >> 4)
                                          : a_desc + static_cast<uint64_t>(C::kABytes >> 4);
            const int b_step = w_res ? c_chunks * C::kBTile : C::kBTile;      // bytes between the tiles of taps dx and dx + 1
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              // tap dx: the same 130-pixel buffer, rows dx .. dx + 127 (start advanced by dx * 128 B, base_offset stays 0)
              const uint64_t da = a_desc + static_cast<uint64_t>((dx * 128) >> 4);
              const uint64_t db = b_desc + static_cast<uint64_t>((dx * b_step) >> 4);
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k)
                umma_bf16(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                          static_cast<uint32_t>((it | dx | k) != 0));
            }
            umma_commit(&empty_bar[stage]);
            if (it == stages_per_tile - 1) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int kColsPerWarp = BN >= 64 ? BN / 2 : BN;
    uint32_t it_tile = 0;
    float* stage_f = reinterpret_cast<float*>(staging + (warp - 2) * 4096);
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it_tile) {
      const long long m_tile = tile / num_n_tiles, n_tile = tile - m_tile * num_n_tiles;
      const uint32_t as = it_tile & 1u, aphase = (it_tile >> 1) & 1u;
      if (ep.tma_out) {                                    // warp-uniform, fixed for the launch
        const int tiles_per_img = ep.tiles_h * ep.tiles_w;
        const int mt32 = static_cast<int>(m_tile);
        const int img = mt32 / tiles_per_img, rem = mt32 - img * tiles_per_img;
        const int hrow = rem / ep.tiles_w, w0 = (rem - hrow * ep.tiles_w) * kBM;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        const uint32_t taddr_t = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
        if (BN >= 64 || chalf == 0)
          epilogue_tma_conv<kColsPerWarp>(ep, &tm_o, taddr_t, staging + (warp - 2) * 4096, lane, chalf * kColsPerWarp, n_tile * BN, w0 + q * 32, hrow, img);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
        continue;
      }
      RowSet rs;
#pragma unroll
      for (int p = 0; p < kPasses; ++p) {
        long long grow = 0;
        rs.ok[p] = tile_row_to_global(ep, true, m_tile, q * 32 + p * 8 + (lane >> 2), grow);
        rs.off[p] = static_cast<unsigned>(grow * ep.ldo);
        rs.off32[p] = static_cast<unsigned>((ep.res_mod > 0 ? grow % ep.res_mod : grow) * ep.ldo);
      }
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(q * 32) << 16);
      if (BN >= 64 || chalf == 0) {
#pragma unroll 1
        for (int c0 = chalf * kColsPerWarp; c0 < (chalf + 1) * kColsPerWarp; c0 += 32) {
          const long long n0 = n_tile * BN + c0;
          if (n0 >= ep.n) break;
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
          epilogue_chunk(ep, acc, stage_f, lane, n0, rs);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
    if (ep.tma_out && lane == 0) bulk_wait_read_all();     // the staging tiles must outlive the last TMA store's reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols, 1>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ SIMT debug kernel
// Straightforward one-thread-per-output kernel with the same epilogue; used by tests only (args->impl == 1) so
// that the rest of the path can be validated independently of the tcgen05 kernel.
__global__ void svla_gemm_simt_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ w,
                                      EpiParams ep, long long k, long long lda, long long ldw, int conv, int c, int cpad) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const bool geglu = (ep.flags & SVLA_GEMM_GEGLU) != 0;
  const long long ncols = geglu ? (ep.n >> 1) : ep.n;
  if (idx >= ep.m * ncols) return;
  const long long row = idx / ncols, col = idx % ncols;
  auto dot = [&](long long n) {
    float acc = 0.f;
    if (!conv) {
      for (long long kk = 0; kk < k; ++kk) acc += __bfloat162float(a[row * lda + kk]) * __bfloat162float(w[n * ldw + kk]);
    } else {
      const int ww = static_cast<int>(row % ep.wd), hh = static_cast<int>((row / ep.wd) % ep.h);
      const long long img = row / (static_cast<long long>(ep.wd) * ep.h);
      for (int tap = 0; tap < 9; ++tap) {
        const int y = hh + tap / 3 - 1, x = ww + tap % 3 - 1;
        if (y < 0 || y >= ep.h || x < 0 || x >= ep.wd) continue;
        const __nv_bfloat16* ap = a + ((img * ep.h + y) * ep.wd + x) * c;
        const __nv_bfloat16* wp = w + n * ldw + tap * cpad;
        for (int ci = 0; ci < c; ++ci) acc += __bfloat162float(ap[ci]) * __bfloat162float(wp[ci]);
      }
    }
    return acc;
  };
  auto pre = [&](float x, long long n) {
    x *= ep.alpha;
    if (ep.bias) x += ep.bias[n];
    if (!geglu) x = apply_act(x, ep.act, ep.act_param);
    if (ep.colscale) x *= ep.colscale[n];
    return x;
  };
  if (geglu) {
    const float g = pre(dot(2 * col), 2 * col), u = pre(dot(2 * col + 1), 2 * col + 1);
    ep.out_bf16[row * ep.ldo + col] = __float2bfloat16(gelu_tanh_f(g) * u);
    return;
  }
  float x = pre(dot(col), col);
  const long long off = row * ep.ldo + col;
  if (ep.res_bf16) x += __bfloat162float(ep.res_bf16[off]);
  if (ep.res2_bf16) x += __bfloat162float(ep.res2_bf16[off]);
  if (ep.res_f32) x += ep.res_f32[(ep.res_mod > 0 ? row % ep.res_mod : row) * ep.ldo + col];
  if (ep.out_f32) {
    if (ep.flags & SVLA_GEMM_ACCUM_F32) x += ep.out_f32[off];
    ep.out_f32[off] = x;
  }
  if (ep.out_bf16) ep.out_bf16[off] = __float2bfloat16(x);
  if (ep.out_relu) ep.out_relu[off] = __float2bfloat16(fmaxf(x, 0.f));
}

// ------------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int encode_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
              uint32_t box_outer) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

// output tile map for the TMA-store epilogue: [M rows, N cols] with row stride ldo, box = 32 rows x 128 bytes, SWIZZLE_128B
int encode_out(CUtensorMap* tm, void* base, bool f32, uint64_t n, uint64_t m, uint64_t ldo) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  const uint64_t es = f32 ? 4 : 2;
  cuuint64_t dims[2] = {n, m};
  cuuint64_t strides[1] = {ldo * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / es), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

int encode_nhwc(CUtensorMap* tm, const void* base, int nb, int h, int w, int c, int box_w = kConvTileW, int box_h = kConvTileH) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                        static_cast<cuuint64_t>(nb)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2,
                           static_cast<cuuint64_t>(h) * w * c * 2};
  cuuint32_t box[4] = {kBK, static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

template <int BN, int CG>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& ta2, const CUtensorMap& tb2,
              const EpiParams& ep, long long mt, long long nt, int kb, int conv, int c_chunks, int kb_main, cudaStream_t st) {
  using C = Cfg<BN, CG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(svla_gemm_tcgen05_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) {
      svla_set_error("svla_gemm: cudaFuncSetAttribute(%d bytes) failed: %s", C::kSmemBytes, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  const long long groups = ((mt + CG - 1) / CG) * nt;
  const int max_groups = svla_num_sms() / CG;
  const int grid = static_cast<int>(groups < max_groups ? groups : max_groups) * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, svla_gemm_tcgen05_kernel<BN, CG>, ta, tb, to, ta2, tb2, ep, mt, nt, kb, conv, c_chunks, kb_main);
  if (e != cudaSuccess) {
    svla_set_error("svla_gemm_tcgen05<%d,%d>: launch failed: %s", BN, CG, cudaGetErrorString(e));
    return -2;
  }
  SVLA_LAUNCH_CHECK("svla_gemm_tcgen05");
  return 0;
}

template <int BN>
int launch_rowtile(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const EpiParams& ep, long long mt, long long nt, int c_chunks,
                   int cpad, cudaStream_t st) {
  using C = RowCfg<BN>;
  constexpr int kBudget = 232448 - 1024 - 256 - C::kStagingBytes;
  // resident weights: one n-tile and 9 x c_chunks weight tiles that leave room for >= 4 activation stages
  const int w_bytes = 9 * c_chunks * C::kBTile;
  static const bool no_res = getenv("SVLA_CONV_NO_WRES") != nullptr;      // A/B switch
  const bool w_res = !no_res && nt == 1 && w_bytes + 4 * C::kABytes <= kBudget;
  int n_stages = w_res ? (kBudget - w_bytes) / C::kABytes : C::kStages;
  if (n_stages > kRowMaxStages) n_stages = kRowMaxStages;
  const int smem = (w_res ? w_bytes + n_stages * C::kABytes : n_stages * C::kStageBytes) + 1024 + 256 + C::kStagingBytes;
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(svla_conv3x3_rowtile_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      svla_set_error("svla_gemm(conv row-tile): cudaFuncSetAttribute(%d bytes) failed: %s", smem, cudaGetErrorString(e));
      return -2;
    }
    configured = smem;
  }
  const long long tiles = mt * nt;
  const int grid = static_cast<int>(tiles < svla_num_sms() ? tiles : svla_num_sms());
  svla_conv3x3_rowtile_kernel<BN><<<grid, kThreads, smem, st>>>(ta, tb, to, ep, mt, nt, c_chunks, cpad, w_res ? 1 : 0, n_stages);
  SVLA_LAUNCH_CHECK("svla_conv3x3_rowtile");
  return 0;
}

int pick_block_n(long long m_tiles, long long n) {
  const int sms = svla_num_sms();
  if (m_tiles == 1) {
    // Skinny (decode, M <= 128): weight-streaming, HBM-bound. Spread N over as many SMs as possible; the per-k-block
    // cost of a CTA is the 16 KB A tile (L2-hot) plus BN*128 B of weights.
    long long best_cost = -1;
    int best = 32;
    for (int bn : {32, 64, 128, 256}) {
      if (bn > 32 && n <= bn / 2) continue;
      const long long tiles = (n + bn - 1) / bn;
      const long long cost = ((tiles + sms - 1) / sms) * (128 + bn);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
  }
  if (n <= 32) return 32;
  if (n <= 64) return 64;
  long long best_cost = -1;
  int best = 128;
  for (int bn : {256, 128}) {
    const long long tiles = m_tiles * ((n + bn - 1) / bn);
    // time per tile ~ BN (MMA bound), but a 128-wide pair tile stages 24 KB per 2 MFLOP (256-wide: 32 KB per 4 MFLOP) and is
    // shared-memory bound: measured 1.4-1.7x slower per column (M=16384 N=1152 K=4304: 0.187 ms at BN=128 vs 0.137 ms at BN=256
    // although N = 4.5 tiles of 256; profiles/gemm_epilogue_r1_v16.txt)
    const long long cost = ((tiles + sms - 1) / sms) * bn * (bn == 128 ? 8 : 5);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

extern "C" int svla_gemm(const SvlaGemmArgs* g, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVLA_REQUIRE(g && g->a && g->w, "svla_gemm: null operand");
  SVLA_REQUIRE(g->m > 0 && g->n > 0 && g->k > 0, "svla_gemm: empty problem m=%lld n=%lld k=%lld", (long long)g->m,
               (long long)g->n, (long long)g->k);
  const bool conv = (g->flags & SVLA_GEMM_CONV3X3) != 0;
  const bool geglu = (g->flags & SVLA_GEMM_GEGLU) != 0;
  SVLA_REQUIRE(!geglu || (g->out_bf16 && !g->out_f32 && !g->out_relu_bf16 && !g->res_bf16 && !g->res2_bf16 && !g->res_f32 && (g->n % 2) == 0),
               "svla_gemm: GEGLU mode needs out_bf16 only and even n");
  SVLA_REQUIRE(g->out_bf16 || g->out_f32 || g->out_relu_bf16, "svla_gemm: no output");
  SVLA_REQUIRE(g->m * g->ldo < (1LL << 32), "svla_gemm: output of %lld x %lld elements exceeds the 32-bit offset range", (long long)g->m, (long long)g->ldo);
  SVLA_REQUIRE(!(g->flags & SVLA_GEMM_ACCUM_F32) || g->out_f32, "svla_gemm: ACCUM_F32 needs out_f32");

  EpiParams ep{};
  ep.bias = g->bias; ep.colscale = g->colscale;
  ep.res_bf16 = static_cast<const __nv_bfloat16*>(g->res_bf16); ep.res_f32 = g->res_f32;
  ep.res2_bf16 = static_cast<const __nv_bfloat16*>(g->res2_bf16); ep.res_mod = g->res_mod;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(g->out_bf16); ep.out_f32 = g->out_f32;
  ep.out_relu = static_cast<__nv_bfloat16*>(g->out_relu_bf16);
  ep.m = g->m; ep.n = g->n; ep.ldo = g->ldo;
  ep.alpha = g->alpha; ep.act_param = g->act_param; ep.act = g->act; ep.flags = g->flags;
  int cpad = 0, c_chunks = 1;
  if (conv) {
    SVLA_REQUIRE(g->nb > 0 && g->h > 0 && g->wd > 0 && g->c > 0 && (g->c % 8) == 0, "svla_gemm: bad conv geometry");
    SVLA_REQUIRE(g->m == static_cast<int64_t>(g->nb) * g->h * g->wd, "svla_gemm: conv m != nb*h*w");
    cpad = (g->c + kBK - 1) / kBK * kBK;
    c_chunks = cpad / kBK;
    SVLA_REQUIRE(g->k == 9LL * cpad && g->ldw >= g->k, "svla_gemm: conv weight must be [n][9][%d]", cpad);
    ep.nb = g->nb; ep.h = g->h; ep.wd = g->wd;
    ep.tiles_h = (g->h + kConvTileH - 1) / kConvTileH;
    ep.tiles_w = (g->wd + kConvTileW - 1) / kConvTileW;
  } else {
    SVLA_REQUIRE(g->lda >= g->k && (g->lda % 8) == 0, "svla_gemm: lda=%lld must be >= k and a multiple of 8", (long long)g->lda);
  }
  SVLA_REQUIRE(g->ldw >= g->k && (g->ldw % 8) == 0, "svla_gemm: ldw=%lld must be >= k and a multiple of 8", (long long)g->ldw);
  SVLA_REQUIRE((reinterpret_cast<uintptr_t>(g->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->w) & 15) == 0,
               "svla_gemm: operands must be 16-byte aligned");

  SVLA_REQUIRE(!g->a2 || (g->impl != 1 && !conv), "svla_gemm: the K extension is implemented by the tcgen05 linear kernel only");
  if (g->impl == 1) {
    const long long ncols = geglu ? g->n / 2 : g->n;
    const long long total = g->m * ncols;
    const int threads = 128;
    const long long blocks = (total + threads - 1) / threads;
    svla_gemm_simt_kernel<<<static_cast<unsigned>(blocks), threads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(g->a), static_cast<const __nv_bfloat16*>(g->w), ep, g->k, g->lda, g->ldw,
        conv ? 1 : 0, g->c, cpad);
    SVLA_LAUNCH_CHECK("svla_gemm_simt");
    return 0;
  }

  // Row-tile conv: few output channels (N <= 128) on wide maps, where the 9x activation re-fetch of the patch-tile kernel is
  // L2-bound.  impl 4 forces it, SVLA_CONV_ROWTILE=0 disables it.
  {
    static const int rowtile_env = getenv("SVLA_CONV_ROWTILE") ? atoi(getenv("SVLA_CONV_ROWTILE")) : 1;
    const int tw = conv ? (g->wd + kBM - 1) / kBM : 0;
    const bool fits = conv && g->n <= 128 && g->wd >= kBM && static_cast<long long>(tw) * kBM * 3 <= 4LL * g->wd;     // <= 33 % padded columns
    if (conv && (g->impl == 4 || (g->impl == 0 && rowtile_env != 0 && fits && !g->block_n))) {
      SVLA_REQUIRE(g->n <= 128, "svla_gemm(conv row-tile): n=%lld > 128", (long long)g->n);
      const int bnr = g->n <= 32 ? 32 : (g->n <= 64 ? 64 : 128);
      ep.rowtile = 1;
      ep.tiles_w = tw;
      ep.tiles_h = g->h;
      CUtensorMap tra, trb;
      int rcr = encode_nhwc(&tra, g->a, g->nb, g->h, g->wd, g->c, kBM + 2, 1);
      SVLA_REQUIRE(rcr == 0, "svla_gemm(conv row-tile): cuTensorMapEncodeTiled(A) failed (%d)", rcr);
      rcr = encode_2d(&trb, g->w, static_cast<uint64_t>(g->k), static_cast<uint64_t>(g->n), static_cast<uint64_t>(g->ldw), kBK, static_cast<uint32_t>(bnr));
      SVLA_REQUIRE(rcr == 0, "svla_gemm(conv row-tile): cuTensorMapEncodeTiled(W) failed (%d)", rcr);
      const long long mt = static_cast<long long>(g->nb) * g->h * tw, ntl = (g->n + bnr - 1) / bnr;
      // bf16 TMA-store epilogue through a 4-D map {channel, x, y, image} when the conv has ONE bf16 output and no residual operands
      // (not for 64-wide tiles: there each warp owns 32 columns, half of a 128-byte store box)
      CUtensorMap tro = tra;
      static const bool legacy_epi_rt = getenv("SVLA_GEMM_LEGACY_EPI") != nullptr;
      if (!legacy_epi_rt && bnr != 64 && !geglu && g->out_bf16 && !g->out_f32 && !g->out_relu_bf16 && !g->res_bf16 && !g->res2_bf16 && !g->res_f32 &&
          (g->ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(g->out_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->bias) & 15) == 0 &&
          (reinterpret_cast<uintptr_t>(g->colscale) & 15) == 0 && mt < (1LL << 31)) {
        auto fn = get_encode_fn();
        SVLA_REQUIRE(fn != nullptr, "svla_gemm(conv row-tile): cuTensorMapEncodeTiled unavailable");
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(g->n), static_cast<cuuint64_t>(g->wd), static_cast<cuuint64_t>(g->h), static_cast<cuuint64_t>(g->nb)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(g->ldo) * 2, static_cast<cuuint64_t>(g->wd) * g->ldo * 2,
                                 static_cast<cuuint64_t>(g->h) * g->wd * g->ldo * 2};
        cuuint32_t box[4] = {64, 32, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = fn(&tro, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, g->out_bf16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SVLA_REQUIRE(r == CUDA_SUCCESS, "svla_gemm(conv row-tile): cuTensorMapEncodeTiled(out) failed (%d)", static_cast<int>(r));
        ep.tma_out = 1;
      }
      if (bnr == 32) return launch_rowtile<32>(tra, trb, tro, ep, mt, ntl, c_chunks, cpad, st);
      if (bnr == 64) return launch_rowtile<64>(tra, trb, tro, ep, mt, ntl, c_chunks, cpad, st);
      return launch_rowtile<128>(tra, trb, tro, ep, mt, ntl, c_chunks, cpad, st);
    }
  }

  const long long m_tiles = conv ? static_cast<long long>(g->nb) * ep.tiles_h * ep.tiles_w : (g->m + kBM - 1) / kBM;
  int bn = g->block_n ? g->block_n : pick_block_n(m_tiles, g->n);
  SVLA_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "svla_gemm: block_n=%d unsupported", bn);
  const long long n_tiles = (g->n + bn - 1) / bn;
  const int kb_main = conv ? 9 * c_chunks : static_cast<int>((g->k + kBK - 1) / kBK);
  const bool has_ext = g->a2 != nullptr;
  if (has_ext) {
    SVLA_REQUIRE(!conv && g->w2 && g->k2 > 0, "svla_gemm: the K extension needs a2, w2, k2 > 0 and no conv mode");
    SVLA_REQUIRE(g->lda2 >= g->k2 && (g->lda2 % 8) == 0 && g->ldw2 >= g->k2 && (g->ldw2 % 8) == 0,
                 "svla_gemm: lda2 / ldw2 must be >= k2 and multiples of 8");
    SVLA_REQUIRE((reinterpret_cast<uintptr_t>(g->a2) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->w2) & 15) == 0,
                 "svla_gemm: extension operands must be 16-byte aligned");
  }
  const int kblocks = kb_main + (has_ext ? static_cast<int>((g->k2 + kBK - 1) / kBK) : 0);

  CUtensorMap ta, tb;
  int rc;
  if (conv) rc = encode_nhwc(&ta, g->a, g->nb, g->h, g->wd, g->c);
  else rc = encode_2d(&ta, g->a, static_cast<uint64_t>(g->k), static_cast<uint64_t>(g->m), static_cast<uint64_t>(g->lda), kBK, kBM);
  SVLA_REQUIRE(rc == 0, "svla_gemm: cuTensorMapEncodeTiled(A) failed (%d)", rc);
  // impl: 0 = auto, 2 = force 1-CTA, 3 = force the 2-CTA (cta_group::2) kernel. Auto uses CTA pairs for large GEMMs.
  const bool pair_ok = (bn == 128 || bn == 256) && m_tiles >= 2;
  static const bool no_pair = getenv("SVLA_GEMM_NO_PAIR") != nullptr;      // debugging switch
  const bool use_pair = pair_ok && (g->impl == 3 || (g->impl == 0 && !no_pair && m_tiles * n_tiles >= 2LL * svla_num_sms()));
  rc = encode_2d(&tb, g->w, static_cast<uint64_t>(g->k), static_cast<uint64_t>(g->n), static_cast<uint64_t>(g->ldw), kBK,
                 static_cast<uint32_t>(use_pair ? bn / 2 : bn));     // a CTA of a pair stages half of the W tile
  SVLA_REQUIRE(rc == 0, "svla_gemm: cuTensorMapEncodeTiled(W) failed (%d)", rc);

  // TMA-store epilogue: one output tensor, no residual operands, 128/256-wide tiles, 16-byte aligned rows and vectors
  // (SVLA_GEMM_LEGACY_EPI=1 keeps the register/LDG/STG epilogue for A/B measurements)
  CUtensorMap to = ta;
  static const bool legacy_epi = getenv("SVLA_GEMM_LEGACY_EPI") != nullptr;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!legacy_epi && !conv && !geglu && bn >= 128 && !g->res_bf16 && !g->res2_bf16 && !g->res_f32 && !g->out_relu_bf16 &&
      al16(g->bias) && al16(g->colscale) && g->m < (1LL << 31) && g->n < (1LL << 31)) {
    if (g->out_bf16 && !g->out_f32 && (g->ldo % 8) == 0 && al16(g->out_bf16)) {
      ep.tma_out = 1;
      rc = encode_out(&to, g->out_bf16, false, static_cast<uint64_t>(g->n), static_cast<uint64_t>(g->m), static_cast<uint64_t>(g->ldo));
    } else if (g->out_f32 && !g->out_bf16 && (g->ldo % 4) == 0 && al16(g->out_f32)) {
      ep.tma_out = (g->flags & SVLA_GEMM_ACCUM_F32) ? 3 : 2;
      rc = encode_out(&to, g->out_f32, true, static_cast<uint64_t>(g->n), static_cast<uint64_t>(g->m), static_cast<uint64_t>(g->ldo));
    }
    SVLA_REQUIRE(rc == 0, "svla_gemm: cuTensorMapEncodeTiled(out) failed (%d)", rc);
  }

  CUtensorMap ta2 = ta, tb2 = tb;
  if (has_ext) {
    rc = encode_2d(&ta2, g->a2, static_cast<uint64_t>(g->k2), static_cast<uint64_t>(g->m), static_cast<uint64_t>(g->lda2), kBK, kBM);
    SVLA_REQUIRE(rc == 0, "svla_gemm: cuTensorMapEncodeTiled(A2) failed (%d)", rc);
    rc = encode_2d(&tb2, g->w2, static_cast<uint64_t>(g->k2), static_cast<uint64_t>(g->n), static_cast<uint64_t>(g->ldw2), kBK,
                   static_cast<uint32_t>(use_pair ? bn / 2 : bn));
    SVLA_REQUIRE(rc == 0, "svla_gemm: cuTensorMapEncodeTiled(W2) failed (%d)", rc);
  }
  const int cv = conv ? 1 : 0;
  if (use_pair) {
    if (bn == 128) return launch_tc<128, 2>(ta, tb, to, ta2, tb2, ep, m_tiles, n_tiles, kblocks, cv, c_chunks, kb_main, st);
    return launch_tc<256, 2>(ta, tb, to, ta2, tb2, ep, m_tiles, n_tiles, kblocks, cv, c_chunks, kb_main, st);
  }
  switch (bn) {
    case 32: return launch_tc<32, 1>(ta, tb, to, ta2, tb2, ep, m_tiles, n_tiles, kblocks, cv, c_chunks, kb_main, st);
    case 64: return launch_tc<64, 1>(ta, tb, to, ta2, tb2, ep, m_tiles, n_tiles, kblocks, cv, c_chunks, kb_main, st);
    case 128: return launch_tc<128, 1>(ta, tb, to, ta2, tb2, ep, m_tiles, n_tiles, kblocks, cv, c_chunks, kb_main, st);
    default: return launch_tc<256, 1>(ta, tb, to, ta2, tb2, ep, m_tiles, n_tiles, kblocks, cv, c_chunks, kb_main, st);
  }
}
