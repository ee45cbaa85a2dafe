// Backward of G2 on the 5th-generation tensor cores (the LoRA fine-tune step, BASELINE.json config #5): tcgen05.mma with the score /
// dP tiles, the bf16 dS / P operand and the gradient accumulator all in TMEM, operands staged by TMA.
// Reference math: model/modeling_gemma2.py:169-195 (GQA, tanh soft-capping) under the training masks of
// model/modeling_spatialvla.py:258-306, HF siglip/modeling_siglip.py:252-312; torch autograd in the reference.  Closed form:
// oracle/backward_ref.softcap_attention_bwd.  With s = scale q.k, c = cap tanh(s / cap), p = exp(c - lse), delta = rowsum(dO * O):
//     dV = P^T dO      dP = dO V^T      dS = P * (dP - delta) * (1 - (c / cap)^2) * scale      dQ = dS K      dK = dS^T Q
//
// ONE kernel template, three sweeps (KIND); a CTA owns 128 rows ("resident" operands, loaded once) of one (batch, head) and streams
// 64-row tiles of the other sequence through shared memory:
//     KIND   rows      resident     streamed      X = R1 T1^T   Y = R2 T2^T   Z (bf16, TMEM)   ACC += Z T3
//     DQ     queries   Q, dO        K_j, V_j      S             dP            dS               dQ += dS K_j
//     DK     keys      K, V         Q_i, dO_i     S^T           dP^T          dS^T             dK += dS^T Q_i     (all heads of the GQA group)
//     DV     keys      K            Q_i, dO_i     S^T           --            P^T              dV += P^T dO_i
// X and Y are SS MMAs (both operands K-major, SWIZZLE_128B) into double-buffered TMEM tiles; the softmax warps (thread = row, two
// warps per TMEM lane quadrant splitting the 64 columns) read them with tcgen05.ld, evaluate p / dS from the forward pass's log2-sum-exp
// (no online softmax: nothing is rescaled) and write Z as packed bf16 over the first 32 columns of the X tile (tcgen05.st); the third
// MMA takes Z from TMEM (TS) against the SAME streamed tile read MN-major.  The issuer runs one tile ahead (X_{j+1}, Y_{j+1} before
// Z_j T3_j), so the tensor core works while tile j is in the softmax warps.  TMEM: 2 x 64 (X) + 2 x 64 (Y) + D (ACC) = 512 columns at
// d = 256; that is why dK and dV are separate sweeps (their two accumulators alone would fill TMEM) -- S^T is recomputed once.
// Shared memory at d = 256: 128 KB resident + 3 x 32 KB stream slots (T1 double-buffered, T2 single: it is free again as soon as
// Y_j retires) = 224 KB; the DV sweep (one resident matrix) double-buffers both.
// Head dims below the tile width (SigLIP 72 -> 128) use the zero-padding 4-D tensor maps of the forward kernel (PAD).
#include <cudaTypedefs.h>
#include "../../include/spatialvla_b200.h"
#include "tc_ptx.cuh"

namespace svla_attn_bwd_tc {
using namespace svla_ptx;

constexpr int kBM = 128;            // resident rows per CTA (UMMA M)
constexpr int kBT = 64;             // streamed rows per tile (UMMA N of X / Y, K of the third MMA)
constexpr int kSoftWarps = 8;
constexpr int kThreads = 64 + 32 * kSoftWarps;
constexpr float kLog2e = 1.4426950408889634f;
enum { KIND_DQ = 0, KIND_DK = 1, KIND_DV = 2 };

struct Params {
  __nv_bfloat16* out;               // dq | dk | dv
  long long o_bs, o_ss;
  const float* lse2;                // fp32 [batch, hq, stat_stride], log2 domain (forward kernel)
  const float* delta;               // fp32 [batch, hq, stat_stride]
  long long stat_stride;
  int hq, hkv, sq, sk, d;
  float scale, softcap;
  int causal, prefix;
  int window;                       // > 0: key j masked for query i when i + (sk - sq) - j >= window (sliding-window layers)
};

template <int D, int KIND> struct Cfg {
  static constexpr int kChunks = D / 64;
  static constexpr int kRBytes = kBM * D * 2;                       // one resident matrix
  static constexpr int kTBytes = kBT * D * 2;                       // one streamed tile
  static constexpr int kNumR = (KIND == KIND_DV) ? 1 : 2;
  // T2 slots: the DV sweep reads T2 (dO) in the LAST MMA of a tile, so it is double-buffered like T1; in the DQ / DK sweeps T2 is
  // only read by Y_j, which retires early -- one slot, and at d = 256 there is no room for a second one next to two resident matrices
  static constexpr int kNumT2 = (KIND == KIND_DV || D <= 128) ? 2 : 1;
  static constexpr bool kHasY = KIND != KIND_DV;
  static constexpr int kTmemX = 0;                                  // 2 x kBT
  static constexpr int kTmemY = 2 * kBT;                            // 2 x kBT
  static constexpr int kTmemAcc = kHasY ? 4 * kBT : 2 * kBT;        // D columns
  // the DV sweep at D <= 128 needs 2 x 64 (X) + D (ACC) <= 256 columns and <= 96 KB of shared memory: two CTAs per SM
  static constexpr int kTmemCols = (kTmemAcc + D <= 256) ? 256 : 512;
  static constexpr int kCtasPerSm = (kTmemCols == 256 && kNumR * kRBytes + (2 + kNumT2) * kTBytes + 1536 <= 113 * 1024) ? 2 : 1;
  static constexpr int kSmemBytes = kNumR * kRBytes + (2 + kNumT2) * kTBytes + 1024 /*align*/ + 512 /*barriers, TMEM slot*/;
  static_assert(kTmemAcc + D <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_poly(float u) {        // exact to fp32 rounding for |u| < 0.35 (same evaluation as the forward)
  const float u2 = u * u;
  if (u2 < 0.1225f) {
    float pl = 62.f / 2835.f;
    pl = fmaf(pl, u2, -17.f / 315.f);
    pl = fmaf(pl, u2, 2.f / 15.f);
    pl = fmaf(pl, u2, -1.f / 3.f);
    pl = fmaf(pl, u2, 1.f);
    return u * pl;
  }
  return tanhf(u);
}

// PAD: real head dim p.d < D, 4-D tensor maps {d, head, token, batch} zero-fill the padding (see attention_tc.cu)
template <int D, int KIND, bool PAD>
__global__ void __launch_bounds__(kThreads, Cfg<D, KIND>::kCtasPerSm)
svla_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_r1, const __grid_constant__ CUtensorMap tm_r2,
                        const __grid_constant__ CUtensorMap tm_t1, const __grid_constant__ CUtensorMap tm_t2, const Params p) {
  using C = Cfg<D, KIND>;
  constexpr bool kRowsAreQueries = KIND == KIND_DQ;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sR1 = smem;                                              // [chunk][128 rows][128 B]
  uint8_t* sR2 = smem + C::kRBytes;                                 // (absent in the DV sweep)
  uint8_t* sT1 = smem + C::kNumR * C::kRBytes;                      // [2][chunk][64 rows][128 B]
  uint8_t* sT2 = sT1 + 2 * C::kTBytes;                              // [kNumT2][...]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT2 + C::kNumT2 * C::kTBytes);
  uint64_t* r_full = bars;                 // 1
  uint64_t* t1_full = bars + 1;            // [2]
  uint64_t* t1_empty = bars + 3;           // [2]
  uint64_t* t2_full = bars + 5;            // [2]
  uint64_t* t2_empty = bars + 7;           // [2]
  uint64_t* x_full = bars + 9;             // [2]  X_j (and Y_j) complete
  uint64_t* z_full = bars + 11;            // [2]  8 softmax warps wrote Z_j
  uint64_t* x_empty = bars + 13;           // [2]  third MMA of tile j complete: X / Y buffers free
  uint64_t* acc_full = bars + 15;          // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, r0 = blockIdx.x * kBM;
  const int G = p.hq / p.hkv;
  // DQ: blockIdx.y = query head; DK / DV: blockIdx.y = kv head, the stream covers the G query heads of its group
  const int h_res = blockIdx.y;                                     // head of the resident operands and of the output
  const int hk = kRowsAreQueries ? h_res / G : h_res;
  const int off = p.sk - p.sq;
  // streamed tiles
  int t_lo = 0, t_hi;                                               // tile range of the streamed sequence (per head)
  if (kRowsAreQueries) {
    t_hi = (p.sk + kBT - 1) / kBT;
    if (p.causal) t_hi = max(1, min(t_hi, (max(min(r0 + kBM, p.sq) + off, p.prefix) + kBT - 1) / kBT));      // keys above the diagonal
  } else {
    t_hi = (p.sq + kBT - 1) / kBT;
    if (p.causal && r0 >= p.prefix) t_lo = min(t_hi - 1, max(0, r0 - off) / kBT);                              // queries that see key r0
  }
  const int tiles_per_head = t_hi - t_lo;
  const int n_tiles = kRowsAreQueries ? tiles_per_head : tiles_per_head * G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_r1);
    tma_prefetch_desc(&tm_t1);
    tma_prefetch_desc(&tm_t2);
    if (C::kNumR == 2) tma_prefetch_desc(&tm_r2);
    mbar_init(r_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&t1_full[s], 1);
      mbar_init(&t1_empty[s], 1);
      mbar_init(&t2_full[s], 1);
      mbar_init(&t2_empty[s], 1);
      mbar_init(&x_full[s], 1);
      mbar_init(&z_full[s], kSoftWarps);
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::kTmemCols, 1>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // streamed tile index -> (head of the streamed operands, first row of the tile)
  auto tile_head = [&](int j) { return kRowsAreQueries ? hk : hk * G + j / tiles_per_head; };
  auto tile_row0 = [&](int j) { return (t_lo + (kRowsAreQueries ? j : j % tiles_per_head)) * kBT; };
  const int t_len = kRowsAreQueries ? p.sk : p.sq;                  // length of the streamed sequence
  auto valid16 = [&](int j) { return min(kBT, (t_len - tile_row0(j) + 15) & ~15); };

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      auto load = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int head, int row, int rows_bytes_per_chunk) {
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
          if constexpr (PAD) tma_load_4d(dst + c * rows_bytes_per_chunk, tm, bar, c * 64, head, row, b);
          else tma_load_3d(dst + c * rows_bytes_per_chunk, tm, bar, head * D + c * 64, row, b);
        }
      };
      mbar_expect_tx(r_full, C::kNumR * C::kRBytes);
      load(sR1, &tm_r1, r_full, kRowsAreQueries ? h_res : hk, r0, kBM * 128);
      if (C::kNumR == 2) load(sR2, &tm_r2, r_full, kRowsAreQueries ? h_res : hk, r0, kBM * 128);
      for (int j = 0; j < n_tiles; ++j) {
        const int s1 = j & 1, s2 = (C::kNumT2 == 2) ? (j & 1) : 0;
        const int th = kRowsAreQueries ? hk : tile_head(j);         // K / V live under the kv head; Q / dO under the query head
        mbar_wait(&t1_empty[s1], ((j >> 1) & 1) ^ 1u);
        mbar_expect_tx(&t1_full[s1], C::kTBytes);
        load(sT1 + s1 * C::kTBytes, &tm_t1, &t1_full[s1], th, tile_row0(j), kBT * 128);
        mbar_wait(&t2_empty[s2], ((C::kNumT2 == 2 ? (j >> 1) : j) & 1) ^ 1u);
        mbar_expect_tx(&t2_full[s2], C::kTBytes);
        load(sT2 + s2 * C::kTBytes, &tm_t2, &t2_full[s2], th, tile_row0(j), kBT * 128);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer: converged warp, an ELECTED lane issues the MMAs and commits of a
    // phase, descriptors are constants + start-address increments (issued from a divergent lane-0 region every tcgen05.mma cost ~70
    // clocks: 36 MMAs per 64-row tile at d = 256 made the issue rate, not the tensor pipe or the softmax warps, the bound)
    {
      const int ksteps = PAD ? ((p.d + 15) >> 4) : D / 16;          // contraction steps over the head dimension
      const uint32_t idesc_acc = make_idesc(kBM, PAD ? ((p.d + 15) & ~15) : D, 1);
      const uint64_t r1_desc = make_kmajor_sw128_desc(smem_u32(sR1)), r2_desc = make_kmajor_sw128_desc(smem_u32(sR2));
      auto issue_acc = [&](int j) {                                 // ACC += Z_j T3_j
        const int st = j & 1, s2 = (C::kNumT2 == 2) ? (j & 1) : 0;
        mbar_wait(&z_full[st], (j >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tbase = (KIND == KIND_DV) ? smem_u32(sT2 + s2 * C::kTBytes) : smem_u32(sT1 + st * C::kTBytes);
          const uint64_t b_desc = make_mnmajor_sw128_desc(tbase, kBT * 128);
          const int ks = valid16(j) >> 4;
#pragma unroll
          for (int kk = 0; kk < kBT / 16; ++kk) {
            if (kk < ks)
              umma_bf16_ts(tmem_base + C::kTmemAcc, tmem_base + C::kTmemX + st * kBT + kk * 8, b_desc + static_cast<uint64_t>((kk * 16 * 128) >> 4),
                           idesc_acc, static_cast<uint32_t>(j > 0 || kk > 0));
          }
          umma_commit(&x_empty[st]);
          umma_commit(&t1_empty[st]);
          if (KIND == KIND_DV) umma_commit(&t2_empty[s2]);
        }
        __syncwarp();
      };
      mbar_wait(r_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1, s2 = (C::kNumT2 == 2) ? (j & 1) : 0;
        mbar_wait(&t1_full[st], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&x_empty[st], ((j - 2) >> 1) & 1);    // the third MMA of tile j-2 has drained Z / Y of this buffer
        tc_fence_after();
        const uint32_t idesc_xy = make_idesc(kBM, valid16(j), 0);
        if (elect_one()) {
          const uint64_t t1_desc = make_kmajor_sw128_desc(smem_u32(sT1 + st * C::kTBytes));
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            if (kk < ksteps)
              umma_bf16(tmem_base + C::kTmemX + st * kBT, r1_desc + static_cast<uint64_t>(((kk >> 2) * (kBM * 128) + (kk & 3) * 32) >> 4),
                        t1_desc + static_cast<uint64_t>(((kk >> 2) * (kBT * 128) + (kk & 3) * 32) >> 4), idesc_xy, static_cast<uint32_t>(kk > 0));
          }
          if (!C::kHasY) umma_commit(&x_full[st]);
        }
        __syncwarp();
        if (C::kHasY) {
          mbar_wait(&t2_full[s2], (C::kNumT2 == 2 ? (j >> 1) : j) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t t2_desc = make_kmajor_sw128_desc(smem_u32(sT2 + s2 * C::kTBytes));
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
              if (kk < ksteps)
                umma_bf16(tmem_base + C::kTmemY + st * kBT, r2_desc + static_cast<uint64_t>(((kk >> 2) * (kBM * 128) + (kk & 3) * 32) >> 4),
                          t2_desc + static_cast<uint64_t>(((kk >> 2) * (kBT * 128) + (kk & 3) * 32) >> 4), idesc_xy, static_cast<uint32_t>(kk > 0));
            }
            umma_commit(&t2_empty[s2]);                             // T2_j is only read by Y_j in the DQ / DK sweeps
            umma_commit(&x_full[st]);
          }
          __syncwarp();
        } else {
          mbar_wait(&t2_full[s2], (j >> 1) & 1);                    // DV: dO_j must have landed before the third MMA reads it
        }
        if (j > 0) issue_acc(j - 1);
      }
      if (n_tiles > 0) issue_acc(n_tiles - 1);
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ============================================================ softmax warps: thread == resident row
    constexpr int HC = kBT / 2;                                     // streamed columns owned by this warp
    const int q = warp & 3, ch = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ri = r0 + row;                                        // query (DQ) or key (DK / DV) index of this thread
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float c1 = p.softcap > 0.f ? p.scale / p.softcap : 0.f, c2 = p.softcap > 0.f ? p.softcap * kLog2e : p.scale * kLog2e;
    float lse_r = 0.f, del_r = 0.f;
    if (kRowsAreQueries) {
      const long long o = (static_cast<long long>(b) * p.hq + h_res) * p.stat_stride + min(ri, p.sq - 1);
      lse_r = p.lse2[o];
      del_r = p.delta[o];
    }
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const int c0 = tile_row0(j) + ch * HC;                        // first streamed index of this warp's columns
      // per-column statistics of the DK / DV sweeps: lane = column, broadcast with shuffles below
      float lse_c = 0.f, del_c = 0.f;
      if (!kRowsAreQueries) {
        const long long o = (static_cast<long long>(b) * p.hq + tile_head(j)) * p.stat_stride + min(c0 + lane, p.sq - 1);
        lse_c = p.lse2[o];
        del_c = p.delta[o];
      }
      mbar_wait(&x_full[st], (j >> 1) & 1);
      tc_fence_after();
      float x[HC], y[HC];
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + C::kTmemX + st * kBT + ch * HC + lane_addr, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
        if (C::kHasY) {
          tmem_ld32(tmem_base + C::kTmemY + st * kBT + ch * HC + lane_addr, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(r[i]);
        }
      }
      // Z overwrites columns [0, 32) of the X tile: both warps of the pair must have their X columns in registers first
      tc_fence_before();
      pair_sync();
      // whole-warp shortcut: every (row, column) of this warp's block is in range and visible
      const int rlo = r0 + q * 32, rhi = rlo + 32;                  // rows of this warp
      bool plain;
      if (kRowsAreQueries) {
        plain = p.window == 0 && rhi <= p.sq && c0 + HC <= p.sk && (!p.causal || (c0 + HC - 1) <= max(rlo + off, p.prefix - 1));
      } else {
        plain = p.window == 0 && rhi <= p.sk && c0 + HC <= p.sq && (!p.causal || (rhi - 1) <= max(c0 + off, p.prefix - 1));
      }
      uint32_t zk[HC / 2];
      // Two columns per FMUL2 / FFMA2 (the scalar loop with a per-element tanh branch and two shuffles per element made the 8
      // softmax warps -- one CTA per SM -- the bound of the sweep).  Column statistics of the DK / DV sweeps are fetched with one
      // shuffle per PAIR and value; the soft-cap polynomial is evaluated unconditionally and a warp vote selects the libm path.
      bool big = false;
      if (p.softcap > 0.f) {
        float u2max = 0.f;
#pragma unroll
        for (int i = 0; i < HC; ++i) { const float u = x[i] * c1; u2max = fmaxf(u2max, u * u); }
        big = __any_sync(0xffffffffu, u2max >= 0.1225f);
      }
      const uint64_t c1p = pack_f32x2(c1, c1), c2p = pack_f32x2(c2, c2), onep = pack_f32x2(1.f, 1.f), scp = pack_f32x2(p.scale, p.scale);
      const uint64_t k9 = pack_f32x2(62.f / 2835.f, 62.f / 2835.f), k7 = pack_f32x2(-17.f / 315.f, -17.f / 315.f);
      const uint64_t k5 = pack_f32x2(2.f / 15.f, 2.f / 15.f), k3 = pack_f32x2(-1.f / 3.f, -1.f / 3.f);
#pragma unroll
      for (int i = 0; i < HC; i += 2) {
        // per-column statistics (DK / DV): lanes i, i + 1 hold the columns' values
        float lse0 = lse_r, lse1 = lse_r, del0 = del_r, del1 = del_r;
        if (!kRowsAreQueries) {
          lse0 = __shfl_sync(0xffffffffu, lse_c, i); lse1 = __shfl_sync(0xffffffffu, lse_c, i + 1);
          if (KIND != KIND_DV) { del0 = __shfl_sync(0xffffffffu, del_c, i); del1 = __shfl_sync(0xffffffffu, del_c, i + 1); }
        }
        const uint64_t xx = pack_f32x2(x[i], x[i + 1]);
        uint64_t s2, fac = onep;
        if (p.softcap > 0.f) {
          uint64_t th;
          if (!big) {
            const uint64_t u = mul_f32x2(xx, c1p), u2 = mul_f32x2(u, u);
            uint64_t pl = fma_f32x2(k9, u2, k7);
            pl = fma_f32x2(pl, u2, k5);
            pl = fma_f32x2(pl, u2, k3);
            pl = fma_f32x2(pl, u2, onep);
            th = mul_f32x2(u, pl);
          } else {
            th = pack_f32x2(tanhf(x[i] * c1), tanhf(x[i + 1] * c1));
          }
          s2 = mul_f32x2(c2p, th);
          fac = fma_f32x2(th, mul_f32x2(th, pack_f32x2(-1.f, -1.f)), onep);          // 1 - th^2
        } else {
          s2 = mul_f32x2(xx, c2p);
        }
        float t0, t1;
        unpack_f32x2(add_f32x2(s2, pack_f32x2(-lse0, -lse1)), t0, t1);
        float pr0 = ex2f(t0), pr1 = ex2f(t1);
        if (!plain) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ci = c0 + i + e;                              // streamed index
            const int qi = kRowsAreQueries ? ri : ci, kj = kRowsAreQueries ? ci : ri;
            const bool masked = qi >= p.sq || kj >= p.sk || (p.causal && kj > max(qi + off, p.prefix - 1)) ||
                                (p.window > 0 && qi + off - kj >= p.window);
            if (masked) { if (e == 0) pr0 = 0.f; else pr1 = 0.f; }
          }
        }
        if (KIND == KIND_DV) {
          zk[i >> 1] = pack_bf16x2(pr0, pr1);
        } else {
          // pr * (y - del) * fac * scale, same association as the scalar expression
          const uint64_t dy = add_f32x2(pack_f32x2(y[i], y[i + 1]), pack_f32x2(-del0, -del1));
          float z0, z1;
          unpack_f32x2(mul_f32x2(mul_f32x2(mul_f32x2(pack_f32x2(pr0, pr1), dy), fac), scp), z0, z1);
          zk[i >> 1] = pack_bf16x2(z0, z1);
        }
      }
      tc_fence_after();
      tmem_st16(tmem_base + C::kTmemX + st * kBT + ch * (HC / 2) + lane_addr, zk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&z_full[st]);
    }
    // ---- epilogue: ACC (fp32, TMEM) -> bf16 -> global; this warp writes its half of the columns
    constexpr int DH = D / 2;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int r_len = kRowsAreQueries ? p.sq : p.sk;
    __nv_bfloat16* og = p.out + b * p.o_bs + static_cast<long long>(ri) * p.o_ss + static_cast<long long>(h_res) * (PAD ? p.d : D);
#pragma unroll 1
    for (int cc = ch * DH; cc < (ch + 1) * DH; cc += 32) {
      if (PAD && cc >= p.d) break;
      uint32_t r[32];
      if (n_tiles > 0) {
        tmem_ld32(tmem_base + C::kTmemAcc + cc + lane_addr, r);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (ri < r_len) {
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          if (PAD && cc + 8 * v8 >= p.d) break;
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[8 * v8 + 0]), __uint_as_float(r[8 * v8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(r[8 * v8 + 2]), __uint_as_float(r[8 * v8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(r[8 * v8 + 4]), __uint_as_float(r[8 * v8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(r[8 * v8 + 6]), __uint_as_float(r[8 * v8 + 7]));
          *reinterpret_cast<uint4*>(og + cc + 8 * v8) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols, 1>(tmem_base);
  }
}

// delta[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d]: one warp per (token, head)
__global__ void __launch_bounds__(256)
svla_attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, long long o_bs, long long o_ss,
                       long long do_bs, long long do_ss, int hq, int sq, int d, float* __restrict__ delta, long long stat_stride, long long total) {
  const int lane = threadIdx.x & 31;
  const long long w = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (w >= total) return;
  const int h = static_cast<int>(w % hq);
  const long long bi = w / hq;
  const int i = static_cast<int>(bi % sq);
  const long long b = bi / sq;
  const __nv_bfloat16* op = o + b * o_bs + static_cast<long long>(i) * o_ss + static_cast<long long>(h) * d;
  const __nv_bfloat16* dp = dout + b * do_bs + static_cast<long long>(i) * do_ss + static_cast<long long>(h) * d;
  float acc = 0.f;
  for (int c = lane * 8; c < d; c += 256) {
    const uint4 a = *reinterpret_cast<const uint4*>(op + c), g = *reinterpret_cast<const uint4*>(dp + c);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      acc += bf16_bits_to_float(aw[k] & 0xFFFFu) * bf16_bits_to_float(gw[k] & 0xFFFFu) + bf16_bits_to_float(aw[k] >> 16) * bf16_bits_to_float(gw[k] >> 16);
  }
  acc = warp_sum(acc);
  if (lane == 0) delta[(b * hq + h) * stat_stride + i] = acc;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}
// {cols, tokens, batch} (3-D) or {d, head, tokens, batch} (4-D, zero-padding) view of a strided bf16 activation; box = 64 columns x rows
static int encode(CUtensorMap* tm, bool pad, const void* base, uint64_t d, uint64_t heads, uint64_t tokens, uint64_t batch,
                  uint64_t token_stride, uint64_t batch_stride, uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) return -1;
  CUresult r;
  if (pad) {
    cuuint64_t dims[4] = {d, heads, tokens, batch};
    cuuint64_t strides[3] = {d * 2, token_stride * 2, batch_stride * 2};
    cuuint32_t box[4] = {64, 1, box_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {d * heads, tokens, batch};
    cuuint64_t strides[2] = {token_stride * 2, batch_stride * 2};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

template <int D, int KIND, bool PAD>
static int launch(const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& t1, const CUtensorMap& t2, const Params& p, dim3 grid,
                  cudaStream_t st) {
  using C = Cfg<D, KIND>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(svla_attn_bwd_tc_kernel<D, KIND, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) {
      svla_set_error("svla_attention_bwd(tcgen05): smem opt-in %d failed: %s", C::kSmemBytes, cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  svla_attn_bwd_tc_kernel<D, KIND, PAD><<<grid, kThreads, C::kSmemBytes, st>>>(r1, r2, t1, t2, p);
  SVLA_LAUNCH_CHECK("svla_attn_bwd_tc");
  return 0;
}

template <int D, bool PAD>
static int run_all(const SvlaAttnBwdArgs* a, cudaStream_t st) {
  const uint64_t nb = static_cast<uint64_t>(a->batch), dd = static_cast<uint64_t>(a->d);
  auto bstride = [&](int64_t bs, int64_t ss, int s) { return static_cast<uint64_t>(nb > 1 ? bs : ss * s); };
  CUtensorMap q128, q64, k128, k64, v128, v64, do128, do64;
  int rc = 0;
  rc |= encode(&q128, PAD, a->q, dd, a->hq, a->sq, nb, a->q_ss, bstride(a->q_bs, a->q_ss, a->sq), kBM);
  rc |= encode(&q64, PAD, a->q, dd, a->hq, a->sq, nb, a->q_ss, bstride(a->q_bs, a->q_ss, a->sq), kBT);
  rc |= encode(&do128, PAD, a->dout, dd, a->hq, a->sq, nb, a->do_ss, bstride(a->do_bs, a->do_ss, a->sq), kBM);
  rc |= encode(&do64, PAD, a->dout, dd, a->hq, a->sq, nb, a->do_ss, bstride(a->do_bs, a->do_ss, a->sq), kBT);
  rc |= encode(&k128, PAD, a->k, dd, a->hkv, a->sk, nb, a->k_ss, bstride(a->k_bs, a->k_ss, a->sk), kBM);
  rc |= encode(&k64, PAD, a->k, dd, a->hkv, a->sk, nb, a->k_ss, bstride(a->k_bs, a->k_ss, a->sk), kBT);
  rc |= encode(&v128, PAD, a->v, dd, a->hkv, a->sk, nb, a->v_ss, bstride(a->v_bs, a->v_ss, a->sk), kBM);
  rc |= encode(&v64, PAD, a->v, dd, a->hkv, a->sk, nb, a->v_ss, bstride(a->v_bs, a->v_ss, a->sk), kBT);
  if (rc != 0) return 1;
  Params p{};
  p.lse2 = a->fwd_lse2; p.delta = a->delta; p.stat_stride = a->lse_stride;
  p.hq = a->hq; p.hkv = a->hkv; p.sq = a->sq; p.sk = a->sk; p.d = a->d;
  p.scale = a->scale; p.softcap = a->softcap; p.causal = a->causal; p.prefix = a->causal ? a->causal_prefix : 0;
  p.window = a->window;
  const long long total = static_cast<long long>(a->batch) * a->sq * a->hq;
  svla_attn_delta_kernel<<<static_cast<unsigned>((total + 7) / 8), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(a->out), static_cast<const __nv_bfloat16*>(a->dout), a->o_bs, a->o_ss, a->do_bs, a->do_ss, a->hq, a->sq, a->d,
      a->delta, a->lse_stride, total);
  SVLA_LAUNCH_CHECK("svla_attn_delta");
  const dim3 gq((a->sq + kBM - 1) / kBM, a->hq, a->batch), gk((a->sk + kBM - 1) / kBM, a->hkv, a->batch);
  p.out = static_cast<__nv_bfloat16*>(a->dq); p.o_bs = a->dq_bs; p.o_ss = a->dq_ss;
  if ((rc = launch<D, KIND_DQ, PAD>(q128, do128, k64, v64, p, gq, st)) != 0) return rc;
  p.out = static_cast<__nv_bfloat16*>(a->dk); p.o_bs = a->dk_bs; p.o_ss = a->dk_ss;
  if ((rc = launch<D, KIND_DK, PAD>(k128, v128, q64, do64, p, gk, st)) != 0) return rc;
  p.out = static_cast<__nv_bfloat16*>(a->dv); p.o_bs = a->dv_bs; p.o_ss = a->dv_ss;
  return launch<D, KIND_DV, PAD>(k128, k128, q64, do64, p, gk, st);
}

}  // namespace svla_attn_bwd_tc

// Returns 1 if the tcgen05 kernels do not cover the problem (the caller then uses the warp-MMA kernels), 0 on success, < 0 on error.
int svla_attention_bwd_tc_try(const SvlaAttnBwdArgs* a, void* stream) {
  using namespace svla_attn_bwd_tc;
  if (!a->fwd_lse2 || !a->delta || a->lse_stride < a->sq) return 1;
  const bool exact = a->d == 256 || a->d == 128 || a->d == 64;
  const bool padded = !exact && a->d < 128 && (a->d % 8) == 0;
  if (!exact && !padded) return 1;
  if (a->causal && a->sk < a->sq) return 1;
  const int64_t strides[] = {a->q_bs, a->q_ss, a->k_bs, a->k_ss, a->v_bs, a->v_ss, a->o_bs, a->o_ss, a->do_bs, a->do_ss,
                             a->dq_bs, a->dq_ss, a->dk_bs, a->dk_ss, a->dv_bs, a->dv_ss};
  for (int64_t s : strides)
    if (s % 8) return 1;
  const void* ptrs[] = {a->q, a->k, a->v, a->out, a->dout, a->dq, a->dk, a->dv};
  for (const void* ptr : ptrs)
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return 1;
  if (a->batch > 1 && (a->q_bs <= 0 || a->k_bs <= 0 || a->v_bs <= 0 || a->do_bs <= 0)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->d == 256) return run_all<256, false>(a, st);
  if (a->d == 128) return run_all<128, false>(a, st);
  if (a->d == 64) return run_all<64, false>(a, st);
  return a->d < 64 ? run_all<64, true>(a, st) : run_all<128, true>(a, st);
}
