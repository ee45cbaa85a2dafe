// Tensor-core kernels of the BACKWARD half of the LoRA fine-tune step (BASELINE.json config #5) that are not
// plain GEMMs:
//   * svla_attention_bwd: backward of softmax(softcap(scale Q K^T) + mask) V (model/modeling_gemma2.py:169-195 with the training
//     masks of model/modeling_spatialvla.py:258-306; HF siglip/modeling_siglip.py:252-312) in the flash-attention formulation:
//     nothing of size [Sq, Sk] is ever stored.  Three launches: (0) row statistics -- logsumexp of every query row (recomputed, the
//     forward kernels keep none) and delta = rowsum(dO * O); (1) dQ: one CTA per 64-query tile walks the key tiles; (2) dK, dV: one
//     CTA per 64-key tile walks the query tiles of every head of its GQA group.  Two sweeps instead of one with atomics: no fp32
//     scratch gradients, deterministic sums.  The closed form is oracle/backward_ref.softcap_attention_bwd.
//   * svla_gemm_tn: out[r, n] += scale * sum_m S[m, r] Y[m, n] -- the rank-r LoRA gradient reductions gA = (dY B)^T X and
//     gB^T = (X A^T)^T dY (train/spatialvla_finetune.py:262-302; oracle/backward_ref.lora_linear_bwd), contraction over the TOKEN
//     dimension, split over CTAs, accumulated with fp32 reductions straight into the gradient arena.
// Warp-level mma.sync (bf16 in, fp32 accumulate): these are ~6 % of the step's FLOPs; the dX GEMMs run on the tcgen05 kernel.
#include <cstdlib>
#include <cstring>
#include "mma_sync.cuh"

namespace {
using namespace svla_mma;

// =================================================================================================== gemm_tn
struct TnGroup {
  float* dst;             // fp32 [rg, ld]
  long long ld;
  int r0, rg;             // rows [r0, r0 + rg) of S^T Y
  int col_start, col_stride, ncols;      // dst column j <- Y column col_start + j * col_stride
};
struct TnParams {
  const __nv_bfloat16* s;
  const __nv_bfloat16* y;
  long long lds, ldy, m, m_per_split;
  int r, n;
  float scale;
  int n_groups;
  TnGroup g[4];
};

constexpr int kTnBN = 128, kTnBM = 64, kTnThreads = 256;

template <int RT>
__global__ void __launch_bounds__(kTnThreads)
svla_gemm_tn_kernel(const TnParams p) {
  constexpr int R = 16 * RT;
  constexpr int LDS = R + 8, LDY = kTnBN + 8;
  extern __shared__ __align__(16) uint8_t smem_tn[];
  __nv_bfloat16* sS = reinterpret_cast<__nv_bfloat16*>(smem_tn);          // [2][64][LDS]
  __nv_bfloat16* sY = sS + 2 * kTnBM * LDS;                               // [2][64][LDY]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int col0 = blockIdx.x * kTnBN;
  const long long m_begin = blockIdx.y * p.m_per_split;
  const long long m_end = (m_begin + p.m_per_split < p.m) ? m_begin + p.m_per_split : p.m;
  if (m_begin >= m_end) return;
  const int n_chunks = static_cast<int>((m_end - m_begin + kTnBM - 1) / kTnBM);

  auto load_chunk = [&](int buf, int c) {
    const long long m0 = m_begin + static_cast<long long>(c) * kTnBM;
    constexpr int SC = R / 8;                       // 16-byte pieces per S row
    for (int i = threadIdx.x; i < kTnBM * SC; i += kTnThreads) {
      const int rr = i / SC, cc = i - rr * SC;
      const bool ok = (m0 + rr) < m_end && cc * 8 < p.r;
      cp_async16(sS + (buf * kTnBM + rr) * LDS + cc * 8, p.s + (ok ? (m0 + rr) : 0) * p.lds + (ok ? cc * 8 : 0), ok);
    }
    constexpr int YC = kTnBN / 8;
    for (int i = threadIdx.x; i < kTnBM * YC; i += kTnThreads) {
      const int rr = i / YC, cc = i - rr * YC;
      const bool ok = (m0 + rr) < m_end && (col0 + cc * 8) < p.n;
      cp_async16(sY + (buf * kTnBM + rr) * LDY + cc * 8, p.y + (ok ? (m0 + rr) : 0) * p.ldy + (ok ? (col0 + cc * 8) : 0), ok);
    }
  };

  float acc[RT][2][4];
#pragma unroll
  for (int a = 0; a < RT; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) { acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f; }

  load_chunk(0, 0);
  cp_async_commit();
  for (int c = 0; c < n_chunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < n_chunks) {
      load_chunk(buf ^ 1, c + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* cS = sS + buf * kTnBM * LDS;
    const __nv_bfloat16* cY = sY + buf * kTnBM * LDY;
#pragma unroll
    for (int ks = 0; ks < kTnBM / 16; ++ks) {
      uint32_t bfr[4];
      ldsm_x4_t(bfr, bkn_addr(cY, LDY, ks * 16, warp * 16, lane));
#pragma unroll
      for (int rt = 0; rt < RT; ++rt) {
        uint32_t a[4];
        ldsm_x4_t(a, akm_addr(cS, LDS, ks * 16, rt * 16, lane));
        mma_bf16(acc[rt][0], a, bfr[0], bfr[1]);
        mma_bf16(acc[rt][1], a, bfr[2], bfr[3]);
      }
    }
    __syncthreads();
  }

  // Epilogue.  A thread owns 4 distinct columns (nt, e & 1) and 2 RT rows; the destination column of every (group, column) pair is
  // resolved ONCE (the per-element group loop with an integer division per group cost ~5000 instructions per thread at r = 96 --
  // more than the main loop of a CTA that reduces 15 chunks: 1.3 TB/s on the fused q|k|v reduction).
  int jc[4][4];
#pragma unroll
  for (int gi = 0; gi < 4; ++gi) {
#pragma unroll
    for (int cv = 0; cv < 4; ++cv) {
      jc[gi][cv] = -1;
      if (gi < p.n_groups) {
        const TnGroup& G = p.g[gi];
        const int col = col0 + warp * 16 + (cv >> 1) * 8 + 2 * t + (cv & 1);
        const int rel = col - G.col_start;
        if (col < p.n && rel >= 0) {
          const int j = rel / G.col_stride;
          if (rel - j * G.col_stride == 0 && j < G.ncols) jc[gi][cv] = j;
        }
      }
    }
  }
#pragma unroll
  for (int gi = 0; gi < 4; ++gi) {
    if (gi >= p.n_groups) break;
    const TnGroup& G = p.g[gi];
#pragma unroll
    for (int rt = 0; rt < RT; ++rt) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int row = rt * 16 + g + hf * 8;
        if (row >= p.r || row < G.r0 || row >= G.r0 + G.rg) continue;
        float* drow = G.dst + static_cast<long long>(row - G.r0) * G.ld;
#pragma unroll
        for (int cv = 0; cv < 4; ++cv)
          if (jc[gi][cv] >= 0) atomicAdd(drow + jc[gi][cv], acc[rt][cv >> 1][hf * 2 + (cv & 1)] * p.scale);
      }
    }
  }
}

// =================================================================================================== attention backward
struct AttnBwdP {
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  __nv_bfloat16 *dq, *dk, *dv;
  long long q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs, o_ss, do_bs, do_ss, dq_bs, dq_ss, dk_bs, dk_ss, dv_bs, dv_ss;
  float* lse;             // fp32 [B, hq, sq]: natural-log logsumexp of every (masked, soft-capped) score row
  float* delta;           // fp32 [B, hq, sq]: sum_d dO * O
  int hq, hkv, sq, sk, d;
  float scale, softcap;
  int causal, prefix;
  int window;             // > 0: key j masked for query i when i + (sk - sq) - j >= window
};

constexpr int kAbThreads = 128;      // 4 warps x 16 rows

// rows [r0, r0 + ROWS) x d of a strided [s, ...] bf16 matrix -> smem [ROWS][DP + 8]; rows >= s zero-filled (cp.async zfill)
template <int DP, int ROWS>
__device__ __forceinline__ void ab_load_tile(__nv_bfloat16* sm, const __nv_bfloat16* g, long long row_stride, int r0, int s, int d) {
  constexpr int LD = DP + 8;
  const int chunks = d >> 3;
  for (int i = threadIdx.x; i < ROWS * chunks; i += kAbThreads) {
    const int r = i / chunks, c = i - r * chunks;
    const bool ok = (r0 + r) < s;
    cp_async16(sm + r * LD + c * 8, g + static_cast<long long>(ok ? (r0 + r) : 0) * row_stride + c * 8, ok);
  }
}
template <int DP>
__device__ __forceinline__ void ab_zero_pad(__nv_bfloat16* sm, int rows, int d) {
  constexpr int LD = DP + 8;
  if (d >= DP) return;
  const int padc = DP - d;
  for (int i = threadIdx.x; i < rows * padc; i += kAbThreads) sm[(i / padc) * LD + d + i % padc] = __float2bfloat16(0.f);
}

// visible(i, j): key j visible to query i
__device__ __forceinline__ bool ab_masked(const AttnBwdP& p, int i, int j) {
  return j >= p.sk || (p.causal && j > max(i + (p.sk - p.sq), p.prefix - 1)) || (p.window > 0 && i + (p.sk - p.sq) - j >= p.window);
}
// tanh for soft-capping: |u| is small (scores / 50), an odd degree-9 polynomial is exact to fp32 rounding below 0.35 (the same
// evaluation the forward kernels use, so the recomputed probabilities are consistent with the forward pass)
__device__ __forceinline__ float ab_tanh(float u) {
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
__device__ __forceinline__ float ab_exp(float x) {        // e^x through MUFU ex2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
// soft-capped score (natural units) and d(capped)/d(raw scaled score)
__device__ __forceinline__ void ab_score(const AttnBwdP& p, float raw, float& c, float& fac) {
  const float u = raw * p.scale;
  if (p.softcap > 0.f) {
    const float th = ab_tanh(u / p.softcap);
    c = p.softcap * th;
    fac = 1.f - th * th;
  } else {
    c = u;
    fac = 1.f;
  }
}
// a (query tile, key tile) pair needs no per-element mask when every key is in range and visible to every query of the tile
__device__ __forceinline__ bool ab_tile_unmasked(const AttnBwdP& p, int q_lo, int q_hi, int k_lo, int k_hi) {
  if (k_hi > p.sk || q_hi > p.sq || p.window > 0) return false;
  if (!p.causal) return true;
  return (k_hi - 1) <= max(q_lo + (p.sk - p.sq), p.prefix - 1);
}

// ---- launch 0: logsumexp per query row + delta = rowsum(dO * O)
template <int DP>
__global__ void __launch_bounds__(kAbThreads)
svla_attn_bwd_stats_kernel(const AttnBwdP p) {
  constexpr int LD = DP + 8, KT = 64;
  extern __shared__ __align__(16) uint8_t smem_ab[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_ab);       // [64][LD]
  __nv_bfloat16* sK = sQ + 64 * LD;                                    // [2][64][LD]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
  const int hk = h / (p.hq / p.hkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* qg = p.q + b * p.q_bs + static_cast<long long>(h) * p.d;
  const __nv_bfloat16* kg = p.k + b * p.k_bs + static_cast<long long>(hk) * p.d;
  ab_zero_pad<DP>(sQ, 64 * 3, p.d);
  const int n_kt = (p.sk + KT - 1) / KT;
  ab_load_tile<DP, 64>(sQ, qg, p.q_ss, q0, p.sq, p.d);
  ab_load_tile<DP, 64>(sK, kg, p.k_ss, 0, p.sk, p.d);
  cp_async_commit();
  const int qi0 = q0 + warp * 16 + g;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  for (int jt = 0; jt < n_kt; ++jt) {
    const int buf = jt & 1;
    if (jt + 1 < n_kt) {
      ab_load_tile<DP, 64>(sK + (buf ^ 1) * 64 * LD, kg, p.k_ss, (jt + 1) * KT, p.sk, p.d);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* cK = sK + buf * 64 * LD;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < DP / 16; ++kt) {
      uint32_t a[4];
      ldsm_x4(a, a_addr(sQ, LD, warp * 16, kt * 16, lane));
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bfr[4];
        ldsm_x4(bfr, bnk_addr(cK, LD, np * 16, kt * 16, lane));
        mma_bf16(s[2 * np], a, bfr[0], bfr[1]);
        mma_bf16(s[2 * np + 1], a, bfr[2], bfr[3]);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = qi0 + (e >> 1) * 8, j = jt * KT + nt * 8 + 2 * t + (e & 1);
        float c, fac;
        ab_score(p, s[nt][e], c, fac);
        s[nt][e] = ab_masked(p, i, j) ? -INFINITY : c;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      l_run[r] *= ab_exp(m_run[r] - m_use);
      m_run[r] = m_new;
      mx[r] = m_use;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      l_run[0] += ab_exp(s[nt][0] - mx[0]) + ab_exp(s[nt][1] - mx[0]);
      l_run[1] += ab_exp(s[nt][2] - mx[1]) + ab_exp(s[nt][3] - mx[1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    const int i = qi0 + 8 * r;
    if (t == 0 && i < p.sq) p.lse[(static_cast<long long>(b) * p.hq + h) * p.sq + i] = m_run[r] + logf(l_run[r]);
  }
  // delta: 16 rows per warp, lanes over d
  const __nv_bfloat16* og = p.o + b * p.o_bs + static_cast<long long>(h) * p.d;
  const __nv_bfloat16* dg = p.dout + b * p.do_bs + static_cast<long long>(h) * p.d;
  for (int r = 0; r < 16; ++r) {
    const int i = q0 + warp * 16 + r;
    if (i >= p.sq) break;
    float acc = 0.f;
    for (int c = lane * 2; c < p.d; c += 64) {
      const __nv_bfloat162 ov = *reinterpret_cast<const __nv_bfloat162*>(og + static_cast<long long>(i) * p.o_ss + c);
      const __nv_bfloat162 dv = *reinterpret_cast<const __nv_bfloat162*>(dg + static_cast<long long>(i) * p.do_ss + c);
      acc += __bfloat162float(ov.x) * __bfloat162float(dv.x) + __bfloat162float(ov.y) * __bfloat162float(dv.y);
    }
    acc = warp_sum(acc);
    if (lane == 0) p.delta[(static_cast<long long>(b) * p.hq + h) * p.sq + i] = acc;
  }
}

// ---- launch 1: dQ.  CTA = 64 queries of one (batch, head); key tiles of KT keys stream through shared memory.
template <int DP, int KT, int NB>
__global__ void __launch_bounds__(kAbThreads, NB == 1 ? 2 : 1)
svla_attn_bwd_dq_kernel(const AttnBwdP p) {
  // NB = K / V buffers: 2 = double-buffered cp.async pipeline; 1 = single buffer so that TWO CTAs fit one SM at d = 256 (101 KB
  // each): the second CTA's MMAs hide this CTA's loads, which measured faster than one CTA with a prefetch pipeline
  constexpr int LD = DP + 8, NTK = KT / 8, NTD = DP / 8;
  extern __shared__ __align__(16) uint8_t smem_ab[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_ab);       // [64][LD]
  __nv_bfloat16* sdO = sQ + 64 * LD;                                   // [64][LD]
  __nv_bfloat16* sK = sdO + 64 * LD;                                   // [NB][KT][LD]
  __nv_bfloat16* sV = sK + NB * KT * LD;                               // [NB][KT][LD]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
  const int hk = h / (p.hq / p.hkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* qg = p.q + b * p.q_bs + static_cast<long long>(h) * p.d;
  const __nv_bfloat16* dog = p.dout + b * p.do_bs + static_cast<long long>(h) * p.d;
  const __nv_bfloat16* kg = p.k + b * p.k_bs + static_cast<long long>(hk) * p.d;
  const __nv_bfloat16* vg = p.v + b * p.v_bs + static_cast<long long>(hk) * p.d;
  ab_zero_pad<DP>(sQ, 128 + 2 * NB * KT, p.d);
  // keys beyond the last one any query of this tile can see are skipped
  int k_hi = p.sk;
  if (p.causal) k_hi = min(p.sk, max(min(q0 + 63, p.sq - 1) + (p.sk - p.sq), p.prefix - 1) + 1);
  const int n_kt = (k_hi + KT - 1) / KT;
  ab_load_tile<DP, 64>(sQ, qg, p.q_ss, q0, p.sq, p.d);
  ab_load_tile<DP, 64>(sdO, dog, p.do_ss, q0, p.sq, p.d);
  ab_load_tile<DP, KT>(sK, kg, p.k_ss, 0, p.sk, p.d);
  ab_load_tile<DP, KT>(sV, vg, p.v_ss, 0, p.sk, p.d);
  cp_async_commit();
  const int qi0 = q0 + warp * 16 + g;
  float lse_r[2], del_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = min(qi0 + 8 * r, p.sq - 1);
    lse_r[r] = p.lse[(static_cast<long long>(b) * p.hq + h) * p.sq + i];
    del_r[r] = p.delta[(static_cast<long long>(b) * p.hq + h) * p.sq + i];
  }
  float dq[NTD][4];
#pragma unroll
  for (int i = 0; i < NTD; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

  for (int jt = 0; jt < n_kt; ++jt) {
    const int buf = (NB == 2) ? (jt & 1) : 0;
    if (NB == 2 && jt + 1 < n_kt) {
      ab_load_tile<DP, KT>(sK + (buf ^ 1) * KT * LD, kg, p.k_ss, (jt + 1) * KT, p.sk, p.d);
      ab_load_tile<DP, KT>(sV + (buf ^ 1) * KT * LD, vg, p.v_ss, (jt + 1) * KT, p.sk, p.d);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* cK = sK + buf * KT * LD;
    const __nv_bfloat16* cV = sV + buf * KT * LD;
    const bool plain = ab_tile_unmasked(p, q0, q0 + 64, jt * KT, (jt + 1) * KT);
    float s[NTK][4], dp[NTK][4];
#pragma unroll
    for (int i = 0; i < NTK; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < DP / 16; ++kt) {
      uint32_t a[4], ad[4];
      ldsm_x4(a, a_addr(sQ, LD, warp * 16, kt * 16, lane));
      ldsm_x4(ad, a_addr(sdO, LD, warp * 16, kt * 16, lane));
#pragma unroll
      for (int np = 0; np < KT / 16; ++np) {
        uint32_t bfr[4];
        ldsm_x4(bfr, bnk_addr(cK, LD, np * 16, kt * 16, lane));
        mma_bf16(s[2 * np], a, bfr[0], bfr[1]);
        mma_bf16(s[2 * np + 1], a, bfr[2], bfr[3]);
        ldsm_x4(bfr, bnk_addr(cV, LD, np * 16, kt * 16, lane));
        mma_bf16(dp[2 * np], ad, bfr[0], bfr[1]);
        mma_bf16(dp[2 * np + 1], ad, bfr[2], bfr[3]);
      }
    }
    uint32_t dsa[NTK][2];
#pragma unroll
    for (int nt = 0; nt < NTK; ++nt) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = qi0 + (e >> 1) * 8, j = jt * KT + nt * 8 + 2 * t + (e & 1);
        float c, fac;
        ab_score(p, s[nt][e], c, fac);
        const float pr = (!plain && (ab_masked(p, i, j) || i >= p.sq)) ? 0.f : ab_exp(c - lse_r[e >> 1]);
        ds[e] = pr * (dp[nt][e] - del_r[e >> 1]) * fac * p.scale;
      }
      dsa[nt][0] = pack_bf16x2(ds[0], ds[1]);
      dsa[nt][1] = pack_bf16x2(ds[2], ds[3]);
    }
#pragma unroll
    for (int ks = 0; ks < KT / 16; ++ks) {
      const uint32_t a[4] = {dsa[2 * ks][0], dsa[2 * ks][1], dsa[2 * ks + 1][0], dsa[2 * ks + 1][1]};
#pragma unroll
      for (int dpi = 0; dpi < NTD / 2; ++dpi) {
        uint32_t bfr[4];
        ldsm_x4_t(bfr, bkn_addr(cK, LD, ks * 16, dpi * 16, lane));
        mma_bf16(dq[2 * dpi], a, bfr[0], bfr[1]);
        mma_bf16(dq[2 * dpi + 1], a, bfr[2], bfr[3]);
      }
    }
    __syncthreads();
    if (NB == 1 && jt + 1 < n_kt) {
      ab_load_tile<DP, KT>(sK, kg, p.k_ss, (jt + 1) * KT, p.sk, p.d);
      ab_load_tile<DP, KT>(sV, vg, p.v_ss, (jt + 1) * KT, p.sk, p.d);
      cp_async_commit();
    }
  }
  __nv_bfloat16* dqg = p.dq + b * p.dq_bs + static_cast<long long>(h) * p.d;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = qi0 + 8 * r;
    if (i >= p.sq) continue;
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (c < p.d) *reinterpret_cast<uint32_t*>(dqg + static_cast<long long>(i) * p.dq_ss + c) = pack_bf16x2(dq[nt][2 * r], dq[nt][2 * r + 1]);
    }
  }
}

// ---- launch 2: dK, dV.  CTA = 64 keys of one (batch, kv head); query tiles of QT queries of every head of the GQA group stream
// through shared memory.  The head dimension is processed in slices of DH columns (one sweep over the queries per slice) so that
// the dK / dV accumulators of a slice fit the register file at d = 256; the score and dP tiles are recomputed per slice.
template <int DP, int QT, int DH, int NB>
__global__ void __launch_bounds__(kAbThreads, NB == 1 ? 2 : 1)
svla_attn_bwd_dkv_kernel(const AttnBwdP p) {
  constexpr int LD = DP + 8, NTQ = QT / 8, NTH = DH / 8;
  extern __shared__ __align__(16) uint8_t smem_ab[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_ab);       // [64][LD]
  __nv_bfloat16* sV = sK + 64 * LD;                                    // [64][LD]
  __nv_bfloat16* sQ = sV + 64 * LD;                                    // [NB][QT][LD]
  __nv_bfloat16* sdO = sQ + NB * QT * LD;                              // [NB][QT][LD]
  float* sLse = reinterpret_cast<float*>(sdO + NB * QT * LD);          // [NB][QT]
  float* sDel = sLse + NB * QT;                                        // [NB][QT]
  const int b = blockIdx.z, hk = blockIdx.y, k0 = blockIdx.x * 64;
  const int G = p.hq / p.hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* kg = p.k + b * p.k_bs + static_cast<long long>(hk) * p.d;
  const __nv_bfloat16* vg = p.v + b * p.v_bs + static_cast<long long>(hk) * p.d;
  ab_zero_pad<DP>(sK, 128 + 2 * NB * QT, p.d);
  // first query that can see a key of this tile
  int q_lo = 0;
  if (p.causal && k0 >= p.prefix) q_lo = max(0, k0 - (p.sk - p.sq));
  const int qt_lo = q_lo / QT, n_qt = (p.sq + QT - 1) / QT;
  const int tiles_per_head = n_qt - qt_lo;
  const int n_it = tiles_per_head > 0 ? tiles_per_head * G : 0;
  const int kj0 = k0 + warp * 16 + g;          // this thread's key rows: kj0, kj0 + 8

  auto load_q = [&](int buf, int it) {
    const int hh = hk * G + it / tiles_per_head, qt = qt_lo + it % tiles_per_head;
    const __nv_bfloat16* qg = p.q + b * p.q_bs + static_cast<long long>(hh) * p.d;
    const __nv_bfloat16* dog = p.dout + b * p.do_bs + static_cast<long long>(hh) * p.d;
    ab_load_tile<DP, QT>(sQ + buf * QT * LD, qg, p.q_ss, qt * QT, p.sq, p.d);
    ab_load_tile<DP, QT>(sdO + buf * QT * LD, dog, p.do_ss, qt * QT, p.sq, p.d);
    if (threadIdx.x < QT) {
      const int i = qt * QT + threadIdx.x;
      const long long o = (static_cast<long long>(b) * p.hq + hh) * p.sq + min(i, p.sq - 1);
      sLse[buf * QT + threadIdx.x] = p.lse[o];
      sDel[buf * QT + threadIdx.x] = p.delta[o];
    }
  };

  ab_load_tile<DP, 64>(sK, kg, p.k_ss, k0, p.sk, p.d);
  ab_load_tile<DP, 64>(sV, vg, p.v_ss, k0, p.sk, p.d);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  __nv_bfloat16* dkg = p.dk + b * p.dk_bs + static_cast<long long>(hk) * p.d;
  __nv_bfloat16* dvg = p.dv + b * p.dv_bs + static_cast<long long>(hk) * p.d;
  for (int dh = 0; dh * DH < p.d; ++dh) {
    float dk[NTH][4], dv[NTH][4];
#pragma unroll
    for (int i = 0; i < NTH; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    if (n_it > 0) {
      load_q(0, 0);
      cp_async_commit();
    }
    for (int it = 0; it < n_it; ++it) {
      const int buf = (NB == 2) ? (it & 1) : 0;
      if (NB == 2 && it + 1 < n_it) {
        load_q(buf ^ 1, it + 1);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const int qt = qt_lo + it % tiles_per_head;
      const bool plain = ab_tile_unmasked(p, qt * QT, (qt + 1) * QT, k0, k0 + 64);
      const __nv_bfloat16* cQ = sQ + buf * QT * LD;
      const __nv_bfloat16* cdO = sdO + buf * QT * LD;
      // S^T = K Q^T and dP^T = V dO^T  (16 keys x QT queries per warp)
      float st[NTQ][4], dpt[NTQ][4];
#pragma unroll
      for (int i = 0; i < NTQ; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
#pragma unroll
      for (int kt = 0; kt < DP / 16; ++kt) {
        uint32_t ak[4], av[4];
        ldsm_x4(ak, a_addr(sK, LD, warp * 16, kt * 16, lane));
        ldsm_x4(av, a_addr(sV, LD, warp * 16, kt * 16, lane));
#pragma unroll
        for (int np = 0; np < QT / 16; ++np) {
          uint32_t bfr[4];
          ldsm_x4(bfr, bnk_addr(cQ, LD, np * 16, kt * 16, lane));
          mma_bf16(st[2 * np], ak, bfr[0], bfr[1]);
          mma_bf16(st[2 * np + 1], ak, bfr[2], bfr[3]);
          ldsm_x4(bfr, bnk_addr(cdO, LD, np * 16, kt * 16, lane));
          mma_bf16(dpt[2 * np], av, bfr[0], bfr[1]);
          mma_bf16(dpt[2 * np + 1], av, bfr[2], bfr[3]);
        }
      }
      uint32_t pa[NTQ][2], dsa[NTQ][2];
#pragma unroll
      for (int nt = 0; nt < NTQ; ++nt) {
        float pv[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = kj0 + (e >> 1) * 8;                       // key
          const int il = nt * 8 + 2 * t + (e & 1), i = qt * QT + il;   // query
          float c, fac;
          ab_score(p, st[nt][e], c, fac);
          const float pr = (!plain && (ab_masked(p, i, j) || i >= p.sq)) ? 0.f : ab_exp(c - sLse[buf * QT + il]);
          pv[e] = pr;
          ds[e] = pr * (dpt[nt][e] - sDel[buf * QT + il]) * fac * p.scale;
        }
        pa[nt][0] = pack_bf16x2(pv[0], pv[1]); pa[nt][1] = pack_bf16x2(pv[2], pv[3]);
        dsa[nt][0] = pack_bf16x2(ds[0], ds[1]); dsa[nt][1] = pack_bf16x2(ds[2], ds[3]);
      }
      // dV += P^T dO,  dK += dS^T Q   (contraction over the QT queries), columns [dh * DH, dh * DH + DH)
#pragma unroll
      for (int ks = 0; ks < QT / 16; ++ks) {
        const uint32_t ap[4] = {pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1]};
        const uint32_t as[4] = {dsa[2 * ks][0], dsa[2 * ks][1], dsa[2 * ks + 1][0], dsa[2 * ks + 1][1]};
#pragma unroll
        for (int dpi = 0; dpi < NTH / 2; ++dpi) {
          uint32_t bfr[4];
          ldsm_x4_t(bfr, bkn_addr(cdO, LD, ks * 16, dh * DH + dpi * 16, lane));
          mma_bf16(dv[2 * dpi], ap, bfr[0], bfr[1]);
          mma_bf16(dv[2 * dpi + 1], ap, bfr[2], bfr[3]);
          ldsm_x4_t(bfr, bkn_addr(cQ, LD, ks * 16, dh * DH + dpi * 16, lane));
          mma_bf16(dk[2 * dpi], as, bfr[0], bfr[1]);
          mma_bf16(dk[2 * dpi + 1], as, bfr[2], bfr[3]);
        }
      }
      __syncthreads();
      if (NB == 1 && it + 1 < n_it) {
        load_q(0, it + 1);
        cp_async_commit();
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = kj0 + 8 * r;
      if (j >= p.sk) continue;
#pragma unroll
      for (int nt = 0; nt < NTH; ++nt) {
        const int c = dh * DH + nt * 8 + 2 * t;
        if (c < p.d) {
          *reinterpret_cast<uint32_t*>(dkg + static_cast<long long>(j) * p.dk_ss + c) = pack_bf16x2(dk[nt][2 * r], dk[nt][2 * r + 1]);
          *reinterpret_cast<uint32_t*>(dvg + static_cast<long long>(j) * p.dv_ss + c) = pack_bf16x2(dv[nt][2 * r], dv[nt][2 * r + 1]);
        }
      }
    }
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    svla_set_error("%s: cudaFuncSetAttribute(%zu bytes) failed: %s", what, bytes, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

template <int DP, int KT, int QT, int DH, int NB>
int launch_attn_bwd(const AttnBwdP& p, int batch, cudaStream_t st) {
  constexpr int LD = DP + 8;
  const size_t sm0 = static_cast<size_t>(3 * 64 * LD) * 2;
  const size_t sm1 = static_cast<size_t>(128 + 2 * NB * KT) * LD * 2;
  const size_t sm2 = static_cast<size_t>(128 + 2 * NB * QT) * LD * 2 + 2 * NB * QT * sizeof(float);
  static bool configured = false;
  if (!configured) {
    if (set_smem(svla_attn_bwd_stats_kernel<DP>, sm0, "svla_attention_bwd(stats)")) return -2;
    if (set_smem(svla_attn_bwd_dq_kernel<DP, KT, NB>, sm1, "svla_attention_bwd(dq)")) return -2;
    if (set_smem(svla_attn_bwd_dkv_kernel<DP, QT, DH, NB>, sm2, "svla_attention_bwd(dkv)")) return -2;
    configured = true;
  }
  const dim3 gq((p.sq + 63) / 64, p.hq, batch), gk((p.sk + 63) / 64, p.hkv, batch);
  svla_attn_bwd_stats_kernel<DP><<<gq, kAbThreads, sm0, st>>>(p);
  SVLA_LAUNCH_CHECK("svla_attn_bwd_stats");
  svla_attn_bwd_dq_kernel<DP, KT, NB><<<gq, kAbThreads, sm1, st>>>(p);
  SVLA_LAUNCH_CHECK("svla_attn_bwd_dq");
  svla_attn_bwd_dkv_kernel<DP, QT, DH, NB><<<gk, kAbThreads, sm2, st>>>(p);
  SVLA_LAUNCH_CHECK("svla_attn_bwd_dkv");
  return 0;
}

}  // namespace

int svla_attention_bwd_tc_try(const SvlaAttnBwdArgs* a, void* stream);      // attention_bwd_tc.cu (tcgen05 / TMEM / TMA)

extern "C" int svla_attention_bwd(const SvlaAttnBwdArgs* a, void* stream) {
  SVLA_REQUIRE(a && a->q && a->k && a->v && a->out && a->dout && a->dq && a->dk && a->dv && a->lse && a->delta, "svla_attention_bwd: null pointer");
  {
    // SVLA_ATTN_IMPL=mma forces the warp-MMA kernels (A/B measurements); they are also the fallback for shapes the tcgen05 sweeps do
    // not cover and for callers without the forward pass's log-sum-exp
    static const bool force_mma = getenv("SVLA_ATTN_IMPL") && strcmp(getenv("SVLA_ATTN_IMPL"), "mma") == 0;
    if (!force_mma && a->fwd_lse2) {
      const int rc = svla_attention_bwd_tc_try(a, stream);
      if (rc != 1) return rc;
    }
  }
  SVLA_REQUIRE(a->batch > 0 && a->hq > 0 && a->hkv > 0 && a->hq % a->hkv == 0 && a->sq > 0 && a->sk > 0, "svla_attention_bwd: bad shape");
  SVLA_REQUIRE(a->d % 8 == 0 && a->d <= 256, "svla_attention_bwd: head dim %d unsupported (multiple of 8, <= 256)", a->d);
  SVLA_REQUIRE(a->batch <= 65535 && a->hq <= 65535, "svla_attention_bwd: grid limits");
  SVLA_REQUIRE(a->causal_prefix == 0 || a->causal, "svla_attention_bwd: causal_prefix needs causal = 1");
  AttnBwdP p;
  p.q = static_cast<const __nv_bfloat16*>(a->q); p.k = static_cast<const __nv_bfloat16*>(a->k); p.v = static_cast<const __nv_bfloat16*>(a->v);
  p.o = static_cast<const __nv_bfloat16*>(a->out); p.dout = static_cast<const __nv_bfloat16*>(a->dout);
  p.dq = static_cast<__nv_bfloat16*>(a->dq); p.dk = static_cast<__nv_bfloat16*>(a->dk); p.dv = static_cast<__nv_bfloat16*>(a->dv);
  p.q_bs = a->q_bs; p.q_ss = a->q_ss; p.k_bs = a->k_bs; p.k_ss = a->k_ss; p.v_bs = a->v_bs; p.v_ss = a->v_ss;
  p.o_bs = a->o_bs; p.o_ss = a->o_ss; p.do_bs = a->do_bs; p.do_ss = a->do_ss;
  p.dq_bs = a->dq_bs; p.dq_ss = a->dq_ss; p.dk_bs = a->dk_bs; p.dk_ss = a->dk_ss; p.dv_bs = a->dv_bs; p.dv_ss = a->dv_ss;
  p.lse = a->lse; p.delta = a->delta;
  p.hq = a->hq; p.hkv = a->hkv; p.sq = a->sq; p.sk = a->sk; p.d = a->d;
  p.scale = a->scale; p.softcap = a->softcap; p.causal = a->causal; p.prefix = a->causal_prefix; p.window = a->window;
  const long long strides[] = {p.q_bs, p.q_ss, p.k_bs, p.k_ss, p.v_bs, p.v_ss, p.do_bs, p.do_ss, p.o_bs, p.o_ss, p.dq_bs, p.dq_ss,
                               p.dk_bs, p.dk_ss, p.dv_bs, p.dv_ss};
  for (long long s : strides) SVLA_REQUIRE(s % 8 == 0, "svla_attention_bwd: strides must be multiples of 8 elements (16-byte rows)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const int nb_env = getenv("SVLA_ATTN_BWD_NB") ? atoi(getenv("SVLA_ATTN_BWD_NB")) : 0;      // A/B switch: 1 | 2 buffers at d = 256
  if (a->d <= 80) return launch_attn_bwd<80, 64, 64, 80, 2>(p, a->batch, st);
  if (a->d <= 128) return launch_attn_bwd<128, 64, 32, 128, 2>(p, a->batch, st);
  if (nb_env == 2) return launch_attn_bwd<256, 32, 32, 128, 2>(p, a->batch, st);
  return launch_attn_bwd<256, 32, 32, 128, 1>(p, a->batch, st);
}

extern "C" int svla_gemm_tn(const SvlaGemmTnArgs* a, void* stream) {
  SVLA_REQUIRE(a && a->s && a->y && a->m > 0 && a->r > 0 && a->n > 0, "svla_gemm_tn: bad arguments");
  SVLA_REQUIRE(a->r % 16 == 0 && a->r <= 128, "svla_gemm_tn: r=%d must be a multiple of 16, <= 128", a->r);
  SVLA_REQUIRE(a->lds % 8 == 0 && a->ldy % 8 == 0 && a->lds >= a->r && a->n % 8 == 0 && a->ldy >= a->n, "svla_gemm_tn: lds / ldy / n must be multiples of 8");
  SVLA_REQUIRE((reinterpret_cast<uintptr_t>(a->s) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->y) & 15) == 0, "svla_gemm_tn: operands must be 16-byte aligned");
  SVLA_REQUIRE(a->n_groups >= 1 && a->n_groups <= 4, "svla_gemm_tn: 1..4 output groups");
  TnParams p;
  p.s = static_cast<const __nv_bfloat16*>(a->s); p.y = static_cast<const __nv_bfloat16*>(a->y);
  p.lds = a->lds; p.ldy = a->ldy; p.m = a->m; p.r = a->r; p.n = a->n; p.scale = a->scale; p.n_groups = a->n_groups;
  for (int i = 0; i < a->n_groups; ++i) {
    const SvlaTnGroup& G = a->groups[i];
    SVLA_REQUIRE(G.dst && G.rows > 0 && G.row0 >= 0 && G.row0 + G.rows <= a->r && G.col_stride >= 1 && G.ncols > 0 && G.col_start >= 0 &&
                 G.col_start + static_cast<long long>(G.ncols - 1) * G.col_stride < a->n && G.ld >= G.ncols, "svla_gemm_tn: bad group %d", i);
    p.g[i].dst = G.dst; p.g[i].ld = G.ld; p.g[i].r0 = G.row0; p.g[i].rg = G.rows;
    p.g[i].col_start = G.col_start; p.g[i].col_stride = G.col_stride; p.g[i].ncols = G.ncols;
  }
  const int col_tiles = (a->n + kTnBN - 1) / kTnBN;
  const long long chunks = (a->m + kTnBM - 1) / kTnBM;
  long long splits = (2LL * svla_num_sms() + col_tiles - 1) / col_tiles;
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const long long chunks_per = (chunks + splits - 1) / splits;
  p.m_per_split = chunks_per * kTnBM;
  splits = (chunks + chunks_per - 1) / chunks_per;
  const dim3 grid(col_tiles, static_cast<unsigned>(splits));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rt = a->r / 16;
  auto smem_for = [](int R) { return static_cast<size_t>(2 * kTnBM) * ((R + 8) + (kTnBN + 8)) * 2; };
#define SVLA_TN_CASE(RT)                                                                                             \
  case RT: {                                                                                                         \
    static bool cfgd = false;                                                                                        \
    if (!cfgd) { if (set_smem(svla_gemm_tn_kernel<RT>, smem_for(16 * RT), "svla_gemm_tn")) return -2; cfgd = true; } \
    svla_gemm_tn_kernel<RT><<<grid, kTnThreads, smem_for(16 * RT), st>>>(p);                                         \
    break;                                                                                                           \
  }
  switch (rt) {
    SVLA_TN_CASE(1) SVLA_TN_CASE(2) SVLA_TN_CASE(3) SVLA_TN_CASE(4) SVLA_TN_CASE(5) SVLA_TN_CASE(6) SVLA_TN_CASE(7) SVLA_TN_CASE(8)
    default: SVLA_REQUIRE(false, "svla_gemm_tn: r=%d unsupported", a->r);
  }
#undef SVLA_TN_CASE
  SVLA_LAUNCH_CHECK("svla_gemm_tn");
  return 0;
}
