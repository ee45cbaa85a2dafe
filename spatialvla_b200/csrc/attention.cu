// G2 / G3: attention kernels.
//
// G2 (svla_attention): flash-style fused softmax(scale*QK^T [softcap] [+relpos bias] [mask]) V.  One CTA = 64
// query rows of one (batch, head); 4 warps x 16 rows; K/V tiles of 64 keys double-buffered with cp.async;
// bf16 mma.sync.m16n8k16 with fp32 accumulation, fp32 online softmax (exp2).  Sequences on this path are
// short (256 / 577 / 278 / 145 tokens), attention is ~3% of the FLOPs of an observation; the tcgen05 port of
// this kernel is scheduled after the GEMM (DESIGN.md "next").  BEiT's relative-position bias is looked up on
// the fly from the per-layer (2w-1)^2+3 table (one head column cached in shared memory) instead of
// materialising a [heads, 577, 577] tensor.
//
// G3 (svla_decode_attention): q_len = 1 over the KV cache, one CTA per (batch, kv head) serving the whole GQA
// group, HBM-bound (reads each K/V row once, 16-byte loads).
#include <cstdlib>
#include <cstring>
#include "svla_common.cuh"
#include "decode_attn_item.cuh"

namespace {

constexpr int kBQ = 128, kBKV = 64, kAttnThreads = 256;   // 8 warps x 16 query rows share each K/V tile

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tanh for soft-capping: |u| is small (scores / 50), so an odd degree-9 Taylor polynomial is exact to fp32
// rounding below 0.35 and the libm path handles the rare large argument.
__device__ __forceinline__ float tanh_small(float u) {
  const float u2 = u * u;
  if (u2 < 0.1225f) {
    float p = 62.f / 2835.f;
    p = fmaf(p, u2, -17.f / 315.f);
    p = fmaf(p, u2, 2.f / 15.f);
    p = fmaf(p, u2, -1.f / 3.f);
    p = fmaf(p, u2, 1.f);
    return u * p;
  }
  return tanhf(u);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnP {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* out;
  long long q_bs, q_ss, k_bs, k_ss, v_bs, v_ss, o_bs, o_ss;
  int hq, hkv, sq, sk, d;
  float scale, softcap;
  int causal;
  const float* relpos;
  int head_major;
  const int* kv_start;
  int win;
  int prefix;          // causal only: keys < prefix are visible to every query (prefix-LM)
  int window;          // > 0: key slot j is masked for query slot i when i - j >= window (sliding-window layers)
};

// Loads rows [r0, r0+ROWS) x d (bf16) of a [s, ...] strided matrix into smem [ROWS][DP+8]; rows >= s are zeroed.
template <int DP, int ROWS>
__device__ __forceinline__ void load_tile(__nv_bfloat16* sm, const __nv_bfloat16* g, long long row_stride, int r0, int s, int d) {
  constexpr int LD = DP + 8;
  const int chunks = d >> 3;
  for (int i = threadIdx.x; i < ROWS * chunks; i += kAttnThreads) {
    const int r = i / chunks, c = i - r * chunks;
    const bool ok = (r0 + r) < s;
    const __nv_bfloat16* src = g + static_cast<long long>(ok ? (r0 + r) : 0) * row_stride + c * 8;
    cp_async16(sm + r * LD + c * 8, src, ok);
  }
}

// MODE: bit0 = BEiT relative-position bias, bit1 = tanh soft-capping, bit2 = causal mask; 8 = decide at run time.
// The specialised instances let the compiler drop the per-score feature checks (the score path, not the MMAs, bounds
// this kernel at these short sequence lengths).
template <int DP, int MODE>
__global__ void __launch_bounds__(kAttnThreads)
svla_flash_attn_kernel(const AttnP p) {
  constexpr bool DYN = (MODE & 8) != 0;
  const bool f_relpos = DYN ? (p.relpos != nullptr) : ((MODE & 1) != 0);
  const bool f_softcap = DYN ? (p.softcap > 0.f) : ((MODE & 2) != 0);
  const bool f_causal = DYN ? (p.causal != 0) : ((MODE & 4) != 0);
  constexpr int LD = DP + 8;
  constexpr int KT = DP / 16;     // k-steps of QK^T
  constexpr int NT = DP / 8;      // n-tiles of the output
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + kBQ * LD;           // 2 buffers
  __nv_bfloat16* sV = sK + 2 * 64 * LD;        // 2 buffers
  float* sTab = reinterpret_cast<float*>(sV + 2 * 64 * LD);
  int* sKterm = reinterpret_cast<int*>(sTab + (f_relpos ? (2 * p.win - 1) * (2 * p.win - 1) + 3 : 0));   // [64]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kBQ;
  const int hk = h / (p.hq / p.hkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

  const __nv_bfloat16* qg = p.q + b * p.q_bs + static_cast<long long>(h) * p.d;
  const __nv_bfloat16* kg = p.k + b * p.k_bs + static_cast<long long>(hk) * p.d;
  const __nv_bfloat16* vg = p.v + b * p.v_bs + static_cast<long long>(hk) * p.d;

  // zero the padding columns [d, DP) once (cp.async never touches them)
  if (p.d < DP) {
    const int padc = DP - p.d;
    for (int i = threadIdx.x; i < (kBQ + 4 * 64) * padc; i += kAttnThreads) {
      const int r = i / padc, c = p.d + i % padc;
      sQ[r * LD + c] = __float2bfloat16(0.f);      // sQ, sK[2], sV[2] are contiguous
    }
  }
  int nrel = 0;
  if (f_relpos) {
    nrel = (2 * p.win - 1) * (2 * p.win - 1) + 3;
    for (int i = threadIdx.x; i < nrel; i += kAttnThreads) sTab[i] = (p.head_major ? p.relpos[static_cast<long long>(h) * nrel + i] : p.relpos[static_cast<long long>(i) * p.hq + h]) * 1.4426950408889634f;
  }

  const int n_kv_tiles = (p.sk + kBKV - 1) / kBKV;
  load_tile<DP, kBQ>(sQ, qg, p.q_ss, q0, p.sq, p.d);
  load_tile<DP, 64>(sK, kg, p.k_ss, 0, p.sk, p.d);
  load_tile<DP, 64>(sV, vg, p.v_ss, 0, p.sk, p.d);
  cp_async_commit();

  float o[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int qi0 = q0 + warp * 16 + g;      // this thread's rows: qi0 and qi0 + 8
  const bool warp_active = (q0 + warp * 16) < p.sq;
  const int causal_off = p.sk - p.sq;
  constexpr float kLog2e = 1.4426950408889634f;
  // BEiT relative-position index = qbase(query) - kterm(key) for patch tokens; CLS row/column are special-cased
  int qbase[2] = {0, 0};
  const int w2 = 2 * p.win - 1;
  if (f_relpos) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int qi = min(qi0 + 8 * r, p.sq - 1);      // rows past the end reuse the last valid row's (in-range) index
      if (qi >= 1) qbase[r] = ((qi - 1) / p.win + p.win - 1) * w2 + (qi - 1) % p.win + p.win - 1;
    }
  }
  const float inv_cap = f_softcap ? 1.f / p.softcap : 0.f;
  const float sl2 = p.scale * kLog2e;

  for (int jt = 0; jt < n_kv_tiles; ++jt) {
    const int buf = jt & 1;
    if (jt + 1 < n_kv_tiles) {
      load_tile<DP, 64>(sK + (buf ^ 1) * 64 * LD, kg, p.k_ss, (jt + 1) * kBKV, p.sk, p.d);
      load_tile<DP, 64>(sV + (buf ^ 1) * 64 * LD, vg, p.v_ss, (jt + 1) * kBKV, p.sk, p.d);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    if (f_relpos && threadIdx.x < 64) {
      const int kj = jt * kBKV + threadIdx.x;
      sKterm[threadIdx.x] = (kj >= 1 && kj < p.sk) ? ((kj - 1) / p.win) * w2 + (kj - 1) % p.win : 0;
    }
    __syncthreads();
    const __nv_bfloat16* cK = sK + buf * 64 * LD;
    const __nv_bfloat16* cV = sV + buf * 64 * LD;

    // ragged tails: warps whose 16 query rows are all >= sq only help with loads/barriers; the last K/V tile only
    // multiplies the 16-key groups that contain valid keys (BEiT: 577 = 9*64 + 1 keys)
    const int np_valid = min(4, (p.sk - jt * kBKV + 15) >> 4);
    if (!warp_active) { __syncthreads(); continue; }
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t a[4];
      ldsm_x4(a, sQ + (warp * 16 + (lane & 15)) * LD + kt * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int np = 0; np < 4; ++np) {       // pairs of key n-tiles
        if (np >= np_valid) break;
        uint32_t bfr[4];
        const int mi = lane >> 3;
        ldsm_x4(bfr, cK + (np * 16 + (mi >> 1) * 8 + (lane & 7)) * LD + kt * 16 + (mi & 1) * 8);
        mma_bf16(s[2 * np], a, bfr[0], bfr[1]);
        mma_bf16(s[2 * np + 1], a, bfr[2], bfr[3]);
      }
    }
    // ---- scores -> log2 domain (y = x * log2 e): scale, soft-cap, rel-pos bias, mask; running max.
    // The score path, not the MMAs, bounds this kernel (ncu: 32 instructions per score element in the first
    // version), so every feature is folded into as few per-element operations as possible and the special cases
    // (CLS row/column of BEiT, ragged/causal mask, large soft-cap arguments) are warp-uniform branches.
    const int kstart = p.kv_start ? p.kv_start[b] : 0;
    const bool need_mask = f_causal || (jt + 1) * kBKV > p.sk || kstart > jt * kBKV || p.window > 0;
    float mx[2] = {-INFINITY, -INFINITY};
    if (f_relpos) {
      const bool has_cls = (jt == 0) || (q0 + warp * 16 == 0);      // only then a CLS key / query is in this block
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int kt0 = sKterm[nt * 8 + 2 * t], kt1 = sKterm[nt * 8 + 2 * t + 1];
        int i00 = qbase[0] - kt0, i01 = qbase[0] - kt1, i10 = qbase[1] - kt0, i11 = qbase[1] - kt1;
        if (has_cls) {
          const int kj0 = jt * kBKV + nt * 8 + 2 * t;
          if (kj0 == 0) { i00 = nrel - 2; i10 = nrel - 2; }
          if (qi0 == 0) { i00 = (kj0 == 0) ? nrel - 1 : nrel - 3; i01 = nrel - 3; }
        }
        s[nt][0] = fmaf(s[nt][0], sl2, sTab[i00]);
        s[nt][1] = fmaf(s[nt][1], sl2, sTab[i01]);
        s[nt][2] = fmaf(s[nt][2], sl2, sTab[i10]);
        s[nt][3] = fmaf(s[nt][3], sl2, sTab[i11]);
      }
    } else if (f_softcap) {
      const float c1 = p.scale * inv_cap, c2 = p.softcap * kLog2e;
      float u2max = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float u = s[nt][e] * c1; u2max = fmaxf(u2max, u * u); }
      }
      if (!__any_sync(0xffffffffu, u2max >= 0.1225f)) {
        // cap * tanh(u) * log2e with a degree-9 odd polynomial (exact to fp32 rounding for |u| < 0.35)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float u = s[nt][e] * c1, u2 = u * u;
            float pl = 62.f / 2835.f;
            pl = fmaf(pl, u2, -17.f / 315.f);
            pl = fmaf(pl, u2, 2.f / 15.f);
            pl = fmaf(pl, u2, -1.f / 3.f);
            pl = fmaf(pl, u2, 1.f);
            s[nt][e] = c2 * u * pl;
          }
        }
      } else {                                  // rare: a large score somewhere in this warp's block -> libm tanh
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) s[nt][e] = c2 * tanhf(s[nt][e] * c1);
        }
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] *= sl2;
      }
    }
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = qi0 + (e >> 1) * 8;
          const int kj = jt * kBKV + nt * 8 + 2 * t + (e & 1);
          const bool masked = (kj >= p.sk) || (kj < kstart) || (f_causal && kj > max(qi + causal_off, p.prefix - 1)) ||
                              (p.window > 0 && qi + causal_off - kj >= p.window);
          s[nt][e] = masked ? -INFINITY : s[nt][e];
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float scale_old[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      scale_old[r] = ex2_approx(m_run[r] - m_use);    // log2 domain; m_run = -inf -> 0
      m_run[r] = m_new;
      mx[r] = m_use;
      l_run[r] *= scale_old[r];
    }
    float ls[2] = {0.f, 0.f};
    uint32_t pa[8][2];      // P as bf16 pairs: [n-tile][row half]
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = ex2_approx(s[nt][0] - mx[0]), p1 = ex2_approx(s[nt][1] - mx[0]);
      const float p2 = ex2_approx(s[nt][2] - mx[1]), p3 = ex2_approx(s[nt][3] - mx[1]);
      ls[0] += p0 + p1;
      ls[1] += p2 + p3;
      pa[nt][0] = pack_bf16x2(p0, p1);
      pa[nt][1] = pack_bf16x2(p2, p3);
    }
    l_run[0] += ls[0];
    l_run[1] += ls[1];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      o[i][0] *= scale_old[0]; o[i][1] *= scale_old[0];
      o[i][2] *= scale_old[1]; o[i][3] *= scale_old[1];
    }
    // ---- O += P V   (k = 64 keys in 4 steps of 16)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks >= np_valid) break;
      const uint32_t a[4] = {pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1]};
#pragma unroll
      for (int dp = 0; dp < NT / 2; ++dp) {   // pairs of output d n-tiles
        uint32_t bfr[4];
        const int mi = lane >> 3;
        ldsm_x4_t(bfr, cV + (ks * 16 + (mi & 1) * 8 + (lane & 7)) * LD + dp * 16 + (mi >> 1) * 8);
        mma_bf16(o[2 * dp], a, bfr[0], bfr[1]);
        mma_bf16(o[2 * dp + 1], a, bfr[2], bfr[3]);
      }
    }
    __syncthreads();
  }

  // ---- finalize: row sums across the quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  __nv_bfloat16* og = p.out + b * p.o_bs + static_cast<long long>(h) * p.d;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int dcol = i * 8 + 2 * t;
    if (dcol < p.d) {
      if (qi0 < p.sq)
        *reinterpret_cast<uint32_t*>(og + static_cast<long long>(qi0) * p.o_ss + dcol) = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
      if (qi0 + 8 < p.sq)
        *reinterpret_cast<uint32_t*>(og + static_cast<long long>(qi0 + 8) * p.o_ss + dcol) = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
    }
  }
}

template <int DP, int MODE>
int launch_attn(const AttnP& p, int batch, cudaStream_t st) {
  constexpr int LD = DP + 8;
  int nrel = p.relpos ? (2 * p.win - 1) * (2 * p.win - 1) + 3 : 0;
  const size_t smem = static_cast<size_t>(kBQ + 4 * 64) * LD * 2 + static_cast<size_t>(nrel) * 4 + 64 * 4 + 16;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(svla_flash_attn_kernel<DP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      svla_set_error("svla_attention: smem opt-in %zu failed: %s", smem, cudaGetErrorString(e));
      return -2;
    }
    configured = smem;
  }
  dim3 grid((p.sq + kBQ - 1) / kBQ, p.hq, batch);
  svla_flash_attn_kernel<DP, MODE><<<grid, kAttnThreads, smem, st>>>(p);
  SVLA_LAUNCH_CHECK("svla_flash_attn");
  return 0;
}

// ------------------------------------------------------------------------------------------ decode (q_len = 1)
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) { f[2 * u] = bf16_bits_to_float(w[u] & 0xFFFFu); f[2 * u + 1] = bf16_bits_to_float(w[u] >> 16); }
}
constexpr int kDecThreads = 256;     // 8 warps, 2 CTAs per SM: all 256 (batch, kv-head) CTAs of a B=64 step are resident in one wave
constexpr int kMaxGroup = 4;

// One CTA per (batch, kv head), 16 warps.  Every warp streams its share of the keys with K and V rows of 4 keys in
// flight together (8 x 16-byte loads per lane), keeps a private online-softmax state (m, l, acc) for each query head
// of the GQA group in registers, and the 16 partial states are merged once through shared memory: a single pass over
// the cache, one block barrier.
template <int D>
__global__ void __launch_bounds__(kDecThreads)
svla_decode_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                        const __nv_bfloat16* __restrict__ vc, __nv_bfloat16* __restrict__ out, int hq, int hkv, int smax,
                        int ctx, float scale, float softcap, const int* __restrict__ kv_start) {
  constexpr int PER = D / 32;
  constexpr int KB = 4;
  constexpr int NW = kDecThreads / 32;
  static_assert(PER == 8, "one 16-byte load per lane per row");
  extern __shared__ float sm_dec[];
  const int grp = hq / hkv;
  float* s_m = sm_dec;                      // [NW][kMaxGroup]
  float* s_l = s_m + NW * kMaxGroup;        // [NW][kMaxGroup]
  float* s_acc = s_l + NW * kMaxGroup;      // [NW][grp][D]
  const int b = blockIdx.y, hk = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this lane's 8 query dims for every head of the group
  float qv[kMaxGroup][PER];
#pragma unroll
  for (int gi = 0; gi < kMaxGroup; ++gi) {
    if (gi < grp) unpack8(__ldg(reinterpret_cast<const uint4*>(q + (static_cast<long long>(b) * hq + hk * grp + gi) * D + lane * PER)), qv[gi]);
    else {
#pragma unroll
      for (int e = 0; e < PER; ++e) qv[gi][e] = 0.f;
    }
  }
  const __nv_bfloat16* kb = kc + (static_cast<long long>(b) * smax * hkv + hk) * D + lane * PER;
  const __nv_bfloat16* vb = vc + (static_cast<long long>(b) * smax * hkv + hk) * D + lane * PER;
  const long long row_stride = static_cast<long long>(hkv) * D;
  float m_run[kMaxGroup], l_run[kMaxGroup], acc[kMaxGroup][PER];
#pragma unroll
  for (int gi = 0; gi < kMaxGroup; ++gi) {
    m_run[gi] = -INFINITY; l_run[gi] = 0.f;
#pragma unroll
    for (int e = 0; e < PER; ++e) acc[gi][e] = 0.f;
  }
  const float inv_cap = softcap > 0.f ? 1.f / softcap : 0.f;
  const int kstart = kv_start ? kv_start[b] : 0;              // left-padded prompt: keys [0, kstart) are masked
  for (int j0 = warp * KB; j0 < ctx; j0 += NW * KB) {
    uint4 kr[KB], vr[KB];
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const bool ok = j0 + u < ctx;
      kr[u] = ok ? __ldg(reinterpret_cast<const uint4*>(kb + (j0 + u) * row_stride)) : make_uint4(0, 0, 0, 0);
      vr[u] = ok ? __ldg(reinterpret_cast<const uint4*>(vb + (j0 + u) * row_stride)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const bool ok = j0 + u < ctx && j0 + u >= kstart;           // warp-uniform
      float kv[PER], vv[PER];
      unpack8(kr[u], kv);
      unpack8(vr[u], vv);
#pragma unroll
      for (int gi = 0; gi < kMaxGroup; ++gi) {
        if (gi < grp && ok) {
          float dot = 0.f;
#pragma unroll
          for (int e = 0; e < PER; ++e) dot += kv[e] * qv[gi][e];
          dot = warp_sum(dot) * scale;
          if (softcap > 0.f) dot = softcap * tanh_small(dot * inv_cap);
          const float m_new = fmaxf(m_run[gi], dot);
          const float corr = __expf(m_run[gi] - m_new);      // exp(-inf) = 0 on the first key
          const float pj = __expf(dot - m_new);
          l_run[gi] = l_run[gi] * corr + pj;
#pragma unroll
          for (int e = 0; e < PER; ++e) acc[gi][e] = acc[gi][e] * corr + pj * vv[e];
          m_run[gi] = m_new;
        }
      }
    }
  }
  // merge the NW partial states
#pragma unroll
  for (int gi = 0; gi < kMaxGroup; ++gi) {
    if (gi < grp) {
      if (lane == 0) { s_m[warp * kMaxGroup + gi] = m_run[gi]; s_l[warp * kMaxGroup + gi] = l_run[gi]; }
#pragma unroll
      for (int e = 0; e < PER; ++e) s_acc[(warp * grp + gi) * D + lane * PER + e] = acc[gi][e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < grp * D; i += kDecThreads) {
    const int gi = i / D, dd = i - gi * D;
    float mmax = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) mmax = fmaxf(mmax, s_m[w * kMaxGroup + gi]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float mw = s_m[w * kMaxGroup + gi];
      const float f = (mw == -INFINITY) ? 0.f : __expf(mw - mmax);
      num += f * s_acc[(w * grp + gi) * D + dd];
      den += f * s_l[w * kMaxGroup + gi];
    }
    out[(static_cast<long long>(b) * hq + hk * grp + gi) * D + dd] = __float2bfloat16(num / den);
  }
}


// G3f: one decode step of one (batch, kv-head): RoPE of the new q/k (from the split-K fp32 partial sums of the qkv
// projection), KV-cache append, soft-capped softmax attention over the cache and the new key -- one launch instead of
// rope_kv + decode_attention.  The item arithmetic lives in decode_attn_item.cuh (shared with the persistent decode kernel);
// here: 6-stage ring (~80 KB in flight per CTA), 2 CTAs per SM -> all 256 CTAs of a B=64 step are resident at once.
constexpr int kFusedStages = 6;
// Few items (batch x kv heads <= #SMs, e.g. the batch-1 latency path: 4 CTAs): one CTA per SM with a 12-stage ring, i.e. ~190 KB
// of the item's ~290 KB of K / V rows in flight at once -- a lone CTA is bound by its memory-level parallelism, not by HBM.
constexpr int kFusedStagesDeep = 12;

template <int GRP, int STAGES>
__global__ void __launch_bounds__(kDecThreads, STAGES <= 6 ? 2 : 1)
svla_decode_attn_fused_kernel(const float* __restrict__ qkv_f32, int n_partials, long long partial_stride,
                              __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc, __nv_bfloat16* __restrict__ out,
                              int hq, int hkv, int smax, int ctx, float theta, float scale, float softcap,
                              const int* __restrict__ kv_start, __nv_bfloat16* __restrict__ out_lo, int window) {
  constexpr int D = 256;
  extern __shared__ __align__(16) uint8_t sm_fused[];
  svla_dec::ItemSmem sm;
  sm.stage = sm_fused;                                                                                   // [STAGES][32][528]
  sm.q = reinterpret_cast<float*>(sm_fused + STAGES * svla_dec::kItemRows * svla_dec::kItemPitch);       // [GRP][256]
  sm.red = sm.q + GRP * D;                                                                               // [GRP][256]
  sm.newk = reinterpret_cast<__nv_bfloat16*>(sm.red + GRP * D);                                          // [256]
  sm.newv = sm.newk + D;                                                                                 // [256]
  sm.inv = reinterpret_cast<float*>(sm.newv + D);                                                        // [4]
  sm.wred = sm.inv + 4;                                                                                  // [16]
  sm.p = sm.wred + 16;                                                                                   // [GRP][ctx_pad]
  svla_dec::ItemArgs a;
  a.b = blockIdx.y; a.hk = blockIdx.x;
  a.qkv = qkv_f32 + static_cast<long long>(a.b) * (hq + 2 * hkv) * D;
  a.n_partials = n_partials; a.partial_stride = partial_stride;
  a.kc = kc; a.vc = vc; a.out = out; a.out_lo = out_lo;
  a.hq = hq; a.hkv = hkv; a.smax = smax; a.ctx = ctx;
  a.kstart = kv_start ? kv_start[a.b] : 0;              // immutable input (not written by the PDL predecessor)
  a.kmask = window > 0 ? max(a.kstart, ctx - window) : a.kstart;     // sliding-window layer: the last `window` slots only
  a.theta = theta; a.scale = scale; a.softcap = softcap;
  pdl_launch_dependents();          // lets the o-projection GEMM start prefetching its weights
  // The cached rows [0, ctx-1) were written by the prefill or by this layer's kernel of an EARLIER decode step, i.e. at least
  // one full layer chain (>= 7 launches) upstream.  A PDL kernel can only overlap predecessors that are still resident and
  // blocked in griddepcontrol.wait; a whole chain of them cannot be resident at once, so those rows are complete and the
  // ring is primed while the qkv projection (the direct predecessor) is still running; griddepcontrol.wait comes before the
  // first read of its partial sums.
  svla_dec::decode_attn_item<GRP, STAGES>(a, sm, static_cast<int>(threadIdx.x), [] { __syncthreads(); }, [] { pdl_wait(); });
}

}  // namespace

int svla_attention_tc_try(const SvlaAttnArgs* a, void* stream);      // attention_tc.cu (tcgen05 / TMEM / TMA)

// SVLA_ATTN_IMPL=mma forces the mma.sync kernel everywhere (A/B measurements, debugging)
static bool attn_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SVLA_ATTN_IMPL");
    v = (e && strcmp(e, "mma") == 0) ? 0 : 1;
  }
  return v == 1;
}

extern "C" int svla_attention(const SvlaAttnArgs* a, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVLA_REQUIRE(a && a->q && a->k && a->v && a->out, "svla_attention: null pointer");
  SVLA_REQUIRE(a->d > 0 && (a->d % 8) == 0 && a->d <= 256, "svla_attention: head dim %d unsupported", a->d);
  SVLA_REQUIRE(a->hkv > 0 && a->hq % a->hkv == 0, "svla_attention: hq %% hkv != 0");
  SVLA_REQUIRE(a->sq > 0 && a->sk > 0 && a->batch > 0, "svla_attention: empty problem");
  SVLA_REQUIRE(!(a->relpos_table && a->softcap > 0.f), "svla_attention: relative-position bias and soft-capping are exclusive");
  SVLA_REQUIRE(!a->relpos_table || (a->relpos_win > 0 && a->sk <= a->relpos_win * a->relpos_win + 1 && a->sq <= a->relpos_win * a->relpos_win + 1),
               "svla_attention: sequence longer than the relative-position window");
  SVLA_REQUIRE((a->q_ss % 8) == 0 && (a->k_ss % 8) == 0 && (a->v_ss % 8) == 0 && (a->o_ss % 2) == 0,
               "svla_attention: row strides must keep 16-byte alignment");
  if (attn_tc_enabled()) {
    const int rc = svla_attention_tc_try(a, stream);      // 1 = shape not covered by the tcgen05 kernel
    if (rc != 1) return rc;
  }
  SVLA_REQUIRE(!a->lse, "svla_attention: the row log-sum-exp output is provided by the tcgen05 kernels only (d=%d, this shape fell back)", a->d);
  AttnP p;
  p.q = static_cast<const __nv_bfloat16*>(a->q); p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v); p.out = static_cast<__nv_bfloat16*>(a->out);
  p.q_bs = a->q_bs; p.q_ss = a->q_ss; p.k_bs = a->k_bs; p.k_ss = a->k_ss; p.v_bs = a->v_bs; p.v_ss = a->v_ss;
  p.o_bs = a->o_bs; p.o_ss = a->o_ss;
  p.hq = a->hq; p.hkv = a->hkv; p.sq = a->sq; p.sk = a->sk; p.d = a->d;
  p.scale = a->scale; p.softcap = a->softcap; p.causal = a->causal; p.relpos = a->relpos_table; p.win = a->relpos_win; p.head_major = a->relpos_head_major; p.kv_start = a->kv_start;
  p.prefix = a->causal ? a->causal_prefix : 0;
  p.window = a->window;
  const int mode = (p.relpos ? 1 : 0) | (p.softcap > 0.f ? 2 : 0) | (p.causal ? 4 : 0);
  if (a->d <= 32) return launch_attn<32, 8>(p, a->batch, st);
  if (a->d <= 64) return mode == 1 ? launch_attn<64, 1>(p, a->batch, st) : launch_attn<64, 8>(p, a->batch, st);
  if (a->d <= 80) return mode == 0 ? launch_attn<80, 0>(p, a->batch, st) : launch_attn<80, 8>(p, a->batch, st);
  if (a->d <= 128) return launch_attn<128, 8>(p, a->batch, st);
  return mode == 2 ? launch_attn<256, 2>(p, a->batch, st) : launch_attn<256, 8>(p, a->batch, st);
}

extern "C" int svla_decode_attention(const void* q, const void* kcache, const void* vcache, void* out, int batch, int hq,
                                     int hkv, int d, int smax, int ctx, float scale, float softcap, const int32_t* kv_start,
                                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVLA_REQUIRE(q && kcache && vcache && out, "svla_decode_attention: null pointer");
  SVLA_REQUIRE(d == 256, "svla_decode_attention: head dim %d unsupported (256 only)", d);
  SVLA_REQUIRE(hkv > 0 && hq % hkv == 0 && hq / hkv <= kMaxGroup, "svla_decode_attention: bad GQA group");
  SVLA_REQUIRE(ctx > 0 && ctx <= smax, "svla_decode_attention: ctx %d out of range", ctx);
  const int grp = hq / hkv;
  const size_t smem = (2 * static_cast<size_t>(kDecThreads / 32) * kMaxGroup + static_cast<size_t>(kDecThreads / 32) * grp * d) * sizeof(float);
  SVLA_REQUIRE(smem <= 200 * 1024, "svla_decode_attention: context too long for shared memory");
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    cudaFuncSetAttribute(svla_decode_attn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  dim3 grid(hkv, batch);
  svla_decode_attn_kernel<256><<<grid, kDecThreads, smem, st>>>(
      static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kcache),
      static_cast<const __nv_bfloat16*>(vcache), static_cast<__nv_bfloat16*>(out), hq, hkv, smax, ctx, scale, softcap, kv_start);
  SVLA_LAUNCH_CHECK("svla_decode_attn");
  return 0;
}

// Decode step of one layer after the qkv projection: RoPE + cache append + attention in one launch (see the kernel).
static int decode_attention_fused_impl(const float* qkv_f32, int n_partials, int64_t partial_stride, void* kcache, void* vcache,
                                       void* out, void* out_lo, int batch, int hq, int hkv, int d, int smax, int ctx, float theta,
                                       float scale, float softcap, const int32_t* kv_start, int window, void* stream);

extern "C" int svla_decode_attention_fused(const float* qkv_f32, int n_partials, int64_t partial_stride, void* kcache, void* vcache,
                                           void* out, int batch, int hq, int hkv, int d, int smax, int ctx, float theta,
                                           float scale, float softcap, const int32_t* kv_start, void* stream) {
  return decode_attention_fused_impl(qkv_f32, n_partials, partial_stride, kcache, vcache, out, nullptr, batch, hq, hkv, d, smax, ctx,
                                     theta, scale, softcap, kv_start, 0, stream);
}

// Extended entry: out_lo != NULL -> the output leaves as two bf16 planes (hi = bf16(o), lo = bf16(o - hi)) for the X_HILO mode of
// the o-projection (svla_gemm_skinny) and the query stays in fp32; window > 0 -> sliding-window layer (even Gemma2 layers,
// model/modeling_gemma2.py:343,441-473): only the last `window` cache slots [ctx - window, ctx) receive weight.
extern "C" int svla_decode_attention_fused_ex(const float* qkv_f32, int n_partials, int64_t partial_stride, void* kcache,
                                              void* vcache, void* out_hi, void* out_lo, int batch, int hq, int hkv, int d, int smax,
                                              int ctx, float theta, float scale, float softcap, const int32_t* kv_start, int window,
                                              void* stream) {
  SVLA_REQUIRE(window >= 0, "svla_decode_attention_fused_ex: negative window");
  return decode_attention_fused_impl(qkv_f32, n_partials, partial_stride, kcache, vcache, out_hi, out_lo, batch, hq, hkv, d, smax, ctx,
                                     theta, scale, softcap, kv_start, window, stream);
}

static int decode_attention_fused_impl(const float* qkv_f32, int n_partials, int64_t partial_stride, void* kcache, void* vcache,
                                       void* out, void* out_lo, int batch, int hq, int hkv, int d, int smax, int ctx, float theta,
                                       float scale, float softcap, const int32_t* kv_start, int window, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVLA_REQUIRE(qkv_f32 && kcache && vcache && out, "svla_decode_attention_fused: null pointer");
  SVLA_REQUIRE(d == 256, "svla_decode_attention_fused: head dim %d unsupported (256 only)", d);
  SVLA_REQUIRE(hkv > 0 && hq % hkv == 0 && (hq / hkv == 1 || hq / hkv == 2), "svla_decode_attention_fused: GQA group must be 1 or 2");
  SVLA_REQUIRE(ctx > 0 && ctx <= smax && n_partials >= 1 && batch > 0 && batch <= 65535, "svla_decode_attention_fused: bad ctx / partials / batch");
  const int grp = hq / hkv;
  const int ctx_pad = (ctx + 31) & ~31;
  auto smem_for = [&](int stages) {
    return static_cast<size_t>(stages) * svla_dec::kItemRows * svla_dec::kItemPitch + (2 * grp * 256 + 4 + 16 + grp * ctx_pad) * sizeof(float) + 2 * 256 * 2;
  };
  const bool deep = static_cast<long long>(batch) * hkv <= svla_num_sms() && smem_for(kFusedStagesDeep) <= 225 * 1024;
  const size_t smem = smem_for(deep ? kFusedStagesDeep : kFusedStages);
  SVLA_REQUIRE(smem <= 225 * 1024, "svla_decode_attention_fused: context %d too long for shared memory", ctx);
  static size_t configured[2][3] = {{0, 0, 0}, {0, 0, 0}};
  if (smem > configured[deep][grp]) {
    cudaError_t e = deep ? (grp == 1 ? cudaFuncSetAttribute(svla_decode_attn_fused_kernel<1, kFusedStagesDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                                     : cudaFuncSetAttribute(svla_decode_attn_fused_kernel<2, kFusedStagesDeep>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)))
                         : (grp == 1 ? cudaFuncSetAttribute(svla_decode_attn_fused_kernel<1, kFusedStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                                     : cudaFuncSetAttribute(svla_decode_attn_fused_kernel<2, kFusedStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SVLA_REQUIRE(e == cudaSuccess, "svla_decode_attention_fused: smem opt-in %zu failed: %s", smem, cudaGetErrorString(e));
    configured[deep][grp] = smem;
  }
  dim3 grid(hkv, batch);
  auto* kcp = static_cast<__nv_bfloat16*>(kcache);
  auto* vcp = static_cast<__nv_bfloat16*>(vcache);
  auto* op = static_cast<__nv_bfloat16*>(out);
  auto* olp = static_cast<__nv_bfloat16*>(out_lo);
  const long long ps = partial_stride;
  cudaError_t le;
  if (deep)
    le = grp == 1 ? svla_launch_pdl(svla_decode_attn_fused_kernel<1, kFusedStagesDeep>, grid, dim3(kDecThreads), smem, st, qkv_f32, n_partials, ps, kcp, vcp, op,
                                    hq, hkv, smax, ctx, theta, scale, softcap, kv_start, olp, window)
                  : svla_launch_pdl(svla_decode_attn_fused_kernel<2, kFusedStagesDeep>, grid, dim3(kDecThreads), smem, st, qkv_f32, n_partials, ps, kcp, vcp, op,
                                    hq, hkv, smax, ctx, theta, scale, softcap, kv_start, olp, window);
  else
    le = grp == 1 ? svla_launch_pdl(svla_decode_attn_fused_kernel<1, kFusedStages>, grid, dim3(kDecThreads), smem, st, qkv_f32, n_partials, ps, kcp, vcp, op,
                                    hq, hkv, smax, ctx, theta, scale, softcap, kv_start, olp, window)
                  : svla_launch_pdl(svla_decode_attn_fused_kernel<2, kFusedStages>, grid, dim3(kDecThreads), smem, st, qkv_f32, n_partials, ps, kcp, vcp, op,
                                    hq, hkv, smax, ctx, theta, scale, softcap, kv_start, olp, window);
  SVLA_REQUIRE(le == cudaSuccess, "svla_decode_attention_fused: launch failed: %s", cudaGetErrorString(le));
  SVLA_LAUNCH_CHECK("svla_decode_attn_fused");
  return 0;
}
