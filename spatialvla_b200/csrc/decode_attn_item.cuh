// One (batch row, kv head) item of a decode step, shared by the stand-alone fused kernel (attention.cu, the PDL chain) and the
// persistent decode kernel (decode_mega.cu): RoPE of the new q / k from the split-K fp32 partial sums of the qkv projection,
// KV-cache append, soft-capped softmax attention over the cached keys and the new one
// (model/modeling_gemma2.py:130-154,169-195,387-395; positions are 1-indexed, model/modeling_spatialvla.py:371-372).
//   * K rows [0, ctx-1) then V rows [0, ctx-1) flow through a STAGES-deep cp.async ring of 32-row tiles (16-byte pieces,
//     528-byte pitch: conflict-free 16-byte reads);
//   * K pass: 8/GRP threads per (key, head) dot product, q slice in registers, two accumulators, 1-3 shuffles per dot;
//   * softmax over the ctx scores in shared memory by all 256 threads (two block reductions), p stays unnormalised;
//   * V pass: thread = (dim pair, key half), p broadcast from shared memory.
// The new token's key / value never round-trip through global memory: its score and value term come from shared memory.
// 256 threads; `sync()` is the barrier of exactly those threads; `after_prime()` runs between the ring priming and the first
// read of the qkv partial sums (griddepcontrol.wait in the chain, nothing in the persistent kernel).
#pragma once
#include "svla_common.cuh"

namespace svla_dec {

constexpr int kItemThreads = 256;
constexpr int kItemD = 256;
constexpr int kItemRows = 32;
constexpr int kItemPitch = 528;     // bytes per staged row (512 + 16)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) { f[2 * u] = bf16_bits_to_float(w[u] & 0xFFFFu); f[2 * u + 1] = bf16_bits_to_float(w[u] >> 16); }
}
// tanh for soft-capping: |u| is small (scores / 50), so an odd degree-9 Taylor polynomial is exact to fp32 rounding below 0.35
// and the libm path handles the rare large argument.
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

struct ItemSmem {
  uint8_t* stage;            // [STAGES][32][528]
  float* q;                  // [GRP][256] (bf16-rounded values)
  float* red;                // [GRP][256]
  __nv_bfloat16* newk;       // [256]
  __nv_bfloat16* newv;       // [256]
  float* inv;                // [GRP] (+ pad to 4)
  float* wred;               // [16] per-warp reduction slots
  float* p;                  // [GRP][ctx_pad]
};

struct ItemArgs {
  const float* qkv;          // this batch row's [(hq + 2 hkv) * D] slice of split 0
  int n_partials;
  long long partial_stride;  // elements between splits
  __nv_bfloat16* kc;         // cache of this layer [B, smax, hkv, D]
  __nv_bfloat16* vc;
  __nv_bfloat16* out;        // [B, hq * D]
  __nv_bfloat16* out_lo;     // NULL, or the lo plane bf16(o - bf16(o)) of the hi/lo activation pair (same layout as out)
  int b, hk, hq, hkv, smax, ctx, kstart;
  int kmask;                 // first cache slot that may receive weight: max(kstart, ctx - sliding window)
  float theta, scale, softcap;
};

template <int GRP, int STAGES, typename Sync, typename AfterPrime>
__device__ __forceinline__ void decode_attn_item(const ItemArgs& a, const ItemSmem& s, int t, Sync sync, AfterPrime after_prime) {
  constexpr int D = kItemD;
  constexpr int TPP = 8 / GRP;            // threads per (key, head) pair
  constexpr int PPT = 32 / TPP;           // 16-byte pieces of a K row per thread
  constexpr int NW = kItemThreads / 32;
  const int hq = a.hq, hkv = a.hkv, ctx = a.ctx, smax = a.smax, b = a.b, hk = a.hk, kstart = a.kstart, kmask = a.kmask;
  const int ctx_pad = (ctx + 31) & ~31;
  const int lane = t & 31, warp = t >> 5;
  const int n_old = ctx - 1;                                  // cached keys; the new token sits at slot ctx - 1
  const int n_chunks = (n_old + kItemRows - 1) / kItemRows;
  const int n_tiles = 2 * n_chunks;
  const long long row_stride = static_cast<long long>(hkv) * D;
  const __nv_bfloat16* kbase = a.kc + (static_cast<long long>(b) * smax * hkv + hk) * D;
  const __nv_bfloat16* vbase = a.vc + (static_cast<long long>(b) * smax * hkv + hk) * D;

  auto issue_tile = [&](int tile) {
    if (tile < n_tiles) {
      const bool is_v = tile >= n_chunks;
      const int key0 = (is_v ? tile - n_chunks : tile) * kItemRows;
      const __nv_bfloat16* base = is_v ? vbase : kbase;
      uint8_t* dst = s.stage + (tile % STAGES) * (kItemRows * kItemPitch);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pc = t + i * kItemThreads;
        const int r = pc >> 5, c16 = pc & 31;
        const bool ok = key0 + r < n_old;
        cp_async16(dst + r * kItemPitch + c16 * 16, base + static_cast<long long>(ok ? key0 + r : 0) * row_stride + c16 * 8, ok);
      }
    }
    cp_async_commit();
  };
  // The cached rows [0, ctx-1) were written by the prefill or by an EARLIER decode step, so the ring may be primed before the
  // qkv projection of this step is complete.
#pragma unroll
  for (int i = 0; i < STAGES - 1; ++i) issue_tile(i);
  after_prime();

  // ---- RoPE of the new token + cache append.  Every split-K partial this thread needs is requested before the first use
  // (one L2 round trip instead of n_partials x 6 serialised ones: 7.6 us -> ~1 us per item at 4 splits); the partial sums are
  // added in split order, like every other consumer of them.
  {
    const float* src = a.qkv;
    const long long cache_row = (static_cast<long long>(b) * smax + n_old) * hkv * D + static_cast<long long>(hk) * D;
    if (t < D / 2) {
      const int j = t;
      float x1[GRP + 1], x2[GRP + 1];
#pragma unroll
      for (int g = 0; g <= GRP; ++g) { x1[g] = 0.f; x2[g] = 0.f; }
      for (int sp0 = 0; sp0 < a.n_partials; sp0 += 4) {
        float v1[4][GRP + 1], v2[4][GRP + 1];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = sp0 + u < a.n_partials;
          const float* sp_src = src + static_cast<long long>(ok ? sp0 + u : 0) * a.partial_stride;
#pragma unroll
          for (int g = 0; g <= GRP; ++g) {
            const long long col = (g < GRP) ? static_cast<long long>(hk * GRP + g) * D : static_cast<long long>(hq + hk) * D;
            v1[u][g] = ok ? __ldcg(sp_src + col + j) : 0.f;
            v2[u][g] = ok ? __ldcg(sp_src + col + j + D / 2) : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int g = 0; g <= GRP; ++g) { x1[g] += v1[u][g]; x2[g] += v2[u][g]; }
        }
      }
      const float inv_freq = 1.0f / powf(a.theta, static_cast<float>(2 * j) / static_cast<float>(D));
      float sn, cs;
      sincosf(static_cast<float>(ctx - kstart) * inv_freq, &sn, &cs);      // position of the new token: ctx - leading pads
#pragma unroll
      for (int g = 0; g <= GRP; ++g) {
        const float r1 = x1[g] * cs - x2[g] * sn, r2 = x2[g] * cs + x1[g] * sn;
        const __nv_bfloat16 o1 = __float2bfloat16(r1), o2 = __float2bfloat16(r2);
        if (g < GRP) {                // the query stays in fp32 on the hi/lo chain (bf16 like the prefill kernels' otherwise)
          s.q[g * D + j] = a.out_lo ? r1 : __bfloat162float(o1);
          s.q[g * D + j + D / 2] = a.out_lo ? r2 : __bfloat162float(o2);
        } else {
          s.newk[j] = o1; s.newk[j + D / 2] = o2;
          a.kc[cache_row + j] = o1; a.kc[cache_row + j + D / 2] = o2;
        }
      }
    } else {
      const int dp = t - D / 2;
      const long long col = static_cast<long long>(hq + hkv + hk) * D + 2 * dp;
      float y0 = 0.f, y1 = 0.f;
      for (int sp0 = 0; sp0 < a.n_partials; sp0 += 4) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = sp0 + u < a.n_partials;
          v[u] = ok ? __ldcg(reinterpret_cast<const float2*>(src + static_cast<long long>(sp0 + u) * a.partial_stride + col)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { y0 += v[u].x; y1 += v[u].y; }
      }
      const __nv_bfloat162 v2 = __floats2bfloat162_rn(y0, y1);
      *reinterpret_cast<__nv_bfloat162*>(s.newv + 2 * dp) = v2;
      *reinterpret_cast<__nv_bfloat162*>(a.vc + cache_row + 2 * dp) = v2;
    }
    for (int i = t; i < GRP * (ctx_pad - n_old); i += kItemThreads) {      // p = 0 behind the last cached key
      const int g = i / (ctx_pad - n_old), k = n_old + i % (ctx_pad - n_old);
      s.p[g * ctx_pad + k] = 0.f;
    }
  }
  sync();

  // ---- K pass
  const int part = t % TPP, g_k = (t / TPP) % GRP, key_l = t >> 3;
  float qr[PPT][8];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const float4 a0 = *reinterpret_cast<const float4*>(s.q + g_k * D + (i * TPP + part) * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(s.q + g_k * D + (i * TPP + part) * 8 + 4);
    qr[i][0] = a0.x; qr[i][1] = a0.y; qr[i][2] = a0.z; qr[i][3] = a0.w;
    qr[i][4] = a1.x; qr[i][5] = a1.y; qr[i][6] = a1.z; qr[i][7] = a1.w;
  }
  const float inv_cap = a.softcap > 0.f ? 1.f / a.softcap : 0.f;
  for (int tile = 0; tile < n_chunks; ++tile) {
    cp_async_wait<STAGES - 2>();
    sync();
    issue_tile(tile + STAGES - 1);
    const uint8_t* row = s.stage + (tile % STAGES) * (kItemRows * kItemPitch) + key_l * kItemPitch;
    float d0 = 0.f, d1 = 0.f;               // two chains: the 8 x PPT dependent FMAs of one accumulator bound this pass
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      float kv[8];
      unpack8(*reinterpret_cast<const uint4*>(row + (i * TPP + part) * 16), kv);
#pragma unroll
      for (int e = 0; e < 8; e += 2) { d0 = fmaf(kv[e], qr[i][e], d0); d1 = fmaf(kv[e + 1], qr[i][e + 1], d1); }
    }
    float dot = d0 + d1;
#pragma unroll
    for (int o = 1; o < TPP; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const int key = tile * kItemRows + key_l;
    if (part == 0 && key < n_old) {
      float sc = dot * a.scale;
      if (a.softcap > 0.f) sc = a.softcap * tanh_small(sc * inv_cap);
      s.p[g_k * ctx_pad + key] = key < kmask ? -INFINITY : sc;        // padded prompt slots / slots behind the sliding window never receive weight
    }
  }
  // ---- new key's score (warp g computes head g)
  if (warp < GRP) {
    float kv[8], dot = 0.f;
    unpack8(*reinterpret_cast<const uint4*>(s.newk + lane * 8), kv);
#pragma unroll
    for (int e = 0; e < 8; ++e) dot = fmaf(kv[e], s.q[warp * D + lane * 8 + e], dot);
    dot = warp_sum(dot) * a.scale;
    if (a.softcap > 0.f) dot = a.softcap * tanh_small(dot * inv_cap);
    if (lane == 0) s.p[warp * ctx_pad + n_old] = dot;
  }
  sync();
  // ---- softmax: NW / GRP warps per head, two block reductions; p stays unnormalised, 1/sum is applied at the end
  {
    constexpr int WPH = NW / GRP;                     // warps per head
    const int g = warp / WPH, tg = t - g * (WPH * 32);
    float* pr = s.p + g * ctx_pad;
    float m = -INFINITY;
    for (int k = tg; k < ctx; k += WPH * 32) m = fmaxf(m, pr[k]);
    m = warp_max(m);
    if (lane == 0) s.wred[warp] = m;
    sync();
#pragma unroll
    for (int w = 0; w < WPH; ++w) m = fmaxf(m, s.wred[g * WPH + w]);
    float sum = 0.f;
    for (int k = tg; k < ctx; k += WPH * 32) {
      const float e = __expf(pr[k] - m);
      pr[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) s.wred[NW + warp] = sum;
    sync();
    if (tg == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < WPH; ++w) tot += s.wred[NW + g * WPH + w];
      s.inv[g] = 1.f / tot;
    }
  }
  // ---- V pass: thread = (dim pair dp, key half kh)   (the first barrier of the loop publishes s.inv / the final p values)
  const int dp = t & 127, kh = t >> 7;
  float acc[GRP][2];
#pragma unroll
  for (int g = 0; g < GRP; ++g) { acc[g][0] = 0.f; acc[g][1] = 0.f; }
  if (n_chunks == 0) sync();
  for (int tile = n_chunks; tile < n_tiles; ++tile) {
    cp_async_wait<STAGES - 2>();
    sync();
    issue_tile(tile + STAGES - 1);
    const uint8_t* st = s.stage + (tile % STAGES) * (kItemRows * kItemPitch);
    const int key0 = (tile - n_chunks) * kItemRows + kh * 16;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(st + (kh * 16 + r) * kItemPitch + dp * 4));
#pragma unroll
      for (int g = 0; g < GRP; ++g) {
        const float pw = s.p[g * ctx_pad + key0 + r];          // 0 for rows behind the last cached key
        acc[g][0] = fmaf(pw, v.x, acc[g][0]);
        acc[g][1] = fmaf(pw, v.y, acc[g][1]);
      }
    }
  }
  cp_async_wait<0>();
  if (kh == 1) {
#pragma unroll
    for (int g = 0; g < GRP; ++g) { s.red[g * D + 2 * dp] = acc[g][0]; s.red[g * D + 2 * dp + 1] = acc[g][1]; }
  }
  sync();
  if (kh == 0) {
    const float2 nv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(s.newv + 2 * dp));
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
      const float pn = s.p[g * ctx_pad + n_old], inv = s.inv[g];
      const float o0 = (acc[g][0] + s.red[g * D + 2 * dp] + pn * nv.x) * inv;
      const float o1 = (acc[g][1] + s.red[g * D + 2 * dp + 1] + pn * nv.y) * inv;
      const long long oi = (static_cast<long long>(b) * hq + hk * GRP + g) * D + 2 * dp;
      const __nv_bfloat162 ohi = __floats2bfloat162_rn(o0, o1);
      *reinterpret_cast<__nv_bfloat162*>(a.out + oi) = ohi;
      if (a.out_lo) {
        const float2 hf = __bfloat1622float2(ohi);
        *reinterpret_cast<__nv_bfloat162*>(a.out_lo + oi) = __floats2bfloat162_rn(o0 - hf.x, o1 - hf.y);
      }
    }
  }
}

}  // namespace svla_dec
