// Memory-bound fused kernels of the predict_action path (M1-M7, M9, ZoeDepth metric-bins tail, Ego3D).
// All of them stream each operand once with 128-bit accesses where the layout allows, keep statistics in
// fp32 and use warp-shuffle reductions.  Reference lines replaced: see include/spatialvla_b200.h.
#include <cstdlib>
#include "svla_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kRowThreads = 256;
constexpr int kMaxVec = 5;   // float4 per thread: rows up to 256*5*4 = 5120 columns

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]);

// ------------------------------------------------------------------------------------------ row loaders
__device__ __forceinline__ int load_row(const float* __restrict__ x, int cols, float4 (&v)[kMaxVec]) {
  const int nv = cols >> 2;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    if (i < nv) { v[k] = reinterpret_cast<const float4*>(x)[i]; cnt = k + 1; }
    else v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return cnt;
}

// ------------------------------------------------------------------------------------------ M1 LayerNorm
__global__ void __launch_bounds__(kRowThreads)
svla_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                      int cols, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int relu) {
  __shared__ float sh[33];
  const long long row = blockIdx.x;
  const float* xr = x + row * cols;
  float4 v[kMaxVec];
  load_row(xr, cols, v);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
  const float mean = block_sum(s, sh) / cols;
  const int nv = cols >> 2;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    if (threadIdx.x + k * kRowThreads < nv) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(block_sum(ss, sh) / cols + eps);
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    if (i < nv) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[i];
      const float4 b = reinterpret_cast<const float4*>(beta)[i];
      float4 o;
      o.x = (v[k].x - mean) * rstd * g.x + b.x;
      o.y = (v[k].y - mean) * rstd * g.y + b.y;
      o.z = (v[k].z - mean) * rstd * g.z + b.z;
      o.w = (v[k].w - mean) * rstd * g.w + b.w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + row * cols)[i] = o;
      if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + row * cols)[i] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

// ------------------------------------------------------------------------------------------ M2 Gemma2 sandwich norm
__global__ void __launch_bounds__(kRowThreads)
svla_rmsnorm_residual_kernel(float* __restrict__ x, const float* __restrict__ branch, const float* __restrict__ w_post,
                             const float* __restrict__ w_pre, float eps, int cols, __nv_bfloat16* __restrict__ out_bf16,
                             int n_partials, long long partial_stride, __nv_bfloat16* __restrict__ out_lo) {
  __shared__ float sh[33];
  pdl_launch_dependents();          // the next kernel (a weight-streaming GEMM in the decode chain) may start its prefetch
  const long long row = blockIdx.x;
  const int nv = cols >> 2;
  // The norm weights are immutable: they are requested BEFORE griddepcontrol.wait (their L2 round trips hide behind the tail of
  // the producing GEMM) and all at once.  Loaded inside the output loops they were one dependent L2 round trip per float4 slice,
  // and with the hi/lo output the compiler serialised the slices behind each other (6.0 -> 7.0 us per launch at 64 rows).
  float4 wpost[kMaxVec], wpre[kMaxVec];
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    wpost[k] = (branch && i < nv) ? __ldg(reinterpret_cast<const float4*>(w_post) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    wpre[k] = (w_pre && i < nv) ? __ldg(reinterpret_cast<const float4*>(w_pre) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();
  float4 xv[kMaxVec];
  load_row(x + row * cols, cols, xv);
  if (branch) {
    float4 bv[kMaxVec];
    load_row(branch + row * cols, cols, bv);
    // split-K partial sums, added in a fixed order; four partial rows are in flight per trip (the decode step is
    // latency-bound: 64 rows, one block each -- serialised L2 round trips were 60% of this kernel)
    for (int sp = 1; sp < n_partials; sp += 4) {
      float4 pv[4][kMaxVec];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (sp + u < n_partials) load_row(branch + (sp + u) * partial_stride + row * cols, cols, pv[u]);
        else {
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k) pv[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < kMaxVec; ++k) { bv[k].x += pv[u][k].x; bv[k].y += pv[u][k].y; bv[k].z += pv[u][k].z; bv[k].w += pv[u][k].w; }
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) ss += bv[k].x * bv[k].x + bv[k].y * bv[k].y + bv[k].z * bv[k].z + bv[k].w * bv[k].w;
    const float r = rsqrtf(block_sum(ss, sh) / cols + eps);
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) {
      const int i = threadIdx.x + k * kRowThreads;
      if (i < nv) {
        const float4 w = wpost[k];
        xv[k].x += bv[k].x * r * (1.f + w.x);
        xv[k].y += bv[k].y * r * (1.f + w.y);
        xv[k].z += bv[k].z * r * (1.f + w.z);
        xv[k].w += bv[k].w * r * (1.f + w.w);
        reinterpret_cast<float4*>(x + row * cols)[i] = xv[k];
      }
    }
  }
  if (w_pre) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) ss += xv[k].x * xv[k].x + xv[k].y * xv[k].y + xv[k].z * xv[k].z + xv[k].w * xv[k].w;
    const float r = rsqrtf(block_sum(ss, sh) / cols + eps);
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) {
      const int i = threadIdx.x + k * kRowThreads;
      if (i < nv) {
        const float4 w = wpre[k];
        const float o0 = xv[k].x * r * (1.f + w.x), o1 = xv[k].y * r * (1.f + w.y);
        const float o2 = xv[k].z * r * (1.f + w.z), o3 = xv[k].w * r * (1.f + w.w);
        const uint2 hi = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        reinterpret_cast<uint2*>(out_bf16 + row * cols)[i] = hi;
        if (out_lo) {       // hi/lo activation planes of the decode chain: lo = bf16(value - hi), consumed by svla_gemm_skinny (X_HILO)
          const float2 h01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.x));
          const float2 h23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi.y));
          reinterpret_cast<uint2*>(out_lo + row * cols)[i] =
              make_uint2(pack_bf16x2(o0 - h01.x, o1 - h01.y), pack_bf16x2(o2 - h23.x, o3 - h23.y));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ warp-per-row variants
// One warp owns one row: no block barriers, every load of the row is in flight before the first reduction, shuffle
// reductions only.  V = float4 per lane (cols <= 128 * V).  The block-per-row kernels above remain for wider rows.
template <int V>
__global__ void __launch_bounds__(128)
svla_layernorm_warp_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                           long long rows, int cols, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int relu) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 4LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = cols >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * cols);
  float4 v[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int i = lane + 32 * k;
    v[k] = (i < nv) ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
  const float mean = warp_sum(s) / cols;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    if (lane + 32 * k < nv) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / cols + eps);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i);
      float4 o;
      o.x = (v[k].x - mean) * rstd * g.x + b.x;
      o.y = (v[k].y - mean) * rstd * g.y + b.y;
      o.z = (v[k].z - mean) * rstd * g.z + b.z;
      o.w = (v[k].w - mean) * rstd * g.w + b.w;
      if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + row * cols)[i] = o;
      if (out_bf16) reinterpret_cast<uint2*>(out_bf16 + row * cols)[i] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

template <int V>
__global__ void __launch_bounds__(128)
svla_rmsnorm_residual_warp_kernel(float* __restrict__ x, const float* __restrict__ branch, const float* __restrict__ w_post,
                                  const float* __restrict__ w_pre, float eps, long long rows, int cols,
                                  __nv_bfloat16* __restrict__ out_bf16, int n_partials, long long partial_stride) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 4LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = cols >> 2;
  float4* xr = reinterpret_cast<float4*>(x + row * cols);
  float4 xv[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int i = lane + 32 * k;
    xv[k] = (i < nv) ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (branch) {
    float4 bv[V];
    const float4* br = reinterpret_cast<const float4*>(branch + row * cols);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = lane + 32 * k;
      bv[k] = (i < nv) ? br[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int sp = 1; sp < n_partials; ++sp) {
      const float4* pr = reinterpret_cast<const float4*>(branch + sp * partial_stride + row * cols);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) { const float4 t = pr[i]; bv[k].x += t.x; bv[k].y += t.y; bv[k].z += t.z; bv[k].w += t.w; }
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) ss += bv[k].x * bv[k].x + bv[k].y * bv[k].y + bv[k].z * bv[k].z + bv[k].w * bv[k].w;
    const float r = rsqrtf(warp_sum(ss) / cols + eps);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = lane + 32 * k;
      if (i < nv) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w_post) + i);
        xv[k].x += bv[k].x * r * (1.f + w.x);
        xv[k].y += bv[k].y * r * (1.f + w.y);
        xv[k].z += bv[k].z * r * (1.f + w.z);
        xv[k].w += bv[k].w * r * (1.f + w.w);
        xr[i] = xv[k];
      }
    }
  }
  if (w_pre) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) ss += xv[k].x * xv[k].x + xv[k].y * xv[k].y + xv[k].z * xv[k].z + xv[k].w * xv[k].w;
    const float r = rsqrtf(warp_sum(ss) / cols + eps);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = lane + 32 * k;
      if (i < nv) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w_pre) + i);
        reinterpret_cast<uint2*>(out_bf16 + row * cols)[i] =
            make_uint2(pack_bf16x2(xv[k].x * r * (1.f + w.x), xv[k].y * r * (1.f + w.y)),
                       pack_bf16x2(xv[k].z * r * (1.f + w.z), xv[k].w * r * (1.f + w.w)));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ M3 RoPE + KV cache write
// one block per token; thread pairs (d, d + D/2) of every head
__global__ void svla_rope_kv_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q_out,
                                    __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc, int s, int hq, int hkv, int d,
                                    int smax, int pos0, float theta, const float* __restrict__ qkv_f32, int n_partials,
                                    long long partial_stride, const int* __restrict__ row_pads) {
  const long long tok = blockIdx.x;
  const int b = static_cast<int>(tok / s), si = static_cast<int>(tok % s);
  const int pos = pos0 + si;                       // cache slot
  // PaliGemma positions are 1-indexed; in a left-padded batch they restart on the row's first real token (padding slots: 2)
  const int pad = row_pads ? row_pads[b] : 0;
  const float fpos = static_cast<float>(pos < pad ? 2 : pos - pad + 1);
  const int half = d >> 1;
  const long long width = static_cast<long long>(hq + 2 * hkv) * d;
  const __nv_bfloat16* src = qkv + tok * width;
  auto load = [&](long long col) -> float {
    if (qkv_f32 == nullptr) return __bfloat162float(src[col]);
    float acc = 0.f;
    for (int sp = 0; sp < n_partials; ++sp) acc += qkv_f32[sp * partial_stride + tok * width + col];
    return acc;
  };
  const long long cache_row = (static_cast<long long>(b) * smax + pos) * hkv * d;
  // thread = (frequency j, head lane): one sincos per thread reused across its heads; heads are spread over
  // blockDim.x / half lanes so that the loads of different heads are independent (decode: 64 tokens only)
  const int hlanes = max(1, static_cast<int>(blockDim.x) / half);
  const int j = threadIdx.x % half, hl = threadIdx.x / half;
  if (hl < hlanes) {
    // inv_freq = 1 / theta^(2j/d) in fp32 exactly like torch: base ** (arange(0,d,2).float()/d)
    const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * j) / static_cast<float>(d));
    const float ang = fpos * inv_freq;
    float sn, cs;
    sincosf(ang, &sn, &cs);
    for (int hh = hl; hh < hq + hkv; hh += hlanes) {
      const float x1 = load(hh * d + j);
      const float x2 = load(hh * d + j + half);
      const float o1 = x1 * cs - x2 * sn;
      const float o2 = x2 * cs + x1 * sn;
      if (hh < hq) {
        q_out[tok * hq * d + hh * d + j] = __float2bfloat16(o1);
        q_out[tok * hq * d + hh * d + j + half] = __float2bfloat16(o2);
      } else {
        const int kh = hh - hq;
        kc[cache_row + kh * d + j] = __float2bfloat16(o1);
        kc[cache_row + kh * d + j + half] = __float2bfloat16(o2);
      }
    }
  }
  const long long voff = static_cast<long long>(hq + hkv) * d;
  for (int i = threadIdx.x; i < hkv * d; i += blockDim.x) vc[cache_row + i] = __float2bfloat16(load(voff + i));
}

// Vectorised prefill variant (bf16 qkv in): the block first computes the token's d/2 (cos, sin) pairs ONCE into shared memory
// (one powf + sincosf per thread -- ncu showed the earlier per-thread 8x powf/sincosf made the kernel issue-bound at 39 % of HBM),
// then a thread owns 8 consecutive frequencies (one 16-byte load per half) of every q/k head it visits; V rows are copied with
// 16-byte accesses.
__global__ void __launch_bounds__(128)
svla_rope_kv_vec_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ kc,
                        __nv_bfloat16* __restrict__ vc, int s, int hq, int hkv, int d, int smax, int pos0, float theta,
                        const int* __restrict__ row_pads) {
  __shared__ __align__(16) float s_cs[128], s_sn[128];      // d <= 256
  const long long tok = blockIdx.x;
  const int b = static_cast<int>(tok / s), si = static_cast<int>(tok % s);
  const int pos = pos0 + si;
  const int pad = row_pads ? row_pads[b] : 0;
  const float fpos = static_cast<float>(pos < pad ? 2 : pos - pad + 1);
  const int half = d >> 1;
  const int tph = half >> 3;                       // threads per head
  const long long width = static_cast<long long>(hq + 2 * hkv) * d;
  const __nv_bfloat16* src = qkv + tok * width;
  const long long cache_row = (static_cast<long long>(b) * smax + pos) * hkv * d;
  // V copy first: independent of the trigonometry, its loads are in flight while the sincos table is built
  {
    const uint4* vsrc = reinterpret_cast<const uint4*>(src + static_cast<long long>(hq + hkv) * d);
    uint4* vdst = reinterpret_cast<uint4*>(vc + cache_row);
    for (int i = threadIdx.x; i < (hkv * d) >> 3; i += blockDim.x) vdst[i] = vsrc[i];
  }
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    // inv_freq = 1 / theta^(2j/d) in fp32 exactly like torch: base ** (arange(0,d,2).float()/d)
    const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * j) / static_cast<float>(d));
    sincosf(fpos * inv_freq, &s_sn[j], &s_cs[j]);
  }
  __syncthreads();
  const int jt = threadIdx.x % tph, hl = threadIdx.x / tph, hlanes = blockDim.x / tph;
  if (hl < hlanes) {
    float sn[8], cs[8];
    const float4 c0 = *reinterpret_cast<const float4*>(s_cs + jt * 8), c1 = *reinterpret_cast<const float4*>(s_cs + jt * 8 + 4);
    const float4 n0 = *reinterpret_cast<const float4*>(s_sn + jt * 8), n1 = *reinterpret_cast<const float4*>(s_sn + jt * 8 + 4);
    cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
    sn[0] = n0.x; sn[1] = n0.y; sn[2] = n0.z; sn[3] = n0.w; sn[4] = n1.x; sn[5] = n1.y; sn[6] = n1.z; sn[7] = n1.w;
    for (int hh = hl; hh < hq + hkv; hh += hlanes) {
      float x1[8], x2[8];
      unpack8(*reinterpret_cast<const uint4*>(src + hh * d + jt * 8), x1);
      unpack8(*reinterpret_cast<const uint4*>(src + hh * d + half + jt * 8), x2);
      uint32_t o1[4], o2[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o1[e] = pack_bf16x2(x1[2 * e] * cs[2 * e] - x2[2 * e] * sn[2 * e], x1[2 * e + 1] * cs[2 * e + 1] - x2[2 * e + 1] * sn[2 * e + 1]);
        o2[e] = pack_bf16x2(x2[2 * e] * cs[2 * e] + x1[2 * e] * sn[2 * e], x2[2 * e + 1] * cs[2 * e + 1] + x1[2 * e + 1] * sn[2 * e + 1]);
      }
      __nv_bfloat16* dst = (hh < hq) ? q_out + tok * hq * d + hh * d : kc + cache_row + (hh - hq) * d;
      *reinterpret_cast<uint4*>(dst + jt * 8) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
      *reinterpret_cast<uint4*>(dst + half + jt * 8) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------ M6 embedding gather
__global__ void svla_embed_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ embed,
                                  const __nv_bfloat16* __restrict__ spatial, const float* __restrict__ img, float* __restrict__ x,
                                  int s, int hdim, long long vocab, long long image_token, long long act_lo, long long n_act,
                                  int n_img, float normalizer, int* __restrict__ status) {
  pdl_launch_dependents();
  pdl_wait();
  const long long tok = blockIdx.x;
  const int b = static_cast<int>(tok / s), si = static_cast<int>(tok % s);
  const long long id = ids[tok];
  float* dst = x + tok * hdim;
  if (img != nullptr && si == s - 1 && status) {
    // the block of a row's last token audits the row: exactly n_img image tokens, or the batch is malformed
    // (model/modeling_spatialvla.py:379-385 raises ValueError; rows with too FEW image tokens would otherwise leave nothing behind)
    int c = 0;
    for (int j = threadIdx.x; j < s; j += blockDim.x) c += (ids[static_cast<long long>(b) * s + j] == image_token);
    __shared__ int total_sh;
    if (threadIdx.x == 0) total_sh = 0;
    __syncthreads();
    if (c) atomicAdd(&total_sh, c);
    __syncthreads();
    if (threadIdx.x == 0 && total_sh != n_img) atomicExch(status, 1);
  }
  if (id == image_token && img != nullptr) {
    // rank of this image token inside its row (the processor puts them first, but stay general)
    __shared__ int rank_sh;
    if (threadIdx.x == 0) rank_sh = 0;
    __syncthreads();
    int cnt = 0;
    for (int j = threadIdx.x; j < si; j += blockDim.x) cnt += (ids[static_cast<long long>(b) * s + j] == image_token);
    if (cnt) atomicAdd(&rank_sh, cnt);
    __syncthreads();
    const int rank = rank_sh;
    if (rank >= n_img) {
      if (threadIdx.x == 0 && status) atomicExch(status, 1);   // image-token / feature count mismatch
      return;
    }
    const float* src = img + (static_cast<long long>(b) * n_img + rank) * hdim;
    for (int i = threadIdx.x; i < hdim; i += blockDim.x) dst[i] = src[i] * normalizer;
    return;
  }
  const __nv_bfloat16* src;
  if (spatial != nullptr && id >= act_lo && id < act_lo + n_act) src = spatial + (id - act_lo) * hdim;
  else {
    if (id < 0 || id >= vocab) {
      if (threadIdx.x == 0 && status) atomicExch(status, 2);
      return;
    }
    src = embed + id * hdim;
  }
  for (int i = threadIdx.x; i < hdim; i += blockDim.x) dst[i] = __bfloat162float(src[i]) * normalizer;
}

// ------------------------------------------------------------------------------------------ M7 argmax
__global__ void __launch_bounds__(256)
svla_argmax_kernel(const float* __restrict__ logits, long long cols, long long ld, long long id_offset,
                   long long* __restrict__ out, long long out_stride) {
  __shared__ float sv[8];
  __shared__ long long si[8];
  pdl_launch_dependents();
  pdl_wait();
  const long long row = blockIdx.x;
  const float* r = logits + row * ld;
  float best = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (long long j = threadIdx.x; j < cols; j += blockDim.x) {
    const float v = r[j];
    if (v > best || (v == best && j < bi)) { best = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    out[row * out_stride] = bi + id_offset;
  }
}

// ------------------------------------------------------------------------------------------ M8 cross entropy over logit rows
// Training / evaluation forward (model/modeling_spatialvla.py:413-430): nn.CrossEntropyLoss over the labelled rows of the
// post-softcap full-vocabulary logits.  One CTA per row (a 265 347-column fp32 row is 1.06 MB): one pass with a per-thread
// online (max, sum exp) pair and the running argmax, 128-bit loads on the 16-byte aligned body of the row (rows of an odd
// vocabulary are not 16-byte aligned: scalar head / tail), four loads in flight per thread and one max update per 16 values
// so that the per-element cost is FSUB + FMUL + ex2.approx + a predicated argmax update (the first version, one expf per
// element and one load in flight, was issue / latency bound at 37 % of HBM peak), warp-shuffle + shared-memory combine.
// HBM-bound: rows * cols * 4 B.
struct CeAcc {
  float m, s, best;      // running max, sum of exp(v - m), best value
  int bi;                // first index of the best value seen by this thread (columns < 2^31)
};
constexpr float kCeLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ce_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp(a - b) with exp(-inf - finite) = 0 and the empty-accumulator case (-inf - -inf) mapped to 0 as well
__device__ __forceinline__ float ce_exp_diff(float a, float b) { return (a == -INFINITY) ? 0.f : ce_ex2((a - b) * kCeLog2e); }
__device__ __forceinline__ void ce_raise_max(CeAcc& a, float cm) {
  if (cm > a.m) {
    a.s *= ce_exp_diff(a.m, cm);
    a.m = cm;
  }
}
__device__ __forceinline__ void ce_add(CeAcc& a, float v, int j) {      // requires v <= a.m
  if (v > a.best) { a.best = v; a.bi = j; }                            // strict: a thread visits its columns in ascending order
  a.s += ce_exp_diff(v, a.m);
}
__device__ __forceinline__ void ce_push(CeAcc& a, float v, int j) {
  ce_raise_max(a, v);
  ce_add(a, v, j);
}
__device__ __forceinline__ void ce_push4(CeAcc& a, const float4& v, int j) {
  ce_add(a, v.x, j);
  ce_add(a, v.y, j + 1);
  ce_add(a, v.z, j + 2);
  ce_add(a, v.w, j + 3);
}
__device__ __forceinline__ void ce_merge(CeAcc& a, float m2, float s2, float b2, int i2) {
  if (b2 > a.best || (b2 == a.best && i2 < a.bi)) { a.best = b2; a.bi = i2; }
  const float M = fmaxf(a.m, m2);
  a.s = a.s * ce_exp_diff(a.m, M) + s2 * ce_exp_diff(m2, M);
  a.m = M;
}

constexpr int kCeThreads = 512;
constexpr int kCeUnroll = 4;      // independent 128-bit loads in flight per thread

__global__ void __launch_bounds__(kCeThreads)
svla_cross_entropy_kernel(const float* __restrict__ logits, long long cols, long long ld, const long long* __restrict__ labels,
                          long long ignore_index, float* __restrict__ row_loss, long long* __restrict__ row_argmax) {
  __shared__ float sm_m[kCeThreads / 32], sm_s[kCeThreads / 32], sm_b[kCeThreads / 32];
  __shared__ int sm_i[kCeThreads / 32];
  const long long row = blockIdx.x;
  const float* r = logits + row * ld;
  const int tid = threadIdx.x;
  CeAcc a{-INFINITY, 0.f, -INFINITY, 0x7fffffff};
  int head = static_cast<int>(((16 - (reinterpret_cast<uintptr_t>(r) & 15)) & 15) >> 2);
  if (head > cols) head = static_cast<int>(cols);
  if (tid < head) ce_push(a, r[tid], tid);
  const int nvec = static_cast<int>((cols - head) >> 2);
  const float4* rv = reinterpret_cast<const float4*>(r + head);
  int i = tid;
  // main loop: kCeUnroll loads issued back to back, ONE max update / rescale per 16 values, then one FSUB + FMUL + MUFU each
  for (; i + (kCeUnroll - 1) * kCeThreads < nvec; i += kCeUnroll * kCeThreads) {
    float4 v[kCeUnroll];
#pragma unroll
    for (int u = 0; u < kCeUnroll; ++u) v[u] = __ldg(rv + i + u * kCeThreads);
    float cm = -INFINITY;
#pragma unroll
    for (int u = 0; u < kCeUnroll; ++u) cm = fmaxf(cm, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
    ce_raise_max(a, cm);
#pragma unroll
    for (int u = 0; u < kCeUnroll; ++u) ce_push4(a, v[u], head + 4 * (i + u * kCeThreads));
  }
  for (; i < nvec; i += kCeThreads) {
    const float4 v = __ldg(rv + i);
    ce_raise_max(a, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    ce_push4(a, v, head + 4 * i);
  }
  const int tail0 = head + 4 * nvec;
  if (tail0 + tid < cols) ce_push(a, r[tail0 + tid], tail0 + tid);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, a.m, o), s2 = __shfl_xor_sync(0xffffffffu, a.s, o);
    const float b2 = __shfl_xor_sync(0xffffffffu, a.best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, a.bi, o);
    ce_merge(a, m2, s2, b2, i2);
  }
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) { sm_m[warp] = a.m; sm_s[warp] = a.s; sm_b[warp] = a.best; sm_i[warp] = a.bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kCeThreads / 32; ++w) ce_merge(a, sm_m[w], sm_s[w], sm_b[w], sm_i[w]);
    const long long lab = labels[row];
    float loss = 0.f;
    if (lab != ignore_index) loss = (lab >= 0 && lab < cols) ? (a.m + logf(a.s)) - r[lab] : __int_as_float(0x7fc00000);
    row_loss[row] = loss;
    row_argmax[row] = a.bi;
  }
}

// Backward of the loss tail (first kernel of the training step's backward half): gradient of the mean cross
// entropy w.r.t. the PRE-soft-cap logits z (logit = cap * tanh(z / cap)), written as the bf16 A operand of the dh = dz * W_head
// GEMM:  dz[r, j] = (softmax(logit[r])[j] - [j == label_r]) * (1 - (logit[r, j] / cap)^2) / count,  0 for ignored rows.
// The row's logsumexp comes from the forward kernel's output (row_loss = lse - logit[label]); count is read from the forward
// summary on the device (no host sync).  Elementwise, HBM-bound: rows * cols * (4 B in + 2 B out).  grid = (column chunks, rows);
// a thread owns adjacent column pairs so the bf16 output leaves as 4-byte stores (128 B per warp instruction).
constexpr int kCeBwdThreads = 256;
constexpr int kCeBwdPairs = 8;      // column pairs per thread -> 4096 columns per CTA

__global__ void __launch_bounds__(kCeBwdThreads)
svla_cross_entropy_bwd_kernel(const float* __restrict__ logits, long long cols, long long ld, const long long* __restrict__ labels,
                              long long ignore_index, const float* __restrict__ row_loss, const float* __restrict__ summary,
                              float softcap, __nv_bfloat16* __restrict__ dz, long long ldo) {
  const long long row = blockIdx.y;
  const float* r = logits + row * ld;
  __nv_bfloat16* o = dz + row * ldo;
  const long long lab = labels[row];
  const bool live = lab != ignore_index && lab >= 0 && lab < cols;
  const float lse = live ? row_loss[row] + r[lab] : 0.f;
  const float inv_count = live ? 1.f / summary[1] : 0.f;
  const float inv_cap2 = softcap > 0.f ? 1.f / (softcap * softcap) : 0.f;
  const long long c0 = static_cast<long long>(blockIdx.x) * (kCeBwdThreads * kCeBwdPairs * 2);
#pragma unroll
  for (int p = 0; p < kCeBwdPairs; ++p) {
    const long long j = c0 + 2LL * (p * kCeBwdThreads + threadIdx.x);
    if (j >= ldo) break;
    float g[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const long long jj = j + e;
      float v = 0.f;
      if (live && jj < cols) {
        const float l = r[jj];
        v = (ce_ex2((l - lse) * kCeLog2e) - (jj == lab ? 1.f : 0.f)) * (1.f - l * l * inv_cap2) * inv_count;
      }
      g[e] = v;                                  // columns in [cols, ldo) are the zero K padding of the GEMM operand
    }
    if (j + 1 < ldo) *reinterpret_cast<__nv_bfloat162*>(o + j) = __floats2bfloat162_rn(g[0], g[1]);
    else o[j] = __float2bfloat16(g[0]);
  }
}

// summary[0] = mean of row_loss over the non-ignored rows (NaN when there is none, like torch), [1] = their count,
// [2] = rows whose argmax equals the label.  One CTA, fixed summation order: deterministic.
__global__ void __launch_bounds__(256)
svla_cross_entropy_summary_kernel(const float* __restrict__ row_loss, const long long* __restrict__ row_argmax,
                                  const long long* __restrict__ labels, long long rows, long long ignore_index,
                                  float* __restrict__ summary) {
  __shared__ float s_sum[256], s_cnt[256], s_hit[256];
  float sum = 0.f, cnt = 0.f, hit = 0.f;
  for (long long i = threadIdx.x; i < rows; i += 256) {
    if (labels[i] != ignore_index) {
      sum += row_loss[i];
      cnt += 1.f;
      hit += (row_argmax[i] == labels[i]) ? 1.f : 0.f;
    }
  }
  s_sum[threadIdx.x] = sum; s_cnt[threadIdx.x] = cnt; s_hit[threadIdx.x] = hit;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
      s_hit[threadIdx.x] += s_hit[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    summary[0] = s_sum[0] / s_cnt[0];
    summary[1] = s_cnt[0];
    summary[2] = s_hit[0];
  }
}

// ------------------------------------------------------------------------------------------ AdamW on the flat LoRA arena
// Optimizer step of the fine-tune config (HF Trainer's default adamw_torch on the ~59.2 M adapter parameters; arithmetic in the
// order of torch.optim.AdamW's single-tensor path: decoupled decay, moment updates, denom = sqrt(v) / sqrt(1 - b2^t) + eps,
// p -= (lr / (1 - b1^t)) * m / denom).  One pass over the flat buffers, 128-bit accesses: 16 B read + 12 B written per element.
__global__ void __launch_bounds__(256)
svla_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                  float decay, float omb1, float b2, float omb2, float eps, float step_size, float bc2_sqrt, float grad_scale,
                  const float* __restrict__ grad_sumsq, float max_norm) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  if (grad_sumsq) {
    // torch.nn.utils.clip_grad_norm_ on the scaled (rank-averaged) gradient: coef = min(1, max_norm / (||g|| + 1e-6)), on the device
    const float total = sqrtf(*grad_sumsq) * grad_scale;
    grad_scale *= fminf(1.f, max_norm / (total + 1e-6f));
  }
  // every scalar (decay = 1 - lr wd, omb1 = 1 - beta1, omb2 = 1 - beta2, step_size = lr / (1 - beta1^t), bc2_sqrt) is formed in
  // DOUBLE on the host like torch does with its Python floats: 1.f - 0.999f alone is off by 1.3e-5 of (1 - beta2)
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    pp *= decay;
    mm = mm + (gg - mm) * omb1;                   // lerp form of torch: m.lerp_(g, 1 - beta1)
    vv = vv * b2 + gg * gg * omb2;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * (mm / denom);
  };
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    upd(pv.x, gv.x, mv.x, vv.x);
    upd(pv.y, gv.y, mv.y, vv.y);
    upd(pv.z, gv.z, mv.z, vv.z);
    upd(pv.w, gv.w, mv.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (nvec << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
    upd(p[i], g[i], m[i], v[i]);
}

// sum of squares of a flat fp32 buffer -> *out += (fp32 atomics over per-block partial sums in double)
__global__ void __launch_bounds__(256)
svla_sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float sh[33];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (nvec << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) acc += x[i] * x[i];
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, t);
}

// ------------------------------------------------------------------------------------------ bicubic helpers (A = -0.75)
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x1 = t, x2 = 1.f - t;
  w[0] = ((A * (x1 + 1.f) - 5.f * A) * (x1 + 1.f) + 8.f * A) * (x1 + 1.f) - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * (x2 + 1.f) - 5.f * A) * (x2 + 1.f) + 8.f * A) * (x2 + 1.f) - 4.f * A;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflect224(int p) {   // index into the 286-wide reflect-padded image -> 224 image
  int i = p - 31;
  if (i < 0) i = -i;
  if (i > 223) i = 446 - i;
  return i;
}

// ------------------------------------------------------------------------------------------ M9 patchify
// siglip: a[(b*256 + py*16+px), c*196 + i*14 + j] = (px[b,c,py*14+i,px*14+j] - .5)/.5 ; cols [588,kpad) = 0
__global__ void svla_siglip_patchify_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ a, int kpad) {
  const int row = blockIdx.x;            // b*256 + patch
  const int b = row >> 8, patch = row & 255, py = patch >> 4, pxx = patch & 15;
  for (int col = threadIdx.x; col < kpad; col += blockDim.x) {
    float v = 0.f;
    if (col < 588) {
      const int c = col / 196, r = col % 196, i = r / 14, j = r % 14;
      v = (px[((static_cast<long long>(b) * 3 + c) * 224 + py * 14 + i) * 224 + pxx * 14 + j] - 0.5f) / 0.5f;
    }
    a[static_cast<long long>(row) * kpad + col] = __float2bfloat16(v);
  }
}

// zoe: 224 -(reflect pad 31)-> 286 -(bicubic, align_corners)-> 384 -> normalise -> im2col(16x16):
// a[(b*576 + py*24+px), c*256 + i*16 + j]
__global__ void svla_zoe_patchify_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ a) {
  const int row = blockIdx.x;
  const int b = row / 576, patch = row % 576, py = patch / 24, pxx = patch % 24;
  const float scale = 285.0f / 383.0f;
  for (int col = threadIdx.x; col < 768; col += blockDim.x) {
    const int c = col >> 8, i = (col >> 4) & 15, j = col & 15;
    const int oy = py * 16 + i, ox = pxx * 16 + j;
    const float sy = scale * oy, sx = scale * ox;
    const int y0 = static_cast<int>(floorf(sy)), x0 = static_cast<int>(floorf(sx));
    float wy[4], wx[4];
    cubic_coeffs(sy - y0, wy);
    cubic_coeffs(sx - x0, wx);
    const float* img = px + (static_cast<long long>(b) * 3 + c) * 224 * 224;
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int yy = reflect224(clampi(y0 - 1 + u, 0, 285));
      float racc = 0.f;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int xx = reflect224(clampi(x0 - 1 + v, 0, 285));
        racc += wx[v] * img[yy * 224 + xx];
      }
      acc += wy[u] * racc;
    }
    a[static_cast<long long>(row) * 768 + col] = __float2bfloat16((acc - 0.5f) / 0.5f);
  }
}

__global__ void svla_beit_assemble_kernel(const float* __restrict__ patches, const float* __restrict__ cls, float* __restrict__ x,
                                          int n, int c) {
  const long long row = blockIdx.x;   // b*(n+1) + t
  const int t = static_cast<int>(row % (n + 1));
  const long long b = row / (n + 1);
  const float* src = (t == 0) ? cls : patches + (b * n + (t - 1)) * c;
  for (int i = threadIdx.x; i < c; i += blockDim.x) x[row * c + i] = src[i];
}

__global__ void svla_readout_concat_kernel(const float* __restrict__ hs, __nv_bfloat16* __restrict__ a, int n, int c) {
  const long long row = blockIdx.x;   // b*n + i
  const long long b = row / n;
  const int i = static_cast<int>(row % n);
  const float* tok = hs + (b * (n + 1) + 1 + i) * c;
  const float* cls = hs + (b * (n + 1)) * c;
  for (int j = threadIdx.x; j < c; j += blockDim.x) {
    a[row * 2 * c + j] = __float2bfloat16(tok[j]);
    a[row * 2 * c + c + j] = __float2bfloat16(cls[j]);
  }
}

// g [B*h*w, f*f*c] -> out NHWC [B, h*f, w*f, c]; 8 bf16 (16 bytes) per thread
__global__ void svla_pixel_shuffle_kernel(const uint4* __restrict__ g, uint4* __restrict__ out, int h, int w, int c8, int f, long long total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int cc = static_cast<int>(idx % c8);
  long long r = idx / c8;
  const int ox = static_cast<int>(r % (w * f)); r /= (w * f);
  const int oy = static_cast<int>(r % (h * f));
  const long long b = r / (h * f);
  const int y = oy / f, i = oy % f, x = ox / f, j = ox % f;
  out[idx] = g[((b * h + y) * w + x) * (static_cast<long long>(f) * f * c8) + (i * f + j) * c8 + cc];
}

__global__ void svla_im2col3x3_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ a, int h, int w, int c8, long long total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int cc = static_cast<int>(idx % c8);
  long long r = idx / c8;
  const int tap = static_cast<int>(r % 9); r /= 9;
  const int oh = h / 2, ow = w / 2;
  const int ox = static_cast<int>(r % ow); r /= ow;
  const int oy = static_cast<int>(r % oh);
  const long long b = r / oh;
  const int y = oy * 2 + tap / 3 - 1, xx = ox * 2 + tap % 3 - 1;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (y >= 0 && y < h && xx >= 0 && xx < w) v = x[((b * h + y) * w + xx) * c8 + cc];
  a[idx] = v;
}

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) { f[2 * u] = bf16_bits_to_float(w[u] & 0xFFFFu); f[2 * u + 1] = bf16_bits_to_float(w[u] >> 16); }
}

// bilinear, align_corners=True, NHWC bf16, 8 channels per thread.  grid = (chunks of an output row, oy, batch): the
// row / batch coordinates come from the block index, so the per-thread index math is one 32-bit division.
__global__ void __launch_bounds__(256)
svla_bilinear_nhwc_kernel(const uint4* __restrict__ x, const uint4* __restrict__ add, uint4* __restrict__ out,
                          uint4* __restrict__ out_relu, int h, int w, int c8, int oh, int ow) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;           // position inside the output row, in uint4 units
  if (i >= static_cast<unsigned>(ow) * c8) return;
  const int ox = static_cast<int>(i / static_cast<unsigned>(c8));
  const int cc = static_cast<int>(i - static_cast<unsigned>(ox) * c8);
  const int oy = blockIdx.y;
  const long long b = blockIdx.z;
  const float sy = (oh > 1) ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) * oy : 0.f;
  const float sx = (ow > 1) ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) * ox : 0.f;
  const int y0 = min(static_cast<int>(sy), h - 1), x0 = min(static_cast<int>(sx), w - 1);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = sy - y0, lx = sx - x0;
  const uint4* r0 = x + (b * h + y0) * static_cast<long long>(w) * c8 + cc;
  const uint4* r1 = x + (b * h + y1) * static_cast<long long>(w) * c8 + cc;
  const uint4 q00 = r0[x0 * c8], q01 = r0[x1 * c8], q10 = r1[x0 * c8], q11 = r1[x1 * c8];
  const long long idx = (b * oh + oy) * static_cast<long long>(ow) * c8 + i;
  float f00[8], f01[8], f10[8], f11[8], o[8];
  unpack8(q00, f00); unpack8(q01, f01); unpack8(q10, f10); unpack8(q11, f11);
#pragma unroll
  for (int e = 0; e < 8; ++e)
    o[e] = (1.f - ly) * ((1.f - lx) * f00[e] + lx * f01[e]) + ly * ((1.f - lx) * f10[e] + lx * f11[e]);
  if (add) {
    float ad[8];
    unpack8(add[idx], ad);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] += ad[e];
  }
  if (out) out[idx] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  if (out_relu)
    out_relu[idx] = make_uint4(pack_bf16x2(fmaxf(o[0], 0.f), fmaxf(o[1], 0.f)), pack_bf16x2(fmaxf(o[2], 0.f), fmaxf(o[3], 0.f)),
                               pack_bf16x2(fmaxf(o[4], 0.f), fmaxf(o[5], 0.f)), pack_bf16x2(fmaxf(o[6], 0.f), fmaxf(o[7], 0.f)));
}

// Tiled bilinear (align_corners=True) for the x2 up-samplings of the DPT neck / relative head / bin embeddings.  ncu on the
// per-pixel kernel above: 33 thread-instructions per output element, issue-bound at 37 % of HBM.  Here a block produces
// 8 output rows x 32 output columns x 64 channels in two passes: (1) the <= floor(7 ry) + 3 source rows it needs are
// interpolated horizontally ONCE into shared memory (fp32), (2) every output is one vertical lerp of two shared-memory
// rows.  Same association as the reference: (1-ly) * ((1-lx) v00 + lx v01) + ly * ((1-lx) v10 + lx v11).
constexpr int kBlRows = 8, kBlCols = 32, kBlGroups = 8;     // output rows, output columns, 8-channel groups per block
__global__ void __launch_bounds__(256)
svla_bilinear_tile_kernel(const uint4* __restrict__ x, const uint4* __restrict__ add, uint4* __restrict__ out,
                          uint4* __restrict__ out_relu, int h, int w, int c8, int oh, int ow, int nrows, int cgroups) {
  extern __shared__ __align__(16) float4 s_bl[];
  float4* s_lo = s_bl;                                            // [nrows][32][8] channels 0-3 of the group
  float4* s_hi = s_bl + nrows * kBlCols * kBlGroups;              //                 channels 4-7
  const int ox0 = blockIdx.x * kBlCols, oy0 = blockIdx.y * kBlRows;
  const int cgb = (blockIdx.z % cgroups) * kBlGroups;
  const long long b = blockIdx.z / cgroups;
  const float ry = (oh > 1) ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.f;
  const float rx = (ow > 1) ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) : 0.f;
  const int ybase = min(static_cast<int>(ry * oy0), h - 1);
  // ---- pass 1: horizontal interpolation of the source rows [ybase, ybase + nrows)
  for (int item = threadIdx.x; item < nrows * kBlCols * kBlGroups; item += 256) {
    const int cg = item & 7, col = (item >> 3) & 31, r = item >> 8;
    const int ox = ox0 + col, y = ybase + r, c = cgb + cg;
    if (ox < ow && y < h && c < c8) {
      const float sx = rx * ox;
      const int x0 = min(static_cast<int>(sx), w - 1), x1 = min(x0 + 1, w - 1);
      const float lx = sx - x0;
      const uint4* rowp = x + (b * h + y) * static_cast<long long>(w) * c8 + c;
      float f0[8], f1[8];
      unpack8(rowp[static_cast<long long>(x0) * c8], f0);
      unpack8(rowp[static_cast<long long>(x1) * c8], f1);
      float hv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) hv[e] = (1.f - lx) * f0[e] + lx * f1[e];
      s_lo[item] = make_float4(hv[0], hv[1], hv[2], hv[3]);
      s_hi[item] = make_float4(hv[4], hv[5], hv[6], hv[7]);
    }
  }
  __syncthreads();
  // ---- pass 2: vertical interpolation, optional add, bf16 (and ReLU copy) stores
  for (int item = threadIdx.x; item < kBlRows * kBlCols * kBlGroups; item += 256) {
    const int cg = item & 7, col = (item >> 3) & 31, rr = item >> 8;
    const int ox = ox0 + col, oy = oy0 + rr, c = cgb + cg;
    if (ox >= ow || oy >= oh || c >= c8) continue;
    const float sy = ry * oy;
    const int y0 = min(static_cast<int>(sy), h - 1), y1 = min(y0 + 1, h - 1);
    const float ly = sy - y0;
    const int i0 = ((y0 - ybase) * kBlCols + col) * kBlGroups + cg, i1 = ((y1 - ybase) * kBlCols + col) * kBlGroups + cg;
    const float4 a0 = s_lo[i0], a1 = s_hi[i0], b0 = s_lo[i1], b1 = s_hi[i1];
    float o[8];
    o[0] = (1.f - ly) * a0.x + ly * b0.x; o[1] = (1.f - ly) * a0.y + ly * b0.y;
    o[2] = (1.f - ly) * a0.z + ly * b0.z; o[3] = (1.f - ly) * a0.w + ly * b0.w;
    o[4] = (1.f - ly) * a1.x + ly * b1.x; o[5] = (1.f - ly) * a1.y + ly * b1.y;
    o[6] = (1.f - ly) * a1.z + ly * b1.z; o[7] = (1.f - ly) * a1.w + ly * b1.w;
    const long long idx = ((b * oh + oy) * static_cast<long long>(ow) + ox) * c8 + c;
    if (add) {
      float ad[8];
      unpack8(add[idx], ad);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += ad[e];
    }
    if (out) out[idx] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    if (out_relu)
      out_relu[idx] = make_uint4(pack_bf16x2(fmaxf(o[0], 0.f), fmaxf(o[1], 0.f)), pack_bf16x2(fmaxf(o[2], 0.f), fmaxf(o[3], 0.f)),
                                 pack_bf16x2(fmaxf(o[4], 0.f), fmaxf(o[5], 0.f)), pack_bf16x2(fmaxf(o[6], 0.f), fmaxf(o[7], 0.f)));
  }
}

__global__ void svla_relu_bf16_kernel(const __nv_bfloat162* __restrict__ x, __nv_bfloat162* __restrict__ out, long long n2) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n2) return;
  const float2 f = __bfloat1622float2(x[idx]);
  out[idx] = __floats2bfloat162_rn(fmaxf(f.x, 0.f), fmaxf(f.y, 0.f));
}

__global__ void svla_softplus_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long n) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx < n) out[idx] = softplus_f(__bfloat162float(x[idx]));
}

// ZoeDepth patch transformer input: PE = cat(sin(pos*div), cos(pos*div)), div_i = exp(2i * (-ln(1e4)/c))
__global__ void svla_zoe_router_embed_kernel(const float* __restrict__ conv, float* __restrict__ e, __nv_bfloat16* __restrict__ eb, int n, int c) {
  const long long row = blockIdx.x;   // b*(n+1) + t
  const int t = static_cast<int>(row % (n + 1));
  const long long b = row / (n + 1);
  const int half = c / 2;
  const float k = -logf(10000.0f) / static_cast<float>(c);
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    const int fi = (i < half) ? i : i - half;
    const float ang = static_cast<float>(t) * expf(static_cast<float>(2 * fi) * k);
    const float pe = (i < half) ? sinf(ang) : cosf(ang);
    const float base = (t == 0) ? 0.f : conv[(b * n + (t - 1)) * c + i];
    e[row * c + i] = base + pe;
    if (eb) eb[row * c + i] = __float2bfloat16(base + pe);
  }
}

// bilinear (align_corners=True) sample of an fp32 NHWC map: channel ch of pixel (oy, ox) of an oh x ow grid
struct Bilin {
  int y0, y1, x0, x1;
  float ly, lx;
};
__device__ __forceinline__ Bilin make_bilin(int oy, int ox, int h, int w, int oh, int ow) {
  Bilin s;
  const float sy = (oh > 1) ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) * oy : 0.f;
  const float sx = (ow > 1) ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) * ox : 0.f;
  s.y0 = min(static_cast<int>(sy), h - 1); s.x0 = min(static_cast<int>(sx), w - 1);
  s.y1 = min(s.y0 + 1, h - 1); s.x1 = min(s.x0 + 1, w - 1);
  s.ly = sy - s.y0; s.lx = sx - s.x0;
  return s;
}
__device__ __forceinline__ float bilin_f32(const float* __restrict__ base, const Bilin& s, int w, int c, int ch) {
  const float v00 = base[(static_cast<long long>(s.y0) * w + s.x0) * c + ch], v01 = base[(static_cast<long long>(s.y0) * w + s.x1) * c + ch];
  const float v10 = base[(static_cast<long long>(s.y1) * w + s.x0) * c + ch], v11 = base[(static_cast<long long>(s.y1) * w + s.x1) * c + ch];
  return (1.f - s.ly) * ((1.f - s.lx) * v00 + s.lx * v01) + s.ly * ((1.f - s.lx) * v10 + s.lx * v11);
}
__device__ __forceinline__ float bilin_bf16(const __nv_bfloat16* __restrict__ base, const Bilin& s, int w, int c, int ch) {
  const float v00 = __bfloat162float(base[(static_cast<long long>(s.y0) * w + s.x0) * c + ch]);
  const float v01 = __bfloat162float(base[(static_cast<long long>(s.y0) * w + s.x1) * c + ch]);
  const float v10 = __bfloat162float(base[(static_cast<long long>(s.y1) * w + s.x0) * c + ch]);
  const float v11 = __bfloat162float(base[(static_cast<long long>(s.y1) * w + s.x1) * c + ch]);
  return (1.f - s.ly) * ((1.f - s.lx) * v00 + s.lx * v01) + s.ly * ((1.f - s.lx) * v10 + s.lx * v11);
}

// Attractor step of one metric-bins stage.  16 lanes per output pixel, one float4 (4 bins) per lane and pass; a block is 16
// consecutive pixels of one output row, (row, batch) come from the block index (no per-pixel integer division).
__global__ void __launch_bounds__(256)
svla_zoe_attractor_kernel(const __nv_bfloat16* __restrict__ attr, const float* __restrict__ prev, float* __restrict__ out, int h,
                          int w, int oh, int ow, int na, int nbins) {
  const int lane = threadIdx.x & 31, sub = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const int ox_raw = blockIdx.x * 16 + grp, oy = blockIdx.y;
  const bool active = ox_raw < ow;
  const int ox = active ? ox_raw : ow - 1;                    // idle groups shadow the last pixel (shuffles stay warp-wide)
  const long long b = blockIdx.z;
  const long long pix = (b * oh + oy) * ow + ox;
  const float a_l = (sub < na) ? softplus_f(__bfloat162float(attr[pix * na + sub])) : 0.f;
  const Bilin s = make_bilin(oy, ox, h, w, oh, ow);
  const int nb4 = nbins >> 2;
  const float4* pb = reinterpret_cast<const float4*>(prev + b * h * w * nbins);
  const float4* p00 = pb + (static_cast<long long>(s.y0) * w + s.x0) * nb4;
  const float4* p01 = pb + (static_cast<long long>(s.y0) * w + s.x1) * nb4;
  const float4* p10 = pb + (static_cast<long long>(s.y1) * w + s.x0) * nb4;
  const float4* p11 = pb + (static_cast<long long>(s.y1) * w + s.x1) * nb4;
  const float inv_na = 1.f / static_cast<float>(na);
  float4* ob = reinterpret_cast<float4*>(out + pix * nbins);
  for (int base = 0; base < nb4; base += 16) {
    const int k4 = base + sub;
    const bool valid = k4 < nb4;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const float4 v00 = p00[k4], v01 = p01[k4], v10 = p10[k4], v11 = p11[k4];
      // same association as bilin_f32: (1-ly)*((1-lx) v00 + lx v01) + ly*((1-lx) v10 + lx v11)
      c.x = (1.f - s.ly) * ((1.f - s.lx) * v00.x + s.lx * v01.x) + s.ly * ((1.f - s.lx) * v10.x + s.lx * v11.x);
      c.y = (1.f - s.ly) * ((1.f - s.lx) * v00.y + s.lx * v01.y) + s.ly * ((1.f - s.lx) * v10.y + s.lx * v11.y);
      c.z = (1.f - s.ly) * ((1.f - s.lx) * v00.z + s.lx * v01.z) + s.ly * ((1.f - s.lx) * v10.z + s.lx * v11.z);
      c.w = (1.f - s.ly) * ((1.f - s.lx) * v00.w + s.lx * v01.w) + s.ly * ((1.f - s.lx) * v10.w + s.lx * v11.w);
    }
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    // dc = dx / (1 + 300 dx^2).  The loop is MUFU-bound with one reciprocal per term, so two attractors share one:
    // dx1/b1 + dx2/b2 = (dx1 b2 + dx2 b1) / (b1 b2)  (b <= 1 + 300 * 80^2: the product stays far inside fp32 range)
    // packed f32x2 form (two bins per FFMA2 / FMUL2, same roundings as the scalar expression): the loop was FMA-issue bound
    auto pair_term2 = [](uint64_t c2, uint64_t a1, uint64_t a2, uint64_t acc) {
      using namespace svla_ptx;
      const uint64_t m1 = pack_f32x2(-1.f, -1.f), k300 = pack_f32x2(300.f, 300.f), one = pack_f32x2(1.f, 1.f);
      const uint64_t d1 = fma_f32x2(c2, m1, a1), d2 = fma_f32x2(c2, m1, a2);                 // a - c (exact: one rounding, like a1 - c0)
      const uint64_t b1 = fma_f32x2(mul_f32x2(k300, d1), d1, one), b2 = fma_f32x2(mul_f32x2(k300, d2), d2, one);
      const uint64_t num = fma_f32x2(d1, b2, mul_f32x2(d2, b1)), den = mul_f32x2(b1, b2);
      float n0, n1, e0, e1;
      unpack_f32x2(num, n0, n1);
      unpack_f32x2(den, e0, e1);
      return add_f32x2(acc, pack_f32x2(__fdividef(n0, e0), __fdividef(n1, e1)));
    };
    uint64_t dxy = svla_ptx::pack_f32x2(0.f, 0.f), dzw = svla_ptx::pack_f32x2(0.f, 0.f);
    const uint64_t cxy = svla_ptx::pack_f32x2(c.x, c.y), czw = svla_ptx::pack_f32x2(c.z, c.w);
    int a = 0;
    for (; a + 1 < na; a += 2) {
      const float a1 = __shfl_sync(0xffffffffu, a_l, (lane & 16) | a), a2 = __shfl_sync(0xffffffffu, a_l, (lane & 16) | (a + 1));
      const uint64_t a1p = svla_ptx::pack_f32x2(a1, a1), a2p = svla_ptx::pack_f32x2(a2, a2);
      dxy = pair_term2(cxy, a1p, a2p, dxy);
      dzw = pair_term2(czw, a1p, a2p, dzw);
    }
    svla_ptx::unpack_f32x2(dxy, d.x, d.y);
    svla_ptx::unpack_f32x2(dzw, d.z, d.w);
    if (a < na) {
      const float av = __shfl_sync(0xffffffffu, a_l, (lane & 16) | a);
      const float dx = av - c.x, dy = av - c.y, dz = av - c.z, dw = av - c.w;
      d.x += __fdividef(dx, fmaf(300.f * dx, dx, 1.f));
      d.y += __fdividef(dy, fmaf(300.f * dy, dy, 1.f));
      d.z += __fdividef(dz, fmaf(300.f * dz, dz, 1.f));
      d.w += __fdividef(dw, fmaf(300.f * dw, dw, 1.f));
    }
    if (valid && active) ob[k4] = make_float4(c.x + d.x * inv_na, c.y + d.y * inv_na, c.z + d.z * inv_na, c.w + d.w * inv_na);
  }
}

// Conditional log-binomial tail: 4 lanes per output pixel (each lane: a quarter of the hidden channels, then a quarter of the
// bins), a block is 64 consecutive pixels of one output row.  Two cheap passes over the bins (max of the logits needs no
// loads, the second pass reads the 4 bilinear taps as float4) replace the 8 warp-wide reductions of a warp-per-pixel map.
__global__ void __launch_bounds__(256)
svla_zoe_depth_tail_kernel(const __nv_bfloat16* __restrict__ t, const __nv_bfloat16* __restrict__ e, const float* __restrict__ b1,
                           const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ bins,
                           float* __restrict__ depth, int h, int w, int oh, int ow, int nh, int nbins, float min_temp,
                           float max_temp) {
  // per-block tables: log C(K-1, k) in the reference's Stirling form (HF zoedepth log_binom, eps = 1e-7), MLP weights
  __shared__ float s_lb[64];
  __shared__ float s_w2[4 * 64];
  __shared__ float s_b1[64];
  const float nn = static_cast<float>(nbins - 1) + 1e-7f;
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
    const float kk = static_cast<float>(k) + 1e-7f;
    s_lb[k] = nn * logf(nn) - kk * logf(kk) - (nn - kk) * logf(nn - kk + 1e-7f);
  }
  for (int i = threadIdx.x; i < 4 * nh; i += blockDim.x) s_w2[i] = w2[i];
  for (int i = threadIdx.x; i < nh; i += blockDim.x) s_b1[i] = b1[i];
  __syncthreads();
  const int sub = threadIdx.x & 3, grp = threadIdx.x >> 2;
  const int ox_raw = blockIdx.x * 64 + grp, oy = blockIdx.y;
  const bool active = ox_raw < ow;
  const int ox = active ? ox_raw : ow - 1;
  const long long b = blockIdx.z;
  const long long pix = (b * oh + oy) * ow + ox;
  const Bilin s = make_bilin(oy, ox, h, w, oh, ow);
  const long long o00 = static_cast<long long>(s.y0) * w + s.x0, o01 = static_cast<long long>(s.y0) * w + s.x1;
  const long long o10 = static_cast<long long>(s.y1) * w + s.x0, o11 = static_cast<long long>(s.y1) * w + s.x1;
  // ---- hidden = gelu(W_a*last + up(W_b*emb) + b1) -> 4 outputs (this lane: channels [ch0, ch1))
  float o4[4] = {0.f, 0.f, 0.f, 0.f};
  const int cpl = (nh + 3) >> 2;
  const int ch0 = sub * cpl, ch1 = min(nh, ch0 + cpl);
  const __nv_bfloat16* eb = e + b * h * w * nh;
  const __nv_bfloat16* tp = t + pix * nh;
  if (((nh | cpl) & 1) == 0) {
    for (int ch = ch0; ch < ch1; ch += 2) {
      const float2 tv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(tp + ch));
      const float2 v00 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(eb + o00 * nh + ch));
      const float2 v01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(eb + o01 * nh + ch));
      const float2 v10 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(eb + o10 * nh + ch));
      const float2 v11 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(eb + o11 * nh + ch));
      const float ex = (1.f - s.ly) * ((1.f - s.lx) * v00.x + s.lx * v01.x) + s.ly * ((1.f - s.lx) * v10.x + s.lx * v11.x);
      const float ey = (1.f - s.ly) * ((1.f - s.lx) * v00.y + s.lx * v01.y) + s.ly * ((1.f - s.lx) * v10.y + s.lx * v11.y);
      const float hx = gelu_erf_fast(tv.x + ex + s_b1[ch]), hy = gelu_erf_fast(tv.y + ey + s_b1[ch + 1]);
#pragma unroll
      for (int q = 0; q < 4; ++q) o4[q] += s_w2[q * nh + ch] * hx + s_w2[q * nh + ch + 1] * hy;
    }
  } else {
    for (int ch = ch0; ch < ch1; ++ch) {
      const float hv = gelu_erf_fast(__bfloat162float(tp[ch]) + bilin_bf16(eb, s, w, nh, ch) + s_b1[ch]);
#pragma unroll
      for (int q = 0; q < 4; ++q) o4[q] += s_w2[q * nh + ch] * hv;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    o4[q] += __shfl_xor_sync(0xffffffffu, o4[q], 1);
    o4[q] += __shfl_xor_sync(0xffffffffu, o4[q], 2);
    o4[q] = softplus_fast(o4[q] + b2[q]);
  }
  const float p0 = o4[0] + 1e-4f, p1 = o4[1] + 1e-4f, t0 = o4[2] + 1e-4f, t1 = o4[3] + 1e-4f;
  const float prob = p0 / (p0 + p1);
  const float temp = (max_temp - min_temp) * (t0 / (t0 + t1)) + min_temp;
  const float lp = __logf(fminf(fmaxf(prob, 1e-4f), 1.f)), lq = __logf(fminf(fmaxf(1.f - prob, 1e-4f), 1.f));
  const float inv_temp = 1.f / temp;
  // ---- y_k = (log C(K-1, k) + k log p + (K-1-k) log(1-p)) / temp; softmax over k; depth = sum_k softmax_k * centre_k
  const int bpl = (nbins + 3) >> 2;
  const int k0 = sub * bpl, k1 = min(nbins, k0 + bpl);
  float mx = -INFINITY;
  for (int k = k0; k < k1; ++k)
    mx = fmaxf(mx, (s_lb[k] + static_cast<float>(k) * lp + static_cast<float>(nbins - 1 - k) * lq) * inv_temp);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float se = 0.f, sc = 0.f;
  const float* bb = bins + b * h * w * nbins;
  if (((nbins | bpl) & 3) == 0) {
    const float4* q00 = reinterpret_cast<const float4*>(bb + o00 * nbins), *q01 = reinterpret_cast<const float4*>(bb + o01 * nbins);
    const float4* q10 = reinterpret_cast<const float4*>(bb + o10 * nbins), *q11 = reinterpret_cast<const float4*>(bb + o11 * nbins);
#pragma unroll 4
    for (int k = k0; k < k1; k += 4) {
      const float4 v00 = q00[k >> 2], v01 = q01[k >> 2], v10 = q10[k >> 2], v11 = q11[k >> 2];
      const float c0 = (1.f - s.ly) * ((1.f - s.lx) * v00.x + s.lx * v01.x) + s.ly * ((1.f - s.lx) * v10.x + s.lx * v11.x);
      const float c1 = (1.f - s.ly) * ((1.f - s.lx) * v00.y + s.lx * v01.y) + s.ly * ((1.f - s.lx) * v10.y + s.lx * v11.y);
      const float c2 = (1.f - s.ly) * ((1.f - s.lx) * v00.z + s.lx * v01.z) + s.ly * ((1.f - s.lx) * v10.z + s.lx * v11.z);
      const float c3 = (1.f - s.ly) * ((1.f - s.lx) * v00.w + s.lx * v01.w) + s.ly * ((1.f - s.lx) * v10.w + s.lx * v11.w);
      const float cs[4] = {c0, c1, c2, c3};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = k + i;
        const float y = (s_lb[kk] + static_cast<float>(kk) * lp + static_cast<float>(nbins - 1 - kk) * lq) * inv_temp;
        const float ex = __expf(y - mx);
        se += ex; sc += ex * cs[i];
      }
    }
  } else {
    for (int k = k0; k < k1; ++k) {
      const float y = (s_lb[k] + static_cast<float>(k) * lp + static_cast<float>(nbins - 1 - k) * lq) * inv_temp;
      const float ex = __expf(y - mx);
      se += ex; sc += ex * bilin_f32(bb, s, w, nbins, k);
    }
  }
  se += __shfl_xor_sync(0xffffffffu, se, 1); se += __shfl_xor_sync(0xffffffffu, se, 2);
  sc += __shfl_xor_sync(0xffffffffu, sc, 1); sc += __shfl_xor_sync(0xffffffffu, sc, 2);
  if (sub == 0 && active) depth[pix] = sc / se;
}

// Fused conditional-log-binomial tail (replaces GEMM [9.4 M x 32] x [32 -> 40] + svla_zoe_depth_tail_kernel): one thread per
// output pixel, a block = 16 x 16 output pixels.  The half-resolution operands every pixel of the tile interpolates (W_b*emb:
// NH bf16 channels, attractor bins: NBINS fp32) are staged ONCE in shared memory (<= 12 x 12 source pixels, padded pixel strides
// keep the float4 tap reads conflict-free); the first MLP layer W_a*x on the 32 relative-head features runs on the FMA pipe with
// warp-broadcast weight reads, so its [pixels, 40] output never exists in HBM (755 MB written + read per 64-image batch before).
// ncu on the previous kernel: 1.86 G warp-instructions (4 lanes per pixel re-doing the tap setup, 64-bit global tap addressing),
// issue-bound at 11 % of HBM.
constexpr int kTailTile = 16;
constexpr int kTailMaxSrc = 12;          // source rows / columns a tile may touch (x2 up-sampling needs 10)
constexpr int kTailTStride = 44;         // floats per pixel of the staged first-layer output (40 + 4: conflict-free float4 rows)

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NX, int NH, int NBINS>
__global__ void __launch_bounds__(256)
svla_zoe_depth_tail_fused_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wa,
                                 const __nv_bfloat16* __restrict__ e, const float* __restrict__ b1, const float* __restrict__ w2,
                                 const float* __restrict__ b2, const float* __restrict__ bins, float* __restrict__ depth, int h, int w,
                                 int oh, int ow, int nr, int nc, float min_temp, float max_temp) {
  static_assert(NX == 32 && NH % 8 == 0 && NBINS % 4 == 0, "tile shapes of the mma.sync first layer");
  constexpr int EST = NH * 2;            // bytes per staged e pixel (NH = 40 -> 80 B: 8 consecutive pixels hit distinct banks)
  constexpr int BST = NBINS + 4;         // floats per staged bins pixel
  extern __shared__ __align__(16) uint8_t s_tail[];
  float* s_bins = reinterpret_cast<float*>(s_tail);                                 // [nr*nc][BST]
  uint8_t* s_e = s_tail + static_cast<size_t>(kTailMaxSrc) * kTailMaxSrc * BST * 4;  // [nr*nc][EST]
  float* s_t = reinterpret_cast<float*>(s_e + kTailMaxSrc * kTailMaxSrc * EST);     // [256 pixels][kTailTStride] first-layer output
  uint32_t* s_wa = reinterpret_cast<uint32_t*>(s_t + 256 * kTailTStride);           // [NH][NX/2] bf16 pairs (B operand, K-major)
  float* s_w2 = reinterpret_cast<float*>(s_wa + NH * NX / 2);                       // [NH][4]
  float* s_b1 = s_w2 + 4 * NH;                                                      // [NH]
  float* s_lb = s_b1 + NH;                                                          // [NBINS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ox0 = blockIdx.x * kTailTile, oy0 = blockIdx.y * kTailTile;
  const long long b = blockIdx.z;
  const float ry = (oh > 1) ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.f;
  const float rx = (ow > 1) ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) : 0.f;
  const int ybase = min(static_cast<int>(ry * oy0), h - 1), xbase = min(static_cast<int>(rx * ox0), w - 1);
  // ---- stage the tables and the source tiles
  const float nn = static_cast<float>(NBINS - 1) + 1e-7f;
  for (int k = tid; k < NBINS; k += 256) {
    const float kk = static_cast<float>(k) + 1e-7f;
    s_lb[k] = nn * logf(nn) - kk * logf(kk) - (nn - kk) * logf(nn - kk + 1e-7f);
  }
  for (int i = tid; i < NH * NX / 2; i += 256) s_wa[i] = reinterpret_cast<const uint32_t*>(wa)[i];
  for (int i = tid; i < 4 * NH; i += 256) s_w2[(i % NH) * 4 + i / NH] = w2[i];       // [q][c] -> [c][q]
  for (int i = tid; i < NH; i += 256) s_b1[i] = b1[i];
  {
    const float4* gb = reinterpret_cast<const float4*>(bins + b * h * w * NBINS);
    for (int i = tid; i < nr * nc * (NBINS / 4); i += 256) {
      const int p = i / (NBINS / 4), q = i - p * (NBINS / 4);
      const int sy = min(ybase + p / nc, h - 1), sx = min(xbase + p % nc, w - 1);
      *reinterpret_cast<float4*>(s_bins + p * BST + 4 * q) = gb[(static_cast<long long>(sy) * w + sx) * (NBINS / 4) + q];
    }
    const uint4* ge = reinterpret_cast<const uint4*>(e + b * h * w * NH);
    for (int i = tid; i < nr * nc * (NH / 8); i += 256) {
      const int p = i / (NH / 8), q = i - p * (NH / 8);
      const int sy = min(ybase + p / nc, h - 1), sx = min(xbase + p % nc, w - 1);
      *reinterpret_cast<uint4*>(s_e + p * EST + 16 * q) = ge[(static_cast<long long>(sy) * w + sx) * (NH / 8) + q];
    }
  }
  __syncthreads();
  // ---- first CLB layer on the tensor cores: each warp owns 2 tile rows = 2 m16 tiles of pixels; T[32 x NH] = X[32 x 32] Wa^T.
  //      A fragments come straight from global memory (4-byte loads, every sector is used by the k / k+8 halves of a row).
  {
    const int g = lane >> 2, tq = lane & 3;
    float acc[2][NH / 8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NH / 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int oyr = oy0 + 2 * warp + mt;                                  // this m-tile = one row of 16 output pixels
      const int oxa = min(ox0 + g, ow - 1), oxb = min(ox0 + g + 8, ow - 1);    // clamped: out-of-image pixels are never stored
      const long long rowbase = (b * oh + min(oyr, oh - 1)) * static_cast<long long>(ow);
      const uint32_t* xa = reinterpret_cast<const uint32_t*>(x + (rowbase + oxa) * NX);
      const uint32_t* xb = reinterpret_cast<const uint32_t*>(x + (rowbase + oxb) * NX);
#pragma unroll
      for (int ks = 0; ks < NX / 16; ++ks) {
        const uint32_t a0 = __ldg(xa + ks * 8 + tq), a1 = __ldg(xb + ks * 8 + tq);
        const uint32_t a2 = __ldg(xa + ks * 8 + 4 + tq), a3 = __ldg(xb + ks * 8 + 4 + tq);
#pragma unroll
        for (int nt = 0; nt < NH / 8; ++nt) {
          const uint32_t* wrow = s_wa + (nt * 8 + g) * (NX / 2) + ks * 8;     // B[k][n] = Wa[n][k]
          mma_bf16_16816(acc[mt][nt], a0, a1, a2, a3, wrow[tq], wrow[4 + tq]);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NH / 8; ++nt) {
        float* ta = s_t + (warp * 32 + mt * 16 + g) * kTailTStride + nt * 8 + 2 * tq;
        *reinterpret_cast<float2*>(ta) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<float2*>(ta + 8 * kTailTStride) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      }
    __syncwarp();                       // a warp only reads back the 32 pixels it has just written
  }
  const int ox = ox0 + (tid & 15), oy = oy0 + (tid >> 4);
  if (ox >= ow || oy >= oh) return;
  const long long pix = (b * oh + oy) * ow + ox;
  const Bilin s = make_bilin(oy, ox, h, w, oh, ow);
  const int p00 = (s.y0 - ybase) * nc + (s.x0 - xbase), p01 = (s.y0 - ybase) * nc + (s.x1 - xbase);
  const int p10 = (s.y1 - ybase) * nc + (s.x0 - xbase), p11 = (s.y1 - ybase) * nc + (s.x1 - xbase);
  // ---- hidden = gelu(W_a x + up(W_b emb) + b1) -> 4 outputs
  float o4[4] = {0.f, 0.f, 0.f, 0.f};
  const float* trow = s_t + tid * kTailTStride;           // pixel tid = (warp * 32 + lane) of the tile
#pragma unroll 1
  for (int c0 = 0; c0 < NH; c0 += 8) {
    float e00[8], e01[8], e10[8], e11[8];
    unpack8(*reinterpret_cast<const uint4*>(s_e + p00 * EST + 2 * c0), e00);
    unpack8(*reinterpret_cast<const uint4*>(s_e + p01 * EST + 2 * c0), e01);
    unpack8(*reinterpret_cast<const uint4*>(s_e + p10 * EST + 2 * c0), e10);
    unpack8(*reinterpret_cast<const uint4*>(s_e + p11 * EST + 2 * c0), e11);
    const float4 ta = *reinterpret_cast<const float4*>(trow + c0), tb = *reinterpret_cast<const float4*>(trow + c0 + 4);
    const float tv[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
      // same association as the reference interpolation: (1-ly)((1-lx) v00 + lx v01) + ly((1-lx) v10 + lx v11)
      float hv[2];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const float ev = (1.f - s.ly) * ((1.f - s.lx) * e00[u + v] + s.lx * e01[u + v]) + s.ly * ((1.f - s.lx) * e10[u + v] + s.lx * e11[u + v]);
        hv[v] = tv[u + v] + ev + s_b1[c0 + u + v];
      }
      svla_ptx::gelu_erf_pair(hv[0], hv[1]);               // two channels per FFMA2 / FMUL2 (the scalar erf-GELU was 18 of ~32 instructions per channel)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const float4 wv = *reinterpret_cast<const float4*>(s_w2 + 4 * (c0 + u + v));      // warp-wide broadcast
        o4[0] = fmaf(wv.x, hv[v], o4[0]); o4[1] = fmaf(wv.y, hv[v], o4[1]); o4[2] = fmaf(wv.z, hv[v], o4[2]); o4[3] = fmaf(wv.w, hv[v], o4[3]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) o4[q] = softplus_fast(o4[q] + b2[q]);
  const float pp0 = o4[0] + 1e-4f, pp1 = o4[1] + 1e-4f, t0 = o4[2] + 1e-4f, t1 = o4[3] + 1e-4f;
  const float prob = pp0 / (pp0 + pp1);
  const float temp = (max_temp - min_temp) * (t0 / (t0 + t1)) + min_temp;
  const float lp = __logf(fminf(fmaxf(prob, 1e-4f), 1.f)), lq = __logf(fminf(fmaxf(1.f - prob, 1e-4f), 1.f));
  const float inv_temp = 1.f / temp;
  // ---- softmax over the bins of (log C(K-1,k) + k log p + (K-1-k) log(1-p)) / T; depth = sum_k softmax_k * centre_k
  // softmax is shift-invariant: any offset close to the row maximum keeps exp() in range.  The log-binomial logits peak at the
  // mode k* ~ (K-1) p, so the offset is the largest of the three logits around it (the Stirling form of log C can move the
  // arg-max by one bin) instead of a 64-term max pass; a neighbour of the true maximum is at most ~0.7/T above it.
  const int kmode = min(NBINS - 2, max(1, __float2int_rn(static_cast<float>(NBINS - 1) * prob)));
  float mx = -INFINITY;
#pragma unroll
  for (int k = kmode - 1; k <= kmode + 1; ++k)
    mx = fmaxf(mx, (s_lb[k] + static_cast<float>(k) * lp + static_cast<float>(NBINS - 1 - k) * lq) * inv_temp);
  float se = 0.f, sc = 0.f;
  const float it2 = inv_temp * 1.4426950408889634f, dk2 = (lp - lq) * it2, c02 = fmaf(static_cast<float>(NBINS - 1) * lq, it2, -mx * 1.4426950408889634f);
  const float4* q00 = reinterpret_cast<const float4*>(s_bins + p00 * BST), *q01 = reinterpret_cast<const float4*>(s_bins + p01 * BST);
  const float4* q10 = reinterpret_cast<const float4*>(s_bins + p10 * BST), *q11 = reinterpret_cast<const float4*>(s_bins + p11 * BST);
#pragma unroll 4
  for (int k4 = 0; k4 < NBINS / 4; ++k4) {
    const float4 v00 = q00[k4], v01 = q01[k4], v10 = q10[k4], v11 = q11[k4];
    const float cs[4] = {(1.f - s.ly) * ((1.f - s.lx) * v00.x + s.lx * v01.x) + s.ly * ((1.f - s.lx) * v10.x + s.lx * v11.x),
                         (1.f - s.ly) * ((1.f - s.lx) * v00.y + s.lx * v01.y) + s.ly * ((1.f - s.lx) * v10.y + s.lx * v11.y),
                         (1.f - s.ly) * ((1.f - s.lx) * v00.z + s.lx * v01.z) + s.ly * ((1.f - s.lx) * v10.z + s.lx * v11.z),
                         (1.f - s.ly) * ((1.f - s.lx) * v00.w + s.lx * v01.w) + s.ly * ((1.f - s.lx) * v10.w + s.lx * v11.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = 4 * k4 + i;
      // (lb_k + k lp + (K-1-k) lq) / T - mx  =  lb_k / T + (k (lp - lq) / T + ((K-1) lq / T - mx)), in log2 units for ex2
      float ex;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(s_lb[kk], it2, fmaf(static_cast<float>(kk), dk2, c02))));
      se += ex; sc = fmaf(ex, cs[i], sc);
    }
  }
  depth[pix] = sc / se;
}

// ------------------------------------------------------------------------------------------ M5 Ego3D
// model/modeling_spatialvla.py:318-323 (bicubic 384 -> 286 with align_corners, crop [31, 255)), :195-223 (7x7 cell mean,
// K^-1 [u v 1] d), :41-97 (sin/cos encoding).  The 7x7 mean of bicubic samples is SEPARABLE: cell(ci, cj) =
// sum_y sum_x W[ci][y] W[cj][x] depth[y][x] with W[c] = the seven 4-tap cubic stencils of the cell's sample coordinates summed
// into one 13-tap stencil (identical table for rows and columns: square map, square crop).  The first version evaluated all
// 49 x 16 taps per cell straight from global memory (784 scattered loads per cell, issue-bound at 21 % of HBM peak); here a
// CTA owns one PATCH row (2 cell rows x 32 cells): the <= 23 source rows it needs are staged once into shared memory with
// coalesced 128-bit loads (columns 40..347 cover every tap of the crop), one vertical 13-tap pass per (cell row, column) and
// one horizontal 13-tap pass per cell follow, and the 16 encoding rows of the patch row are assembled in shared memory and
// leave as one contiguous 128-bit store stream.
constexpr int kEgoTaps = 13;
constexpr int kEgoX0 = 40, kEgoCols = 308, kEgoRows = 24;
constexpr int kEgoMaxKpad = 256;

__global__ void __launch_bounds__(256)
svla_ego3d_kernel(const float* __restrict__ depth384, const float* __restrict__ intr, int k_stride, float* __restrict__ xyz,
                  __nv_bfloat16* __restrict__ enc, int kpad, int n_freqs) {
  __shared__ float sW[32][kEgoTaps];
  __shared__ int sStart[32];
  __shared__ __align__(16) float sD[kEgoRows][kEgoCols];
  __shared__ float sV[2][kEgoCols];
  __shared__ float sCell[2][32];
  __shared__ __align__(16) __nv_bfloat16 sEnc[16 * kEgoMaxKpad];
  const int pr = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float scale = 383.0f / 285.0f;
  // 13-tap stencils (fp32 coordinates exactly as the 49-sample formulation computes them)
  if (tid < 32) {
    const int first = static_cast<int>(floorf(scale * (tid * 7 + 31))) - 1;
    sStart[tid] = first;
    for (int k = 0; k < kEgoTaps; ++k) sW[tid][k] = 0.f;
    for (int r = 0; r < 7; ++r) {
      const float sc = scale * (tid * 7 + r + 31);
      const int i0 = static_cast<int>(floorf(sc));
      float c[4];
      cubic_coeffs(sc - i0, c);
#pragma unroll
      for (int u = 0; u < 4; ++u) sW[tid][i0 - 1 + u - first] += c[u];
    }
  }
  for (int k = tid; k < 16 * kpad / 2; k += 256) reinterpret_cast<uint32_t*>(sEnc)[k] = 0u;      // K padding columns stay zero
  __syncthreads();
  const int ybase = sStart[2 * pr];
  const int nrows = sStart[2 * pr + 1] + kEgoTaps - ybase;                    // <= 23, last source row <= 344
  const float* dm = depth384 + static_cast<long long>(b) * 384 * 384;
  constexpr int kVecPerRow = kEgoCols / 4;
  for (int k = tid; k < nrows * kVecPerRow; k += 256) {
    const int row = k / kVecPerRow, v = k % kVecPerRow;
    reinterpret_cast<float4*>(&sD[row][0])[v] = __ldg(reinterpret_cast<const float4*>(dm + (ybase + row) * 384 + kEgoX0) + v);
  }
  __syncthreads();
  // vertical pass first: for each of the 2 cell rows and every staged column, the 13-tap column sum.  Consecutive threads take
  // consecutive columns (conflict-free shared-memory reads, the weight is a warp-wide broadcast); 2.2x fewer shared-memory
  // reads than the horizontal-first order, whose 9.4-float lane stride also cost 35 % bank-conflict replays (ncu).
  for (int idx = tid; idx < 2 * kEgoCols; idx += 256) {
    const int half = idx >= kEgoCols ? 1 : 0, x = idx - half * kEgoCols;
    const int ci = 2 * pr + half, r0 = sStart[ci] - ybase;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < kEgoTaps; ++k) v += sW[ci][k] * sD[r0 + k][x];
    sV[half][x] = v;
  }
  __syncthreads();
  if (tid < 64) {
    const int half = tid >> 5, cj = tid & 31;
    const float* src = &sV[half][sStart[cj] - kEgoX0];
    float d = 0.f;
#pragma unroll
    for (int x = 0; x < kEgoTaps; ++x) d += sW[cj][x] * src[x];
    sCell[half][cj] = d / 49.f;
  }
  __syncthreads();
  if (tid < 192) {
    const int half = tid / 96, rem = tid % 96, cj = rem / 3, sub = rem % 3, ci = 2 * pr + half;
    const float d = sCell[half][cj];
    // inverse of the 3x3 intrinsic matrix (adjugate / det) in fp32
    const float* K = intr + static_cast<long long>(b) * k_stride;
    const float a00 = K[0], a01 = K[1], a02 = K[2], a10 = K[3], a11 = K[4], a12 = K[5], a20 = K[6], a21 = K[7], a22 = K[8];
    const float c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const float c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
    const float c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
    const float det = a00 * c00 + a01 * c10 + a02 * c20;
    const float u = cj * 7 + 3.5f, v = ci * 7 + 3.5f;
    float r;
    if (sub == 0) r = (c00 * u + c01 * v + c02) / det;
    else if (sub == 1) r = (c10 * u + c11 * v + c12) / det;
    else r = (c20 * u + c21 * v + c22) / det;
    const float val = r * d;
    const int pl = cj >> 1;                                   // patch within this patch row
    const int m = (half * 2 + (cj & 1)) * 3 + sub;            // (sub_row, sub_col, xyz)
    xyz[(static_cast<long long>(b) * 256 + pr * 16 + pl) * 12 + m] = val;
    const float xn = (val - (sub == 2 ? 2.f : 0.f)) / 2.f;
    __nv_bfloat16* er = sEnc + pl * kpad + m * (2 * n_freqs + 1);
    er[0] = __float2bfloat16(xn);
    float f = 1.f;
    for (int k = 0; k < n_freqs; ++k) {
      float sv, cv;
      sincosf(xn * f, &sv, &cv);               // one shared range reduction (the 16 sinf / cosf calls were half of the issue slots)
      er[1 + k] = __float2bfloat16(sv);
      er[1 + n_freqs + k] = __float2bfloat16(cv);
      f *= 2.f;
    }
  }
  __syncthreads();
  __nv_bfloat16* dst = enc + (static_cast<long long>(b) * 256 + pr * 16) * kpad;      // 16 consecutive patch rows: contiguous
  if ((kpad & 7) == 0 && (reinterpret_cast<uintptr_t>(enc) & 15) == 0) {
    for (int k = tid; k < 16 * kpad / 8; k += 256) reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(sEnc)[k];
  } else {
    for (int k = tid; k < 16 * kpad; k += 256) dst[k] = sEnc[k];
  }
}

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace

// ================================================================================================ C ABI
extern "C" int svla_layernorm(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int cols,
                              void* out_bf16, float* out_f32, int relu, void* stream) {
  SVLA_REQUIRE(x && gamma && beta && (out_bf16 || out_f32), "svla_layernorm: null pointer");
  SVLA_REQUIRE(rows > 0 && cols > 0 && (cols % 4) == 0 && cols <= kRowThreads * kMaxVec * 4, "svla_layernorm: cols=%d unsupported", cols);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out_bf16);
  const unsigned wblocks = static_cast<unsigned>((rows + 3) / 4);
  const int need = (cols / 4 + 31) / 32;      // float4 per lane for the warp-per-row kernels
  if (rows >= 256 && need <= 2) svla_layernorm_warp_kernel<2><<<wblocks, 128, 0, st>>>(x, gamma, beta, eps, rows, cols, ob, out_f32, relu);
  else if (rows >= 256 && need <= 4) svla_layernorm_warp_kernel<4><<<wblocks, 128, 0, st>>>(x, gamma, beta, eps, rows, cols, ob, out_f32, relu);
  else if (rows >= 256 && need <= 9) svla_layernorm_warp_kernel<9><<<wblocks, 128, 0, st>>>(x, gamma, beta, eps, rows, cols, ob, out_f32, relu);
  else if (rows >= 256 && need <= 18) svla_layernorm_warp_kernel<18><<<wblocks, 128, 0, st>>>(x, gamma, beta, eps, rows, cols, ob, out_f32, relu);
  else svla_layernorm_kernel<<<static_cast<unsigned>(rows), kRowThreads, 0, st>>>(x, gamma, beta, eps, cols, ob, out_f32, relu);
  SVLA_LAUNCH_CHECK("svla_layernorm");
  return 0;
}

static int rmsnorm_residual_impl(float* x, const float* branch, const float* w_post, const float* w_pre, float eps, int64_t rows, int cols,
                                 void* out_bf16, void* out_lo_bf16, int n_partials, int64_t partial_stride, void* stream);

extern "C" int svla_rmsnorm_residual(float* x, const float* branch, const float* w_post, const float* w_pre, float eps,
                                     int64_t rows, int cols, void* out_bf16, int n_partials, int64_t partial_stride, void* stream) {
  return rmsnorm_residual_impl(x, branch, w_post, w_pre, eps, rows, cols, out_bf16, nullptr, n_partials, partial_stride, stream);
}

// Decode-chain variant: the normalised activations leave as TWO bf16 planes, hi = bf16(v) and lo = bf16(v - hi) (16 mantissa
// bits together), for the X_HILO mode of svla_gemm_skinny.  Few-row (block-per-row) kernel only.
extern "C" int svla_rmsnorm_residual_hilo(float* x, const float* branch, const float* w_post, const float* w_pre, float eps,
                                          int64_t rows, int cols, void* out_hi_bf16, void* out_lo_bf16, int n_partials,
                                          int64_t partial_stride, void* stream) {
  SVLA_REQUIRE(out_hi_bf16 && out_lo_bf16 && w_pre, "svla_rmsnorm_residual_hilo: needs w_pre and both output planes");
  SVLA_REQUIRE(rows < 2048, "svla_rmsnorm_residual_hilo: decode-sized inputs only (rows=%lld)", (long long)rows);
  return rmsnorm_residual_impl(x, branch, w_post, w_pre, eps, rows, cols, out_hi_bf16, out_lo_bf16, n_partials, partial_stride, stream);
}

static int rmsnorm_residual_impl(float* x, const float* branch, const float* w_post, const float* w_pre, float eps, int64_t rows, int cols,
                                 void* out_bf16, void* out_lo_bf16, int n_partials, int64_t partial_stride, void* stream) {
  SVLA_REQUIRE(x, "svla_rmsnorm_residual: null x");
  SVLA_REQUIRE((branch == nullptr) == (w_post == nullptr), "svla_rmsnorm_residual: branch and w_post go together");
  SVLA_REQUIRE((w_pre == nullptr) == (out_bf16 == nullptr), "svla_rmsnorm_residual: w_pre and out_bf16 go together");
  SVLA_REQUIRE(rows > 0 && cols > 0 && (cols % 4) == 0 && cols <= kRowThreads * kMaxVec * 4, "svla_rmsnorm_residual: cols=%d unsupported", cols);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out_bf16);
  const int np = n_partials < 1 ? 1 : n_partials;
  const unsigned wblocks = static_cast<unsigned>((rows + 3) / 4);
  const int need = (cols / 4 + 31) / 32;
  // many rows (prefill): warp-per-row; few rows (decode, 64 rows): block-per-row keeps all SMs busy
  if (rows >= 2048 && need <= 4) svla_rmsnorm_residual_warp_kernel<4><<<wblocks, 128, 0, st>>>(x, branch, w_post, w_pre, eps, rows, cols, ob, np, partial_stride);
  else if (rows >= 2048 && need <= 9) svla_rmsnorm_residual_warp_kernel<9><<<wblocks, 128, 0, st>>>(x, branch, w_post, w_pre, eps, rows, cols, ob, np, partial_stride);
  else if (rows >= 2048 && need <= 18) svla_rmsnorm_residual_warp_kernel<18><<<wblocks, 128, 0, st>>>(x, branch, w_post, w_pre, eps, rows, cols, ob, np, partial_stride);
  else {
    cudaError_t le = svla_launch_pdl(svla_rmsnorm_residual_kernel, dim3(static_cast<unsigned>(rows)), dim3(kRowThreads), 0, st, x, branch, w_post,
                                     w_pre, eps, cols, ob, np, static_cast<long long>(partial_stride), static_cast<__nv_bfloat16*>(out_lo_bf16));
    SVLA_REQUIRE(le == cudaSuccess, "svla_rmsnorm_residual: launch failed: %s", cudaGetErrorString(le));
  }
  SVLA_LAUNCH_CHECK("svla_rmsnorm_residual");
  return 0;
}

extern "C" int svla_rope_kv(const void* qkv, void* q_out, void* kcache, void* vcache, int batch, int s, int hq, int hkv, int d,
                            int smax, int pos0, float theta, const float* qkv_f32, int n_partials, int64_t partial_stride,
                            const int32_t* row_pads, void* stream) {
  SVLA_REQUIRE((qkv || qkv_f32) && q_out && kcache && vcache, "svla_rope_kv: null pointer");
  SVLA_REQUIRE(batch > 0 && s > 0 && (d % 2) == 0 && pos0 >= 0 && pos0 + s <= smax, "svla_rope_kv: bad geometry (pos0=%d s=%d smax=%d)", pos0, s, smax);
  SVLA_REQUIRE(d / 2 <= 512, "svla_rope_kv: head dim too large");
  if (qkv_f32 == nullptr && (d % 16) == 0 && d <= 256 && 128 % (d / 16) == 0 && batch * s > 1024) {
    svla_rope_kv_vec_kernel<<<static_cast<unsigned>(batch) * s, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(q_out), static_cast<__nv_bfloat16*>(kcache),
        static_cast<__nv_bfloat16*>(vcache), s, hq, hkv, d, smax, pos0, theta, row_pads);
    SVLA_LAUNCH_CHECK("svla_rope_kv_vec");
    return 0;
  }
  const int rope_threads = (d / 2) * ((batch * s <= 1024) ? max(1, 512 / (d / 2)) : 1);     // decode: spread heads over lanes
  svla_rope_kv_kernel<<<static_cast<unsigned>(batch) * s, rope_threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(q_out), static_cast<__nv_bfloat16*>(kcache),
      static_cast<__nv_bfloat16*>(vcache), s, hq, hkv, d, smax, pos0, theta, qkv_f32, n_partials < 1 ? 1 : n_partials, partial_stride,
      row_pads);
  SVLA_LAUNCH_CHECK("svla_rope_kv");
  return 0;
}

extern "C" int svla_embed_tokens(const int64_t* ids, const void* embed, const void* spatial_embed, const float* image_feats,
                                 float* x, int batch, int s, int hdim, int64_t vocab, int64_t image_token, int64_t act_lo,
                                 int64_t n_act, int n_img, float normalizer, int* status_flag, void* stream) {
  SVLA_REQUIRE(ids && embed && x, "svla_embed_tokens: null pointer");
  SVLA_REQUIRE(batch > 0 && s > 0 && hdim > 0, "svla_embed_tokens: empty problem");
  cudaError_t le = svla_launch_pdl(svla_embed_kernel, dim3(static_cast<unsigned>(batch) * s), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                   reinterpret_cast<const long long*>(ids), static_cast<const __nv_bfloat16*>(embed),
                                   static_cast<const __nv_bfloat16*>(spatial_embed), image_feats, x, s, hdim, static_cast<long long>(vocab),
                                   static_cast<long long>(image_token), static_cast<long long>(act_lo), static_cast<long long>(n_act), n_img,
                                   normalizer, status_flag);
  SVLA_REQUIRE(le == cudaSuccess, "svla_embed_tokens: launch failed: %s", cudaGetErrorString(le));
  SVLA_LAUNCH_CHECK("svla_embed_tokens");
  return 0;
}

// ZoeDepth metric-head router decision ON THE DEVICE (HF zoedepth/modeling_zoedepth.py:1059-1067: argmax of the batch-summed
// domain logits, which HF reads back with `.item()`): every CTA recomputes the (tiny) vote and copies its slice of the selected
// head's packed weight arena into the ACTIVE arena the following kernels read, so one CUDA graph covers the whole step.
__global__ void __launch_bounds__(256)
svla_zoe_select_head_kernel(const float* __restrict__ dlog, int batch, int n_heads, int forced, const uint4* __restrict__ const* __restrict__ srcs,
                            uint4* __restrict__ dst, long long n16, int* __restrict__ head_out) {
  __shared__ int head_sh;
  if (threadIdx.x == 0) {
    int best = 0;
    if (forced >= 0) best = forced;
    else {
      float bv = -INFINITY;
      for (int h = 0; h < n_heads; ++h) {
        float v = 0.f;
        for (int b = 0; b < batch; ++b) v += dlog[b * n_heads + h];      // torch.sum over dim 0, then first maximum wins
        if (v > bv) { bv = v; best = h; }
      }
    }
    head_sh = best;
    if (blockIdx.x == 0 && head_out) *head_out = best;
  }
  __syncthreads();
  const uint4* __restrict__ src = srcs[head_sh];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

extern "C" int svla_zoe_select_head(const float* domain_logits, int batch, int n_heads, int forced_head, const void* const* head_arenas_dev,
                                    void* active_arena, int64_t bytes, int* head_out, void* stream) {
  SVLA_REQUIRE(domain_logits && head_arenas_dev && active_arena && batch > 0 && n_heads > 0 && n_heads <= 8, "svla_zoe_select_head: bad arguments");
  SVLA_REQUIRE(bytes > 0 && (bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(active_arena) & 15) == 0, "svla_zoe_select_head: arena must be 16-byte granular");
  SVLA_REQUIRE(forced_head < n_heads, "svla_zoe_select_head: forced head %d out of range", forced_head);
  const long long n16 = bytes / 16;
  long long blocks = (n16 + 255) / 256;
  if (blocks > 2LL * svla_num_sms()) blocks = 2LL * svla_num_sms();
  svla_zoe_select_head_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      domain_logits, batch, n_heads, forced_head, reinterpret_cast<const uint4* const*>(head_arenas_dev), static_cast<uint4*>(active_arena), n16,
      head_out);
  SVLA_LAUNCH_CHECK("svla_zoe_select_head");
  return 0;
}

extern "C" int svla_argmax_rows(const float* logits, int64_t rows, int64_t cols, int64_t ld, int64_t id_offset, int64_t* out_ids,
                                int64_t out_stride, void* stream) {
  SVLA_REQUIRE(logits && out_ids && rows > 0 && cols > 0, "svla_argmax_rows: bad arguments");
  cudaError_t le = svla_launch_pdl(svla_argmax_kernel, dim3(static_cast<unsigned>(rows)), dim3(256), 0, static_cast<cudaStream_t>(stream), logits,
                                   static_cast<long long>(cols), static_cast<long long>(ld), static_cast<long long>(id_offset),
                                   reinterpret_cast<long long*>(out_ids), static_cast<long long>(out_stride));
  SVLA_REQUIRE(le == cudaSuccess, "svla_argmax_rows: launch failed: %s", cudaGetErrorString(le));
  SVLA_LAUNCH_CHECK("svla_argmax_rows");
  return 0;
}

extern "C" int svla_cross_entropy_rows(const float* logits, int64_t rows, int64_t cols, int64_t ld, const int64_t* labels,
                                       int64_t ignore_index, float* row_loss, int64_t* row_argmax, int64_t row_offset, float* summary,
                                       void* stream) {
  SVLA_REQUIRE(logits && labels && row_loss && row_argmax, "svla_cross_entropy_rows: null operand");
  SVLA_REQUIRE(rows > 0 && cols > 0 && ld >= cols && row_offset >= 0, "svla_cross_entropy_rows: bad shape (rows %lld cols %lld ld %lld)",
               static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld));
  SVLA_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 3) == 0, "svla_cross_entropy_rows: logits must be 4-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  svla_cross_entropy_kernel<<<static_cast<unsigned>(rows), kCeThreads, 0, st>>>(
      logits, static_cast<long long>(cols), static_cast<long long>(ld), reinterpret_cast<const long long*>(labels) + row_offset,
      static_cast<long long>(ignore_index), row_loss + row_offset, reinterpret_cast<long long*>(row_argmax) + row_offset);
  SVLA_LAUNCH_CHECK("svla_cross_entropy_rows");
  if (summary) {
    svla_cross_entropy_summary_kernel<<<1, 256, 0, st>>>(row_loss, reinterpret_cast<const long long*>(row_argmax),
                                                         reinterpret_cast<const long long*>(labels),
                                                         static_cast<long long>(row_offset + rows), static_cast<long long>(ignore_index), summary);
    SVLA_LAUNCH_CHECK("svla_cross_entropy_summary");
  }
  return 0;
}

extern "C" int svla_cross_entropy_bwd(const float* logits, int64_t rows, int64_t cols, int64_t ld, const int64_t* labels,
                                      int64_t ignore_index, const float* row_loss, int64_t row_offset, const float* summary,
                                      float softcap, void* dz_bf16, int64_t ldo, void* stream) {
  SVLA_REQUIRE(logits && labels && row_loss && summary && dz_bf16, "svla_cross_entropy_bwd: null operand");
  SVLA_REQUIRE(rows > 0 && rows <= 65535 && cols > 0 && ld >= cols && ldo >= cols && (ldo % 2) == 0 && row_offset >= 0,
               "svla_cross_entropy_bwd: bad shape (rows %lld cols %lld ld %lld ldo %lld)", static_cast<long long>(rows),
               static_cast<long long>(cols), static_cast<long long>(ld), static_cast<long long>(ldo));
  SVLA_REQUIRE((reinterpret_cast<uintptr_t>(dz_bf16) & 3) == 0, "svla_cross_entropy_bwd: dz must be 4-byte aligned");
  const long long per_cta = kCeBwdThreads * kCeBwdPairs * 2;
  dim3 grid(static_cast<unsigned>((ldo + per_cta - 1) / per_cta), static_cast<unsigned>(rows));
  svla_cross_entropy_bwd_kernel<<<grid, kCeBwdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, static_cast<long long>(cols), static_cast<long long>(ld), reinterpret_cast<const long long*>(labels) + row_offset,
      static_cast<long long>(ignore_index), row_loss + row_offset, summary, softcap, static_cast<__nv_bfloat16*>(dz_bf16),
      static_cast<long long>(ldo));
  SVLA_LAUNCH_CHECK("svla_cross_entropy_bwd");
  return 0;
}

extern "C" int svla_sumsq(const float* x, int64_t n, float* out, void* stream) {
  SVLA_REQUIRE(x && out && n > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "svla_sumsq: bad arguments");
  const long long blocks = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(svla_num_sms()) * 4;
  svla_sumsq_kernel<<<static_cast<unsigned>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<long long>(n), out);
  SVLA_LAUNCH_CHECK("svla_sumsq");
  return 0;
}

extern "C" int svla_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                               double beta2, double eps, double weight_decay, int64_t step, double grad_scale,
                               const float* grad_sumsq_dev, double max_grad_norm, void* stream) {
  SVLA_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "svla_adamw_step: bad arguments");
  SVLA_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "svla_adamw_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  const long long blocks = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(svla_num_sms()) * 8;
  svla_adamw_kernel<<<static_cast<unsigned>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, static_cast<long long>(n), static_cast<float>(1.0 - lr * weight_decay),
      static_cast<float>(1.0 - beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta2), static_cast<float>(eps),
      static_cast<float>(lr / bc1), static_cast<float>(sqrt(bc2)), static_cast<float>(grad_scale), grad_sumsq_dev,
      static_cast<float>(max_grad_norm));
  SVLA_LAUNCH_CHECK("svla_adamw_step");
  return 0;
}

extern "C" int svla_siglip_patchify(const float* px, void* a, int batch, int kpad, void* stream) {
  SVLA_REQUIRE(px && a && batch > 0 && kpad >= 588, "svla_siglip_patchify: bad arguments");
  svla_siglip_patchify_kernel<<<static_cast<unsigned>(batch) * 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      px, static_cast<__nv_bfloat16*>(a), kpad);
  SVLA_LAUNCH_CHECK("svla_siglip_patchify");
  return 0;
}

extern "C" int svla_zoe_patchify(const float* px, void* a, int batch, void* stream) {
  SVLA_REQUIRE(px && a && batch > 0, "svla_zoe_patchify: bad arguments");
  svla_zoe_patchify_kernel<<<static_cast<unsigned>(batch) * 576, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      px, static_cast<__nv_bfloat16*>(a));
  SVLA_LAUNCH_CHECK("svla_zoe_patchify");
  return 0;
}

extern "C" int svla_beit_assemble(const float* patches, const float* cls, float* x, int batch, int n, int c, void* stream) {
  SVLA_REQUIRE(patches && cls && x && batch > 0 && n > 0 && c > 0, "svla_beit_assemble: bad arguments");
  svla_beit_assemble_kernel<<<static_cast<unsigned>(batch) * (n + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(patches, cls, x, n, c);
  SVLA_LAUNCH_CHECK("svla_beit_assemble");
  return 0;
}

extern "C" int svla_readout_concat(const float* hs, void* a, int batch, int n, int c, void* stream) {
  SVLA_REQUIRE(hs && a && batch > 0 && n > 0 && c > 0, "svla_readout_concat: bad arguments");
  svla_readout_concat_kernel<<<static_cast<unsigned>(batch) * n, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      hs, static_cast<__nv_bfloat16*>(a), n, c);
  SVLA_LAUNCH_CHECK("svla_readout_concat");
  return 0;
}

extern "C" int svla_pixel_shuffle(const void* g, void* out, int batch, int h, int w, int c, int f, void* stream) {
  SVLA_REQUIRE(g && out && batch > 0 && (c % 8) == 0 && f > 0, "svla_pixel_shuffle: bad arguments");
  const long long total = static_cast<long long>(batch) * h * f * w * f * (c / 8);
  svla_pixel_shuffle_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(g), static_cast<uint4*>(out), h, w, c / 8, f, total);
  SVLA_LAUNCH_CHECK("svla_pixel_shuffle");
  return 0;
}

extern "C" int svla_im2col3x3_s2(const void* x, void* a, int batch, int h, int w, int c, void* stream) {
  SVLA_REQUIRE(x && a && batch > 0 && (c % 8) == 0 && (h % 2) == 0 && (w % 2) == 0, "svla_im2col3x3_s2: bad arguments");
  const long long total = static_cast<long long>(batch) * (h / 2) * (w / 2) * 9 * (c / 8);
  svla_im2col3x3_s2_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(a), h, w, c / 8, total);
  SVLA_LAUNCH_CHECK("svla_im2col3x3_s2");
  return 0;
}

extern "C" int svla_bilinear_nhwc(const void* x, const void* add, void* out, void* out_relu, int batch, int h, int w, int c,
                                  int oh, int ow, void* stream) {
  SVLA_REQUIRE(x && (out || out_relu) && batch > 0 && (c % 8) == 0, "svla_bilinear_nhwc: bad arguments");
  SVLA_REQUIRE(oh <= 65535 && batch <= 65535 && static_cast<long long>(ow) * (c / 8) < (1LL << 31), "svla_bilinear_nhwc: grid too large");
  // tiled two-pass kernel for up-sampling ratios <= ~1.2 (every use on this path is x2); the per-pixel kernel covers the rest
  const float ry = (oh > 1) ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.f;
  const int nrows = static_cast<int>(floorf(7.f * ry)) + 3;
  const int c8 = c / 8, cgroups = (c8 + kBlGroups - 1) / kBlGroups;
  static const bool legacy_bilinear = getenv("SVLA_BILINEAR_LEGACY") != nullptr;
  if (!legacy_bilinear && nrows <= 12 && static_cast<long long>(batch) * cgroups <= 65535) {
    const size_t smem = static_cast<size_t>(nrows) * kBlCols * kBlGroups * 2 * sizeof(float4);
    static size_t configured = 0;
    if (smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(svla_bilinear_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      SVLA_REQUIRE(e == cudaSuccess, "svla_bilinear_nhwc: smem opt-in %zu failed: %s", smem, cudaGetErrorString(e));
      configured = smem;
    }
    dim3 tgrid((ow + kBlCols - 1) / kBlCols, (oh + kBlRows - 1) / kBlRows, batch * cgroups);
    svla_bilinear_tile_kernel<<<tgrid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), static_cast<const uint4*>(add), static_cast<uint4*>(out), static_cast<uint4*>(out_relu), h, w,
        c8, oh, ow, nrows, cgroups);
    SVLA_LAUNCH_CHECK("svla_bilinear_tile");
    return 0;
  }
  dim3 grid(blocks_for(static_cast<long long>(ow) * (c / 8), 256), oh, batch);
  svla_bilinear_nhwc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<const uint4*>(add), static_cast<uint4*>(out), static_cast<uint4*>(out_relu), h, w,
      c / 8, oh, ow);
  SVLA_LAUNCH_CHECK("svla_bilinear_nhwc");
  return 0;
}

extern "C" int svla_relu_bf16(const void* x, void* out, int64_t n, void* stream) {
  SVLA_REQUIRE(x && out && n > 0 && (n % 2) == 0, "svla_relu_bf16: bad arguments");
  svla_relu_bf16_kernel<<<blocks_for(n / 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat162*>(x), static_cast<__nv_bfloat162*>(out), n / 2);
  SVLA_LAUNCH_CHECK("svla_relu_bf16");
  return 0;
}

extern "C" int svla_softplus_f32(const void* x, float* out, int64_t n, void* stream) {
  SVLA_REQUIRE(x && out && n > 0, "svla_softplus_f32: bad arguments");
  svla_softplus_f32_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), out, n);
  SVLA_LAUNCH_CHECK("svla_softplus_f32");
  return 0;
}

extern "C" int svla_zoe_router_embed(const float* conv, float* e, void* e_bf16, int batch, int n, int c, void* stream) {
  SVLA_REQUIRE(conv && e && batch > 0 && n > 0 && (c % 2) == 0, "svla_zoe_router_embed: bad arguments");
  svla_zoe_router_embed_kernel<<<static_cast<unsigned>(batch) * (n + 1), 128, 0, static_cast<cudaStream_t>(stream)>>>(conv, e, static_cast<__nv_bfloat16*>(e_bf16), n, c);
  SVLA_LAUNCH_CHECK("svla_zoe_router_embed");
  return 0;
}

extern "C" int svla_zoe_attractor(const void* attr, const float* prev, float* out, int batch, int h, int w, int oh, int ow, int na,
                                  int nbins, void* stream) {
  SVLA_REQUIRE(attr && prev && out && batch > 0 && na > 0 && na <= 16 && nbins > 0 && (nbins % 4) == 0,
               "svla_zoe_attractor: bad arguments (need na <= 16, nbins %% 4 == 0)");
  SVLA_REQUIRE(oh <= 65535 && batch <= 65535, "svla_zoe_attractor: grid too large");
  dim3 grid((ow + 15) / 16, oh, batch);
  svla_zoe_attractor_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(attr), prev, out, h, w, oh, ow, na, nbins);
  SVLA_LAUNCH_CHECK("svla_zoe_attractor");
  return 0;
}

extern "C" int svla_zoe_depth_tail(const void* t, const void* e, const float* b1, const float* w2, const float* b2,
                                   const float* bins, float* depth, int batch, int h, int w, int oh, int ow, int nh, int nbins,
                                   float min_temp, float max_temp, void* stream) {
  SVLA_REQUIRE(t && e && b1 && w2 && b2 && bins && depth, "svla_zoe_depth_tail: null pointer");
  SVLA_REQUIRE(batch > 0 && nbins > 0 && nbins <= 64 && nh > 0 && nh <= 64, "svla_zoe_depth_tail: bad geometry");
  SVLA_REQUIRE(oh <= 65535 && batch <= 65535, "svla_zoe_depth_tail: grid too large");
  dim3 grid((ow + 63) / 64, oh, batch);
  svla_zoe_depth_tail_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(t), static_cast<const __nv_bfloat16*>(e), b1, w2, b2, bins, depth, h, w, oh, ow, nh, nbins,
      min_temp, max_temp);
  SVLA_LAUNCH_CHECK("svla_zoe_depth_tail");
  return 0;
}

extern "C" int svla_zoe_depth_tail_fused(const void* x, const void* wa, const void* e, const float* b1, const float* w2,
                                         const float* b2, const float* bins, float* depth, int batch, int h, int w, int oh, int ow,
                                         int nx, int nh, int nbins, float min_temp, float max_temp, void* stream) {
  SVLA_REQUIRE(x && wa && e && b1 && w2 && b2 && bins && depth, "svla_zoe_depth_tail_fused: null pointer");
  SVLA_REQUIRE(nx == 32 && nh == 40 && nbins == 64, "svla_zoe_depth_tail_fused: specialised for 32 features -> 40 hidden, 64 bins (got %d, %d, %d)", nx, nh, nbins);
  SVLA_REQUIRE(batch > 0 && batch <= 65535 && h > 0 && w > 0 && oh > 0 && ow > 0, "svla_zoe_depth_tail_fused: bad geometry");
  const float ry = (oh > 1) ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.f;
  const float rx = (ow > 1) ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) : 0.f;
  const int nr = static_cast<int>(floorf((kTailTile - 1) * ry)) + 3, nc = static_cast<int>(floorf((kTailTile - 1) * rx)) + 3;
  SVLA_REQUIRE(nr <= kTailMaxSrc && nc <= kTailMaxSrc, "svla_zoe_depth_tail_fused: resampling ratio %.3f too large (up-sampling only)", ry);
  const size_t smem = static_cast<size_t>(kTailMaxSrc) * kTailMaxSrc * ((64 + 4) * 4 + 40 * 2) + 256 * kTailTStride * sizeof(float) +
                      (40 * 32 / 2 + 4 * 40 + 40 + 64) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(svla_zoe_depth_tail_fused_kernel<32, 40, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    SVLA_REQUIRE(ce == cudaSuccess, "svla_zoe_depth_tail_fused: smem opt-in failed: %s", cudaGetErrorString(ce));
    configured = true;
  }
  dim3 grid((ow + kTailTile - 1) / kTailTile, (oh + kTailTile - 1) / kTailTile, batch);
  svla_zoe_depth_tail_fused_kernel<32, 40, 64><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(wa), static_cast<const __nv_bfloat16*>(e), b1, w2, b2, bins,
      depth, h, w, oh, ow, nr, nc, min_temp, max_temp);
  SVLA_LAUNCH_CHECK("svla_zoe_depth_tail_fused");
  return 0;
}

extern "C" int svla_ego3d_encode(const float* depth384, const float* intrinsic, int k_stride, float* xyz, void* enc, int batch,
                                 int kpad, int n_freqs, void* stream) {
  SVLA_REQUIRE(depth384 && intrinsic && xyz && enc && batch > 0, "svla_ego3d_encode: bad arguments");
  SVLA_REQUIRE(kpad >= 12 * (2 * n_freqs + 1) && (k_stride == 0 || k_stride == 9), "svla_ego3d_encode: bad kpad / k_stride");
  SVLA_REQUIRE(kpad <= kEgoMaxKpad && (kpad % 2) == 0 && n_freqs >= 0, "svla_ego3d_encode: kpad must be even and <= %d", kEgoMaxKpad);
  SVLA_REQUIRE((reinterpret_cast<uintptr_t>(depth384) & 15) == 0, "svla_ego3d_encode: depth maps must be 16-byte aligned");
  dim3 grid(16, batch);
  svla_ego3d_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(depth384, intrinsic, k_stride, xyz,
                                                                        static_cast<__nv_bfloat16*>(enc), kpad, n_freqs);
  SVLA_LAUNCH_CHECK("svla_ego3d_encode");
  return 0;
}
