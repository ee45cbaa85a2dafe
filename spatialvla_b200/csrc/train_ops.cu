// Memory-bound kernels of the LoRA fine-tune step (BASELINE.json config #5): the training-mode forward
// pieces that must keep what the backward needs, and the backward of every non-GEMM op of Gemma2 / SigLIP / the Ego3D head.
//   reference forward:  model/modeling_gemma2.py:60-77 (RMSNorm (1 + w)), :80-92 (GeGLU), :95-154 (RoPE), :451-506 (sandwich layer);
//                       HF siglip/modeling_siglip.py:330-362 (pre-LN block, gelu_pytorch_tanh MLP); model/modeling_spatialvla.py:59-64
//                       (Ego3D head: Linear, LayerNorm, ReLU, Linear).
//   backward:           torch autograd in the reference (train/spatialvla_finetune.py + HF Trainer); the closed forms are written
//                       out and checked against autograd in oracle/backward_ref.py.
// Layout rules of the step: residual streams and their gradients are fp32 [tokens, hidden]; every GEMM operand (activations and
// activation gradients) is bf16; norm weights are frozen (LoRA adapts Linear layers only), so no weight gradients are formed here.
// All kernels: one warp per row, 128-bit accesses, fp32 statistics, warp-shuffle reductions; rows are re-read from L1/L2 for the
// second pass instead of being held in registers, so any hidden size that is a multiple of 4 works.
#include "svla_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ float4 ld4(const float* p, int i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ float4 ld4_bf16(const __nv_bfloat16* p, int i) {
  const uint2 t = reinterpret_cast<const uint2*>(p)[i];
  return make_float4(bf16_bits_to_float(t.x & 0xFFFFu), bf16_bits_to_float(t.x >> 16), bf16_bits_to_float(t.y & 0xFFFFu),
                     bf16_bits_to_float(t.y >> 16));
}
__device__ __forceinline__ void st4_bf16(__nv_bfloat16* p, int i, float4 v) {
  reinterpret_cast<uint2*>(p)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__device__ __forceinline__ float gelu_tanh_grad_f(float z) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (z + k1 * z * z * z);
  const float t = tanhf(u);
  return 0.5f * (1.f + t) + 0.5f * z * (1.f - t * t) * k0 * (1.f + 3.f * k1 * z * z);
}

// ------------------------------------------------------------------------------------------ Gemma2 sandwich norm, training forward
// x_out = x_in + rms(branch) * (1 + w_post)   (branch == NULL: x_out is not written, x = x_in)
// h     = rms(x) * (1 + w_pre)  in bf16       (w_pre == NULL: skipped)
// Out of place: every layer keeps its own residual tensor for the backward pass (180 GB of HBM: nothing is recomputed).
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
svla_rmsnorm_train_fwd_kernel(const float* __restrict__ x_in, const float* __restrict__ branch, const float* __restrict__ w_post,
                              const float* __restrict__ w_pre, float eps, long long rows, int cols, float* __restrict__ x_out,
                              __nv_bfloat16* __restrict__ h) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = cols >> 2;
  const float* xi = x_in + row * cols;
  const float* xs = xi;                 // the row the pre-norm reads
  if (branch) {
    const float* br = branch + row * cols;
    float ss = 0.f;
    for (int i = lane; i < nv; i += 32) { const float4 b = ld4(br, i); ss += dot4(b, b); }
    const float r = rsqrtf(warp_sum(ss) / cols + eps);
    float* xo = x_out + row * cols;
    for (int i = lane; i < nv; i += 32) {
      const float4 b = ld4(br, i), w = __ldg(reinterpret_cast<const float4*>(w_post) + i);
      float4 x = ld4(xi, i);
      x.x += b.x * r * (1.f + w.x); x.y += b.y * r * (1.f + w.y); x.z += b.z * r * (1.f + w.z); x.w += b.w * r * (1.f + w.w);
      reinterpret_cast<float4*>(xo)[i] = x;
    }
    __syncwarp();
    xs = xo;
  }
  if (w_pre) {
    float ss = 0.f;
    for (int i = lane; i < nv; i += 32) { const float4 x = ld4(xs, i); ss += dot4(x, x); }
    const float r = rsqrtf(warp_sum(ss) / cols + eps);
    __nv_bfloat16* ho = h + row * cols;
    for (int i = lane; i < nv; i += 32) {
      const float4 x = ld4(xs, i), w = __ldg(reinterpret_cast<const float4*>(w_pre) + i);
      st4_bf16(ho, i, make_float4(x.x * r * (1.f + w.x), x.y * r * (1.f + w.y), x.z * r * (1.f + w.z), x.w * r * (1.f + w.w)));
    }
  }
}

// ------------------------------------------------------------------------------------------ RMSNorm backward
// y = x * rsqrt(mean(x^2) + eps) * (1 + w):   dx = r * (g - x * r^2 * mean(g * x)),  g = dy * (1 + w),  r = rsqrt(mean(x^2) + eps)
// dy: bf16 (a GEMM output) or fp32 (the residual-stream gradient).  row_idx: dy row i belongs to x / dx row row_idx[i] (the labelled
// rows of the final norm).  Output: dx_accum[row] += dx (fp32 residual-stream gradient) and / or dx_bf16[row] = dx (GEMM operand).
template <bool DY_F32>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
svla_rmsnorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const void* __restrict__ dy_, const long long* __restrict__ row_idx,
                        float eps, long long rows, int cols, float* __restrict__ dx_accum, __nv_bfloat16* __restrict__ dx_bf16) {
  const int lane = threadIdx.x & 31;
  const long long i_row = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (i_row >= rows) return;
  const long long row = row_idx ? row_idx[i_row] : i_row;
  const int nv = cols >> 2;
  const float* xr = x + row * cols;
  auto load_dy = [&](int i) -> float4 {
    if constexpr (DY_F32) return ld4(static_cast<const float*>(dy_) + i_row * cols, i);
    else return ld4_bf16(static_cast<const __nv_bfloat16*>(dy_) + i_row * cols, i);
  };
  float ss = 0.f, gx = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 xv = ld4(xr, i), d = load_dy(i), wv = __ldg(reinterpret_cast<const float4*>(w) + i);
    ss += dot4(xv, xv);
    gx += d.x * (1.f + wv.x) * xv.x + d.y * (1.f + wv.y) * xv.y + d.z * (1.f + wv.z) * xv.z + d.w * (1.f + wv.w) * xv.w;
  }
  const float r = rsqrtf(warp_sum(ss) / cols + eps);
  const float c = warp_sum(gx) / cols * r * r;
  for (int i = lane; i < nv; i += 32) {
    const float4 xv = ld4(xr, i), d = load_dy(i), wv = __ldg(reinterpret_cast<const float4*>(w) + i);
    float4 o;
    o.x = r * (d.x * (1.f + wv.x) - xv.x * c);
    o.y = r * (d.y * (1.f + wv.y) - xv.y * c);
    o.z = r * (d.z * (1.f + wv.z) - xv.z * c);
    o.w = r * (d.w * (1.f + wv.w) - xv.w * c);
    if (dx_accum) {
      float4* p = reinterpret_cast<float4*>(dx_accum + row * cols) + i;
      float4 a = *p;
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
      *p = a;
    }
    if (dx_bf16) st4_bf16(dx_bf16 + row * cols, i, o);
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm backward
// y = (x - mu) * rstd * gamma + beta [, relu]:  g = dy * gamma [* (y > 0)],  xh = (x - mu) * rstd,
//                                               dx = rstd * (g - mean(g) - xh * mean(g * xh))
// dx_accum += dx (fp32 residual gradient; NULL = skip), copy_bf16 = bf16 of the UPDATED dx_accum row (the next GEMM's operand),
// dx_bf16 = dx itself (no accumulation; the Ego3D head, whose input gradient only feeds a GEMM).
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
svla_layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const __nv_bfloat16* __restrict__ dy, float eps, long long rows, int cols, int relu,
                          float* __restrict__ dx_accum, __nv_bfloat16* __restrict__ copy_bf16, __nv_bfloat16* __restrict__ dx_bf16) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = cols >> 2;
  const float* xr = x + row * cols;
  const __nv_bfloat16* dr = dy + row * cols;
  float s = 0.f;
  for (int i = lane; i < nv; i += 32) { const float4 v = ld4(xr, i); s += v.x + v.y + v.z + v.w; }
  const float mu = warp_sum(s) / cols;
  float ss = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = ld4(xr, i);
    const float a = v.x - mu, b = v.y - mu, c = v.z - mu, d = v.w - mu;
    ss += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / cols + eps);
  auto grad_in = [&](int i, float4& xh) -> float4 {
    const float4 v = ld4(xr, i), d = ld4_bf16(dr, i), gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    xh = make_float4((v.x - mu) * rstd, (v.y - mu) * rstd, (v.z - mu) * rstd, (v.w - mu) * rstd);
    float4 g = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
    if (relu) {
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
      if (xh.x * gm.x + bt.x <= 0.f) g.x = 0.f;
      if (xh.y * gm.y + bt.y <= 0.f) g.y = 0.f;
      if (xh.z * gm.z + bt.z <= 0.f) g.z = 0.f;
      if (xh.w * gm.w + bt.w <= 0.f) g.w = 0.f;
    }
    return g;
  };
  float sg = 0.f, sgx = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float4 xh;
    const float4 g = grad_in(i, xh);
    sg += g.x + g.y + g.z + g.w;
    sgx += dot4(g, xh);
  }
  const float mg = warp_sum(sg) / cols, mgx = warp_sum(sgx) / cols;
  for (int i = lane; i < nv; i += 32) {
    float4 xh;
    const float4 g = grad_in(i, xh);
    float4 o = make_float4(rstd * (g.x - mg - xh.x * mgx), rstd * (g.y - mg - xh.y * mgx), rstd * (g.z - mg - xh.z * mgx),
                           rstd * (g.w - mg - xh.w * mgx));
    if (dx_bf16) st4_bf16(dx_bf16 + row * cols, i, o);
    if (dx_accum) {
      float4* p = reinterpret_cast<float4*>(dx_accum + row * cols) + i;
      float4 a = *p;
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
      *p = a;
      if (copy_bf16) st4_bf16(copy_bf16 + row * cols, i, a);
    }
  }
}

// ------------------------------------------------------------------------------------------ GeGLU / GELU element-wise
// gu bf16 [rows, 2 * inter] with columns 2j = gate_j, 2j + 1 = up_j (the interleaved rows of the fused gate/up weight)
__global__ void svla_geglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ act, long long n4) {
  // one thread: 8 gu values -> 4 act values
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 t = reinterpret_cast<const uint4*>(gu)[i];
    const float g0 = bf16_bits_to_float(t.x & 0xFFFFu), u0 = bf16_bits_to_float(t.x >> 16);
    const float g1 = bf16_bits_to_float(t.y & 0xFFFFu), u1 = bf16_bits_to_float(t.y >> 16);
    const float g2 = bf16_bits_to_float(t.z & 0xFFFFu), u2 = bf16_bits_to_float(t.z >> 16);
    const float g3 = bf16_bits_to_float(t.w & 0xFFFFu), u3 = bf16_bits_to_float(t.w >> 16);
    reinterpret_cast<uint2*>(act)[i] = make_uint2(pack_bf16x2(gelu_tanh_f(g0) * u0, gelu_tanh_f(g1) * u1),
                                                  pack_bf16x2(gelu_tanh_f(g2) * u2, gelu_tanh_f(g3) * u3));
  }
}
// d gate = d act * up * gelu'(gate),  d up = d act * gelu(gate)
__global__ void svla_geglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, const __nv_bfloat16* __restrict__ dact,
                                      __nv_bfloat16* __restrict__ dgu, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 t = reinterpret_cast<const uint4*>(gu)[i];
    const uint2 da = reinterpret_cast<const uint2*>(dact)[i];
    const float g[4] = {bf16_bits_to_float(t.x & 0xFFFFu), bf16_bits_to_float(t.y & 0xFFFFu), bf16_bits_to_float(t.z & 0xFFFFu),
                        bf16_bits_to_float(t.w & 0xFFFFu)};
    const float u[4] = {bf16_bits_to_float(t.x >> 16), bf16_bits_to_float(t.y >> 16), bf16_bits_to_float(t.z >> 16), bf16_bits_to_float(t.w >> 16)};
    const float d[4] = {bf16_bits_to_float(da.x & 0xFFFFu), bf16_bits_to_float(da.x >> 16), bf16_bits_to_float(da.y & 0xFFFFu),
                        bf16_bits_to_float(da.y >> 16)};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2(d[k] * u[k] * gelu_tanh_grad_f(g[k]), d[k] * gelu_tanh_f(g[k]));
    reinterpret_cast<uint4*>(dgu)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void svla_gelu_tanh_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ f, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 t = reinterpret_cast<const uint4*>(z)[i];
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2(gelu_tanh_f(bf16_bits_to_float(w[k] & 0xFFFFu)), gelu_tanh_f(bf16_bits_to_float(w[k] >> 16)));
    reinterpret_cast<uint4*>(f)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void svla_gelu_tanh_bwd_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ df,
                                          __nv_bfloat16* __restrict__ dz, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 t = reinterpret_cast<const uint4*>(z)[i], d = reinterpret_cast<const uint4*>(df)[i];
    const uint32_t w[4] = {t.x, t.y, t.z, t.w}, dw[4] = {d.x, d.y, d.z, d.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = pack_bf16x2(bf16_bits_to_float(dw[k] & 0xFFFFu) * gelu_tanh_grad_f(bf16_bits_to_float(w[k] & 0xFFFFu)),
                         bf16_bits_to_float(dw[k] >> 16) * gelu_tanh_grad_f(bf16_bits_to_float(w[k] >> 16)));
    reinterpret_cast<uint4*>(dz)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------ RoPE backward
// y = x cos + rotate_half(x) sin is a rotation, so dx = dy cos - rotate_half(dy) sin: the inverse rotation, applied IN PLACE to the
// q and k column blocks of the gradient of the (rotated) qkv tensor bf16 [B*S, (hq + 2 hkv) * d]; position of row t = t % S + 1.
__global__ void svla_rope_bwd_kernel(__nv_bfloat16* __restrict__ dqkv, int s, int hq, int hkv, int d, float theta) {
  const long long tok = blockIdx.x;
  const int pos = static_cast<int>(tok % s) + 1;
  const int half = d >> 1;
  const int heads = hq + hkv;
  __nv_bfloat16* row = dqkv + tok * static_cast<long long>(hq + 2 * hkv) * d;
  for (int i = threadIdx.x; i < heads * half; i += blockDim.x) {
    const int hh = i / half, j = i - hh * half;
    const float inv = powf(theta, -2.f * static_cast<float>(j) / static_cast<float>(d));
    float sn, cs;
    sincosf(static_cast<float>(pos) * inv, &sn, &cs);
    __nv_bfloat16* p = row + hh * d;
    const float a = __bfloat162float(p[j]), b = __bfloat162float(p[j + half]);
    // forward: y1 = x1 c - x2 s, y2 = x2 c + x1 s   ->   dx1 = dy1 c + dy2 s, dx2 = dy2 c - dy1 s
    p[j] = __float2bfloat16(a * cs + b * sn);
    p[j + half] = __float2bfloat16(b * cs - a * sn);
  }
}

// ------------------------------------------------------------------------------------------ row gather / cast
// out_bf16[i, :] = scale * src[row_idx ? row_idx[i] : i, :]   (fp32 -> bf16 GEMM operand; the image-token rows of d(embeddings))
__global__ void svla_rows_cast_kernel(const float* __restrict__ src, const long long* __restrict__ row_idx, float scale, long long rows,
                                      int cols, __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long i_row = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (i_row >= rows) return;
  const long long row = row_idx ? row_idx[i_row] : i_row;
  const int nv = cols >> 2;
  for (int i = lane; i < nv; i += 32) {
    const float4 v = ld4(src + row * cols, i);
    st4_bf16(out + i_row * cols, i, make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale));
  }
}

// ------------------------------------------------------------------------------------------ LoRA operand packing
// One launch per step turns the fp32 master copy of every adapter (the flat arena AdamW updates) into the bf16 operand layouts the
// GEMMs read: descriptor d copies the [rows, cols] fp32 matrix at arena + src_off to pool + dst_off + i * stride_i + j * stride_j
// (transposes, block placement inside the fused q|k|v and gate/up operands and row interleaving are all strides).
struct PackDesc { long long src_off, dst_off, stride_i, stride_j; int rows, cols, tile0, pad; };

__global__ void __launch_bounds__(256)
svla_lora_pack_kernel(const float* __restrict__ arena, __nv_bfloat16* __restrict__ pool, const PackDesc* __restrict__ descs, int n_desc,
                      int total_tiles) {
  __shared__ float tile[32][33];
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    // binary search of the descriptor owning tile t (tile0 = first tile of a descriptor, ascending)
    int lo = 0, hi = n_desc - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].tile0 <= t) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = descs[lo];
    const int tiles_j = (d.cols + 31) >> 5;
    const int lt = t - d.tile0, ti = lt / tiles_j, tj = lt - ti * tiles_j;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = ti * 32 + ty + 8 * r, j = tj * 32 + tx;
      tile[ty + 8 * r][tx] = (i < d.rows && j < d.cols) ? arena[d.src_off + static_cast<long long>(i) * d.cols + j] : 0.f;
    }
    __syncthreads();
    // write with the faster-varying destination index on threadIdx.x
    const bool j_fast = (d.stride_j <= d.stride_i);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int a = ty + 8 * r, b = tx;
      const int li = j_fast ? a : b, lj = j_fast ? b : a;
      const int i = ti * 32 + li, j = tj * 32 + lj;
      if (i < d.rows && j < d.cols) pool[d.dst_off + i * d.stride_i + j * d.stride_j] = __float2bfloat16(tile[li][lj]);
    }
  }
}

}  // namespace

// ================================================================================================== C ABI
static inline unsigned warp_row_blocks(long long rows) { return static_cast<unsigned>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock); }
static inline unsigned ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = static_cast<long long>(svla_num_sms()) * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int svla_rmsnorm_train_fwd(const float* x_in, const float* branch, const float* w_post, const float* w_pre, float eps,
                                      int64_t rows, int cols, float* x_out, void* h_bf16, void* stream) {
  SVLA_REQUIRE(x_in && rows > 0 && cols > 0 && (cols % 4) == 0, "svla_rmsnorm_train_fwd: bad arguments");
  SVLA_REQUIRE((branch == nullptr) == (w_post == nullptr) && (branch == nullptr) == (x_out == nullptr),
               "svla_rmsnorm_train_fwd: branch, w_post and x_out go together");
  SVLA_REQUIRE((w_pre == nullptr) == (h_bf16 == nullptr), "svla_rmsnorm_train_fwd: w_pre and h go together");
  svla_rmsnorm_train_fwd_kernel<<<warp_row_blocks(rows), 32 * kWarpsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      x_in, branch, w_post, w_pre, eps, rows, cols, x_out, static_cast<__nv_bfloat16*>(h_bf16));
  SVLA_LAUNCH_CHECK("svla_rmsnorm_train_fwd");
  return 0;
}

extern "C" int svla_rmsnorm_bwd(const float* x, const float* w, const void* dy, int dy_is_f32, const int64_t* row_idx, float eps,
                                int64_t rows, int cols, float* dx_accum, void* dx_bf16, void* stream) {
  SVLA_REQUIRE(x && w && dy && rows > 0 && cols > 0 && (cols % 4) == 0 && (dx_accum || dx_bf16), "svla_rmsnorm_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dy_is_f32)
    svla_rmsnorm_bwd_kernel<true><<<warp_row_blocks(rows), 32 * kWarpsPerBlock, 0, st>>>(
        x, w, dy, reinterpret_cast<const long long*>(row_idx), eps, rows, cols, dx_accum, static_cast<__nv_bfloat16*>(dx_bf16));
  else
    svla_rmsnorm_bwd_kernel<false><<<warp_row_blocks(rows), 32 * kWarpsPerBlock, 0, st>>>(
        x, w, dy, reinterpret_cast<const long long*>(row_idx), eps, rows, cols, dx_accum, static_cast<__nv_bfloat16*>(dx_bf16));
  SVLA_LAUNCH_CHECK("svla_rmsnorm_bwd");
  return 0;
}

extern "C" int svla_layernorm_bwd(const float* x, const float* gamma, const float* beta, const void* dy_bf16, float eps, int64_t rows,
                                  int cols, int relu, float* dx_accum, void* copy_bf16, void* dx_bf16, void* stream) {
  SVLA_REQUIRE(x && gamma && dy_bf16 && rows > 0 && cols > 0 && (cols % 4) == 0 && (dx_accum || dx_bf16), "svla_layernorm_bwd: bad arguments");
  SVLA_REQUIRE(!relu || beta, "svla_layernorm_bwd: the ReLU mask needs beta");
  SVLA_REQUIRE(!copy_bf16 || dx_accum, "svla_layernorm_bwd: copy_bf16 mirrors dx_accum");
  svla_layernorm_bwd_kernel<<<warp_row_blocks(rows), 32 * kWarpsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      x, gamma, beta, static_cast<const __nv_bfloat16*>(dy_bf16), eps, rows, cols, relu, dx_accum, static_cast<__nv_bfloat16*>(copy_bf16),
      static_cast<__nv_bfloat16*>(dx_bf16));
  SVLA_LAUNCH_CHECK("svla_layernorm_bwd");
  return 0;
}

extern "C" int svla_geglu_fwd(const void* gu, void* act, int64_t rows, int64_t inter, void* stream) {
  SVLA_REQUIRE(gu && act && rows > 0 && inter > 0 && (inter % 4) == 0, "svla_geglu_fwd: bad arguments (inter must be a multiple of 4)");
  const long long n4 = rows * inter / 4;
  svla_geglu_fwd_kernel<<<ew_blocks(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(gu),
                                                                                      static_cast<__nv_bfloat16*>(act), n4);
  SVLA_LAUNCH_CHECK("svla_geglu_fwd");
  return 0;
}
extern "C" int svla_geglu_bwd(const void* gu, const void* dact, void* dgu, int64_t rows, int64_t inter, void* stream) {
  SVLA_REQUIRE(gu && dact && dgu && rows > 0 && inter > 0 && (inter % 4) == 0, "svla_geglu_bwd: bad arguments (inter must be a multiple of 4)");
  const long long n4 = rows * inter / 4;
  svla_geglu_bwd_kernel<<<ew_blocks(n4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(gu), static_cast<const __nv_bfloat16*>(dact), static_cast<__nv_bfloat16*>(dgu), n4);
  SVLA_LAUNCH_CHECK("svla_geglu_bwd");
  return 0;
}
extern "C" int svla_gelu_tanh_fwd(const void* z, void* f, int64_t n, void* stream) {
  SVLA_REQUIRE(z && f && n > 0 && (n % 8) == 0, "svla_gelu_tanh_fwd: n must be a positive multiple of 8");
  svla_gelu_tanh_fwd_kernel<<<ew_blocks(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(z),
                                                                                             static_cast<__nv_bfloat16*>(f), n / 8);
  SVLA_LAUNCH_CHECK("svla_gelu_tanh_fwd");
  return 0;
}
extern "C" int svla_gelu_tanh_bwd(const void* z, const void* df, void* dz, int64_t n, void* stream) {
  SVLA_REQUIRE(z && df && dz && n > 0 && (n % 8) == 0, "svla_gelu_tanh_bwd: n must be a positive multiple of 8");
  svla_gelu_tanh_bwd_kernel<<<ew_blocks(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(df), static_cast<__nv_bfloat16*>(dz), n / 8);
  SVLA_LAUNCH_CHECK("svla_gelu_tanh_bwd");
  return 0;
}

extern "C" int svla_rope_bwd(void* dqkv, int batch, int s, int hq, int hkv, int d, float theta, void* stream) {
  SVLA_REQUIRE(dqkv && batch > 0 && s > 0 && hq > 0 && hkv > 0 && d > 0 && (d % 2) == 0, "svla_rope_bwd: bad arguments");
  svla_rope_bwd_kernel<<<static_cast<unsigned>(batch) * s, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(dqkv), s, hq,
                                                                                                        hkv, d, theta);
  SVLA_LAUNCH_CHECK("svla_rope_bwd");
  return 0;
}

extern "C" int svla_rows_cast(const float* src, const int64_t* row_idx, float scale, int64_t rows, int cols, void* out_bf16, void* stream) {
  SVLA_REQUIRE(src && out_bf16 && rows > 0 && cols > 0 && (cols % 4) == 0, "svla_rows_cast: bad arguments");
  svla_rows_cast_kernel<<<warp_row_blocks(rows), 32 * kWarpsPerBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<const long long*>(row_idx), scale, rows, cols, static_cast<__nv_bfloat16*>(out_bf16));
  SVLA_LAUNCH_CHECK("svla_rows_cast");
  return 0;
}

extern "C" int svla_lora_pack(const float* arena, void* pool_bf16, const void* descs_dev, int n_desc, int total_tiles, void* stream) {
  SVLA_REQUIRE(arena && pool_bf16 && descs_dev && n_desc > 0 && total_tiles > 0, "svla_lora_pack: bad arguments");
  const int cap = svla_num_sms() * 8;
  svla_lora_pack_kernel<<<total_tiles < cap ? total_tiles : cap, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      arena, static_cast<__nv_bfloat16*>(pool_bf16), static_cast<const PackDesc*>(descs_dev), n_desc, total_tiles);
  SVLA_LAUNCH_CHECK("svla_lora_pack");
  return 0;
}

extern "C" int svla_fill_zero(void* ptr, int64_t bytes, void* stream) {
  SVLA_REQUIRE(ptr && bytes >= 0, "svla_fill_zero: bad arguments");
  cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), static_cast<cudaStream_t>(stream));
  SVLA_REQUIRE(e == cudaSuccess, "svla_fill_zero: %s", cudaGetErrorString(e));
  return 0;
}
