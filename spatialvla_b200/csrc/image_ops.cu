// M9b: observation images on the device (reference model/processing_spatialvla.py:174 -> HF SiglipImageProcessor
// at transformers 4.47: PIL Image.resize((224, 224), BICUBIC) on uint8, x * (1/255) [, (x - mean) / std], channels first).
// The resampler is Pillow's ImagingResample: separable, antialiased (filter support scaled by the down-sampling ratio), 22-bit
// fixed-point coefficients, horizontal pass first, each pass rounded and clipped to uint8 -- integer arithmetic, so the device
// result is bit-identical.  The coefficient tables (bounds / kk per output column and row) are built on the host per input size
// (spatialvla_b200/image_processing.py); the uint8 -> float32 value map (rescale, optional normalisation: 256 x 3 floats computed
// on the host with numpy's own arithmetic) is a lookup table, so the floats are the reference's bit for bit as well.
// HBM-bound: 3 B per input pixel in, 12 B per output pixel out.
#include "svla_common.cuh"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: one CTA per (image row, image); the source row is staged in shared memory with coalesced 16-byte loads.
__global__ void __launch_bounds__(256)
svla_resize_h_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ tmp, const int* __restrict__ bounds,
                     const int* __restrict__ kk, int ksize, int h, int w, int ow) {
  extern __shared__ __align__(16) unsigned char srow[];
  const int y = blockIdx.x, b = blockIdx.y;
  const long long row_bytes = static_cast<long long>(w) * 3;
  const unsigned char* src = in + (static_cast<long long>(b) * h + y) * row_bytes;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n16 = static_cast<int>(row_bytes >> 4);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) reinterpret_cast<uint4*>(srow)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    for (int i = (n16 << 4) + threadIdx.x; i < row_bytes; i += blockDim.x) srow[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) srow[i] = src[i];
  }
  __syncthreads();
  unsigned char* dst = tmp + (static_cast<long long>(b) * h + y) * ow * 3;
  for (int i = threadIdx.x; i < ow * 3; i += blockDim.x) {
    const int ox = i / 3, c = i - ox * 3;
    const int x0 = bounds[2 * ox], n = bounds[2 * ox + 1];
    const int* k = kk + ox * ksize;
    int acc = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < n; ++t) acc += static_cast<int>(srow[(x0 + t) * 3 + c]) * __ldg(k + t);
    dst[i] = clip8(acc);
  }
}

// Vertical pass + value map + HWC -> CHW: one CTA per (output row, image); thread = (channel, output column), so the fp32 stores
// of one channel row are contiguous.  tmp rows are re-read by ~ksize/scale neighbouring CTAs out of L2.
__global__ void __launch_bounds__(256)
svla_resize_v_kernel(const unsigned char* __restrict__ tmp, float* __restrict__ out, const int* __restrict__ bounds,
                     const int* __restrict__ kk, int ksize, const float* __restrict__ lut, int h, int ow, int oh) {
  __shared__ float slut[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) slut[i] = lut[i];
  __syncthreads();
  const int oy = blockIdx.x, b = blockIdx.y;
  const int y0 = bounds[2 * oy], n = bounds[2 * oy + 1];
  const int* k = kk + oy * ksize;
  const unsigned char* src = tmp + (static_cast<long long>(b) * h + y0) * ow * 3;
  for (int i = threadIdx.x; i < 3 * ow; i += blockDim.x) {
    const int c = i / ow, ox = i - c * ow;
    int acc = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < n; ++t) acc += static_cast<int>(src[static_cast<long long>(t) * ow * 3 + ox * 3 + c]) * __ldg(k + t);
    out[((static_cast<long long>(b) * 3 + c) * oh + oy) * ow + ox] = slut[c * 256 + clip8(acc)];
  }
}

// No resampling on an axis (input already at the output size): identity "pass" so the two-kernel structure stays uniform.
__global__ void svla_u8_to_chw_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, const float* __restrict__ lut,
                                      int oh, int ow) {
  __shared__ float slut[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) slut[i] = lut[i];
  __syncthreads();
  const int oy = blockIdx.x, b = blockIdx.y;
  const unsigned char* src = in + (static_cast<long long>(b) * oh + oy) * ow * 3;
  for (int i = threadIdx.x; i < 3 * ow; i += blockDim.x) {
    const int c = i / ow, ox = i - c * ow;
    out[((static_cast<long long>(b) * 3 + c) * oh + oy) * ow + ox] = slut[c * 256 + src[ox * 3 + c]];
  }
}

// Barycentric gather: out[t, :] = sum_v w[t, v] * src[rows[t, v], :] (double accumulation), NaN rows where rows[t, 0] < 0.
__global__ void __launch_bounds__(256)
svla_barycentric_gather_kernel(const float* __restrict__ src, const int* __restrict__ rows, const double* __restrict__ w,
                               float* __restrict__ out, int e) {
  const long long t = blockIdx.x;
  const int r0 = rows[4 * t], r1 = rows[4 * t + 1], r2 = rows[4 * t + 2], r3 = rows[4 * t + 3];
  float* dst = out + t * e;
  if (r0 < 0) {
    for (int i = threadIdx.x; i < e; i += blockDim.x) dst[i] = __int_as_float(0x7fc00000);
    return;
  }
  const double w0 = w[4 * t], w1 = w[4 * t + 1], w2 = w[4 * t + 2], w3 = w[4 * t + 3];
  const float *s0 = src + static_cast<long long>(r0) * e, *s1 = src + static_cast<long long>(r1) * e,
              *s2 = src + static_cast<long long>(r2) * e, *s3 = src + static_cast<long long>(r3) * e;
  for (int i = threadIdx.x; i < e; i += blockDim.x) {
    // same order as scipy's LinearNDInterpolator: out = sum_j c[j] * values[simplex[j]] accumulated from zero in double
    double acc = __dmul_rn(w0, static_cast<double>(s0[i]));
    acc = __dadd_rn(acc, __dmul_rn(w1, static_cast<double>(s1[i])));
    acc = __dadd_rn(acc, __dmul_rn(w2, static_cast<double>(s2[i])));
    acc = __dadd_rn(acc, __dmul_rn(w3, static_cast<double>(s3[i])));
    dst[i] = static_cast<float>(acc);
  }
}

}  // namespace

extern "C" int svla_barycentric_gather(const float* src, int64_t n_src, const int32_t* rows, const double* weights, float* out,
                                       int64_t n_out, int e, void* stream) {
  SVLA_REQUIRE(src && rows && weights && out && n_src > 0 && n_out > 0 && e > 0, "svla_barycentric_gather: bad arguments");
  svla_barycentric_gather_kernel<<<static_cast<unsigned>(n_out), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, rows, weights, out, e);
  SVLA_LAUNCH_CHECK("svla_barycentric_gather");
  return 0;
}

extern "C" int svla_image_preprocess(const void* images_u8, int batch, int h, int w, void* tmp_u8, float* out, int oh, int ow,
                                     const int32_t* bounds_h, const int32_t* kk_h, int ksize_h, const int32_t* bounds_v,
                                     const int32_t* kk_v, int ksize_v, const float* value_lut, void* stream) {
  SVLA_REQUIRE(images_u8 && out && value_lut && batch > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "svla_image_preprocess: bad arguments");
  SVLA_REQUIRE(batch <= 65535, "svla_image_preprocess: batch %d > 65535", batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* cur = static_cast<const unsigned char*>(images_u8);
  if (w != ow) {
    SVLA_REQUIRE(tmp_u8 && bounds_h && kk_h && ksize_h > 0, "svla_image_preprocess: horizontal pass needs tmp / tables");
    const size_t smem = (static_cast<size_t>(w) * 3 + 15) / 16 * 16;
    SVLA_REQUIRE(smem <= 96 * 1024, "svla_image_preprocess: image rows of %d pixels are not supported", w);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(svla_resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
    svla_resize_h_kernel<<<dim3(h, batch), 256, smem, st>>>(cur, static_cast<unsigned char*>(tmp_u8), bounds_h, kk_h, ksize_h, h, w, ow);
    SVLA_LAUNCH_CHECK("svla_resize_h");
    cur = static_cast<const unsigned char*>(tmp_u8);
  }
  if (h != oh) {
    SVLA_REQUIRE(bounds_v && kk_v && ksize_v > 0, "svla_image_preprocess: vertical pass needs tables");
    svla_resize_v_kernel<<<dim3(oh, batch), 256, 0, st>>>(cur, out, bounds_v, kk_v, ksize_v, value_lut, h, ow, oh);
    SVLA_LAUNCH_CHECK("svla_resize_v");
  } else {
    svla_u8_to_chw_kernel<<<dim3(oh, batch), 256, 0, st>>>(cur, out, value_lut, oh, ow);
    SVLA_LAUNCH_CHECK("svla_u8_to_chw");
  }
  return 0;
}
