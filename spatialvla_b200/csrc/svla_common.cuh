// Shared helpers for the spatialvla_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/spatialvla_b200.h"

void svla_set_error(const char* fmt, ...);
void svla_count_launch(int n = 1);

// Every launch goes through this: records the launch (gpu_launches claim) and maps CUDA errors to the ABI.
#define SVLA_LAUNCH_CHECK(name)                                                          \
  do {                                                                                   \
    svla_count_launch();                                                                 \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      svla_set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));             \
      return -2;                                                                         \
    }                                                                                    \
  } while (0)

#define SVLA_REQUIRE(cond, ...)                                                          \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      svla_set_error(__VA_ARGS__);                                                       \
      return -1;                                                                         \
    }                                                                                    \
  } while (0)

// ---- programmatic dependent launch (PDL) for the launch-bound decode chain ------------------------------------------------
// A kernel launched through svla_launch_pdl may start while its predecessor in the stream is still running.  Contract of
// every such kernel: (1) before griddepcontrol.wait it touches only immutable data (weights, tensor maps) and its own
// shared memory / TMEM; (2) it executes griddepcontrol.wait before the first access to anything an earlier kernel wrote or
// still reads; (3) it executes griddepcontrol.launch_dependents at its top so that the NEXT kernel's prologue (and, for the
// weight-streaming GEMM, its first pipeline stages of weights) overlaps this kernel.  Both instructions are no-ops in a
// kernel that was launched without the attribute.  SVLA_PDL=0 turns the attribute off (plain stream order).
bool svla_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t svla_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = svla_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, for a kernel that runs as thread-block clusters of `cluster_x` CTAs along x (cta_group::2 pairs)
template <typename... KArgs, typename... Args>
static inline cudaError_t svla_launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x,
                                                  Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = svla_pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster_x;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

static inline int svla_num_sms() {
  static int cache[64] = {0};        // per device: a process may drive more than one GPU
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cache[(dev >= 0 && dev < 64) ? dev : 0];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `sh` needs 33 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? sh[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
}
// MUFU tanh (max rel. error 2^-11): below the bf16 rounding of every consumer it is used for
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanh_approx(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, below fp32 GELU rounding for bf16 consumers)
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  float t;      // branch-free reciprocal: __frcp_rn's slow-path branch serialises the 32 independent elements of a chunk
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float p = 1.061405429f;
  p = fmaf(p, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-ax * ax * 1.4426950408889634f));
  const float r = 1.f - p * t * e;
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf_fast(float x) { return 0.5f * x * (1.f + erf_fast(x * 0.7071067811865476f)); }
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_fast(float x) { return x > 15.f ? x : __logf(1.f + __expf(x)); }
