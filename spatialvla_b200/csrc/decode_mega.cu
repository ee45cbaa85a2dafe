// G4: ONE decode step of the whole Gemma2 stack in ONE persistent launch (batch <= 64) -- the tensor-core, any-batch successor
// of the CUDA-core small-batch experiment of round 1.  OPT-IN (SVLA_DECODE=mega): measured on B200 it is bit-identical to the
// 7-kernels-per-layer PDL chain (gemm_skinny.cu + attention.cu + fused_ops.cu) but slower (2.43-2.57 ms against 1.99 ms per
// batch-64 step, profiles/decode_mega_r2.txt), see "What the measurement says" below.
//
// A decode step at batch 64 is HBM-bound on paper: 4.05 GB of layer weights + 1.9 GB of KV cache, ~0.9 ms at the measured copy
// bandwidth; the chain needs 2 ms because each of its 182 kernels pays ~5 us of dependency latency during which HBM idles.
// The dependencies themselves cannot go away (every norm is a full-row reduction), so this kernel keeps the seven phases of a
// layer but takes the WEIGHT STREAM out of the dependency chain:
//
//   * one CTA per SM (cooperative launch: co-residency is checked by the driver), 16 warps with fixed roles;
//   * warp 0 = weight producer: walks the CTA's static schedule of [128 rows x 64 k] weight tiles of ALL phases of ALL
//     layers and keeps a 7-slot TMA ring (112 KB per SM, 16.6 MB chip-wide) full.  It waits for ring slots only, never for a
//     phase boundary: while a grid barrier, a norm or the attention phase resolves, the next GEMM's weights are already
//     arriving;
//   * warp 2 = activation producer: [64 rows x 64 k] tiles of the phase's input (L2-hot) through a second, 4-slot ring; it
//     does wait for the phase boundary;
//   * warp 1 = MMA issuer: swap-AB tcgen05.mma 128 x 64 x 16 (weights on the M side), fp32 accumulators double-buffered in
//     TMEM; split-K exactly as svla_gemm_skinny (n_tiles x splits ~ #SMs);
//   * warps 4-7 = epilogue: TMEM -> registers -> split-K partial sums / GeGLU, then the phase arrival;
//   * warps 8-15 = 256 workers for the two phases that are not GEMMs: RoPE + KV append + soft-capped attention of one
//     (batch row, kv head) item per pass (decode_attn_item.cuh, shared with the chain's kernel), and the sandwich norms (one
//     batch row per CTA);
//   * warp 3 lane 0 = barrier watcher: polls the global arrival counter and publishes the completed phase number in shared
//     memory; the roles wait on that word (with nanosleep back-off), not on L2.
//
// Phases (grid barrier after each):  N0 | per layer: QKV, ATTN, O, NORM1, GATE/UP(+GeGLU), DOWN, NORM2.
// Numerics are those of the chain: bf16 operands, fp32 accumulation, split-K partials summed in split order, fp32 residual
// stream and statistics with the same reduction trees -- the two paths produce bit-identical hidden states and cache rows
// (tests/test_e2e_gpu.py::test_persistent_decode_kernel_matches_chain, tools/decode_mega_check.py).
//
// What the measurement says (layer 5, CTA 0, globaltimer stamps): ~94 us per layer = attention 28 us (256 items on 148 CTAs:
// two rounds, and one 8-warp worker group per SM where the chain has two resident CTAs) + gate/up 18 + down 11 + qkv 9 + o 6 +
// norms 2 x 4.5 + 7 barriers x 1.3-1.9 us.  The ring does its job -- the o projection's weights are resident 30 us before its
// phase starts, gate/up and down stream at 6.5 TB/s -- but an in-kernel grid barrier (bar.sync + fence.proxy.async +
// red.release, poll, fence.acq_rel, shared-memory publish, ~2.5 us end to end) costs what a PDL kernel boundary costs, and
// each GEMM phase still pays ~2.3 us from the epoch to its first accumulator plus ~3 us of epilogue.  Seven DEPENDENT phases per
// layer bound the step in both designs; launch overhead was never the limiter.
// Reference ops replaced: model/modeling_gemma2.py:80-92 (MLP), :169-195,351-413 (attention), :451-506 (decoder layer).
#include <cstdlib>
#include <cstring>
#include <cudaTypedefs.h>
#include "tc_ptx.cuh"
#include "decode_attn_item.cuh"

namespace {
using namespace svla_ptx;

constexpr int kWM = 128;           // weight rows per tile (UMMA M)
constexpr int kBK = 64;            // K elements per tile (128 B rows, SWIZZLE_128B)
constexpr int kNB = 64;            // activation rows (UMMA N): batch <= 64
constexpr int kD = 256;            // head dim
constexpr int kWBytes = kWM * kBK * 2;
constexpr int kXBytes = kNB * kBK * 2;
constexpr int kWStages = 7;
constexpr int kXStages = 4;
constexpr int kKvStages = 4, kKvRows = 32, kKvPitch = 528;
constexpr int kThreads = 512;
constexpr int kWorkers = 256;
constexpr int kMaxCtxPad = 1024;
constexpr int kTmemCols = 2 * kNB;

// shared-memory map (offsets from the 1024-aligned base)
constexpr int kOffW = 0;
constexpr int kOffX = kOffW + kWStages * kWBytes;
constexpr int kOffKv = kOffX + kXStages * kXBytes;
constexpr int kOffQ = kOffKv + kKvStages * kKvRows * kKvPitch;      // float [2][256]
constexpr int kOffAcc = kOffQ + 2 * kD * 4;                          // float [2][256]
constexpr int kOffNewK = kOffAcc + 2 * kD * 4;                       // bf16 [256]
constexpr int kOffNewV = kOffNewK + kD * 2;                          // bf16 [256]
constexpr int kOffInv = kOffNewV + kD * 2;                           // float [4]
constexpr int kOffRed = kOffInv + 16;                                // float [16] per-warp reduction slots
constexpr int kOffP = kOffRed + 64;                                  // float [2][ctx_pad]
constexpr int kOffBars = kOffP + 2 * kMaxCtxPad * 4;
constexpr int kNumBars = 2 * kWStages + 2 * kXStages + 4;
constexpr int kOffMisc = kOffBars + kNumBars * 8;                    // tmem slot, epoch word
constexpr int kSmemBytes = kOffMisc + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert((kOffQ % 16) == 0 && (kOffP % 16) == 0 && (kOffBars % 8) == 0, "alignment");

struct GemmPhase { int n, k, n_tiles, splits, kbps, nkb; };

struct MegaParams {
  const CUtensorMap* maps;        // [4 * n_layers] weight maps (qkv, o, gate/up, down per layer) + [3] activation maps (h, ctx, act)
  const float* const* norm_w;     // [4 * n_layers]: ln_in, ln_post_attn, ln_pre_ff, ln_post_ff
  int n_layers;
  float* x;                       // [batch][H] residual stream (in/out)
  const float* final_w;
  __nv_bfloat16* h_out;           // [batch][H]
  __nv_bfloat16* kcache; __nv_bfloat16* vcache;
  long long cache_layer_stride;   // elements between layers of the cache
  __nv_bfloat16* h; __nv_bfloat16* ctxb; __nv_bfloat16* act;
  float* part;
  unsigned* counter;
  unsigned long long* timing;     // optional [64] phase timestamps of CTA 0, layer timing_layer
  const int* kv_start;
  GemmPhase g[4];
  int batch, H, hq, hkv, FF, smax, ctx, timing_layer;
  float theta, scale, softcap, eps;
};

struct Unit { int n_tile, split, kb0, nkb, rot; };
__device__ __forceinline__ Unit make_unit(const GemmPhase& g, int u) {
  Unit r;
  r.n_tile = u % g.n_tiles;
  r.split = u / g.n_tiles;
  r.kb0 = r.split * g.kbps;
  const int kb1 = min(r.kb0 + g.kbps, g.nkb);
  r.nkb = kb1 - r.kb0;
  // every CTA walks its K range from a different block so that they do not all ask L2 for the same activation tile at once
  r.rot = r.nkb > 0 ? (r.n_tile * 5) % r.nkb : 0;
  return r;
}
// global phase index of GEMM g of layer l (phase 0 is the initial norm)
__device__ __forceinline__ int gemm_phase(int l, int g) { return 1 + 7 * l + (g == 0 ? 0 : (g == 1 ? 2 : (g == 2 ? 4 : 5))); }

__device__ __forceinline__ void wsync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }        // the 256 workers
__device__ __forceinline__ void esync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }        // the 4 epilogue warps
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Waiting roles must not burn issue slots: with 16 warps on 4 schedulers a hot try_wait / shared-memory spin loop takes every
// other issue slot from the warp that does the phase's work (measured: the epilogue's 64 stores took 3 us, one attention item
// 14 us).  Every wait of this kernel therefore backs off with nanosleep between polls.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  unsigned long long t0 = 0;
#pragma unroll 1
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    __nanosleep(32);
    if ((spin & 1023u) == 1023u) {
      const unsigned long long now = now_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { printf("svla_decode_mega: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
  }
}

// phase `q` may start when every CTA has finished phase q-1: the watcher publishes the number of completed phases
__device__ __forceinline__ void wait_epoch(const int* s_epoch, int q) {
  if (q <= 0) return;
  const uint32_t a = smem_u32(s_epoch);
  unsigned long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    if (v >= q) return;
    __nanosleep(64);
    if ((spin & 4095u) == 4095u) {
      const unsigned long long now = now_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { printf("svla_decode_mega: phase %d never completed (block %d thread %d)\n", q - 1, blockIdx.x, threadIdx.x); __trap(); }
    }
  }
}
// one arrival per CTA and phase; the release orders this CTA's phase output (made visible to the arriving thread by the
// role's barrier) before the count
__device__ __forceinline__ void arrive_phase(unsigned* counter) {
  fence_proxy_async_all();        // generic-proxy stores of this phase -> visible to the TMA loads of the next one
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}

__device__ __forceinline__ float worker_sum(float v, float* s_red, int t) {
  v = warp_sum(v);
  wsync();
  if ((t & 31) == 0) s_red[t >> 5] = v;
  wsync();
  // the summation tree of block_sum() in svla_common.cuh (8 warp partials, butterfly 4-2-1): bit-identical to the chain's kernel
  static_assert(kWorkers == 256, "eight warp partials");
  return ((s_red[0] + s_red[4]) + (s_red[2] + s_red[6])) + ((s_red[1] + s_red[5]) + (s_red[3] + s_red[7]));
}

// x_row += rms(sum of split-K partial rows)(1 + w_post)   [part != null];   out = bf16(rms(x_row)(1 + w_pre))
// (model/modeling_gemma2.py:60-77,475-496; the arithmetic and summation order of svla_rmsnorm_residual_kernel)
template <int V>
__device__ __forceinline__ void norm_row(float* xrow, const float* part, int n_part, long long pstride, const float* w_post,
                                         const float* w_pre, __nv_bfloat16* out, int H, float eps, float* s_red, int t) {
  const int nv = H >> 2;
  float4 xv[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int i = t + k * kWorkers;
    xv[k] = i < nv ? __ldcg(reinterpret_cast<const float4*>(xrow) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (part) {
    float4 bv[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = t + k * kWorkers;
      bv[k] = i < nv ? __ldcg(reinterpret_cast<const float4*>(part) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int sp = 1; sp < n_part; sp += 4) {
      float4 pv[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const int i = t + k * kWorkers;
          pv[u][k] = (sp + u < n_part && i < nv) ? __ldcg(reinterpret_cast<const float4*>(part + (sp + u) * pstride) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < V; ++k) { bv[k].x += pv[u][k].x; bv[k].y += pv[u][k].y; bv[k].z += pv[u][k].z; bv[k].w += pv[u][k].w; }
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) ss += bv[k].x * bv[k].x + bv[k].y * bv[k].y + bv[k].z * bv[k].z + bv[k].w * bv[k].w;
    const float r = rsqrtf(worker_sum(ss, s_red, t) / H + eps);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = t + k * kWorkers;
      if (i < nv) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(w_post) + i);
        xv[k].x += bv[k].x * r * (1.f + w.x);
        xv[k].y += bv[k].y * r * (1.f + w.y);
        xv[k].z += bv[k].z * r * (1.f + w.z);
        xv[k].w += bv[k].w * r * (1.f + w.w);
        reinterpret_cast<float4*>(xrow)[i] = xv[k];
      }
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) ss += xv[k].x * xv[k].x + xv[k].y * xv[k].y + xv[k].z * xv[k].z + xv[k].w * xv[k].w;
  const float r = rsqrtf(worker_sum(ss, s_red, t) / H + eps);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int i = t + k * kWorkers;
    if (i < nv) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(w_pre) + i);
      reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(xv[k].x * r * (1.f + w.x), xv[k].y * r * (1.f + w.y)),
                                                    pack_bf16x2(xv[k].z * r * (1.f + w.z), xv[k].w * r * (1.f + w.w)));
    }
  }
}

__device__ __forceinline__ void norm_phase(const MegaParams& p, const float* part, int n_part, const float* w_post, const float* w_pre,
                                           __nv_bfloat16* out, float* s_red, int t) {
  const long long pstride = static_cast<long long>(kNB) * p.H;
  for (int row = blockIdx.x; row < p.batch; row += gridDim.x) {
    float* xrow = p.x + static_cast<long long>(row) * p.H;
    const float* prow = part ? part + static_cast<long long>(row) * p.H : nullptr;
    __nv_bfloat16* orow = out + static_cast<long long>(row) * p.H;
    const int need = ((p.H >> 2) + kWorkers - 1) / kWorkers;
    if (need <= 1) norm_row<1>(xrow, prow, n_part, pstride, w_post, w_pre, orow, p.H, p.eps, s_red, t);
    else if (need <= 3) norm_row<3>(xrow, prow, n_part, pstride, w_post, w_pre, orow, p.H, p.eps, s_red, t);
    else norm_row<4>(xrow, prow, n_part, pstride, w_post, w_pre, orow, p.H, p.eps, s_red, t);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
svla_decode_mega_kernel(const MegaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem + kOffW;
  uint8_t* sX = smem + kOffX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + kWStages;
  uint64_t* x_full = bars + 2 * kWStages;
  uint64_t* x_empty = x_full + kXStages;
  uint64_t* acc_full = x_empty + kXStages;        // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffMisc);
  int* s_epoch = reinterpret_cast<int*>(smem + kOffMisc + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int n_phases = 1 + 7 * p.n_layers;
  const CUtensorMap* xmaps = p.maps + 4 * p.n_layers;      // h, ctx, act

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < kXStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    *s_epoch = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemCols, 1>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ weight producer: never waits for a phase boundary
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int l = 0; l < p.n_layers; ++l) {
        for (int g = 0; g < 4; ++g) {
          const GemmPhase gp = p.g[g];
          const CUtensorMap* tm = p.maps + l * 4 + g;
          for (int u = blockIdx.x; u < gp.n_tiles * gp.splits; u += G) {
            const Unit un = make_unit(gp, u);
            for (int it = 0; it < un.nkb; ++it) {
              const int kb = un.kb0 + (it + un.rot) % un.nkb;
              mbar_wait_backoff(&w_empty[slot], ph ^ 1u);
              mbar_expect_tx(&w_full[slot], kWBytes);
              tma_load_2d(sW + slot * kWBytes, tm, &w_full[slot], kb * kBK, un.n_tile * kWM);
              if (++slot == kWStages) { slot = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================================================ activation producer
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int l = 0; l < p.n_layers; ++l) {
        for (int g = 0; g < 4; ++g) {
          const GemmPhase gp = p.g[g];
          if (static_cast<int>(blockIdx.x) >= gp.n_tiles * gp.splits) continue;
          const CUtensorMap* tm = xmaps + (g == 1 ? 1 : (g == 3 ? 2 : 0));
          wait_epoch(s_epoch, gemm_phase(l, g));
          fence_proxy_async_all();
          for (int u = blockIdx.x; u < gp.n_tiles * gp.splits; u += G) {
            const Unit un = make_unit(gp, u);
            for (int it = 0; it < un.nkb; ++it) {
              const int kb = un.kb0 + (it + un.rot) % un.nkb;
              mbar_wait_backoff(&x_empty[slot], ph ^ 1u);
              mbar_expect_tx(&x_full[slot], kXBytes);
              tma_load_2d(sX + slot * kXBytes, tm, &x_full[slot], kb * kBK, 0);
              if (++slot == kXStages) { slot = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kWM, kNB);
      int ws = 0, xs = 0, buf = 0;
      uint32_t wph = 0, xph = 0, aph = 0;
      for (int l = 0; l < p.n_layers; ++l) {
        for (int g = 0; g < 4; ++g) {
          const GemmPhase gp = p.g[g];
          for (int u = blockIdx.x; u < gp.n_tiles * gp.splits; u += G) {
            const Unit un = make_unit(gp, u);
            mbar_wait_backoff(&acc_empty[buf], aph ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * kNB;
            for (int it = 0; it < un.nkb; ++it) {
              mbar_wait_backoff(&w_full[ws], wph);
              mbar_wait_backoff(&x_full[xs], xph);
              tc_fence_after();
              const uint64_t dw = make_kmajor_sw128_desc(smem_u32(sW + ws * kWBytes));
              const uint64_t dx = make_kmajor_sw128_desc(smem_u32(sX + xs * kXBytes));
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                umma_bf16(tmem_d, dw + static_cast<uint64_t>(k * 2), dx + static_cast<uint64_t>(k * 2), idesc, static_cast<uint32_t>(it > 0 || k != 0));
              umma_commit(&w_empty[ws]);
              umma_commit(&x_empty[xs]);
              if (++ws == kWStages) { ws = 0; wph ^= 1u; }
              if (++xs == kXStages) { xs = 0; xph ^= 1u; }
            }
            umma_commit(&acc_full[buf]);
            if (++buf == 2) { buf = 0; aph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ================================================================ barrier watcher
    if (lane == 0) {
      for (int q = 0; q + 1 < n_phases; ++q) {
        const unsigned target = static_cast<unsigned>(q + 1) * static_cast<unsigned>(G);
        unsigned v;
        unsigned long long t0 = 0;
        for (unsigned spin = 0;; ++spin) {
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.counter) : "memory");
          if (v >= target) break;
          if ((spin & 4095u) == 4095u) {
            const unsigned long long now = now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) { printf("svla_decode_mega: grid barrier %d timed out (block %d, count %u)\n", q, blockIdx.x, v); __trap(); }
          }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(s_epoch)), "r"(q + 1) : "memory");
      }
    }
  } else if (warp < 8) {
    // ================================================================ epilogue: lane <-> weight row n, register j <-> batch row m
    const int q4 = warp & 3;
    int buf = 0;
    uint32_t aph = 0;
    for (int l = 0; l < p.n_layers; ++l) {
      for (int g = 0; g < 4; ++g) {
        const GemmPhase gp = p.g[g];
        const int phase = gemm_phase(l, g);
        // an arrival may only be counted once the previous phase is complete (CTAs without a unit would otherwise run ahead)
        if (lane == 0) wait_epoch(s_epoch, phase);
        __syncwarp();
        const bool tme = p.timing && blockIdx.x == 0 && l == p.timing_layer && warp == 4 && lane == 0;
        if (tme) p.timing[20 + 4 * g] = now_ns();
        for (int u = blockIdx.x; u < gp.n_tiles * gp.splits; u += G) {
          const Unit un = make_unit(gp, u);
          const long long n = static_cast<long long>(un.n_tile) * kWM + q4 * 32 + lane;
          const bool n_ok = n < gp.n;
          mbar_wait_backoff(&acc_full[buf], aph);
          tc_fence_after();
          if (tme) p.timing[21 + 4 * g] = now_ns();
          const uint32_t taddr = tmem_base + buf * kNB + (static_cast<uint32_t>(q4 * 32) << 16);
          float* pout = p.part + static_cast<long long>(un.split) * kNB * gp.n;
          // predicated stores, no data-dependent exit inside the unrolled row loop: the 32 rows of a chunk overlap
#pragma unroll 1
          for (int c0 = 0; c0 < kNB; c0 += 32) {
            if (c0 >= p.batch) break;                       // warp-uniform
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            const int rows = p.batch - c0;
            if (g == 2) {                                   // GeGLU: rows 2j = gate_j, 2j+1 = up_j
              __nv_bfloat16* dst = p.act + static_cast<long long>(c0) * p.FF + (n >> 1);
              const bool wr = n_ok && (lane & 1) == 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                const float other = __shfl_xor_sync(0xffffffffu, v, 1);
                const __nv_bfloat16 o = __float2bfloat16(gelu_tanh_fast(v) * other);
                if (wr && j < rows) dst[static_cast<long long>(j) * p.FF] = o;
              }
            } else {
              float* dst = pout + static_cast<long long>(c0) * gp.n + n;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n_ok && j < rows) dst[static_cast<long long>(j) * gp.n] = __uint_as_float(r[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
          if (++buf == 2) { buf = 0; aph ^= 1u; }
        }
        if (tme) p.timing[22 + 4 * g] = now_ns();
        esync();
        if (warp == 4 && lane == 0) {
          arrive_phase(p.counter);
          if (p.timing && blockIdx.x == 0 && l == p.timing_layer) p.timing[8 + g] = now_ns();
        }
      }
    }
  } else {
    // ================================================================ workers: norms and attention
    const int t = threadIdx.x - (kThreads - kWorkers);
    float* s_red = reinterpret_cast<float*>(smem + kOffRed);
    const int grp = p.hq / p.hkv;
    auto finish = [&](int slot_, int l_) {
      wsync();
      if (t == 0) {
        arrive_phase(p.counter);
        if (p.timing && blockIdx.x == 0 && l_ == p.timing_layer) p.timing[slot_] = now_ns();
      }
    };
    // phase 0: h = bf16(rms(x)(1 + ln_in[0]))
    norm_phase(p, nullptr, 0, nullptr, p.norm_w[0], p.h, s_red, t);
    finish(0, -2);
    for (int l = 0; l < p.n_layers; ++l) {
      const int base = 1 + 7 * l;
      // ---- ATTN
      if (t == 0) wait_epoch(s_epoch, base + 1);
      wsync();
      if (t == 0 && p.timing && blockIdx.x == 0 && l == p.timing_layer) p.timing[1] = now_ns();
      {
        __nv_bfloat16* kc = p.kcache + static_cast<long long>(l) * p.cache_layer_stride;
        __nv_bfloat16* vc = p.vcache + static_cast<long long>(l) * p.cache_layer_stride;
        svla_dec::ItemSmem sm;
        sm.stage = smem + kOffKv;
        sm.q = reinterpret_cast<float*>(smem + kOffQ);
        sm.red = reinterpret_cast<float*>(smem + kOffAcc);
        sm.newk = reinterpret_cast<__nv_bfloat16*>(smem + kOffNewK);
        sm.newv = reinterpret_cast<__nv_bfloat16*>(smem + kOffNewV);
        sm.inv = reinterpret_cast<float*>(smem + kOffInv);
        sm.wred = s_red;
        sm.p = reinterpret_cast<float*>(smem + kOffP);
        const long long width = static_cast<long long>(p.hq + 2 * p.hkv) * kD;
        for (int item = blockIdx.x; item < p.batch * p.hkv; item += G) {
          svla_dec::ItemArgs ia;
          ia.b = item / p.hkv; ia.hk = item % p.hkv;
          ia.qkv = p.part + static_cast<long long>(ia.b) * width;
          ia.n_partials = p.g[0].splits;
          ia.partial_stride = static_cast<long long>(kNB) * width;
          ia.kc = kc; ia.vc = vc; ia.out = p.ctxb; ia.out_lo = nullptr;
          ia.hq = p.hq; ia.hkv = p.hkv; ia.smax = p.smax; ia.ctx = p.ctx;
          ia.kstart = p.kv_start ? p.kv_start[ia.b] : 0;
          ia.kmask = ia.kstart;
          ia.theta = p.theta; ia.scale = p.scale; ia.softcap = p.softcap;
          if (grp == 2) svla_dec::decode_attn_item<2, kKvStages>(ia, sm, t, wsync, [] {});
          else svla_dec::decode_attn_item<1, kKvStages>(ia, sm, t, wsync, [] {});
          wsync();        // the shared buffers are reused by the next item
          if (t == 0 && p.timing && blockIdx.x == 0 && l == p.timing_layer && item == 0) p.timing[19] = now_ns();
        }
      }
      finish(2, l);
      // ---- NORM1: x += rms(o)(1 + ln_post_attn); h = bf16(rms(x)(1 + ln_pre_ff))
      if (t == 0) wait_epoch(s_epoch, base + 3);
      wsync();
      if (t == 0 && p.timing && blockIdx.x == 0 && l == p.timing_layer) p.timing[3] = now_ns();
      norm_phase(p, p.part, p.g[1].splits, p.norm_w[4 * l + 1], p.norm_w[4 * l + 2], p.h, s_red, t);
      finish(4, l);
      // ---- NORM2: x += rms(down)(1 + ln_post_ff); h = bf16(rms(x)(1 + next ln_in | final norm))
      if (t == 0) wait_epoch(s_epoch, base + 6);
      wsync();
      if (t == 0 && p.timing && blockIdx.x == 0 && l == p.timing_layer) p.timing[5] = now_ns();
      const bool last = l + 1 == p.n_layers;
      norm_phase(p, p.part, p.g[3].splits, p.norm_w[4 * l + 3], last ? p.final_w : p.norm_w[4 * (l + 1)], last ? p.h_out : p.h, s_red, t);
      finish(6, l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols, 1>(tmem_base);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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

int encode_kmajor(CUtensorMap* tm, const void* base, uint64_t k, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {kBK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

// split-K plan of one projection: the rule of svla_gemm_skinny_splits (n_tiles x splits ~ #SMs, >= 4 K steps per CTA)
GemmPhase plan_gemm(int n, int k, bool allow_split) {
  GemmPhase g;
  g.n = n; g.k = k;
  g.n_tiles = (n + kWM - 1) / kWM;
  g.nkb = (k + kBK - 1) / kBK;
  int splits = allow_split ? svla_num_sms() / g.n_tiles : 1;
  if (splits < 1) splits = 1;
  while (splits > 1 && (g.nkb + splits - 1) / splits < 4) --splits;
  g.kbps = (g.nkb + splits - 1) / splits;
  g.splits = (g.nkb + g.kbps - 1) / g.kbps;
  return g;
}

struct Layout {
  GemmPhase g[4];
  size_t off_h, off_ctx, off_act, off_part, off_counter, off_timing, total;
};
Layout make_layout(int hidden, int hq, int hkv, int d, int ff) {
  Layout L;
  L.g[0] = plan_gemm((hq + 2 * hkv) * d, hidden, true);
  L.g[1] = plan_gemm(hidden, hq * d, true);
  L.g[2] = plan_gemm(2 * ff, hidden, false);
  L.g[3] = plan_gemm(hidden, ff, true);
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o = 0;
  L.off_h = o; o = al(o + static_cast<size_t>(kNB) * hidden * 2);
  L.off_ctx = o; o = al(o + static_cast<size_t>(kNB) * hq * d * 2);
  L.off_act = o; o = al(o + static_cast<size_t>(kNB) * ff * 2);
  size_t part = 0;
  for (int i : {0, 1, 3}) {
    const size_t b = static_cast<size_t>(L.g[i].splits) * kNB * L.g[i].n * 4;
    if (b > part) part = b;
  }
  L.off_part = o; o = al(o + part);
  L.off_counter = o; o = al(o + 256);
  L.off_timing = o; o = al(o + 64 * 8);
  L.total = o;
  return L;
}

}  // namespace

extern "C" int64_t svla_decode_mega_scratch_bytes(int hidden, int hq, int hkv, int d, int ff) {
  return static_cast<int64_t>(make_layout(hidden, hq, hkv, d, ff).total);
}
extern "C" int64_t svla_decode_mega_maps_bytes(int n_layers) { return static_cast<int64_t>(4 * n_layers + 3) * static_cast<int64_t>(sizeof(CUtensorMap)); }

extern "C" int svla_decode_mega_supported(int batch, int hidden, int hq, int hkv, int d, int ff, int ctx) {
  return (batch >= 1 && batch <= kNB && d == kD && hkv > 0 && hq % hkv == 0 && (hq / hkv == 1 || hq / hkv == 2) && (hidden % 8) == 0 &&
          hidden <= 4 * 4 * kWorkers && (ff % 8) == 0 && ctx >= 1 && ((ctx + 31) & ~31) <= kMaxCtxPad) ? 1 : 0;
}

// Host side, once per engine: the TMA descriptors of the 4 weight matrices of every layer and of the three activation buffers
// inside `scratch_dev`, written to HOST memory `maps_host` (svla_decode_mega_maps_bytes); the caller copies them to the device.
extern "C" int svla_decode_mega_plan(void* maps_host, const void* const* weights, int n_layers, int hidden, int hq, int hkv, int d, int ff,
                                     void* scratch_dev) {
  SVLA_REQUIRE(maps_host && weights && scratch_dev && n_layers > 0, "svla_decode_mega_plan: null pointer");
  SVLA_REQUIRE(svla_decode_mega_supported(1, hidden, hq, hkv, d, ff, 1), "svla_decode_mega_plan: unsupported geometry");
  const Layout L = make_layout(hidden, hq, hkv, d, ff);
  CUtensorMap* maps = static_cast<CUtensorMap*>(maps_host);
  for (int l = 0; l < n_layers; ++l) {
    for (int g = 0; g < 4; ++g) {
      const void* w = weights[l * 4 + g];
      SVLA_REQUIRE(w && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "svla_decode_mega_plan: weight %d of layer %d is null / misaligned", g, l);
      CUtensorMap tm;
      const int rc = encode_kmajor(&tm, w, static_cast<uint64_t>(L.g[g].k), static_cast<uint64_t>(L.g[g].n), static_cast<uint64_t>(L.g[g].k), kWM);
      SVLA_REQUIRE(rc == 0, "svla_decode_mega_plan: cuTensorMapEncodeTiled(W) failed (%d)", rc);
      memcpy(&maps[l * 4 + g], &tm, sizeof(tm));
    }
  }
  uint8_t* s = static_cast<uint8_t*>(scratch_dev);
  const void* xb[3] = {s + L.off_h, s + L.off_ctx, s + L.off_act};
  const uint64_t xk[3] = {static_cast<uint64_t>(hidden), static_cast<uint64_t>(hq) * d, static_cast<uint64_t>(ff)};
  for (int i = 0; i < 3; ++i) {
    CUtensorMap tm;
    const int rc = encode_kmajor(&tm, xb[i], xk[i], kNB, xk[i], kNB);
    SVLA_REQUIRE(rc == 0, "svla_decode_mega_plan: cuTensorMapEncodeTiled(X) failed (%d)", rc);
    memcpy(&maps[4 * n_layers + i], &tm, sizeof(tm));
  }
  return 0;
}

extern "C" int svla_decode_mega_step(const void* maps_dev, const void* norm_w_dev, int n_layers, float* x, const float* final_norm_w,
                                     void* h_out_bf16, void* kcache, void* vcache, int64_t cache_layer_stride, void* scratch_dev, int batch,
                                     int hidden, int hq, int hkv, int d, int ff, int smax, int ctx, float theta, float scale, float softcap,
                                     float eps, const int32_t* kv_start, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVLA_REQUIRE(maps_dev && norm_w_dev && x && final_norm_w && h_out_bf16 && kcache && vcache && scratch_dev, "svla_decode_mega_step: null pointer");
  SVLA_REQUIRE(svla_decode_mega_supported(batch, hidden, hq, hkv, d, ff, ctx), "svla_decode_mega_step: unsupported geometry (batch %d, ctx %d)", batch, ctx);
  SVLA_REQUIRE(ctx <= smax && n_layers > 0 && (reinterpret_cast<uintptr_t>(maps_dev) & 63) == 0, "svla_decode_mega_step: bad ctx / layers / map alignment");
  const Layout L = make_layout(hidden, hq, hkv, d, ff);
  uint8_t* s = static_cast<uint8_t*>(scratch_dev);
  MegaParams p{};
  p.maps = static_cast<const CUtensorMap*>(maps_dev);
  p.norm_w = static_cast<const float* const*>(norm_w_dev);
  p.n_layers = n_layers;
  p.x = x; p.final_w = final_norm_w; p.h_out = static_cast<__nv_bfloat16*>(h_out_bf16);
  p.kcache = static_cast<__nv_bfloat16*>(kcache); p.vcache = static_cast<__nv_bfloat16*>(vcache);
  p.cache_layer_stride = cache_layer_stride;
  p.h = reinterpret_cast<__nv_bfloat16*>(s + L.off_h);
  p.ctxb = reinterpret_cast<__nv_bfloat16*>(s + L.off_ctx);
  p.act = reinterpret_cast<__nv_bfloat16*>(s + L.off_act);
  p.part = reinterpret_cast<float*>(s + L.off_part);
  p.counter = reinterpret_cast<unsigned*>(s + L.off_counter);
  static const int timing_layer = getenv("SVLA_DECODE_MEGA_TIMING") ? atoi(getenv("SVLA_DECODE_MEGA_TIMING")) : -1;
  p.timing = timing_layer >= 0 ? reinterpret_cast<unsigned long long*>(s + L.off_timing) : nullptr;
  p.timing_layer = timing_layer;
  p.kv_start = kv_start;
  for (int i = 0; i < 4; ++i) p.g[i] = L.g[i];
  p.batch = batch; p.H = hidden; p.hq = hq; p.hkv = hkv; p.FF = ff; p.smax = smax; p.ctx = ctx;
  p.theta = theta; p.scale = scale; p.softcap = softcap; p.eps = eps;
  cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned), st);
  SVLA_REQUIRE(e == cudaSuccess, "svla_decode_mega_step: memset failed: %s", cudaGetErrorString(e));
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(svla_decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    SVLA_REQUIRE(e == cudaSuccess, "svla_decode_mega_step: smem opt-in failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  // Cooperative launch: the driver starts the grid only when ALL its CTAs can be resident at once, so the in-kernel grid
  // barriers cannot deadlock when something else shares the GPU.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(svla_num_sms()));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, svla_decode_mega_kernel, p);
  SVLA_REQUIRE(e == cudaSuccess, "svla_decode_mega_step: cooperative launch failed: %s", cudaGetErrorString(e));
  SVLA_LAUNCH_CHECK("svla_decode_mega_step");
  return 0;
}
