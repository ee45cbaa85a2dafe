// G4: persistent decode step for SMALL batches (B <= 4, the latency configuration of BASELINE.json: "p50 latency at bs=1").
//
// At batch 1 a decode step is 5.2 GB of weights read once; the 186-launch chain of the batched path (7 kernels per layer) spends
// most of its ~1.7 ms in per-launch fixed cost (launch, prologue, first-byte latency, drain), not in HBM time (0.8 ms).  With at
// most 4 activation rows none of the projections needs tensor cores: this kernel runs ALL layers of one decode step in ONE
// launch -- one CTA per SM, every phase a weight-streaming GEMV over all SMs (128-bit non-allocating loads, fp32 accumulation,
// warp-shuffle reductions), phases separated by a grid barrier (monotonic global counter, release/acquire).  The small vectors
// between phases (residual stream, qkv, attention partial states, MLP activations) live in an L2-resident scratch buffer and
// every CTA recomputes the cheap row-wise work (RMSNorm statistics, softmax combine) it needs instead of waiting for another
// kernel to do it.  Phases of a layer (model/modeling_gemma2.py:364-413,451-506):
//   A  h = rms(x)(1+w_in)                      -> qkv = Wqkv h
//   B  RoPE(q, k_new), cache append, soft-capped attention over the cache (split over key ranges, flash-decoding style)
//   C  ctx = combine(partials)                 -> o = Wo ctx
//   D  x += rms(o)(1+w_post); h = rms(x)(1+w_pre)  -> act = gelu_tanh(gate) * up
//   E                                          -> dn = Wd act        (x += rms(dn)(1+w_post_ff) is folded into the next A)
// Numerics follow the batched kernels: operands rounded to bf16 where those round (h, ctx, act, q, k, v), fp32 accumulation,
// fp32 residual stream and norm statistics.
#include <cstdlib>
#include "svla_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSplits = 8;             // key-range splits of one (batch row, kv head) attention item
constexpr int kD = 256;                // head dim (Gemma2)

struct Params {
  const SvlaDecodeLayer* layers;
  int n_layers;
  float* x0;             // [B][H] residual stream in (buffer 0)
  float* x1;             // [B][H] ping-pong residual buffer
  float* qkv;            // [B][(hq+2hkv)*D]
  float* part;           // [B*hq][kSplits][2 + D] attention partial states (m, l, acc)
  float* o;              // [B][H]
  float* act;            // [B][FF]
  float* dn;             // [B][H]
  const float* final_w;
  __nv_bfloat16* h_out;  // [B][H] final-normed hidden state
  unsigned* barrier;
  const int* kv_start;
  int H, hq, hkv, FF, smax, ctx;
  unsigned long long* timing;   // optional [32] timestamps (ns) of CTA 0 in layer `timing_layer` (profiling aid, may be null)
  int vec_bytes;         // bytes of the activation-vector region at the start of dynamic shared memory (16-byte multiple)
  float theta, scale, softcap, eps;
};

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16(v)); }
__device__ __forceinline__ float tanh_cap(float u) {
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

// All CTAs of the grid are co-resident (grid <= #SMs, one CTA per SM by shared-memory size), so a spin barrier is safe; the spin
// is wall-clock bounded so that a scheduling surprise fails the launch instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // arrive: one release reduction (orders this CTA's phase output before the count); poll with RELAXED loads -- an acquire
    // load per spin iteration invalidates the L1 every time (CCTL.IVALL, 346 K of them per step in the first version) -- and
    // take the acquire fence once, after the count has been seen
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned v;
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if ((spin & 4095u) == 4095u) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) { printf("svla_decode_step_small: grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
      }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ float block_sum256(float v, float* sh) {      // sh: 8 floats; all threads get the sum
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += sh[w];
  return t;
}

// s_vec[b][k] (fp32 values already rounded to bf16) x W[n][k] (bf16, row-major, ld = K) -> epilogue(n, acc[NB]).
// Output rows are dealt round-robin to all warps of the grid, two rows per warp and work item.  Weight bytes reach the SM through
// a per-warp ring of kSlots shared-memory slots filled by 1-D bulk copies (cp.async.bulk, completion on an mbarrier): a warp keeps
// kSlots-1 items (2 rows x <= 1152 columns, 4.6 KB each) in flight while it multiplies the current one, so ~110 KB per SM are
// always outstanding -- HBM latency is covered by memory-level parallelism, not by occupancy (one 8-warp CTA per SM).
constexpr int kSeg = 1152;             // columns per ring item (K = 2304 -> 2 segments, 9216 -> 8)
constexpr int kSlots = 4;
constexpr int kSlotBytes = 2 * kSeg * 2;

struct GemvRing {
  uint8_t* slots;        // this warp's kSlots * kSlotBytes bytes
  uint64_t* bars;        // this warp's kSlots mbarriers
  uint32_t parity;       // bit s = phase of slot s the consumer waits for next
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ring_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (spin > (1u << 26)) { printf("svla_decode_step_small: weight ring wait timed out\n"); __trap(); }
  }
}

template <int NB, typename Epi>
__device__ __forceinline__ void gemv_rows(const __nv_bfloat16* __restrict__ W, int N, int K, const float* s_vec, GemvRing& ring, bool prime_only,
                                          Epi epi) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5), nw = gridDim.x * kWarps;
  const int seg = K < kSeg ? K : kSeg;
  const int nseg = (K + seg - 1) / seg;
  const int my_pairs = (gw * 2 < N) ? ((N - gw * 2 + nw * 2 - 1) / (nw * 2)) : 0;
  const int n_items = my_pairs * nseg;
  auto issue = [&](int item) {               // lane 0: request item `item` into slot item % kSlots
    const int slot = item % kSlots, it = item / nseg, sg = item - it * nseg;
    const int n0 = gw * 2 + it * nw * 2;
    const int cols = min(seg, K - sg * seg);
    const bool two = n0 + 1 < N;
    const uint32_t bytes = static_cast<uint32_t>(cols) * 2u;
    const uint32_t bar = smem_addr(ring.bars + slot), dst = smem_addr(ring.slots + slot * kSlotBytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(two ? 2u * bytes : bytes) : "memory");
    const __nv_bfloat16* src = W + static_cast<long long>(n0) * K + sg * seg;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    if (two)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst + kSeg * 2), "l"(src + K), "r"(bytes), "r"(bar) : "memory");
  };
  if (prime_only) {        // weights are immutable: the first items of a phase are requested one phase (and one barrier) ahead
    if (lane == 0)
      for (int i = 0; i < kSlots - 1 && i < n_items; ++i) issue(i);
    return;
  }
  float a0[NB], a1[NB];
  for (int item = 0; item < n_items; ++item) {
    const int slot = item % kSlots, it = item / nseg, sg = item - it * nseg;
    if (sg == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) { a0[b] = 0.f; a1[b] = 0.f; }
    }
    __syncwarp();                                        // every lane is done with the slot that is refilled next
    if (lane == 0 && item + kSlots - 1 < n_items) issue(item + kSlots - 1);
    ring_wait(ring.bars + slot, (ring.parity >> slot) & 1u);
    ring.parity ^= 1u << slot;
    const int cols = min(seg, K - sg * seg);
    const uint4* r0 = reinterpret_cast<const uint4*>(ring.slots + slot * kSlotBytes);
    const uint4* r1 = r0 + kSeg * 2 / 16;
    const float* xv = s_vec + sg * seg;
    for (int k8 = lane; k8 < (cols >> 3); k8 += 32) {
      const uint4 q0 = r0[k8], q1 = r1[k8];
      const uint32_t u0[4] = {q0.x, q0.y, q0.z, q0.w}, u1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 xa = *reinterpret_cast<const float4*>(xv + b * K + k8 * 8);
        const float4 xb = *reinterpret_cast<const float4*>(xv + b * K + k8 * 8 + 4);
        const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a0[b] = fmaf(bf16_bits_to_float(u0[i] & 0xFFFFu), xs[2 * i], a0[b]);
          a0[b] = fmaf(bf16_bits_to_float(u0[i] >> 16), xs[2 * i + 1], a0[b]);
          a1[b] = fmaf(bf16_bits_to_float(u1[i] & 0xFFFFu), xs[2 * i], a1[b]);
          a1[b] = fmaf(bf16_bits_to_float(u1[i] >> 16), xs[2 * i + 1], a1[b]);
        }
      }
    }
    if (sg == nseg - 1) {
      const int n0 = gw * 2 + it * nw * 2;
      float t0[NB], t1[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) { t0[b] = warp_sum(a0[b]); t1[b] = warp_sum(a1[b]); }
      if (lane == 0) epi(n0, t0, n0 + 1 < N, t1);
    }
  }
  __syncwarp();
}

// x_out = x_in + rms(branch)(1+w_post) (branch may be null: x_out = x_in); s_vec = bf16round(rms(x_out)(1+w_pre)).
// Every CTA computes all NB rows; CTA 0 additionally stores x_out.
constexpr int kNormVec = 16;            // elements per thread: hidden <= 4096
template <int NB, typename Prime>
__device__ __forceinline__ void norm_prologue(const Params& p, const float* x_in, const float* branch, const float* w_post,
                                              const float* w_pre, float* x_out, float* s_vec, float* s_red, Prime prime_next) {
  // one pass: every operand is loaded once, all loads are issued before the first use (one L2 round trip), the two row
  // statistics are two block reductions over registers
  const int H = p.H;
  // A warp issues in order: a use placed between two loads would stall the second load behind the first one's latency (ncu on
  // the first version: 36 serialised L2 round trips, ~6 us).  So: every load first (clamped index instead of predicates, nothing
  // but loads in this loop), the arithmetic afterwards.
  float xv[NB][kNormVec], bv[NB][kNormVec], wq[kNormVec], wp[kNormVec];
  const float* br = branch ? branch : x_in;               // dummy source when there is no branch (values unused)
  const float* wpp = branch ? w_post : w_pre;
#pragma unroll
  for (int j = 0; j < kNormVec; ++j) {
    const int i = min(threadIdx.x + j * kThreads, H - 1);
    wq[j] = __ldg(w_pre + i);
    wp[j] = __ldg(wpp + i);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      xv[b][j] = __ldcg(x_in + b * H + i);
      bv[b][j] = __ldcg(br + b * H + i);
    }
  }
#pragma unroll
  for (int j = 0; j < kNormVec; ++j) {
    const bool ok = threadIdx.x + j * kThreads < H;
    wq[j] = ok ? 1.f + wq[j] : 0.f;
    wp[j] = (ok && branch) ? 1.f + wp[j] : 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (!ok) xv[b][j] = 0.f;
      if (!ok || !branch) bv[b][j] = 0.f;
    }
  }
  // The weight ring of the following GEMV is primed only now: its bulk copies (110 KB per SM) queue up behind the few small
  // loads above instead of in front of them (primed before the barrier, the prologue's loads waited ~5 us behind 16 MB of weights)
  prime_next();
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (branch) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < kNormVec; ++j) ss = fmaf(bv[b][j], bv[b][j], ss);
      const float inv1 = rsqrtf(block_sum256(ss, s_red) / static_cast<float>(H) + p.eps);
#pragma unroll
      for (int j = 0; j < kNormVec; ++j) xv[b][j] += bv[b][j] * inv1 * wp[j];
    }
    float ss2 = 0.f;
#pragma unroll
    for (int j = 0; j < kNormVec; ++j) ss2 = fmaf(xv[b][j], xv[b][j], ss2);
    const float inv2 = rsqrtf(block_sum256(ss2, s_red) / static_cast<float>(H) + p.eps);
#pragma unroll
    for (int j = 0; j < kNormVec; ++j) {
      const int i = threadIdx.x + j * kThreads;
      if (i < H) {
        if (blockIdx.x == 0 && x_out) x_out[b * H + i] = xv[b][j];
        s_vec[b * H + i] = bf16r(xv[b][j] * inv2 * wq[j]);
      }
    }
  }
  __syncthreads();
}

template <int NB>
__global__ void __launch_bounds__(kThreads, 1)
svla_decode_step_small_kernel(const Params p) {
  extern __shared__ __align__(16) float s_dyn[];
  float* s_vec = s_dyn;                                     // [NB][max(H, FF, hq*D)] GEMV input vectors
  __shared__ float s_red[kWarps];
  __shared__ float s_q[2 * kD], s_knew[kD], s_vnew[kD];     // attention phase (GQA group <= 2)
  __shared__ float s_m[kWarps][2], s_l[kWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = p.H, FF = p.FF, hq = p.hq, hkv = p.hkv, grp = hq / hkv;
  // weight ring of this warp (behind the activation vectors) and its mbarriers
  GemvRing ring;
  {
    uint8_t* ring_base = reinterpret_cast<uint8_t*>(s_dyn) + p.vec_bytes;
    ring.slots = ring_base + warp * kSlots * kSlotBytes;
    ring.bars = reinterpret_cast<uint64_t*>(ring_base + kWarps * kSlots * kSlotBytes) + warp * kSlots;
    ring.parity = 0u;
    if (lane == 0)
      for (int s2 = 0; s2 < kSlots; ++s2)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(ring.bars + s2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }
  const int QW = (hq + 2 * hkv) * kD;
  auto noop = [](int, const float*, bool, const float*) {};
  auto prime = [&](const void* W, int N, int K) { gemv_rows<NB>(static_cast<const __nv_bfloat16*>(W), N, K, s_vec, ring, true, noop); };
  unsigned epoch = 0;
  float* xa = p.x0;
  float* xb = p.x1;
  const float* prev_dn = nullptr;
  const float* prev_post_ff = nullptr;

  int tix = 0;
  auto stamp = [&](int li_) {
    if (p.timing && blockIdx.x == 0 && threadIdx.x == 0 && li_ == 5 && tix < 32) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      p.timing[tix++] = now;
    }
  };
  for (int li = 0; li < p.n_layers; ++li) {
    const SvlaDecodeLayer L = p.layers[li];
    stamp(li);
    // ---------------------------------------------------------------- phase A: (residual of the previous MLP) + input norm + qkv
    norm_prologue<NB>(p, xa, prev_dn, prev_post_ff, L.ln_in, prev_dn ? xb : nullptr, s_vec, s_red, [&]() { prime(L.wqkv, QW, H); });
    if (prev_dn) { float* t = xa; xa = xb; xb = t; }
    stamp(li);
    {
      float* qkv = p.qkv;
      gemv_rows<NB>(static_cast<const __nv_bfloat16*>(L.wqkv), QW, H, s_vec, ring, false, [&](int n, const float* a0, bool two, const float* a1) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          qkv[b * QW + n] = a0[b];
          if (two) qkv[b * QW + n + 1] = a1[b];
        }
      });
      prime(L.wo, H, hq * kD);
    }
    stamp(li);
    grid_barrier(p.barrier, (++epoch) * gridDim.x);
    stamp(li);
    // ---------------------------------------------------------------- phase B: RoPE + cache append + attention partials
    {
      const int items = NB * hkv * kSplits;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int split = item % kSplits, bk = item / kSplits, hk = bk % hkv, b = bk / hkv;
        const int kstart = p.kv_start ? p.kv_start[b] : 0;
        const int n_old = p.ctx - 1;
        const float* src = p.qkv + b * QW;
        __syncthreads();
        if (threadIdx.x < kD / 2) {
          const int j = threadIdx.x;
          const float inv_freq = 1.0f / powf(p.theta, static_cast<float>(2 * j) / static_cast<float>(kD));
          float sn, cs;
          sincosf(static_cast<float>(p.ctx - kstart) * inv_freq, &sn, &cs);
          float xr1[3], xr2[3];
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const int gg = min(g, grp);
            const int col = (gg < grp) ? (hk * grp + gg) * kD : (hq + hk) * kD;
            xr1[g] = __ldcg(src + col + j); xr2[g] = __ldcg(src + col + j + kD / 2);
          }
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            if (g > grp) continue;
            const float x1 = xr1[g], x2 = xr2[g];
            const float o1 = bf16r(x1 * cs - x2 * sn), o2 = bf16r(x2 * cs + x1 * sn);
            if (g < grp) { s_q[g * kD + j] = o1; s_q[g * kD + j + kD / 2] = o2; }
            else { s_knew[j] = o1; s_knew[j + kD / 2] = o2; }
          }
        } else {
          const int dpos = (threadIdx.x - kD / 2) * 2;
          s_vnew[dpos] = bf16r(__ldcg(src + (hq + hkv + hk) * kD + dpos));
          s_vnew[dpos + 1] = bf16r(__ldcg(src + (hq + hkv + hk) * kD + dpos + 1));
        }
        __syncthreads();
        __nv_bfloat16* kc = static_cast<__nv_bfloat16*>(L.kcache) + (static_cast<long long>(b) * p.smax * hkv + hk) * kD;
        __nv_bfloat16* vc = static_cast<__nv_bfloat16*>(L.vcache) + (static_cast<long long>(b) * p.smax * hkv + hk) * kD;
        const long long row_stride = static_cast<long long>(hkv) * kD;
        if (split == 0) {                                   // append the new token's key / value at slot ctx - 1
          kc[n_old * row_stride + threadIdx.x] = __float2bfloat16(s_knew[threadIdx.x]);
          vc[n_old * row_stride + threadIdx.x] = __float2bfloat16(s_vnew[threadIdx.x]);
        }
        // keys of this split: an even share of [kstart, n_old); the last split also owns the new key (from shared memory)
        const int n_valid = max(n_old - kstart, 0);
        const int per = (n_valid + kSplits - 1) / kSplits;
        const int lo = kstart + split * per, hi = min(lo + per, n_old);
        float qv[2][8];
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int e = 0; e < 8; ++e) qv[g][e] = (g < grp) ? s_q[g * kD + lane * 8 + e] : 0.f;
        float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f}, acc[2][8];
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[g][e] = 0.f;
        const float inv_cap = p.softcap > 0.f ? 1.f / p.softcap : 0.f;
        auto visit = [&](const float (&kv)[8], const float (&vv)[8]) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (g < grp) {
              float dot = 0.f;
#pragma unroll
              for (int e = 0; e < 8; ++e) dot = fmaf(kv[e], qv[g][e], dot);
              dot = warp_sum(dot) * p.scale;
              if (p.softcap > 0.f) dot = p.softcap * tanh_cap(dot * inv_cap);
              const float m_new = fmaxf(m_run[g], dot);
              const float corr = __expf(m_run[g] - m_new), pj = __expf(dot - m_new);
              l_run[g] = l_run[g] * corr + pj;
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[g][e] = acc[g][e] * corr + pj * vv[e];
              m_run[g] = m_new;
            }
          }
        };
        // this warp's keys, kKeyBatch at a time: all K / V row loads of a batch are issued before the first dot product
        constexpr int kKeyBatch = 6;
        for (int j0 = lo + warp; j0 < hi; j0 += kWarps * kKeyBatch) {
          uint4 kr[kKeyBatch], vr[kKeyBatch];
#pragma unroll
          for (int u = 0; u < kKeyBatch; ++u) {
            const int j = min(j0 + u * kWarps, hi - 1);
            kr[u] = __ldcg(reinterpret_cast<const uint4*>(kc + j * row_stride + lane * 8));
            vr[u] = __ldcg(reinterpret_cast<const uint4*>(vc + j * row_stride + lane * 8));
          }
#pragma unroll
          for (int u = 0; u < kKeyBatch; ++u) {
            if (j0 + u * kWarps < hi) {                     // warp-uniform
              float kv[8], vv[8];
              const uint32_t ku[4] = {kr[u].x, kr[u].y, kr[u].z, kr[u].w}, vu[4] = {vr[u].x, vr[u].y, vr[u].z, vr[u].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                kv[2 * i] = bf16_bits_to_float(ku[i] & 0xFFFFu); kv[2 * i + 1] = bf16_bits_to_float(ku[i] >> 16);
                vv[2 * i] = bf16_bits_to_float(vu[i] & 0xFFFFu); vv[2 * i + 1] = bf16_bits_to_float(vu[i] >> 16);
              }
              visit(kv, vv);
            }
          }
        }
        if (split == kSplits - 1 && warp == 0) {
          float kv[8], vv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { kv[e] = s_knew[lane * 8 + e]; vv[e] = s_vnew[lane * 8 + e]; }
          visit(kv, vv);
        }
        // merge the warps' states through shared memory (s_vec is free in this phase), then one partial per (row, head, split)
        float* s_acc = s_vec;                                // [kWarps][2][kD]
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (lane == 0) { s_m[warp][g] = m_run[g]; s_l[warp][g] = l_run[g]; }
#pragma unroll
          for (int e = 0; e < 8; ++e) s_acc[(warp * 2 + g) * kD + lane * 8 + e] = acc[g][e];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < grp * kD; i += kThreads) {
          const int g = i / kD, dd = i - g * kD;
          float mmax = -INFINITY;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) mmax = fmaxf(mmax, s_m[w][g]);
          float num = 0.f, den = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) {
            const float f = (s_m[w][g] == -INFINITY) ? 0.f : __expf(s_m[w][g] - mmax);
            num += f * s_acc[(w * 2 + g) * kD + dd];
            den += f * s_l[w][g];
          }
          float* dst = p.part + (static_cast<long long>(b * hq + hk * grp + g) * kSplits + split) * (2 + kD);
          dst[2 + dd] = num;
          if (dd == 0) { dst[0] = mmax; dst[1] = den; }
        }
      }
    }
    stamp(li);
    grid_barrier(p.barrier, (++epoch) * gridDim.x);
    stamp(li);
    // ---------------------------------------------------------------- phase C: combine the attention partials, o projection
    {
      const int CW = hq * kD;
      // per (row, head): softmax weights of the kSplits partial states, computed once, then one weighted sum per element
      float* s_f = s_vec + NB * CW;                        // [NB*hq][kSplits]
      for (int i = threadIdx.x; i < NB * hq; i += kThreads) {
        const float* base = p.part + static_cast<long long>(i) * kSplits * (2 + kD);
        float ms[kSplits], ls[kSplits], mmax = -INFINITY, den = 0.f;
#pragma unroll
        for (int sp = 0; sp < kSplits; ++sp) { ms[sp] = __ldcg(base + sp * (2 + kD)); ls[sp] = __ldcg(base + sp * (2 + kD) + 1); }
#pragma unroll
        for (int sp = 0; sp < kSplits; ++sp) mmax = fmaxf(mmax, ms[sp]);
#pragma unroll
        for (int sp = 0; sp < kSplits; ++sp) {
          ms[sp] = (ms[sp] == -INFINITY) ? 0.f : __expf(ms[sp] - mmax);
          den += ms[sp] * ls[sp];
        }
#pragma unroll
        for (int sp = 0; sp < kSplits; ++sp) s_f[i * kSplits + sp] = ms[sp] / den;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < NB * CW; i += kThreads) {
        const int bh = i / kD, dd = i - bh * kD;             // bh = b * hq + head, and i = b * CW + head * kD + dd
        const float* base = p.part + static_cast<long long>(bh) * kSplits * (2 + kD) + 2 + dd;
        float pv[kSplits], num = 0.f;
#pragma unroll
        for (int sp = 0; sp < kSplits; ++sp) pv[sp] = __ldcg(base + sp * (2 + kD));
#pragma unroll
        for (int sp = 0; sp < kSplits; ++sp) num = fmaf(s_f[bh * kSplits + sp], pv[sp], num);
        s_vec[i] = bf16r(num);
      }
      __syncthreads();
      stamp(li);
      float* o = p.o;
      gemv_rows<NB>(static_cast<const __nv_bfloat16*>(L.wo), H, CW, s_vec, ring, false, [&](int n, const float* a0, bool two, const float* a1) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          o[b * H + n] = a0[b];
          if (two) o[b * H + n + 1] = a1[b];
        }
      });
    }
    stamp(li);
    grid_barrier(p.barrier, (++epoch) * gridDim.x);
    stamp(li);
    // ---------------------------------------------------------------- phase D: post-attention norm + residual, pre-FF norm, gate/up
    norm_prologue<NB>(p, xa, p.o, L.ln_post_attn, L.ln_pre_ff, xb, s_vec, s_red, [&]() { prime(L.wgu, 2 * FF, H); });
    { float* t = xa; xa = xb; xb = t; }
    stamp(li);
    {
      float* act = p.act;
      // rows 2j = gate_j, 2j+1 = up_j (the engine's interleaved layout): one warp iteration yields one GeGLU output
      gemv_rows<NB>(static_cast<const __nv_bfloat16*>(L.wgu), 2 * FF, H, s_vec, ring, false, [&](int n, const float* a0, bool two, const float* a1) {
#pragma unroll
        for (int b = 0; b < NB; ++b) act[b * FF + (n >> 1)] = bf16r(gelu_tanh_fast(a0[b]) * a1[b]);
      });
      prime(L.wd, H, FF);
    }
    stamp(li);
    grid_barrier(p.barrier, (++epoch) * gridDim.x);
    stamp(li);
    // ---------------------------------------------------------------- phase E: down projection
    {
      for (int i0 = 0; i0 < NB * FF; i0 += 12 * kThreads) {          // 12 loads in flight per thread, then the 12 stores
        float t[12];
#pragma unroll
        for (int u = 0; u < 12; ++u) t[u] = __ldcg(p.act + min(i0 + u * kThreads + static_cast<int>(threadIdx.x), NB * FF - 1));
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const int i = i0 + u * kThreads + threadIdx.x;
          if (i < NB * FF) s_vec[i] = t[u];
        }
      }
      __syncthreads();
      stamp(li);
      float* dn = p.dn;
      gemv_rows<NB>(static_cast<const __nv_bfloat16*>(L.wd), H, FF, s_vec, ring, false, [&](int n, const float* a0, bool two, const float* a1) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          dn[b * H + n] = a0[b];
          if (two) dn[b * H + n + 1] = a1[b];
        }
      });
    }
    stamp(li);
    grid_barrier(p.barrier, (++epoch) * gridDim.x);
    stamp(li);
    prev_dn = p.dn;
    prev_post_ff = L.ln_post_ff;
  }
  // ------------------------------------------------------------------ final: last MLP residual + final norm -> h_out (CTA 0)
  if (blockIdx.x == 0) {
    norm_prologue<NB>(p, xa, prev_dn, prev_post_ff, p.final_w, nullptr, s_vec, s_red, []() {});
    for (int i = threadIdx.x; i < NB * H; i += kThreads) p.h_out[i] = __float2bfloat16(s_vec[i]);
  }
}

}  // namespace

extern "C" int64_t svla_decode_step_small_scratch_floats(int batch, int hidden, int hq, int hkv, int d, int ff) {
  const int64_t B = batch;
  return B * hidden /*x1*/ + B * (hq + 2 * hkv) * d /*qkv*/ + B * hq * kSplits * (2 + d) /*partials*/ + B * hidden /*o*/ + B * ff /*act*/ +
         B * hidden /*dn*/ + 16 /*barrier counter, padded*/ + 64 /*optional phase timestamps*/;
}

extern "C" int svla_decode_step_small(const SvlaDecodeLayer* layers_dev, int n_layers, float* x, const float* final_norm_w,
                                      void* h_out_bf16, float* scratch, int batch, int hidden, int hq, int hkv, int d, int ff,
                                      int smax, int ctx, float theta, float scale, float softcap, float eps, const int32_t* kv_start,
                                      void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVLA_REQUIRE(layers_dev && x && final_norm_w && h_out_bf16 && scratch, "svla_decode_step_small: null pointer");
  SVLA_REQUIRE(batch == 1 || batch == 2 || batch == 4, "svla_decode_step_small: batch %d not in {1, 2, 4}", batch);
  SVLA_REQUIRE(d == kD && hkv > 0 && hq % hkv == 0 && hq / hkv <= 2, "svla_decode_step_small: head dim 256 and a GQA group <= 2 only");
  SVLA_REQUIRE((hidden % 8) == 0 && hidden <= kNormVec * kThreads && (ff % 8) == 0 && n_layers > 0 && ctx >= 1 && ctx <= smax,
               "svla_decode_step_small: bad geometry (hidden <= %d)", kNormVec * kThreads);
  const int nb = batch <= 1 ? 1 : (batch <= 2 ? 2 : 4);
  Params p{};
  p.layers = layers_dev; p.n_layers = n_layers;
  p.x0 = x;
  float* s = scratch;
  p.x1 = s; s += static_cast<int64_t>(batch) * hidden;
  p.qkv = s; s += static_cast<int64_t>(batch) * (hq + 2 * hkv) * d;
  p.part = s; s += static_cast<int64_t>(batch) * hq * kSplits * (2 + d);
  p.o = s; s += static_cast<int64_t>(batch) * hidden;
  p.act = s; s += static_cast<int64_t>(batch) * ff;
  p.dn = s; s += static_cast<int64_t>(batch) * hidden;
  p.barrier = reinterpret_cast<unsigned*>(s);
  static const bool want_timing = getenv("SVLA_DECODE_SMALL_TIMING") != nullptr;       // profiling aid: 32 x u64 behind the counter
  p.timing = want_timing ? reinterpret_cast<unsigned long long*>(s + 16) : nullptr;
  p.final_w = final_norm_w; p.h_out = static_cast<__nv_bfloat16*>(h_out_bf16); p.kv_start = kv_start;
  p.H = hidden; p.hq = hq; p.hkv = hkv; p.FF = ff; p.smax = smax; p.ctx = ctx;
  p.theta = theta; p.scale = scale; p.softcap = softcap; p.eps = eps;
  int widest = hidden > ff ? hidden : ff;
  if (hq * d > widest) widest = hq * d;
  if (kWarps * 2 * kD > widest * nb) widest = (kWarps * 2 * kD + nb - 1) / nb;       // warp-state merge buffer of the attention phase
  widest += hq * kSplits;                                                              // softmax weights of the combine (phase C)
  // activation vectors + per-warp weight rings + mbarriers; >= 114 KB keeps the kernel at one CTA per SM, so that the
  // grid <= #SMs CTAs are all co-resident (spin barrier)
  p.vec_bytes = static_cast<int>((static_cast<size_t>(nb) * widest * sizeof(float) + 15) / 16 * 16);
  size_t smem = static_cast<size_t>(p.vec_bytes) + static_cast<size_t>(kWarps) * kSlots * kSlotBytes + kWarps * kSlots * sizeof(uint64_t);
  if (smem < 116 * 1024) smem = 116 * 1024;
  SVLA_REQUIRE(smem <= 227 * 1024, "svla_decode_step_small: %zu bytes of shared memory needed (batch %d, width %d)", smem, batch, widest);
  cudaError_t e = cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), st);
  SVLA_REQUIRE(e == cudaSuccess, "svla_decode_step_small: memset failed: %s", cudaGetErrorString(e));
  const int grid = svla_num_sms();
  auto launch = [&](auto kernel) -> cudaError_t {
    static_assert(sizeof(p) < 4000, "kernel parameter block");
    cudaError_t ce = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce != cudaSuccess) return ce;
    // Cooperative launch: the driver starts the grid only when ALL its CTAs can be resident at once (or fails with
    // cudaErrorCooperativeLaunchTooLarge), so the spin barriers cannot deadlock when another stream / process shares the GPU.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
  };
  if (nb == 1) e = launch(svla_decode_step_small_kernel<1>);
  else if (nb == 2) e = launch(svla_decode_step_small_kernel<2>);
  else e = launch(svla_decode_step_small_kernel<4>);
  SVLA_REQUIRE(e == cudaSuccess, "svla_decode_step_small: cooperative launch (smem %zu) failed: %s", smem, cudaGetErrorString(e));
  SVLA_LAUNCH_CHECK("svla_decode_step_small");
  return 0;
}
