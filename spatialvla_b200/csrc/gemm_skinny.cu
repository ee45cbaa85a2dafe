// G1s: skinny (decode) GEMM  out[m, n] = epilogue(sum_k X[m,k] W[n,k])  for M <= 128 activation rows.
//
// At batch 64 the decode GEMMs are weight-streaming, HBM-bound (AI ~ 64 FLOP/B): the job of the kernel is to pull
// the weight matrix through all 148 SMs once, as fast as HBM delivers it.  So the operands are SWAPPED relative to
// the big-GEMM kernel: the 128-row UMMA M dimension walks the WEIGHT rows (16 KB of weights per K step per CTA, all
// useful bytes), the activations are the small N-side operand (NB = 16/64/128 columns, L2-hot), and K is SPLIT across
// CTAs so that n_tiles x splits ~ 148 even for the 2304-row projections.  Accumulators are D^T[n, m] in TMEM: TMEM
// lane = weight row n, so for a fixed m the 32 lanes of a warp store 32 consecutive n -- fully coalesced without any
// shared-memory transpose.  Split-K partial sums go to out_f32[split][m][n]; the consumers (svla_rmsnorm_residual,
// svla_rope_kv) add the partials while they read them, deterministically.
// hi/lo activation pairs (flags X_HILO / OUT_HILO): the decode chain hands its activations over as two bf16 planes, hi = bf16(v)
// and lo = bf16(v - hi).  The planes are staged as the two halves of the activation tile, the MMA computes both products against
// the SAME weight tile (streamed once) and the epilogue adds the accumulator halves: hi.W^T + lo.W^T in fp32.  Rounding the decode
// row's own activations to 8 mantissa bits in front of every Linear was the bf16-vs-fp32 logit noise of the path (full-size parity
// 98.9 % -> 99.8 % of 832 positions, profiles/decode_hilo_r2.txt); the wider tile costs 4-7 % of a decode step.
// Reference ops replaced: the q/k/v/o/gate/up/down/lm_head nn.Linear calls of a decode step
// (model/modeling_gemma2.py:80-92,351-354,993).
#include <cstdlib>
#include <cudaTypedefs.h>
#include "tc_ptx.cuh"

namespace {
using namespace svla_ptx;

constexpr int kWM = 128;       // weight rows per CTA (UMMA M)
constexpr int kBK = 64;
constexpr int kThreads = 192;
constexpr int kDefaultCG = 1;  // CTAs per MMA unless SVLA_SKINNY_CG says otherwise

// CG = CTAs per MMA: 2 = a CTA pair (cluster of two consecutive weight-row tiles of the same K split) issues ONE
// tcgen05.mma.cta_group::2 of 256 weight rows per K step; each CTA stages its own 128 weight rows but only HALF of the activation
// tile (with hi/lo planes: rank 0 the hi plane, rank 1 the lo plane), so the L2 -> shared-memory traffic of the activations,
// which equals the weight traffic once the tile is 128 rows wide, is halved.
template <int NB, int CG = 1> struct SCfg {
  static constexpr int kWBytes = kWM * kBK * 2;
  static constexpr int kXBytes = (NB / CG) * kBK * 2;
  static constexpr int kStageBytes = kWBytes + kXBytes;
  static constexpr int kMaxStages = 8;
  static constexpr int kTmemCols = NB < 32 ? 32 : NB;
  static constexpr int smem_bytes(int stages) { return stages * kStageBytes + 1024 + 256; }
};
// Pipeline depth: 8 x 24 KB (NB <= 64) or 6 x 32 KB (NB = 128) in flight per SM.  SVLA_SKINNY_STAGES overrides it for A/B
// measurements; measured on the in-situ decode step (tools/decode_step_perf.py, B=64): 8 stages 2176 us, 4 stages (two CTAs
// of consecutive PDL kernels co-resident per SM) 2238 us, 3 stages 2490 us -- depth beats co-residency.  Also tried and
// dropped: an L2 prefetch chain (every GEMM issuing cp.async.bulk.prefetch.L2 for the next kernel's weights / this layer's KV
// slabs): 2511-2628 us against 2246 us on the same box -- the prefetches compete with the demand stream instead of filling gaps.
inline int skinny_stages(int stage_bytes) {
  static const int env = getenv("SVLA_SKINNY_STAGES") ? atoi(getenv("SVLA_SKINNY_STAGES")) : 0;
  if (env >= 2 && env <= 8) return env;
  return (stage_bytes > 24576) ? 6 : 8;
}

struct SkinnyParams {
  const float* bias;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  long long m, n, ldo, partial_stride;
  float alpha, act_param;
  int act, flags;
  int n_tiles, kb_per_split, num_k_blocks;
  int n_tiles_grid; // n_tiles rounded up to the cluster size (a surplus CTA computes an all-zero tile and stores nothing)
  int stages;      // pipeline depth (<= SCfg::kMaxStages)
  int x_lo_row;    // > 0 (X_HILO): X holds two bf16 planes, rows [0, m) = hi, rows [m, 2m) = lo = bf16(x - hi); each plane is
                   // loaded into one half of the NB-row activation tile and the epilogue adds the two accumulator halves
  int w_tiled;     // W is stored tile-major [n_tile][k_block][128 rows][64 cols]: every 16 KB stage is one contiguous read
};

template <int NB, int CG>
__global__ void __launch_bounds__(kThreads, 2)
svla_gemm_skinny_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const SkinnyParams p) {
  using C = SCfg<NB, CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  const int kStages = p.stages;
  uint8_t* smem_x = smem + kStages * C::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kMaxStages;
  uint64_t* acc_bar = bars + 2 * C::kMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = n_tiles_grid x splits, n_tiles_grid = n_tiles rounded up to a multiple of CG: the CTAs of a pair share the K split
  const int n_tile = blockIdx.x % p.n_tiles_grid, split = blockIdx.x / p.n_tiles_grid;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
#pragma unroll 1
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::kTmemCols, CG>(tmem_ptr_smem);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();      // the pair's barriers exist before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      // every CTA walks its K range from a different starting block (rotation by n_tile) so that the CTAs do not all ask
      // for the same activation tile at the same time
      const int nkb = kb1 - kb0;
      const int rot = nkb > 0 ? ((n_tile / CG) * 5) % nkb : 0;          // the same order in both CTAs of a pair
      auto load_w = [&](int it, int stage) {
        const int kb = kb0 + (it + rot) % nkb;
        if constexpr (CG == 1) {
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          if (p.w_tiled) tma_load_2d(smem_w + stage * C::kWBytes, &tm_w, &full_bar[stage], 0, (n_tile * p.num_k_blocks + kb) * kWM);
          else tma_load_2d(smem_w + stage * C::kWBytes, &tm_w, &full_bar[stage], kb * kBK, n_tile * kWM);
        } else {
          // the leader arms its barrier for the bytes of BOTH CTAs; the peer's TMA signals the leader's barrier
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
          if (p.w_tiled) tma_load_2d_cg2(smem_w + stage * C::kWBytes, &tm_w, &full_bar[stage], 0, (n_tile * p.num_k_blocks + kb) * kWM);
          else tma_load_2d_cg2(smem_w + stage * C::kWBytes, &tm_w, &full_bar[stage], kb * kBK, n_tile * kWM);
        }
      };
      auto load_x = [&](int it, int stage) {
        const int kb = kb0 + (it + rot) % nkb;
        if constexpr (CG == 1) {
          tma_load_2d(smem_x + stage * C::kXBytes, &tm_x, &full_bar[stage], kb * kBK, 0);
          if (p.x_lo_row > 0)        // hi/lo planes: two boxes of NB/2 rows (rows past 2m are zero-filled by the TMA unit)
            tma_load_2d(smem_x + stage * C::kXBytes + C::kXBytes / 2, &tm_x, &full_bar[stage], kb * kBK, p.x_lo_row);
        } else {
          // this CTA's half of the activation tile: rows [rank * NB/2, ...) -- with hi/lo planes rank 0 takes hi, rank 1 lo
          const int row0 = p.x_lo_row > 0 ? static_cast<int>(cta_rank) * p.x_lo_row : static_cast<int>(cta_rank) * (NB / 2);
          tma_load_2d_cg2(smem_x + stage * C::kXBytes, &tm_x, &full_bar[stage], kb * kBK, row0);
        }
      };
      // PDL: the weights are immutable, so the first kStages weight tiles are requested BEFORE waiting for the kernel that
      // produces the activations; their HBM latency (and this kernel's launch + prologue) hides behind that kernel
      const int pre = nkb < kStages ? nkb : kStages;
      for (int it = 0; it < pre; ++it) load_w(it, it);
      pdl_wait();
      for (int it = 0; it < pre; ++it) load_x(it, it);
      int stage = 0;
      uint32_t phase = 1;                          // the ring has been filled once
      for (int it = pre; it < nkb; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        load_w(it, stage);
        load_x(it, stage);
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (leader) {                          // converged warp, an elected lane issues (tc_ptx.cuh: elect_one); CG == 2: the leader only
      constexpr uint32_t idesc = make_idesc_bf16(kWM * CG, NB < 16 ? 16 : NB);
      const uint64_t dw0 = make_kmajor_sw128_desc(smem_u32(smem_w)), dx0 = make_kmajor_sw128_desc(smem_u32(smem_x));
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dw = dw0 + static_cast<uint64_t>((stage * C::kWBytes) >> 4);
          const uint64_t dx = dx0 + static_cast<uint64_t>((stage * C::kXBytes) >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            if constexpr (CG == 1)
              umma_bf16(tmem_base, dw + static_cast<uint64_t>(k * 2), dx + static_cast<uint64_t>(k * 2), idesc,
                        static_cast<uint32_t>((kb > kb0) || k != 0));
            else
              umma_bf16_cg2(tmem_base, dw + static_cast<uint64_t>(k * 2), dx + static_cast<uint64_t>(k * 2), idesc,
                            static_cast<uint32_t>((kb > kb0) || k != 0));
          }
          if constexpr (CG == 1) umma_commit(&empty_bar[stage]); else umma_commit_cg2(&empty_bar[stage]);   // frees the stage in both CTAs
          if (kb == kb1 - 1) {
            if constexpr (CG == 1) umma_commit(acc_bar); else umma_commit_cg2(acc_bar);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      if (kb1 <= kb0) {                    // empty K range (cannot happen with the host's split choice): release the epilogue
        if (elect_one()) { if constexpr (CG == 1) umma_commit(acc_bar); else umma_commit_cg2(acc_bar); }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ epilogue: lane <-> weight row n, register j <-> activation row m
    const int q = warp & 3;
    const long long n = static_cast<long long>(n_tile) * kWM + q * 32 + lane;
    const bool n_ok = n < p.n;
    const bool partial = (p.flags & 2) != 0, geglu = (p.flags & 1) != 0;
    const bool hilo_in = p.x_lo_row > 0, hilo_out = (p.flags & 16) != 0;
    const int ncols = hilo_in ? NB / 2 : (NB < 32 ? 32 : NB);      // accumulator columns that carry activation rows
    const long long lo_plane = p.m * p.ldo;                         // OUT_HILO: the lo plane follows the hi plane
    const float bias = (p.bias != nullptr && n_ok && !partial) ? __ldg(p.bias + n) : 0.f;
    pdl_wait();
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* pout = partial ? p.out_f32 + static_cast<long long>(split) * p.partial_stride : p.out_f32;
    // No data-dependent exit inside the unrolled row loop: with a `break` per row every iteration is its own basic block and the
    // address / activation / store chains of the 32 rows run back to back (measured in the persistent decode kernel: 3 us to
    // store one 64-row partial tile, 8 us for the GeGLU tile); predicated stores let the 32 rows overlap.
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      if (c0 >= p.m) break;                         // warp-uniform
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      if (hilo_in) {                                // x.W = hi.W + lo.W: the lo plane's accumulators sit NB/2 columns further
        if constexpr (NB == 16) {                   // m <= 8: both planes are inside the one 32-column load
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r[j + 8]));
        } else {
          uint32_t r2[32];
          tmem_ld32(taddr + NB / 2 + c0, r2);
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
      }
      const int rows = static_cast<int>(p.m) - c0;  // rows of this chunk that exist (>= 1)
      if (partial) {
        float* dst = pout + static_cast<long long>(c0) * p.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n_ok && j < rows) dst[static_cast<long long>(j) * p.ldo] = __uint_as_float(r[j]);
      } else if (geglu) {
        __nv_bfloat16* dst = p.out_bf16 + static_cast<long long>(c0) * p.ldo + (n >> 1);
        const bool wr = n_ok && (lane & 1) == 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = __uint_as_float(r[j]) * p.alpha + bias;
          const float other = __shfl_xor_sync(0xffffffffu, v, 1);
          // tanh.approx (2^-11 relative) is below the bf16 rounding of a single plane but not of a hi/lo pair: there
          // 1 + tanh(u) = 2 - 2 / (exp(2u) + 1) through ex2 + rcp (absolute error ~1e-7, saturates correctly at +-inf)
          float a;
          if (hilo_out) {
            const float e = __expf(1.5957691216057308f * (v + 0.044715f * v * v * v));
            a = v * (1.f - __fdividef(1.f, e + 1.f)) * other;
          } else {
            a = gelu_tanh_fast(v) * other;
          }
          const __nv_bfloat16 o = __float2bfloat16(a);
          if (wr && j < rows) {
            dst[static_cast<long long>(j) * p.ldo] = o;
            if (hilo_out) dst[lo_plane + static_cast<long long>(j) * p.ldo] = __float2bfloat16(a - __bfloat162float(o));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = __uint_as_float(r[j]) * p.alpha + bias;
          if (p.act == SVLA_ACT_SOFTCAP) v = p.act_param * tanhf(v / p.act_param);
          else if (p.act == SVLA_ACT_GELU_TANH) v = gelu_tanh_fast(v);
          else if (p.act == SVLA_ACT_RELU) v = fmaxf(v, 0.f);
          if (n_ok && j < rows) {
            const long long o = static_cast<long long>(c0 + j) * p.ldo + n;
            if (p.out_f32) p.out_f32[o] = v;
            if (p.out_bf16) {
              const __nv_bfloat16 hi = __float2bfloat16(v);
              p.out_bf16[o] = hi;
              if (hilo_out) p.out_bf16[lo_plane + o] = __float2bfloat16(v - __bfloat162float(hi));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols, CG>(tmem_base);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
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
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

template <int NB, int CG>
int launch_skinny(const CUtensorMap& tw, const CUtensorMap& tx, SkinnyParams p, int ctas, cudaStream_t st) {
  using C = SCfg<NB, CG>;
  static bool configured = false;
  // never more stages than K steps per CTA: the o-projection (4 K steps per split) then needs 97 KB instead of 193 KB and a
  // CTA of it fits beside a running attention CTA, so its whole weight slice is requested before griddepcontrol.wait
  p.stages = skinny_stages(C::kStageBytes);
  if (p.kb_per_split < p.stages) p.stages = p.kb_per_split < 2 ? 2 : p.kb_per_split;
  static const int qkv_stages = getenv("SVLA_SKINNY_QKV_STAGES") ? atoi(getenv("SVLA_SKINNY_QKV_STAGES")) : 0;   // A/B switch
  if (qkv_stages >= 2 && qkv_stages <= 8 && p.kb_per_split == 9 && p.stages > qkv_stages) p.stages = qkv_stages;
  const int smem_bytes = C::smem_bytes(p.stages);
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(svla_gemm_skinny_kernel<NB, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::smem_bytes(skinny_stages(C::kStageBytes)));
    if (e != cudaSuccess) {
      svla_set_error("svla_gemm_skinny: smem opt-in failed: %s", cudaGetErrorString(e));
      return -2;
    }
    configured = true;
  }
  cudaError_t le = CG == 1 ? svla_launch_pdl(svla_gemm_skinny_kernel<NB, CG>, dim3(ctas), dim3(kThreads), smem_bytes, st, tw, tx, p)
                           : svla_launch_pdl_cluster(svla_gemm_skinny_kernel<NB, CG>, dim3(ctas), dim3(kThreads), smem_bytes, st, CG, tw, tx, p);
  if (le != cudaSuccess) {
    svla_set_error("svla_gemm_skinny: launch failed: %s", cudaGetErrorString(le));
    return -2;
  }
  SVLA_LAUNCH_CHECK("svla_gemm_skinny");
  return 0;
}

}  // namespace

extern "C" int svla_gemm_skinny_splits(int64_t n, int64_t k) {
  const int n_tiles = static_cast<int>((n + kWM - 1) / kWM);
  const int kblocks = static_cast<int>((k + kBK - 1) / kBK);
  int splits = svla_num_sms() / n_tiles;
  if (splits < 1) splits = 1;
  while (splits > 1 && (kblocks + splits - 1) / splits < 4) --splits;     // at least 4 K steps (64 KB of weights) per CTA
  const int per = (kblocks + splits - 1) / splits;
  return (kblocks + per - 1) / per;                                        // no empty trailing split
}

extern "C" int svla_gemm_skinny(const SvlaSkinnyArgs* g, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SVLA_REQUIRE(g && g->x && g->w, "svla_gemm_skinny: null operand");
  SVLA_REQUIRE(g->m > 0 && g->m <= 128 && g->n > 0 && g->k > 0, "svla_gemm_skinny: need 0 < m <= 128 (m=%lld)", (long long)g->m);
  SVLA_REQUIRE((g->ldx % 8) == 0 && g->ldx >= g->k && (((g->flags & 4) != 0) || ((g->ldw % 8) == 0 && g->ldw >= g->k)), "svla_gemm_skinny: bad leading dimensions");
  const bool geglu = (g->flags & 1) != 0, partial = (g->flags & 2) != 0, w_tiled = (g->flags & 4) != 0;
  const bool x_hilo = (g->flags & 8) != 0, out_hilo = (g->flags & 16) != 0;
  SVLA_REQUIRE(!x_hilo || g->m <= 64, "svla_gemm_skinny: X_HILO needs m <= 64 (m=%lld)", (long long)g->m);
  SVLA_REQUIRE(!out_hilo || (g->out_bf16 && !partial), "svla_gemm_skinny: OUT_HILO needs a bf16 output");
  const int splits = g->splits > 0 ? g->splits : 1;
  SVLA_REQUIRE(partial || splits == 1, "svla_gemm_skinny: split-K needs the PARTIAL flag (fp32 partial sums)");
  SVLA_REQUIRE(!partial || (g->out_f32 && !g->out_bf16 && !geglu), "svla_gemm_skinny: PARTIAL writes out_f32 only");
  SVLA_REQUIRE(!geglu || (g->out_bf16 && !g->out_f32 && (g->n % 2) == 0), "svla_gemm_skinny: GEGLU needs out_bf16 only and even n");
  SVLA_REQUIRE(g->out_bf16 || g->out_f32, "svla_gemm_skinny: no output");
  SkinnyParams p{};
  p.bias = g->bias; p.out_bf16 = static_cast<__nv_bfloat16*>(g->out_bf16); p.out_f32 = g->out_f32;
  p.m = g->m; p.n = g->n; p.ldo = g->ldo; p.partial_stride = g->partial_stride;
  p.alpha = g->alpha; p.act_param = g->act_param; p.act = g->act; p.flags = g->flags;
  p.n_tiles = static_cast<int>((g->n + kWM - 1) / kWM);
  p.num_k_blocks = static_cast<int>((g->k + kBK - 1) / kBK);
  p.kb_per_split = (p.num_k_blocks + splits - 1) / splits;
  SVLA_REQUIRE(static_cast<long long>(p.kb_per_split) * (splits - 1) < p.num_k_blocks, "svla_gemm_skinny: too many splits (%d) for k=%lld", splits, (long long)g->k);
  // X_HILO: the two planes of m <= 8 (32, 64) rows fill the halves of a 16 (64, 128) column tile
  const int nb = x_hilo ? (g->m <= 8 ? 16 : (g->m <= 32 ? 64 : 128)) : (g->m <= 16 ? 16 : (g->m <= 64 ? 64 : 128));
  p.x_lo_row = x_hilo ? static_cast<int>(g->m) : 0;
  CUtensorMap tw, tx;
  p.w_tiled = w_tiled ? 1 : 0;
  int rc = w_tiled ? encode_kmajor(&tw, g->w, kBK, static_cast<uint64_t>(p.n_tiles) * p.num_k_blocks * kWM, kBK, kWM)
                   : encode_kmajor(&tw, g->w, static_cast<uint64_t>(g->k), static_cast<uint64_t>(g->n), static_cast<uint64_t>(g->ldw), kWM);
  SVLA_REQUIRE(rc == 0, "svla_gemm_skinny: cuTensorMapEncodeTiled(W) failed (%d)", rc);
  // CTA pairs (cta_group::2, flag 32 or SVLA_SKINNY_CG=2) stage half an activation tile each.  Built to halve the L2 -> shared
  // memory traffic of the hi/lo activation tile and MEASURED on B200 (batch-64 decode step, tools/decode_step_perf.py): correct
  // (all skinny cases green with pairs) but not faster -- hi/lo chain 2097 us with pairs vs 2072 us without, plain chain 2556
  // vs 1990 us: co-scheduling two CTAs per cluster costs the PDL chain more than the activation traffic saves.  Opt-in only.
  static const int cg_env = getenv("SVLA_SKINNY_CG") ? atoi(getenv("SVLA_SKINNY_CG")) : kDefaultCG;
  const bool pair = (cg_env == 2 || (g->flags & 32) != 0) && nb >= 64 && p.n_tiles >= 2;
  rc = encode_kmajor(&tx, g->x, static_cast<uint64_t>(g->k), static_cast<uint64_t>(x_hilo ? 2 * g->m : g->m), static_cast<uint64_t>(g->ldx),
                     static_cast<uint32_t>((x_hilo || pair) ? nb / 2 : nb));
  SVLA_REQUIRE(rc == 0, "svla_gemm_skinny: cuTensorMapEncodeTiled(X) failed (%d)", rc);
  const int cg = pair ? 2 : 1;
  p.n_tiles_grid = (p.n_tiles + cg - 1) / cg * cg;
  const int ctas = p.n_tiles_grid * splits;
  if (nb == 16) return launch_skinny<16, 1>(tw, tx, p, ctas, st);
  if (nb == 64) return cg == 2 ? launch_skinny<64, 2>(tw, tx, p, ctas, st) : launch_skinny<64, 1>(tw, tx, p, ctas, st);
  return cg == 2 ? launch_skinny<128, 2>(tw, tx, p, ctas, st) : launch_skinny<128, 1>(tw, tx, p, ctas, st);
}
