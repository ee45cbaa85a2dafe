// G2 on the 5th-generation tensor cores: flash attention with tcgen05.mma, accumulators and P in TMEM, operands by TMA.
//
// One CTA = 128 query ROWS of one (batch, head group); 10 warps:
//   warp 0   : TMA producer   Q once (5-D map {d, g, head group, token, batch}); K_j and V_j tiles through SEPARATE mbarrier
//                             rings (4-D maps {d, head, token, batch}): a K stage is released by Q K_j^T, a V stage by P_j V_j, so
//                             K_{j+1} is requested two softmax periods before it is needed (with one shared ring of depth 2 at
//                             d = 256 every tile paid a full TMA round trip: profiles/attn_timeline_r2_before.txt)
//   warp 1   : MMA issuer     S_j = Q K_j^T          (SS: A = Q smem, B = K_j smem, both K-major, SWIZZLE_128B)
//                             O  += P_j V_j          (TS: A = P_j in TMEM (bf16), B = V_j smem, MN-major, SWIZZLE_128B)
//              issue order QK_0 QK_1 PV_0 QK_2 PV_1 ...: the tensor core computes S_{j+1} during the softmax of tile j
//   warps 2-9: softmax        two warps per TMEM lane quadrant: warp w and w+4 own the same 32 rows and split the 64 keys of a
//                             tile (32 columns each); the row max is exchanged through shared memory under a 64-thread named
//                             barrier; thread == row (tcgen05.ld 32x32b): row max / row sum need NO shuffles; scores are
//                             transformed in the log2 domain (scale, BEiT rel-pos bias, tanh soft-capping, masks); P_j is written
//                             as bf16 pairs OVER the first 32 columns of S_j (tcgen05.st) and becomes the A operand of the second
//                             MMA; lazy rescaling of O (only when the running max grows by more than 8 log2 units).
//   epilogue : O / l -> bf16 -> the (dead) Q tile in shared memory, SWIZZLE_128B -> ONE TMA store per 64-column chunk (the
//              per-thread 16-byte global stores of the first version cost 3 us of a 15 us CTA at d = 256).
// Grouped-query packing: with hq = 2 hkv (Gemma2) the two query heads of a KV head share one CTA: row r = 2 * token + g, fetched
// and stored through the g dimension of the 5-D maps (no copies).  A 278-token prefill then fills 4.3 tiles of 128 rows per KV
// head instead of 2 x 2.2, and a causal tile needs only the keys up to ITS 64 tokens.
// TMEM budget (columns): S double buffer 2*64 (P_j aliased onto S_j: tcgen05.mma executes in issue order, so P_j V_j has read
// P_j before Q K_{j+2}^T overwrites the buffer) | O: D   ->  256 (D <= 128, 2 CTAs/SM) / 512 (D = 256).
// Shapes on this path: BEiT d=64 (S=577, rel-pos bias), Gemma2 prefill d=256 (S=278, GQA 8:4, soft-cap), SigLIP d=72 (128-wide
// tile) and the ZoeDepth router d=32 (64-wide tile): d is its own tensor-map dimension, so columns >= d are zero-filled on load
// and clipped on store by the TMA unit -- no padded copies in HBM.
// Reference ops replaced: model/modeling_gemma2.py:169-195, HF beit/modeling_beit.py:225-306,511-590.
#include <cudaTypedefs.h>
#include <cstdlib>
#include "../../include/spatialvla_b200.h"
#include "tc_ptx.cuh"

namespace svla_attn_tc {
using namespace svla_ptx;

#ifdef SVLA_ATTN_TIMELINE
// profiling build (tools/attn_timeline.py): clock64() stamps of one CTA's producer / issuer / two softmax warps
__device__ unsigned long long g_tl[2048];
#define STAMP(slot) do { if (tl) tl[slot] = clock64(); } while (0)
#else
#define STAMP(slot) do { } while (0)
#endif

constexpr int kBQ = 128;
constexpr int kBKV = 64;
constexpr int kSoftWarps = 8;
constexpr int kThreads = 64 + 32 * kSoftWarps;
constexpr float kLog2e = 1.4426950408889634f;

struct Params {
  int hq, hkv, sq, sk, d;
  int gshift;          // log2(query heads packed into one CTA): row = (token << gshift) | g
  float scale, softcap;
  int causal;
  const float* relpos;
  int win, relpitch;   // rel-pos table rows are stored with pitch `relpitch` floats in shared memory (bank-conflict-free gathers)
  int head_major;      // relpos is [hq][nrel] (coalesced per-head row) instead of HF's [nrel][hq]
  const int* kv_start; // [batch] or null: keys < kv_start[b] are masked (left-padded prompts)
  int prefix;          // causal only: keys < prefix are visible to every query (prefix-LM training mask)
  int window;          // > 0: key slot j is masked for query slot i when i - j >= window (Gemma2 sliding-window layers)
  float* lse;          // null or fp32 [batch, hq, lse_stride]: log2-domain logsumexp of every query row (kept for the backward pass)
  long long lse_stride;
};

template <int D> struct Cfg {
  static constexpr int kChunks = D / 64;                      // 64-column (128-byte) swizzle chunks per row
  static constexpr int kQBytes = kBQ * D * 2;
  static constexpr int kKBytes = kBKV * D * 2;                // one K (or V) tile
  static constexpr int kStages = (D <= 64) ? 4 : 2;           // depth of the K ring and of the V ring
  static constexpr int kTmemS = 0;                            // 2 * kBKV columns; P_j (kBKV / 2 columns) aliases S_j
  static constexpr int kTmemO = 2 * kBKV;                     // D columns
  static constexpr int kTmemUsed = 2 * kBKV + D;
  static constexpr int kTmemCols = kTmemUsed <= 256 ? 256 : 512;
  static constexpr int kCtasPerSm = kTmemCols <= 256 ? 2 : 1;
  static constexpr int kSmemBytes = kQBytes + 2 * kStages * kKBytes + 1024 /*align slack*/ + 256 /*barriers + TMEM slot*/;
};

// kind::f16 instruction descriptor (D = f32, A = B = bf16) with selectable B major-ness (bit 16: 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
// MN-major SWIZZLE_128B operand, canonical layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in elements: 128-byte rows hold
// 64 consecutive MN elements of one k; 8 k rows per swizzle atom (SBO = 1024 B between 8-row groups); LBO = byte
// distance between consecutive 64-element MN blocks.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
      "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// MODE: bit0 rel-pos bias, bit1 soft-cap, bit2 causal.  The real head dimension p.d may be SMALLER than the tile width D.
template <int D, int MODE>
__global__ void __launch_bounds__(kThreads, Cfg<D>::kCtasPerSm)
svla_flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                          const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const Params p) {
  using C = Cfg<D>;
  constexpr int BKV = kBKV;
  constexpr bool kRelpos = (MODE & 1) != 0, kSoftcap = (MODE & 2) != 0, kCausal = (MODE & 4) != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                            // [chunk][128 rows][128 B]; re-used as the O staging tile
  uint8_t* sK = smem + C::kQBytes;                               // [stage][chunk][BKV rows][128 B]
  uint8_t* sV = sK + C::kStages * C::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::kStages * C::kKBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // [4]
  uint64_t* k_empty = bars + 5;            // [4]  (Q K_j^T done)
  uint64_t* v_full = bars + 9;             // [4]
  uint64_t* v_empty = bars + 13;           // [4]  (P_j V_j done)
  uint64_t* s_full = bars + 17;            // [2]
  uint64_t* p_full = bars + 19;            // [2]  (8 softmax warps)
  uint64_t* o_done = bars + 21;            // [2]  (P_j V_j done: O may be rescaled / read)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 23);
  // rel-pos table (pre-multiplied by log2 e) and per-key index terms live after the barriers
  const int nrel_rows = kRelpos ? 2 * p.win - 1 : 0;
  float* sXch = reinterpret_cast<float*>(bars + 24);              // [2 parities][2 column halves][128 rows] row-max exchange
  int* sKterm = reinterpret_cast<int*>(sXch + 4 * kBQ);           // [n_tiles_all * BKV] per-key index terms (rel-pos only)
  const int n_tiles_all = (p.sk + BKV - 1) / BKV;
  float* sTab = reinterpret_cast<float*>(sKterm + (kRelpos ? n_tiles_all * BKV : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gsh = p.gshift;
  const int bqt = kBQ >> gsh;                                     // tokens per CTA
  const int b = blockIdx.z, hg = blockIdx.y, q0 = blockIdx.x * bqt;
  const int hk = gsh ? hg : hg / (p.hq / p.hkv);
  const int causal_off = p.sk - p.sq;
  int n_tiles = n_tiles_all;
  if (kCausal) n_tiles = max(1, min(n_tiles, (max(min(q0 + bqt, p.sq) + causal_off, p.prefix) + BKV - 1) / BKV));   // tiles above the diagonal are skipped
#ifdef SVLA_ATTN_TIMELINE
  unsigned long long* tl = nullptr;
  if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 20 && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 6))
    tl = g_tl + (warp == 6 ? 3 : warp) * 512;
#endif
  STAMP(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], kSoftWarps);
      mbar_init(&o_done[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::kTmemCols, 1>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  STAMP(1);

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, C::kQBytes);
#pragma unroll
      for (int c = 0; c < C::kChunks; ++c) tma_load_5d(sQ + c * (kBQ * 128), &tm_q, q_full, c * 64, 0, hg, q0, b);
      auto load_k = [&](int j) {
        const int st = j % C::kStages;
        mbar_wait_suspend(&k_empty[st], ((j / C::kStages) & 1) ^ 1u);
        STAMP(8 + j * 8);
        mbar_expect_tx(&k_full[st], C::kKBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(sK + st * C::kKBytes + c * (BKV * 128), &tm_k, &k_full[st], c * 64, hk, j * BKV, b);
      };
      auto load_v = [&](int j) {
        const int st = j % C::kStages;
        mbar_wait_suspend(&v_empty[st], ((j / C::kStages) & 1) ^ 1u);
        STAMP(8 + j * 8 + 1);
        mbar_expect_tx(&v_full[st], C::kKBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(sV + st * C::kKBytes + c * (BKV * 128), &tm_v, &v_full[st], c * 64, hk, j * BKV, b);
      };
      load_k(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    // The whole warp runs the loop (converged); one ELECTED lane issues the MMAs and commits of a phase.  Issued from a
    // divergent `lane == 0` region every tcgen05.mma became an elect/branch loop plus a descriptor rebuild (~70 clocks per MMA,
    // 16 + 4 MMAs per tile at d = 256: the issue rate, not the tensor pipe, bounded the chain).  Descriptors: 64-bit constants
    // built once; a k-step only adds to the 14-bit start-address field.
    {
      const int d16 = (p.d + 15) & ~15;
      const uint32_t idesc_pv = make_idesc(kBQ, d16, 1);                              // P V output columns
      const int ksteps_qk = d16 >> 4;                                                 // Q K^T contraction steps
      // the last K/V tile only multiplies the 16-key groups that hold valid keys (BEiT: 577 = 9 * 64 + 1 keys)
      auto valid16 = [&](int j) { return min(BKV, (p.sk - j * BKV + 15) & ~15); };
      const uint64_t q_desc = make_kmajor_sw128_desc(smem_u32(sQ));
      auto issue_pv = [&](int j) {
        const int st = j & 1, kst = j % C::kStages;
        mbar_wait_suspend(&v_full[kst], (j / C::kStages) & 1);
        mbar_wait_suspend(&p_full[st], (j >> 1) & 1);
        STAMP(8 + j * 8 + 3);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t v_desc = make_mnmajor_sw128_desc(smem_u32(sV + kst * C::kKBytes), BKV * 128);
          const uint32_t p_addr = tmem_base + C::kTmemS + st * BKV;
          const int ksteps = valid16(j) >> 4;
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk) {
            if (kk < ksteps)
              umma_bf16_ts(tmem_base + C::kTmemO, p_addr + kk * 8, v_desc + static_cast<uint64_t>(kk * (16 * 128 >> 4)), idesc_pv,
                           static_cast<uint32_t>(j > 0 || kk > 0));
          }
          umma_commit(&o_done[st]);        // P_j consumed, O updated
          umma_commit(&v_empty[kst]);      // V_j stage free
        }
        __syncwarp();
        STAMP(8 + j * 8 + 4);
      };
      mbar_wait_suspend(q_full, 0);
      STAMP(2);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1, kst = j % C::kStages;
        mbar_wait_suspend(&k_full[kst], (j / C::kStages) & 1);
        STAMP(8 + j * 8);
        tc_fence_after();
        // S buffer st: its previous contents (S_{j-2}, then P_{j-2}) were consumed by P_{j-2} V_{j-2}, issued before this MMA
        if (elect_one()) {
          const uint64_t k_desc = make_kmajor_sw128_desc(smem_u32(sK + kst * C::kKBytes));
          const uint32_t idesc_qk = make_idesc(kBQ, valid16(j), 0);
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            if (kk < ksteps_qk)
              umma_bf16(tmem_base + C::kTmemS + st * BKV, q_desc + static_cast<uint64_t>(((kk >> 2) * (kBQ * 128) + (kk & 3) * 32) >> 4),
                        k_desc + static_cast<uint64_t>(((kk >> 2) * (BKV * 128) + (kk & 3) * 32) >> 4), idesc_qk, static_cast<uint32_t>(kk > 0));
          }
          umma_commit(&s_full[st]);
          umma_commit(&k_empty[kst]);      // K_j stage free
        }
        __syncwarp();
        STAMP(8 + j * 8 + 2);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(n_tiles - 1);
    }
  } else {
    // ============================================================ softmax / correction / epilogue: thread == row
    constexpr int HC = BKV / 2;                       // key columns of a tile owned by this warp
    constexpr int DH = D / 2;                         // O columns rescaled / written by this warp
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;                   // column half: warps 2-5 -> 0, warps 6-9 -> 1
    const int row = q * 32 + lane;
    const int qi = q0 + (row >> gsh);                 // token of this row
    const int h = (hg << gsh) | (row & ((1 << gsh) - 1));       // query head of this row
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    int qbase = 0;
    const bool cls_q = kRelpos && (qi == 0);
    const int cls_off = nrel_rows * p.relpitch;       // the three CLS entries follow the (2w-1) x pitch grid
    if (kRelpos) {
      const int qc = min(qi, p.sq - 1);               // rows past the end reuse the last valid row's (in-range) index
      const int qp = max(qc - 1, 0);                  // the CLS query row uses an in-range dummy (its bias is a constant, see below)
      qbase = (qp / p.win + p.win - 1) * p.relpitch + qp % p.win + p.win - 1;
    }
    const float sl2 = p.scale * kLog2e;
    const float c1 = kSoftcap ? p.scale / p.softcap : 0.f, c2 = kSoftcap ? p.softcap * kLog2e : 0.f;
    const int kstart = p.kv_start ? p.kv_start[b] : 0;
    if constexpr (kRelpos) {
      // this head's bias table (x log2 e) and the per-key index terms, loaded by the softmax warps only: the TMA producer and
      // the MMA issuer are already running.  Table rows are re-pitched (pitch == win mod 32) so that the 32 rows of a warp,
      // which span up to three grid rows, gather from 32 different banks.
      const int tid = threadIdx.x - 64;
      constexpr int NT = 32 * kSoftWarps;
      const int nr2 = nrel_rows * nrel_rows;
      const float* src = p.relpos + (p.head_major ? static_cast<long long>(h) * (nr2 + 3) : h);
      const long long sstride = p.head_major ? 1 : p.hq;
      if (p.head_major && ((nr2 + 3) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        // contiguous per-head table: every thread issues its (up to 3) 16-byte loads BEFORE the first use (one global round
        // trip for the whole table; the strided gather below cost ~3000 clocks of every CTA's prologue), then scatters the
        // four elements into the pitched rows (row / column by a float reciprocal: exact below 2^20); the three CLS entries at the
        // end of the row land at r == nrel_rows, i.e. at cls_off + c
        const float inv_rows = 1.f / static_cast<float>(nrel_rows);
        float4 t[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int i4 = tid + u * NT;
          t[u] = (4 * i4 < nr2 + 3) ? __ldg(reinterpret_cast<const float4*>(src) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int i0 = 4 * (tid + u * NT);
          if (i0 < nr2 + 3) {
            const float tv[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = i0 + e, r = __float2int_rd((static_cast<float>(i) + 0.5f) * inv_rows), c = i - r * nrel_rows;
              sTab[r * p.relpitch + c] = tv[e] * kLog2e;
            }
          }
        }
        for (int i4 = tid + 3 * NT; 4 * i4 < nr2 + 3; i4 += NT) {          // tables beyond 3072 entries (not on this path)
          const float4 tt = __ldg(reinterpret_cast<const float4*>(src) + i4);
          const float tv[4] = {tt.x, tt.y, tt.z, tt.w};
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e, r = i / nrel_rows, c = i - r * nrel_rows;
            sTab[r * p.relpitch + c] = tv[e] * kLog2e;
          }
        }
      } else {
        // thread (r4, c) copies column c of table rows r4, r4 + 4, ...
        for (int c = tid & 63; c < nrel_rows; c += 64)
          for (int r = tid >> 6; r < nrel_rows; r += NT / 64)
            sTab[r * p.relpitch + c] = __ldg(src + (r * nrel_rows + c) * sstride) * kLog2e;
        if (tid < 3) sTab[cls_off + tid] = __ldg(src + (nr2 + tid) * sstride) * kLog2e;
      }
      const float inv_win = 1.f / static_cast<float>(p.win);
      for (int kj = tid; kj < n_tiles_all * BKV; kj += NT) {
        const int g0 = kj - 1, ky = __float2int_rd((static_cast<float>(g0) + 0.5f) * inv_win), kx = g0 - ky * p.win;      // exact for g0 < 2^20
        sKterm[kj] = (kj >= 1 && kj < p.sk) ? 4 * (ky * p.relpitch + kx) : 0;      // byte offsets
      }
      asm volatile("bar.sync 5, 256;" ::: "memory");      // softmax warps only (ids 1-4 are the pair barriers)
    }
    // the two warps of a quadrant meet on named barrier 1 + q (64 threads)
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const int k0 = j * BKV + ch * HC;               // first key of this warp's columns
      mbar_wait_suspend(&s_full[st], (j >> 1) & 1);
      STAMP(8 + j * 8);
      tc_fence_after();
      float s[HC];
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + C::kTmemS + st * BKV + ch * HC + lane_addr, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
      }
      STAMP(8 + j * 8 + 1);
      // ---- scores -> log2 domain (skipped when every column of this warp lies behind the last key, e.g. BEiT's 577 = 9*64 + 1)
      if (k0 >= p.sk) {
#pragma unroll
        for (int i = 0; i < HC; ++i) s[i] = -INFINITY;
      } else if (kRelpos) {
        // bias[q, k] = table[qbase - kterm[k]]: the per-key terms are BYTE offsets (one int4 broadcast load per 4 keys), the
        // per-thread base is a byte address, so an element costs IADD + LDS + half an FFMA2.  The CLS query row (constant bias
        // over all keys) exists in one warp pair of the first query tile only: that warp takes the select-per-element path.
        const float raw0 = s[0];
        const int4* kt4 = reinterpret_cast<const int4*>(sKterm + k0);
        const char* qtab = reinterpret_cast<const char*>(sTab) + 4 * qbase;
        if (q0 == 0 && q == 0) {
          const char* tab0 = reinterpret_cast<const char*>(sTab);
          const int clsb = 4 * cls_off, qoff = 4 * qbase;
#pragma unroll
          for (int i4 = 0; i4 < HC / 4; ++i4) {
            const int4 kt = kt4[i4];                          // warp-wide broadcast
            const int i0 = cls_q ? clsb : qoff - kt.x, i1 = cls_q ? clsb : qoff - kt.y;
            const int i2 = cls_q ? clsb : qoff - kt.z, i3 = cls_q ? clsb : qoff - kt.w;
            s[4 * i4 + 0] = fmaf(s[4 * i4 + 0], sl2, *reinterpret_cast<const float*>(tab0 + i0));
            s[4 * i4 + 1] = fmaf(s[4 * i4 + 1], sl2, *reinterpret_cast<const float*>(tab0 + i1));
            s[4 * i4 + 2] = fmaf(s[4 * i4 + 2], sl2, *reinterpret_cast<const float*>(tab0 + i2));
            s[4 * i4 + 3] = fmaf(s[4 * i4 + 3], sl2, *reinterpret_cast<const float*>(tab0 + i3));
          }
        } else {
          const uint64_t sl22 = pack_f32x2(sl2, sl2);
#pragma unroll
          for (int i4 = 0; i4 < HC / 4; ++i4) {
            const int4 kt = kt4[i4];                          // warp-wide broadcast
            const float b0 = *reinterpret_cast<const float*>(qtab - kt.x), b1 = *reinterpret_cast<const float*>(qtab - kt.y);
            const float b2 = *reinterpret_cast<const float*>(qtab - kt.z), b3 = *reinterpret_cast<const float*>(qtab - kt.w);
            unpack_f32x2(fma_f32x2(pack_f32x2(s[4 * i4 + 0], s[4 * i4 + 1]), sl22, pack_f32x2(b0, b1)), s[4 * i4 + 0], s[4 * i4 + 1]);
            unpack_f32x2(fma_f32x2(pack_f32x2(s[4 * i4 + 2], s[4 * i4 + 3]), sl22, pack_f32x2(b2, b3)), s[4 * i4 + 2], s[4 * i4 + 3]);
          }
        }
        if (k0 == 0) s[0] = fmaf(raw0, sl2, sTab[cls_off + (cls_q ? 2 : 1)]);      // CLS key column
      } else if (kSoftcap) {
        // cap * tanh(u) * log2e, u = s * scale / cap, with a degree-9 odd polynomial (exact to fp32 rounding for |u| < 0.35), two
        // scores per FMUL2 / FFMA2 (the scalar chain was ~11 of this warp's ~14 instructions per score at d = 256)
        const uint64_t c1p = pack_f32x2(c1, c1), c2p = pack_f32x2(c2, c2);
        const uint64_t k9 = pack_f32x2(62.f / 2835.f, 62.f / 2835.f), k7 = pack_f32x2(-17.f / 315.f, -17.f / 315.f);
        const uint64_t k5 = pack_f32x2(2.f / 15.f, 2.f / 15.f), k3 = pack_f32x2(-1.f / 3.f, -1.f / 3.f), k1 = pack_f32x2(1.f, 1.f);
        float t[HC];
        float u2max = 0.f;
#pragma unroll
        for (int i = 0; i < HC; i += 2) {
          const uint64_t u = mul_f32x2(pack_f32x2(s[i], s[i + 1]), c1p), u2 = mul_f32x2(u, u);
          float q0, q1;
          unpack_f32x2(u2, q0, q1);
          u2max = fmaxf(u2max, fmaxf(q0, q1));
          uint64_t pl = fma_f32x2(k9, u2, k7);
          pl = fma_f32x2(pl, u2, k5);
          pl = fma_f32x2(pl, u2, k3);
          pl = fma_f32x2(pl, u2, k1);
          unpack_f32x2(mul_f32x2(mul_f32x2(c2p, u), pl), t[i], t[i + 1]);
        }
        if (!__any_sync(0xffffffffu, u2max >= 0.1225f)) {
#pragma unroll
          for (int i = 0; i < HC; ++i) s[i] = t[i];
        } else {                                  // rare: a large score somewhere in this warp's rows -> libm tanh
#pragma unroll
          for (int i = 0; i < HC; ++i) s[i] = c2 * tanhf(s[i] * c1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < HC; ++i) s[i] *= sl2;
      }
      if (kCausal || k0 + HC > p.sk || kstart > k0 || p.window > 0) {
#pragma unroll
        for (int i = 0; i < HC; ++i) {
          const int kj = k0 + i;
          if (kj >= p.sk || kj < kstart || (kCausal && kj > max(qi + causal_off, p.prefix - 1)) ||
              (p.window > 0 && qi + causal_off - kj >= p.window)) s[i] = -INFINITY;
        }
      }
      float mj = -INFINITY;
#pragma unroll
      for (int i = 0; i < HC; ++i) mj = fmaxf(mj, s[i]);
      // ---- row max over the whole tile: exchange with the partner warp (parity-double-buffered, one barrier per tile).  The
      // barrier also orders this pair's S_j loads before the P_j stores below (P_j aliases the first 32 columns of S_j).
      sXch[(st * 2 + ch) * kBQ + row] = mj;
      tc_fence_before();
      pair_sync();
      tc_fence_after();
      mj = fmaxf(mj, sXch[(st * 2 + (ch ^ 1)) * kBQ + row]);
      STAMP(8 + j * 8 + 2);
      // ---- lazy rescale: keep the stale max unless it grows by more than 8 (P <= 2^8: bf16/fp32 lose nothing)
      float alpha = 1.f;
      bool rescale = false;
      if (mj > m_run + 8.f || (m_run == -INFINITY && mj != -INFINITY)) {
        alpha = (m_run == -INFINITY) ? 0.f : ex2f(m_run - mj);
        rescale = (j > 0) && (m_run != -INFINITY);
        m_run = mj;
      }
      const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
      uint32_t pk[HC / 2];
      uint64_t lsum2 = pack_f32x2(0.f, 0.f);
      const uint64_t negm2 = pack_f32x2(-m_use, -m_use);
#pragma unroll
      for (int i = 0; i < HC; i += 2) {
        float t0, t1;
        unpack_f32x2(add_f32x2(pack_f32x2(s[i], s[i + 1]), negm2), t0, t1);
        const float p0 = ex2f(t0), p1 = ex2f(t1);
        lsum2 = add_f32x2(lsum2, pack_f32x2(p0, p1));
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      float lsum, lsum_hi;
      unpack_f32x2(lsum2, lsum, lsum_hi);
      lsum += lsum_hi;
      l_run = l_run * alpha + lsum;                          // partial row sum over this warp's columns
      STAMP(8 + j * 8 + 3);
      // a rescale needs P_{j-1} V_{j-1} (the last writer of O) done
      if (__any_sync(0xffffffffu, rescale)) {                // identical decision in both warps of the pair (same rows, same max)
        mbar_wait_suspend(&o_done[st ^ 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = ch * DH; c0 < (ch + 1) * DH; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + C::kTmemO + c0 + lane_addr, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(tmem_base + C::kTmemO + c0 + lane_addr, r);
        }
      }
      STAMP(8 + j * 8 + 4);
      tmem_st16(tmem_base + C::kTmemS + st * BKV + ch * (HC / 2) + lane_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
      STAMP(8 + j * 8 + 5);
    }
    // ---- epilogue: total row sum (both column halves), wait for the last P V, O / l -> bf16 -> shared memory -> TMA store
    pair_sync();                                             // the partner has finished reading the exchange buffers
    sXch[ch * kBQ + row] = l_run;
    pair_sync();
    l_run += sXch[(ch ^ 1) * kBQ + row];
    // training: keep log2-sum-exp of the row (scores are in the log2 domain; m_run is the -- possibly stale -- max l_run refers to)
    if (p.lse && ch == 0 && qi < p.sq)
      p.lse[(static_cast<long long>(b) * p.hq + h) * p.lse_stride + qi] = (l_run > 0.f) ? m_run + log2f(l_run) : -INFINITY;
    const int last = n_tiles - 1;
    STAMP(3);
    mbar_wait_suspend(&o_done[last & 1], (last >> 1) & 1);   // every MMA of this CTA has completed: the Q tile is dead
    STAMP(4);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const int d16 = (p.d + 15) & ~15;
    const uint32_t row_s = smem_u32(sQ) + static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
#pragma unroll 1
    for (int c0 = ch * DH; c0 < (ch + 1) * DH; c0 += 32) {
      if (c0 >= d16) break;
      uint32_t r[32];
      tmem_ld32(tmem_base + C::kTmemO + c0 + lane_addr, r);
      const uint32_t chunk_s = row_s + static_cast<uint32_t>(c0 >> 6) * (kBQ * 128);
      const uint32_t u0 = static_cast<uint32_t>(c0 & 63) >> 3;           // first 16-byte unit of these 32 columns in the 128-byte row
#pragma unroll
      for (uint32_t v8 = 0; v8 < 4; ++v8)
        st_shared_v4(chunk_s + (((u0 + v8) ^ sw) << 4),
                     pack_bf16x2(__uint_as_float(r[8 * v8 + 0]) * inv, __uint_as_float(r[8 * v8 + 1]) * inv),
                     pack_bf16x2(__uint_as_float(r[8 * v8 + 2]) * inv, __uint_as_float(r[8 * v8 + 3]) * inv),
                     pack_bf16x2(__uint_as_float(r[8 * v8 + 4]) * inv, __uint_as_float(r[8 * v8 + 5]) * inv),
                     pack_bf16x2(__uint_as_float(r[8 * v8 + 6]) * inv, __uint_as_float(r[8 * v8 + 7]) * inv));
    }
    // each warp pair stores ITS 32 rows as soon as both column halves are staged (boxes of 32 rows: no CTA-wide barrier, four
    // independent store streams)
    fence_proxy_async_smem();
    pair_sync();
    if (ch == 0 && lane == 0) {
#pragma unroll
      for (int c = 0; c < C::kChunks; ++c)
        if (c * 64 < p.d) tma_store_5d(&tm_o, sQ + c * (kBQ * 128) + q * (32 * 128), c * 64, 0, hg, q0 + ((q * 32) >> gsh), b);     // rows >= sq / columns >= d are clipped
      bulk_commit_group();
      bulk_wait_read_all();                                  // shared memory must stay valid until the TMA unit has read it
    }
    STAMP(5);
  }
  tc_fence_before();
  __syncthreads();
  STAMP(6);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols, 1>(tmem_base);
  }
  STAMP(7);
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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

// 4-D view {d, head, tokens, batch} of a strided bf16 activation; box = 64 columns x box_rows tokens of one head.  Columns past
// the real head dimension are out of bounds -> zero fill.
static int encode_kv(CUtensorMap* tm, const void* base, uint64_t d, uint64_t heads, uint64_t tokens, uint64_t batch,
                     uint64_t token_stride, uint64_t batch_stride, uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[4] = {d, heads, tokens, batch};
  cuuint64_t strides[3] = {d * 2, token_stride * 2, batch_stride * 2};
  cuuint32_t box[4] = {64, 1, box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

// 5-D view {d, g, head group, tokens, batch} of Q / O: a box holds 128 / G tokens x G heads x 64 columns, i.e. shared-memory
// row = token * G + g (G query heads of one KV head packed into the M dimension of the MMA)
static int encode_q(CUtensorMap* tm, const void* base, uint64_t d, uint64_t g, uint64_t groups, uint64_t tokens, uint64_t batch,
                    uint64_t token_stride, uint64_t batch_stride, uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[5] = {d, g, groups, tokens, batch};
  cuuint64_t strides[4] = {d * 2, g * d * 2, token_stride * 2, batch_stride * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(g), 1, static_cast<cuuint32_t>(box_rows / g), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

template <int D, int MODE>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, const Params& p, int batch,
                  cudaStream_t st) {
  using C = Cfg<D>;
  const int tab_floats = (MODE & 1) ? (2 * p.win - 1) * p.relpitch + 4 : 0;
  const int n_tiles_all = (p.sk + kBKV - 1) / kBKV;
  const size_t smem = C::kSmemBytes + 4 * kBQ * 4 + ((MODE & 1) ? static_cast<size_t>(n_tiles_all) * kBKV * 4 : 0) + static_cast<size_t>(tab_floats) * 4;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(svla_flash_attn_tc_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      svla_set_error("svla_attention(tcgen05): smem opt-in %zu failed: %s", smem, cudaGetErrorString(e));
      return -2;
    }
    configured = smem;
  }
  const int bqt = kBQ >> p.gshift;
  dim3 grid((p.sq + bqt - 1) / bqt, p.hq >> p.gshift, batch);
  svla_flash_attn_tc_kernel<D, MODE><<<grid, kThreads, smem, st>>>(tq, tk, tv, to, p);
  SVLA_LAUNCH_CHECK("svla_flash_attn_tc");
  return 0;
}

}  // namespace svla_attn_tc

// Returns 1 if the tcgen05 kernel does not cover the problem (the caller then uses the mma.sync kernel), 0 on success,
// < 0 on error.  Covered: the feature sets the model uses -- plain or rel-pos bias at d <= 128 (BEiT 64, SigLIP 72, router 32),
// soft-cap with / without causal mask at d = 256 (Gemma2) -- with 16-byte aligned bases and strides.
int svla_attention_tc_try(const SvlaAttnArgs* a, void* stream) {
  using namespace svla_attn_tc;
  const int mode = (a->relpos_table ? 1 : 0) | (a->softcap > 0.f ? 2 : 0) | (a->causal ? 4 : 0);
  const bool small = (mode == 0 && a->d <= 128 && !a->kv_start) || (mode == 1 && a->d == 64);
  const bool gemma = a->d == 256 && (mode == 2 || mode == 6);
  if (!small && !gemma) return 1;
  if ((a->d % 8) != 0) return 1;
  if ((a->q_ss % 8) || (a->k_ss % 8) || (a->v_ss % 8) || (a->q_bs % 8) || (a->k_bs % 8) || (a->v_bs % 8) || (a->o_ss % 8) || (a->o_bs % 8)) return 1;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->out)) & 15) return 1;
  if (a->causal && a->sk < a->sq) return 1;
  Params p{};
  p.hq = a->hq; p.hkv = a->hkv; p.sq = a->sq; p.sk = a->sk; p.d = a->d;
  p.gshift = (a->hq == 2 * a->hkv && !a->relpos_table) ? 1 : 0;      // grouped-query packing: both query heads of a KV head in one CTA
  p.scale = a->scale; p.softcap = a->softcap; p.causal = a->causal; p.relpos = a->relpos_table; p.win = a->relpos_win; p.head_major = a->relpos_head_major; p.kv_start = a->kv_start;
  if (a->relpos_table) {
    int pitch = 2 * a->relpos_win - 1;
    while ((pitch & 31) != (a->relpos_win & 31)) ++pitch;           // pitch == win (mod 32): conflict-free gathers (see the kernel)
    p.relpitch = pitch;
  }
  p.prefix = a->causal ? a->causal_prefix : 0;
  p.window = a->window;
  p.lse = a->lse;
  p.lse_stride = a->lse_stride;
  // a batch stride of 0 is not expressible in a tensor map; batch == 1 problems get a dummy stride
  const uint64_t nb = static_cast<uint64_t>(a->batch);
  auto bstride = [&](int64_t bs, int64_t ss, int s) { return static_cast<uint64_t>(nb > 1 ? bs : ss * s); };
  if (nb > 1 && (a->q_bs <= 0 || a->k_bs <= 0 || a->v_bs <= 0 || a->o_bs <= 0)) return 1;
  CUtensorMap tq, tk, tv, to;
  const uint64_t dd = static_cast<uint64_t>(a->d), g = 1ull << p.gshift;
  if (encode_q(&tq, a->q, dd, g, a->hq / g, a->sq, nb, a->q_ss, bstride(a->q_bs, a->q_ss, a->sq), kBQ) != 0) return 1;
  if (encode_q(&to, a->out, dd, g, a->hq / g, a->sq, nb, a->o_ss, bstride(a->o_bs, a->o_ss, a->sq), 32) != 0) return 1;
  if (encode_kv(&tk, a->k, dd, a->hkv, a->sk, nb, a->k_ss, bstride(a->k_bs, a->k_ss, a->sk), kBKV) != 0) return 1;
  if (encode_kv(&tv, a->v, dd, a->hkv, a->sk, nb, a->v_ss, bstride(a->v_bs, a->v_ss, a->sk), kBKV) != 0) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gemma) return mode == 2 ? launch<256, 2>(tq, tk, tv, to, p, a->batch, st) : launch<256, 6>(tq, tk, tv, to, p, a->batch, st);
  if (mode == 1) return launch<64, 1>(tq, tk, tv, to, p, a->batch, st);
  return a->d <= 64 ? launch<64, 0>(tq, tk, tv, to, p, a->batch, st) : launch<128, 0>(tq, tk, tv, to, p, a->batch, st);
}

#ifdef SVLA_ATTN_TIMELINE
extern "C" int svla_dbg_attn_timeline(unsigned long long* host2048) {
  return cudaMemcpyFromSymbol(host2048, svla_attn_tc::g_tl, sizeof(unsigned long long) * 2048) == cudaSuccess ? 0 : -1;
}
#endif
