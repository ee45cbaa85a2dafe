// G2 on the 5th-generation tensor cores: flash attention with tcgen05.mma, accumulators and P in TMEM, operands by TMA.
//
// One CTA = 128 query rows of one (batch, head); 10 warps:
//   warp 0   : TMA producer   (Q once; K_j / V_j tiles through a 2-stage mbarrier ring; 3-D tensor maps
//                              {head columns, tokens, batch} so ragged tails are zero-filled by the hardware)
//   warp 1   : MMA issuer     S_j = Q K_j^T          (SS: A = Q smem, B = K_j smem, both K-major, SWIZZLE_128B)
//                             O  += P_j V_j          (TS: A = P_j in TMEM (bf16), B = V_j smem, MN-major, SWIZZLE_128B)
//              software-pipelined: QK_{j+1} is issued before P_j V_j so the tensor core overlaps the softmax of tile j
//   warps 2-9: softmax        two warps per TMEM lane quadrant: warp w and w+4 own the same 32 query rows and split the 64
//                             keys of a tile (32 columns each), so 16 softmax warps per SM (2 CTAs) hide the TMEM / MUFU /
//                             shared-memory latencies that bound the 4-warp version (ncu: one warp per scheduler, issue slots
//                             idle); the row max is exchanged through shared memory under a 64-thread named barrier;
//                             thread == query row (tcgen05.ld 32x32b): row max / row sum need NO shuffles;
//                             scores are transformed in the log2 domain (scale, BEiT rel-pos bias from the per-head
//                             table, tanh soft-capping, ragged / causal mask), P is written back to TMEM as bf16 pairs
//                             (tcgen05.st) and becomes the A operand of the second MMA;
//                             lazy rescaling: O is only multiplied by 2^(m_old - m_new) when the running max grows by
//                             more than 8 (log2 units), the common case leaves O untouched in TMEM;
//                             epilogue: O / l -> bf16 -> global.
// TMEM budget (columns): S double buffer 2*64 | O: D | P double buffer 2*32  ->  256 (D=64, 2 CTAs/SM) / 512 (D=256).
// Shapes on this path: BEiT d=64 (S=577, rel-pos bias), Gemma2 prefill d=256 (S=278, GQA 8:4, soft-cap), and -- through the
// zero-padding 4-D tensor maps of the PAD variant -- SigLIP d=72 (128-wide tile) and the ZoeDepth router d=32 (64-wide tile).
// Reference ops replaced: model/modeling_gemma2.py:169-195, HF beit/modeling_beit.py:225-306,511-590.
#include <cudaTypedefs.h>
#include "../../include/spatialvla_b200.h"
#include "tc_ptx.cuh"

namespace svla_attn_tc {
using namespace svla_ptx;

constexpr int kBQ = 128;
constexpr int kBKV = 64;
constexpr int kSoftWarps = 8;
constexpr int kThreads = 64 + 32 * kSoftWarps;
constexpr float kLog2e = 1.4426950408889634f;

struct Params {
  __nv_bfloat16* out;
  long long o_bs, o_ss;
  int hq, hkv, sq, sk, d;
  float scale, softcap;
  int causal;
  const float* relpos;
  int win;
  int head_major;      // relpos is [hq][nrel] (coalesced per-head row) instead of HF's [nrel][hq]
  const int* kv_start; // [batch] or null: keys < kv_start[b] are masked (left-padded prompts)
  int prefix;          // causal only: keys < prefix are visible to every query (prefix-LM training mask)
  float* lse;          // null or fp32 [batch, hq, lse_stride]: log2-domain logsumexp of every query row (kept for the backward pass)
  long long lse_stride;
};

template <int D> struct Cfg {
  static constexpr int kChunks = D / 64;                      // 64-column (128-byte) swizzle chunks per row
  static constexpr int kQBytes = kBQ * D * 2;
  static constexpr int kKBytes = kBKV * D * 2;
  // K/V ring depth: the producer may only refill a stage after P_j V_j has retired it, so with 2 stages every tile pays a full
  // TMA round trip (measured: ~2 us per 64-key tile at d=64 against ~0.15 us of MMA); d=64 has the shared memory for 4 stages
  static constexpr int kStages = (D <= 128) ? 4 : 2;
  static constexpr int kTmemS = 0;                            // 2 * kBKV columns
  static constexpr int kTmemO = 2 * kBKV;                     // D columns
  static constexpr int kTmemP = 2 * kBKV + D;                 // 2 * kBKV/2 columns
  static constexpr int kTmemUsed = 2 * kBKV + D + kBKV;
  static constexpr int kTmemCols = kTmemUsed <= 256 ? 256 : 512;
  static constexpr int kCtasPerSm = kTmemCols <= 256 ? 2 : 1;
  static constexpr int kSmemBytes = kQBytes + kStages * 2 * kKBytes + 1024 /*align slack*/ + 256 /*barriers + TMEM slot*/;
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

// MODE: bit0 rel-pos bias, bit1 soft-cap, bit2 causal.
// PAD: the real head dimension p.d is SMALLER than the tile width D (SigLIP: 72 in a 128-wide tile; the ZoeDepth router: 32 in a
// 64-wide tile).  The operands are then fetched through 4-D tensor maps {d, head, token, batch}: the box is still 64 columns wide,
// but columns >= p.d lie outside dimension 0 of the tensor and are ZERO-FILLED by the TMA unit, so Q K^T contracts over
// ceil(p.d / 16) k-steps of exact zeros-padded operands and P V produces round16(p.d) output columns -- no padded copies in HBM.
template <int D, int MODE, bool PAD = false>
__global__ void __launch_bounds__(kThreads, Cfg<D>::kCtasPerSm)
svla_flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                          const __grid_constant__ CUtensorMap tm_v, const Params p) {
  using C = Cfg<D>;
  constexpr int BKV = kBKV;
  constexpr bool kRelpos = (MODE & 1) != 0, kSoftcap = (MODE & 2) != 0, kCausal = (MODE & 4) != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                            // [chunk][128 rows][128 B]
  uint8_t* sK = smem + C::kQBytes;                               // [stage][chunk][BKV rows][128 B]
  uint8_t* sV = sK + C::kStages * C::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::kStages * C::kKBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // [4]
  uint64_t* kv_empty = bars + 5;           // [4]
  uint64_t* s_full = bars + 9;             // [2]
  uint64_t* s_empty = bars + 11;           // [2]  (8 softmax warps)
  uint64_t* p_full = bars + 13;            // [2]  (8 softmax warps)
  uint64_t* p_empty = bars + 15;           // [2]  (PV MMA of that tile done)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 17);
  // rel-pos table (pre-multiplied by log2 e) and per-key index terms live after the barriers
  const int nrel = kRelpos ? (2 * p.win - 1) * (2 * p.win - 1) + 3 : 0;
  float* sXch = reinterpret_cast<float*>(bars + 20);              // [2 parities][2 column halves][128 rows] row-max exchange
  int* sKterm = reinterpret_cast<int*>(sXch + 4 * kBQ);           // [n_tiles_all * BKV] per-key index terms (rel-pos only)
  const int n_tiles_all = (p.sk + BKV - 1) / BKV;
  float* sTab = reinterpret_cast<float*>(sKterm + (kRelpos ? n_tiles_all * BKV : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kBQ;
  const int hk = h / (p.hq / p.hkv);
  const int causal_off = p.sk - p.sq;
  int n_tiles = (p.sk + BKV - 1) / BKV;
  if (kCausal) n_tiles = max(1, min(n_tiles, (max(min(q0 + kBQ, p.sq) + causal_off, p.prefix) + BKV - 1) / BKV));   // tiles above the diagonal are skipped

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], kSoftWarps);
      mbar_init(&p_full[s], kSoftWarps);
      mbar_init(&p_empty[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::kTmemCols, 1>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, C::kQBytes);
#pragma unroll
      for (int c = 0; c < C::kChunks; ++c) {
        if constexpr (PAD) tma_load_4d(sQ + c * (kBQ * 128), &tm_q, q_full, c * 64, h, q0, b);
        else tma_load_3d(sQ + c * (kBQ * 128), &tm_q, q_full, h * D + c * 64, q0, b);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % C::kStages;
        mbar_wait_suspend(&kv_empty[st], ((j / C::kStages) & 1) ^ 1u);
        mbar_expect_tx(&kv_full[st], 2 * C::kKBytes);
#pragma unroll
        for (int c = 0; c < C::kChunks; ++c) {
          if constexpr (PAD) {
            tma_load_4d(sK + st * C::kKBytes + c * (BKV * 128), &tm_k, &kv_full[st], c * 64, hk, j * BKV, b);
            tma_load_4d(sV + st * C::kKBytes + c * (BKV * 128), &tm_v, &kv_full[st], c * 64, hk, j * BKV, b);
          } else {
            tma_load_3d(sK + st * C::kKBytes + c * (BKV * 128), &tm_k, &kv_full[st], hk * D + c * 64, j * BKV, b);
            tma_load_3d(sV + st * C::kKBytes + c * (BKV * 128), &tm_v, &kv_full[st], hk * D + c * 64, j * BKV, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_pv = make_idesc(kBQ, PAD ? ((p.d + 15) & ~15) : D, 1);     // P V output columns
      const int ksteps_qk = PAD ? ((p.d + 15) >> 4) : D / 16;                         // Q K^T contraction steps
      // the last K/V tile only multiplies the 16-key groups that hold valid keys (BEiT: 577 = 9 * 64 + 1 keys)
      auto valid16 = [&](int j) { return min(BKV, (p.sk - j * BKV + 15) & ~15); };
      auto issue_pv = [&](int j) {
        const int st = j & 1, kst = j % C::kStages;
        mbar_wait_suspend(&p_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t vbase = smem_u32(sV + kst * C::kKBytes);
        const int ksteps = valid16(j) >> 4;
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t dv = make_mnmajor_sw128_desc(vbase + kk * 16 * 128, BKV * 128);
          umma_bf16_ts(tmem_base + C::kTmemO, tmem_base + C::kTmemP + st * (BKV / 2) + kk * 8, dv, idesc_pv,
                       static_cast<uint32_t>(j > 0 || kk > 0));
        }
        umma_commit(&p_empty[st]);       // P_j consumed, O updated
        umma_commit(&kv_empty[kst]);     // K_j / V_j stage free
      };
      mbar_wait_suspend(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1, kst = j % C::kStages;
        mbar_wait_suspend(&kv_full[kst], (j / C::kStages) & 1);
        mbar_wait_suspend(&s_empty[st], ((j >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t qbase = smem_u32(sQ), kbase = smem_u32(sK + kst * C::kKBytes);
        const uint32_t idesc_qk = make_idesc(kBQ, valid16(j), 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          if (PAD && kk >= ksteps_qk) break;
          const uint64_t da = make_kmajor_sw128_desc(qbase + (kk >> 2) * (kBQ * 128) + (kk & 3) * 32);
          const uint64_t db = make_kmajor_sw128_desc(kbase + (kk >> 2) * (BKV * 128) + (kk & 3) * 32);
          umma_bf16(tmem_base + C::kTmemS + st * BKV, da, db, idesc_qk, static_cast<uint32_t>(kk > 0));
        }
        umma_commit(&s_full[st]);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(n_tiles - 1);
    }
  } else {
    // ============================================================ softmax / correction / epilogue: thread == query row
    constexpr int HC = BKV / 2;                       // key columns of a tile owned by this warp
    constexpr int DH = D / 2;                         // O columns rescaled / written by this warp
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;                   // column half: warps 2-5 -> 0, warps 6-9 -> 1
    const int row = q * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    int qbase = 0;
    const bool cls_q = kRelpos && (qi == 0);
    if (kRelpos) {
      const int qc = min(qi, p.sq - 1);               // rows past the end reuse the last valid row's (in-range) index
      const int qp = max(qc - 1, 0);                  // the CLS query row uses an in-range dummy (its bias is a constant, see below)
      qbase = (qp / p.win + p.win - 1) * (2 * p.win - 1) + qp % p.win + p.win - 1;
    }
    const float sl2 = p.scale * kLog2e;
    const float c1 = kSoftcap ? p.scale / p.softcap : 0.f, c2 = kSoftcap ? p.softcap * kLog2e : 0.f;
    const int kstart = p.kv_start ? p.kv_start[b] : 0;
    if (kRelpos) {
      // this head's bias table (x log2 e) and the per-key index terms, loaded by the softmax warps only: the TMA producer and
      // the MMA issuer are already running (ncu on the first version: the strided table gather + CTA-wide barrier in front of
      // the first TMA cost ~4 us of a ~20 us CTA lifetime)
      const int tid = threadIdx.x - 64;
      constexpr int NT = 32 * kSoftWarps;
      if (p.head_major) {
        const float* src = p.relpos + static_cast<long long>(h) * nrel;
        if ((nrel & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          for (int i = tid; i < (nrel >> 2); i += NT) {
            float4 t = __ldg(reinterpret_cast<const float4*>(src) + i);
            t.x *= kLog2e; t.y *= kLog2e; t.z *= kLog2e; t.w *= kLog2e;
            reinterpret_cast<float4*>(sTab)[i] = t;
          }
        } else {
          for (int i = tid; i < nrel; i += NT) sTab[i] = __ldg(src + i) * kLog2e;
        }
      } else {
        for (int i = tid; i < nrel; i += NT) sTab[i] = p.relpos[static_cast<long long>(i) * p.hq + h] * kLog2e;
      }
      for (int kj = tid; kj < n_tiles_all * BKV; kj += NT)
        sKterm[kj] = (kj >= 1 && kj < p.sk) ? 4 * (((kj - 1) / p.win) * (2 * p.win - 1) + (kj - 1) % p.win) : 0;      // byte offsets
      asm volatile("bar.sync 5, 256;" ::: "memory");      // softmax warps only (ids 1-4 are the pair barriers)
    }
    // the two warps of a quadrant meet on named barrier 1 + q (64 threads)
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const int k0 = j * BKV + ch * HC;               // first key of this warp's columns
      mbar_wait_suspend(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      float s[HC];
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + C::kTmemS + st * BKV + ch * HC + lane_addr, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);             // this warp's half of S_j is in registers
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
          const int cls_off = 4 * (nrel - 3), qoff = 4 * qbase;
#pragma unroll
          for (int i4 = 0; i4 < HC / 4; ++i4) {
            const int4 kt = kt4[i4];                          // warp-wide broadcast
            const int i0 = cls_q ? cls_off : qoff - kt.x, i1 = cls_q ? cls_off : qoff - kt.y;
            const int i2 = cls_q ? cls_off : qoff - kt.z, i3 = cls_q ? cls_off : qoff - kt.w;
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
        if (k0 == 0) s[0] = fmaf(raw0, sl2, sTab[cls_q ? nrel - 1 : nrel - 2]);      // CLS key column
      } else if (kSoftcap) {
        float u2max = 0.f;
#pragma unroll
        for (int i = 0; i < HC; ++i) { const float u = s[i] * c1; u2max = fmaxf(u2max, u * u); }
        if (!__any_sync(0xffffffffu, u2max >= 0.1225f)) {
          // cap * tanh(u) * log2e with a degree-9 odd polynomial (exact to fp32 rounding for |u| < 0.35)
#pragma unroll
          for (int i = 0; i < HC; ++i) {
            const float u = s[i] * c1, u2 = u * u;
            float pl = 62.f / 2835.f;
            pl = fmaf(pl, u2, -17.f / 315.f);
            pl = fmaf(pl, u2, 2.f / 15.f);
            pl = fmaf(pl, u2, -1.f / 3.f);
            pl = fmaf(pl, u2, 1.f);
            s[i] = c2 * u * pl;
          }
        } else {                                  // rare: a large score somewhere in this warp's rows -> libm tanh
#pragma unroll
          for (int i = 0; i < HC; ++i) s[i] = c2 * tanhf(s[i] * c1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < HC; ++i) s[i] *= sl2;
      }
      if (kCausal || k0 + HC > p.sk || kstart > k0) {
#pragma unroll
        for (int i = 0; i < HC; ++i) {
          const int kj = k0 + i;
          if (kj >= p.sk || kj < kstart || (kCausal && kj > max(qi + causal_off, p.prefix - 1))) s[i] = -INFINITY;
        }
      }
      float mj = -INFINITY;
#pragma unroll
      for (int i = 0; i < HC; ++i) mj = fmaxf(mj, s[i]);
      // ---- row max over the whole tile: exchange with the partner warp (parity-double-buffered, one barrier per tile)
      sXch[(st * 2 + ch) * kBQ + row] = mj;
      pair_sync();
      mj = fmaxf(mj, sXch[(st * 2 + (ch ^ 1)) * kBQ + row]);
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
      // P buffer st was read by P_{j-2} V_{j-2}; a rescale additionally needs P_{j-1} V_{j-1} (the last writer of O) done
      if (j >= 2) mbar_wait_suspend(&p_empty[st], ((j - 2) >> 1) & 1);
      if (__any_sync(0xffffffffu, rescale)) {                // identical decision in both warps of the pair (same rows, same max)
        mbar_wait_suspend(&p_empty[st ^ 1], ((j - 1) >> 1) & 1);
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
      tc_fence_after();
      tmem_st16(tmem_base + C::kTmemP + st * (BKV / 2) + ch * (HC / 2) + lane_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    }
    // ---- epilogue: total row sum (both column halves), wait for the last P V, O / l -> bf16 -> global
    pair_sync();                                             // the partner has finished reading the exchange buffers
    sXch[ch * kBQ + row] = l_run;
    pair_sync();
    l_run += sXch[(ch ^ 1) * kBQ + row];
    // training: keep log2-sum-exp of the row (scores are in the log2 domain; m_run is the -- possibly stale -- max l_run refers to)
    if (p.lse && ch == 0 && qi < p.sq)
      p.lse[(static_cast<long long>(b) * p.hq + h) * p.lse_stride + qi] = (l_run > 0.f) ? m_run + log2f(l_run) : -INFINITY;
    const int last = n_tiles - 1;
    mbar_wait_suspend(&p_empty[last & 1], (last >> 1) & 1);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __nv_bfloat16* og = p.out + b * p.o_bs + static_cast<long long>(qi) * p.o_ss + static_cast<long long>(h) * (PAD ? p.d : D);
#pragma unroll 1
    for (int c0 = ch * DH; c0 < (ch + 1) * DH; c0 += 32) {
      if (PAD && c0 >= p.d) break;
      uint32_t r[32];
      tmem_ld32(tmem_base + C::kTmemO + c0 + lane_addr, r);
      if (qi < p.sq) {
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          if (PAD && c0 + 8 * v8 >= p.d) break;
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[8 * v8 + 0]) * inv, __uint_as_float(r[8 * v8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(r[8 * v8 + 2]) * inv, __uint_as_float(r[8 * v8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(r[8 * v8 + 4]) * inv, __uint_as_float(r[8 * v8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(r[8 * v8 + 6]) * inv, __uint_as_float(r[8 * v8 + 7]) * inv);
          *reinterpret_cast<uint4*>(og + c0 + 8 * v8) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols, 1>(tmem_base);
  }
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

// 3-D view {cols (head-major features), tokens, batch} of a strided bf16 activation; box = 64 columns x box_rows tokens
static int encode_tokens(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t tokens, uint64_t batch, uint64_t token_stride,
                         uint64_t batch_stride, uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[3] = {cols, tokens, batch};
  cuuint64_t strides[2] = {token_stride * 2, batch_stride * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

// 4-D view {d, head, tokens, batch} of the same activation: columns past the real head dimension are out of bounds -> zero fill
static int encode_tokens_padded(CUtensorMap* tm, const void* base, uint64_t d, uint64_t heads, uint64_t tokens, uint64_t batch,
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

template <int D, int MODE, bool PAD = false>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Params& p, int batch, cudaStream_t st) {
  using C = Cfg<D>;
  const int nrel = (MODE & 1) ? (2 * p.win - 1) * (2 * p.win - 1) + 3 : 0;
  const int n_tiles_all = (p.sk + kBKV - 1) / kBKV;
  const size_t smem = C::kSmemBytes + 4 * kBQ * 4 + ((MODE & 1) ? static_cast<size_t>(n_tiles_all) * kBKV * 4 : 0) + static_cast<size_t>(nrel) * 4;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(svla_flash_attn_tc_kernel<D, MODE, PAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      svla_set_error("svla_attention(tcgen05): smem opt-in %zu failed: %s", smem, cudaGetErrorString(e));
      return -2;
    }
    configured = smem;
  }
  dim3 grid((p.sq + kBQ - 1) / kBQ, p.hq, batch);
  svla_flash_attn_tc_kernel<D, MODE, PAD><<<grid, kThreads, smem, st>>>(tq, tk, tv, p);
  SVLA_LAUNCH_CHECK("svla_flash_attn_tc");
  return 0;
}

}  // namespace svla_attn_tc

// Returns 1 if the tcgen05 kernel does not cover the problem (the caller then uses the mma.sync kernel), 0 on success,
// < 0 on error.  Covered: d in {64, 256} with the feature sets the model uses (BEiT: rel-pos or plain; Gemma2: soft-cap
// with / without causal mask), 16-byte aligned bases and strides, output rows 16-byte aligned.
int svla_attention_tc_try(const SvlaAttnArgs* a, void* stream) {
  using namespace svla_attn_tc;
  const int mode = (a->relpos_table ? 1 : 0) | (a->softcap > 0.f ? 2 : 0) | (a->causal ? 4 : 0);
  // head dims that are not a tile width (SigLIP 72, router 32): zero-padded by the TMA unit inside a 128- / 64-wide tile
  const bool padded = mode == 0 && a->d != 64 && a->d <= 128 && (a->d % 8) == 0 && !a->kv_start;
  if (!padded && !((a->d == 64 && (mode == 1 || mode == 0)) || (a->d == 256 && (mode == 2 || mode == 6)))) return 1;
  if ((a->q_ss % 8) || (a->k_ss % 8) || (a->v_ss % 8) || (a->q_bs % 8) || (a->k_bs % 8) || (a->v_bs % 8) || (a->o_ss % 8) || (a->o_bs % 8)) return 1;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->out)) & 15) return 1;
  if (a->causal && a->sk < a->sq) return 1;
  Params p{};
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.o_bs = a->o_bs; p.o_ss = a->o_ss;
  p.hq = a->hq; p.hkv = a->hkv; p.sq = a->sq; p.sk = a->sk; p.d = a->d;
  p.scale = a->scale; p.softcap = a->softcap; p.causal = a->causal; p.relpos = a->relpos_table; p.win = a->relpos_win; p.head_major = a->relpos_head_major; p.kv_start = a->kv_start;
  p.prefix = a->causal ? a->causal_prefix : 0;
  p.lse = a->lse;
  p.lse_stride = a->lse_stride;
  // a batch stride of 0 is not expressible in a tensor map; batch == 1 problems get a dummy stride
  const uint64_t nb = static_cast<uint64_t>(a->batch);
  auto bstride = [&](int64_t bs, int64_t ss, int s) { return static_cast<uint64_t>(nb > 1 ? bs : ss * s); };
  if (nb > 1 && (a->q_bs <= 0 || a->k_bs <= 0 || a->v_bs <= 0)) return 1;
  CUtensorMap tq, tk, tv;
  if (padded) {
    const uint64_t dd = static_cast<uint64_t>(a->d);
    if (encode_tokens_padded(&tq, a->q, dd, a->hq, a->sq, nb, a->q_ss, bstride(a->q_bs, a->q_ss, a->sq), kBQ) != 0) return 1;
    if (encode_tokens_padded(&tk, a->k, dd, a->hkv, a->sk, nb, a->k_ss, bstride(a->k_bs, a->k_ss, a->sk), kBKV) != 0) return 1;
    if (encode_tokens_padded(&tv, a->v, dd, a->hkv, a->sk, nb, a->v_ss, bstride(a->v_bs, a->v_ss, a->sk), kBKV) != 0) return 1;
    cudaStream_t stp = static_cast<cudaStream_t>(stream);
    return a->d < 64 ? launch<64, 0, true>(tq, tk, tv, p, a->batch, stp) : launch<128, 0, true>(tq, tk, tv, p, a->batch, stp);
  }
  if (encode_tokens(&tq, a->q, static_cast<uint64_t>(a->hq) * a->d, a->sq, nb, a->q_ss, bstride(a->q_bs, a->q_ss, a->sq), kBQ) != 0) return 1;
  if (encode_tokens(&tk, a->k, static_cast<uint64_t>(a->hkv) * a->d, a->sk, nb, a->k_ss, bstride(a->k_bs, a->k_ss, a->sk), kBKV) != 0) return 1;
  if (encode_tokens(&tv, a->v, static_cast<uint64_t>(a->hkv) * a->d, a->sk, nb, a->v_ss, bstride(a->v_bs, a->v_ss, a->sk), kBKV) != 0) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->d == 64) return mode == 1 ? launch<64, 1>(tq, tk, tv, p, a->batch, st) : launch<64, 0>(tq, tk, tv, p, a->batch, st);
  return mode == 2 ? launch<256, 2>(tq, tk, tv, p, a->batch, st) : launch<256, 6>(tq, tk, tv, p, a->batch, st);
}
