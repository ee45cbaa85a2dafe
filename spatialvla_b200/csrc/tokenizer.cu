// M8: SpatialActionTokenizer adaptive-grid lookup and inverse in FP64.
// Reference: model/action_tokenizer.py:105-137 (translation, spherical), :177-202 (rotation), :227-243 (gripper),
// :305-333 (SpatialActionTokenizer.__call__ / decode_token_ids_to_actions).
// HBM-bound: encode reads 56 B and writes 12 B per action, decode reads 24 B and writes 56 B.
// IEEE add/mul/sqrt are issued through the _rn intrinsics so nvcc cannot contract them into FMAs: the sums of
// squares are then bit-identical to numpy's; only atan2 / sin / cos can differ from glibc in the last ulp.
#include "svla_common.cuh"

namespace {

constexpr int kMaxEdges = 6 * 65;

struct TokGrid {
  int nb[7];
  int off[6];
};

__device__ __forceinline__ int digitize_right_open(double v, const double* __restrict__ e, int n) {
  // number of edges e[0..n) that are <= v  (== np.digitize(v, e) for increasing e; NaN -> n like numpy's searchsorted)
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return (v != v) ? n : lo;
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// The DECODE kernel moves whole tiles of kTokTile actions between global and shared memory with unit-stride accesses (a warp
// instruction covers 256 contiguous bytes); the per-action records (3 ids in, 7 doubles out) are read and written in shared
// memory.  Its first version let every thread load / store its own record straight from global memory: every 8-byte access
// touched its own 32-byte sector (ncu: 7 store sectors per action instead of 1.75; L1 at 66 % / L2 at 54 % with DRAM at 41 %):
// 0.487 -> 0.336 ms on 16 M actions (3.8 TB/s).  The ENCODE kernel keeps per-thread records: it is issue-bound on FP64 library
// code (2 x atan2, 2 x sqrt, 6 binary searches: 806 thread-instructions per action, 78 % of issue slots busy at 35 % of the
// HBM peak), its 7-sector record loads hit L1, and the staged variant measured SLOWER (0.549 vs 0.455 ms on 16 M actions:
// extra staging instructions and three barriers per tile in an issue-bound loop).
constexpr int kTokTile = 256;

// ---- exact angular binning without atan2 -----------------------------------------------------------------------------------
// The reference bins theta = atan2(rho, z) and phi = atan2(y, x) with np.digitize, i.e. it asks "fl(atan2(a, b)) >= e" for an
// edge e.  For a correctly rounded atan2 that is "atan2(a, b) >= m", m = (pred(e) + e) / 2 the rounding boundary below e, and
//     atan2(a, b) >= m   <=>   sin(atan2(a, b) - m) >= 0   <=>   a cos m - b sin m >= 0        (|angle - m| < pi)
// a 2-term cross product.  cos m / sin m are tabulated per interior edge on the host with 200-bit arithmetic as double-double
// pairs (SpatialActionTokenizer._edge_trig).  Fast path: one rounded evaluation in double whose sign is certain unless it is
// within 8 roundings of zero; slow path (only then): error-free products and sums, 2^-104 relative -- the decision of a
// correctly rounded atan2, which is what glibc returns on every near-edge vector of the golden files (checked with mpmath).
// CUDA's own atan2 (2 ulp) differs from glibc within a few ulp of an edge; it is kept for degenerate vectors only.
// This also removes ~500 of the 806 FP64 instructions per action the atan2 version spent.
struct DD4 { double ch, cl, sh, sl; };

__device__ __forceinline__ bool angle_ge(double a, double b, const DD4 t) {
  // sign of a * cos(m) - b * sin(m)
  const double p1 = __dmul_rn(a, t.ch), p2 = __dmul_rn(b, t.sh);
  const double d = __dadd_rn(p1, -p2);
  const double bound = 1.8e-15 * __dadd_rn(fabs(p1), fabs(p2));      // > 8 * 2^-53 * (|p1| + |p2|)
  if (fabs(d) > bound) return d > 0.0;
  // error-free transformation: p1 + e1 = a*ch, p2 + e2 = b*sh exactly; s + e3 = p1 - p2 exactly
  const double e1 = __fma_rn(a, t.ch, -p1), e2 = __fma_rn(b, t.sh, -p2);
  const double s = d;
  const double bb = __dadd_rn(s, -p1);
  const double e3 = __dadd_rn(__dadd_rn(p1, -__dadd_rn(s, -bb)), __dadd_rn(-p2, -bb));
  const double lo = __dadd_rn(__dadd_rn(__dadd_rn(e1, -e2), e3), __dadd_rn(__dmul_rn(a, t.cl), -__dmul_rn(b, t.sl)));
  const double tot = __dadd_rn(s, lo);
  return tot >= 0.0;
}

// ---- table-driven binning (round 2) -----------------------------------------------------------------------------------------
// ncu on the binary-search version: 909 thread-instructions per action, 78 % of the issue slots busy at 0.32 of the HBM peak,
// FP64 pipe 9 %: the loop overhead of six data-dependent binary searches (IMAD / ISETP / BRA / BSSY) and the NaN-aware
// fmin/fmax clamps, not FP64 arithmetic, were the bound.  Each axis now has a small lookup table over uniform cells of a cheap
// fp32 key that is MONOTONE in the binned quantity; a table entry is a guaranteed LOWER bound of the bin count (it counts the
// edges below the cell BEFORE the key's cell, which absorbs every fp32 rounding of the key), and a forward scan with the exact
// test finishes the count -- one failing test in the common case.  The result is the exact np.digitize for every input:
//   * r / roll / pitch / yaw: key = the value itself, exact test e[lo] <= v in double;
//   * theta / phi: key = the "diamond angle" p(a, b) = sign(a) (1 - b / (|a| + |b|)) in [-2, 2] (monotone in atan2(a, b)), exact
//     test = sign of a cos m - b sin m: first in fp32, where it differs from the true value by < 4.5 * 2^-24 (|p1| + |p2|) (operand
//     roundings, sqrtf, products, difference); inside that band (a vector within ~1e-6 rad of an edge) the double / double-double
//     test above decides.
constexpr int kLutPlain = 256, kLutAng = 512;

__device__ __noinline__ bool angle_ge_slow(double a, double b, double ch, double cl, double sh, double sl) { return angle_ge(a, b, DD4{ch, cl, sh, sl}); }
__device__ __noinline__ bool angle_ge_slow_sqrt(double a2, double b, double ch, double cl, double sh, double sl) { return angle_ge(sqrt(a2), b, DD4{ch, cl, sh, sl}); }

// NaN-propagating clip (np.clip); fmin(fmax()) costs 17 instructions per value for its NaN rules
__device__ __forceinline__ double clipd(double v, double lo, double hi) { return v > hi ? hi : (v < lo ? lo : v); }

struct PlainAxis { float base, inv_w; };

__device__ __forceinline__ int digitize_lut(double v, const double* __restrict__ e, int n, const unsigned char* __restrict__ lut, PlainAxis ax) {
  const float vf = __double2float_rn(v);
  const int cell = min(max(__float2int_rd((vf - ax.base) * ax.inv_w), 0), kLutPlain - 1);      // NaN -> 0
  int lo = lut[cell];
  while (lo < n && e[lo] <= v) ++lo;
  return (v != v) ? n : lo;
}

// number of interior edges of one axis (all of [0, lo_min) are known to be below, none of [n, ..) can be) whose rounding boundary is <= the angle of the vector (b, a), a = a_arg (or
// sqrt(a_arg) with SQRT: the double square root is only taken when the fp32 screen cannot decide)
template <bool SQRT>
__device__ __forceinline__ int angle_count_lut(float af, float bf, double a_arg, double b, const DD4* __restrict__ tab, const float2* __restrict__ tabf,
                                               int lo_min, int n, const unsigned char* __restrict__ lut, float key_base, float key_inv_w) {
  float q = 1.f - __fdividef(bf, fabsf(af) + fabsf(bf));
  q = af < 0.f ? -q : q;
  const int cell = min(max(__float2int_rd((q - key_base) * key_inv_w), 0), kLutAng - 1);
  // only edges within two cells below the angle up to the first edge above it are tested, all inside [lo_min, n): the sign test
  // needs |angle - m| < pi, which holds for edges on the angle's side of zero
  int lo = max(static_cast<int>(lut[cell]), lo_min);
  while (lo < n) {
    const float2 t = tabf[lo];
    const float p1 = __fmul_rn(af, t.x), p2 = __fmul_rn(bf, t.y);
    const float d = __fadd_rn(p1, -p2);
    bool ge = d > 0.f;
    if (!(fabsf(d) > 6e-7f * (fabsf(p1) + fabsf(p2)))) {
      const DD4 w = tab[lo];
      ge = SQRT ? angle_ge_slow_sqrt(a_arg, b, w.ch, w.cl, w.sh, w.sl) : angle_ge_slow(a_arg, b, w.ch, w.cl, w.sh, w.sl);
    }
    if (!ge) break;
    ++lo;
  }
  return lo;
}

// One tile = kTokTile actions: the 7 doubles of every action are moved global -> registers -> shared memory with unit-stride
// 8-byte accesses (the loads of tile t + 1 are in flight while tile t is binned), then each thread bins one action from shared
// memory (stride-7 double reads are bank-conflict-free).
__global__ void __launch_bounds__(kTokTile)
svla_tok_encode_kernel(const double* __restrict__ actions, const double* __restrict__ edges, const double* __restrict__ trig,
                       TokGrid g, int* __restrict__ ids, long long n, double amin, double amax, int spherical, int phi_nonpos,
                       int phi_neg) {
  __shared__ double se[kMaxEdges];
  __shared__ DD4 st[128];                 // interior theta edges then interior phi edges (<= 63 + 63)
  __shared__ float2 stf[128];             // (cos m, sin m) rounded to float
  __shared__ double spm[128];             // diamond angle of every interior theta / phi edge
  __shared__ unsigned char lut_p[4][kLutPlain];      // r (interior edges), roll, pitch, yaw (all edges)
  __shared__ unsigned char lut_a[2][kLutAng];        // theta, phi
  __shared__ double sa[kTokTile * 7];
  int total = g.off[5] + g.nb[5] + 1;
  for (int i = threadIdx.x; i < total; i += blockDim.x) se[i] = edges[i];
  const int n_ti = g.nb[0] - 1, n_pi = g.nb[1] - 1;
  const bool exact = trig != nullptr && spherical;
  if (exact)
    for (int i = threadIdx.x; i < n_ti + n_pi; i += blockDim.x) {
      const DD4 t = DD4{trig[4 * i], trig[4 * i + 1], trig[4 * i + 2], trig[4 * i + 3]};
      st[i] = t;
      stf[i] = make_float2(__double2float_rn(t.ch), __double2float_rn(t.sh));
      const double q = 1.0 - t.ch / (fabs(t.sh) + fabs(t.ch));
      spm[i] = t.sh < 0.0 ? -q : q;
    }
  __syncthreads();
  // axis descriptors (uniform): first edge / edge count of the four plain axes, uniform cells over [first edge, last edge]
  const double* e_ax[4] = {se + g.off[2] + 1, se + g.off[3], se + g.off[4], se + g.off[5]};
  const int n_ax[4] = {g.nb[2] - 1, g.nb[3] + 1, g.nb[4] + 1, g.nb[5] + 1};
  PlainAxis pax[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double lo_d = n_ax[c] > 0 ? e_ax[c][0] : 0.0, hi_d = n_ax[c] > 0 ? e_ax[c][n_ax[c] - 1] : 1.0;
    const double w = hi_d > lo_d ? (hi_d - lo_d) / kLutPlain : 1.0;
    pax[c].base = __double2float_rn(lo_d);
    pax[c].inv_w = __double2float_rn(1.0 / w);
    // entry = #edges <= lower bound of the PREVIOUS cell (in terms of the fp32 key map actually used at run time)
    for (int k = threadIdx.x; k < kLutPlain; k += blockDim.x) {
      const double lower = static_cast<double>(pax[c].base) + (k - 1) / static_cast<double>(pax[c].inv_w);
      lut_p[c][k] = static_cast<unsigned char>(digitize_right_open(lower, e_ax[c], n_ax[c]));
    }
  }
  const float th_base = 0.f, th_inv = kLutAng / 2.f, ph_base = -2.f, ph_inv = kLutAng / 4.f;
  if (exact)
    for (int k = threadIdx.x; k < 2 * kLutAng; k += blockDim.x) {
      const int axis = k / kLutAng, cidx = k - axis * kLutAng;
      const double lower = axis == 0 ? (cidx - 1) * (2.0 / kLutAng) : -2.0 + (cidx - 1) * (4.0 / kLutAng);
      lut_a[axis][cidx] = static_cast<unsigned char>(digitize_right_open(lower, spm + (axis ? n_ti : 0), axis ? n_pi : n_ti));
    }
  const long long n_tiles = (n + kTokTile - 1) / kTokTile;
  const int nb1 = g.nb[1], nb2 = g.nb[2], nb3 = g.nb[3], nb4 = g.nb[4], nb5 = g.nb[5];
  const int n_trans = g.nb[0] * nb1 * nb2, n_rot = nb3 * nb4 * nb5;
  double pf[7];
  auto prefetch = [&](long long tile) {
    const long long base = tile * kTokTile * 7, lim = n * 7;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      const long long k = base + c * kTokTile + threadIdx.x;
      pf[c] = (tile < n_tiles && k < lim) ? actions[k] : 0.0;
    }
  };
  auto clip = [&](double v) { return clipd(v, amin, amax); };
  prefetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();                                   // previous tile binned; tables built
#pragma unroll
    for (int c = 0; c < 7; ++c) sa[c * kTokTile + threadIdx.x] = pf[c];
    __syncthreads();
    prefetch(tile + gridDim.x);
    const long long i = tile * kTokTile + threadIdx.x;
    if (i >= n) continue;
    const double* a = sa + threadIdx.x * 7;
    const double x = clip(a[0]), y = clip(a[1]), z = clip(a[2]);
    const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), zz = __dmul_rn(z, z);
    const double sxy = __dadd_rn(xx, yy);
    // use_spherical=False (model/action_tokenizer.py:112-113) bins the clipped Cartesian components directly
    const double r = spherical ? sqrt(__dadd_rn(sxy, zz)) : z;
    // translation uses the interior edges e[1:-1]
    int dt, dp;
    if (exact) {
      const float xf = __double2float_rn(x), yf = __double2float_rn(y), zf = __double2float_rn(z);
      // degenerate vectors (a zero component decides the quadrant through the sign of zero) and NaNs: library atan2, exact special values
      if ((sxy == 0.0 && z == 0.0) || sxy != sxy || z != z) dt = digitize_right_open(atan2(sqrt(sxy), z), se + g.off[0] + 1, n_ti);
      else dt = angle_count_lut<true>(sqrtf(__double2float_rn(sxy)), zf, sxy, z, st, stf, 0, n_ti, lut_a[0], th_base, th_inv);      // theta in [0, pi]
      if (y == 0.0 || x != x || y != y) dp = digitize_right_open(atan2(y, x), se + g.off[1] + 1, n_pi);
      else if (y > 0.0) dp = angle_count_lut<false>(yf, xf, y, x, st + n_ti, stf + n_ti, phi_nonpos, n_pi, lut_a[1], ph_base, ph_inv);   // phi in (0, pi): every edge <= 0 is below it
      else dp = angle_count_lut<false>(yf, xf, y, x, st + n_ti, stf + n_ti, 0, phi_neg, lut_a[1], ph_base, ph_inv);                      // phi in (-pi, 0): only negative edges can be
    } else {
      const double theta = spherical ? atan2(sqrt(sxy), z) : x;
      const double phi = spherical ? atan2(y, x) : y;
      dt = digitize_right_open(theta, se + g.off[0] + 1, g.nb[0] - 1);
      dp = digitize_right_open(phi, se + g.off[1] + 1, g.nb[1] - 1);
    }
    const int dr = digitize_lut(r, e_ax[0], n_ax[0], lut_p[0], pax[0]);
    const int tid = dt * (nb1 * nb2) + dp * nb2 + dr;
    const int d0 = min(max(digitize_lut(clip(a[3]), e_ax[1], n_ax[1], lut_p[1], pax[1]) - 1, 0), nb3 - 1);
    const int d1 = min(max(digitize_lut(clip(a[4]), e_ax[2], n_ax[2], lut_p[2], pax[2]) - 1, 0), nb4 - 1);
    const int d2 = min(max(digitize_lut(clip(a[5]), e_ax[3], n_ax[3], lut_p[3], pax[3]) - 1, 0), nb5 - 1);
    const int rid = d0 * (nb4 * nb5) + d1 * nb5 + d2 + n_trans;
    const int gid = (clip(a[6]) >= 0.5 ? 1 : 0) + n_trans + n_rot;
    ids[i * 3 + 0] = tid;
    ids[i * 3 + 1] = rid;
    ids[i * 3 + 2] = gid;
  }
}

__global__ void __launch_bounds__(kTokTile)
svla_tok_decode_kernel(const long long* __restrict__ ids, const double* __restrict__ edges, const double* __restrict__ ctrig,
                       TokGrid g, long long begin, double* __restrict__ actions, long long n, int spherical) {
  __shared__ double se[kMaxEdges];
  __shared__ double sc[2 * 128];          // (sin, cos) of the theta bin centres, then of the phi bin centres
  __shared__ double so[kTokTile * 7];
  __shared__ long long sid[kTokTile * 3];
  int total = g.off[5] + g.nb[5] + 1;
  for (int i = threadIdx.x; i < total; i += blockDim.x) se[i] = edges[i];
  if (ctrig)
    for (int i = threadIdx.x; i < 2 * (g.nb[0] + g.nb[1]); i += blockDim.x) sc[i] = ctrig[i];
  const long long n_trans = static_cast<long long>(g.nb[0]) * g.nb[1] * g.nb[2];
  const long long n_rot = static_cast<long long>(g.nb[3]) * g.nb[4] * g.nb[5];
  const long long n_tiles = (n + kTokTile - 1) / kTokTile;
  long long pf[3];
  auto prefetch = [&](long long tile) {                 // the ids of the NEXT tile are in flight while this tile is decoded
    const long long base = tile * kTokTile * 3, lim = n * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long k = base + c * kTokTile + threadIdx.x;
      pf[c] = (tile < n_tiles && k < lim) ? ids[k] : 0;
    }
  };
  prefetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long i0 = tile * kTokTile;
    const int cnt = static_cast<int>(n - i0 < kTokTile ? n - i0 : kTokTile);
    __syncthreads();                                   // previous tile's actions are out; se[] is loaded
#pragma unroll
    for (int c = 0; c < 3; ++c) sid[c * kTokTile + threadIdx.x] = pf[c];
    __syncthreads();
    prefetch(tile + gridDim.x);
    if (threadIdx.x < cnt) {
      // clip each id into its sub-range first (model/action_tokenizer.py:126,195,240)
      long long t = sid[threadIdx.x * 3 + 0] - begin;
      t = t < 0 ? 0 : (t > n_trans - 1 ? n_trans - 1 : t);
      const int ti = static_cast<int>(t);
      const int np_ = g.nb[1] * g.nb[2];
      const int a = ti / np_, b = (ti % np_) / g.nb[2], c = ti % g.nb[2];
      const double th = 0.5 * __dadd_rn(se[g.off[0] + a], se[g.off[0] + a + 1]);
      const double ph = 0.5 * __dadd_rn(se[g.off[1] + b], se[g.off[1] + b + 1]);
      const double rr = 0.5 * __dadd_rn(se[g.off[2] + c], se[g.off[2] + c + 1]);
      double x = th, y = ph, z = rr;
      if (spherical) {
        double st, ct, sp, cp;
        if (ctrig) {
          // the inverse only ever evaluates sin / cos at the bin centres: tabulated once on the host with the reference's own libm
          st = sc[2 * a]; ct = sc[2 * a + 1];
          sp = sc[2 * (g.nb[0] + b)]; cp = sc[2 * (g.nb[0] + b) + 1];
        } else {
          sincos(th, &st, &ct);
          sincos(ph, &sp, &cp);
        }
        x = __dmul_rn(__dmul_rn(rr, st), cp);
        y = __dmul_rn(__dmul_rn(rr, st), sp);
        z = __dmul_rn(rr, ct);
      }
      double* o = so + threadIdx.x * 7;
      o[0] = clampd(x, -1.0, 1.0);
      o[1] = clampd(y, -1.0, 1.0);
      o[2] = clampd(z, -1.0, 1.0);
      long long r = sid[threadIdx.x * 3 + 1] - begin;
      r = r < n_trans ? n_trans : (r > n_trans + n_rot - 1 ? n_trans + n_rot - 1 : r);
      const int ri = static_cast<int>(r - n_trans);
      const int nq = g.nb[4] * g.nb[5];
      const int r0 = ri / nq, r1 = (ri % nq) / g.nb[5], r2 = ri % g.nb[5];
      o[3] = 0.5 * __dadd_rn(se[g.off[3] + r0], se[g.off[3] + r0 + 1]);
      o[4] = 0.5 * __dadd_rn(se[g.off[4] + r1], se[g.off[4] + r1 + 1]);
      o[5] = 0.5 * __dadd_rn(se[g.off[5] + r2], se[g.off[5] + r2 + 1]);
      long long gi = sid[threadIdx.x * 3 + 2] - begin;
      const long long glo = n_trans + n_rot, ghi = n_trans + n_rot + g.nb[6] - 1;
      gi = gi < glo ? glo : (gi > ghi ? ghi : gi);
      o[6] = (gi - glo == 0) ? 0.0 : 1.0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < cnt * 7; k += kTokTile) actions[i0 * 7 + k] = so[k];
  }
}

int make_grid(const int32_t* nbins_host, TokGrid& g) {
  int off = 0;
  for (int i = 0; i < 7; ++i) g.nb[i] = nbins_host[i];
  for (int i = 0; i < 6; ++i) {
    if (g.nb[i] < 1 || g.nb[i] > 64) return -1;
    g.off[i] = off;
    off += g.nb[i] + 1;
  }
  return off;
}

unsigned grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148LL * 8;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// nbins is read on the HOST in every variant (it is 7 ints of configuration, never per-sample data).
extern "C" int svla_tok_encode(const double* actions, const double* edges, const int32_t* nbins, int32_t* ids, int64_t n,
                               double min_action, double max_action, int use_spherical, const double* edge_trig,
                               int phi_nonpos, int phi_neg, void* stream) {
  SVLA_REQUIRE(edges && nbins && (n == 0 || (actions && ids)), "svla_tok_encode: null pointer");
  SVLA_REQUIRE(n >= 0, "svla_tok_encode: negative n");
  if (n == 0) return 0;
  TokGrid g;
  SVLA_REQUIRE(make_grid(nbins, g) > 0, "svla_tok_encode: bins per axis must be in [1, 64]");
  SVLA_REQUIRE(!edge_trig || (phi_neg >= 0 && phi_neg <= phi_nonpos && phi_nonpos <= g.nb[1] - 1),
               "svla_tok_encode: phi_neg=%d <= phi_nonpos=%d <= %d interior phi edges expected", phi_neg, phi_nonpos, g.nb[1] - 1);
  svla_tok_encode_kernel<<<grid_for(n), kTokTile, 0, static_cast<cudaStream_t>(stream)>>>(actions, edges, edge_trig, g, ids, n, min_action,
                                                                                     max_action, use_spherical, phi_nonpos, phi_neg);
  SVLA_LAUNCH_CHECK("svla_tok_encode");
  return 0;
}

extern "C" int svla_tok_decode(const int64_t* ids, const double* edges, const int32_t* nbins, int64_t begin, double* actions,
                               int64_t n, int use_spherical, const double* center_trig, void* stream) {
  SVLA_REQUIRE(edges && nbins && (n == 0 || (actions && ids)), "svla_tok_decode: null pointer");
  SVLA_REQUIRE(n >= 0, "svla_tok_decode: negative n");
  if (n == 0) return 0;
  TokGrid g;
  SVLA_REQUIRE(make_grid(nbins, g) > 0, "svla_tok_decode: bins per axis must be in [1, 64]");
  svla_tok_decode_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), edges, center_trig, g, begin, actions, n, use_spherical);
  SVLA_LAUNCH_CHECK("svla_tok_decode");
  return 0;
}

#define SVLA_CUDA_OK(call)                                                        \
  do {                                                                            \
    cudaError_t _e = (call);                                                      \
    if (_e != cudaSuccess) {                                                      \
      svla_set_error("%s failed: %s", #call, cudaGetErrorString(_e));             \
      rc = -2;                                                                    \
      goto done;                                                                  \
    }                                                                             \
  } while (0)

extern "C" int svla_tok_encode_host(const double* actions_host, const double* edges_host, const int32_t* nbins_host,
                                    int32_t* ids_host, int64_t n, double min_action, double max_action,
                                    int use_spherical, const double* edge_trig_host, int phi_nonpos, int phi_neg) {
  if (n == 0) return 0;
  TokGrid g;
  const int ne = make_grid(nbins_host, g);
  SVLA_REQUIRE(ne > 0, "svla_tok_encode_host: bins per axis must be in [1, 64]");
  const int nt = edge_trig_host ? 4 * (g.nb[0] - 1 + g.nb[1] - 1) : 0;
  int rc = 0;
  double *d_a = nullptr, *d_e = nullptr;
  int32_t* d_i = nullptr;
  SVLA_CUDA_OK(cudaMalloc(&d_a, sizeof(double) * 7 * n));
  SVLA_CUDA_OK(cudaMalloc(&d_e, sizeof(double) * (ne + nt)));
  SVLA_CUDA_OK(cudaMalloc(&d_i, sizeof(int32_t) * 3 * n));
  SVLA_CUDA_OK(cudaMemcpy(d_a, actions_host, sizeof(double) * 7 * n, cudaMemcpyHostToDevice));
  SVLA_CUDA_OK(cudaMemcpy(d_e, edges_host, sizeof(double) * ne, cudaMemcpyHostToDevice));
  if (nt) SVLA_CUDA_OK(cudaMemcpy(d_e + ne, edge_trig_host, sizeof(double) * nt, cudaMemcpyHostToDevice));
  rc = svla_tok_encode(d_a, d_e, nbins_host, d_i, n, min_action, max_action, use_spherical, nt ? d_e + ne : nullptr, phi_nonpos,
                       phi_neg, nullptr);
  if (rc == 0) SVLA_CUDA_OK(cudaMemcpy(ids_host, d_i, sizeof(int32_t) * 3 * n, cudaMemcpyDeviceToHost));
done:
  cudaFree(d_a); cudaFree(d_e); cudaFree(d_i);
  return rc;
}

extern "C" int svla_tok_decode_host(const int64_t* ids_host, const double* edges_host, const int32_t* nbins_host, int64_t begin,
                                    double* actions_host, int64_t n, int use_spherical, const double* center_trig_host) {
  if (n == 0) return 0;
  TokGrid g;
  const int ne = make_grid(nbins_host, g);
  SVLA_REQUIRE(ne > 0, "svla_tok_decode_host: bins per axis must be in [1, 64]");
  const int nt = center_trig_host ? 2 * (g.nb[0] + g.nb[1]) : 0;
  int rc = 0;
  double *d_a = nullptr, *d_e = nullptr;
  int64_t* d_i = nullptr;
  SVLA_CUDA_OK(cudaMalloc(&d_a, sizeof(double) * 7 * n));
  SVLA_CUDA_OK(cudaMalloc(&d_e, sizeof(double) * (ne + nt)));
  SVLA_CUDA_OK(cudaMalloc(&d_i, sizeof(int64_t) * 3 * n));
  SVLA_CUDA_OK(cudaMemcpy(d_i, ids_host, sizeof(int64_t) * 3 * n, cudaMemcpyHostToDevice));
  SVLA_CUDA_OK(cudaMemcpy(d_e, edges_host, sizeof(double) * ne, cudaMemcpyHostToDevice));
  if (nt) SVLA_CUDA_OK(cudaMemcpy(d_e + ne, center_trig_host, sizeof(double) * nt, cudaMemcpyHostToDevice));
  rc = svla_tok_decode(d_i, d_e, nbins_host, begin, d_a, n, use_spherical, nt ? d_e + ne : nullptr, nullptr);
  if (rc == 0) SVLA_CUDA_OK(cudaMemcpy(actions_host, d_a, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost));
done:
  cudaFree(d_a); cudaFree(d_e); cudaFree(d_i);
  return rc;
}
