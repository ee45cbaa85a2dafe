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

// number of interior edges [lo, hi) of one axis whose rounding boundary is <= the angle of the vector (b, a)
__device__ __forceinline__ int angle_count(double a, double b, const DD4* __restrict__ tab, int lo, int hi) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (angle_ge(a, b, tab[mid])) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
svla_tok_encode_kernel(const double* __restrict__ actions, const double* __restrict__ edges, const double* __restrict__ trig,
                       TokGrid g, int* __restrict__ ids, long long n, double amin, double amax, int spherical, int phi_nonpos,
                       int phi_neg) {
  __shared__ double se[kMaxEdges];
  __shared__ DD4 st[128];                 // interior theta edges then interior phi edges (<= 63 + 63)
  int total = g.off[5] + g.nb[5] + 1;
  for (int i = threadIdx.x; i < total; i += blockDim.x) se[i] = edges[i];
  const int n_ti = g.nb[0] - 1, n_pi = g.nb[1] - 1;
  const bool exact = trig != nullptr && spherical;
  if (exact)
    for (int i = threadIdx.x; i < n_ti + n_pi; i += blockDim.x) st[i] = DD4{trig[4 * i], trig[4 * i + 1], trig[4 * i + 2], trig[4 * i + 3]};
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const double* a = actions + i * 7;
    const double x = clampd(a[0], amin, amax), y = clampd(a[1], amin, amax), z = clampd(a[2], amin, amax);
    const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), zz = __dmul_rn(z, z);
    const double sxy = __dadd_rn(xx, yy);
    // use_spherical=False (model/action_tokenizer.py:112-113) bins the clipped Cartesian components directly
    const double r = spherical ? sqrt(__dadd_rn(sxy, zz)) : z;
    // translation uses the interior edges e[1:-1]
    int dt, dp;
    if (exact) {
      const double rho = sqrt(sxy);
      // degenerate vectors (a zero component decides the quadrant through the sign of zero): library atan2, exact special values
      if (rho == 0.0 && z == 0.0) dt = digitize_right_open(atan2(rho, z), se + g.off[0] + 1, n_ti);
      else dt = angle_count(rho, z, st, 0, n_ti);                         // theta in [0, pi], every interior edge in (0, pi)
      if (y == 0.0 || x != x || y != y) dp = digitize_right_open(atan2(y, x), se + g.off[1] + 1, n_pi);
      else if (y > 0.0) dp = angle_count(y, x, st + n_ti, phi_nonpos, n_pi);   // phi in (0, pi): every edge <= 0 is below it
      else dp = angle_count(y, x, st + n_ti, 0, phi_neg);                      // phi in (-pi, 0): only negative edges can be
    } else {
      const double theta = spherical ? atan2(sqrt(sxy), z) : x;
      const double phi = spherical ? atan2(y, x) : y;
      dt = digitize_right_open(theta, se + g.off[0] + 1, g.nb[0] - 1);
      dp = digitize_right_open(phi, se + g.off[1] + 1, g.nb[1] - 1);
    }
    const int dr = digitize_right_open(r, se + g.off[2] + 1, g.nb[2] - 1);
    const int tid = dt * (g.nb[1] * g.nb[2]) + dp * g.nb[2] + dr;
    int d3[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v = clampd(a[3 + c], amin, amax);
      int d = digitize_right_open(v, se + g.off[3 + c], g.nb[3 + c] + 1) - 1;
      d3[c] = min(max(d, 0), g.nb[3 + c] - 1);
    }
    const int n_trans = g.nb[0] * g.nb[1] * g.nb[2];
    const int n_rot = g.nb[3] * g.nb[4] * g.nb[5];
    const int rid = d3[0] * (g.nb[4] * g.nb[5]) + d3[1] * g.nb[5] + d3[2] + n_trans;
    const int gid = (clampd(a[6], amin, amax) >= 0.5 ? 1 : 0) + n_trans + n_rot;
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
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long i0 = tile * kTokTile;
    const int cnt = static_cast<int>(n - i0 < kTokTile ? n - i0 : kTokTile);
    __syncthreads();                                   // previous tile's actions are out; se[] is loaded
    for (int k = threadIdx.x; k < cnt * 3; k += kTokTile) sid[k] = ids[i0 * 3 + k];
    __syncthreads();
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
  svla_tok_encode_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(actions, edges, edge_trig, g, ids, n, min_action,
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
