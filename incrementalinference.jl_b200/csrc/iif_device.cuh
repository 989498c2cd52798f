// iif_device.cuh — device-side building blocks shared by the convolution and product kernels.
//
// Everything here is FP64 (the reference path is Float64 end to end) and written for one CTA
// of IIF_NT threads working on one belief (N <= IIF_MAX_POINTS particles resident in
// shared memory).  Reference citations are to IncrementalInference.jl v0.35.6.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/iifb200.h"

// CTAs are launched with 128..512 threads (the host picks per wave: small CTAs for wide, throughput-bound
// waves, large CTAs for narrow, latency-bound ones); device code reads the size from blockDim.
#define IIF_MAX_THREADS 512
#define IIF_MAX_WARPS (IIF_MAX_THREADS / 32)
#define IIF_NT ((int)blockDim.x)
#define IIF_NW ((int)(blockDim.x >> 5))
#define IIF_PI 3.14159265358979323846
#define IIF_TWO_PI 6.28318530717958647692
#define IIF_RED_KMAX 8
#define IIF_RED_DOUBLES (2 * IIF_MAX_WARPS * IIF_RED_KMAX)

// random-stream ids: Philox4x32-10 counter word 1 (same constants as the test oracle)
enum {
  IIF_RS_MEAS = 1,
  IIF_RS_MIXLABEL = 2,
  IIF_RS_LABEL = 3,
  IIF_RS_INFLATE = 4,
  IIF_RS_ANYN = 5,
  IIF_RS_GIBBS_U = 6,
  IIF_RS_GIBBS_N = 7,
  IIF_RS_OLDPAD = 8
};

// ---- ball-tree structure of a balanced median-split tree over N points.  It depends on N
// only (left child gets ceil(n/2) points), so the host builds it once per N and caches it.
struct TreeStruct {
  int32_t L;               // levels below the root: floor(log2(N) + 1)  (KDE Nlevels)
  int32_t nn;              // total entries over all level lists
  const int16_t* lev_off;  // L+2 offsets; level l list = [lev_off[l], lev_off[l+1])
  const int16_t* lo;       // node range [lo, hi] in the per-density permutation
  const int16_t* hi;
  const int16_t* child;    // position (within level l+1) of the node's first child
  const int16_t* node_at;  // L x N: node_at[l*N + pos] = level-l list entry holding position pos
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG: counter = (idx, stream, call, tag), key = seed
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void rs_pair(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx,
                                        double& ua, double& ub) {
  uint32_t o[4];
  philox4x32(idx, stream, call, 0x1F1B200u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  uint64_t a = ((uint64_t)o[1] << 32) | o[0];
  uint64_t b = ((uint64_t)o[3] << 32) | o[2];
  ua = (double)(a >> 11) * (1.0 / 9007199254740992.0);
  ub = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double rs_uniform(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx) {
  double a, b;
  rs_pair(seed, call, stream, idx, a, b);
  return a;
}

// Box-Muller on the two uniforms of one Philox block
__device__ __forceinline__ double rs_normal(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx) {
  double a, b;
  rs_pair(seed, call, stream, idx, a, b);
  return sqrt(-2.0 * log(1.0 - a)) * cos(IIF_TWO_PI * b);
}

// ------------------------------------------------------------------------------------------
// Manifold helpers: coordinates of TranslationGroup(1) x RealCircleGroup products
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double wrap_pi(double a) {  // Manifolds.sym_rem: [-pi, pi)
  double r = a - IIF_TWO_PI * floor((a + IIF_PI) / IIF_TWO_PI);
  if (r >= IIF_PI) r -= IIF_TWO_PI;
  if (r < -IIF_PI) r += IIF_TWO_PI;
  return r;
}
__device__ __forceinline__ bool is_circ(int32_t mask, int c) { return (mask >> c) & 1; }
__device__ __forceinline__ double mdiff(double a, double b, bool circ) { return circ ? wrap_pi(a - b) : a - b; }
__device__ __forceinline__ double madd(double a, double t, bool circ) { return circ ? wrap_pi(a + t) : a + t; }

// ------------------------------------------------------------------------------------------
// Warp / block reductions.  `red` is shared scratch of IIF_RED_DOUBLES doubles split in two
// parity halves; `parity` alternates so that one __syncthreads per reduction suffices (the two
// halves never overlap whatever K is).  All threads get the result, summed in a fixed order
// (deterministic run to run).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* red, int& parity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* buf = red + parity * (IIF_MAX_WARPS * IIF_RED_KMAX);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) buf[warp * K + k] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0;
    for (int w = 0; w < IIF_NW; ++w) s += buf[w * K + k];
    v[k] = s;
  }
  parity ^= 1;
}
__device__ __forceinline__ double block_sum1(double x, double* red, int& parity) {
  double v[1] = {x};
  block_sum<1>(v, red, parity);
  return v[0];
}
__device__ __forceinline__ double block_min1(double x, double* red, int& parity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* buf = red + parity * (IIF_MAX_WARPS * IIF_RED_KMAX);
  double s = warp_min(x);
  if (lane == 0) buf[warp] = s;
  __syncthreads();
  double m = buf[0];
  for (int w = 1; w < IIF_NW; ++w) m = fmin(m, buf[w]);
  parity ^= 1;
  return m;
}

// ------------------------------------------------------------------------------------------
// exp(x) for x <= 0 in FP64 (the only use on this path: Gaussian kernel weights).  Cody-Waite
// range reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor/Horner (truncation 4e-18) with
// the coefficients taken straight from the constant bank (no 64-bit immediates to materialise),
// two-step power-of-two scaling so results down to the denormal range stay correct.
// ------------------------------------------------------------------------------------------
__constant__ double IIF_EXPC[16] = {
    1.6059043836821613e-10, 2.08767569878681e-09,  2.505210838544172e-08,  2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05,  1.984126984126984e-04,  1.388888888888889e-03,
    8.333333333333333e-03,  4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
    1.4426950408889634074,  -6.93147180369123816490e-01, -1.90821492927058770002e-10, 6755399441055744.0};

__device__ __forceinline__ double exp_neg(double x) {
  double t = fma(x, IIF_EXPC[12], IIF_EXPC[15]);
  const int k = __double2loint(t);
  t -= IIF_EXPC[15];
  double r = fma(t, IIF_EXPC[13], x);
  r = fma(t, IIF_EXPC[14], r);
  double p = IIF_EXPC[0];
#pragma unroll
  for (int c = 1; c < 12; ++c) p = fma(p, r, IIF_EXPC[c]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  // 2^k in two normal-range factors; k clamped so that anything below the denormal range is 0
  const int kc = max(k, -1100), k1 = kc >> 1, k2 = kc - k1;
  const double s1 = __hiloint2double((k1 + 1023) << 20, 0);
  const double s2 = __hiloint2double((k2 + 1023) << 20, 0);
  return (p * s1) * s2;
}

// Four exponentials in lockstep: every polynomial step is issued for all four arguments before the
// next one, so the four FP64 dependency chains interleave and each coefficient is fetched once.
__device__ __forceinline__ void exp_neg4(const double (&x)[4], double (&out)[4]) {
  const double L2E = IIF_EXPC[12], LN2H = IIF_EXPC[13], LN2L = IIF_EXPC[14], MAGIC = IIF_EXPC[15];
  double t[4], r[4], p[4];
  int k[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) t[u] = fma(x[u], L2E, MAGIC);
#pragma unroll
  for (int u = 0; u < 4; ++u) { k[u] = __double2loint(t[u]); t[u] -= MAGIC; }
#pragma unroll
  for (int u = 0; u < 4; ++u) r[u] = fma(t[u], LN2H, x[u]);
#pragma unroll
  for (int u = 0; u < 4; ++u) r[u] = fma(t[u], LN2L, r[u]);
  {
    const double c0 = IIF_EXPC[0], c1 = IIF_EXPC[1];
#pragma unroll
    for (int u = 0; u < 4; ++u) p[u] = fma(c0, r[u], c1);
  }
#pragma unroll
  for (int c = 2; c < 12; ++c) {
    const double cc = IIF_EXPC[c];
#pragma unroll
    for (int u = 0; u < 4; ++u) p[u] = fma(p[u], r[u], cc);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) p[u] = fma(p[u], r[u], 1.0);
#pragma unroll
  for (int u = 0; u < 4; ++u) p[u] = fma(p[u], r[u], 1.0);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int kc = max(k[u], -1100), k1 = kc >> 1, k2 = kc - k1;
    const double s1 = __hiloint2double((k1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((k2 + 1023) << 20, 0);
    out[u] = (p[u] * s1) * s2;
  }
}

// ------------------------------------------------------------------------------------------
// a14: KDE bandwidth by leave-one-out likelihood cross-validation (AMP.manikde! ->
// getKDEManifoldBandwidths -> KDE kde!(x) "lcv"; call sites ApproxConv.jl:36-42,
// GraphProductOperations.jl:53).  Exact O(N^2) evaluation (KDE.setForceEvalDirect!(true),
// src/IncrementalInference.jl:104).
//
// Objective -1/N sum_i log( 1/(N-1) sum_{j!=i} N(x_i - x_j; 0, h^2) ).
// Thread layout: lane <-> row i, the G = threads/roundup(N,32) thread segments split the columns,
// four independent exp chains per thread hide the FP64 latency, no shuffles in the hot loop.
//  * symmetric form (N <= IIF_LOO_SYM_MAX): the kernel matrix is symmetric with a zero diagonal,
//    so every unordered pair is evaluated ONCE: row i covers the circulant half j = i+1..i+N/2
//    (mod N), keeps its row part in a register and stores e(i,k) to the shared tile E[k][i]; after
//    one barrier thread j gathers the mirrored column part sum_k E[k][j-k].  Plain stores and
//    loads only (conflict-free: consecutive lanes touch consecutive addresses).
//  * full form (larger N, tile would not fit): every thread sums its share of row i directly.
// scr layout: part[IIF_LOO_PARTS][N] segment partials, then the E tile (symmetric form only).
// ------------------------------------------------------------------------------------------
#define IIF_LOO_SYM_MAX 160
#define IIF_LOO_PARTS 8
__host__ __device__ inline bool loo_sym(int N) { return N <= IIF_LOO_SYM_MAX; }
__host__ __device__ inline int loo_scratch_doubles(int N) {
  return IIF_LOO_PARTS * N + (loo_sym(N) ? ((N >> 1) + 1) * N : 0);
}
#define IIF_LOO_SCRATCH_N(N) loo_scratch_doubles(N)

template <bool CIRC>
__device__ __forceinline__ double loo_thread_sum(const double* __restrict__ x, int N, double c, double* E,
                                                 int i, int seg, int G, bool active) {
  double acc = 0.0;
  if (!active) return acc;
  const double xi = x[i];
  if (loo_sym(N)) {
    const int hN = N >> 1;
    const bool even = (N & 1) == 0;
    const int nk = ((N - 1) >> 1) + (even ? 1 : 0);  // k = 1..nk ; k == N/2 (even N) only for i < N/2
    const int per = (nk + G - 1) / G;
    const int k0 = 1 + seg * per, k1 = min(nk, (seg + 1) * per);
    // the antipodal column (even N, k == N/2) counts once: rows i >= N/2 store and add 0 there
    const bool cut = even && k1 == hN && i >= hN;
    int k = k0;
    for (; k + 3 <= k1; k += 4) {  // full groups: no per-element checks except the antipodal one
      double a[4], e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int j = i + k + u;
        j -= (j >= N) ? N : 0;
        const double dl = mdiff(xi, x[j], CIRC);
        a[u] = dl * dl * c;
      }
      exp_neg4(a, e);
      if (cut && k + 3 == k1) e[3] = 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc += e[u];
        E[(k + u - 1) * N + i] = e[u];
      }
    }
    for (; k <= k1; ++k) {  // tail (< 4 columns)
      int j = i + k;
      j -= (j >= N) ? N : 0;
      const double dl = mdiff(xi, x[j], CIRC);
      double e = exp_neg(dl * dl * c);
      if (cut && k == k1) e = 0.0;
      acc += e;
      E[(k - 1) * N + i] = e;
    }
  } else {
    const int per = (N + G - 1) / G;
    const int j0 = seg * per, j1 = min(N, (seg + 1) * per) - 1;
    for (int j = j0; j <= j1; j += 4) {
      double a[4], e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = min(j + u, j1);
        const double dl = mdiff(xi, x[jj], CIRC);
        a[u] = dl * dl * c;
      }
      exp_neg4(a, e);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += ((j + u <= j1) && (j + u != i)) ? e[u] : 0.0;
    }
  }
  return acc;
}

// KTAG gives every kernel its own instantiation (and register budget) of the noinline search code.
template <int KTAG>
__device__ __noinline__ double loo_nll(const double* __restrict__ x, int N, int circ, double h, double* scr,
                                       double* red, int* parity_io) {
  const double c = -1.0 / (2.0 * h * h);
  const double lognorm = log((double)(N - 1) * sqrt(IIF_TWO_PI) * h);
  const int Npad = (N + 31) & ~31;
  int G = IIF_NT / Npad;
  G = G > IIF_LOO_PARTS ? IIF_LOO_PARTS : G;
  const int seg = threadIdx.x / Npad, i = threadIdx.x - seg * Npad;
  const bool active = (seg < G) && (i < N);
  double* part = scr;
  double* E = scr + IIF_LOO_PARTS * N;
  double acc = circ ? loo_thread_sum<true>(x, N, c, E, i, seg, G, active)
                    : loo_thread_sum<false>(x, N, c, E, i, seg, G, active);
  if (loo_sym(N)) {
    __syncthreads();
    if (active) {  // mirrored column part: sum_k E[k][j - k], same k-range as this thread's row part
      const int hN = N >> 1;
      const int nk = ((N - 1) >> 1) + (((N & 1) == 0) ? 1 : 0);
      const int per = (nk + G - 1) / G;
      const int k0 = 1 + seg * per, k1 = min(nk, (seg + 1) * per);
      double a0 = 0.0, a1 = 0.0;
      int k = k0;
      for (; k + 1 <= k1; k += 2) {
        int r0 = i - k, r1 = i - k - 1;
        r0 += (r0 < 0) ? N : 0;
        r1 += (r1 < 0) ? N : 0;
        a0 += E[(k - 1) * N + r0];
        a1 += E[k * N + r1];
      }
      if (k <= k1) {
        int r0 = i - k;
        r0 += (r0 < 0) ? N : 0;
        a0 += E[(k - 1) * N + r0];
      }
      acc += a0 + a1;
      (void)hN;
    }
  }
  double term = 0.0;
  if (G > 1) {
    if (active) part[seg * N + i] = acc;
    __syncthreads();
    if (seg == 0 && i < N) {
      double t = 0.0;
      for (int g = 0; g < G; ++g) t += part[g * N + i];
      term = log(t) - lognorm;
    }
  } else if (active) {
    term = log(acc) - lognorm;
  }
  int parity = *parity_io;
  const double tot = block_sum1(term, red, parity);  // its barrier also protects scr for the next call
  *parity_io = parity;
  return -tot / (double)N;
}

// Numerical-Recipes golden section as used by KDE `golden(npd, nLOO_LL, ax, bx, cx, tol)`;
// the search variable scales the base bandwidth h0.  Uniform control flow across the CTA.
template <int KTAG>
__device__ double golden_nr(const double* x, int N, double h0, double ax, double bx, double cx, double tol,
                            double* scr, double* red, int& parity) {
  const double C = (3.0 - sqrt(5.0)) / 2.0, R = 1.0 - C;
  double x0 = ax, x3 = cx, x1, x2;
  if (fabs(cx - bx) > fabs(bx - ax)) { x1 = bx; x2 = bx + C * (cx - bx); }
  else { x2 = bx; x1 = bx - C * (bx - ax); }
  double f1 = loo_nll<KTAG>(x, N, 0, x1 * h0, scr, red, &parity);
  double f2 = loo_nll<KTAG>(x, N, 0, x2 * h0, scr, red, &parity);
  for (int it = 0; it < 200 && fabs(x3 - x0) > tol * (fabs(x1) + fabs(x2)); ++it) {
    const bool right = f2 < f1;
    double xn;
    if (right) { x0 = x1; x1 = x2; x2 = R * x1 + C * x3; f1 = f2; xn = x2; }
    else { x3 = x2; x2 = x1; x1 = R * x2 + C * x0; f2 = f1; xn = x1; }
    const double fn = loo_nll<KTAG>(x, N, 0, xn * h0, scr, red, &parity);
    if (right) f2 = fn; else f1 = fn;
  }
  return (f1 < f2) ? x1 : x2;
}

// Optim.jl GoldenSection on [lo, hi] as used by AMP kde!_CircularNaiveCV
template <int KTAG>
__device__ double golden_optim(const double* x, int N, double lo, double hi, double rel_tol, double* scr,
                               double* red, int& parity) {
  const double gr = 0.5 * (3.0 - sqrt(5.0));
  const double abs_tol = 2.220446049250313e-16;
  double xm = lo + gr * (hi - lo);
  double fm = loo_nll<KTAG>(x, N, 1, xm, scr, red, &parity);
  for (int it = 0; it < 200; ++it) {
    double tolx = rel_tol * fabs(xm) + abs_tol;
    double mid = 0.5 * (hi + lo);
    if (fabs(xm - mid) <= 2 * tolx - 0.5 * (hi - lo)) break;
    const bool up = hi - xm > xm - lo;
    const double xn = up ? xm + gr * (hi - xm) : xm - gr * (xm - lo);
    const double fn = loo_nll<KTAG>(x, N, 1, xn, scr, red, &parity);
    if (up) {
      if (fn < fm) { lo = xm; xm = xn; fm = fn; } else hi = xn;
    } else {
      if (fn < fm) { hi = xm; xm = xn; fm = fn; } else lo = xn;
    }
  }
  return xm;
}

// Per-dimension bandwidth of the N x d points in `pts` (shared or global memory).
// xa, xb: shared scratch of N doubles each; scr: IIF_NW*N + N doubles.  Result bw[c] is returned
// to every thread.
template <int KTAG>
__device__ void block_kde_bandwidth(const double* pts, int N, int d, int32_t circ_mask, const TreeStruct& T,
                                    double* xa, double* xb, double* scr, double* red, int& parity, double* bw) {
  for (int c = 0; c < d; ++c) {
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += IIF_NT) xa[i] = pts[i * d + c];
    __syncthreads();
    if (is_circ(circ_mask, c)) {
      bw[c] = golden_optim<KTAG>(xa, N, 1e-3, IIF_TWO_PI, 1e-3, scr, red, parity);
    } else {
      // rank sort (ties by index) -> xb ascending
      for (int i = threadIdx.x; i < N; i += IIF_NT) {
        double xi = xa[i];
        int r = 0;
        for (int k = 0; k < N; ++k) {
          double xk = xa[k];
          r += (xk < xi) || (xk == xi && k < i);
        }
        xb[r] = xi;
      }
      __syncthreads();
      // KDE neighborMinMax: root ball diameter and smallest internal ball diameter (>= 1e-6)
      double maxm = xb[N - 1] - xb[0];
      double m = maxm;
      for (int z = threadIdx.x; z < T.nn; z += IIF_NT) {
        int lo = T.lo[z], hi = T.hi[z];
        if (hi > lo) m = fmin(m, xb[hi] - xb[lo]);
      }
      double minm = block_min1(m, red, parity);
      if (minm < 1e-6) minm = 1e-6;
      double h0 = 0.5 * (minm + maxm);
      double a = golden_nr<KTAG>(xb, N, h0, 2.0 * minm / (minm + maxm), 1.0, 2.0 * maxm / (minm + maxm), 1e-2, scr,
                           red, parity);
      bw[c] = a * h0;
    }
  }
}
