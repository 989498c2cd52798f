// iif_device.cuh — device-side building blocks shared by the convolution and product kernels.
//
// Everything here is FP64 (the reference path is Float64 end to end) and written for one CTA
// of IIF_THREADS threads working on one belief (N <= IIF_MAX_POINTS particles resident in
// shared memory).  Reference citations are to IncrementalInference.jl v0.35.6.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/iifb200.h"

#define IIF_THREADS 256
#define IIF_WARPS (IIF_THREADS / 32)
#define IIF_PI 3.14159265358979323846
#define IIF_TWO_PI 6.28318530717958647692
#define IIF_RED_KMAX 8
#define IIF_RED_DOUBLES (2 * IIF_WARPS * IIF_RED_KMAX)

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
  double* buf = red + parity * (IIF_WARPS * IIF_RED_KMAX);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) buf[warp * K + k] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < IIF_WARPS; ++w) s += buf[w * K + k];
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
  double* buf = red + parity * (IIF_WARPS * IIF_RED_KMAX);
  double s = warp_min(x);
  if (lane == 0) buf[warp] = s;
  __syncthreads();
  double m = buf[0];
#pragma unroll
  for (int w = 1; w < IIF_WARPS; ++w) m = fmin(m, buf[w]);
  parity ^= 1;
  return m;
}

// ------------------------------------------------------------------------------------------
// a14: KDE bandwidth by leave-one-out likelihood cross-validation (AMP.manikde! ->
// getKDEManifoldBandwidths -> KDE kde!(x) "lcv"; call sites ApproxConv.jl:36-42,
// GraphProductOperations.jl:53).  Exact O(N^2) evaluation (KDE.setForceEvalDirect!(true),
// src/IncrementalInference.jl:104).
//
// Objective -1/N sum_i log( 1/(N-1) sum_{j!=i} N(x_i - x_j; 0, h^2) ): one warp per row i,
// lanes stride the columns j, warp-shuffle reduction of the kernel sums, one log per row
// (lane k keeps the sum of the warp's k-th row), block reduction of the log terms.
// ------------------------------------------------------------------------------------------
__device__ double loo_nll(const double* __restrict__ x, int N, bool circ, double h, double* red, int& parity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double ninv2h2 = -1.0 / (2.0 * h * h);
  const double lognorm = log((double)(N - 1) * sqrt(IIF_TWO_PI) * h);
  double acc = 0.0;
  for (int i0 = warp; i0 < N; i0 += IIF_WARPS * 32) {
    double mine = 1.0;  // log(1) = 0 for lanes without a row
    int k = 0;
    for (int i = i0; i < N && k < 32; i += IIF_WARPS, ++k) {
      const double xi = x[i];
      double s = 0.0;
      for (int j = lane; j < N; j += 32) {
        double dl = mdiff(xi, x[j], circ);
        double e = exp(dl * dl * ninv2h2);
        s += (j == i) ? 0.0 : e;
      }
      s = warp_sum(s);
      if (lane == k) mine = s;
    }
    if (lane < k) acc += log(mine) - lognorm;
  }
  double tot = block_sum1(acc, red, parity);
  return -tot / (double)N;
}

// Numerical-Recipes golden section as used by KDE `golden(npd, nLOO_LL, ax, bx, cx, tol)`;
// the search variable scales the base bandwidth h0.  Uniform control flow across the CTA.
__device__ double golden_nr(const double* x, int N, double h0, double ax, double bx, double cx, double tol,
                            double* red, int& parity) {
  const double C = (3.0 - sqrt(5.0)) / 2.0, R = 1.0 - C;
  double x0 = ax, x3 = cx, x1, x2;
  if (fabs(cx - bx) > fabs(bx - ax)) { x1 = bx; x2 = bx + C * (cx - bx); }
  else { x2 = bx; x1 = bx - C * (bx - ax); }
  double f1 = loo_nll(x, N, false, x1 * h0, red, parity);
  double f2 = loo_nll(x, N, false, x2 * h0, red, parity);
  for (int it = 0; it < 200 && fabs(x3 - x0) > tol * (fabs(x1) + fabs(x2)); ++it) {
    if (f2 < f1) {
      x0 = x1; x1 = x2; x2 = R * x1 + C * x3;
      f1 = f2; f2 = loo_nll(x, N, false, x2 * h0, red, parity);
    } else {
      x3 = x2; x2 = x1; x1 = R * x2 + C * x0;
      f2 = f1; f1 = loo_nll(x, N, false, x1 * h0, red, parity);
    }
  }
  return (f1 < f2) ? x1 : x2;
}

// Optim.jl GoldenSection on [lo, hi] as used by AMP kde!_CircularNaiveCV
__device__ double golden_optim(const double* x, int N, double lo, double hi, double rel_tol, double* red,
                               int& parity) {
  const double gr = 0.5 * (3.0 - sqrt(5.0));
  const double abs_tol = 2.220446049250313e-16;
  double xm = lo + gr * (hi - lo);
  double fm = loo_nll(x, N, true, xm, red, parity);
  for (int it = 0; it < 200; ++it) {
    double tolx = rel_tol * fabs(xm) + abs_tol;
    double mid = 0.5 * (hi + lo);
    if (fabs(xm - mid) <= 2 * tolx - 0.5 * (hi - lo)) break;
    if (hi - xm > xm - lo) {
      double xn = xm + gr * (hi - xm);
      double fn = loo_nll(x, N, true, xn, red, parity);
      if (fn < fm) { lo = xm; xm = xn; fm = fn; } else hi = xn;
    } else {
      double xn = xm - gr * (xm - lo);
      double fn = loo_nll(x, N, true, xn, red, parity);
      if (fn < fm) { hi = xm; xm = xn; fm = fn; } else lo = xn;
    }
  }
  return xm;
}

// Per-dimension bandwidth of the N x d points in `pts` (shared or global memory).
// xa, xb: shared scratch of N doubles each.  Result bw[c] is returned to every thread.
__device__ void block_kde_bandwidth(const double* pts, int N, int d, int32_t circ_mask, const TreeStruct& T,
                                    double* xa, double* xb, double* red, int& parity, double* bw) {
  for (int c = 0; c < d; ++c) {
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += IIF_THREADS) xa[i] = pts[i * d + c];
    __syncthreads();
    if (is_circ(circ_mask, c)) {
      bw[c] = golden_optim(xa, N, 1e-3, IIF_TWO_PI, 1e-3, red, parity);
    } else {
      // rank sort (ties by index) -> xb ascending
      for (int i = threadIdx.x; i < N; i += IIF_THREADS) {
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
      for (int z = threadIdx.x; z < T.nn; z += IIF_THREADS) {
        int lo = T.lo[z], hi = T.hi[z];
        if (hi > lo) m = fmin(m, xb[hi] - xb[lo]);
      }
      double minm = block_min1(m, red, parity);
      if (minm < 1e-6) minm = 1e-6;
      double h0 = 0.5 * (minm + maxm);
      double a = golden_nr(xb, N, h0, 2.0 * minm / (minm + maxm), 1.0, 2.0 * maxm / (minm + maxm), 1e-2, red,
                           parity);
      bw[c] = a * h0;
    }
  }
}
