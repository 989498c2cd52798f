// iif_device.cuh — device-side building blocks shared by the convolution and product kernels.
//
// Everything here is FP64 (the reference path is Float64 end to end) and written for one CTA
// of IIF_NT threads working on one belief (N <= IIF_MAX_POINTS particles resident in
// shared memory).  Reference citations are to IncrementalInference.jl v0.35.6.
#pragma once
#include <cooperative_groups.h>
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

// Phase clocks for kernel development (profiles/phase_probe.py builds a copy of the library with
// -DIIF_PHASES; compiled out of the product library): thread IIF_PHASE_TID of block 0 accumulates the
// clock64() time between consecutive marks.
#ifdef IIF_PHASES
#ifndef IIF_PHASE_TID
#define IIF_PHASE_TID 0
#endif
__device__ long long g_iif_phase[32];
__shared__ long long s_iif_phase[32];  // shared accumulators: a mark costs one LDS/STS round trip, not a global one
#define IIF_PHASE_BEGIN() long long ph_t_ = clock64()
#define IIF_PHASE(k)                                             \
  do {                                                           \
    if (threadIdx.x == IIF_PHASE_TID && blockIdx.x == 0) {       \
      const long long t_ = clock64();                            \
      s_iif_phase[k] += t_ - ph_t_;                              \
      ph_t_ = clock64();                                         \
    }                                                            \
  } while (0)
#define IIF_PHASE_ZERO()                                                                        \
  do {                                                                                          \
    if (threadIdx.x == IIF_PHASE_TID && blockIdx.x == 0)                                        \
      for (int k_ = 0; k_ < 32; ++k_) s_iif_phase[k_] = 0;                                      \
  } while (0)
#define IIF_PHASE_FLUSH()                                                                       \
  do {                                                                                          \
    if (threadIdx.x == IIF_PHASE_TID && blockIdx.x == 0)                                        \
      for (int k_ = 0; k_ < 32; ++k_) g_iif_phase[k_] += s_iif_phase[k_];                       \
  } while (0)
#else
#define IIF_PHASE_BEGIN()
#define IIF_PHASE(k)
#define IIF_PHASE_ZERO()
#define IIF_PHASE_FLUSH()
#endif

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

// SpecialOrthogonal(3): points travel as rotation vectors omega = vee(log(eps, R)), |omega| <= pi.  Group operations go
// through unit quaternions (w, v) = (cos(|omega|/2), sin(|omega|/2) omega/|omega|).
__device__ __forceinline__ bool is_so3(int32_t mask) { return (mask & IIF_MANI_SO3) != 0; }
struct Quat { double w, x, y, z; };
__device__ __forceinline__ Quat so3_quat(const double* om) {
  const double t2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double s, w;
  if (t2 < 1e-16) { s = 0.5 - t2 / 48.0; w = 1.0 - t2 / 8.0; }
  else { const double t = sqrt(t2); sincos(0.5 * t, &s, &w); s /= t; }
  Quat q = {w, s * om[0], s * om[1], s * om[2]};
  return q;
}
__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r = {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
  return r;
}
__device__ __forceinline__ void so3_rotvec(Quat q, double* om) {  // Log: rotation angle in [0, pi]
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  const double vn = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  const double k = vn < 1e-12 ? 2.0 : 2.0 * atan2(vn, q.w) / vn;
  om[0] = k * q.x; om[1] = k * q.y; om[2] = k * q.z;
}
// out = Log(Exp(a) Exp(b)): p Exp(X) for a point p = Exp(a) and a tangent X = hat(b) (exp / retract at p, compose)
__device__ __forceinline__ void so3_compose(const double* a, const double* b, double* out) {
  so3_rotvec(quat_mul(so3_quat(a), so3_quat(b)), out);
}
// out = Log(Exp(a)^T Exp(b)) = vee(log(p, q)) for p = Exp(a), q = Exp(b)
__device__ __forceinline__ void so3_between(const double* a, const double* b, double* out) {
  Quat qa = so3_quat(a);
  qa.x = -qa.x; qa.y = -qa.y; qa.z = -qa.z;
  so3_rotvec(quat_mul(qa, so3_quat(b)), out);
}

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

// `nwa`: warps that can hold a non-zero contribution (items 0..n-1 dealt to threads round-robin occupy the
// first ceil(n/32) warps); the others skip the shuffles and are not read back.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* red, int& parity, int nwa = IIF_MAX_WARPS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* buf = red + parity * (IIF_MAX_WARPS * IIF_RED_KMAX);
  nwa = min(nwa, IIF_NW);
  if (warp < nwa) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double s = warp_sum(v[k]);
      if (lane == 0) buf[warp * K + k] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0;
    for (int w = 0; w < nwa; ++w) s += buf[w * K + k];
    v[k] = s;
  }
  parity ^= 1;
}
__device__ __forceinline__ double block_sum1(double x, double* red, int& parity, int nwa = IIF_MAX_WARPS) {
  double v[1] = {x};
  block_sum<1>(v, red, parity, nwa);
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
// exp(x) for x <= 0 in FP64 (the only use on this path: Gaussian kernel weights).
//   x = k ln2 + r, |r| <= ln2/2 (one FMA against the double nearest ln2: the reduction error
//   |k| 2.3e-17 is below one ulp of the result wherever the result matters), degree-11 minimax
//   polynomial 1 + r + r^2 q(r) (approximation error 1.3e-17, coefficients below from a Chebyshev fit
//   of (e^r - 1 - r)/r^2, read straight from the constant bank), and 2^k applied by an integer add
//   on the exponent field.  Arguments below -708 (subnormal results), -inf and NaN take the careful
//   path (two-step power-of-two scaling, exact gradual underflow like libm).
// ------------------------------------------------------------------------------------------
__constant__ double IIF_EXPC[16] = {
    2.510037583256132e-08,  2.762007587998348e-07,  2.7557268480310024e-06, 2.4801521322368692e-05,
    1.9841269863040545e-04, 1.3888888917196719e-03, 8.333333333330065e-03,  4.1666666666624164e-02,
    1.6666666666666669e-01, 5.000000000000001e-01,  0.0,                    0.0,
    1.4426950408889634074,  -6.931471805599453094e-01, 0.0,                 6755399441055744.0};
#define IIF_EXP_NQ 10  // coefficients of q, highest degree first

__device__ __forceinline__ double exp_scale2(double p, int k) {  // p * 2^k, any k <= 0, gradual underflow
  const int kc = max(k, -1100), k1 = kc >> 1, k2 = kc - k1;
  const double s1 = __hiloint2double((k1 + 1023) << 20, 0);
  const double s2 = __hiloint2double((k2 + 1023) << 20, 0);
  return (p * s1) * s2;
}

__device__ __forceinline__ double exp_neg(double x) {  // careful scalar form
  if (x < -1000.0) return 0.0;                         // also -inf; NaN falls through and propagates
  double t = fma(x, IIF_EXPC[12], IIF_EXPC[15]);
  const int k = __double2loint(t);
  t -= IIF_EXPC[15];
  const double r = fma(t, IIF_EXPC[13], x);
  double p = IIF_EXPC[0];
#pragma unroll
  for (int c = 1; c < IIF_EXP_NQ; ++c) p = fma(p, r, IIF_EXPC[c]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return exp_scale2(p, k);
}

// U exponentials in lockstep: every step is issued for all U arguments before the next one, so the U
// FP64 dependency chains interleave and each coefficient is fetched once.
template <int U>
__device__ __forceinline__ void exp_negU(const double (&x)[U], double (&out)[U]) {
  const double L2E = IIF_EXPC[12], NLN2 = IIF_EXPC[13], MAGIC = IIF_EXPC[15];
  double t[U], r[U], p[U];
  unsigned worst = 0u;  // largest high word: negative arguments sort by magnitude as unsigned
#pragma unroll
  for (int u = 0; u < U; ++u) worst = max(worst, (unsigned)__double2hiint(x[u]));
#pragma unroll
  for (int u = 0; u < U; ++u) t[u] = fma(x[u], L2E, MAGIC);
#pragma unroll
  for (int u = 0; u < U; ++u) r[u] = t[u] - MAGIC;
#pragma unroll
  for (int u = 0; u < U; ++u) r[u] = fma(r[u], NLN2, x[u]);
  {
    const double c0 = IIF_EXPC[0], c1 = IIF_EXPC[1];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(c0, r[u], c1);
  }
#pragma unroll
  for (int c = 2; c < IIF_EXP_NQ; ++c) {
    const double cc = IIF_EXPC[c];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], cc);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], 1.0);
#pragma unroll
  for (int u = 0; u < U; ++u) p[u] = fma(p[u], r[u], 1.0);
  if (worst < 0xC0862000u) {  // every argument in (-708, +0]: 2^k is a plain exponent-field add
#pragma unroll
    for (int u = 0; u < U; ++u)
      out[u] = __hiloint2double(__double2hiint(p[u]) + (int)((unsigned)__double2loint(t[u]) << 20), __double2loint(p[u]));
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) out[u] = exp_neg(x[u]);
  }
}

// Gaussian kernel values exp(-z^2 ln2/256) for U scaled differences z (z = delta * sqrt(128 log2 e) / h gives
// exp(-delta^2 / 2h^2)), in lockstep.  -z^2 = n + f with n = round(-z^2) taken by the magic-number FMA
// straight from the exact product, f = fma(-z, z, -n) in [-1/2, 1/2] (single rounding), and
//   2^((n + f)/256) = 2^(n >> 8) * T[n & 255] * P(f),   T[j] = 2^(j/256) (256-entry shared table),
// P = degree-4 Taylor polynomial of 2^(f/256) (|f ln2/256| <= 1.36e-3: truncation error 3.8e-17).
// 11 FP64 instructions per kernel value including the difference and the accumulation (16-entry table + degree 6: 13,
// libm-style exp: 20+); the price is a table lookup with bank conflicts, on the load/store pipe the FP64-bound loops
// leave idle.  Values below 2^-1020 (and inf / NaN arguments) take the careful scalar path.
#define IIF_GTAB_N 256
#define IIF_GTAB_BITS 8
__constant__ double IIF_G16C[8] = {2.239395190875157e-12, 3.308302680541371e-09, 3.665565596910106e-06,
                                   0.0027076061740622863, 6755399441055744.0, 0.0, 0.0, 0.0};
__constant__ double IIF_EXP2TAB[IIF_GTAB_N] = {
    1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175,
    1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738,
    1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041,
    1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316,
    1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264,
    1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645,
    1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791,
    1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697,
    1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222,
    1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175,
    1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302,
    1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055,
    1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179,
    1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988,
    1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373,
    1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938,
    1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806,
    1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012,
    1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455,
    1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165,
    1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162,
    1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063,
    1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693,
    1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895,
    1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773,
    1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575,
    1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458,
    1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345,
    1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012,
    1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238,
    1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003,
    1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707,
    1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942,
    1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257,
    1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438,
    1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675,
    1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397,
    1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065,
    1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146,
    1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568,
    1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887,
    1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414,
    1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164,
    1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771,
    1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289,
    1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063,
    1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162,
    1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785,
    1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326,
    1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926,
    1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311,
    1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782,
    1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547,
    1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401,
    1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032,
    1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973,
    1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508,
    1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386,
    1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392,
    1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899,
    1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627,
    1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783,
    1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343,
    1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402};
#define IIF_GSCALE 13.589148804608305  // sqrt(128 log2(e))
#define IIF_GLN2 2.7076061740622863e-03  // ln2 / 256
#define IIF_GCLAMP_HI 0x407FE000       // high word of 510.0: arguments are clamped there (2^(-510^2/256) < 1e-305)

template <int U>
__device__ __forceinline__ void gauss_negU(const double (&z)[U], const double* __restrict__ tab, double (&out)[U]) {
  const double MAGIC = IIF_G16C[4];
  double t[U], f[U], p[U];
#pragma unroll
  for (int u = 0; u < U; ++u) t[u] = fma(-z[u], z[u], MAGIC);
#pragma unroll
  for (int u = 0; u < U; ++u) f[u] = t[u] - MAGIC;  // n as a double (<= 0)
  unsigned worst = 0u;  // largest high word of n: negative values sort by magnitude as unsigned
#pragma unroll
  for (int u = 0; u < U; ++u) worst = max(worst, (unsigned)__double2hiint(f[u]));
#pragma unroll
  for (int u = 0; u < U; ++u) f[u] = fma(-z[u], z[u], -f[u]);
  {
    const double c0 = IIF_G16C[0], c1 = IIF_G16C[1];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(c0, f[u], c1);
  }
#pragma unroll
  for (int c = 2; c < 4; ++c) {
    const double cc = IIF_G16C[c];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = fma(p[u], f[u], cc);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) p[u] = fma(p[u], f[u], 1.0);
  if (worst <= (unsigned)__double2hiint(-261120.0)) {  // every value >= 2^-1020: 2^(n>>8) is an exponent-field add
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int n = __double2loint(t[u]);
      const double v = p[u] * tab[n & (IIF_GTAB_N - 1)];
      out[u] = __hiloint2double(__double2hiint(v) + (n >> IIF_GTAB_BITS) * 1048576, __double2loint(v));
    }
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) out[u] = exp_neg(-(z[u] * z[u]) * IIF_GLN2);
  }
}

// stage the 2^(j/256) table in shared memory (all threads of the CTA; the caller synchronises before the first use)
__device__ __forceinline__ void gauss_stage_table(double* tab) {
  for (int j = threadIdx.x; j < IIF_GTAB_N; j += IIF_NT) tab[j] = IIF_EXP2TAB[j];
}

// ------------------------------------------------------------------------------------------
// a14: KDE bandwidth by leave-one-out likelihood cross-validation (AMP.manikde! ->
// getKDEManifoldBandwidths -> KDE kde!(x) "lcv"; call sites ApproxConv.jl:36-42,
// GraphProductOperations.jl:53).  Exact O(N^2) evaluation (KDE.setForceEvalDirect!(true),
// src/IncrementalInference.jl:104).
//
// Objective -1/N sum_i log( 1/(N-1) sum_{j!=i} N(x_i - x_j; 0, h^2) ).
// The kernel matrix is symmetric with a zero diagonal, so every unordered pair is evaluated ONCE:
// row i covers the circulant half j = i+1..i+N/2 (mod N).  Thread t owns row i = t mod N and the
// column segment seg = t / N (G = threads/N segments, columns split evenly so that only a thread's
// last lock-step group can hold padding).  It reads its partners from a doubled copy of the
// coordinates (x2[i+k], no index wrap), keeps its row part in a register and stores e(i,k) to the
// shared tile E[k][i]; after one barrier the same thread gathers the mirrored column part
// sum_k E[k][i-k] (a stride N-1 walk with one wrap).  Plain conflict-free loads and stores only.
// For even N the antipodal column k = N/2 is evaluated by both of its rows and never stored.
// Beliefs whose half matrix exceeds the tile are processed in column chunks.
// Everything that depends only on (N, thread) is computed once per belief (LooThread), outside the
// ~16 objective evaluations of the golden-section search.
// scr layout: part[IIF_LOO_PARTS][N] segment partials, then the E tile of loo_tile_rows(N) rows.
// ------------------------------------------------------------------------------------------
#define IIF_LOO_PARTS 8
#define IIF_LOO_TILE_DOUBLES 12800  // 100 KB: the whole half matrix up to N = 160
#define IIF_LOO_XPAD 8              // slack behind the doubled coordinates (padding lanes of the last group)
__host__ __device__ inline int loo_nk(int N) { return ((N - 1) >> 1) + (((N & 1) == 0) ? 1 : 0); }
__host__ __device__ inline int loo_chunks(int N) {
  const int nk = loo_nk(N), cap = IIF_LOO_TILE_DOUBLES / N;
  return (nk + cap - 1) / cap;
}
__host__ __device__ inline int loo_tile_rows(int N) {  // equal-sized column chunks
  const int nc = loo_chunks(N);
  return (loo_nk(N) + nc - 1) / nc;
}
__host__ __device__ inline int loo_scratch_doubles(int N) { return IIF_LOO_PARTS * N + loo_tile_rows(N) * N; }
__host__ __device__ inline int loo_x2_doubles(int N) { return 2 * N + IIF_LOO_XPAD; }
// the coordinate buffer `xa` in front of it: the compiler vectorises the rank sort's reads of xa (LDS.128, eight
// values per trip) and may read up to 7 values past xa[N-1]; they are discarded, the slack keeps them off xb
__host__ __device__ inline int loo_xa_doubles(int N) { return N + 8; }
#define IIF_LOO_SCRATCH_N(N) loo_scratch_doubles(N)

struct LooCfg {  // per belief size N and CTA size (uniform)
  int N, nk, nks, Kc, nchunks, G, U, nrw;
};

__device__ __forceinline__ LooCfg loo_config(int N) {
  LooCfg L;
  L.N = N;
  L.nk = loo_nk(N);
  L.nks = (N - 1) >> 1;  // stored (mirrored) columns: all but the antipodal one
  L.Kc = loo_tile_rows(N);
  L.nchunks = loo_chunks(N);
  int G = IIF_NT / N;
  G = G > IIF_LOO_PARTS ? IIF_LOO_PARTS : G;
  G = G > L.Kc ? L.Kc : G;
  L.G = G < 1 ? 1 : G;
  const int per = (L.Kc + L.G - 1) / L.G;
  // lock-step width: least padded work among 4, 5, 6 (ties: the widest, more independent chains)
  int best = 4, cost = ((per + 3) / 4) * 4;
  if (((per + 4) / 5) * 5 <= cost) { best = 5; cost = ((per + 4) / 5) * 5; }
  if (((per + 5) / 6) * 6 <= cost) { best = 6; cost = ((per + 5) / 6) * 6; }
  L.U = best;
  L.nrw = (N + 31) >> 5;
  return L;
}

struct LooThread {  // this thread's share of one column chunk
  int cnt;          // columns evaluated
  int cnts;         // columns stored and gathered (<= cnt)
  int wrap;         // column (0-based within the thread) at which the partner row index i + k wraps
  int xoff;         // x2 offset of the first partner: i + kb
  int soff;         // tile offset of the first store:  (kb - cf) * N + (i + kb)  [- N once wrapped]
  int goff;         // tile offset of the first gather: (kb - cf) * N + i
};

__device__ __forceinline__ LooThread loo_thread(const LooCfg& L, int seg, int i, int ch) {
  LooThread T;
  const int cf = 1 + ch * L.Kc;
  const int cols = min(L.Kc, L.nk - cf + 1);
  const int kb = cf + (seg * cols) / L.G, ke = cf + ((seg + 1) * cols) / L.G - 1;
  const bool on = seg < L.G && kb <= ke;
  T.cnt = on ? ke - kb + 1 : 0;
  T.cnts = on ? max(min(ke, L.nks) - kb + 1, 0) : 0;
  T.wrap = L.N - (i + kb);              // < 0: wrapped from the first column on
  T.xoff = i + kb;
  T.soff = (kb - cf) * L.N + (i + kb) - (T.wrap < 0 ? L.N : 0);
  T.goff = (kb - cf) * L.N + i;
  return T;
}

// row part: sum_k e(i, i+k) over this thread's columns; e(i, i+k) is stored to tile row k at the position
// of the partner row j = (i + k) mod N, so that the gather below is a plain stride-N walk
template <int U, bool CIRC>
__device__ __forceinline__ double loo_rows(const double* __restrict__ xp, double xi, double sc,
                                           const double* __restrict__ tab, double* __restrict__ ep, int N, int cnt,
                                           int cnts, int wrap) {
  double acc0 = 0.0, acc1 = 0.0;
  int done = 0;
  for (; done + U <= cnts; done += U) {  // full groups: every column valid and stored
    double a[U], e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = mdiff(xi, xp[u], CIRC) * sc;
    gauss_negU<U>(a, tab, e);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (u & 1) acc1 += e[u]; else acc0 += e[u];
      if (done + u == wrap) ep -= N;
      *ep = e[u];
      ep += N + 1;
    }
    xp += U;
  }
  if (done < cnt) {  // last group: padding lanes are dropped, the antipodal column is not stored
    double a[U], e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = mdiff(xi, xp[u], CIRC) * sc;
    gauss_negU<U>(a, tab, e);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (done + u < cnt) acc0 += e[u];
      if (done + u == wrap) ep -= N;
      if (done + u < cnts) *ep = e[u];
      ep += N + 1;
    }
  }
  return acc0 + acc1;
}

// mirrored column part: sum_k E[k][i] over the stored columns of this thread
__device__ __forceinline__ double loo_gather(const double* __restrict__ gp, int N, int cnts) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int m = 0;
#pragma unroll 1
  for (; m + 3 < cnts; m += 4) {
    a0 += gp[0];
    a1 += gp[N];
    a2 += gp[2 * N];
    a3 += gp[3 * N];
    gp += 4 * N;
  }
#pragma unroll 1
  for (; m < cnts; ++m) { a0 += gp[0]; gp += N; }
  return (a0 + a1) + (a2 + a3);
}

// one objective evaluation; x2 = doubled coordinates in shared memory.  invK = 1 / ((N-1) sqrt(2 pi)).
template <int U, bool CIRC>
__device__ __forceinline__ double loo_nll(const double* __restrict__ x2, const LooCfg& L, const LooThread& T0, int seg,
                                          int i, double xi, double h, double invK, double negInvN,
                                          const double* __restrict__ tab, double* scr, double* red, int& parity) {
  const int N = L.N;
  IIF_PHASE_BEGIN();
  const double rh = 1.0 / h;
  const double sc = IIF_GSCALE * rh;  // exp(-delta^2 / 2h^2) = 2^(-(delta sc)^2 / 256)
  const double norm = rh * invK;  // row sums are normalised before the log: one log per row, none for the constant
  double* part = scr;
  double* E = scr + IIF_LOO_PARTS * N;
  double acc = 0.0;
  for (int ch = 0; ch < L.nchunks; ++ch) {
    LooThread T = T0;
    if (ch > 0) {
      T = loo_thread(L, seg, i, ch);
      __syncthreads();  // the previous chunk's tile is no longer read
    }
    if (T.cnt > 0) acc += loo_rows<U, CIRC>(x2 + T.xoff, xi, sc, tab, E + T.soff, N, T.cnt, T.cnts, T.wrap);
    IIF_PHASE(0);
    __syncthreads();
    IIF_PHASE(1);
    if (T.cnts > 0) acc += loo_gather(E + T.goff, N, T.cnts);
    IIF_PHASE(2);
  }
  if (L.G > 1) {
    if (seg < L.G) part[seg * N + i] = acc;
    __syncthreads();
    IIF_PHASE(3);
    if (threadIdx.x < N) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int g = 0; g < IIF_LOO_PARTS; g += 2) {
        if (g < L.G) s0 += part[g * N + threadIdx.x];
        if (g + 1 < L.G) s1 += part[(g + 1) * N + threadIdx.x];
      }
      acc = s0 + s1;
    }
  }
  // rows live in the first ceil(N/32) warps: sum their log terms
  double* buf = red + parity * (IIF_MAX_WARPS * IIF_RED_KMAX);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < L.nrw) {
    double term = (threadIdx.x < N) ? log(acc * norm) : 0.0;
    term = warp_sum(term);
    if (lane == 0) buf[warp] = term;
  }
  IIF_PHASE(4);
  __syncthreads();  // also protects scr for the next evaluation
  IIF_PHASE(5);
  double t0 = 0.0, t1 = 0.0;
#pragma unroll
  for (int w = 0; w < IIF_MAX_POINTS / 32; w += 2) {
    if (w < L.nrw) t0 += buf[w];
    if (w + 1 < L.nrw) t1 += buf[w + 1];
  }
  parity ^= 1;
  IIF_PHASE(6);
  IIF_PHASE(30);
  IIF_PHASE(31);
  return (t0 + t1) * negInvN;
}

// ------------------------------------------------------------------------------------------
// Golden-section searches, sequential and cluster-speculative.
//
// A search is ~18 strictly sequential objective evaluations, so a lone CTA pays 18 evaluation latencies.
// In narrow launches (tree top: fewer beliefs than SMs / 3) every belief is given a thread-block CLUSTER of
// three CTAs that run the kernel redundantly (same inputs, same Philox streams => identical state) and split
// the search speculatively: rank 0 evaluates the next point, ranks 1 and 2 the two points the step after
// could ask for (they depend only on the outcome of one comparison).  One value per CTA is exchanged per
// round through distributed shared memory + a cluster barrier, and every CTA advances the search by two
// steps.  The evaluated-and-used points and all comparisons are those of the sequential search, so the
// result is bit-identical; only rank 0 writes outputs.
// ------------------------------------------------------------------------------------------
#define IIF_SPEC_CLUSTER 3
struct ClusterCtx {
  int C, rank;   // cluster size (1 or IIF_SPEC_CLUSTER) and this CTA's rank
  double* fx;    // shared exchange buffer [2][8]
  int xpar;      // round parity (double buffering)
};

__device__ __forceinline__ void cluster_exchange(ClusterCtx& cc, double f, double (&F)[IIF_SPEC_CLUSTER]) {
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0) {
    for (int p = 0; p < cc.C; ++p) {
      double* remote = cl.map_shared_rank(cc.fx, p);
      remote[cc.xpar * 8 + cc.rank] = f;
    }
  }
  cl.sync();
#pragma unroll
  for (int k = 0; k < IIF_SPEC_CLUSTER; ++k) F[k] = cc.fx[cc.xpar * 8 + k];
  cc.xpar ^= 1;
}

// Numerical-Recipes golden section as used by KDE `golden(npd, nLOO_LL, ax, bx, cx, tol)`; the search
// variable scales the base bandwidth h0.  Uniform control flow across the CTA (and the cluster); the
// objective has ONE call site per variant so that it is inlined once.
struct GoldenNR {
  double x0, x1, x2, x3, f1, f2;
};
__device__ __forceinline__ bool gnr_done(const GoldenNR& s, double tol) {
  return !(fabs(s.x3 - s.x0) > tol * (fabs(s.x1) + fabs(s.x2)));
}
__device__ __forceinline__ double gnr_advance(GoldenNR& s, bool right) {  // new evaluation point
  const double C = (3.0 - sqrt(5.0)) / 2.0, R = 1.0 - C;
  if (right) { s.x0 = s.x1; s.x1 = s.x2; s.x2 = R * s.x1 + C * s.x3; s.f1 = s.f2; return s.x2; }
  s.x3 = s.x2; s.x2 = s.x1; s.x1 = R * s.x2 + C * s.x0; s.f2 = s.f1;
  return s.x1;
}
__device__ __forceinline__ void gnr_set(GoldenNR& s, bool right, double fn) {
  if (right) s.f2 = fn; else s.f1 = fn;
}

template <int U>
__device__ __forceinline__ double golden_nr(const double* x2, const LooCfg& L, const LooThread& T0, int seg, int i,
                                            double h0, double ax, double bx, double cx, double tol,
                                            const double* tab, double* scr, double* red, int& parity,
                                            ClusterCtx& cc) {
  const double C = (3.0 - sqrt(5.0)) / 2.0;
  const double xi = x2[i];
  const double invK = 1.0 / ((double)(L.N - 1) * sqrt(IIF_TWO_PI)), negInvN = -1.0 / (double)L.N;
  GoldenNR s;
  s.x0 = ax; s.x3 = cx; s.f1 = 0.0; s.f2 = 0.0;
  if (fabs(cx - bx) > fabs(bx - ax)) { s.x1 = bx; s.x2 = bx + C * (cx - bx); }
  else { s.x2 = bx; s.x1 = bx - C * (bx - ax); }
  if (cc.C < IIF_SPEC_CLUSTER) {  // ---- sequential (iterations -2 and -1 are the two initial evaluations)
    bool right = false;
    for (int it = -2; it < 200; ++it) {
      double xn;
      if (it >= 0) {
        if (gnr_done(s, tol)) break;
        right = s.f2 < s.f1;
        xn = gnr_advance(s, right);
      } else {
        xn = (it == -2) ? s.x1 : s.x2;
      }
      const double fn = loo_nll<U, false>(x2, L, T0, seg, i, xi, xn * h0, invK, negInvN, tab, scr, red, parity);
      if (it == -2) s.f1 = fn;
      else if (it == -1) s.f2 = fn;
      else gnr_set(s, right, fn);
    }
  } else {  // ---- cluster-speculative: two steps per round
    bool right = false, stop = false;
    GoldenNR sA = s;
    for (int round = -1; round < 100 && !stop; ++round) {
      double xe;
      if (round < 0) {
        xe = (cc.rank == 0) ? s.x1 : s.x2;
      } else {
        if (gnr_done(s, tol)) break;
        right = s.f2 < s.f1;
        sA = s;
        const double xA = gnr_advance(sA, right);
        GoldenNR sT = sA, sF = sA;
        const double xT = gnr_advance(sT, true), xF = gnr_advance(sF, false);
        xe = (cc.rank == 0) ? xA : (cc.rank == 1) ? xT : xF;
      }
      const double fn = loo_nll<U, false>(x2, L, T0, seg, i, xi, xe * h0, invK, negInvN, tab, scr, red, parity);
      double F[IIF_SPEC_CLUSTER];
      cluster_exchange(cc, fn, F);
      if (round < 0) { s.f1 = F[0]; s.f2 = F[1]; continue; }
      s = sA;
      gnr_set(s, right, F[0]);
      if (gnr_done(s, tol)) { stop = true; continue; }
      const bool right2 = s.f2 < s.f1;
      gnr_advance(s, right2);
      gnr_set(s, right2, right2 ? F[1] : F[2]);
    }
  }
  return (s.f1 < s.f2) ? s.x1 : s.x2;
}

// Optim.jl GoldenSection on [lo, hi] as used by AMP kde!_CircularNaiveCV
struct GoldenOptim {
  double lo, hi, xm, fm;
};
__device__ __forceinline__ bool gop_done(const GoldenOptim& s, double rel_tol) {
  const double abs_tol = 2.220446049250313e-16;
  const double tolx = rel_tol * fabs(s.xm) + abs_tol;
  const double mid = 0.5 * (s.hi + s.lo);
  return fabs(s.xm - mid) <= 2 * tolx - 0.5 * (s.hi - s.lo);
}
__device__ __forceinline__ double gop_next(const GoldenOptim& s, bool& up) {
  const double gr = 0.5 * (3.0 - sqrt(5.0));
  up = s.hi - s.xm > s.xm - s.lo;
  return up ? s.xm + gr * (s.hi - s.xm) : s.xm - gr * (s.xm - s.lo);
}
__device__ __forceinline__ void gop_apply(GoldenOptim& s, bool up, double xn, bool less, double fn) {
  if (up) {
    if (less) { s.lo = s.xm; s.xm = xn; s.fm = fn; } else s.hi = xn;
  } else {
    if (less) { s.hi = s.xm; s.xm = xn; s.fm = fn; } else s.lo = xn;
  }
}

template <int U>
__device__ __forceinline__ double golden_optim(const double* x2, const LooCfg& L, const LooThread& T0, int seg, int i,
                                               double lo, double hi, double rel_tol, const double* tab, double* scr,
                                               double* red, int& parity, ClusterCtx& cc) {
  const double gr = 0.5 * (3.0 - sqrt(5.0));
  const double xi = x2[i];
  const double invK = 1.0 / ((double)(L.N - 1) * sqrt(IIF_TWO_PI)), negInvN = -1.0 / (double)L.N;
  GoldenOptim s;
  s.lo = lo; s.hi = hi; s.xm = lo + gr * (hi - lo); s.fm = 0.0;
  if (cc.C < IIF_SPEC_CLUSTER) {  // ---- sequential (iteration -1 is the initial evaluation)
    bool up = false;
    for (int it = -1; it < 200; ++it) {
      double xn = s.xm;
      if (it >= 0) {
        if (gop_done(s, rel_tol)) break;
        xn = gop_next(s, up);
      }
      const double fn = loo_nll<U, true>(x2, L, T0, seg, i, xi, xn, invK, negInvN, tab, scr, red, parity);
      if (it < 0) s.fm = fn;
      else gop_apply(s, up, xn, fn < s.fm, fn);
    }
  } else {  // ---- cluster-speculative: the step after next depends only on whether fn < fm
    bool up = false, stop = false;
    double xA = s.xm, xT = s.xm, xF = s.xm;
    bool upT = false, upF = false;
    for (int round = -1; round < 100 && !stop; ++round) {
      double xe = s.xm;
      if (round >= 0) {
        if (gop_done(s, rel_tol)) break;
        xA = gop_next(s, up);
        GoldenOptim sT = s, sF = s;
        gop_apply(sT, up, xA, true, 0.0);
        gop_apply(sF, up, xA, false, 0.0);
        xT = gop_next(sT, upT);
        xF = gop_next(sF, upF);
        xe = (cc.rank == 0) ? xA : (cc.rank == 1) ? xT : xF;
      }
      const double fn = loo_nll<U, true>(x2, L, T0, seg, i, xi, xe, invK, negInvN, tab, scr, red, parity);
      double F[IIF_SPEC_CLUSTER];
      cluster_exchange(cc, fn, F);
      if (round < 0) { s.fm = F[0]; continue; }
      const bool less = F[0] < s.fm;
      gop_apply(s, up, xA, less, F[0]);
      if (gop_done(s, rel_tol)) { stop = true; continue; }
      if (less) gop_apply(s, upT, xT, F[1] < s.fm, F[1]);
      else gop_apply(s, upF, xF, F[2] < s.fm, F[2]);
    }
  }
  return s.xm;
}

// bandwidth of one coordinate; xa = the N coordinates, xb = loo_x2_doubles(N) scratch
template <int U>
__device__ __forceinline__ double coord_bandwidth(bool circ, const LooCfg& L, const LooThread& T0, int seg, int i,
                                                  const TreeStruct& T, const double* xa, double* xb,
                                                  const double* tab, double* scr, double* red, int& parity,
                                                  ClusterCtx& cc) {
  const int N = L.N;
  if (circ) {
    for (int m = threadIdx.x; m < 2 * N + IIF_LOO_XPAD; m += IIF_NT) xb[m] = xa[m % N];
    __syncthreads();
    return golden_optim<U>(xb, L, T0, seg, i, 1e-3, IIF_TWO_PI, 1e-3, tab, scr, red, parity, cc);
  }
  // rank sort (ties by index) -> xb ascending, doubled
  for (int m = threadIdx.x; m < N; m += IIF_NT) {
    const double xi = xa[m];
    int r = 0;
    for (int k = 0; k < N; ++k) {
      const double xk = xa[k];
      r += (xk < xi) || (xk == xi && k < m);
    }
    xb[r] = xi;
    xb[r + N] = xi;
    if (r < IIF_LOO_XPAD) xb[r + 2 * N] = xi;
  }
  __syncthreads();
  // KDE neighborMinMax: root ball diameter and smallest internal ball diameter (>= 1e-6)
  const double maxm = xb[N - 1] - xb[0];
  double m = maxm;
  for (int z = threadIdx.x; z < T.nn; z += IIF_NT) {
    const int lo = T.lo[z], hi = T.hi[z];
    if (hi > lo) m = fmin(m, xb[hi] - xb[lo]);
  }
  double minm = block_min1(m, red, parity);
  if (minm < 1e-6) minm = 1e-6;
  const double h0 = 0.5 * (minm + maxm);
  const double a = golden_nr<U>(xb, L, T0, seg, i, h0, 2.0 * minm / (minm + maxm), 1.0, 2.0 * maxm / (minm + maxm),
                                1e-2, tab, scr, red, parity, cc);
  return a * h0;
}

// Per-dimension bandwidth of the N x d points in `pts` (shared or global memory).
// xa: N doubles, xb: loo_x2_doubles(N) doubles, scr: loo_scratch_doubles(N) doubles of shared scratch.
// Result bw[c] is returned to every thread.  KTAG gives every kernel its own instantiation.
template <int KTAG>
__device__ __noinline__ void block_kde_bandwidth(const double* pts, int N, int d, int32_t circ_mask,
                                                 const TreeStruct* Tp, double* xa, double* xb, double* scr,
                                                 double* red, int* parity_io, double* bw) {
  const TreeStruct T = *Tp;
  const LooCfg L = loo_config(N);
  const int seg = threadIdx.x / N, i = threadIdx.x - seg * N;
  const LooThread T0 = loo_thread(L, seg, i, 0);
  int parity = *parity_io;
  __shared__ double tab[IIF_GTAB_N];  // 2^(j/256), see gauss_negU
  __shared__ double fx[16];   // cluster exchange buffer (speculative search)
  gauss_stage_table(tab);
  ClusterCtx cc;
  {
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    cc.C = (int)cl.num_blocks();
    cc.rank = (int)cl.block_rank();
    cc.fx = fx;
    cc.xpar = 0;
  }
  for (int c = 0; c < d; ++c) {
    __syncthreads();
    for (int m = threadIdx.x; m < N; m += IIF_NT) xa[m] = pts[m * d + c];
    __syncthreads();
    const bool circ = is_circ(circ_mask, c);
    bw[c] = (L.U == 4)   ? coord_bandwidth<4>(circ, L, T0, seg, i, T, xa, xb, tab, scr, red, parity, cc)
            : (L.U == 5) ? coord_bandwidth<5>(circ, L, T0, seg, i, T, xa, xb, tab, scr, red, parity, cc)
                         : coord_bandwidth<6>(circ, L, T0, seg, i, T, xa, xb, tab, scr, red, parity, cc);
  }
  *parity_io = parity;
}
