/*
 * iif_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See iif_oracle.h for
 * the parity status: a14 (KDE bandwidth) and a15 (KDE product) are **PARITY UNPINNED**
 * restatements of published algorithms (their Julia source is not under /root/reference).
 *
 * Every function cites the reference file:line (IncrementalInference.jl v0.35.6) it follows.
 * Scalar loops, one thread, no cleverness: this is the thing the CUDA path is checked against.
 */
#include "iif_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846
#define TWO_PI 6.28318530717958647692

static int64_t g_conv_count = 0;
int64_t iifo_conv_count(void) { return g_conv_count; }
void iifo_reset_counters(void) { g_conv_count = 0; }

/* ------------------------------------------------------------------------------------ */
/* Random streams: Philox4x32-10 (Salmon et al., SC'11), counter = (idx, stream, call, tag) */
/* ------------------------------------------------------------------------------------ */
void iifo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                     uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void rs_pair(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx, double* ua,
                    double* ub) {
  uint32_t o[4];
  iifo_philox4x32(idx, stream, call, 0x1F1B200u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  uint64_t a = ((uint64_t)o[1] << 32) | o[0];
  uint64_t b = ((uint64_t)o[3] << 32) | o[2];
  *ua = (double)(a >> 11) * (1.0 / 9007199254740992.0);
  *ub = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

double iifo_uniform(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx) {
  double a, b;
  rs_pair(seed, call, stream, idx, &a, &b);
  return a;
}

/* Box-Muller on the two uniforms of one Philox block */
double iifo_normal(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx) {
  double a, b;
  rs_pair(seed, call, stream, idx, &a, &b);
  return sqrt(-2.0 * log(1.0 - a)) * cos(TWO_PI * b);
}

/* ------------------------------------------------------------------------------------ */
/* Manifold helpers: products of TranslationGroup(1) / RealCircleGroup coordinates        */
/* ------------------------------------------------------------------------------------ */
/* Manifolds.sym_rem: wrap to [-pi, pi) */
static double wrap_pi(double a) {
  double r = a - TWO_PI * floor((a + PI) / TWO_PI);
  if (r >= PI) r -= TWO_PI;
  if (r < -PI) r += TWO_PI;
  return r;
}
static inline int is_circ(int32_t mask, int c) { return (mask >> c) & 1; }
static inline double mdiff(double a, double b, int circ) { /* log_b(a) coordinate */
  return circ ? wrap_pi(a - b) : a - b;
}
static inline double madd(double a, double t, int circ) { /* exp_a(t) coordinate */
  return circ ? wrap_pi(a + t) : a + t;
}

int64_t iifo_layout(int32_t nslots, iif_slot_desc* slots) {
  int64_t off = 0;
  for (int s = 0; s < nslots; ++s) {
    slots[s].pts_off = (int32_t)off;
    off += (int64_t)slots[s].cap * slots[s].dim;
  }
  return off;
}

/* mean(M, pts, GeodesicInterpolation()) — Manifolds.jl sequential geodesic interpolation,
 * called from calcStdBasicSpread (src/services/VariableStatistics.jl:30). */
/* SpecialOrthogonal(3) (test/testSpecialOrthogonalMani.jl:75-140): points as rotation vectors omega = vee(log(eps, R));
 * group operations through unit quaternions.  AMP treats the coordinates as (:Euclid, :Euclid, :Euclid) (:80-81). */
static inline int is_so3(int32_t mask) { return (mask & IIF_MANI_SO3) != 0; }
typedef struct { double w, x, y, z; } quat;
static quat so3_quat(const double* om) {
  double t2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], s, w;
  if (t2 < 1e-16) { s = 0.5 - t2 / 48.0; w = 1.0 - t2 / 8.0; }
  else { double t = sqrt(t2); s = sin(0.5 * t) / t; w = cos(0.5 * t); }
  quat q = {w, s * om[0], s * om[1], s * om[2]};
  return q;
}
static quat quat_mul(quat a, quat b) {
  quat r = {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
  return r;
}
static void so3_rotvec(quat q, double* om) {
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  double vn = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  double k = vn < 1e-12 ? 2.0 : 2.0 * atan2(vn, q.w) / vn;
  om[0] = k * q.x; om[1] = k * q.y; om[2] = k * q.z;
}
static void so3_compose(const double* a, const double* b, double* out) { /* Log(Exp(a) Exp(b)) = p Exp(X) */
  so3_rotvec(quat_mul(so3_quat(a), so3_quat(b)), out);
}
static void so3_between(const double* a, const double* b, double* out) { /* Log(Exp(a)^T Exp(b)) = vee(log(p, q)) */
  quat qa = so3_quat(a);
  qa.x = -qa.x; qa.y = -qa.y; qa.z = -qa.z;
  so3_rotvec(quat_mul(qa, so3_quat(b)), out);
}

static void geodesic_mean(const double* pts, int n, int d, int32_t cm, double* mu) {
  if (is_so3(cm)) { /* mu <- mu Exp(Log(mu^T p_i) / (i + 1)) */
    for (int c = 0; c < 3; ++c) mu[c] = n > 0 ? pts[c] : 0.0;
    for (int i = 1; i < n; ++i) {
      double dl[3], nx[3], t = 1.0 / (double)(i + 1);
      so3_between(mu, pts + i * 3, dl);
      for (int c = 0; c < 3; ++c) dl[c] *= t;
      so3_compose(mu, dl, nx);
      for (int c = 0; c < 3; ++c) mu[c] = nx[c];
    }
    return;
  }
  for (int c = 0; c < d; ++c) mu[c] = n > 0 ? pts[c] : 0.0;
  for (int i = 1; i < n; ++i) {
    double t = 1.0 / (double)(i + 1);
    for (int c = 0; c < d; ++c) {
      double v = mdiff(pts[i * d + c], mu[c], is_circ(cm, c));
      mu[c] = madd(mu[c], t * v, is_circ(cm, c));
    }
  }
}

/* mean(M, pts) default estimator: arithmetic (Euclid) / extrinsic atan2 (Circle) */
static void default_mean(const double* pts, int n, int d, int32_t cm, double* mu) {
  for (int c = 0; c < d; ++c) {
    if (is_circ(cm, c)) {
      double s = 0, k = 0;
      for (int i = 0; i < n; ++i) { s += sin(pts[i * d + c]); k += cos(pts[i * d + c]); }
      mu[c] = n > 0 ? atan2(s, k) : 0.0;
    } else {
      double s = 0;
      for (int i = 0; i < n; ++i) s += pts[i * d + c];
      mu[c] = n > 0 ? s / n : 0.0;
    }
  }
}

/* calcStdBasicSpread — src/services/VariableStatistics.jl:22-36 */
double iifo_std_basic_spread(const double* pts, int32_t n, int32_t d, int32_t cm) {
  if (n < 2) return 1.0; /* std of <2 points is NaN in Julia => `1e-10 < NaN` false => 1.0 */
  double mu[IIF_MAX_DIM];
  geodesic_mean(pts, n, d, cm, mu);
  double acc = 0;
  if (is_so3(cm)) { /* distance(M, mu, p)^2 = |log(mu, p)|_F^2 = 2 angle^2: ASSUMPTION (Manifolds' Frobenius metric on the
                       skew matrices); only scales the entropy spread, never the roots */
    for (int i = 0; i < n; ++i) {
      double dl[3];
      so3_between(mu, pts + i * 3, dl);
      acc += 2.0 * (dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    }
  } else {
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < d; ++c) {
        double v = mdiff(pts[i * d + c], mu[c], is_circ(cm, c));
        acc += v * v;
      }
  }
  double sigma = sqrt(acc / (double)(n - 1));
  return (1e-10 < sigma) ? sigma : 1.0;
}

/* ------------------------------------------------------------------------------------ */
/* a6: _prepareHypoRecipe!  — src/services/ExplicitDiscreteMarginalizations.jl:142-289    */
/* ------------------------------------------------------------------------------------ */
static int in_list(const int32_t* l, int n, int v) {
  for (int i = 0; i < n; ++i)
    if (l[i] == v) return 1;
  return 0;
}

static int32_t categorical(const double* p, int np, double u) { /* 1-based label */
  double c = 0;
  int last = 1;
  for (int k = 0; k < np; ++k) {
    if (p[k] > 0) last = k + 1;
    c += p[k];
    if (u < c) return k + 1;
  }
  return last;
}

int32_t iifo_hypo_recipe(const double* mh, int32_t lenXi, int32_t maxlen, int32_t sfidx,
                         const int32_t* isinit, double nullhypo, const double* u,
                         const int32_t* mhidx_in, int32_t* mhidx, int32_t* nbuckets,
                         int32_t* bucket_hypo, int32_t* bucket_nvars, int32_t* bucket_vars,
                         int32_t* certain, int32_t* ncertain) {
  if (lenXi < 1 || lenXi > IIF_MAX_ARITY || sfidx < 1 || sfidx > lenXi) return IIF_ERR_ARG;
  if (mh == NULL) {
    /* `Nothing` method, :234-289 */
    for (int n = 0; n < maxlen; ++n) {
      if (mhidx_in) mhidx[n] = mhidx_in[n];
      else if (nullhypo == 0) mhidx[n] = 1;                 /* :261 ones(Int, maxlen) */
      else {
        double p[2] = {nullhypo, 1.0 - nullhypo};           /* :254-255 */
        mhidx[n] = categorical(p, 2, u[n]) - 1;             /* :261 rand(nhh) .- 1 */
      }
    }
    *ncertain = lenXi;
    for (int i = 0; i < lenXi; ++i) certain[i] = i + 1;      /* :264 certainidx = 1:lenXi */
    *nbuckets = lenXi + 1;
    for (int b = 0; b <= lenXi; ++b) {                       /* :269-283 */
      bucket_hypo[b] = b;
      if (b == 0) { bucket_nvars[b] = 1; bucket_vars[b * IIF_MAX_ARITY] = sfidx; }
      else if (b == 1) {
        bucket_nvars[b] = lenXi;
        for (int i = 0; i < lenXi; ++i) bucket_vars[b * IIF_MAX_ARITY + i] = i + 1;
      } else bucket_nvars[b] = 0;
    }
    return IIF_OK;
  }
  /* Categorical method, :142-232 */
  int32_t uncertn[IIF_MAX_ARITY];
  int nc = 0, nu = 0;
  for (int i = 0; i < lenXi; ++i) {                          /* getHypothesesVectors :17-24 */
    if (mh[i] == 0.0) certain[nc++] = i + 1;
    else if (mh[i] > 0.0) uncertn[nu++] = i + 1;
  }
  *ncertain = nc;
  double p[IIF_MAX_ARITY + 1];
  int np = lenXi;
  for (int i = 0; i < lenXi; ++i) p[i] = mh[i];
  int ninit = 0;
  for (int i = 0; i < lenXi; ++i) ninit += isinit ? (isinit[i] != 0) : 1;
  if (ninit < lenXi - 1) {                                   /* :161-172 */
    double s = 0;
    for (int i = 0; i < lenXi; ++i) {
      int suppress = isinit && !isinit[i] && (i + 1 != sfidx);
      if (suppress) p[i] = 0.0;
      s += p[i];
    }
    for (int i = 0; i < lenXi; ++i) p[i] /= s;
  }
  int sf_uncertain = in_list(uncertn, nu, sfidx);
  if (sf_uncertain) {                                        /* :176-183 */
    double nhw = (double)(nu + 1);
    double q[IIF_MAX_ARITY + 1];
    q[0] = 1.0 / nhw;
    double s = q[0];
    for (int i = 0; i < lenXi; ++i) { q[i + 1] = (double)nu / nhw * p[i]; s += q[i + 1]; }
    for (int i = 0; i <= lenXi; ++i) p[i] = q[i] / s;
    np = lenXi + 1;
  }
  for (int n = 0; n < maxlen; ++n) {                         /* :186-192 */
    if (mhidx_in) mhidx[n] = mhidx_in[n];
    else mhidx[n] = categorical(p, np, u[n]) - (sf_uncertain ? 1 : 0);
  }
  int pidx = sf_uncertain ? -1 : 0;
  int sfincer = in_list(certain, nc, sfidx);
  int nb = 0;
  for (int k = 0; k < np; ++k) {                             /* :195-224 */
    pidx += 1;
    int pidxincer = in_list(certain, nc, pidx);
    int32_t* vars = bucket_vars + nb * IIF_MAX_ARITY;
    int nv = 0;
    if (!pidxincer && sfincer && pidx != 0) {                /* :201-203 sort(union(certainidx,pidx)) */
      for (int v = 1; v <= lenXi; ++v)
        if (in_list(certain, nc, v) || v == pidx) vars[nv++] = v;
    } else if (((pidxincer && !sfincer) || sfidx == pidx) && pidx != 0) { /* :205-207 */
      for (int v = 1; v <= lenXi; ++v)
        if (in_list(certain, nc, v) || v == sfidx) vars[nv++] = v;
    } else if (pidxincer && sfincer && pidx != 0) {          /* :209-211 both certain: empty */
      nv = 0;
    } else if (!pidxincer && !sfincer && pidx != 0) {        /* :212-213 iterah = uncertnidx */
      for (int i = 0; i < nu; ++i) vars[nv++] = uncertn[i];
    } else if (pidx == 0) {                                  /* :214-216 */
      vars[nv++] = sfidx;
    } else return IIF_ERR_ARG;
    bucket_hypo[nb] = pidx;
    bucket_nvars[nb] = nv;
    nb++;
  }
  *nbuckets = nb;
  return IIF_OK;
}

/* ------------------------------------------------------------------------------------ */
/* a13: built-in residuals  — src/Factors/ *.jl                                           */
/* ------------------------------------------------------------------------------------ */
int32_t iifo_residual(int32_t kind, int32_t d, int32_t cm, int32_t zdim, const double* z,
                      int32_t arity, const double* x, double* res) {
  switch (kind) {
    case IIF_F_PRIOR:         /* DefaultPrior.jl:17  z .- x1 */
    case IIF_F_MSG_PRIOR:     /* MsgPrior.jl:36      z .- x1 */
      for (int c = 0; c < zdim; ++c) res[c] = z[c] - x[c];
      return IIF_OK;
    case IIF_F_PRIOR_CIRCULAR: /* Circular.jl:70-74  vee(log(M, p, m)) */
      for (int c = 0; c < zdim; ++c) res[c] = wrap_pi(z[c] - x[c]);
      return IIF_OK;
    case IIF_F_LINEAR_RELATIVE: /* LinearRelative.jl:42-49  z .- (x2 .- x1) */
      if (arity != 2) return IIF_ERR_ARG;
      for (int c = 0; c < zdim; ++c) res[c] = z[c] - (x[d + c] - x[c]);
      return IIF_OK;
    case IIF_F_CIRCULAR_CIRCULAR: /* Circular.jl:24-28 -> GenericFunctions.jl:47-52
                                     qhat = exp(M,p,X); vee(log(M,q,qhat)) */
      if (arity != 2) return IIF_ERR_ARG;
      for (int c = 0; c < zdim; ++c) {
        double qhat = madd(x[c], z[c], is_circ(cm, c));
        res[c] = mdiff(qhat, x[d + c], is_circ(cm, c));
      }
      return IIF_OK;
    case IIF_F_EUCLID_DISTANCE: { /* EuclidDistance.jl:20  z .- norm(x2 .- x1) */
      if (arity != 2) return IIF_ERR_ARG;
      double s = 0;
      for (int c = 0; c < d; ++c) s += (x[d + c] - x[c]) * (x[d + c] - x[c]);
      res[0] = z[0] - sqrt(s);
      return IIF_OK;
    }
    case IIF_F_MANIFOLD_PRIOR: /* GenericFunctions.jl:209-214  vee(M, p, log(M, p, m)) in (Euclid.., angle..) coords */
      for (int c = 0; c < zdim; ++c) res[c] = mdiff(z[c], x[c], is_circ(cm, c));
      return IIF_OK;
    case IIF_F_SO3_RELATIVE: { /* qhat = p Exp(X); vee(log(q, qhat)) = Log(q^T qhat) */
      if (arity != 2 || d != 3) return IIF_ERR_ARG;
      double qh[3];
      so3_compose(x, z, qh);
      so3_between(x + d, qh, res);
      return IIF_OK;
    }
    case IIF_F_SE2_RELATIVE: { /* GenericFunctions.jl:39-44: qhat = compose(p, exp(M, eps, X)); vee(M, q, log(M, q, qhat));
                                  hybrid tangent representation (testSpecialEuclidean2Mani.jl:14): exp(eps, X) = (X_t, R(X_th)),
                                  log(q, qhat) = (t_qhat - t_q, th_qhat - th_q) */
      if (arity != 2 || d != 3) return IIF_ERR_ARG;
      const double* p = x; const double* q = x + d;
      double sn = sin(p[2]), cs = cos(p[2]);
      double qh0 = p[0] + cs * z[0] - sn * z[1], qh1 = p[1] + sn * z[0] + cs * z[1];
      res[0] = qh0 - q[0];
      res[1] = qh1 - q[1];
      res[2] = wrap_pi(wrap_pi(p[2] + z[2]) - q[2]);
      return IIF_OK;
    }
    default: return IIF_ERR_UNSUPPORTED;
  }
}

/* Per-sample solve  — _solveCCWNumeric! / _solveLambdaNumeric
 * (src/services/NumericalCalculations.jl:413-452, :90-133): argmin over the solve-for point of
 * sum(res.^2), started at u0.  For the built-ins with a unique root the minimiser is the
 * analytic root; EuclidDistance has a ring of roots and a descent method started
 * at u0 moves radially, so its minimiser is the radial projection of u0 onto the ring.
 * `xa`,`xb`: the two active points in factor order; `sf_second`: solving for the 2nd. */
static void solve_binary(int kind, int d, int32_t cm, const double* z, const double* other,
                         int sf_second, const double* u0, double* out) {
  if (kind == IIF_F_LINEAR_RELATIVE || kind == IIF_F_CIRCULAR_CIRCULAR) {
    for (int c = 0; c < d; ++c) {
      int circ = is_circ(cm, c);
      out[c] = sf_second ? madd(other[c], z[c], circ) : madd(other[c], -z[c], circ);
    }
  } else if (kind == IIF_F_SO3_RELATIVE) { /* q = p Exp(X); p = q Exp(-X) */
    if (sf_second) so3_compose(other, z, out);
    else { double nz[3] = {-z[0], -z[1], -z[2]}; so3_compose(other, nz, out); }
  } else if (kind == IIF_F_SE2_RELATIVE) {
    /* unique root of the residual above: solving q: q = p o (X_t, R(X_th)); solving p: th_p = th_q - X_th,
     * t_p = t_q - R(th_p) X_t */
    double th = sf_second ? other[2] : wrap_pi(other[2] - z[2]);
    double sn = sin(th), cs = cos(th);
    double rx = cs * z[0] - sn * z[1], ry = sn * z[0] + cs * z[1];
    if (sf_second) { out[0] = other[0] + rx; out[1] = other[1] + ry; out[2] = wrap_pi(other[2] + z[2]); }
    else { out[0] = other[0] - rx; out[1] = other[1] - ry; out[2] = th; }
  } else { /* IIF_F_EUCLID_DISTANCE */
    double dir[IIF_MAX_DIM], nrm = 0;
    for (int c = 0; c < d; ++c) { dir[c] = u0[c] - other[c]; nrm += dir[c] * dir[c]; }
    nrm = sqrt(nrm);
    double r = z[0] > 0 ? z[0] : 0.0;
    for (int c = 0; c < d; ++c) {
      double unit = nrm > 0 ? dir[c] / nrm : (c == 0 ? 1.0 : 0.0);
      out[c] = other[c] + r * unit;
    }
  }
}

/* sum(res.^2) of a binary relative factor with the solve-for point at x (CalcFactorNormSq, NumericalCalculations.jl:68-72) */
static double cost_binary(int kind, int d, int32_t cm, int zdim, const double* z, const double* x, const double* other,
                          int sf_second) {
  double xx[2 * IIF_MAX_DIM], res[IIF_MAX_DIM] = {0};
  for (int c = 0; c < d; ++c) { xx[c] = sf_second ? other[c] : x[c]; xx[d + c] = sf_second ? x[c] : other[c]; }
  iifo_residual(kind, d, cm, zdim, z, 2, xx, res);
  double s = 0;
  for (int c = 0; c < zdim; ++c) s += res[c] * res[c];
  return s;
}

/* Optim.NelderMead as _solveLambdaNumeric configures it (NumericalCalculations.jl:49-72, :90-133): default
 * AdaptiveParameters (alpha 1, beta 1 + 2/n, gamma 0.75 - 1/2n, delta 1 - 1/n), AffineSimplexer (vertex j+1 =
 * x0 with coordinate j -> 1.5 x0_j + 0.025), convergence sqrt(var(f_simplex) n/(n+1)) < g_abstol = 1e-8, at most 1000
 * iterations; the result is the better of the best vertex and the centroid of the n best.  Optim.jl is not vendored:
 * restated from its published algorithm (PARITY UNPINNED against Julia; GPU and oracle follow the same steps). */
static void nelder_mead_binary(int kind, int d, int32_t cm, int zdim, const double* z, const double* other,
                               int sf_second, const double* x0, double* out) {
  const int n = d, m = d + 1;
  const double alpha = 1.0, beta = 1.0 + 2.0 / n, gamma = 0.75 - 1.0 / (2.0 * n), delta = 1.0 - 1.0 / n;
  double S[IIF_MAX_DIM + 1][IIF_MAX_DIM], f[IIF_MAX_DIM + 1];
  int ord[IIF_MAX_DIM + 1];
  for (int i = 0; i < m; ++i) {
    for (int c = 0; c < n; ++c) S[i][c] = x0[c];
    if (i > 0) S[i][i - 1] = 1.5 * x0[i - 1] + 0.025;
    f[i] = cost_binary(kind, d, cm, zdim, z, S[i], other, sf_second);
  }
  double cen[IIF_MAX_DIM], xr[IIF_MAX_DIM], xe[IIF_MAX_DIM], xc[IIF_MAX_DIM];
  for (int it = 0; it < 1000; ++it) {
    for (int i = 0; i < m; ++i) ord[i] = i;                     /* sortperm (stable insertion sort) */
    for (int i = 1; i < m; ++i) {
      int k = ord[i], j = i - 1;
      while (j >= 0 && f[ord[j]] > f[k]) { ord[j + 1] = ord[j]; --j; }
      ord[j + 1] = k;
    }
    const int lo = ord[0], hi = ord[m - 1], sh = ord[m - 2];
    for (int c = 0; c < n; ++c) {
      double s = 0;
      for (int i = 0; i < m - 1; ++i) s += S[ord[i]][c];
      cen[c] = s / n;
    }
    for (int c = 0; c < n; ++c) xr[c] = cen[c] + alpha * (cen[c] - S[hi][c]);
    const double fr = cost_binary(kind, d, cm, zdim, z, xr, other, sf_second);
    int shrink = 0;
    if (fr < f[lo]) {
      for (int c = 0; c < n; ++c) xe[c] = cen[c] + beta * (xr[c] - cen[c]);
      const double fe = cost_binary(kind, d, cm, zdim, z, xe, other, sf_second);
      if (fe < fr) { for (int c = 0; c < n; ++c) S[hi][c] = xe[c]; f[hi] = fe; }
      else { for (int c = 0; c < n; ++c) S[hi][c] = xr[c]; f[hi] = fr; }
    } else if (fr < f[sh]) {
      for (int c = 0; c < n; ++c) S[hi][c] = xr[c];
      f[hi] = fr;
    } else if (fr < f[hi]) {                                   /* outside contraction */
      for (int c = 0; c < n; ++c) xc[c] = cen[c] + gamma * (xr[c] - cen[c]);
      const double fc = cost_binary(kind, d, cm, zdim, z, xc, other, sf_second);
      if (fc <= fr) { for (int c = 0; c < n; ++c) S[hi][c] = xc[c]; f[hi] = fc; } else shrink = 1;
    } else {                                                   /* inside contraction */
      for (int c = 0; c < n; ++c) xc[c] = cen[c] - gamma * (xr[c] - cen[c]);
      const double fc = cost_binary(kind, d, cm, zdim, z, xc, other, sf_second);
      if (fc < f[hi]) { for (int c = 0; c < n; ++c) S[hi][c] = xc[c]; f[hi] = fc; } else shrink = 1;
    }
    if (shrink)
      for (int i = 1; i < m; ++i) {
        const int k = ord[i];
        for (int c = 0; c < n; ++c) S[k][c] = S[lo][c] + delta * (S[k][c] - S[lo][c]);
        f[k] = cost_binary(kind, d, cm, zdim, z, S[k], other, sf_second);
      }
    double mean = 0, var = 0;
    for (int i = 0; i < m; ++i) mean += f[i];
    mean /= m;
    for (int i = 0; i < m; ++i) var += (f[i] - mean) * (f[i] - mean);
    if (sqrt(var / m) < 1e-8) break;                          /* sqrt(var_corrected * n / (n + 1)) */
  }
  int best = 0;
  for (int i = 1; i < m; ++i) if (f[i] < f[best]) best = i;
  int worst = 0;
  for (int i = 1; i < m; ++i) if (f[i] > f[worst]) worst = i;
  for (int c = 0; c < n; ++c) {
    double s = 0;
    for (int i = 0; i < m; ++i) if (i != worst) s += S[i][c];
    cen[c] = s / n;
  }
  const double fcen = cost_binary(kind, d, cm, zdim, z, cen, other, sf_second);
  const double* r = fcen < f[best] ? cen : S[best];
  for (int c = 0; c < n; ++c) out[c] = is_circ(cm, c) ? wrap_pi(r[c]) : r[c];     /* exp(M, eps, hat(minimizer)) */
  if (is_so3(cm)) { double zero[3] = {0, 0, 0}, pr[3]; so3_compose(out, zero, pr); for (int c = 0; c < 3; ++c) out[c] = pr[c]; }   /* principal rotation vector */
}

/* ------------------------------------------------------------------------------------ */
/* Measurement sampling — sampleFactor! SolverUtilities.jl:50-76, getSample                */
/* ManifoldSampling.jl:121-145, Mixture.jl:114-155, MsgPrior.jl:21-30, Circular.jl:30-33,62-68 */
/* ------------------------------------------------------------------------------------ */
static void sample_simple(int kind, int dim, const double* prm, uint64_t seed, uint32_t call,
                          int n, int zdim, double* z) {
  if (kind == IIF_D_NORMAL) {
    z[0] = prm[0] + prm[1] * iifo_normal(seed, call, IIF_RS_MEAS, (uint32_t)(n * zdim));
  } else if (kind == IIF_D_UNIFORM) {
    z[0] = prm[0] + (prm[1] - prm[0]) * iifo_uniform(seed, call, IIF_RS_MEAS, (uint32_t)(n * zdim));
  } else { /* MVNORMAL: mu + L*eps */
    double e[IIF_MAX_DIM];
    for (int c = 0; c < dim; ++c) e[c] = iifo_normal(seed, call, IIF_RS_MEAS, (uint32_t)(n * zdim + c));
    for (int r = 0; r < dim; ++r) {
      double acc = prm[r];
      for (int c = 0; c <= r; ++c) acc += prm[dim + r * dim + c] * e[c];
      z[r] = acc;
    }
  }
}

static int simple_block_len(int kind, int dim) {
  return kind == IIF_D_MVNORMAL ? dim + dim * dim : 2;
}

static int32_t sample_measurement(const iifo_graph* g, const iif_factor_desc* f, uint32_t call,
                                  int n, double* z) {
  const iif_dist_desc* D = &g->dists[f->dist];
  const double* prm = g->dparams + D->poff;
  uint64_t seed = g->sp.seed;
  switch (D->kind) {
    case IIF_D_NORMAL:
    case IIF_D_UNIFORM:
    case IIF_D_MVNORMAL: sample_simple(D->kind, D->dim, prm, seed, call, n, f->zdim, z); break;
    case IIF_D_MIXTURE: { /* Mixture.jl:137 labels = rand(diversity); :141-151 per-sample draw */
      double u = iifo_uniform(seed, call, IIF_RS_MIXLABEL, (uint32_t)n);
      int lbl = categorical(prm, D->ncomp, u) - 1;
      const double* cp = prm + D->ncomp + lbl * simple_block_len(D->comp_kind, D->dim);
      sample_simple(D->comp_kind, D->dim, cp, seed, call, n, f->zdim, z);
      break;
    }
    case IIF_D_SAMPLES: { /* rand(Z) of any host-side distribution, pre-drawn into a table; uniform resampling */
      double u = iifo_uniform(seed, call, IIF_RS_MIXLABEL, (uint32_t)n);
      int k = (int)(u * D->ncomp);
      if (k >= D->ncomp) k = D->ncomp - 1;
      for (int c = 0; c < D->dim; ++c) z[c] = prm[k * D->dim + c];
      break;
    }
    case IIF_D_KDE: { /* MsgPrior.jl:27-30 samplePoint(mkd): pick a kernel, add bw-scaled jitter */
      const iif_slot_desc* S = &g->slots[D->slot];
      int np = g->npts[D->slot];
      if (np <= 0) return IIF_ERR_STATE;
      double u = iifo_uniform(seed, call, IIF_RS_MIXLABEL, (uint32_t)n);
      int k = (int)(u * np);
      if (k >= np) k = np - 1;
      for (int c = 0; c < S->dim; ++c) {
        double e = iifo_normal(seed, call, IIF_RS_MEAS, (uint32_t)(n * f->zdim + c));
        z[c] = madd(g->pts[S->pts_off + k * S->dim + c], g->bw[D->slot * IIF_MAX_DIM + c] * e,
                    is_circ(S->circ_mask, c));
      }
      break;
    }
    default: return IIF_ERR_UNSUPPORTED;
  }
  return IIF_OK;
}

/* ------------------------------------------------------------------------------------ */
/* point estimates — calcPPE (src/services/FGOSUtils.jl:237-278): mean = calcMean(P),   */
/* max = getKDEMax(P).  getKDEMax / getKDERange live in KernelDensityEstimate.jl (un-vendored,     */
/* PARITY UNPINNED): per coordinate, the marginal KDE is evaluated on a 200-point grid over the    */
/* point range extended by 10 % on both sides and the first grid point of maximal density is taken. */
/* ------------------------------------------------------------------------------------ */
#define IIFO_PPE_GRID 200
int32_t iifo_ppe(const double* pts, int32_t n, int32_t d, int32_t cm, const double* bw,
                 double* mean_out, double* max_out) {
  if (n < 1 || d < 1 || d > IIF_MAX_DIM) return IIF_ERR_ARG;
  for (int c = 0; c < d; ++c) {
    const int circ = is_circ(cm, c);
    double s = 0, sn = 0, cs = 0, lo = pts[c], hi = pts[c];
    for (int i = 0; i < n; ++i) {
      double v = pts[i * d + c];
      if (circ) { sn += sin(v); cs += cos(v); } else s += v;
      if (v < lo) lo = v;
      if (v > hi) hi = v;
    }
    mean_out[c] = circ ? atan2(sn, cs) : s / (double)n;
    const double dr = 0.1 * (hi - lo);
    const double a = lo - dr, b = hi + dr, step = (b - a) / (double)(IIFO_PPE_GRID - 1);
    const double inv2h2 = 1.0 / (2.0 * bw[c] * bw[c]);
    double best = -1.0, xbest = a;
    for (int gi = 0; gi < IIFO_PPE_GRID; ++gi) {
      const double X = a + step * (double)gi;
      double y = 0;
      for (int i = 0; i < n; ++i) {
        double dl = mdiff(X, pts[i * d + c], circ);
        y += exp(-dl * dl * inv2h2);
      }
      if (y > best) { best = y; xbest = X; }
    }
    max_out[c] = circ ? wrap_pi(xbest) : xbest;
  }
  return IIF_OK;
}

/* ------------------------------------------------------------------------------------ */
/* a14: KDE bandwidth — AMP.manikde! -> getKDEManifoldBandwidths -> KDE.kde!(x) "lcv"      */
/* PARITY UNPINNED (upstream packages not vendored).                                       */
/* ------------------------------------------------------------------------------------ */
/* Negative average leave-one-out log likelihood, exact O(N^2) (KDE.setForceEvalDirect!(true),
 * src/IncrementalInference.jl:104):  -1/N sum_i log( 1/(N-1) sum_{j!=i} N(x_i - x_j; 0, h^2) ) */
double iifo_loo_nll(const double* x, int32_t n, int32_t circular, double h) {
  double inv2h2 = 1.0 / (2.0 * h * h);
  double lognorm = log((double)(n - 1) * sqrt(TWO_PI) * h);
  double acc = 0;
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int j = 0; j < n; ++j) {
      if (j == i) continue;
      double dlt = mdiff(x[i], x[j], circular);
      s += exp(-dlt * dlt * inv2h2);
    }
    acc += log(s) - lognorm;
  }
  return -acc / (double)n;
}

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* KDE neighborMinMax: maxm = root ball diameter, minm = smallest internal ball-tree node
 * diameter (>= 1e-6); 1-D median-split tree over the sorted points. */
static void node_min_diam(const double* xs, int lo, int hi, double* minm) {
  if (lo >= hi) return;
  double dm = xs[hi] - xs[lo];
  if (dm < *minm) *minm = dm;
  int mid = (lo + hi) / 2;
  node_min_diam(xs, lo, mid, minm);
  node_min_diam(xs, mid + 1, hi, minm);
}

/* Numerical-Recipes golden section as used by KDE `golden(npd, nLOO_LL, ax, bx, cx, tol)` */
static double golden_nr(const double* x, int n, double h0, double ax, double bx, double cx,
                        double tol) {
  const double C = (3.0 - sqrt(5.0)) / 2.0, R = 1.0 - C;
  double x0 = ax, x3 = cx, x1, x2;
  if (fabs(cx - bx) > fabs(bx - ax)) { x1 = bx; x2 = bx + C * (cx - bx); }
  else { x2 = bx; x1 = bx - C * (bx - ax); }
  double f1 = iifo_loo_nll(x, n, 0, x1 * h0), f2 = iifo_loo_nll(x, n, 0, x2 * h0);
  for (int it = 0; it < 200 && fabs(x3 - x0) > tol * (fabs(x1) + fabs(x2)); ++it) {
    if (f2 < f1) {
      x0 = x1; x1 = x2; x2 = R * x1 + C * x3;
      f1 = f2; f2 = iifo_loo_nll(x, n, 0, x2 * h0);
    } else {
      x3 = x2; x2 = x1; x1 = R * x2 + C * x0;
      f2 = f1; f1 = iifo_loo_nll(x, n, 0, x1 * h0);
    }
  }
  return (f1 < f2) ? x1 : x2;
}

/* Optim.jl GoldenSection on [lo, hi] as used by AMP kde!_CircularNaiveCV */
static double golden_optim(const double* x, int n, double lo, double hi, double rel_tol) {
  const double gr = 0.5 * (3.0 - sqrt(5.0));
  const double abs_tol = 2.220446049250313e-16;
  double xm = lo + gr * (hi - lo);
  double fm = iifo_loo_nll(x, n, 1, xm);
  for (int it = 0; it < 200; ++it) {
    double tolx = rel_tol * fabs(xm) + abs_tol;
    double mid = 0.5 * (hi + lo);
    if (fabs(xm - mid) <= 2 * tolx - 0.5 * (hi - lo)) break;
    if (hi - xm > xm - lo) {
      double xn = xm + gr * (hi - xm);
      double fn = iifo_loo_nll(x, n, 1, xn);
      if (fn < fm) { lo = xm; xm = xn; fm = fn; } else hi = xn;
    } else {
      double xn = xm - gr * (xm - lo);
      double fn = iifo_loo_nll(x, n, 1, xn);
      if (fn < fm) { hi = xm; xm = xn; fm = fn; } else lo = xn;
    }
  }
  return xm;
}

int32_t iifo_kde_bandwidth(const double* pts, int32_t n, int32_t d, int32_t cm, double* bw_out) {
  if (n < 2 || d < 1 || d > IIF_MAX_DIM) return IIF_ERR_ARG;
  double* x = (double*)malloc(sizeof(double) * 2 * n);
  double* xs = x + n;
  for (int c = 0; c < d; ++c) {
    for (int i = 0; i < n; ++i) x[i] = pts[i * d + c];
    if (is_circ(cm, c)) {
      bw_out[c] = golden_optim(x, n, 1e-3, TWO_PI, 1e-3);
    } else {
      memcpy(xs, x, sizeof(double) * n);
      qsort(xs, n, sizeof(double), cmp_double);
      double maxm = xs[n - 1] - xs[0], minm = maxm;
      node_min_diam(xs, 0, n - 1, &minm);
      if (minm < 1e-6) minm = 1e-6;
      double h0 = 0.5 * (minm + maxm);
      double a = golden_nr(xs, n, h0, 2.0 * minm / (minm + maxm), 1.0, 2.0 * maxm / (minm + maxm), 1e-2);
      bw_out[c] = a * h0;
    }
  }
  free(x);
  return IIF_OK;
}

/* ------------------------------------------------------------------------------------ */
/* approxDeconv (src/services/DeconvUtils.jl:32-162) — the inverse of the convolution:  */
/* for every sample n solve the factor residual for the MEASUREMENT given particle n of every        */
/* variable (first hypothesis only, :76,:139), starting from a sampled measurement.  Returns          */
/* (predicted, sampled) measurements.  The built-in residuals are linear in z, so the root is closed  */
/* form: prior z = x1, LinearRelative z = x2 - x1, CircularCircular z = wrap(x2 - x1), EuclidDistance */
/* z = |x2 - x1|.  Multihypo factors are not supported by the reference either (:22, #1096).          */
/* ------------------------------------------------------------------------------------ */
int32_t iifo_deconv(const iifo_graph* g, int32_t factor, int32_t N, int32_t call_id,
                    double* out_pred, double* out_meas) {
  if (factor < 0 || factor >= g->nfactors || N < 1) return IIF_ERR_ARG;
  const iif_factor_desc* f = &g->factors[factor];
  if (f->nmh != 0) return IIF_ERR_UNSUPPORTED;
  if (!(f->arity == 1 || f->arity == 2)) return IIF_ERR_UNSUPPORTED;
  const int zd = f->zdim;
  const iif_slot_desc* S1 = &g->slots[f->slot[0]];
  const iif_slot_desc* S2 = f->arity == 2 ? &g->slots[f->slot[1]] : NULL;
  for (int n = 0; n < N; ++n) {
    double z[IIF_MAX_DIM] = {0, 0, 0, 0};
    int32_t st = sample_measurement(g, f, (uint32_t)call_id, n, z);
    if (st != IIF_OK) return st;
    for (int c = 0; c < zd; ++c) out_meas[n * zd + c] = z[c];
    const double* x[2] = {NULL, NULL};
    for (int v = 0; v < f->arity; ++v) {       /* _getindex_anyn, NumericalCalculations.jl:377-381 */
      const iif_slot_desc* S = v == 0 ? S1 : S2;
      int len = g->npts[f->slot[v]];
      if (len <= 0) return IIF_ERR_STATE;
      int m = n;
      if (n >= len) {
        double u = iifo_uniform(g->sp.seed, (uint32_t)call_id, IIF_RS_ANYN, (uint32_t)(v * N + n));
        m = (int)(u * len);
        if (m >= len) m = len - 1;
      }
      x[v] = g->pts + S->pts_off + (size_t)m * S->dim;
    }
    double* p = out_pred + (size_t)n * zd;
    switch (f->kind) {
      case IIF_F_PRIOR: case IIF_F_MSG_PRIOR:
        for (int c = 0; c < zd; ++c) p[c] = x[0][c];
        break;
      case IIF_F_PRIOR_CIRCULAR:
        p[0] = wrap_pi(x[0][0]);
        break;
      case IIF_F_PARTIAL_PRIOR: {
        int k = 0;
        for (int c = 0; c < S1->dim; ++c) if ((f->partial_mask >> c) & 1) p[k++] = x[0][c];
        break;
      }
      case IIF_F_MANIFOLD_PRIOR: {
        int k = 0;
        for (int c = 0; c < S1->dim; ++c)
          if (!f->partial_mask || ((f->partial_mask >> c) & 1))
            p[k++] = is_circ(S1->circ_mask, c) ? wrap_pi(x[0][c]) : x[0][c];
        break;
      }
      case IIF_F_SO3_PRIOR: so3_between(f->aux, x[0], p); break;
      case IIF_F_SO3_RELATIVE: so3_between(x[0], x[1], p); break;
      case IIF_F_SE2_RELATIVE: { /* X = vee(log(eps, p^-1 o q)) */
        double sn = sin(x[0][2]), cs = cos(x[0][2]);
        double dx = x[1][0] - x[0][0], dy = x[1][1] - x[0][1];
        p[0] = cs * dx + sn * dy;
        p[1] = -sn * dx + cs * dy;
        p[2] = wrap_pi(x[1][2] - x[0][2]);
        break;
      }
      case IIF_F_LINEAR_RELATIVE: case IIF_F_CIRCULAR_CIRCULAR:
        for (int c = 0; c < zd; ++c) p[c] = mdiff(x[1][c], x[0][c], is_circ(S1->circ_mask, c));
        break;
      case IIF_F_EUCLID_DISTANCE: {
        double s = 0;
        for (int c = 0; c < S1->dim; ++c) s += (x[1][c] - x[0][c]) * (x[1][c] - x[0][c]);
        p[0] = sqrt(s);
        break;
      }
      default: return IIF_ERR_UNSUPPORTED;
    }
  }
  return IIF_OK;
}

/* mmd (src/services/SolverUtilities.jl:25-47 -> AMP.mmd!, un-vendored, PARITY UNPINNED): kernel-embedding */
/* distance with ker(p,q) = exp(-bw * dist(p,q)^2):  sum k(a,a)/Na^2 - 2 sum k(a,b)/(Na Nb) + sum k(b,b)/Nb^2 */
double iifo_mmd(const double* a, int32_t na, const double* b, int32_t nb, int32_t d, int32_t cm, double bw) {
  double saa = 0, sab = 0, sbb = 0;
  for (int pass = 0; pass < 3; ++pass) {
    const double* P = pass == 2 ? b : a;
    const double* Q = pass == 0 ? a : b;
    const int np = pass == 2 ? nb : na, nq = pass == 0 ? na : nb;
    double s = 0;
    for (int i = 0; i < np; ++i)
      for (int j = 0; j < nq; ++j) {
        double d2 = 0;
        for (int c = 0; c < d; ++c) {
          double dl = mdiff(P[i * d + c], Q[j * d + c], is_circ(cm, c));
          d2 += dl * dl;
        }
        s += exp(-bw * d2);
      }
    if (pass == 0) saa = s; else if (pass == 1) sab = s; else sbb = s;
  }
  return saa / ((double)na * na) - 2.0 * sab / ((double)na * nb) + sbb / ((double)nb * nb);
}

/* ------------------------------------------------------------------------------------ */
/* a3/a4/a5: approxConvBelief -> evalFactor -> evalPotentialSpecific                        */
/* ------------------------------------------------------------------------------------ */
static int is_prior_kind(int k) {
  return k == IIF_F_PRIOR || k == IIF_F_PRIOR_CIRCULAR || k == IIF_F_MSG_PRIOR || k == IIF_F_PARTIAL_PRIOR ||
         k == IIF_F_MANIFOLD_PRIOR || k == IIF_F_SO3_PRIOR;
}

static double inflate_u(const iifo_graph* g, const iif_conv_op* op, const double* uinf, int cyc,
                        int n, int c, int N, int d) {
  uint32_t idx = (uint32_t)((cyc * N + n) * d + c);
  if (uinf && op->uinf_off >= 0) return uinf[op->uinf_off + idx];
  return iifo_uniform(g->sp.seed, (uint32_t)op->call_id, IIF_RS_INFLATE, idx);
}

/* addEntropyOnManifold! — src/services/EvalFactor.jl:95-132 on the elements with label `hyp` */
static void add_entropy(const iifo_graph* g, const iif_conv_op* op, const double* uinf, double* dest,
                        const int32_t* mhidx, int hyp, int N, int d, int32_t cm, int32_t dimmask,
                        double spread, int cyc) {
  for (int n = 0; n < N; ++n) {
    if (mhidx[n] != hyp) continue;
    if (is_so3(cm)) { /* retract(M, p, get_vector(M, p, Xc)) = p Exp(Xc) */
      double dl[3], nx[3];
      for (int c = 0; c < 3; ++c) dl[c] = ((dimmask >> c) & 1) ? spread * (inflate_u(g, op, uinf, cyc, n, c, N, d) - 0.5) : 0.0;
      so3_compose(dest + n * 3, dl, nx);
      for (int c = 0; c < 3; ++c) dest[n * 3 + c] = nx[c];
      continue;
    }
    for (int c = 0; c < d; ++c) {
      if (!((dimmask >> c) & 1)) continue;
      double u = inflate_u(g, op, uinf, cyc, n, c, N, d);
      dest[n * d + c] = madd(dest[n * d + c], spread * (u - 0.5), is_circ(cm, c));
    }
  }
}

/* calcVariableDistanceExpectedFractional — src/services/EvalFactor.jl:40-92 */
static double spread_distance(const iifo_graph* g, const iif_factor_desc* f, int sfidx,
                              const double* dest, int N, const int32_t* certain, int nc,
                              double kappa) {
  const iif_slot_desc* Ssf = &g->slots[f->slot[sfidx - 1]];
  int d = Ssf->dim;
  if (in_list(certain, nc, sfidx))                           /* :50-54 */
    return kappa * iifo_std_basic_spread(dest, N, d, Ssf->circ_mask);
  double ref[IIF_MAX_DIM], m[IIF_MAX_DIM];
  if (is_so3(Ssf->circ_mask)) geodesic_mean(dest, N, d, Ssf->circ_mask, ref);
  else default_mean(dest, N, d, Ssf->circ_mask, ref);        /* :71-74 */
  double best = 1e-2;                                        /* :90 */
  for (int v = 1; v <= f->arity; ++v) {
    const iif_slot_desc* S = &g->slots[f->slot[v - 1]];
    const double* p = (v == sfidx) ? dest : g->pts + S->pts_off;
    int np = (v == sfidx) ? N : g->npts[f->slot[v - 1]];
    if (in_list(certain, nc, v) || is_so3(S->circ_mask)) geodesic_mean(p, np, S->dim, S->circ_mask, m);  /* :84-88 */
    else default_mean(p, np, S->dim, S->circ_mask, m);       /* :64-67 */
    double s = 0;
    for (int c = 0; c < d && c < S->dim; ++c) s += (ref[c] - m[c]) * (ref[c] - m[c]);
    s = sqrt(s);
    if (s > best) best = s;
  }
  return kappa * best;
}

int32_t iifo_conv(const iifo_graph* g, const iif_conv_op* op, const double* meas,
                  const int32_t* mhidx_in, const double* uinf, double* out_pts, double* out_bw,
                  double* out_ipc, int32_t* out_mhidx, int32_t* nan_count) {
  if (op->factor < 0 || op->factor >= g->nfactors) return IIF_ERR_ARG;
  const iif_factor_desc* f = &g->factors[op->factor];
  int sfidx = op->sfidx, N = op->N;
  if (sfidx < 1 || sfidx > f->arity || N < 2 || N > IIF_MAX_POINTS) return IIF_ERR_ARG;
  int sslot = f->slot[sfidx - 1];
  const iif_slot_desc* S = &g->slots[sslot];
  int d = S->dim;
  int32_t cm = S->circ_mask;
  uint64_t seed = g->sp.seed;
  uint32_t call = (uint32_t)op->call_id;
  int nnan = 0;
#pragma omp atomic
  g_conv_count++;

  /* _beforeSolveCCW! (CalcFactor.jl:519-617): dest = deepcopy(X_sf) resized to N, new slots =
   * identity; maxlen = max(N, lengths...) — other variables longer than N are not supported. */
  int len_sf = g->npts[sslot] < N ? g->npts[sslot] : N;
  for (int v = 0; v < f->arity; ++v)
    if (v + 1 != sfidx && g->npts[f->slot[v]] > N) return IIF_ERR_ARG;
  double* dest = out_pts;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < d; ++c) dest[n * d + c] = n < len_sf ? g->pts[S->pts_off + n * d + c] : 0.0;

  /* fresh measurements (CalcFactor.jl:578) */
  double* z = (double*)malloc(sizeof(double) * (size_t)N * IIF_MAX_DIM);
  for (int n = 0; n < N; ++n) {
    if (meas && op->meas_off >= 0) {
      for (int c = 0; c < f->zdim; ++c) z[n * IIF_MAX_DIM + c] = meas[op->meas_off + n * f->zdim + c];
    } else {
      int32_t st = sample_measurement(g, f, call, n, z + n * IIF_MAX_DIM);
      if (st != IIF_OK) { free(z); return st; }
    }
  }

  /* hypothesis recipe (EvalFactor.jl:347-355 / :426-431) */
  double runnull = f->nullhypo > op->nullSurplus ? f->nullhypo : op->nullSurplus;
  int32_t isinit[IIF_MAX_ARITY];
  for (int v = 0; v < f->arity; ++v) isinit[v] = g->flags[f->slot[v]] & 1;
  double* ul = (double*)malloc(sizeof(double) * N);
  for (int n = 0; n < N; ++n) ul[n] = iifo_uniform(seed, call, IIF_RS_LABEL, (uint32_t)n);
  int32_t* mhidx = (int32_t*)malloc(sizeof(int32_t) * N);
  int32_t nb, bh[IIF_MAX_ARITY + 2], bnv[IIF_MAX_ARITY + 2], bv[(IIF_MAX_ARITY + 2) * IIF_MAX_ARITY];
  int32_t certain[IIF_MAX_ARITY], nc;
  const int32_t* lab_in = (mhidx_in && op->mhidx_off >= 0) ? mhidx_in + op->mhidx_off : NULL;
  int32_t st = iifo_hypo_recipe(f->nmh ? f->mh : NULL, f->arity, N, sfidx, isinit, runnull, ul,
                                lab_in, mhidx, &nb, bh, bnv, bv, certain, &nc);
  free(ul);
  if (st != IIF_OK) { free(z); free(mhidx); return st; }

  int32_t fullmask = (1 << d) - 1;
  int32_t pmask = f->partial_mask ? f->partial_mask : fullmask;
  int C = g->sp.inflateCycles;

  if (is_prior_kind(f->kind)) {
    /* evalPotentialSpecific, AbstractPrior — EvalFactor.jl:400-542 */
    double spreadDist = g->sp.spreadNH * iifo_std_basic_spread(dest, N, d, cm); /* :464 */
    int wrap = (f->kind == IIF_F_PRIOR_CIRCULAR || f->kind == IIF_F_MSG_PRIOR || f->kind == IIF_F_MANIFOLD_PRIOR);
    for (int n = 0; n < N; ++n) {
      if (mhidx[n] != 1) continue;                           /* ahmask :438 */
      if (f->kind == IIF_F_SO3_PRIOR) {                      /* retract(M, p, hat(Z)) = p Exp(z), GenericFunctions.jl:186-195 */
        so3_compose(f->aux, z + n * IIF_MAX_DIM, dest + n * 3);
      } else if (!f->partial_mask) {                         /* setPointsMani! :469-474 */
        for (int c = 0; c < d; ++c) {
          double v = z[n * IIF_MAX_DIM + c];
          dest[n * d + c] = (wrap && is_circ(cm, c)) ? wrap_pi(v) : v;
        }
      } else {                                               /* setPointPartial! :505-515 */
        int k = 0;
        for (int c = 0; c < d; ++c)
          if ((pmask >> c) & 1) {
            double v = z[n * IIF_MAX_DIM + (k++)];   /* ManifoldPriorPartial: point = exp(eps, hat(Xc)) => angles wrapped */
            dest[n * d + c] = (f->kind == IIF_F_MANIFOLD_PRIOR && is_circ(cm, c)) ? wrap_pi(v) : v;
          }
      }
    }
    /* null-hypothesis elements get entropy (:476 full, :531 partial dims only) */
    add_entropy(g, op, uinf, dest, mhidx, 0, N, d, cm, pmask, spreadDist, C);
  } else {
    /* evalPotentialSpecific, AbstractRelative — EvalFactor.jl:321-395
     * computeAcrossHypothesis! — EvalFactor.jl:145-237 */
    if (f->arity < 2) { free(z); free(mhidx); return IIF_ERR_ARG; }
    int sfincer = in_list(certain, nc, sfidx);
    for (int b = 0; b < nb; ++b) {
      int hyp = bh[b];
      const int32_t* vars = bv + b * IIF_MAX_ARITY;
      if ((sfincer && hyp != 0) || in_list(certain, nc, hyp) || hyp == sfidx) {   /* :171 */
        int nel = 0;
        for (int n = 0; n < N; ++n) nel += (mhidx[n] == hyp);
        /* active variables: need exactly (other, sf) for the binary residual library */
        int other = -1, sf_second = 0;
        if (nel > 0) {
          if (bnv[b] != 2 || !in_list(vars, 2, sfidx)) { free(z); free(mhidx); return IIF_ERR_UNSUPPORTED; }
          other = (vars[0] == sfidx) ? vars[1] : vars[0];
          sf_second = (vars[1] == sfidx);
        }
        for (int cyc = 0; cyc < C; ++cyc) {                  /* :184 inflateCycles */
          double sp = spread_distance(g, f, sfidx, dest, N, certain, nc, f->inflation); /* :186 */
          add_entropy(g, op, uinf, dest, mhidx, hyp, N, d, cm, pmask, sp, cyc);         /* :193 */
          if (nel == 0) continue;
          const iif_slot_desc* So = &g->slots[f->slot[other - 1]];
          int lo = g->npts[f->slot[other - 1]];
          for (int n = 0; n < N; ++n) {                      /* approxConvOnElements! :14-27 */
            if (mhidx[n] != hyp) continue;
            int m = n;                                       /* _getindex_anyn NumericalCalculations.jl:377-381 */
            if (n >= lo) {
              if (lo <= 0) { free(z); free(mhidx); return IIF_ERR_STATE; }
              double u = iifo_uniform(seed, call, IIF_RS_ANYN, (uint32_t)((other - 1) * N + n));
              m = (int)(u * lo);
              if (m >= lo) m = lo - 1;
            }
            double r[IIF_MAX_DIM];
            /* islen1 (1-D: BFGS in the reference) keeps the closed form; in several dimensions EuclidDistance (a ring
             * of roots) and factors flagged solver = 1 run the restated Nelder-Mead from the inflated start */
            if (d > 1 && (f->solver == 1 || f->kind == IIF_F_EUCLID_DISTANCE))
              nelder_mead_binary(f->kind, d, cm, f->zdim, z + n * IIF_MAX_DIM, g->pts + So->pts_off + m * So->dim,
                                 sf_second, dest + n * d, r);
            else
              solve_binary(f->kind, d, cm, z + n * IIF_MAX_DIM, g->pts + So->pts_off + m * So->dim,
                           sf_second, dest + n * d, r);
            int bad = 0;
            for (int c = 0; c < d; ++c) bad |= isnan(r[c]);
            if (bad) { nnan++; continue; }                   /* NumericalCalculations.jl:348-351 */
            for (int c = 0; c < d; ++c)
              if ((pmask >> c) & 1) dest[n * d + c] = r[c];
          }
        }
      } else {
        /* other-hypothesis (:208-220) and null-hypothesis (:222-231): entropy only, all dims */
        double sp = spread_distance(g, f, sfidx, dest, N, certain, nc, g->sp.spreadNH);
        add_entropy(g, op, uinf, dest, mhidx, hyp, N, d, cm, fullmask, sp, C);
      }
    }
  }

  /* ipc (:383-391 relative, :478,:535-536 prior) */
  for (int c = 0; c < IIF_MAX_DIM; ++c) out_ipc[c] = 0.0;
  for (int c = 0; c < d; ++c) out_ipc[c] = ((pmask >> c) & 1) ? 1.0 : 0.0;

  /* approxConvBelief: manikde!(M, pts; partial) — ApproxConv.jl:31-42; AMP sets the bandwidth of
   * non-partial coordinates to 1.0 when a partial is given and bw is auto-selected. */
  for (int c = 0; c < IIF_MAX_DIM; ++c) out_bw[c] = 0.0;
  double bw[IIF_MAX_DIM];
  st = iifo_kde_bandwidth(dest, N, d, cm, bw);
  for (int c = 0; c < d; ++c) out_bw[c] = ((pmask >> c) & 1) ? bw[c] : 1.0;
  if (out_mhidx) memcpy(out_mhidx, mhidx, sizeof(int32_t) * N);
  if (nan_count) *nan_count = nnan;
  free(z);
  free(mhidx);
  return st;
}

/* ------------------------------------------------------------------------------------ */
/* a15: AMP.manifoldProduct -> KDE.prodAppxMSGibbsS (multiscale Gibbs product)             */
/* PARITY UNPINNED (upstream packages not vendored).  Restated from Ihler et al. NIPS 2003. */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  int nlev;       /* number of levels below the root */
  int* lev_off;   /* nlev+2 offsets into node arrays; level 0 = root */
  int* lo;        /* node range [lo,hi] into perm */
  int* hi;
  int* child;     /* position (within next level) of the first child; leaf: its own copy */
  double* mean;   /* nnodes * d */
  double* var;    /* nnodes * d  (kernel variance + spread of the members) */
  double* wt;     /* nnodes */
  int* perm;      /* N point indices */
} ms_tree;

typedef struct { double v; int i; } keyed;
static int cmp_keyed(const void* a, const void* b) {
  const keyed* x = (const keyed*)a;
  const keyed* y = (const keyed*)b;
  if (x->v < y->v) return -1;
  if (x->v > y->v) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

/* KDE BallTree build: split each ball at the median of its most-spread coordinate
 * (left child gets ceil(n/2) points), level by level; node statistics are the moment-matched
 * Gaussian of the member kernels. */
static void tree_build(ms_tree* T, const double* pts, const double* bw, int N, int d, int32_t mask,
                       int nlev) {
  int maxnodes = (nlev + 1) * N + 1;
  T->nlev = nlev;
  T->lev_off = (int*)malloc(sizeof(int) * (nlev + 2));
  T->lo = (int*)malloc(sizeof(int) * maxnodes);
  T->hi = (int*)malloc(sizeof(int) * maxnodes);
  T->child = (int*)malloc(sizeof(int) * maxnodes);
  T->mean = (double*)malloc(sizeof(double) * maxnodes * d);
  T->var = (double*)malloc(sizeof(double) * maxnodes * d);
  T->wt = (double*)malloc(sizeof(double) * maxnodes);
  T->perm = (int*)malloc(sizeof(int) * N);
  keyed* tmp = (keyed*)malloc(sizeof(keyed) * N);
  for (int i = 0; i < N; ++i) T->perm[i] = i;
  T->lev_off[0] = 0;
  T->lo[0] = 0; T->hi[0] = N - 1;
  T->lev_off[1] = 1;
  for (int l = 0; l < nlev; ++l) {
    int w = T->lev_off[l + 1];
    for (int z = T->lev_off[l]; z < T->lev_off[l + 1]; ++z) {
      int lo = T->lo[z], hi = T->hi[z];
      T->child[z] = w - T->lev_off[l + 1];
      if (lo == hi) { T->lo[w] = lo; T->hi[w] = hi; w++; continue; }
      int best = -1; double bs = -1.0;
      for (int c = 0; c < d; ++c) {
        if (!((mask >> c) & 1)) continue;
        double mn = INFINITY, mx = -INFINITY;
        for (int i = lo; i <= hi; ++i) {
          double v = pts[T->perm[i] * d + c];
          if (v < mn) mn = v;
          if (v > mx) mx = v;
        }
        if (mx - mn > bs) { bs = mx - mn; best = c; }
      }
      for (int i = lo; i <= hi; ++i) { tmp[i - lo].v = pts[T->perm[i] * d + best]; tmp[i - lo].i = T->perm[i]; }
      qsort(tmp, hi - lo + 1, sizeof(keyed), cmp_keyed);
      for (int i = lo; i <= hi; ++i) T->perm[i] = tmp[i - lo].i;
      int mid = (lo + hi) / 2;
      T->lo[w] = lo; T->hi[w] = mid; w++;
      T->lo[w] = mid + 1; T->hi[w] = hi; w++;
    }
    T->lev_off[l + 2] = w;
  }
  for (int z = T->lev_off[nlev]; z < T->lev_off[nlev + 1]; ++z) T->child[z] = z - T->lev_off[nlev];
  int nn = T->lev_off[nlev + 1];
  for (int z = 0; z < nn; ++z) {
    int lo = T->lo[z], hi = T->hi[z], cnt = hi - lo + 1;
    T->wt[z] = (double)cnt / (double)N;
    for (int c = 0; c < d; ++c) {
      double s = 0;
      for (int i = lo; i <= hi; ++i) s += pts[T->perm[i] * d + c];
      double m = s / cnt, q = 0;
      for (int i = lo; i <= hi; ++i) { double e = pts[T->perm[i] * d + c] - m; q += e * e; }
      T->mean[z * d + c] = m;
      T->var[z * d + c] = bw[c] * bw[c] + q / cnt;
    }
  }
  free(tmp);
}

static void tree_free(ms_tree* T) {
  free(T->lev_off); free(T->lo); free(T->hi); free(T->child);
  free(T->mean); free(T->var); free(T->wt); free(T->perm);
}

/* product of the Gaussians currently selected in densities != skip, coordinate c.
 * AMP getManiMu/getManiLam: Euclid = precision-weighted mean; Circular = precision-weighted
 * extrinsic (atan2) mean.  Returns total precision (0 if no density informs c). */
static double cond_gauss(int F, const ms_tree* T, const int* node, const int32_t* masks, int skip,
                         int d, int c, int circ, double* mu) {
  double lam = 0, a = 0, sn = 0, cs = 0;
  for (int k = 0; k < F; ++k) {
    if (k == skip || !((masks[k] >> c) & 1)) continue;
    double l = 1.0 / T[k].var[node[k] * d + c];
    double m = T[k].mean[node[k] * d + c];
    lam += l;
    if (circ) { sn += l * sin(m); cs += l * cos(m); } else a += l * m;
  }
  if (lam > 0) *mu = circ ? atan2(sn, cs) : a / lam;
  return lam;
}

int32_t iifo_product(int32_t d, int32_t cm, int32_t F, int32_t N, const double* dens_pts,
                     const double* dens_bw, const int32_t* dens_mask, const double* old_pts,
                     uint64_t seed, uint32_t call_id, int32_t niter, const double* randU,
                     const double* randN, double* out_pts, double* out_bw, int32_t* out_labels) {
  if (F < 1 || F > IIF_MAX_FACTORS || N < 2 || d < 1 || d > IIF_MAX_DIM) return IIF_ERR_ARG;
  int32_t fullmask = (1 << d) - 1;
  int32_t masks[IIF_MAX_FACTORS];
  for (int j = 0; j < F; ++j) masks[j] = (dens_mask && dens_mask[j]) ? dens_mask[j] : fullmask;
  if (F == 1 && masks[0] == fullmask) {
    /* manifoldProduct with one density is a pass-through (no Gibbs, no re-bandwidth) */
    memcpy(out_pts, dens_pts, sizeof(double) * N * d);
    for (int c = 0; c < IIF_MAX_DIM; ++c) out_bw[c] = c < d ? dens_bw[c] : 0.0;
    if (out_labels) for (int s = 0; s < N; ++s) out_labels[s] = s;
    return IIF_OK;
  }
  int L = (int)floor(log((double)N) / log(2.0) + 1.0);       /* KDE: Nlevels */
  ms_tree* T = (ms_tree*)malloc(sizeof(ms_tree) * F);
  for (int j = 0; j < F; ++j)
    tree_build(&T[j], dens_pts + (size_t)j * N * d, dens_bw + j * IIF_MAX_DIM, N, d, masks[j], L);
  double* logw = (double*)malloc(sizeof(double) * N);
  /* Stream layout (AMP.manifoldProduct keyword vectors `_randU` / `_randN`; KDE.jl sizes them
   * Np*Ndens*(Niter+2)*Nlevels and Ndim*Np*(Nlevels+1)): per output sample the uniforms are consumed in the order
   * initIndices (F), then per level sampleIndices (F) and Niter sweeps of sampleIndex (F each); the normals are the
   * Nlevels+1 samplePoint draws (d each).  Here every sample owns a fixed-size block so that samples are
   * independent of each other: U block = F*(1 + L*(niter+1)), N block = d*(L+1). */
  const int ublk = F * (1 + L * (niter + 1)), nblk = d * (L + 1);
  for (int s = 0; s < N; ++s) {
    int node[IIF_MAX_FACTORS];
    uint32_t uc = (uint32_t)(s * ublk), nc = (uint32_t)(s * nblk);
    /* levelInit / initIndices: every level list holds the root only; the draw is trivial but consumes its uniform */
    for (int j = 0; j < F; ++j) { node[j] = 0; uc++; }
    double X[IIF_MAX_DIM];
    for (int l = 1; l <= L; ++l) {
      /* samplePoint(X): a point from the product of the currently selected (coarse) Gaussians */
      for (int c = 0; c < d; ++c) {
        double mu = 0;
        double lam = cond_gauss(F, T, node, masks, -1, d, c, is_circ(cm, c), &mu);
        uint32_t idx = nc++;
        double e = randN ? randN[idx] : iifo_normal(seed, call_id, IIF_RS_GIBBS_N, idx);
        X[c] = lam > 0 ? madd(mu, sqrt(1.0 / lam) * e, is_circ(cm, c)) : 0.0;
      }
      /* levelDown: every density's level list becomes the children of the previous list (a leaf stays).  The label
       * would follow the last child pushed, but sampleIndices below re-draws every label over the whole new list. */
      /* sampleIndices(X): label of every density given the point, p(z) ~ w_z N(X; mean_z, var_z) */
      for (int j = 0; j < F; ++j) {
        int z0 = T[j].lev_off[l], z1 = T[j].lev_off[l + 1], nz = z1 - z0;
        double best = INFINITY;
        for (int z = 0; z < nz; ++z) {
          double p = 0;
          for (int c = 0; c < d; ++c) {
            if (!((masks[j] >> c) & 1)) continue;
            double dl = mdiff(X[c], T[j].mean[(z0 + z) * d + c], is_circ(cm, c));
            double v = T[j].var[(z0 + z) * d + c];
            p += dl * dl / v + log(v);
          }
          logw[z] = p;
          if (p < best) best = p;
        }
        double tot = 0;
        for (int z = 0; z < nz; ++z) { logw[z] = exp(-0.5 * (logw[z] - best)) * T[j].wt[z0 + z]; tot += logw[z]; }
        uint32_t idx = uc++;
        double u = randU ? randU[idx] : iifo_uniform(seed, call_id, IIF_RS_GIBBS_U, idx);
        double thr = u * tot, cum = 0;
        int pick = nz - 1;
        for (int z = 0; z < nz; ++z) { cum += logw[z]; if (thr < cum) { pick = z; break; } }
        node[j] = z0 + pick;
      }
      for (int it = 0; it < niter; ++it) {
        for (int j = 0; j < F; ++j) {                        /* sampleIndex(j): label | the other densities' labels */
          double cmu[IIF_MAX_DIM] = {0}, clam[IIF_MAX_DIM];
          for (int c = 0; c < d; ++c)
            clam[c] = ((masks[j] >> c) & 1) ? cond_gauss(F, T, node, masks, j, d, c, is_circ(cm, c), &cmu[c]) : 0.0;
          int z0 = T[j].lev_off[l], z1 = T[j].lev_off[l + 1], nz = z1 - z0;
          double best = INFINITY;
          for (int z = 0; z < nz; ++z) {
            double p = 0;
            for (int c = 0; c < d; ++c) {
              if (!(clam[c] > 0)) continue;
              double dl = mdiff(T[j].mean[(z0 + z) * d + c], cmu[c], is_circ(cm, c));
              double v = T[j].var[(z0 + z) * d + c] + 1.0 / clam[c];
              p += dl * dl / v + log(v);
            }
            logw[z] = p;
            if (p < best) best = p;
          }
          double tot = 0;
          for (int z = 0; z < nz; ++z) { logw[z] = exp(-0.5 * (logw[z] - best)) * T[j].wt[z0 + z]; tot += logw[z]; }
          uint32_t idx = uc++;
          double u = randU ? randU[idx] : iifo_uniform(seed, call_id, IIF_RS_GIBBS_U, idx);
          double thr = u * tot, cum = 0;
          int pick = nz - 1;
          for (int z = 0; z < nz; ++z) { cum += logw[z]; if (thr < cum) { pick = z; break; } }
          node[j] = z0 + pick;
        }
      }
    }
    /* final samplePoint: draw from the product of the selected leaf kernels */
    for (int c = 0; c < d; ++c) {
      double mu = 0;
      double lam = cond_gauss(F, T, node, masks, -1, d, c, is_circ(cm, c), &mu);
      uint32_t idx = nc++;
      if (lam > 0) {
        double e = randN ? randN[idx] : iifo_normal(seed, call_id, IIF_RS_GIBBS_N, idx);
        out_pts[s * d + c] = madd(mu, sqrt(1.0 / lam) * e, is_circ(cm, c));
      } else {
        out_pts[s * d + c] = old_pts ? old_pts[s * d + c] : 0.0; /* no proposal informs c */
      }
    }
    if (out_labels)
      for (int j = 0; j < F; ++j) out_labels[s * F + j] = T[j].perm[T[j].lo[node[j]]];
  }
  free(logw);
  for (int j = 0; j < F; ++j) tree_free(&T[j]);
  free(T);
  for (int c = 0; c < IIF_MAX_DIM; ++c) out_bw[c] = 0.0;
  return iifo_kde_bandwidth(out_pts, N, d, cm, out_bw);      /* getKDEManifoldBandwidths on the result */
}

/* ------------------------------------------------------------------------------------ */
/* a1/a2: propagateBelief + proposalbeliefs! + setBelief!                                   */
/* GraphProductOperations.jl:16-64, ApproxConv.jl:238-304, SolveTree.jl:63-74               */
/* ------------------------------------------------------------------------------------ */
int32_t iifo_propagate(iifo_graph* g, const iif_prop_op* op) {
  int F = op->nfactors, N = op->N;
  if (F < 1 || F > IIF_MAX_FACTORS) return IIF_ERR_ARG;
  const iif_slot_desc* S = &g->slots[op->target_slot];
  const iif_slot_desc* O = &g->slots[op->out_slot];
  int d = S->dim;
  if (O->dim != d || O->cap < N) return IIF_ERR_ARG;
  double* prop = (double*)malloc(sizeof(double) * (size_t)F * N * d);
  double pbw[IIF_MAX_FACTORS * IIF_MAX_DIM], pipc[IIF_MAX_DIM];
  int32_t pmask[IIF_MAX_FACTORS];
  for (int fi = 0; fi < F; ++fi) {
    const iif_factor_desc* f = &g->factors[op->factor[fi]];
    iif_conv_op c;
    memset(&c, 0, sizeof(c));
    c.factor = op->factor[fi];
    c.sfidx = op->sfidx[fi];
    c.N = N;
    c.call_id = op->call_id + 1 + fi;
    /* ApproxConv.jl:256-265: relative, non-multihypo siblings of a multihypo factor */
    c.nullSurplus = (op->any_multihypo && !is_prior_kind(f->kind) && !f->nmh) ? g->sp.nullSurplusAdd : 0.0;
    c.meas_off = c.mhidx_off = c.uinf_off = -1;
    int32_t st = iifo_conv(g, &c, NULL, NULL, NULL, prop + (size_t)fi * N * d, pbw + fi * IIF_MAX_DIM,
                           pipc, NULL, NULL);
    if (st != IIF_OK) { free(prop); return st; }
    pmask[fi] = f->partial_mask;
  }
  /* oldPoints: current belief padded to N by sampling it (GraphProductOperations.jl:37-45) */
  double* oldp = (double*)malloc(sizeof(double) * (size_t)N * d);
  int len = g->npts[op->target_slot];
  for (int n = 0; n < N; ++n) {
    if (n < len) {
      for (int c = 0; c < d; ++c) oldp[n * d + c] = g->pts[S->pts_off + n * d + c];
    } else if (len > 0) {
      double u = iifo_uniform(g->sp.seed, (uint32_t)op->call_id, IIF_RS_OLDPAD, (uint32_t)(n * (d + 1)));
      int k = (int)(u * len);
      if (k >= len) k = len - 1;
      for (int c = 0; c < d; ++c) {
        double e = iifo_normal(g->sp.seed, (uint32_t)op->call_id, IIF_RS_OLDPAD, (uint32_t)(n * (d + 1) + 1 + c));
        oldp[n * d + c] = madd(g->pts[S->pts_off + k * d + c], g->bw[op->target_slot * IIF_MAX_DIM + c] * e,
                               is_circ(S->circ_mask, c));
      }
    } else {
      for (int c = 0; c < d; ++c) oldp[n * d + c] = 0.0;
    }
  }
  double* post = (double*)malloc(sizeof(double) * (size_t)N * d);
  double bw[IIF_MAX_DIM];
  int32_t st = iifo_product(d, S->circ_mask, F, N, prop, pbw, pmask, oldp, g->sp.seed,
                            (uint32_t)op->call_id, g->sp.gibbsNiter, NULL, NULL, post, bw, NULL);
  if (st == IIF_OK) {
    /* setBelief!/setValKDE! (FactorGraph.jl:237-286): val, bw, infoPerCoord, initialized */
    memcpy(g->pts + O->pts_off, post, sizeof(double) * (size_t)N * d);
    for (int c = 0; c < IIF_MAX_DIM; ++c) {
      g->bw[op->out_slot * IIF_MAX_DIM + c] = c < d ? bw[c] : 0.0;
      g->ipc[op->out_slot * IIF_MAX_DIM + c] = c < d ? (double)F : 0.0; /* ApproxConv.jl:296-300 */
    }
    g->npts[op->out_slot] = N;
    g->flags[op->out_slot] |= 1;
  }
  free(prop); free(oldp); free(post);
  return st;
}

/* IIF_S_DECONV — addLikelihoodsDifferentialCHILD! (src/services/TreeMessageUtils.jl:314-321):
 *   pred_X, = approxDeconv(tfg, afc.label); pts = exp.(M, e0, pred_X); newBel = manikde!(sft, pts)
 * the belief lands in `out_slot` (a measurement density the parent's relative factor samples from). */
int32_t iifo_deconv_to_slot(iifo_graph* g, const iif_deconv_op* op) {
  if (op->factor < 0 || op->factor >= g->nfactors || op->out_slot < 0 || op->out_slot >= g->nslots) return IIF_ERR_ARG;
  const iif_factor_desc* f = &g->factors[op->factor];
  const iif_slot_desc* O = &g->slots[op->out_slot];
  const int N = op->N, d = O->dim;
  if (f->nmh != 0 || f->arity != 2 || f->zdim != d || O->cap < N) return IIF_ERR_UNSUPPORTED;
  double* pred = (double*)malloc(sizeof(double) * 2 * (size_t)N * d);
  double* meas = pred + (size_t)N * d;
  int32_t st = iifo_deconv(g, op->factor, N, op->call_id, pred, meas);
  if (st != IIF_OK) { free(pred); return st; }
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < d; ++c) {
      double v = pred[n * d + c];
      g->pts[O->pts_off + n * d + c] = is_circ(O->circ_mask, c) ? wrap_pi(v) : v;   /* exp(M, eps, X) */
    }
  free(pred);
  double bw[IIF_MAX_DIM] = {0, 0, 0, 0};
  st = iifo_kde_bandwidth(g->pts + O->pts_off, N, d, O->circ_mask, bw);
  if (st != IIF_OK) return st;
  for (int c = 0; c < IIF_MAX_DIM; ++c) {
    g->bw[op->out_slot * IIF_MAX_DIM + c] = c < d ? bw[c] : 0.0;
    g->ipc[op->out_slot * IIF_MAX_DIM + c] = c < d ? 1.0 : 0.0;
  }
  g->npts[op->out_slot] = N;
  g->flags[op->out_slot] |= 1;
  return IIF_OK;
}

/* a16: clique Gibbs sweeps expressed as a wave schedule (fmcmc! SolveTree.jl:89-142 etc.) */
int32_t iifo_schedule_run(iifo_graph* g, int32_t nwaves, const int32_t* wave_off,
                          const iif_sched_op* ops, const iif_prop_op* props, int32_t first_wave,
                          int32_t last_wave) {
  return iifo_schedule_run_ex(g, nwaves, wave_off, ops, props, NULL, first_wave, last_wave);
}

int32_t iifo_schedule_run_ex(iifo_graph* g, int32_t nwaves, const int32_t* wave_off,
                             const iif_sched_op* ops, const iif_prop_op* props, const iif_deconv_op* deconvs,
                             int32_t first_wave, int32_t last_wave) {
  if (first_wave < 0) first_wave = 0;
  if (last_wave > nwaves) last_wave = nwaves;
  for (int w = first_wave; w < last_wave; ++w) {
    /* ops inside a wave are mutually independent (distinct destination slots, no op reads a slot
     * another op of the wave writes), so the multithreaded CPU baseline may run them concurrently —
     * the same clique/factor-level parallelism solveTree!(; multithread=true) exposes. */
    int32_t werr = IIF_OK;
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = wave_off[w]; k < wave_off[w + 1]; ++k) {
      const iif_sched_op* o = &ops[k];
      int32_t st = IIF_OK;
      if (o->kind == IIF_S_PROPAGATE) {
        st = iifo_propagate(g, &props[o->a]);
      } else if (o->kind == IIF_S_COPY) {
        const iif_slot_desc* A = &g->slots[o->a];
        const iif_slot_desc* B = &g->slots[o->b];
        if (A->dim != B->dim || B->cap < g->npts[o->a]) st = IIF_ERR_ARG;
        else {
          memcpy(g->pts + B->pts_off, g->pts + A->pts_off, sizeof(double) * (size_t)g->npts[o->a] * A->dim);
          memcpy(g->bw + o->b * IIF_MAX_DIM, g->bw + o->a * IIF_MAX_DIM, sizeof(double) * IIF_MAX_DIM);
          memcpy(g->ipc + o->b * IIF_MAX_DIM, g->ipc + o->a * IIF_MAX_DIM, sizeof(double) * IIF_MAX_DIM);
          g->npts[o->b] = g->npts[o->a];
          g->flags[o->b] = g->flags[o->a];
        }
      } else if (o->kind == IIF_S_DECONV && deconvs) {
        st = iifo_deconv_to_slot(g, &deconvs[o->a]);
      } else st = IIF_ERR_ARG;
      if (st != IIF_OK) {
#pragma omp critical
        werr = st;
      }
    }
    if (werr != IIF_OK) return werr;
  }
  return IIF_OK;
}
