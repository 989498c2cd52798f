// iif_conv.cuh — the Chapman-Kolmogorov convolution kernel (one CTA per approxConvBelief).
//
// Fuses, per convolution:
//   measurement sampling (sampleFactor!)            SolverUtilities.jl:50-76, Mixture.jl:114-155
//   hypothesis recipe + per-particle label draw      ExplicitDiscreteMarginalizations.jl:142-289
//   inflation / null-hypothesis entropy              EvalFactor.jl:40-132
//   per-particle residual solve                      NumericalCalculations.jl:413-452
//   prior proposal                                   EvalFactor.jl:400-542
//   KDE bandwidth (manikde!)                         ApproxConv.jl:36-42
// One thread owns one particle; block reductions provide the spread statistics.
#pragma once
#include "iif_device.cuh"

struct DeviceGraph {
  const iif_slot_desc* slots;
  const iif_factor_desc* factors;
  const iif_dist_desc* dists;
  const double* dparams;
  double* pts;
  double* bw;
  double* ipc;
  int32_t* npts;
  int32_t* flags;
  const iif_solver_params* sp;  // device memory: updated between graph replays without re-capture
  int32_t nslots, nfactors, ndists, _pad;
};

struct ConvTask {
  iif_conv_op op;
  double* out_pts;     // N*d
  double* out_bw;      // IIF_MAX_DIM
  double* out_ipc;     // IIF_MAX_DIM
  int32_t* out_mhidx;  // N or NULL
  int32_t* out_nan;    // 1 or NULL
  int32_t* out_status; // 1 or NULL (device-side error code for this task)
  int32_t out_slot;    // >= 0: single-factor propagateBelief — the proposal IS the posterior (manifoldProduct of
                       // one density is a pass-through), so it is written straight into this belief slot
  int32_t _pad;
};

__host__ __device__ inline size_t conv_smem_bytes(int N) {
  return sizeof(double) * ((size_t)N * IIF_MAX_DIM + (size_t)loo_xa_doubles(N) + (size_t)loo_x2_doubles(N) + (size_t)loo_scratch_doubles(N));
}

struct HypoRecipe {  // HypoRecipe, src/entities/HypoRecipe.jl:4-9 (elements are implicit: mhidx == hyp)
  int32_t nb;
  int32_t hyp[IIF_MAX_ARITY + 2];
  int32_t nv[IIF_MAX_ARITY + 2];
  int32_t vars[IIF_MAX_ARITY + 2][IIF_MAX_ARITY];
  int32_t ncert;
  int32_t cert[IIF_MAX_ARITY];
  int32_t np;     // categorical support size (0: all labels are 1)
  int32_t shift;  // subtract from the 1-based categorical draw
  double p[IIF_MAX_ARITY + 1];
  int32_t status;
};

__device__ __forceinline__ bool in_list(const int32_t* l, int n, int v) {
  for (int i = 0; i < n; ++i)
    if (l[i] == v) return true;
  return false;
}

__device__ __forceinline__ int categorical(const double* p, int np, double u) {  // 1-based
  double c = 0;
  int last = 1;
  for (int k = 0; k < np; ++k) {
    if (p[k] > 0) last = k + 1;
    c += p[k];
    if (u < c) return k + 1;
  }
  return last;
}

// _prepareHypoRecipe! structure (scalar, thread 0) — ExplicitDiscreteMarginalizations.jl:142-289
__device__ __noinline__ void build_recipe(HypoRecipe& R, const iif_factor_desc& f, int sfidx, const int32_t* isinit,
                             double nullhypo) {
  const int lenXi = f.arity;
  R.status = IIF_OK;
  if (f.nmh == 0) {  // `Nothing` method :234-289
    R.np = (nullhypo == 0) ? 0 : 2;
    R.p[0] = nullhypo;
    R.p[1] = 1.0 - nullhypo;
    R.shift = 1;
    R.ncert = lenXi;
    for (int i = 0; i < lenXi; ++i) R.cert[i] = i + 1;
    R.nb = lenXi + 1;
    for (int b = 0; b <= lenXi; ++b) {
      R.hyp[b] = b;
      if (b == 0) { R.nv[b] = 1; R.vars[b][0] = sfidx; }
      else if (b == 1) { R.nv[b] = lenXi; for (int i = 0; i < lenXi; ++i) R.vars[b][i] = i + 1; }
      else R.nv[b] = 0;
    }
    return;
  }
  int32_t uncertn[IIF_MAX_ARITY];
  int nc = 0, nu = 0;
  for (int i = 0; i < lenXi; ++i) {  // getHypothesesVectors :17-24
    if (f.mh[i] == 0.0) R.cert[nc++] = i + 1;
    else if (f.mh[i] > 0.0) uncertn[nu++] = i + 1;
  }
  R.ncert = nc;
  double p[IIF_MAX_ARITY + 1];
  int np = lenXi, ninit = 0;
  for (int i = 0; i < lenXi; ++i) { p[i] = f.mh[i]; ninit += (isinit[i] != 0); }
  if (ninit < lenXi - 1) {  // :161-172 suppress uninitialised hypotheses
    double s = 0;
    for (int i = 0; i < lenXi; ++i) {
      if (!isinit[i] && (i + 1 != sfidx)) p[i] = 0.0;
      s += p[i];
    }
    for (int i = 0; i < lenXi; ++i) p[i] /= s;
  }
  const bool sf_unc = in_list(uncertn, nu, sfidx);
  if (sf_unc) {  // :176-183 prepend the bad-init null class
    double nhw = (double)(nu + 1), q[IIF_MAX_ARITY + 1];
    q[0] = 1.0 / nhw;
    double s = q[0];
    for (int i = 0; i < lenXi; ++i) { q[i + 1] = (double)nu / nhw * p[i]; s += q[i + 1]; }
    for (int i = 0; i <= lenXi; ++i) p[i] = q[i] / s;
    np = lenXi + 1;
  }
  R.np = np;
  R.shift = sf_unc ? 1 : 0;
  for (int i = 0; i < np; ++i) R.p[i] = p[i];
  int pidx = sf_unc ? -1 : 0, nb = 0;
  const bool sfincer = in_list(R.cert, nc, sfidx);
  for (int k = 0; k < np; ++k) {  // :195-224
    pidx += 1;
    const bool pc = in_list(R.cert, nc, pidx);
    int nv = 0;
    int32_t* vars = R.vars[nb];
    if (!pc && sfincer && pidx != 0) {
      for (int v = 1; v <= lenXi; ++v) if (in_list(R.cert, nc, v) || v == pidx) vars[nv++] = v;
    } else if (((pc && !sfincer) || sfidx == pidx) && pidx != 0) {
      for (int v = 1; v <= lenXi; ++v) if (in_list(R.cert, nc, v) || v == sfidx) vars[nv++] = v;
    } else if (pc && sfincer && pidx != 0) {
      nv = 0;
    } else if (!pc && !sfincer && pidx != 0) {
      for (int i = 0; i < nu; ++i) vars[nv++] = uncertn[i];
    } else if (pidx == 0) {
      vars[nv++] = sfidx;
    } else R.status = IIF_ERR_ARG;
    R.hyp[nb] = pidx;
    R.nv[nb] = nv;
    nb++;
  }
  R.nb = nb;
}

// ---- measurement sampling ---------------------------------------------------------------
__device__ __forceinline__ void sample_simple(int kind, int dim, const double* prm, uint64_t seed,
                                              uint32_t call, int n, int zdim, double* z) {
  if (kind == IIF_D_NORMAL) {
    z[0] = prm[0] + prm[1] * rs_normal(seed, call, IIF_RS_MEAS, (uint32_t)(n * zdim));
  } else if (kind == IIF_D_UNIFORM) {
    z[0] = prm[0] + (prm[1] - prm[0]) * rs_uniform(seed, call, IIF_RS_MEAS, (uint32_t)(n * zdim));
  } else {  // MVNORMAL: mu + L*eps
    double e[IIF_MAX_DIM];
    for (int c = 0; c < dim; ++c) e[c] = rs_normal(seed, call, IIF_RS_MEAS, (uint32_t)(n * zdim + c));
    for (int r = 0; r < dim; ++r) {
      double acc = prm[r];
      for (int c = 0; c <= r; ++c) acc += prm[dim + r * dim + c] * e[c];
      z[r] = acc;
    }
  }
}

__device__ __noinline__ int sample_measurement(const DeviceGraph& g, const iif_factor_desc& f, uint32_t call, int n,
                                  double* z) {
  const iif_dist_desc D = g.dists[f.dist];
  const double* prm = g.dparams + D.poff;
  const uint64_t seed = g.sp->seed;
  switch (D.kind) {
    case IIF_D_NORMAL:
    case IIF_D_UNIFORM:
    case IIF_D_MVNORMAL: sample_simple(D.kind, D.dim, prm, seed, call, n, f.zdim, z); return IIF_OK;
    case IIF_D_MIXTURE: {  // Mixture.jl:137-151
      double u = rs_uniform(seed, call, IIF_RS_MIXLABEL, (uint32_t)n);
      int lbl = categorical(prm, D.ncomp, u) - 1;
      int blk = (D.comp_kind == IIF_D_MVNORMAL) ? D.dim + D.dim * D.dim : 2;
      sample_simple(D.comp_kind, D.dim, prm + D.ncomp + lbl * blk, seed, call, n, f.zdim, z);
      return IIF_OK;
    }
    case IIF_D_KDE: {  // MsgPrior.jl:27-30 -> samplePoint(mkd): kernel pick + bandwidth-scaled jitter
      const iif_slot_desc S = g.slots[D.slot];
      int np = g.npts[D.slot];
      if (np <= 0) return IIF_ERR_STATE;
      double u = rs_uniform(seed, call, IIF_RS_MIXLABEL, (uint32_t)n);
      int k = (int)(u * np);
      if (k >= np) k = np - 1;
      for (int c = 0; c < S.dim; ++c) {
        double e = rs_normal(seed, call, IIF_RS_MEAS, (uint32_t)(n * f.zdim + c));
        z[c] = madd(g.pts[S.pts_off + k * S.dim + c], g.bw[D.slot * IIF_MAX_DIM + c] * e,
                    is_circ(S.circ_mask, c));
      }
      return IIF_OK;
    }
    case IIF_D_SAMPLES: {  // host-drawn sample table of any distribution: uniform resampling with replacement
      const double u = rs_uniform(seed, call, IIF_RS_MIXLABEL, (uint32_t)n);
      const int k = min((int)(u * D.ncomp), D.ncomp - 1);
      for (int c = 0; c < D.dim; ++c) z[c] = prm[k * D.dim + c];
      return IIF_OK;
    }
    default: return IIF_ERR_UNSUPPORTED;
  }
}

// ---- spread statistics (block-wide) -----------------------------------------------------
// mean(M, pts, GeodesicInterpolation()): Euclid coordinates reduce in parallel (the sequential
// running mean equals the arithmetic mean up to rounding); circular coordinates follow the
// sequential geodesic recurrence on one thread (order dependent).  Result in mu_s (shared).
__device__ void block_geodesic_mean(const double* pts, int n, int d, int32_t cm, double* mu_s, double* red,
                                    int& parity) {
  if (is_so3(cm)) {  // mean(M, pts, GeodesicInterpolation()) on SO(3): mu <- mu Exp(Log(mu^T p_i) / (i + 1)), one thread
    if (threadIdx.x == 0) {
      double mu[3] = {0, 0, 0};
      if (n > 0) for (int c = 0; c < 3; ++c) mu[c] = pts[c];
      for (int i = 1; i < n; ++i) {
        double dl[3], nx[3];
        so3_between(mu, pts + i * 3, dl);
        const double t = 1.0 / (double)(i + 1);
        for (int c = 0; c < 3; ++c) dl[c] *= t;
        so3_compose(mu, dl, nx);
        for (int c = 0; c < 3; ++c) mu[c] = nx[c];
      }
      for (int c = 0; c < 3; ++c) mu_s[c] = mu[c];
    }
    __syncthreads();
    return;
  }
  double s[IIF_MAX_DIM] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += IIF_NT)
    for (int c = 0; c < d; ++c) s[c] += pts[i * d + c];
  const int nwa = (n + 31) >> 5;
  if (d == 1) s[0] = block_sum1(s[0], red, parity, nwa);
  else block_sum<IIF_MAX_DIM>(s, red, parity, nwa);
  if (threadIdx.x == 0) {
    for (int c = 0; c < d; ++c) {
      if (!is_circ(cm, c)) { mu_s[c] = n > 0 ? s[c] / n : 0.0; continue; }
      double mu = n > 0 ? pts[c] : 0.0;
      for (int i = 1; i < n; ++i) {
        double t = 1.0 / (double)(i + 1);
        mu = wrap_pi(mu + t * wrap_pi(pts[i * d + c] - mu));
      }
      mu_s[c] = mu;
    }
  }
  __syncthreads();
}

// mean(M, pts) default estimator: arithmetic (Euclid) / extrinsic atan2 (Circle)
__device__ void block_default_mean(const double* pts, int n, int d, int32_t cm, double* mu, double* red,
                                   int& parity) {
  double s[2 * IIF_MAX_DIM] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += IIF_NT)
    for (int c = 0; c < d; ++c) {
      double v = pts[i * d + c];
      if (is_circ(cm, c)) { s[c] += sin(v); s[IIF_MAX_DIM + c] += cos(v); } else s[c] += v;
    }
  block_sum<2 * IIF_MAX_DIM>(s, red, parity, (n + 31) >> 5);
  for (int c = 0; c < d; ++c)
    mu[c] = n > 0 ? (is_circ(cm, c) ? atan2(s[c], s[IIF_MAX_DIM + c]) : s[c] / n) : 0.0;
}

// calcStdBasicSpread — VariableStatistics.jl:22-36
__device__ __noinline__ double block_std_basic_spread(const double* pts, int n, int d, int32_t cm, double* mu_s, double* red,
                                         int& parity) {
  if (n < 2) return 1.0;
  block_geodesic_mean(pts, n, d, cm, mu_s, red, parity);
  double acc = 0;
  if (is_so3(cm)) {  // distance(M, mu, p)^2 = |log(mu, p)|_F^2 = 2 angle^2 (Frobenius metric of the skew matrices)
    for (int i = threadIdx.x; i < n; i += IIF_NT) {
      double dl[3];
      so3_between(mu_s, pts + i * 3, dl);
      acc += 2.0 * (dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += IIF_NT)
      for (int c = 0; c < d; ++c) {
        double v = mdiff(pts[i * d + c], mu_s[c], is_circ(cm, c));
        acc += v * v;
      }
  }
  acc = block_sum1(acc, red, parity, (n + 31) >> 5);
  double sigma = sqrt(acc / (double)(n - 1));
  return (1e-10 < sigma) ? sigma : 1.0;
}

// calcVariableDistanceExpectedFractional — EvalFactor.jl:40-92
__device__ __noinline__ double block_spread_distance(const DeviceGraph& g, const iif_factor_desc& f, int sfidx,
                                        const double* dest, int N, const HypoRecipe& R, double kappa,
                                        double* mu_s, double* red, int& parity) {
  const iif_slot_desc Ssf = g.slots[f.slot[sfidx - 1]];
  const int d = Ssf.dim;
  if (in_list(R.cert, R.ncert, sfidx)) return kappa * block_std_basic_spread(dest, N, d, Ssf.circ_mask, mu_s, red, parity);
  double ref[IIF_MAX_DIM], m[IIF_MAX_DIM];
  if (is_so3(Ssf.circ_mask)) {
    __syncthreads();
    block_geodesic_mean(dest, N, d, Ssf.circ_mask, mu_s, red, parity);
    for (int c = 0; c < d; ++c) ref[c] = mu_s[c];
    __syncthreads();
  } else {
    block_default_mean(dest, N, d, Ssf.circ_mask, ref, red, parity);
  }
  double best = 1e-2;
  for (int v = 1; v <= f.arity; ++v) {
    const iif_slot_desc S = g.slots[f.slot[v - 1]];
    const double* p = (v == sfidx) ? dest : g.pts + S.pts_off;
    const int np = (v == sfidx) ? N : g.npts[f.slot[v - 1]];
    if (in_list(R.cert, R.ncert, v) || is_so3(S.circ_mask)) {
      __syncthreads();
      block_geodesic_mean(p, np, S.dim, S.circ_mask, mu_s, red, parity);
      for (int c = 0; c < S.dim; ++c) m[c] = mu_s[c];
      __syncthreads();
    } else {
      block_default_mean(p, np, S.dim, S.circ_mask, m, red, parity);
    }
    double s = 0;
    for (int c = 0; c < d && c < S.dim; ++c) s += (ref[c] - m[c]) * (ref[c] - m[c]);
    s = sqrt(s);
    if (s > best) best = s;
  }
  return kappa * best;
}

// Per-sample solve of the binary residual library (see oracle solve_binary for the derivation):
// unique-root factors in closed form, EuclidDistance by radial projection of the start point.
__device__ __forceinline__ void solve_binary(int kind, int d, int32_t cm, const double* z, const double* other,
                                             bool sf_second, const double* u0, double* out) {
  if (kind == IIF_F_LINEAR_RELATIVE || kind == IIF_F_CIRCULAR_CIRCULAR) {
    for (int c = 0; c < d; ++c) out[c] = madd(other[c], sf_second ? z[c] : -z[c], is_circ(cm, c));
  } else if (kind == IIF_F_SE2_RELATIVE) {
    // ManifoldFactor{SpecialEuclidean(2)}: q = p o exp(eps, X)  (GenericFunctions.jl:39-44, hybrid tangent
    // representation: exp(eps, X) = (X_t, R(X_theta))).  Solving q: theta_q = theta_p + X_theta,
    // t_q = t_p + R(theta_p) X_t; solving p: theta_p = theta_q - X_theta, t_p = t_q - R(theta_p) X_t.
    const double th = sf_second ? other[2] : wrap_pi(other[2] - z[2]);
    double sn, cs;
    sincos(th, &sn, &cs);
    const double rx = cs * z[0] - sn * z[1], ry = sn * z[0] + cs * z[1];
    if (sf_second) { out[0] = other[0] + rx; out[1] = other[1] + ry; out[2] = wrap_pi(other[2] + z[2]); }
    else { out[0] = other[0] - rx; out[1] = other[1] - ry; out[2] = th; }
  } else {
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

// sum(res.^2) of a binary relative factor with the solve-for point at x (CalcFactorNormSq, NumericalCalculations.jl:68-72);
// residuals as in the oracle's iifo_residual (src/Factors/*.jl)
__device__ __noinline__ double cost_binary(int kind, int d, int32_t cm, const double* z, const double* x, const double* other,
                                           bool sf_second) {
  const double* p = sf_second ? other : x;
  const double* q = sf_second ? x : other;
  double s = 0.0;
  if (kind == IIF_F_LINEAR_RELATIVE) {
    for (int c = 0; c < d; ++c) { const double r = __dsub_rn(z[c], __dsub_rn(q[c], p[c])); s = __dadd_rn(s, __dmul_rn(r, r)); }
  } else if (kind == IIF_F_CIRCULAR_CIRCULAR) {
    for (int c = 0; c < d; ++c) {
      const double r = mdiff(madd(p[c], z[c], is_circ(cm, c)), q[c], is_circ(cm, c));
      s = __dadd_rn(s, __dmul_rn(r, r));
    }
  } else if (kind == IIF_F_EUCLID_DISTANCE) {
    double n2 = 0.0;
    for (int c = 0; c < d; ++c) { const double e = __dsub_rn(q[c], p[c]); n2 = __dadd_rn(n2, __dmul_rn(e, e)); }
    const double r = __dsub_rn(z[0], sqrt(n2));
    s = __dmul_rn(r, r);
  } else if (kind == IIF_F_SO3_RELATIVE) {  // |Log(q^T p Exp(z))|^2
    double qh[3], r[3];
    so3_compose(p, z, qh);
    so3_between(q, qh, r);
    s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  } else {  // IIF_F_SE2_RELATIVE
    double sn, cs;
    sincos(p[2], &sn, &cs);
    const double r0 = __dsub_rn(__dsub_rn(__dadd_rn(p[0], __dmul_rn(cs, z[0])), __dmul_rn(sn, z[1])), q[0]);
    const double r1 = __dsub_rn(__dadd_rn(__dadd_rn(p[1], __dmul_rn(sn, z[0])), __dmul_rn(cs, z[1])), q[1]);
    const double r2 = wrap_pi(wrap_pi(p[2] + z[2]) - q[2]);
    s = __dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2));
  }
  return s;
}

// (explicit _rn arithmetic: no FMA contraction, so the simplex follows the oracle's path bit for bit)
// Optim.NelderMead as _solveLambdaNumeric configures it (NumericalCalculations.jl:49-72, :90-133): AdaptiveParameters,
// AffineSimplexer, sqrt(var(f_simplex) n/(n+1)) < 1e-8, <= 1000 iterations, result = better of best vertex and centroid
// of the n best.  Same steps as the oracle's nelder_mead_binary.  One thread per particle; a rare path (EuclidDistance in
// more than one dimension, or factors that ask for the numeric solve), so the simplex lives in local memory.
__device__ __noinline__ void nelder_mead_binary(int kind, int d, int32_t cm, const double* z, const double* other,
                                                bool sf_second, const double* x0, double* out) {
  const int n = d, m = d + 1;
  const double alpha = 1.0, beta = 1.0 + 2.0 / n, gamma = 0.75 - 1.0 / (2.0 * n), delta = 1.0 - 1.0 / n;
  double S[IIF_MAX_DIM + 1][IIF_MAX_DIM], f[IIF_MAX_DIM + 1];
  int ord[IIF_MAX_DIM + 1];
  for (int i = 0; i < m; ++i) {
    for (int c = 0; c < n; ++c) S[i][c] = x0[c];
    if (i > 0) S[i][i - 1] = __dadd_rn(__dmul_rn(1.5, x0[i - 1]), 0.025);
    f[i] = cost_binary(kind, d, cm, z, S[i], other, sf_second);
  }
  double cen[IIF_MAX_DIM], xr[IIF_MAX_DIM], xt[IIF_MAX_DIM];
  for (int it = 0; it < 1000; ++it) {
    for (int i = 0; i < m; ++i) ord[i] = i;
    for (int i = 1; i < m; ++i) {
      const int k = ord[i];
      int j = i - 1;
      while (j >= 0 && f[ord[j]] > f[k]) { ord[j + 1] = ord[j]; --j; }
      ord[j + 1] = k;
    }
    const int lo = ord[0], hi = ord[m - 1], sh = ord[m - 2];
    for (int c = 0; c < n; ++c) {
      double s = 0;
      for (int i = 0; i < m - 1; ++i) s += S[ord[i]][c];
      cen[c] = s / n;
    }
    for (int c = 0; c < n; ++c) xr[c] = __dadd_rn(cen[c], __dmul_rn(alpha, __dsub_rn(cen[c], S[hi][c])));
    const double fr = cost_binary(kind, d, cm, z, xr, other, sf_second);
    bool shrink = false;
    if (fr < f[lo]) {
      for (int c = 0; c < n; ++c) xt[c] = __dadd_rn(cen[c], __dmul_rn(beta, __dsub_rn(xr[c], cen[c])));
      const double fe = cost_binary(kind, d, cm, z, xt, other, sf_second);
      if (fe < fr) { for (int c = 0; c < n; ++c) S[hi][c] = xt[c]; f[hi] = fe; }
      else { for (int c = 0; c < n; ++c) S[hi][c] = xr[c]; f[hi] = fr; }
    } else if (fr < f[sh]) {
      for (int c = 0; c < n; ++c) S[hi][c] = xr[c];
      f[hi] = fr;
    } else if (fr < f[hi]) {
      for (int c = 0; c < n; ++c) xt[c] = __dadd_rn(cen[c], __dmul_rn(gamma, __dsub_rn(xr[c], cen[c])));
      const double fc = cost_binary(kind, d, cm, z, xt, other, sf_second);
      if (fc <= fr) { for (int c = 0; c < n; ++c) S[hi][c] = xt[c]; f[hi] = fc; } else shrink = true;
    } else {
      for (int c = 0; c < n; ++c) xt[c] = __dsub_rn(cen[c], __dmul_rn(gamma, __dsub_rn(xr[c], cen[c])));
      const double fc = cost_binary(kind, d, cm, z, xt, other, sf_second);
      if (fc < f[hi]) { for (int c = 0; c < n; ++c) S[hi][c] = xt[c]; f[hi] = fc; } else shrink = true;
    }
    if (shrink)
      for (int i = 1; i < m; ++i) {
        const int k = ord[i];
        for (int c = 0; c < n; ++c) S[k][c] = __dadd_rn(S[lo][c], __dmul_rn(delta, __dsub_rn(S[k][c], S[lo][c])));
        f[k] = cost_binary(kind, d, cm, z, S[k], other, sf_second);
      }
    double mean = 0, var = 0;
    for (int i = 0; i < m; ++i) mean += f[i];
    mean /= m;
    for (int i = 0; i < m; ++i) { const double e = __dsub_rn(f[i], mean); var = __dadd_rn(var, __dmul_rn(e, e)); }
    if (sqrt(var / m) < 1e-8) break;
  }
  int best = 0, worst = 0;
  for (int i = 1; i < m; ++i) { if (f[i] < f[best]) best = i; if (f[i] > f[worst]) worst = i; }
  for (int c = 0; c < n; ++c) {
    double s = 0;
    for (int i = 0; i < m; ++i) if (i != worst) s += S[i][c];
    cen[c] = s / n;
  }
  const double fcen = cost_binary(kind, d, cm, z, cen, other, sf_second);
  const double* r = fcen < f[best] ? cen : S[best];
  for (int c = 0; c < n; ++c) out[c] = is_circ(cm, c) ? wrap_pi(r[c]) : r[c];
  if (is_so3(cm)) {  // exp(M, eps, hat(minimizer)): the principal rotation vector of the minimiser
    const double zero[3] = {0, 0, 0};
    double pr[3];
    so3_compose(out, zero, pr);
    for (int c = 0; c < 3; ++c) out[c] = pr[c];
  }
}

__device__ __forceinline__ bool is_prior_kind(int k) {
  return k == IIF_F_PRIOR || k == IIF_F_PRIOR_CIRCULAR || k == IIF_F_MSG_PRIOR || k == IIF_F_PARTIAL_PRIOR ||
         k == IIF_F_MANIFOLD_PRIOR || k == IIF_F_SO3_PRIOR;
}

// Rare per-particle paths (SpecialOrthogonal(3) composition, the numeric Nelder-Mead solve), out of line and with BY-VALUE
// arguments so that the per-particle state of the caller stays in registers.
struct Vec4 { double v[IIF_MAX_DIM]; };
__device__ __forceinline__ Vec4 pack4(const double* a) { Vec4 r = {{a[0], a[1], a[2], a[3]}}; return r; }
__device__ __noinline__ Vec4 so3_compose_v(Vec4 a, Vec4 b) {
  Vec4 r = {{0, 0, 0, 0}};
  so3_compose(a.v, b.v, r.v);
  return r;
}
__device__ __noinline__ Vec4 solve_rare(int kind, int d, int32_t cm, Vec4 z, Vec4 other, bool sf_second, Vec4 u0, bool numeric) {
  Vec4 r = {{0, 0, 0, 0}};
  if (numeric) nelder_mead_binary(kind, d, cm, z.v, other.v, sf_second, u0.v, r.v);
  else if (kind == IIF_F_SO3_RELATIVE) {  // ManifoldFactor{SpecialOrthogonal(3)}: q = p Exp(X); p = q Exp(X)^-1 = q Exp(-X)
    if (sf_second) so3_compose(other.v, z.v, r.v);
    else { const double nz_[3] = {-z.v[0], -z.v[1], -z.v[2]}; so3_compose(other.v, nz_, r.v); }
  } else solve_binary(kind, d, cm, z.v, other.v, sf_second, u0.v, r.v);
  return r;
}
// does this convolution need the RARE variant of the kernel?  (host and device agree through this one predicate)
__host__ __device__ inline bool conv_needs_rare(const iif_factor_desc& F, const iif_slot_desc* slots, int sfidx) {
  if (F.kind == IIF_F_SO3_PRIOR || F.kind == IIF_F_SO3_RELATIVE) return true;
  for (int v = 0; v < F.arity; ++v)
    if (slots[F.slot[v]].circ_mask & IIF_MANI_SO3) return true;
  const int d = slots[F.slot[sfidx - 1]].dim;
  const bool relative = !(F.kind == IIF_F_PRIOR || F.kind == IIF_F_PRIOR_CIRCULAR || F.kind == IIF_F_MSG_PRIOR ||
                          F.kind == IIF_F_PARTIAL_PRIOR || F.kind == IIF_F_MANIFOLD_PRIOR);
  return relative && d > 1 && (F.solver == 1 || F.kind == IIF_F_EUCLID_DISTANCE);
}

// The kernel exists in two variants.  RARE = false is the hot one: everything the common factors need and nothing
// else, so the per-particle phase keeps its state in registers (the SO(3) / Nelder-Mead paths cost the common kernel
// 8 % when they were compiled into it).  RARE = true adds those paths; the host picks the variant per launch
// (conv_needs_rare), and a task that needs them but lands in the lean kernel fails with IIF_ERR_STATE.
template <bool RARE>
__device__ __forceinline__ void conv_body(DeviceGraph g, const ConvTask* __restrict__ tasks, const double* __restrict__ meas,
                const int32_t* __restrict__ mhidx_in, const double* __restrict__ uinf,
                const TreeStruct* __restrict__ trees) {
  // dynamic shared memory, sized by the host for the launch's largest N (conv_smem_bytes)
  extern __shared__ __align__(16) double conv_smem[];
  __shared__ double red[IIF_RED_DOUBLES];
  __shared__ double mu_s[IIF_MAX_DIM];
  __shared__ HypoRecipe R;
  __shared__ iif_factor_desc f;
  __shared__ int s_status;

  IIF_PHASE_ZERO();
  IIF_PHASE_BEGIN();
  // narrow launches give every convolution a cluster of CTAs that run redundantly and share the bandwidth
  // search (iif_device.cuh, "cluster-speculative"); only rank 0 writes
  const int cC = (int)cooperative_groups::this_cluster().num_blocks();
  const bool wr = cooperative_groups::this_cluster().block_rank() == 0;
  if (cC > 1) cooperative_groups::this_cluster().sync();  // every rank is resident before any remote shared-memory access
  const ConvTask t = tasks[blockIdx.x / cC];
  const iif_conv_op op = t.op;
  const int n = threadIdx.x;
  int parity = 0;
  double* dest = conv_smem;                         // N * IIF_MAX_DIM
  double* xa = dest + (size_t)op.N * IIF_MAX_DIM;   // N
  double* xb = xa + loo_xa_doubles(op.N);           // loo_x2_doubles(N)
  double* scr = xb + loo_x2_doubles(op.N);          // loo_scratch_doubles(N)
  if (n == 0) {
    f = g.factors[op.factor];
    s_status = IIF_OK;
  }
  __syncthreads();
  const int sfidx = op.sfidx, N = op.N;
  const int sslot = f.slot[sfidx - 1];
  const iif_slot_desc S = g.slots[sslot];
  const int d = S.dim;
  const int32_t cm = S.circ_mask;
  const uint64_t seed = g.sp->seed;
  const uint32_t call = (uint32_t)op.call_id;
  const bool active = n < N;

  // _beforeSolveCCW!: dest = deepcopy(X_sf) resized to N, new slots = identity (CalcFactor.jl:543-565)
  double my[IIF_MAX_DIM] = {0, 0, 0, 0};
  {
    int len_sf = min(g.npts[sslot], N);
    if (active) {
      for (int c = 0; c < d; ++c) {
        my[c] = n < len_sf ? g.pts[S.pts_off + n * d + c] : 0.0;
        dest[n * d + c] = my[c];
      }
    }
  }
  IIF_PHASE(8);
  // fresh measurements (CalcFactor.jl:578)
  double z[IIF_MAX_DIM] = {0, 0, 0, 0};
  if (active) {
    if (meas != nullptr && op.meas_off >= 0) {
      for (int c = 0; c < f.zdim; ++c) z[c] = meas[op.meas_off + n * f.zdim + c];
    } else {
      int st = sample_measurement(g, f, call, n, z);
      if (st != IIF_OK) s_status = st;
    }
  }
  // hypothesis recipe structure + per-particle label
  if (n == 0) {
    int32_t isinit[IIF_MAX_ARITY];
    for (int v = 0; v < f.arity; ++v) isinit[v] = g.flags[f.slot[v]] & 1;
    double runnull = fmax(f.nullhypo, op.nullSurplus);  // EvalFactor.jl:352
    build_recipe(R, f, sfidx, isinit, runnull);
    if (R.status != IIF_OK) s_status = R.status;
    for (int v = 0; v < f.arity; ++v)
      if (v + 1 != sfidx && g.npts[f.slot[v]] > N) s_status = IIF_ERR_ARG;
    if (!RARE && conv_needs_rare(f, g.slots, sfidx)) s_status = IIF_ERR_STATE;   // host picked the wrong variant
  }
  __syncthreads();
  int label = 1;
  if (active) {
    if (mhidx_in != nullptr && op.mhidx_off >= 0) label = mhidx_in[op.mhidx_off + n];
    else if (R.np > 0) label = categorical(R.p, R.np, rs_uniform(seed, call, IIF_RS_LABEL, (uint32_t)n)) - R.shift;
  }
  const int32_t fullmask = (1 << d) - 1;
  const int32_t pmask = f.partial_mask ? f.partial_mask : fullmask;
  const int C = g.sp->inflateCycles;
  int nnan = 0;
  IIF_PHASE(9);

  auto inflate_u = [&](int cyc, int c) -> double {
    uint32_t idx = (uint32_t)((cyc * N + n) * d + c);
    if (uinf != nullptr && op.uinf_off >= 0) return uinf[op.uinf_off + idx];
    return rs_uniform(seed, call, IIF_RS_INFLATE, idx);
  };
  // addEntropyOnManifold! on this thread's particle (EvalFactor.jl:95-132)
  auto add_entropy = [&](int hyp, int32_t dimmask, double spread, int cyc) {
    if (active && label == hyp) {
      if (RARE && is_so3(cm)) {  // retract(M, p, get_vector(M, p, Xc)) = p Exp(Xc)
        Vec4 dl = {{0, 0, 0, 0}};
        for (int c = 0; c < 3; ++c) dl.v[c] = ((dimmask >> c) & 1) ? spread * (inflate_u(cyc, c) - 0.5) : 0.0;
        const Vec4 nx = so3_compose_v(pack4(my), dl);
        for (int c = 0; c < 3; ++c) { my[c] = nx.v[c]; dest[n * d + c] = nx.v[c]; }
      } else {
        for (int c = 0; c < d; ++c) {
          if (!((dimmask >> c) & 1)) continue;
          my[c] = madd(my[c], spread * (inflate_u(cyc, c) - 0.5), is_circ(cm, c));
          dest[n * d + c] = my[c];
        }
      }
    }
  };

  if (s_status == IIF_OK) {
    if (is_prior_kind(f.kind)) {
      // evalPotentialSpecific (AbstractPrior) — EvalFactor.jl:400-542
      double spreadDist = g.sp->spreadNH * block_std_basic_spread(dest, N, d, cm, mu_s, red, parity);  // :464
      const bool wrap = (f.kind == IIF_F_PRIOR_CIRCULAR || f.kind == IIF_F_MSG_PRIOR || f.kind == IIF_F_MANIFOLD_PRIOR);
      if (active && label == 1) {
        if (RARE && f.kind == IIF_F_SO3_PRIOR) {
          const Vec4 nx = so3_compose_v(pack4(f.aux), pack4(z));     // retract(M, p, hat(Z)) = p Exp(z)
          for (int c = 0; c < 3; ++c) my[c] = nx.v[c];
        } else if (!f.partial_mask) {
          for (int c = 0; c < d; ++c) my[c] = (wrap && is_circ(cm, c)) ? wrap_pi(z[c]) : z[c];
        } else {
          int k = 0;
          for (int c = 0; c < d; ++c)
            if ((pmask >> c) & 1) {
              const double v = z[k++];
              my[c] = (f.kind == IIF_F_MANIFOLD_PRIOR && is_circ(cm, c)) ? wrap_pi(v) : v;
            }
        }
        for (int c = 0; c < d; ++c) dest[n * d + c] = my[c];
      }
      add_entropy(0, pmask, spreadDist, C);
    } else if (f.arity < 2) {
      if (n == 0) s_status = IIF_ERR_ARG;
    } else {
      // evalPotentialSpecific (AbstractRelative) + computeAcrossHypothesis! — EvalFactor.jl:145-237,321-395
      const bool sfincer = in_list(R.cert, R.ncert, sfidx);
      for (int b = 0; b < R.nb; ++b) {
        const int hyp = R.hyp[b];
        // a bucket without elements changes nothing (its spread would only feed its own entropy): skip it
        const int nel = __syncthreads_count(active && label == hyp);
        if (nel == 0) continue;
        if ((sfincer && hyp != 0) || in_list(R.cert, R.ncert, hyp) || hyp == sfidx) {
          if (R.nv[b] != 2 || !in_list(R.vars[b], 2, sfidx)) {
            if (n == 0) s_status = IIF_ERR_UNSUPPORTED;
            break;
          }
          const int other = (R.vars[b][0] == sfidx) ? R.vars[b][1] : R.vars[b][0];
          const bool sf_second = (R.vars[b][1] == sfidx);
          // the partner particle of this sample does not change between the inflation cycles: fetch it once
          double o[IIF_MAX_DIM] = {0, 0, 0, 0};
          bool ok = false;
          if (active && label == hyp) {
            const int oslot = f.slot[other - 1];
            const iif_slot_desc So = g.slots[oslot];
            const int lo = g.npts[oslot];
            int m = n;  // _getindex_anyn, NumericalCalculations.jl:377-381
            ok = true;
            if (n >= lo) {
              if (lo <= 0) { s_status = IIF_ERR_STATE; ok = false; }
              else {
                double u = rs_uniform(seed, call, IIF_RS_ANYN, (uint32_t)((other - 1) * N + n));
                m = min((int)(u * lo), lo - 1);
              }
            }
            if (ok)
              for (int c = 0; c < So.dim; ++c) o[c] = g.pts[So.pts_off + m * So.dim + c];
          }
          for (int cyc = 0; cyc < C; ++cyc) {
            __syncthreads();
            double sp = block_spread_distance(g, f, sfidx, dest, N, R, f.inflation, mu_s, red, parity);
            add_entropy(hyp, pmask, sp, cyc);
            if (ok) {
              double r[IIF_MAX_DIM];
              // islen1 (1-D: BFGS in the reference) keeps the closed form; in several dimensions EuclidDistance (a ring of
              // roots) and factors that ask for it run the restated Nelder-Mead from the inflated start
              if (RARE) {
                const bool numeric = d > 1 && (f.solver == 1 || f.kind == IIF_F_EUCLID_DISTANCE);
                const Vec4 r4 = solve_rare(f.kind, d, cm, pack4(z), pack4(o), sf_second, pack4(my), numeric);
                for (int c = 0; c < IIF_MAX_DIM; ++c) r[c] = r4.v[c];
              } else solve_binary(f.kind, d, cm, z, o, sf_second, my, r);
              bool bad = false;
              for (int c = 0; c < d; ++c) bad |= isnan(r[c]);
              if (bad) nnan++;  // NumericalCalculations.jl:348-351: particle left unchanged
              else
                for (int c = 0; c < d; ++c)
                  if ((pmask >> c) & 1) { my[c] = r[c]; dest[n * d + c] = r[c]; }
            }
          }
        } else {
          // other-hypothesis (:208-220) and null-hypothesis (:222-231): entropy only, all dims
          __syncthreads();
          double sp = block_spread_distance(g, f, sfidx, dest, N, R, g.sp->spreadNH, mu_s, red, parity);
          add_entropy(hyp, fullmask, sp, C);
        }
      }
    }
  }
  __syncthreads();
  IIF_PHASE(10);
  const int status = s_status;
  if (t.out_status != nullptr && n == 0 && wr) *t.out_status = status;
  if (status != IIF_OK) return;

  // proposal points out
  if (active && wr)
    for (int c = 0; c < d; ++c) t.out_pts[n * d + c] = dest[n * d + c];
  if (t.out_mhidx != nullptr && active && wr) t.out_mhidx[n] = label;
  if (t.out_nan != nullptr) {
    double tot = block_sum1((double)nnan, red, parity);
    if (n == 0 && wr) *t.out_nan = (int32_t)tot;
  }
  // approxConvBelief: manikde!(M, pts; partial) — ApproxConv.jl:31-42
  double bw[IIF_MAX_DIM];
  IIF_PHASE(11);
  block_kde_bandwidth<0>(dest, N, d, cm, &trees[N], xa, xb, scr, red, &parity, bw);
  IIF_PHASE(12);
  if (!wr) return;
  if (n == 0) {
    for (int c = 0; c < IIF_MAX_DIM; ++c) {
      t.out_bw[c] = c < d ? (((pmask >> c) & 1) ? bw[c] : 1.0) : 0.0;
      t.out_ipc[c] = c < d ? (((pmask >> c) & 1) ? 1.0 : 0.0) : 0.0;
    }
  }
  if (t.out_slot >= 0) {  // setBelief! of a one-factor propagateBelief (SolveTree.jl:74)
    const iif_slot_desc O = g.slots[t.out_slot];
    if (active)
      for (int c = 0; c < d; ++c) g.pts[O.pts_off + n * d + c] = dest[n * d + c];
    if (n < IIF_MAX_DIM) {
      g.bw[t.out_slot * IIF_MAX_DIM + n] = n < d ? bw[n] : 0.0;
      g.ipc[t.out_slot * IIF_MAX_DIM + n] = n < d ? 1.0 : 0.0;
    }
    if (n == 0) {
      g.npts[t.out_slot] = N;
      g.flags[t.out_slot] |= 1;
    }
  }
  IIF_PHASE(13);
  IIF_PHASE_FLUSH();
}

__global__ void __launch_bounds__(IIF_MAX_THREADS, 1)
iif_conv_kernel(DeviceGraph g, const ConvTask* __restrict__ tasks, const double* __restrict__ meas,
                const int32_t* __restrict__ mhidx_in, const double* __restrict__ uinf, const TreeStruct* __restrict__ trees) {
  conv_body<false>(g, tasks, meas, mhidx_in, uinf, trees);
}
__global__ void __launch_bounds__(IIF_MAX_THREADS, 1)
iif_conv_kernel_rare(DeviceGraph g, const ConvTask* __restrict__ tasks, const double* __restrict__ meas,
                     const int32_t* __restrict__ mhidx_in, const double* __restrict__ uinf, const TreeStruct* __restrict__ trees) {
  conv_body<true>(g, tasks, meas, mhidx_in, uinf, trees);
}
