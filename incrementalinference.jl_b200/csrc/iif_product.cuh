// iif_product.cuh — KDE product kernel (one CTA per AMP.manifoldProduct, call site
// GraphProductOperations.jl:53-60) fused with setBelief! (SolveTree.jl:74) and the re-bandwidth.
//
// Multiscale Gibbs sampling from a product of F kernel density estimates (Ihler, Sudderth,
// Freeman & Willsky, NIPS 2003; KDE.jl prodAppxMSGibbsS):
//   1. every density gets a median-split ball tree (rank sort inside each node along its
//      most-spread coordinate; node statistics = moment-matched Gaussian of the member kernels);
//   2. one warp per output sample walks the levels coarse->fine; per level and density the
//      lanes evaluate the candidate nodes' weights against the product of the other densities'
//      selected nodes, a warp scan + ballot picks the label by inverse CDF;
//   3. the sample is drawn from the product of the selected leaf kernels;
//   4. the posterior's bandwidth comes from block_kde_bandwidth (same LOO-CV as the proposals).
#pragma once
#include "iif_conv.cuh"

struct ProdTask {
  int32_t dim, circ_mask, F, N;
  int32_t call_id, randu_off, randn_off;
  int32_t target_slot;          // >=0: oldPoints come from this slot (padded on device)
  int32_t out_slot;             // >=0: posterior written into this slot (setBelief!)
  int32_t mask[IIF_MAX_FACTORS];  // partial masks (0 = full)
  const double* dens_pts;       // F * N * d
  const double* dens_bw;        // F * IIF_MAX_DIM
  const double* old_pts;        // explicit oldPoints N*d (target_slot < 0) or NULL
  double* out_pts;              // explicit outputs (may be NULL when out_slot >= 0)
  double* out_bw;
  int32_t* out_labels;          // N * F or NULL
  const int32_t* conv_status;   // F status words of the producing convolutions or NULL
  int32_t* out_status;
};

// dynamic shared memory layout (doubles first, then int16)
struct ProdSmem {
  double* P;      // F*N*d proposals
  double* mean;   // F*nn*d
  double* var;    // F*nn*d
  double* irs;    // F*nn*d  rsqrt(var)
  double* cm;     // N*d  per-sample conditional mean of the current label draw
  double* cv;     // N*d  per-sample conditional variance
  double* us;     // N*IIF_GIBBS_RND_PHASES  the level's uniforms of every owned sample (two-density products)
  double* zs;     // N*d  normals of the next samplePoint
  double* post;   // N*d
  double* xa;     // N
  double* xb;     // loo_x2_doubles(N)
  double* red;    // IIF_RED_DOUBLES
  double* scr;    // IIF_LOO_SCRATCH_N(N)  (leave-one-out scratch)
  double* bwk;    // F*IIF_MAX_DIM kernel bandwidths
  double* wt;     // nn node weights (count / N), shared by all densities
  double* minvar; // F*(L+1)*d smallest node variance per level
  const double* tab;  // 16-entry 2^(j/16) table
  int16_t* perm;  // 2*F*N
};

// the label draws keep one total per piece of four candidates and output sample in the scratch region that the
// leave-one-out search uses afterwards
#define IIF_GIBBS_FUSED_MAX 128  // largest N whose four tables of piece totals fit in shared memory
__host__ __device__ inline bool gibbs_fused(int F, int N) { return F == 2 && N <= IIF_GIBBS_FUSED_MAX; }
#define IIF_GIBBS_RND_PHASES 6   // staged uniforms per sample and level: F (1 + Niter) <= 6 for two densities, Niter <= 2
__host__ __device__ inline size_t prod_scratch_doubles(int F, int N) {
  // two densities build the level's four tables of piece totals at once
  const size_t loo = (size_t)IIF_LOO_SCRATCH_N(N), pt = (size_t)(gibbs_fused(F, N) ? 4 : 1) * N * (size_t)(((N + 3) >> 2) | 1);
  return loo > pt ? loo : pt;
}

__host__ __device__ inline size_t prod_smem_bytes(int F, int N, int d, int nn, int L) {
  size_t dbl = (size_t)F * N * d + 3 * (size_t)F * nn * d + 4 * (size_t)N * d + (size_t)N * IIF_GIBBS_RND_PHASES + (size_t)loo_xa_doubles(N) +
               (size_t)loo_x2_doubles(N) + IIF_RED_DOUBLES + (size_t)F * IIF_MAX_DIM + (size_t)nn + (size_t)F * (L + 1) * d;
  dbl += prod_scratch_doubles(F, N);
  size_t i16 = 2 * (size_t)F * N + (size_t)(L + 2) + 3 * (size_t)nn + (size_t)L * N;  // permutations + tree structure
  return dbl * sizeof(double) + ((i16 * sizeof(int16_t) + 7) / 8) * 8;
}

// Product of the Gaussians selected in densities != skip along coordinate c
// (AMP getManiMu/getManiLam: Euclid precision-weighted mean; Circular precision-weighted atan2 mean).
__device__ __forceinline__ double cond_gauss(int F, const double* mean, const double* var, const int* node,
                                             const int32_t* masks, int skip, int nn, int d, int c, bool circ,
                                             double& mu) {
  double lam = 0, a = 0, sn = 0, cs = 0;
  for (int k = 0; k < F; ++k) {
    if (k == skip || !((masks[k] >> c) & 1)) continue;
    double l = 1.0 / var[(k * nn + node[k]) * d + c];
    double m = mean[(k * nn + node[k]) * d + c];
    lam += l;
    if (circ) { sn += l * sin(m); cs += l * cos(m); } else a += l * m;
  }
  if (lam > 0) mu = circ ? atan2(sn, cs) : a / lam;
  return lam;
}

// |a| clamped to 510 by integer min on the high word (the low word keeps its bits: |result| < 510.0001).  The Gaussian
// kernel value of a clamped argument is below 1e-300, i.e. nothing, and the lock-step evaluation stays on its fast path.
__device__ __forceinline__ double clamp_gauss_arg(double a) {
  return __hiloint2double(min(__double2hiint(a) & 0x7fffffff, IIF_GCLAMP_HI), __double2loint(a));
}

// Weights of the four candidate nodes [zb, zb+4) of one density at one level given a Gaussian (m, cv) per coordinate
// (see the Gibbs section of the kernel).  mean / var / irs / wt point at the level's first node.  POINT: cv == 0
// (sampleIndices), the stored rsqrt(var) is used; LEAF: every candidate has the same variance, the conditional rsqrt
// is hoisted.  One-coordinate Euclid beliefs (the chain benchmark):
template <bool POINT, bool LEAF>
__device__ __forceinline__ void gibbs_piece_d1(const double* __restrict__ mean, const double* __restrict__ var,
                                               const double* __restrict__ irs, const double* __restrict__ wt, int nz,
                                               int zb, double m0, double cv0, const double* __restrict__ tab,
                                               double (&w)[4]) {
  double a[4], pre[4], e[4];
  const double rsl = (!POINT && LEAF) ? rsqrt(var[0] + cv0) : 0.0;
  if (zb + 4 <= nz) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int z = zb + u;
      const double rs = POINT ? irs[z] : LEAF ? rsl : rsqrt(var[z] + cv0);
      a[u] = clamp_gauss_arg((mean[z] - m0) * (rs * IIF_GSCALE));
      pre[u] = wt[z] * rs;
    }
  } else {  // last piece of the level: clamp the index, padding lanes weigh nothing
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int z = min(zb + u, nz - 1);
      const double rs = POINT ? irs[z] : LEAF ? rsl : rsqrt(var[z] + cv0);
      a[u] = clamp_gauss_arg((mean[z] - m0) * (rs * IIF_GSCALE));
      pre[u] = (zb + u < nz) ? wt[z] * rs : 0.0;
    }
  }
  gauss_negU<4>(a, tab, e);
#pragma unroll
  for (int u = 0; u < 4; ++u) w[u] = e[u] * pre[u];
}

// general form: several coordinates, circular coordinates, partial masks
__device__ __forceinline__ void gibbs_piece_nd(const double* __restrict__ mean, const double* __restrict__ var,
                                               const double* __restrict__ irs, const double* __restrict__ wt, int nz, int zb,
                                               int d, int32_t cmask, int32_t hasmask, bool point, bool leaf,
                                               const double* __restrict__ m, const double* __restrict__ cv, double (&w)[4]) {
  double rsl[IIF_MAX_DIM] = {0, 0, 0, 0};
  if (!point && leaf)
    for (int c = 0; c < d; ++c) rsl[c] = rsqrt(var[c] + cv[c]);
  double arg[4], pre[4], e[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int z = min(zb + u, nz - 1);
    double q = 0.0, pr = (zb + u < nz) ? wt[z] : 0.0;
    for (int c = 0; c < d; ++c) {
      if (!((hasmask >> c) & 1)) continue;
      const double rs = point ? irs[z * d + c] : leaf ? rsl[c] : rsqrt(var[z * d + c] + cv[c]);
      const double tq = mdiff(mean[z * d + c], m[c], is_circ(cmask, c)) * rs;
      q = fma(tq, tq, q);
      pr *= rs;
    }
    arg[u] = fmax(-0.5 * q, -700.0);
    pre[u] = pr;
  }
  exp_negU<4>(arg, e);
#pragma unroll
  for (int u = 0; u < 4; ++u) w[u] = e[u] * pre[u];
}

// runtime dispatch of one piece
__device__ __forceinline__ void gibbs_piece(bool d1, const double* mean, const double* var, const double* irs,
                                            const double* wt, int nz, int zb, int d, int32_t cmask, int32_t hasmask,
                                            bool point, bool leaf, const double* m, const double* cv, const double* tab,
                                            double (&w)[4]) {
  if (d1 && (hasmask & 1)) {
    if (point) gibbs_piece_d1<true, false>(mean, var, irs, wt, nz, zb, m[0], 0.0, tab, w);
    else if (leaf) gibbs_piece_d1<false, true>(mean, var, irs, wt, nz, zb, m[0], cv[0], tab, w);
    else gibbs_piece_d1<false, false>(mean, var, irs, wt, nz, zb, m[0], cv[0], tab, w);
  } else {
    gibbs_piece_nd(mean, var, irs, wt, nz, zb, d, cmask, hasmask, point, leaf, m, cv, w);
  }
}

// piece totals of `rows` rows x np pieces into PT (row stride NPs); row r has the conditional Gaussian
// (mrow + r * mstride, cvrow + r * mstride).  All threads of the CTA; the mode is uniform, so the loop is specialised.
template <bool POINT, bool LEAF>
__device__ __forceinline__ void gibbs_build_d1(const double* mean, const double* var, const double* irs, const double* wt,
                                               int nz, int np, int rows, const double* mrow, const double* cvrow,
                                               int mstride, const double* tab, double* PT, int NPs) {
  // consecutive lanes take consecutive ROWS of the same piece: the candidates' node data is a broadcast read, the
  // rows' conditional parameters are contiguous, and the totals land on distinct banks (odd row stride)
  const float inv_rows = 1.0f / (float)rows;
  for (int p = threadIdx.x; p < rows * np; p += IIF_NT) {
    const int pc = (int)(((float)p + 0.5f) * inv_rows), row = p - pc * rows;   // exact: p < 2^14
    double w[4];
    gibbs_piece_d1<POINT, LEAF>(mean, var, irs, wt, nz, pc << 2, mrow[row * mstride], POINT ? 0.0 : cvrow[row * mstride], tab, w);
    PT[row * NPs + pc] = (w[0] + w[1]) + (w[2] + w[3]);
  }
}
__device__ __forceinline__ void gibbs_build(bool d1, const double* mean, const double* var, const double* irs,
                                            const double* wt, int nz, int np, int rows, int d, int32_t cmask,
                                            int32_t hasmask, bool point, bool leaf, const double* mrow,
                                            const double* cvrow, const double* tab, double* PT, int NPs) {
  if (d1 && (hasmask & 1)) {
    if (point) gibbs_build_d1<true, false>(mean, var, irs, wt, nz, np, rows, mrow, cvrow, 1, tab, PT, NPs);
    else if (leaf) gibbs_build_d1<false, true>(mean, var, irs, wt, nz, np, rows, mrow, cvrow, 1, tab, PT, NPs);
    else gibbs_build_d1<false, false>(mean, var, irs, wt, nz, np, rows, mrow, cvrow, 1, tab, PT, NPs);
    return;
  }
  const float inv_rows = 1.0f / (float)rows;
  for (int p = threadIdx.x; p < rows * np; p += IIF_NT) {
    const int pc = (int)(((float)p + 0.5f) * inv_rows), row = p - pc * rows;
    double w[4];
    gibbs_piece_nd(mean, var, irs, wt, nz, pc << 2, d, cmask, hasmask, point, leaf, mrow + row * d, cvrow + row * d, w);
    PT[row * NPs + pc] = (w[0] + w[1]) + (w[2] + w[3]);
  }
}

// Label pick when every weight of a draw underflowed: p_z = sum_c dl^2 / v + log v, weights exp(-(p_z - min p)/2) wt_z,
// sequential inverse CDF — the form the oracle always uses.  Rare (a sample hundreds of bandwidths from every node).
__device__ __noinline__ int gibbs_pick_exact(const double* mean, const double* var, const double* wt, int nz, int d,
                                             int32_t cmask, int32_t hasmask, const double* m, const double* cv, double u) {
  auto expo = [&](int z) {
    double p = 0.0;
    for (int c = 0; c < d; ++c) {
      if (!((hasmask >> c) & 1)) continue;
      const double dl = mdiff(mean[z * d + c], m[c], is_circ(cmask, c));
      const double v = var[z * d + c] + cv[c];
      p += dl * dl / v + log(v);
    }
    return p;
  };
  double best = INFINITY;
  for (int z = 0; z < nz; ++z) best = fmin(best, expo(z));
  double tot = 0.0;
  for (int z = 0; z < nz; ++z) tot += exp(-0.5 * (expo(z) - best)) * wt[z];
  const double thr = u * tot;
  double cum = 0.0;
  for (int z = 0; z < nz; ++z) {
    cum += exp(-0.5 * (expo(z) - best)) * wt[z];
    if (thr < cum) return z;
  }
  return nz - 1;
}

__global__ void __launch_bounds__(IIF_MAX_THREADS, 1)
iif_product_kernel(DeviceGraph g, const ProdTask* __restrict__ tasks, const double* __restrict__ randU,
                   const double* __restrict__ randN, const TreeStruct* __restrict__ trees) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IIF_PHASE_ZERO();
  IIF_PHASE_BEGIN();
  // narrow launches: a cluster of CTAs per product runs redundantly and shares the bandwidth search; rank 0 writes
  const int cC = (int)cooperative_groups::this_cluster().num_blocks();
  const bool wr = cooperative_groups::this_cluster().block_rank() == 0;
  if (cC > 1) cooperative_groups::this_cluster().sync();  // every rank is resident before any remote shared-memory access
  const ProdTask t = tasks[blockIdx.x / cC];
  const int F = t.F, N = t.N, d = t.dim;
  const int32_t cm = t.circ_mask;
  TreeStruct T = trees[N];
  const int nn = T.nn, L = T.L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int parity = 0;

  // upstream failure: propagate the status and leave the destination untouched
  if (t.conv_status != nullptr) {
    int st = IIF_OK;
    for (int j = 0; j < F; ++j) if (t.conv_status[j] != IIF_OK) st = t.conv_status[j];
    if (st != IIF_OK) {
      if (tid == 0 && t.out_status && wr) *t.out_status = st;
      return;
    }
  }

  ProdSmem sm;
  {
    double* p = reinterpret_cast<double*>(smem_raw);
    sm.P = p; p += (size_t)F * N * d;
    sm.mean = p; p += (size_t)F * nn * d;
    sm.var = p; p += (size_t)F * nn * d;
    sm.irs = p; p += (size_t)F * nn * d;
    sm.cm = p; p += (size_t)N * d;
    sm.cv = p; p += (size_t)N * d;
    sm.us = p; p += (size_t)N * IIF_GIBBS_RND_PHASES;
    sm.zs = p; p += (size_t)N * d;
    sm.post = p; p += (size_t)N * d;
    sm.xa = p; p += loo_xa_doubles(N);
    sm.xb = p; p += loo_x2_doubles(N);
    sm.red = p; p += IIF_RED_DOUBLES;
    sm.bwk = p; p += F * IIF_MAX_DIM;
    sm.wt = p; p += nn;
    sm.minvar = p; p += (size_t)F * (L + 1) * d;
    sm.scr = p;  // last: leave-one-out scratch, used for the label draws' piece totals before
    p += prod_scratch_doubles(F, N);
    sm.perm = reinterpret_cast<int16_t*>(p);
  }
  {
    // the tree structure (level lists, node ranges, children) is walked at every level by every thread: stage
    // the whole blob (lev_off | lo | hi | child | node_at, contiguous in global memory) in shared memory
    int16_t* ts = sm.perm + 2 * (size_t)F * N;
    const int16_t* gsrc = T.lev_off;
    const int tot = (L + 2) + 3 * nn + L * N;
    for (int i = tid; i < tot; i += IIF_NT) ts[i] = gsrc[i];
    T.lev_off = ts;
    T.lo = ts + (L + 2);
    T.hi = T.lo + nn;
    T.child = T.hi + nn;
    T.node_at = T.child + nn;
  }
  const int32_t fullmask = (1 << d) - 1;
  __shared__ int32_t masks[IIF_MAX_FACTORS];
  __shared__ double gtab[IIF_GTAB_N];  // 2^(j/256) for gauss_negU
  gauss_stage_table(gtab);
  sm.tab = gtab;
  if (tid < F) masks[tid] = t.mask[tid] ? t.mask[tid] : fullmask;
  for (int i = tid; i < F * N * d; i += IIF_NT) sm.P[i] = t.dens_pts[i];
  for (int i = tid; i < F * IIF_MAX_DIM; i += IIF_NT) sm.bwk[i] = t.dens_bw[i];
  __syncthreads();

  double bw[IIF_MAX_DIM] = {0, 0, 0, 0};
  const bool passthrough = (F == 1 && masks[0] == fullmask);
  if (passthrough) {
    // manifoldProduct of one density returns it unchanged (no Gibbs, no re-bandwidth)
    for (int i = tid; i < N * d; i += IIF_NT) sm.post[i] = sm.P[i];
    for (int c = 0; c < d; ++c) bw[c] = sm.bwk[c];
    if (t.out_labels && wr) for (int s = tid; s < N; s += IIF_NT) t.out_labels[s] = s;
    __syncthreads();
  } else {
    IIF_PHASE(8);
    // ---- 1. ball trees: per level, rank-sort every node along its most-spread coordinate
    int16_t* permA = sm.perm;
    int16_t* permB = sm.perm + F * N;
    for (int i = tid; i < F * N; i += IIF_NT) permA[i] = (int16_t)(i % N);
    __syncthreads();
    // one coordinate: the root's rank sort already orders every descendant node (same key, same tie-break),
    // so the deeper levels would be identity permutations
    const int Lsort = (d == 1) ? 1 : L;
    for (int l = 0; l < Lsort; ++l) {
      for (int it = tid; it < F * N; it += IIF_NT) {
        const int j = it / N, pos = it - j * N;
        const int z = T.lev_off[l] + T.node_at[l * N + pos];
        const int lo = T.lo[z], hi = T.hi[z];
        const int16_t* pj = permA + j * N;
        const double* Pj = sm.P + (size_t)j * N * d;
        const int me = pj[pos];
        if (lo == hi) { permB[j * N + pos] = (int16_t)me; continue; }
        int best = (d == 1) ? 0 : -1;
        double bs = -1.0;
        for (int c = 0; c < d && d > 1; ++c) {
          if (!((masks[j] >> c) & 1)) continue;
          double mn = INFINITY, mx = -INFINITY;
          for (int i = lo; i <= hi; ++i) {
            double v = Pj[pj[i] * d + c];
            mn = fmin(mn, v);
            mx = fmax(mx, v);
          }
          if (mx - mn > bs) { bs = mx - mn; best = c; }
        }
        const double vi = Pj[me * d + best];
        int r = 0;
        for (int i = lo; i <= hi; ++i) {
          const int o = pj[i];
          const double vo = Pj[o * d + best];
          r += (vo < vi) || (vo == vi && o < me);
        }
        permB[j * N + lo + r] = (int16_t)me;
      }
      __syncthreads();
      int16_t* tmp = permA; permA = permB; permB = tmp;
    }
    IIF_PHASE(9);
    // ---- node statistics: mean and (kernel variance + member spread) per level-list entry.
    // Two-pass mean / squared deviation; nodes with more than 32 members are reduced by a whole warp
    // (lanes stride the members), the rest by one thread each.
    {
      const int nbig = (nn > 0) ? min(nn, 8) : 0;  // candidates: the first entries of the list are the large nodes
      // warp tasks: (density, big node, coordinate)
      for (int task = warp; task < F * nbig * d; task += IIF_NW) {
        const int c = task % d, z = (task / d) % nbig, j = task / (d * nbig);
        const int lo = T.lo[z], hi = T.hi[z], cnt = hi - lo + 1;
        if (cnt <= 32) continue;
        const int16_t* pj = permA + j * N;
        const double* Pj = sm.P + (size_t)j * N * d;
        double s1 = 0;
        for (int i = lo + lane; i <= hi; i += 32) s1 += Pj[pj[i] * d + c];
        s1 = warp_sum(s1);
        const double m = s1 / cnt;
        double q = 0;
        for (int i = lo + lane; i <= hi; i += 32) { double e = Pj[pj[i] * d + c] - m; q += e * e; }
        q = warp_sum(q);
        if (lane == 0) {
          const double h = sm.bwk[j * IIF_MAX_DIM + c];
          sm.mean[((size_t)j * nn + z) * d + c] = m;
          sm.var[((size_t)j * nn + z) * d + c] = h * h + q / cnt;
        }
      }
      for (int it = tid; it < F * nn; it += IIF_NT) {
        const int j = it / nn, z = it - j * nn;
        const int lo = T.lo[z], hi = T.hi[z], cnt = hi - lo + 1;
        if (z < nbig && cnt > 32) continue;  // done by a warp above
        const int16_t* pj = permA + j * N;
        const double* Pj = sm.P + (size_t)j * N * d;
        for (int c = 0; c < d; ++c) {
          double s1 = 0;
          for (int i = lo; i <= hi; ++i) s1 += Pj[pj[i] * d + c];
          const double m = s1 / cnt;
          double q = 0;
          for (int i = lo; i <= hi; ++i) { double e = Pj[pj[i] * d + c] - m; q += e * e; }
          const double h = sm.bwk[j * IIF_MAX_DIM + c];
          sm.mean[(size_t)it * d + c] = m;
          sm.var[(size_t)it * d + c] = h * h + q / cnt;
        }
      }
    }
    for (int z = tid; z < nn; z += IIF_NT) sm.wt[z] = (double)(T.hi[z] - T.lo[z] + 1) / (double)N;
    __syncthreads();

    // reciprocal standard deviations of every node (the label draws below multiply instead of dividing)
    for (int i = tid; i < F * nn * d; i += IIF_NT) sm.irs[i] = rsqrt(sm.var[i]);
    __syncthreads();
    IIF_PHASE(10);
    // ---- 2./3. multiscale Gibbs (KDE.jl gibbs1; Ihler, Sudderth, Freeman & Willsky 2003), per output sample:
    //   labels := roots; for every level: X = samplePoint(product of the selected nodes), levelDown,
    //   sampleIndices(X) = every density's label given X, then Niter sweeps of sampleIndex(j) = label of density j
    //   given the other densities' selected nodes; the sample is samplePoint(selected leaf kernels).
    // Every label draw is an inverse-CDF pick over the nz candidate nodes of the level with weights
    //   w_z = wt_z prod_c rsqrt(v_zc) exp(-1/2 sum_c (mean_zc - m_c)^2 / v_zc),   v_zc = var_zc + cv_c,
    // (m, cv) = (X, 0) for sampleIndices and the product Gaussian of the other densities' nodes for sampleIndex.
    // Mapping: thread i < nloc OWNS output sample s (labels in registers, the random draws, the pick).  The N x nz
    // weight evaluations of a draw — the expensive part — are dealt flat to ALL threads of the CTA in pieces of four
    // candidates (one lock-step gauss_negU / exp_negU group); only the piece totals are stored (in the idle
    // leave-one-out scratch).  The owner walks its row of totals, finds the piece holding u * total and re-evaluates
    // that piece alone.  In a cluster launch the output samples are dealt round-robin to the ranks.
    const int niter = g.sp->gibbsNiter;
    const uint64_t seed = g.sp->seed;
    const uint32_t call = (uint32_t)t.call_id;
    const int crank = (int)cooperative_groups::this_cluster().block_rank();
    const int nloc = (N - crank + cC - 1) / cC;
    const bool own = tid < nloc;
    const int s = tid * cC + crank;
    int node[IIF_MAX_FACTORS];
    for (int j = 0; j < F; ++j) node[j] = 0;  // levelInit / initIndices: roots (their uniforms are skipped)
    // Random-stream layout (AMP.manifoldProduct `_randU` / `_randN`; KDE.jl sizes them Np*Ndens*(Niter+2)*Nlevels and
    // Ndim*Np*(Nlevels+1)): per output sample the uniforms go to initIndices (F), then per level to sampleIndices (F)
    // and Niter sweeps of sampleIndex (F each); the normals are the Nlevels+1 samplePoint draws (d each).
    const uint32_t ublk = (uint32_t)(F * (1 + L * (niter + 1))), nblk = (uint32_t)(d * (L + 1));
    auto gibbs_u = [&](uint32_t idx) -> double {
      return (randU != nullptr && t.randu_off >= 0) ? randU[t.randu_off + idx] : rs_uniform(seed, call, IIF_RS_GIBBS_U, idx);
    };
    auto gibbs_n = [&](uint32_t idx) -> double {
      return (randN != nullptr && t.randn_off >= 0) ? randN[t.randn_off + idx] : rs_normal(seed, call, IIF_RS_GIBBS_N, idx);
    };
    const int NP = ((N + 3) >> 2) | 1;   // row stride of the piece totals (odd: rows on distinct banks)
    const bool d1 = (d == 1) && !is_circ(cm, 0);
    int32_t anyother[IIF_MAX_FACTORS];  // coordinates some OTHER density informs (uniform over the samples)
    for (int j = 0; j < F; ++j) {
      int32_t m = 0;
      for (int k = 0; k < F; ++k) if (k != j) m |= masks[k];
      anyother[j] = m;
    }
    // inverse-CDF pick in a row of piece totals: four interleaved partial sums locate the quarter, a short walk the
    // piece, `eval` re-evaluates that piece's four weights (same function, same arguments => same bits as the build)
    auto pick_row = [&](const double* rowp, int np, int nz, double u, auto&& eval, auto&& exact) -> int {
      const int ch = (np + 3) >> 2;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int k = 0; k < ch; ++k) {
        s0 += rowp[k];
        if (k + ch < np) s1 += rowp[k + ch];
        if (k + 2 * ch < np) s2 += rowp[k + 2 * ch];
        if (k + 3 * ch < np) s3 += rowp[k + 3 * ch];
      }
      const double c1 = s0 + s1, c2 = c1 + s2, tot = c2 + s3;
      IIF_PHASE(21);
      if (!(tot > 1e-280)) return exact();   // every weight underflowed (or NaN): the oracle's exact form
      const double thr = u * tot;
      double run = 0.0;
      int pc = 0;
      if (thr >= s0) { run = s0; pc = ch; }
      if (thr >= c1) { run = c1; pc = 2 * ch; }
      if (thr >= c2) { run = c2; pc = 3 * ch; }
      pc = min(pc, np - 1);
      for (; pc < np - 1; ++pc) {
        const double nx = run + rowp[pc];
        if (thr < nx) break;
        run = nx;
      }
      IIF_PHASE(22);
      double w[4];
      eval(pc << 2, w);
      IIF_PHASE(23);
      int pick = min((pc << 2) + 3, nz - 1);  // rounding left thr >= total: the last candidate (as the oracle does)
      double cum = run;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        cum += w[q];
        if ((pc << 2) + q < nz && thr < cum) { pick = (pc << 2) + q; break; }
      }
      IIF_PHASE(24);
      return pick;
    };
    // per-level random numbers of the owned sample, drawn by ALL threads during the build pass
    const int nph = F * (1 + niter);
    const bool fused = gibbs_fused(F, N);
    const bool rnd_staged = fused && nph <= IIF_GIBBS_RND_PHASES;
    if (rnd_staged) {
      for (int r = tid; r < nloc * d; r += IIF_NT) {  // normals of the first samplePoint
        const int i = r / d, c = r - i * d;
        sm.zs[r] = gibbs_n((uint32_t)(i * cC + crank) * nblk + (uint32_t)c);
      }
      __syncthreads();
    }
    for (int l = 1; l <= L; ++l) {
      const int z0 = T.lev_off[l], nz = T.lev_off[l + 1] - z0, np = (nz + 3) >> 2;
      const bool leaf = (l == L);
      if (own) {  // samplePoint from the product of the nodes selected at the coarser level
        for (int c = 0; c < d; ++c) {
          double mu = 0;
          const double lam = cond_gauss(F, sm.mean, sm.var, node, masks, -1, nn, d, c, is_circ(cm, c), mu);
          const double e = rnd_staged ? sm.zs[tid * d + c] : gibbs_n((uint32_t)s * nblk + (uint32_t)((l - 1) * d + c));
          sm.cm[tid * d + c] = lam > 0 ? madd(mu, sqrt(1.0 / lam) * e, is_circ(cm, c)) : 0.0;
          sm.cv[tid * d + c] = 0.0;
        }
      }
      IIF_PHASE(20);
      // levelDown: the level lists become the children of the previous lists (leaves stay); every label is re-drawn
      if (fused) {
        // ---- two densities: none of the level's four weight tables depends on a label drawn at this level — the
        // sampleIndices rows are the samples (given X), the sampleIndex rows of density j are the nz candidate nodes of
        // the OTHER density (its selected node is the whole condition) — so all piece totals are built in one flat
        // pass between two barriers, together with the level's uniforms and the next samplePoint's normals.
        __syncthreads();  // X of every sample visible; the previous level's picks are done
        IIF_PHASE(17);
        const size_t ptq = (size_t)N * NP;   // table q at sm.scr + q * ptq
        for (int q = 0; q < 4; ++q) {
          const int j = q & 1;
          const bool point = q < 2;
          const int32_t hasmask = masks[j] & (point ? fullmask : anyother[j]);
          const size_t bj = ((size_t)j * nn + z0) * d, bo = ((size_t)(1 - j) * nn + z0) * d;
          gibbs_build(d1, sm.mean + bj, sm.var + bj, sm.irs + bj, sm.wt + z0, nz, np, point ? nloc : nz, d, cm, hasmask, point,
                      leaf, point ? sm.cm : sm.mean + bo, point ? sm.cv : sm.var + bo, sm.tab, sm.scr + q * ptq, NP);
        }
        if (rnd_staged) {
          for (int r = tid; r < nloc * nph; r += IIF_NT) {
            const int i = r / nph, ph = r - i * nph;
            sm.us[r] = gibbs_u((uint32_t)(i * cC + crank) * ublk + (uint32_t)(F + (l - 1) * nph + ph));
          }
          for (int r = tid; r < nloc * d; r += IIF_NT) {
            const int i = r / d, c = r - i * d;
            sm.zs[r] = gibbs_n((uint32_t)(i * cC + crank) * nblk + (uint32_t)(l * d + c));
          }
        }
        IIF_PHASE(7);
        __syncthreads();
        IIF_PHASE(18);
        if (own) {
          int nd0 = node[0], nd1 = node[1];   // scalars: no dynamically indexed local array in the hot path
          for (int phase = 0; phase < nph; ++phase) {
            const int j = phase & 1;
            const bool point = phase < 2;
            const int32_t hasmask = masks[j] & (point ? fullmask : anyother[j]);
            const size_t bj = ((size_t)j * nn + z0) * d;
            const double* mj = sm.mean + bj;
            const double* vj = sm.var + bj;
            const double* rj = sm.irs + bj;
            const double* wl = sm.wt + z0;
            const int other = j ? nd0 : nd1;
            const int row = point ? tid : other - z0;
            const double* mrow = point ? sm.cm + tid * d : sm.mean + (size_t)other * d + (size_t)(1 - j) * nn * d;
            const double* cvrow = point ? sm.cv + tid * d : sm.var + (size_t)other * d + (size_t)(1 - j) * nn * d;
            const double u = rnd_staged ? sm.us[tid * nph + phase] : gibbs_u((uint32_t)s * ublk + (uint32_t)(F + (l - 1) * nph + phase));
            const int pick = pick_row(
                sm.scr + (point ? j : 2 + j) * ptq + (size_t)row * NP, np, nz, u,
                [&](int zb, double (&w)[4]) { gibbs_piece(d1, mj, vj, rj, wl, nz, zb, d, cm, hasmask, point, leaf, mrow, cvrow, sm.tab, w); },
                [&]() { return gibbs_pick_exact(mj, vj, wl, nz, d, cm, hasmask, mrow, cvrow, u); });
            if (j) nd1 = z0 + pick; else nd0 = z0 + pick;
          }
          node[0] = nd0;
          node[1] = nd1;
        }
        IIF_PHASE(15);
        continue;
      }
      // ---- any number of densities: one draw after the other; the conditional of sampleIndex(j) (product of the
      // other densities' selected nodes) is per sample, staged in shared memory by the owner
      double* PT = sm.scr;
      for (int phase = 0; phase < nph; ++phase) {
        const int j = phase % F;
        const bool point = phase < F;   // sampleIndices (given X) first, then the sampleIndex sweeps
        const int32_t hasmask = masks[j] & (point ? fullmask : anyother[j]);
        if (!point && own) {
          for (int c = 0; c < d; ++c) {
            double mu = 0;
            const double lam = ((hasmask >> c) & 1) ? cond_gauss(F, sm.mean, sm.var, node, masks, j, nn, d, c, is_circ(cm, c), mu) : 0.0;
            sm.cm[tid * d + c] = mu;
            sm.cv[tid * d + c] = lam > 0 ? 1.0 / lam : 0.0;
          }
        }
        IIF_PHASE(16);
        __syncthreads();  // (cm, cv) of every sample visible; the previous draw's totals are no longer read
        IIF_PHASE(17);
        const size_t bj = ((size_t)j * nn + z0) * d;
        const double* mj = sm.mean + bj;
        const double* vj = sm.var + bj;
        const double* rj = sm.irs + bj;
        const double* wl = sm.wt + z0;
        gibbs_build(d1, mj, vj, rj, wl, nz, np, nloc, d, cm, hasmask, point, leaf, sm.cm, sm.cv, sm.tab, PT, NP);
        IIF_PHASE(7);
        __syncthreads();
        IIF_PHASE(18);
        if (own) {
          const double u = gibbs_u((uint32_t)s * ublk + (uint32_t)(F + (l - 1) * nph + phase));
          const double* mrow = sm.cm + tid * d;
          const double* cvrow = sm.cv + tid * d;
          const int pick = pick_row(
              PT + (size_t)tid * NP, np, nz, u,
              [&](int zb, double (&w)[4]) { gibbs_piece(d1, mj, vj, rj, wl, nz, zb, d, cm, hasmask, point, leaf, mrow, cvrow, sm.tab, w); },
              [&]() { return gibbs_pick_exact(mj, vj, wl, nz, d, cm, hasmask, mrow, cvrow, u); });
          node[j] = z0 + pick;
        }
        IIF_PHASE(15);
      }
    }
    IIF_PHASE(11);
    if (own) {
      // final samplePoint: draw from the product of the selected leaf kernels
      for (int c = 0; c < d; ++c) {
        double mu = 0;
        const double lam = cond_gauss(F, sm.mean, sm.var, node, masks, -1, nn, d, c, is_circ(cm, c), mu);
        double x;
        if (lam > 0) {
          const double e = rnd_staged ? sm.zs[tid * d + c] : gibbs_n((uint32_t)s * nblk + (uint32_t)(L * d + c));
          x = madd(mu, sqrt(1.0 / lam) * e, is_circ(cm, c));
        } else if (t.target_slot >= 0) {
          // coordinates no proposal informs keep oldPoints (GraphProductOperations.jl:37-45)
          const iif_slot_desc S = g.slots[t.target_slot];
          const int len = g.npts[t.target_slot];
          if (s < len) x = g.pts[S.pts_off + s * d + c];
          else if (len > 0) {
            double uu = rs_uniform(seed, call, IIF_RS_OLDPAD, (uint32_t)(s * (d + 1)));
            int k = min((int)(uu * len), len - 1);
            double e = rs_normal(seed, call, IIF_RS_OLDPAD, (uint32_t)(s * (d + 1) + 1 + c));
            x = madd(g.pts[S.pts_off + k * d + c], g.bw[t.target_slot * IIF_MAX_DIM + c] * e, is_circ(cm, c));
          } else x = 0.0;
        } else {
          x = t.old_pts ? t.old_pts[s * d + c] : 0.0;
        }
        if (cC > 1) {
          for (int r = 0; r < cC; ++r) cooperative_groups::this_cluster().map_shared_rank(sm.post, r)[s * d + c] = x;
        } else {
          sm.post[s * d + c] = x;
        }
      }
      if (t.out_labels != nullptr)   // every sample belongs to exactly one rank
        for (int j = 0; j < F; ++j) t.out_labels[s * F + j] = permA[j * N + T.lo[node[j]]];
    }
    if (cC > 1) cooperative_groups::this_cluster().sync();   // all ranks hold all posterior samples
    else __syncthreads();
    IIF_PHASE(12);
    // ---- 4. re-bandwidth of the posterior (getKDEManifoldBandwidths on the result)
    block_kde_bandwidth<1>(sm.post, N, d, cm, &T, sm.xa, sm.xb, sm.scr, sm.red, &parity, bw);
    IIF_PHASE(13);
  }

  // ---- outputs: explicit buffers and / or setBelief! into the destination slot
  if (!wr) return;
  if (t.out_pts != nullptr)
    for (int i = tid; i < N * d; i += IIF_NT) t.out_pts[i] = sm.post[i];
  if (t.out_bw != nullptr && tid < IIF_MAX_DIM) t.out_bw[tid] = tid < d ? bw[tid] : 0.0;
  if (t.out_slot >= 0) {
    const iif_slot_desc O = g.slots[t.out_slot];
    for (int i = tid; i < N * d; i += IIF_NT) g.pts[O.pts_off + i] = sm.post[i];
    if (tid < IIF_MAX_DIM) {
      g.bw[t.out_slot * IIF_MAX_DIM + tid] = tid < d ? bw[tid] : 0.0;
      g.ipc[t.out_slot * IIF_MAX_DIM + tid] = tid < d ? (double)F : 0.0;  // ApproxConv.jl:296-300
    }
    if (tid == 0) {
      g.npts[t.out_slot] = N;
      g.flags[t.out_slot] |= 1;
    }
  }
  if (tid == 0 && t.out_status) *t.out_status = IIF_OK;
  IIF_PHASE(14);
  IIF_PHASE_FLUSH();
}

// separator-message adoption: slot b := slot a (updateSubFgFromDownMsgs!, TreeMessageUtils.jl:66)
__global__ void iif_copy_kernel(DeviceGraph g, const int32_t* __restrict__ pairs, int npairs) {
  const int k = blockIdx.x;
  if (k >= npairs) return;
  const int a = pairs[2 * k], b = pairs[2 * k + 1];
  const iif_slot_desc A = g.slots[a], B = g.slots[b];
  const int n = g.npts[a];
  for (int i = threadIdx.x; i < n * A.dim; i += blockDim.x) g.pts[B.pts_off + i] = g.pts[A.pts_off + i];
  if (threadIdx.x < IIF_MAX_DIM) {
    g.bw[b * IIF_MAX_DIM + threadIdx.x] = g.bw[a * IIF_MAX_DIM + threadIdx.x];
    g.ipc[b * IIF_MAX_DIM + threadIdx.x] = g.ipc[a * IIF_MAX_DIM + threadIdx.x];
  }
  if (threadIdx.x == 0) {
    g.npts[b] = n;
    g.flags[b] = g.flags[a];
  }
}

// ---- multi-GPU separator messages inside the captured graph -------------------------------------------------
// A message = one belief slot copied into the SAME slot of a peer GPU's arena (mapped through CUDA IPC), followed by a
// flag raise in the peer's memory; the receiver's graph holds a WAIT node in front of the kernels that read the slot.
// Flags carry the replay's epoch (a device counter every rank advances once per replay), so they never need clearing.
struct PushTask {
  int32_t slot, _pad;
  double* r_pts;      // peer arena: points of the slot
  double* r_bw;       // peer arena: bandwidth row of the slot
  double* r_ipc;
  int32_t* r_npts;
  int32_t* r_flags;
  int32_t* r_msgflag; // peer flag of this message
};
__global__ void iif_push_kernel(DeviceGraph g, const PushTask* __restrict__ tasks, int ntasks, const int32_t* __restrict__ epoch) {
  const int k = blockIdx.x;
  if (k >= ntasks) return;
  const PushTask t = tasks[k];
  const iif_slot_desc S = g.slots[t.slot];
  const int n = g.npts[t.slot];
  for (int i = threadIdx.x; i < n * S.dim; i += blockDim.x) t.r_pts[i] = g.pts[S.pts_off + i];
  if (threadIdx.x < IIF_MAX_DIM) {
    t.r_bw[threadIdx.x] = g.bw[t.slot * IIF_MAX_DIM + threadIdx.x];
    t.r_ipc[threadIdx.x] = g.ipc[t.slot * IIF_MAX_DIM + threadIdx.x];
  }
  if (threadIdx.x == 0) {
    *t.r_npts = n;
    *t.r_flags = g.flags[t.slot];
  }
  __threadfence_system();   // every thread: its stores are visible system-wide before the barrier releases thread 0
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *(volatile int32_t*)t.r_msgflag = *epoch;
  }
}
__global__ void iif_wait_kernel(const int32_t* flags, const int32_t* __restrict__ ids, int n, const int32_t* __restrict__ epoch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t want = *epoch;
  const volatile int32_t* f = flags + ids[i];
  while (*f < want) __nanosleep(200);
  __threadfence_system();
}
__global__ void iif_epoch_kernel(int32_t* epoch) { *epoch += 1; }

// standalone bandwidth kernel (manikde! with bw === nothing on K point sets)
struct BwTask {
  const double* pts;
  double* out_bw;
  int32_t N, dim, circ_mask, _pad;
};
__global__ void __launch_bounds__(IIF_MAX_THREADS, 1)
iif_bandwidth_kernel(const BwTask* __restrict__ tasks, const TreeStruct* __restrict__ trees) {
  extern __shared__ __align__(16) double bw_smem[];  // conv_smem_bytes(N)
  __shared__ double red[IIF_RED_DOUBLES];
  const int cC = (int)cooperative_groups::this_cluster().num_blocks();
  const bool wr = cooperative_groups::this_cluster().block_rank() == 0;
  if (cC > 1) cooperative_groups::this_cluster().sync();  // every rank is resident before any remote shared-memory access
  const BwTask t = tasks[blockIdx.x / cC];
  int parity = 0;
  double* pts = bw_smem;
  double* xa = pts + (size_t)t.N * IIF_MAX_DIM;
  double* xb = xa + loo_xa_doubles(t.N);
  double* scr = xb + loo_x2_doubles(t.N);
  IIF_PHASE_ZERO();
  IIF_PHASE_BEGIN();
  for (int i = threadIdx.x; i < t.N * t.dim; i += IIF_NT) pts[i] = t.pts[i];
  __syncthreads();
  IIF_PHASE(8);
  double bw[IIF_MAX_DIM] = {0, 0, 0, 0};
  block_kde_bandwidth<2>(pts, t.N, t.dim, t.circ_mask, &trees[t.N], xa, xb, scr, red, &parity, bw);
  if (threadIdx.x < IIF_MAX_DIM && wr) t.out_bw[threadIdx.x] = threadIdx.x < t.dim ? bw[threadIdx.x] : 0.0;
  IIF_PHASE(13);
  IIF_PHASE_FLUSH();
}
