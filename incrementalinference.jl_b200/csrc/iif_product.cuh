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

// F == 2 products tabulate the level's pair-weight matrix (N^2 doubles at the leaf level) in the scratch
// region shared with the leave-one-out tile
#define IIF_GIBBS_TAB_MAX 150
__host__ __device__ inline bool gibbs_tab(int N) { return N <= IIF_GIBBS_TAB_MAX; }

__host__ __device__ inline size_t prod_smem_bytes(int F, int N, int d, int nn, int L) {
  size_t dbl = (size_t)F * N * d + 2 * (size_t)F * nn * d + (size_t)N * d + (size_t)loo_xa_doubles(N) + (size_t)loo_x2_doubles(N) + IIF_RED_DOUBLES +
               (size_t)F * IIF_MAX_DIM + (size_t)nn + (size_t)F * (L + 1) * d;
  size_t scr = (size_t)IIF_LOO_SCRATCH_N(N);
  if (F == 2 && gibbs_tab(N) && (size_t)N * N > scr) scr = (size_t)N * N;
  dbl += scr;
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
    sm.post = p; p += (size_t)N * d;
    sm.xa = p; p += loo_xa_doubles(N);
    sm.xb = p; p += loo_x2_doubles(N);
    sm.red = p; p += IIF_RED_DOUBLES;
    sm.bwk = p; p += F * IIF_MAX_DIM;
    sm.wt = p; p += nn;
    sm.minvar = p; p += (size_t)F * (L + 1) * d;
    sm.scr = p;  // last: loo scratch, aliased by the F == 2 pair-weight matrix (sized by prod_smem_bytes)
    {
      size_t scrn = (size_t)IIF_LOO_SCRATCH_N(N);
      if (F == 2 && gibbs_tab(N) && (size_t)N * N > scrn) scrn = (size_t)N * N;
      p += scrn;
    }
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
  __shared__ double gtab[16];  // 2^(j/16) for gauss_negU
  if (tid < 16) gtab[tid] = IIF_EXP2TAB[tid];
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

    IIF_PHASE(10);
    // ---- 2./3. multiscale Gibbs: G lanes per output sample (G = largest power of two <= threads/N).
    // Every lane owns a contiguous block of the level's candidate nodes, accumulates its weights in
    // chunks (kept in registers), a group scan locates the lane and chunk holding the inverse-CDF
    // crossing and only that chunk is re-evaluated.  Weights are taken relative to an analytic lower
    // bound of the exponent so a single pass suffices; the exact-minimum form (what the oracle
    // computes) is the fallback when every weight underflows.
    const int niter = g.sp->gibbsNiter;
    const uint64_t seed = g.sp->seed;
    const uint32_t call = (uint32_t)t.call_id;
    int G = 1;
    while (G < 32 && 2 * G * N <= IIF_NT) G <<= 1;  // small CTAs (wide waves): G = 1, least total work
    // in a cluster launch the output samples are dealt round-robin to the ranks (every rank builds the same
    // trees and tables, draws only its own samples and broadcasts their points before the bandwidth search)
    const int crank = (int)cooperative_groups::this_cluster().block_rank();
    const int smp = (tid / G) * cC + crank, gl = tid % G;   // sample, lane within the group
    const bool live = smp < N;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    // whole warps without a sample skip; partial groups never occur (G divides 32)
    const int s = smp;
    int node[IIF_MAX_FACTORS];
    for (int j = 0; j < F; ++j) node[j] = 0;  // roots
    const bool tab = (F == 2) && gibbs_tab(N);
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
    // samplePoint: a point from the product of the currently selected Gaussians (every lane of a group computes it)
    auto sample_point = [&](int l, double* X) {
      for (int c = 0; c < d; ++c) {
        double mu = 0;
        const double lam = cond_gauss(F, sm.mean, sm.var, node, masks, -1, nn, d, c, is_circ(cm, c), mu);
        X[c] = lam > 0 ? madd(mu, sqrt(1.0 / lam) * gibbs_n((uint32_t)s * nblk + (uint32_t)((l - 1) * d + c)), is_circ(cm, c)) : 0.0;
      }
    };
    // Label draw of density j at level l given a Gaussian (cmu, cvar) per coordinate (`has[c]`: coordinate takes part):
    // weight_z = wt_z * prod_c rsqrt(v) * exp(-1/2 sum_c dl^2 / v),  v = var_z + cvar  (the oracle's
    // exp(-(p_z - min p)/2) with p_z = sum dl^2/v + log v, up to the common factor).  sampleIndex(j) passes the product
    // of the other densities' selected nodes, sampleIndices passes the point X with cvar = 0.  The G lanes of the
    // sample's group own contiguous blocks of candidates, accumulate chunk sums in registers, a group scan locates the
    // lane and chunk holding the inverse-CDF crossing and only that chunk is re-evaluated.  Weights are relative to an
    // analytic lower bound of the exponent (single pass); exact-minimum fallback when every weight underflows.
    auto draw_label = [&](int j, int l, const double* cmu, const double* cvar, const bool* has, double u) -> int {
      const int z0 = T.lev_off[l], nz = T.lev_off[l + 1] - z0;
      const bool leaf = (l == L);
      const int B = (nz + G - 1) / G;
      const int zb = gl * B, ze = min(zb + B, nz);
      const int CH = (B + 15) >> 4;  // <= 16 chunks per lane; CH == 1 (B <= 16) needs no re-evaluation
      const double* mj = sm.mean + ((size_t)j * nn + z0) * d;
      const double* vj = sm.var + ((size_t)j * nn + z0) * d;
      const double* wj = sm.wt + z0;
      // at the leaf level v is the same for every candidate, so 1/v is hoisted and rsqrt(v) cancels
      double iv[IIF_MAX_DIM] = {0, 0, 0, 0};
      for (int c = 0; c < d; ++c)
        if (has[c] && leaf) iv[c] = 1.0 / (vj[c] + cvar[c]);
      auto cand = [&](int z, double& pre) -> double {  // exponent (>= 0) and prefactor of candidate z
        double p = 0.0;
        pre = wj[z];
        for (int c = 0; c < d; ++c) {
          if (!has[c]) continue;
          const double dl = mdiff(mj[z * d + c], cmu[c], is_circ(cm, c));
          if (leaf) p = fma(dl * dl, iv[c], p);
          else {
            const double rs = rsqrt(vj[z * d + c] + cvar[c]);
            p = fma(dl * dl, rs * rs, p);
            pre *= rs;
          }
        }
        return p;
      };
      auto weight = [&](int z, double base) -> double {
        double pre;
        const double p = cand(z, pre);
        return exp_neg(fmin(-0.5 * (p - base), 0.0)) * pre;
      };
      double ct[16];
      double Tl = 0.0, off = 0.0, tot = 0.0, base = 0.0;
      for (int attempt = 0; attempt < 2; ++attempt) {
        Tl = 0.0;
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          const int cb = zb + ch * CH;
          double sacc = 0.0;
          if (cb < ze) {
            const int ce = min(cb + CH, ze);
            for (int z = cb; z < ce; ++z) sacc += weight(z, base);
          }
          ct[ch] = sacc;
          Tl += sacc;
        }
        double inc = Tl;  // inclusive scan over the group's lanes
        for (int o = 1; o < G; o <<= 1) {
          const double y = __shfl_up_sync(gmask, inc, o, G);
          if (gl >= o) inc += y;
        }
        off = inc - Tl;
        tot = __shfl_sync(gmask, inc, G - 1, G);
        if (tot > 1e-280 || attempt == 1) break;
        // every weight underflowed: redo relative to the exact minimum exponent
        double pm = INFINITY;
        for (int z = zb; z < ze; ++z) { double pre; pm = fmin(pm, cand(z, pre)); }
        for (int o = G >> 1; o > 0; o >>= 1) pm = fmin(pm, __shfl_xor_sync(gmask, pm, o, G));
        base = pm;
      }
      const double thr = u * tot;
      // every lane searches its own chunks (no divergent owner path); exactly one lane finds the crossing
      int mypick = -1;
      if ((zb < ze) && (thr >= off) && (thr < off + Tl)) {
        double run = off;
        mypick = ze - 1;
        bool found = false;
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          const int cb = zb + ch * CH;
          if (!found && cb < ze) {
            if (thr < run + ct[ch]) {
              const int ce = min(cb + CH, ze);
              mypick = ce - 1;
              if (CH > 1) {
                double cum = run;
                for (int z = cb; z < ce; ++z) {
                  cum += weight(z, base);
                  if (thr < cum) { mypick = z; break; }
                }
              }
              found = true;
            }
            run += ct[ch];
          }
        }
      }
      const unsigned hit = __ballot_sync(gmask, mypick >= 0) & gmask;
      int pick = nz - 1;  // rounding left thr >= total: the oracle falls back to the last candidate
      if (hit) pick = __shfl_sync(gmask, mypick, (__ffs(hit) - 1) & (G - 1), G);
      return z0 + pick;
    };
    // sampleIndices(X): the label of every density given the point X, over the whole level list
    auto sample_indices = [&](int l, const double* X) {
      const double zero[IIF_MAX_DIM] = {0, 0, 0, 0};
      for (int j = 0; j < F; ++j) {
        bool has[IIF_MAX_DIM];
        for (int c = 0; c < IIF_MAX_DIM; ++c) has[c] = c < d && ((masks[j] >> c) & 1);
        node[j] = draw_label(j, l, X, zero, has, gibbs_u((uint32_t)s * ublk + (uint32_t)(F + (l - 1) * F * (niter + 1) + j)));
      }
    };
    if (tab) {
      // ---- F == 2: the conditional of one density depends only on the node selected in the other, so the
      // level's pair-weight matrix K[a][b] = prod_c rsqrt(va+vb) exp(-(ma-mb)^2 / 2(va+vb)) is tabulated ONCE
      // per level by the whole CTA (shared with the leave-one-out tile) and every sample's two label draws
      // become a weighted column / row scan of K — no transcendental per sample.
      double* K = sm.scr;
      const int32_t both = masks[0] & masks[1];
      const double* m0 = sm.mean;
      const double* m1 = sm.mean + (size_t)nn * d;
      const double* v0 = sm.var;
      const double* v1 = sm.var + (size_t)nn * d;
      for (int l = 1; l <= L; ++l) {
        const int z0 = T.lev_off[l], nz = T.lev_off[l + 1] - z0;
        // samplePoint from the product of the nodes selected at the coarser level; levelDown then replaces every
        // level list by its children (leaves stay) and sampleIndices re-draws every label over the new list
        double X[IIF_MAX_DIM] = {0, 0, 0, 0};
        if (live) sample_point(l, X);
        __syncthreads();  // the previous level's K is no longer read
        if (l == L && d == 1 && !is_circ(cm, 0)) {
          // leaf level, one Euclid coordinate: every node is a single kernel of variance h_j^2, so the table is
          // the Gaussian kernel matrix between the two point sets: rs exp(-(ma - mb)^2 rs^2 / 2), rs constant
          const int nzz = nz * nz;
          const float invnz = 1.0f / (float)nz;
          const double rs = rsqrt(v0[z0] + v1[z0]);
          const double sc = IIF_GSCALE * rs;
          for (int base = tid; base < nzz; base += 4 * IIF_NT) {
            double zz[4], e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int idx = min(base + u * IIF_NT, nzz - 1);
              const int a = (int)(((float)idx + 0.5f) * invnz), b = idx - a * nz;
              zz[u] = (m0[z0 + a] - m1[z0 + b]) * sc;
            }
            gauss_negU<4>(zz, sm.tab, e);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (base + u * IIF_NT < nzz) K[base + u * IIF_NT] = e[u] * rs;
          }
        } else {
          // four table entries per thread in lockstep (independent rsqrt / exp chains)
          const int nzz = nz * nz;
          const float invnz = 1.0f / (float)nz;
          for (int base = tid; base < nzz; base += 4 * IIF_NT) {
            double arg[4], pre[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int idx = min(base + u * IIF_NT, nzz - 1);
              const int a = (int)(((float)idx + 0.5f) * invnz), b = idx - a * nz;  // exact: nz <= 256
              double pexp = 0.0, pr = 1.0;
              for (int c = 0; c < d; ++c) {
                if (!((both >> c) & 1)) continue;
                const double dl = mdiff(m0[(z0 + a) * d + c], m1[(z0 + b) * d + c], is_circ(cm, c));
                const double rs = rsqrt(v0[(z0 + a) * d + c] + v1[(z0 + b) * d + c]);
                pexp = fma(dl * dl, rs * rs, pexp);
                pr *= rs;
              }
              arg[u] = -0.5 * pexp;
              pre[u] = pr;
            }
            double e[4];
            exp_negU<4>(arg, e);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (base + u * IIF_NT < nzz) K[base + u * IIF_NT] = e[u] * pre[u];
          }
        }
        __syncthreads();
        IIF_PHASE(7);
        if (live) {
          sample_indices(l, X);
          const double* wl = sm.wt + z0;
          const int B = (nz + G - 1) / G;
          const int zb = gl * B, ze = min(zb + B, nz);
          const int CHL = (B + 7) >> 3;  // every lane splits its block into <= 8 chunks of CHL candidates
          for (int it = 0; it < niter; ++it) {
            for (int j = 0; j < 2; ++j) {
              // density 0 scans column b of K (stride nz), density 1 scans row a (stride 1)
              const int other = node[1 - j] - z0;
              const double* Kb = (j == 0) ? K + other : K + (size_t)other * nz;
              const int stride = (j == 0) ? nz : 1;
              const double u = gibbs_u((uint32_t)s * ublk + (uint32_t)(2 + (l - 1) * 2 * (niter + 1) + 2 + it * 2 + j));
              IIF_PHASE(16);
              // chunk sums: eight independent accumulation chains
              double ct[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              for (int e = 0; e < CHL; ++e) {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                  const int z = zb + ch * CHL + e;
                  if (z < ze) ct[ch] = fma(wl[z], Kb[(size_t)z * stride], ct[ch]);
                }
              }
              const double Tl = ((ct[0] + ct[1]) + (ct[2] + ct[3])) + ((ct[4] + ct[5]) + (ct[6] + ct[7]));
              IIF_PHASE(17);
              double inc = Tl;
              for (int o = 1; o < G; o <<= 1) {
                const double y = __shfl_up_sync(gmask, inc, o, G);
                if (gl >= o) inc += y;
              }
              const double off = inc - Tl;
              const double tot = __shfl_sync(gmask, inc, G - 1, G);
              IIF_PHASE(18);
              int pick;
              if (tot > 1e-290) {
                const double thr = u * tot;
                int mypick = -1;
                if ((zb < ze) && (thr >= off) && (thr < off + Tl)) {
                  // the chunk holding the crossing, then only that chunk is walked again
                  double run = off;
                  int cb = zb;
                  bool go = true;
#pragma unroll
                  for (int ch = 0; ch < 7; ++ch) {
                    go = go && (thr >= run + ct[ch]) && (cb + CHL < ze);
                    run += go ? ct[ch] : 0.0;
                    cb += go ? CHL : 0;
                  }
                  const int ce = min(cb + CHL, ze);
                  mypick = ce - 1;
                  double cum = run;
                  for (int z = cb; z < ce; ++z) {
                    cum = fma(wl[z], Kb[(size_t)z * stride], cum);
                    if (thr < cum) { mypick = z; break; }
                  }
                }
                const unsigned hit = __ballot_sync(gmask, mypick >= 0) & gmask;
                pick = nz - 1;
                if (hit) pick = __shfl_sync(gmask, mypick, (__ffs(hit) - 1) & (G - 1), G);
              } else {
                // every pair weight underflowed: relative to the smallest exponent (what the oracle does) the
                // closest candidate carries all the mass
                double best = INFINITY;
                int bz = 0;
                for (int z = 0; z < nz; ++z) {
                  const int a = (j == 0) ? z : other, b = (j == 0) ? other : z;
                  double pexp = 0.0;
                  for (int c = 0; c < d; ++c) {
                    if (!((both >> c) & 1)) continue;
                    const double dl = mdiff(m0[(z0 + a) * d + c], m1[(z0 + b) * d + c], is_circ(cm, c));
                    const double v = v0[(z0 + a) * d + c] + v1[(z0 + b) * d + c];
                    pexp += dl * dl / v + log(v);
                  }
                  if (pexp < best) { best = pexp; bz = z; }
                }
                pick = bz;
              }
              IIF_PHASE(20);
              node[j] = z0 + pick;
            }
          }
        }
        IIF_PHASE(15);
      }
      __syncthreads();  // K (aliases the leave-one-out scratch) is dead from here on
    } else if (live) {
      for (int l = 1; l <= L; ++l) {
        double X[IIF_MAX_DIM] = {0, 0, 0, 0};
        sample_point(l, X);      // samplePoint, then levelDown (implicit) and sampleIndices over the new level list
        sample_indices(l, X);
        for (int it = 0; it < niter; ++it) {
          for (int j = 0; j < F; ++j) {  // sampleIndex(j): label given the other densities' selected nodes
            double cmu[IIF_MAX_DIM] = {0, 0, 0, 0}, cvar[IIF_MAX_DIM] = {0, 0, 0, 0};
            bool has[IIF_MAX_DIM] = {false, false, false, false};
            for (int c = 0; c < d; ++c) {
              if (!((masks[j] >> c) & 1)) continue;
              const double lam = cond_gauss(F, sm.mean, sm.var, node, masks, j, nn, d, c, is_circ(cm, c), cmu[c]);
              has[c] = lam > 0;
              cvar[c] = has[c] ? 1.0 / lam : 0.0;
            }
            node[j] = draw_label(j, l, cmu, cvar, has,
                                 gibbs_u((uint32_t)s * ublk + (uint32_t)(F + (l - 1) * F * (niter + 1) + F + it * F + j)));
          }
        }
      }
    }
    IIF_PHASE(11);
    if (live) {
      // samplePoint: draw from the product of the selected leaf kernels
      if (gl == 0) {
        for (int c = 0; c < d; ++c) {
          double mu = 0;
          const double lam = cond_gauss(F, sm.mean, sm.var, node, masks, -1, nn, d, c, is_circ(cm, c), mu);
          double x;
          if (lam > 0) {
            const double e = gibbs_n((uint32_t)s * nblk + (uint32_t)(L * d + c));
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
