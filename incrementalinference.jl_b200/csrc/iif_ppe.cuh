// iif_ppe.cuh — point estimates of a belief: calcPPE, src/services/FGOSUtils.jl:237-278,
// run per frontal at CSM step 5 (CliqueStateMachine.jl:933-939) and in doautoinit! (GraphInit.jl:180).
//   mean = calcMean(P)   : arithmetic mean (Euclid) / extrinsic atan2 mean (circle)
//   max  = getKDEMax(P)  : per coordinate, the marginal KDE evaluated on a 200-point grid over the point
//                          range extended by 10 % on both sides; first grid point of maximal density
//                          (KernelDensityEstimate.jl getKDEMax / getKDERange — un-vendored, parity unpinned).
// One 256-thread CTA per belief: one thread per grid point, Gaussian kernel values by gauss_negU.
#pragma once
#include "iif_conv.cuh"

#define IIF_PPE_GRID 200
#define IIF_PPE_THREADS 256

struct PpeTask {
  int32_t slot, _pad;
  double* out_mean;  // IIF_MAX_DIM
  double* out_max;   // IIF_MAX_DIM
};

__global__ void __launch_bounds__(IIF_PPE_THREADS, 1)
iif_ppe_kernel(DeviceGraph g, const PpeTask* __restrict__ tasks) {
  __shared__ double x[IIF_MAX_POINTS + 8];
  __shared__ double red[IIF_RED_DOUBLES];
  __shared__ double tab[IIF_GTAB_N];
  __shared__ double wv[IIF_PPE_THREADS / 32];
  __shared__ int wi[IIF_PPE_THREADS / 32];
  const PpeTask t = tasks[blockIdx.x];
  const iif_slot_desc S = g.slots[t.slot];
  const int n = g.npts[t.slot], d = S.dim;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int parity = 0;
  gauss_stage_table(tab);
  for (int c = 0; c < d; ++c) {
    const bool circ = is_circ(S.circ_mask, c);
    __syncthreads();
    for (int i = tid; i < n + 8; i += IIF_PPE_THREADS) x[i] = g.pts[S.pts_off + min(i, n - 1) * d + c];  // padded with the last point
    __syncthreads();
    // mean, range
    double v[3] = {0.0, 0.0, 0.0};
    double lo = INFINITY, hi = -INFINITY;
    for (int i = tid; i < n; i += IIF_PPE_THREADS) {
      const double xi = x[i];
      if (circ) { v[1] += sin(xi); v[2] += cos(xi); } else v[0] += xi;
      lo = fmin(lo, xi);
      hi = fmax(hi, xi);
    }
    block_sum<3>(v, red, parity);
    lo = block_min1(lo, red, parity);
    hi = -block_min1(-hi, red, parity);
    const double mean = circ ? atan2(v[1], v[2]) : v[0] / (double)n;
    // marginal density on the grid
    const double dr = 0.1 * (hi - lo);
    const double a = lo - dr, b = hi + dr, step = (b - a) / (double)(IIF_PPE_GRID - 1);
    const double h = g.bw[t.slot * IIF_MAX_DIM + c];
    const double sc = IIF_GSCALE / h;
    const double X = a + step * (double)tid;
    double y = -1.0;
    if (tid < IIF_PPE_GRID) {
      y = 0.0;
      for (int i = 0; i < n; i += 4) {
        double z[4], e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) z[u] = mdiff(X, x[i + u], circ) * sc;
        gauss_negU<4>(z, tab, e);
#pragma unroll
        for (int u = 0; u < 4; ++u) y += (i + u < n) ? e[u] : 0.0;
      }
    }
    // first grid point of maximal density
    int bi = tid;
    for (int o = 16; o > 0; o >>= 1) {
      const double yo = __shfl_xor_sync(0xffffffffu, y, o);
      const int io = __shfl_xor_sync(0xffffffffu, bi, o);
      if (yo > y || (yo == y && io < bi)) { y = yo; bi = io; }
    }
    if (lane == 0) { wv[warp] = y; wi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double by = wv[0];
      int bidx = wi[0];
      for (int w = 1; w < IIF_PPE_THREADS / 32; ++w)
        if (wv[w] > by || (wv[w] == by && wi[w] < bidx)) { by = wv[w]; bidx = wi[w]; }
      const double xm = a + step * (double)bidx;
      t.out_mean[c] = mean;
      t.out_max[c] = circ ? wrap_pi(xm) : xm;
    }
  }
  if (tid == 0)
    for (int c = d; c < IIF_MAX_DIM; ++c) { t.out_mean[c] = 0.0; t.out_max[c] = 0.0; }
}
