// iif_deconv.cuh — SURVEY.md §8f-2: approxDeconv (src/services/DeconvUtils.jl:32-162) and the mmd
// kernel-embedding distance that pins it (SolverUtilities.jl:25-47 -> AMP.mmd!; test/testDefaultDeconv.jl:10-45).
//
// approxDeconv is the inverse of the convolution: for every sample n, solve the factor residual for the
// MEASUREMENT given particle n of every variable (first hypothesis only, DeconvUtils.jl:76,139), starting from a
// sampled measurement; it returns (predicted, sampled).  The built-in residuals are linear in z, so the root is
// closed form (prior z = x1, LinearRelative z = x2 - x1, CircularCircular z = wrap(x2 - x1), EuclidDistance
// z = |x2 - x1|).  One CTA per factor, one thread per sample.
#pragma once
#include "iif_conv.cuh"

struct DeconvTask {
  int32_t factor, N, call_id, _pad;
  double* out_pred;    // N * zdim
  double* out_meas;    // N * zdim
  int32_t* out_status;
};

// approxDeconv of sample n: samples the factor's measurement into `meas` (zd doubles) and writes the measurement
// that zeroes the residual for particle n of every variable into `pred`.  Returns a status code.
__device__ __forceinline__ int deconv_sample(const DeviceGraph& g, const iif_factor_desc& f, uint32_t call_id, int N, int n,
                                             double* pred, double* meas) {
  const int zd = f.zdim;
  const iif_slot_desc S1 = g.slots[f.slot[0]];
  double z[IIF_MAX_DIM] = {0, 0, 0, 0};
  int st = sample_measurement(g, f, call_id, n, z);
  if (st != IIF_OK) return st;
  for (int c = 0; c < zd; ++c) meas[c] = z[c];
  double x[2][IIF_MAX_DIM];
  for (int v = 0; v < f.arity; ++v) {  // _getindex_anyn, NumericalCalculations.jl:377-381
    const iif_slot_desc S = g.slots[f.slot[v]];
    const int len = g.npts[f.slot[v]];
    if (len <= 0) return IIF_ERR_STATE;
    int m = n;
    if (n >= len) {
      const double u = rs_uniform(g.sp->seed, call_id, IIF_RS_ANYN, (uint32_t)(v * N + n));
      m = min((int)(u * len), len - 1);
    }
    for (int c = 0; c < S.dim; ++c) x[v][c] = g.pts[S.pts_off + m * S.dim + c];
  }
  double* p = pred;
  switch (f.kind) {
    case IIF_F_PRIOR:
    case IIF_F_MSG_PRIOR:
      for (int c = 0; c < zd; ++c) p[c] = x[0][c];
      break;
    case IIF_F_PRIOR_CIRCULAR: p[0] = wrap_pi(x[0][0]); break;
    case IIF_F_PARTIAL_PRIOR: {
      int k = 0;
      for (int c = 0; c < S1.dim; ++c)
        if ((f.partial_mask >> c) & 1) p[k++] = x[0][c];
      break;
    }
    case IIF_F_MANIFOLD_PRIOR: {  // the sampled measurement IS a point: z = x1 (on the informed coordinates)
      int k = 0;
      for (int c = 0; c < S1.dim; ++c)
        if (!f.partial_mask || ((f.partial_mask >> c) & 1)) p[k++] = is_circ(S1.circ_mask, c) ? wrap_pi(x[0][c]) : x[0][c];
      break;
    }
    case IIF_F_SO3_PRIOR: so3_between(f.aux, x[0], p); break;       // z with p Exp(z) = x1
    case IIF_F_SO3_RELATIVE: so3_between(x[0], x[1], p); break;     // X = Log(p^T q)
    case IIF_F_SE2_RELATIVE: {    // X = vee(log(eps, p^-1 o q)): X_t = R(theta_p)^T (t_q - t_p), X_theta = theta_q - theta_p
      double sn, cs;
      sincos(x[0][2], &sn, &cs);
      const double dx = x[1][0] - x[0][0], dy = x[1][1] - x[0][1];
      p[0] = cs * dx + sn * dy;
      p[1] = -sn * dx + cs * dy;
      p[2] = wrap_pi(x[1][2] - x[0][2]);
      break;
    }
    case IIF_F_LINEAR_RELATIVE:
    case IIF_F_CIRCULAR_CIRCULAR:
      for (int c = 0; c < zd; ++c) p[c] = mdiff(x[1][c], x[0][c], is_circ(S1.circ_mask, c));
      break;
    case IIF_F_EUCLID_DISTANCE: {
      double s = 0;
      for (int c = 0; c < S1.dim; ++c) s += (x[1][c] - x[0][c]) * (x[1][c] - x[0][c]);
      p[0] = sqrt(s);
      break;
    }
    default: return IIF_ERR_UNSUPPORTED;
  }
  return IIF_OK;
}

__global__ void iif_deconv_kernel(DeviceGraph g, const DeconvTask* __restrict__ tasks) {
  const DeconvTask t = tasks[blockIdx.x];
  const iif_factor_desc f = g.factors[t.factor];
  const int zd = f.zdim;
  __shared__ int s_status;
  if (threadIdx.x == 0) s_status = (f.nmh != 0 || f.arity > 2) ? IIF_ERR_UNSUPPORTED : IIF_OK;
  __syncthreads();
  for (int n = threadIdx.x; n < t.N && s_status == IIF_OK; n += blockDim.x) {
    double pred[IIF_MAX_DIM], meas[IIF_MAX_DIM];
    const int st = deconv_sample(g, f, (uint32_t)t.call_id, t.N, n, pred, meas);
    if (st != IIF_OK) { s_status = st; break; }
    for (int c = 0; c < zd; ++c) {
      t.out_pred[(size_t)n * zd + c] = pred[c];
      t.out_meas[(size_t)n * zd + c] = meas[c];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && t.out_status) *t.out_status = s_status;
}

// IIF_S_DECONV: the differential likelihood of an up message, built on the device (addLikelihoodsDifferentialCHILD!,
// TreeMessageUtils.jl:279-335): approxDeconv of the dummy relative factor over two separator beliefs, the predicted
// measurements mapped to points (exp at the identity) and their manikde! bandwidth, written into a belief slot that a
// relative factor of the parent clique samples its measurements from (IIF_D_KDE).  One CTA (or speculative cluster,
// see block_kde_bandwidth) per differential, one thread per sample.
struct DeconvSlotTask {
  iif_deconv_op op;
  int32_t* out_status;
};

__global__ void __launch_bounds__(IIF_MAX_THREADS, 1)
iif_deconv_slot_kernel(DeviceGraph g, const DeconvSlotTask* __restrict__ tasks, const TreeStruct* __restrict__ trees) {
  extern __shared__ __align__(16) double dcv_smem[];  // conv_smem_bytes(N)
  __shared__ double red[IIF_RED_DOUBLES];
  __shared__ int s_status;
  const int cC = (int)cooperative_groups::this_cluster().num_blocks();
  const bool wr = cooperative_groups::this_cluster().block_rank() == 0;
  if (cC > 1) cooperative_groups::this_cluster().sync();  // every rank is resident before any remote shared-memory access
  const DeconvSlotTask t = tasks[blockIdx.x / cC];
  const iif_factor_desc f = g.factors[t.op.factor];
  const iif_slot_desc O = g.slots[t.op.out_slot];
  const int N = t.op.N, d = O.dim;
  int parity = 0;
  double* pts = dcv_smem;
  double* xa = pts + (size_t)N * IIF_MAX_DIM;
  double* xb = xa + loo_xa_doubles(N);
  double* scr = xb + loo_x2_doubles(N);
  if (threadIdx.x == 0)
    s_status = (f.nmh != 0 || f.arity != 2 || f.zdim != d || O.cap < N) ? IIF_ERR_UNSUPPORTED : IIF_OK;
  __syncthreads();
  const int n = threadIdx.x;
  if (n < N && s_status == IIF_OK) {
    double pred[IIF_MAX_DIM], meas[IIF_MAX_DIM];
    const int st = deconv_sample(g, f, (uint32_t)t.op.call_id, N, n, pred, meas);
    if (st != IIF_OK) s_status = st;
    else
      for (int c = 0; c < d; ++c) pts[n * d + c] = is_circ(O.circ_mask, c) ? wrap_pi(pred[c]) : pred[c];  // exp(M, eps, X)
  }
  __syncthreads();
  const int status = s_status;
  if (t.out_status != nullptr && threadIdx.x == 0 && wr) *t.out_status = status;
  if (status != IIF_OK) return;
  double bw[IIF_MAX_DIM] = {0, 0, 0, 0};
  block_kde_bandwidth<3>(pts, N, d, O.circ_mask, &trees[N], xa, xb, scr, red, &parity, bw);
  if (!wr) return;
  if (n < N)
    for (int c = 0; c < d; ++c) g.pts[O.pts_off + n * d + c] = pts[n * d + c];
  if (n < IIF_MAX_DIM) {
    g.bw[t.op.out_slot * IIF_MAX_DIM + n] = n < d ? bw[n] : 0.0;
    g.ipc[t.op.out_slot * IIF_MAX_DIM + n] = n < d ? 1.0 : 0.0;
  }
  if (n == 0) {
    g.npts[t.op.out_slot] = N;
    g.flags[t.op.out_slot] |= 1;
  }
}

// mmd: sum k(a,a)/Na^2 - 2 sum k(a,b)/(Na Nb) + sum k(b,b)/Nb^2 with k(p,q) = exp(-bw dist(p,q)^2).
// One CTA per pair of point sets; thread i owns row i of each of the three kernel matrices; Gaussian kernel
// values four at a time by exp_negU.
struct MmdTask {
  const double* a;
  const double* b;
  int32_t na, nb, dim, circ_mask;
  double bw;
  double* out;
};

__device__ __forceinline__ double mmd_rows(const double* __restrict__ P, int np, const double* __restrict__ Q, int nq,
                                           int d, int32_t cm, double bw) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    double pi[IIF_MAX_DIM];
    for (int c = 0; c < d; ++c) pi[c] = P[i * d + c];
    for (int j = 0; j < nq; j += 4) {
      double a[4], e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = min(j + u, nq - 1);
        double d2 = 0.0;
        for (int c = 0; c < d; ++c) {
          const double dl = mdiff(pi[c], Q[jj * d + c], is_circ(cm, c));
          d2 = fma(dl, dl, d2);
        }
        a[u] = -bw * d2;
      }
      exp_negU<4>(a, e);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += (j + u < nq) ? e[u] : 0.0;
    }
  }
  return acc;
}

__global__ void __launch_bounds__(256, 1) iif_mmd_kernel(const MmdTask* __restrict__ tasks) {
  extern __shared__ __align__(16) double mmd_smem[];  // a (na*d) then b (nb*d)
  __shared__ double red[IIF_RED_DOUBLES];
  const MmdTask t = tasks[blockIdx.x];
  double* A_ = mmd_smem;
  double* B_ = mmd_smem + (size_t)t.na * t.dim;
  for (int i = threadIdx.x; i < t.na * t.dim; i += blockDim.x) A_[i] = t.a[i];
  for (int i = threadIdx.x; i < t.nb * t.dim; i += blockDim.x) B_[i] = t.b[i];
  __syncthreads();
  int parity = 0;
  double v[3];
  v[0] = mmd_rows(A_, t.na, A_, t.na, t.dim, t.circ_mask, t.bw);
  v[1] = mmd_rows(A_, t.na, B_, t.nb, t.dim, t.circ_mask, t.bw);
  v[2] = mmd_rows(B_, t.nb, B_, t.nb, t.dim, t.circ_mask, t.bw);
  block_sum<3>(v, red, parity);
  if (threadIdx.x == 0)
    *t.out = v[0] / ((double)t.na * t.na) - 2.0 * v[1] / ((double)t.na * t.nb) + v[2] / ((double)t.nb * t.nb);
}
