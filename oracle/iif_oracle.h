/*
 * iif_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar, single-threaded plain-C restatement of the IncrementalInference.jl v0.35.6
 * clique belief-convolution hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libiifb200.so) never links, imports or calls it.
 *
 * PARITY STATUS (see DESIGN.md "Oracle"):
 *   - hypothesis recipe (a6), residual roots (a13), inflation / spread (a7,a8), prior and
 *     relative proposal logic (a4,a5): restated from source under /root/reference, pinned
 *     by the reference's own known-answer tests (test/testExplicitMultihypo.jl,
 *     test/testApproxConv.jl:24-32) and analytic roots.
 *   - KDE bandwidth (a14) and KDE product (a15): the algorithms live in the un-vendored
 *     packages ApproxManifoldProducts v0.9 / KernelDensityEstimate v0.5.x (Project.toml:68,85),
 *     whose source is NOT under /root/reference and Julia is not installed here.  They are
 *     restated from the published algorithms (leave-one-out likelihood cross-validation with
 *     golden-section search; Ihler, Sudderth, Freeman & Willsky, "Efficient multiscale
 *     sampling from products of Gaussian mixtures", NIPS 2003).  Sample-level parity with
 *     Julia for these two is therefore **PARITY UNPINNED**; they are pinned only by the
 *     reference's statistical acceptance bands (test/testBasicGraphs.jl etc.).
 *
 * Descriptor structs are shared with include/iifb200.h (data layout only).
 */
#ifndef IIF_ORACLE_H
#define IIF_ORACLE_H

#include "../include/iifb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* random-stream ids (Philox4x32-10 counter word 1); identical constants in the CUDA code */
enum {
  IIF_RS_MEAS = 1,     /* idx = n*zdim + c : standard normal / uniform for the measurement */
  IIF_RS_MIXLABEL = 2, /* idx = n          : mixture component or KDE kernel pick        */
  IIF_RS_LABEL = 3,    /* idx = n          : hypothesis label                             */
  IIF_RS_INFLATE = 4,  /* idx = (cycle*N + n)*d + c ; cycle == inflateCycles: null entropy */
  IIF_RS_ANYN = 5,     /* idx = v*N + n    : _getindex_anyn random partner                */
  IIF_RS_GIBBS_U = 6,  /* idx = ((s*L + l)*Niter + it)*F + j                               */
  IIF_RS_GIBBS_N = 7,  /* idx = s*d + c                                                    */
  IIF_RS_OLDPAD = 8    /* idx = n*(d+1) + {0: kernel pick, 1+c: jitter}                   */
};

/* host mirror of the device arena */
typedef struct {
  int32_t nslots;
  const iif_slot_desc* slots; /* pts_off must be filled (iifo_layout) */
  double* pts;                /* packed by pts_off */
  double* bw;                 /* nslots * IIF_MAX_DIM */
  double* ipc;                /* nslots * IIF_MAX_DIM */
  int32_t* npts;              /* nslots */
  int32_t* flags;             /* nslots, bit0 = initialized */
  int32_t nfactors;
  const iif_factor_desc* factors;
  int32_t ndists;
  const iif_dist_desc* dists;
  const double* dparams;
  iif_solver_params sp;
} iifo_graph;

/* fills pts_off, returns total doubles */
int64_t iifo_layout(int32_t nslots, iif_slot_desc* slots);

/* Philox-based streams */
void iifo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                     uint32_t k1, uint32_t out[4]);
double iifo_uniform(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx);
double iifo_normal(uint64_t seed, uint32_t call, uint32_t stream, uint32_t idx);

/* a6: hypothesis recipe.  mh == NULL => the `Nothing` method.  sfidx 1-based.
 * u (maxlen uniforms) drives the categorical draw; if mhidx_in != NULL labels are taken from it.
 * Outputs: mhidx[maxlen]; nbuckets; bucket_hypo[b]; bucket_nvars[b]; bucket_vars[b*IIF_MAX_ARITY+..]
 * (1-based variable indices, sorted); certain[ncertain] (1-based). */
int32_t iifo_hypo_recipe(const double* mh, int32_t lenXi, int32_t maxlen, int32_t sfidx,
                         const int32_t* isinit, double nullhypo, const double* u,
                         const int32_t* mhidx_in, int32_t* mhidx, int32_t* nbuckets,
                         int32_t* bucket_hypo, int32_t* bucket_nvars, int32_t* bucket_vars,
                         int32_t* certain, int32_t* ncertain);

/* a13: residual of a built-in factor; x = arity points (each d doubles, packed) */
int32_t iifo_residual(int32_t kind, int32_t d, int32_t circ_mask, int32_t zdim, const double* z,
                      int32_t arity, const double* x, double* res);

/* a7: calcStdBasicSpread */
double iifo_std_basic_spread(const double* pts, int32_t n, int32_t d, int32_t circ_mask);

/* a14: per-dimension leave-one-out likelihood bandwidth (manikde! with bw === nothing) */
/* approxDeconv (DeconvUtils.jl:32-162): predicted and sampled measurements, N x zdim each */
int32_t iifo_deconv(const iifo_graph* g, int32_t factor, int32_t N, int32_t call_id,
                    double* out_pred, double* out_meas);
/* mmd (SolverUtilities.jl:25-47 / AMP.mmd!): kernel-embedding distance of two point sets */
double iifo_mmd(const double* a, int32_t na, const double* b, int32_t nb, int32_t d, int32_t cm, double bw);
/* calcPPE (FGOSUtils.jl:237-278): mean and KDE-max point estimates of one belief */
int32_t iifo_ppe(const double* pts, int32_t n, int32_t d, int32_t cm, const double* bw,
                 double* mean_out, double* max_out);
int32_t iifo_kde_bandwidth(const double* pts, int32_t n, int32_t d, int32_t circ_mask,
                           double* bw_out);
/* LOO objective (negative average log likelihood) of one coordinate at bandwidth h */
double iifo_loo_nll(const double* x, int32_t n, int32_t circular, double h);

/* a3: one approxConvBelief.  out_pts N*d, out_bw/out_ipc IIF_MAX_DIM, out_mhidx N (or NULL).
 * meas/mhidx_in/uinf: explicit streams or NULL.  Returns status; *nan_count particles left
 * unchanged because the solve produced NaN. */
int32_t iifo_conv(const iifo_graph* g, const iif_conv_op* op, const double* meas,
                  const int32_t* mhidx_in, const double* uinf, double* out_pts, double* out_bw,
                  double* out_ipc, int32_t* out_mhidx, int32_t* nan_count);

/* a15: manifoldProduct of F densities (points packed F x N x d, bw F x IIF_MAX_DIM). */
int32_t iifo_product(int32_t d, int32_t circ_mask, int32_t F, int32_t N, const double* dens_pts,
                     const double* dens_bw, const int32_t* dens_mask, const double* old_pts,
                     uint64_t seed, uint32_t call_id, int32_t niter, const double* randU,
                     const double* randN, double* out_pts, double* out_bw, int32_t* out_labels);

/* a1: propagateBelief + setBelief! on the host mirror */
int32_t iifo_propagate(iifo_graph* g, const iif_prop_op* op);

/* a16: run a whole schedule (waves are executed in order; ops inside a wave in order) */
int32_t iifo_schedule_run(iifo_graph* g, int32_t nwaves, const int32_t* wave_off,
                          const iif_sched_op* ops, const iif_prop_op* props, int32_t first_wave,
                          int32_t last_wave);
int32_t iifo_schedule_run_ex(iifo_graph* g, int32_t nwaves, const int32_t* wave_off,
                             const iif_sched_op* ops, const iif_prop_op* props, const iif_deconv_op* deconvs,
                             int32_t first_wave, int32_t last_wave);
/* IIF_S_DECONV: differential likelihood of an up message (TreeMessageUtils.jl:314-321) into a belief slot */
int32_t iifo_deconv_to_slot(iifo_graph* g, const iif_deconv_op* op);

/* counters for the CPU baseline */
int64_t iifo_conv_count(void);
void iifo_reset_counters(void);

#ifdef __cplusplus
}
#endif
#endif
