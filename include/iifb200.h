/*
 * iifb200.h — C ABI of libiifb200.so: B200-native (sm_100a) nonparametric clique
 * belief-convolution hot path of IncrementalInference.jl (IIF).
 *
 * This is the drop-in boundary a Julia shim binds with `ccall`.
 * All entry points are `extern "C"`, take plain pointers and sizes, return an int32
 * status (0 = ok, <0 = error, message via iifb200_last_error) and never throw.
 * No torch / Julia types cross this boundary.
 *
 * Reference interfaces replaced (file:line relative to the IIF v0.35.6 tree):
 *   iifb200_conv_batch       <- approxConvBelief            src/services/ApproxConv.jl:4-45
 *                               evalFactor                  src/services/EvalFactor.jl:571-603
 *                               evalPotentialSpecific       src/services/EvalFactor.jl:321-395, 400-542
 *                               computeAcrossHypothesis!    src/services/EvalFactor.jl:145-237
 *                               _prepareHypoRecipe!         src/services/ExplicitDiscreteMarginalizations.jl:142-289
 *                               approxConvOnElements!       src/services/EvalFactor.jl:14-27
 *                               _solveCCWNumeric!           src/services/NumericalCalculations.jl:413-452
 *                               manikde! (AMP, call site)   src/services/ApproxConv.jl:36-42
 *   iifb200_product_batch    <- AMP.manifoldProduct (call)  src/services/GraphProductOperations.jl:53-60
 *   iifb200_propagate_batch  <- propagateBelief             src/services/GraphProductOperations.jl:16-64
 *                               proposalbeliefs!            src/services/ApproxConv.jl:238-304
 *   iifb200_schedule_*       <- upGibbsCliqueDensity        src/services/SolveTree.jl:164-239
 *                               fmcmc! / doFMCIteration     src/services/SolveTree.jl:47-142
 *                               localProductAndUpdate!      src/services/GraphProductOperations.jl:136-155
 *                               solveCliqDownFrontalProducts! src/CliqueStateMachine/services/CliqStateMachineUtils.jl:479-571
 *   iifb200_kde_bandwidth    <- AMP.manikde! bandwidth (call sites ApproxConv.jl:38-41, FGOSUtils.jl:118-128)
 *   iifb200_ppe_batch        <- calcPPE / setPPE!           src/services/FGOSUtils.jl:237-278
 *   iifb200_deconv_batch     <- approxDeconv                src/services/DeconvUtils.jl:32-162
 *   IIF_S_DECONV ops         <- addLikelihoodsDifferentialCHILD!  src/services/TreeMessageUtils.jl:279-335
 *   iifb200_mmd              <- mmd (AMP.mmd!)              src/services/SolverUtilities.jl:25-47
 *   belief slots             <- VariableNodeData.val/.bw, TreeBelief   src/entities/BeliefTypes.jl:47-57
 *   iif_factor_desc          <- CommonConvWrapper           src/entities/FactorOperationalMemory.jl:21-70
 *   iif_solver_params        <- SolverParams                src/entities/SolverParams.jl:12-75
 *
 * Data layout. A "belief slot" is the device-resident analogue of one variable's
 * (val, bw, infoPerCoord) inside one (sub)graph: `cap` points of `dim` doubles each,
 * point-major (`pts[n*dim + c]`, identical to Julia's Vector{SVector{dim,Float64}}),
 * plus `dim` bandwidths and `dim` infoPerCoord values.  Points are stored as
 * coordinates at the group identity (TranslationGroup(d), RealCircleGroup and their
 * products: point representation == coordinates; circular coordinates live in [-pi,pi)).
 * SpecialEuclidean(2) points ArrayPartition(t, R) are stored as (t1, t2, atan2(R21, R11)) with
 * dim = 3, circ_mask = 0b100 (vee(log(M, eps, p)) in the hybrid tangent representation the reference
 * tests use, test/testSpecialEuclidean2Mani.jl:14); KDE bandwidth / product / statistics work on these
 * coordinates, only the relative residual (IIF_F_SE2_RELATIVE) composes on the group.
 */
#ifndef IIFB200_H
#define IIFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IIF_MAX_DIM 4           /* max variable dimension handled by the kernels */
#define IIF_MAX_ARITY 6         /* max variables per factor (four-door multihypo uses 5) */
#define IIF_MAX_FACTORS 16      /* max factors contributing to one propagateBelief (more: IIF_ERR_ARG, never truncated) */
#define IIF_MAX_POINTS 256      /* max particles per belief (kernel shared-memory budget) */

/* status codes */
#define IIF_OK 0
#define IIF_ERR_ARG (-1)         /* bad argument / descriptor */
#define IIF_ERR_CUDA (-2)        /* CUDA runtime error (message has details) */
#define IIF_ERR_UNSUPPORTED (-3) /* factor / distribution kind with no device implementation */
#define IIF_ERR_STATE (-4)       /* call order (e.g. run before graph upload) */

/* factor kinds: device residual library (src/Factors/ *.jl) */
enum iif_factor_kind {
  IIF_F_PRIOR = 1,             /* Prior            DefaultPrior.jl:8,17      z - x1              */
  IIF_F_LINEAR_RELATIVE = 2,   /* LinearRelative   LinearRelative.jl:13,42   z - (x2 - x1)       */
  IIF_F_PRIOR_CIRCULAR = 3,    /* PriorCircular    Circular.jl:54,70                             */
  IIF_F_CIRCULAR_CIRCULAR = 4, /* CircularCircular Circular.jl:13,24  vee(log(q, exp(p,X)))      */
  IIF_F_EUCLID_DISTANCE = 5,   /* EuclidDistance   EuclidDistance.jl:9,20    z - norm(x2 - x1)   */
  IIF_F_MSG_PRIOR = 6,         /* MsgPrior         MsgPrior.jl:10,36         z - x1              */
  IIF_F_PARTIAL_PRIOR = 7,     /* PartialPrior     PartialPrior.jl:11-21  (prior on partial_mask dims) */
  /* generic-manifold factors (src/Factors/GenericFunctions.jl) on groups whose point coordinates
   * are (Euclid..., angle...): TranslationGroup(d), RealCircleGroup, SpecialEuclidean(2) as (x, y, theta). */
  IIF_F_MANIFOLD_PRIOR = 8,    /* ManifoldPrior :163-214 / ManifoldPriorPartial :288-303: sample = retract(M, p, hat(Z))
                                  = p (+) Z coordinate-wise (angles wrapped); p is folded into Z's mean by the host;
                                  partial_mask != 0 selects the informed coordinates */
  IIF_F_SE2_RELATIVE = 9,      /* ManifoldFactor{SpecialEuclidean(2)} :64-100 with distanceTangent2Point :39-44:
                                  qhat = p o exp(eps, X), residual vee(log(q, qhat)); X = (dx, dy, dtheta) in p's frame */
  /* SpecialOrthogonal(3) (test/testSpecialOrthogonalMani.jl:75-140).  Points are stored as rotation vectors
   * omega = vee(log(eps, R)) (slot flag IIF_MANI_SO3); AMP treats these coordinates as (:Euclid, :Euclid, :Euclid)
   * (same test, :80-81), so bandwidth and product work on them unchanged; the factors compose on the group. */
  IIF_F_SO3_PRIOR = 10,        /* ManifoldPrior{SpecialOrthogonal(3)} :163-214: sample = retract(M, p, hat(Z)) = p Exp(z);
                                  p's rotation vector in iif_factor_desc.aux */
  IIF_F_SO3_RELATIVE = 11      /* ManifoldFactor{SpecialOrthogonal(3)}: qhat = p Exp(X), residual vee(log(q, qhat)) = Log(q^T qhat) */
};
/* iif_slot_desc.circ_mask, bit 8: the three coordinates are an SO(3) rotation vector — entropy, spread statistics and the
 * group factors compose on the manifold; KDE bandwidth / product treat them as Euclid coordinates (AMP's convention). */
#define IIF_MANI_SO3 0x100

/* measurement distributions (SamplableBelief) */
enum iif_dist_kind {
  IIF_D_NORMAL = 1,   /* params: [mu, sigma]                                           */
  IIF_D_MVNORMAL = 2, /* params: [mu(dim), L(dim*dim) row-major lower Cholesky factor]   */
  IIF_D_MIXTURE = 3,  /* params: [w(ncomp)] then ncomp component blocks of comp_kind     */
  IIF_D_KDE = 4,      /* ManifoldKernelDensity held in belief slot `slot` (MsgPrior)     */
  IIF_D_UNIFORM = 5,  /* params: [a, b]                                                  */
  IIF_D_SAMPLES = 6   /* ANY SamplableBelief the host can draw from (sampleFactor! is `rand(Z)` per sample,
                         SolverUtilities.jl:50-76; Rayleigh, Gamma, user types ...): params hold ncomp samples of dim values,
                         row-major, drawn by the host; the device resamples the table uniformly with replacement */
};

typedef struct {
  int32_t kind;      /* iif_dist_kind */
  int32_t dim;       /* sample dimension z */
  int32_t ncomp;     /* MIXTURE: number of components */
  int32_t comp_kind; /* MIXTURE: kind of every component (NORMAL or MVNORMAL) */
  int32_t slot;      /* KDE: belief slot holding the kernel centres and bandwidths */
  int32_t poff;      /* offset (doubles) of this distribution's block in the params array */
} iif_dist_desc;

typedef struct {
  int32_t dim;       /* variable dimension d (== doubles per point) */
  int32_t circ_mask; /* bit c set: coordinate c is circular (RealCircleGroup), else Euclid */
  int32_t cap;       /* capacity in points (>= N) */
  int32_t pts_off;   /* OUT (filled by iifb200_set_graph): offset in doubles into the arena */
} iif_slot_desc;

typedef struct {
  int32_t kind;                /* iif_factor_kind */
  int32_t arity;               /* number of variables */
  int32_t zdim;                /* measurement dimension */
  int32_t dist;                /* index into the distribution table */
  int32_t slot[IIF_MAX_ARITY]; /* belief slot of each variable, in factor variable order */
  int32_t nmh;                 /* 0: no multihypo; else == arity */
  int32_t partial_mask;        /* 0: full; else bit c set: factor informs coordinate c */
  int32_t solver;              /* per-sample solve of a relative factor (_solveLambdaNumeric, NumericalCalculations.jl:49-133):
                                  0 = closed-form root where the residual has a unique one (what the optimiser converges
                                  to), Nelder-Mead for EuclidDistance in more than one dimension (a ring of roots: the
                                  result IS the optimiser's path); 1 = always Nelder-Mead over the solve-for point's
                                  coordinates, restating Optim.NelderMead as the reference configures it (adaptive
                                  parameters, affine initial simplex, g_abstol 1e-8, 1000 iterations) */
  int32_t _pad;
  double mh[IIF_MAX_ARITY];    /* parseusermultihypo output (FactorGraph.jl:639-654): 0.0 = certain */
  double nullhypo;             /* CCW.nullhypo   FactorOperationalMemory.jl:51 */
  double inflation;            /* CCW.inflation  FactorOperationalMemory.jl:53 (default 5.0) */
  double aux[IIF_MAX_DIM];     /* IIF_F_SO3_PRIOR: coordinates of the prior's point p (ManifoldPrior.p); else unused */
} iif_factor_desc;

/* SolverParams subset used on the hot path (src/entities/SolverParams.jl) */
typedef struct {
  double spreadNH;       /* :57 default 3.0 */
  double nullSurplusAdd; /* :61 default 0.3 */
  int32_t inflateCycles; /* :63 default 3   */
  int32_t gibbsNiter;    /* AMP.manifoldProduct Niter, GraphProductOperations.jl:56 == 1 */
  uint64_t seed;         /* Philox key for every device-drawn random stream */
} iif_solver_params;

/* One Chapman-Kolmogorov convolution = one approxConvBelief (the metric's unit of work). */
typedef struct {
  int32_t factor;     /* index into factor table */
  int32_t sfidx;      /* 1-based index of the solve-for variable in the factor's variable order */
  int32_t N;          /* number of proposal points */
  int32_t call_id;    /* unique id: selects this op's Philox streams */
  double nullSurplus; /* ApproxConv.jl:256-265 */
  /* optional explicit random streams (offset in elements into the arrays passed with the
   * batch call; -1 => drawn on device from Philox).  These reproduce a host RNG exactly:
   *   meas:  N*zdim doubles      (sampleFactor!, SolverUtilities.jl:50)
   *   mhidx: N int32 labels      (_prepareHypoRecipe! rand(Categorical), EDM.jl:186,261)
   *   uinf:  (inflateCycles+1)*N*d uniforms in [0,1): cycle-major, particle, coord
   *          (addEntropyOnManifold! rand(), EvalFactor.jl:115-118); the extra block feeds
   *          the null / other-hypothesis entropy. */
  int32_t meas_off, mhidx_off, uinf_off, _pad;
} iif_conv_op;

/* One propagateBelief: F convolutions + manifoldProduct, posterior written to out_slot. */
typedef struct {
  int32_t target_slot;             /* destination variable (read: current particles)     */
  int32_t out_slot;                /* where posterior (pts,bw,ipc) is written (may == target) */
  int32_t nfactors;                /* F */
  int32_t N;
  int32_t factor[IIF_MAX_FACTORS]; /* factor table indices */
  int32_t sfidx[IIF_MAX_FACTORS];  /* 1-based solve-for index inside each factor */
  int32_t call_id;                 /* base id; convolution f uses call_id+1+f, product uses call_id */
  int32_t any_multihypo;           /* proposalbeliefs! sibling nullSurplus rule, ApproxConv.jl:256-265 */
} iif_prop_op;

/* schedule op kinds (one clique solve = a few of these; one wave = independent ops) */
enum iif_sched_kind {
  IIF_S_PROPAGATE = 1, /* propagateBelief + setBelief!  (SolveTree.jl:63-74)                */
  IIF_S_COPY = 2,      /* slot := slot (separator message adoption, TreeMessageUtils.jl:66) */
  IIF_S_DECONV = 3,    /* differential likelihood of an up message (useMsgLikelihoods=true):
                          slot := manikde!(exp(M, eps, approxDeconv(dummy factor)))
                          addLikelihoodsDifferentialCHILD!, TreeMessageUtils.jl:279-335            */
  /* multi-GPU (one process per GPU, peer arenas attached with iifb200_ipc_attach): a separator message that crosses
   * a GPU boundary (prepCliqueMsgUp TreeMessageUtils.jl:667-703 going up, CliqDownMessage CliqueStateMachine.jl:672-691
   * coming down) is a PUSH node on the sender and a WAIT node on the receiver INSIDE the captured CUDA graphs: no host
   * call, no graph split.  a = index into the iif_xfer_op table. */
  IIF_S_PUSH = 4,      /* copy belief slot xfer.slot into the same slot of peer xfer.peer over NVLink, then raise the
                          peer's flag xfer.msg (runs after the segment's kernels)                   */
  IIF_S_WAIT = 5       /* spin until the local flag xfer.msg has been raised in this replay (runs before the segment's
                          kernels)                                                                 */
};
typedef struct {
  int32_t slot; /* belief slot (same index and layout on every rank: all ranks build the same slot table) */
  int32_t peer; /* PUSH: destination rank;  WAIT: source rank (informational) */
  int32_t msg;  /* message id == flag index, unique per (slot, destination) in a pass, < nflags */
  int32_t _pad;
} iif_xfer_op;

/* One differential-likelihood construction (TreeMessageUtils.jl:314-321): approxDeconv of the (dummy)
 * relative factor `factor` over two separator beliefs, the predicted measurements mapped to points
 * (exp at the identity: angles wrapped) and written, with their KDE bandwidth, into `out_slot`
 * (dim == the factor's zdim).  The belief in out_slot then serves as the measurement density
 * (IIF_D_KDE) of the relative factor the parent clique adds (`_sft(newBel)`, :321-324). */
typedef struct {
  int32_t factor;   /* index into the factor table: binary relative factor, no multihypo */
  int32_t out_slot; /* belief slot receiving N points of dim zdim (+ bw, ipc = 1, initialized) */
  int32_t N;        /* samples (approxDeconv uses the point count of the first variable, DeconvUtils.jl:194-196) */
  int32_t call_id;  /* Philox call id (partner draws for short beliefs, measurement start samples) */
} iif_deconv_op;

typedef struct iifb200_ctx iifb200_ctx;

/* ---- lifecycle -------------------------------------------------------------------- */
int32_t iifb200_init(int32_t device_ordinal, iifb200_ctx** ctx_out);
void iifb200_free(iifb200_ctx* ctx);
const char* iifb200_last_error(const iifb200_ctx* ctx); /* ctx may be NULL: last init error */
int32_t iifb200_version(void);

/* ---- graph upload (descriptor tables; copied, caller keeps ownership) ------------- */
/* Allocates the device arena.  If ext_arena != NULL the caller supplies device memory of at
 * least iifb200_arena_bytes(...) bytes (e.g. a torch CUDA tensor used as NCCL buffer). */
int64_t iifb200_arena_bytes(int32_t nslots, const iif_slot_desc* slots);
int32_t iifb200_set_graph(iifb200_ctx* ctx, int32_t nslots, iif_slot_desc* slots /* in/out */,
                          int32_t nfactors, const iif_factor_desc* factors, int32_t ndists,
                          const iif_dist_desc* dists, int32_t nparams, const double* dparams,
                          const iif_solver_params* sp, void* ext_arena /* device ptr or NULL */);
int32_t iifb200_set_solver_params(iifb200_ctx* ctx, const iif_solver_params* sp);

/* ---- belief I/O (host <-> device) ------------------------------------------------- */
int32_t iifb200_upload_belief(iifb200_ctx* ctx, int32_t slot, int32_t npts, const double* pts,
                              const double* bw /* dim or NULL */, int32_t initialized);
int32_t iifb200_download_belief(iifb200_ctx* ctx, int32_t slot, int32_t* npts, double* pts,
                                double* bw, double* ipc);
/* all slots at once: pts packed by pts_off, bw/ipc as nslots*IIF_MAX_DIM, npts/flags nslots */
int32_t iifb200_upload_all(iifb200_ctx* ctx, const double* pts, const double* bw,
                           const int32_t* npts, const int32_t* flags);
int32_t iifb200_download_all(iifb200_ctx* ctx, double* pts, double* bw, double* ipc,
                             int32_t* npts);
/* contiguous slot range [first, first+count): pts packed back to back starting at slot `first`'s
 * offset, bw count*IIF_MAX_DIM, npts/flags count.  Asynchronous on the ctx stream when the host
 * buffers are pinned (iifb200_host_alloc); call iifb200_sync before reading downloaded data. */
int32_t iifb200_upload_slots(iifb200_ctx* ctx, int32_t first, int32_t count, const double* pts,
                             const double* bw, const int32_t* npts, const int32_t* flags);
int32_t iifb200_download_slots(iifb200_ctx* ctx, int32_t first, int32_t count, double* pts, double* bw,
                               double* ipc, int32_t* npts);
/* pinned host memory for the transfers above */
int32_t iifb200_host_alloc(iifb200_ctx* ctx, int64_t bytes, void** ptr_out);
int32_t iifb200_host_free(iifb200_ctx* ctx, void* ptr);
/* raw device pointer of a slot's points (for NCCL send/recv of separator messages);
 * layout: pts[cap*dim]; bw/ipc live in the tail region, see iifb200_slot_msg_ptr */
int32_t iifb200_slot_device_ptr(iifb200_ctx* ctx, int32_t slot, void** pts_ptr, void** bw_ptr);

/* ---- hot path ---------------------------------------------------------------------- */
/* K independent convolutions (approxConvBelief).  Outputs are HOST buffers, caller-owned:
 *   out_pts   K x N x d (packed back to back, op k at sum_{j<k} N_j*d_j)
 *   out_bw    K x IIF_MAX_DIM,  out_ipc K x IIF_MAX_DIM
 *   out_mhidx packed like out_pts with N_j ints each (may be NULL)
 *   out_nan   K ints: particles whose solve produced NaN and were left unchanged
 *             (NumericalCalculations.jl:348-351)  (may be NULL)
 * Source/target particles are the current contents of the belief slots; the target slot is
 * NOT modified (ApproxConv.jl:17, test/testMultiHypo3Door.jl:59-90). */
int32_t iifb200_conv_batch(iifb200_ctx* ctx, int32_t K, const iif_conv_op* ops,
                           const double* meas, const int32_t* mhidx, const double* uinf,
                           double* out_pts, double* out_bw, double* out_ipc,
                           int32_t* out_mhidx, int32_t* out_nan);

/* V independent KDE products (AMP.manifoldProduct).  Inputs HOST buffers:
 *   dens_pts  packed F_v x N x d proposals per product, dens_bw F_v x IIF_MAX_DIM,
 *   dens_mask F_v partial masks (0 = full), old_pts N x d,
 *   randU / randN optional explicit Gibbs streams (NULL => Philox from call_id), the analogue of AMP.manifoldProduct's
 *   `_randU` / `_randN` keyword vectors.  With L = floor(log2(N) + 1) levels and Niter Gibbs sweeps per level, output
 *   sample s owns randU[s*F*(1 + L*(Niter+1)) ...]: F uniforms for initIndices, then per level F for sampleIndices
 *   (labels given the point) and Niter*F for sampleIndex (labels given the other densities' labels); and
 *   randN[s*d*(L+1) ...]: d normals for the samplePoint of every level and d for the final sample. */
typedef struct {
  int32_t dim, circ_mask, nfactors, N;
  int32_t call_id;
  int32_t randu_off, randn_off; /* -1 => Philox */
  int32_t _pad;
} iif_product_op;
int32_t iifb200_product_batch(iifb200_ctx* ctx, int32_t V, const iif_product_op* ops,
                              const double* dens_pts, const double* dens_bw,
                              const int32_t* dens_mask, const double* old_pts,
                              const double* randU, const double* randN, double* out_pts,
                              double* out_bw, int32_t* out_labels /* V x N x F or NULL */);

/* KDE leave-one-out likelihood bandwidth of K point sets (manikde! with bw === nothing). */
int32_t iifb200_kde_bandwidth(iifb200_ctx* ctx, int32_t K, const int32_t* N, const int32_t* dim,
                              const int32_t* circ_mask, const double* pts /* packed */,
                              double* out_bw /* K x IIF_MAX_DIM */);

/* Point estimates of V device-resident beliefs (calcPPE, src/services/FGOSUtils.jl:237-278; setPPE! at CSM
 * step 5, CliqueStateMachine.jl:933-939): out_mean = calcMean, out_max = getKDEMax (per coordinate, first
 * maximum of the marginal KDE on a 200-point grid over the 10 %-extended point range).  Outputs are HOST
 * buffers of V x IIF_MAX_DIM doubles (coordinates; the "suggested" estimate is the mean). */
int32_t iifb200_ppe_batch(iifb200_ctx* ctx, int32_t V, const int32_t* slots, double* out_mean,
                          double* out_max);

/* approxDeconv of K factors on device-resident beliefs (src/services/DeconvUtils.jl:32-162): for sample n the
 * factor residual is solved for the measurement given particle n of every variable (closed form for the
 * built-in residuals; multihypo factors are unsupported, as in the reference :22).  Outputs are HOST buffers
 * packed per factor, N_k x zdim_k doubles each: out_pred = predicted, out_meas = sampled measurements. */
int32_t iifb200_deconv_batch(iifb200_ctx* ctx, int32_t K, const int32_t* factors, const int32_t* N,
                             const int32_t* call_ids, double* out_pred, double* out_meas);

/* mmd kernel-embedding distance of K pairs of HOST point sets (src/services/SolverUtilities.jl:25-47 ->
 * AMP.mmd!): sum k(a,a)/Na^2 - 2 sum k(a,b)/(Na Nb) + sum k(b,b)/Nb^2, k(p,q) = exp(-bw dist(p,q)^2)
 * (bw = 0.001 in the reference).  a / b are packed Na_k x dim_k / Nb_k x dim_k; out holds K doubles. */
int32_t iifb200_mmd(iifb200_ctx* ctx, int32_t K, const int32_t* na, const int32_t* nb, const int32_t* dim,
                    const int32_t* circ_mask, const double* a, const double* b, double bw, double* out);

/* V independent propagateBelief calls on device-resident slots (one launch sequence).
 * Posteriors are written into out_slot on the device; nothing is copied to the host. */
int32_t iifb200_propagate_batch(iifb200_ctx* ctx, int32_t V, const iif_prop_op* ops);
/* ONE propagateBelief per call, host buffers in and out (boundary B3, GraphProductOperations.jl:16-64): what
 * iifb200_set_graph + iifb200_upload_slots(0, nslots) + iifb200_propagate_batch(1) + iifb200_download_belief(op->out_slot)
 * do, with one stream synchronisation instead of five.  pts / bw / npts / flags cover all nslots slots, packed as for
 * iifb200_upload_slots; out_pts receives op->N points of the destination's dimension. */
int32_t iifb200_propagate_once(iifb200_ctx* ctx, int32_t nslots, iif_slot_desc* slots, int32_t nfactors,
                               const iif_factor_desc* factors, int32_t ndists, const iif_dist_desc* dists, int32_t nparams,
                               const double* dparams, const iif_solver_params* sp, const double* pts, const double* bw,
                               const int32_t* npts, const int32_t* flags, const iif_prop_op* op, int32_t* out_npts,
                               double* out_pts, double* out_bw, double* out_ipc);

/* ---- clique / tree schedule (throughput mode, boundary B4) ------------------------ */
/* A schedule is a sequence of waves; wave w holds ops [wave_off[w], wave_off[w+1]) that are
 * mutually independent.  Ops are PROPAGATE (index into props) or COPY (src,dst slot).
 * The schedule is captured once into a CUDA graph and replayed by iifb200_schedule_run. */
typedef struct {
  int32_t kind; /* iif_sched_kind */
  int32_t a;    /* PROPAGATE: index into props;  COPY: source slot;  DECONV: index into deconvs */
  int32_t b;    /* COPY: destination slot */
  int32_t lane; /* 0: no lane (a wave holding such an op is a barrier).  1..8: independent sub-tree ("lane") the op
                   belongs to.  Ops of different lanes never touch a common slot unless a barrier wave lies between
                   them (the caller guarantees it); inside the captured CUDA graph every lane is its own branch, so a
                   lane's next wave starts as soon as ITS previous wave has drained. */
} iif_sched_op;
int32_t iifb200_schedule_build(iifb200_ctx* ctx, int32_t nwaves, const int32_t* wave_off,
                               int32_t nops, const iif_sched_op* ops, int32_t nprops,
                               const iif_prop_op* props, int32_t* schedule_id_out);
/* same, with IIF_S_DECONV ops (tree solves with SolverParams.useMsgLikelihoods = true) */
int32_t iifb200_schedule_build_ex(iifb200_ctx* ctx, int32_t nwaves, const int32_t* wave_off,
                                  int32_t nops, const iif_sched_op* ops, int32_t nprops,
                                  const iif_prop_op* props, int32_t ndeconvs,
                                  const iif_deconv_op* deconvs, int32_t* schedule_id_out);
/* same, with IIF_S_PUSH / IIF_S_WAIT ops for schedules whose passes exchange beliefs with peer GPUs (call
 * iifb200_ipc_attach first).  Such a schedule must be run whole (first_wave = 0, last_wave = -1): every replay advances
 * the epoch the flags are compared with, on every rank alike. */
int32_t iifb200_schedule_build_dist(iifb200_ctx* ctx, int32_t nwaves, const int32_t* wave_off, int32_t nops,
                                    const iif_sched_op* ops, int32_t nprops, const iif_prop_op* props,
                                    int32_t ndeconvs, const iif_deconv_op* deconvs, int32_t nxfers,
                                    const iif_xfer_op* xfers, int32_t* schedule_id_out);
/* Peer memory.  iifb200_ipc_export (after iifb200_set_graph, library-owned arena) allocates `nflags` message flags
 * and returns the 64-byte CUDA IPC handles of the arena and of the flags; the caller all-gathers them (e.g. with
 * torch.distributed) and hands every rank's handles to iifb200_ipc_attach, which maps the peers' arenas and flags. */
int32_t iifb200_ipc_export(iifb200_ctx* ctx, int32_t nflags, void* arena_handle64, void* flags_handle64);
int32_t iifb200_ipc_attach(iifb200_ctx* ctx, int32_t world, int32_t rank, const void* arena_handles,
                           const void* flags_handles);
/* Runs schedule asynchronously on the ctx stream; `first_wave,last_wave` select a wave range
 * (multi-GPU: run to a cut level, exchange separator messages with NCCL, continue). */
int32_t iifb200_schedule_run(iifb200_ctx* ctx, int32_t schedule_id, int32_t first_wave,
                             int32_t last_wave);
int32_t iifb200_schedule_free(iifb200_ctx* ctx, int32_t schedule_id);
int32_t iifb200_sync(iifb200_ctx* ctx);

/* ---- tree planner (boundary B4 for a reference-side caller) ------------------------------------------------------
 * The reference's own offload unit is a CLIQUE: `upGibbsCliqueDensity(dfg, cliq, ...)` (src/services/SolveTree.jl:164-173,
 * shipped to workers by CliqStateMachineUtils.jl:369-385) and `solveCliqDownFrontalProducts!` (:479-571).  The planner
 * takes what `tree.bt` and the DFG already hold — per clique: parent, frontals, separators, potentials and the four Gibbs
 * variable classes of setCliqMCIDs! (JunctionTreeUtils.jl:1352-1523) — and produces the whole pass inside the library:
 * clique-local belief slots, one iif_prop_op per propagateBelief of every fmcmc! sweep, MsgPrior instances for the up
 * messages, separator adoption and addDownVariableFactors! for the down pass, the frontal write-back, then waves of
 * independent ops, copy forwarding and lanes.  Pure host code (no GPU needed); iifb200_plan_upload puts it on a device.
 * Variables are numbered 0..nvars-1 in graph order and factor descriptors refer to them through `slot[]`; in the plan the
 * main-graph belief of variable v is slot v. */
typedef struct {
  int32_t nvars;
  const iif_slot_desc* vars;       /* dim, circ_mask, cap (>= current point count) of every graph variable */
  int32_t nfactors;
  const iif_factor_desc* factors;  /* graph order (getFactors); slot[k] = variable index */
  int32_t ndists;
  const iif_dist_desc* dists;
  int32_t nparams;
  const double* dparams;
  /* only read when iif_plan_opts.useMsgLikelihoods != 0 (differential separator messages, TreeMessageUtils.jl:279-446);
   * may be NULL otherwise.  Type ids are the caller's own numbering of factor / variable TYPES (e.g. an index into a
   * table of type names): the joint-message rules compare them only for equality. */
  const int32_t* factor_type;           /* per factor: id of typeof(getFactorType(fct)) */
  const int32_t* var_type;              /* per variable: id of its variable type */
  const int32_t* var_relative_kind;     /* per variable type T: iif_factor_kind of selectFactorType(T, T) (DefaultNodeTypes.jl:
                                           12-31: Position{N} -> LinearRelative{N}, Circular -> CircularCircular) or 0 */
  const int32_t* var_relative_type;     /* ... and that factor type's id in the numbering of `factor_type` */
  int32_t msgprior_type;                /* type id of MsgPrior in the numbering of `factor_type` */
  int32_t _pad;
} iif_graph_desc;

/* Bayes tree in CSR form; cliques are numbered parents first (parent id < child id, roots have parent -1), children of a
 * clique are taken in increasing id.  All lists hold variable indices except `potentials` (factor indices). */
typedef struct {
  int32_t ncliques;
  const int32_t* parent;
  const int32_t *frontal_off, *frontals;               /* getCliqFrontalVarIds */
  const int32_t *separator_off, *separators;           /* getCliqSeparatorVarIds */
  const int32_t *potential_off, *potentials;           /* getCliqueData(cliq).potentials */
  const int32_t *directFrtlMsg_off, *directFrtlMsg;    /* .directFrtlMsgIDs */
  const int32_t *msgskip_off, *msgskip;                /* .msgskipIDs */
  const int32_t *itervar_off, *itervar;                /* .itervarIDs */
  const int32_t *directPriorMsg_off, *directPriorMsg;  /* .directPriorMsgIDs */
} iif_tree_desc;

typedef struct {
  int32_t N;                  /* particles per belief */
  int32_t gibbsIters;         /* SolverParams.gibbsIters (3) */
  int32_t downIters;          /* MCIters of the down solve (3) */
  int32_t downsolve;          /* SolverParams.downsolve */
  int32_t lanes;              /* parallel graph branches for independent sub-trees (0 = none, <= 8) */
  int32_t forward_copies;     /* collapse separator copy chains */
  int32_t useMsgLikelihoods;  /* SolverParams.useMsgLikelihoods: joint up messages (differentials as IIF_S_DECONV ops + one
                                 MsgPrior per class), kept for the down solve (CliqueStateMachine.jl:825-834) */
  int32_t call_base;          /* first Philox call id of the pass (prop k uses call_base + 16 k) */
  double inflation;           /* SolverParams.inflation, for the message priors */
} iif_plan_opts;

typedef struct iifb200_plan iifb200_plan;
int32_t iifb200_plan_tree(const iif_graph_desc* graph, const iif_tree_desc* tree, const iif_plan_opts* opts,
                          iifb200_plan** plan_out);
const char* iifb200_plan_error(void); /* message of the last failed iifb200_plan_tree on this thread */
void iifb200_plan_free(iifb200_plan* plan);
/* counts[16]: nslots, nfactors, ndists, nparams, nprops, nops, nwaves, n_conv, n_prod, n_msgs, up_last_wave, nvars,
 * ndeconvs, 0.. */
int32_t iifb200_plan_counts(const iifb200_plan* plan, int32_t* counts);
/* copies the plan's tables into caller arrays sized from iifb200_plan_counts (any pointer may be NULL) */
int32_t iifb200_plan_export(const iifb200_plan* plan, iif_slot_desc* slots, iif_factor_desc* factors,
                            iif_dist_desc* dists, double* dparams, iif_prop_op* props, iif_sched_op* ops,
                            int32_t* wave_off /* nwaves + 1 */);
int32_t iifb200_plan_export_deconvs(const iifb200_plan* plan, iif_deconv_op* deconvs /* counts[12] */);
/* iifb200_set_graph + iifb200_schedule_build of a plan in one call; beliefs of the graph variables then go to slots
 * 0..nvars-1 (iifb200_upload_slots), iifb200_schedule_run solves, iifb200_download_slots reads the posteriors. */
int32_t iifb200_plan_upload(iifb200_ctx* ctx, const iifb200_plan* plan, const iif_solver_params* sp,
                            void* ext_arena, int32_t* schedule_id_out);

/* Nested-dissection variable elimination order for `solveTree!(fg; eliminationOrder = ...)` (SolverAPI.jl:338;
 * getEliminationOrder BayesNet.jl:19-65 is the reference's default).  The Bayes tree of the default (QR / natural)
 * order of a pose chain is a path — every clique waits for its child, nothing runs in parallel — while recursive
 * bisection (BFS level sets from a peripheral variable, separators eliminated last) gives a tree of depth O(log n)
 * whose levels are the wide waves of iifb200_plan_tree.  Variables are 0..nvars-1 in graph order, factor f touches
 * fac_vars[fac_off[f] .. fac_off[f+1]); order_out receives the nvars variable ids, first eliminated first.  Host only. */
int32_t iifb200_elimination_order_nd(int32_t nvars, int32_t nfactors, const int32_t* fac_off, const int32_t* fac_vars,
                                     int32_t* order_out);
/* Independent-set order (generalised odd-even / cyclic reduction), the recommended one: rounds of eliminating a maximal
 * independent set of the variables whose degree in the current elimination graph is <= minimum + slack (greedy in
 * variable order; fill-in edges join the neighbours of every eliminated variable).  The variables of a round are
 * pairwise non-adjacent, so their cliques are siblings (one wide wave), the rounds number O(log n) on chains and grids,
 * and the degree bound keeps cliques small (few frontals => short in-clique Gibbs iterations).  Measured against the
 * bisection above: 1000-pose chain 5.33 vs 5.81 ms per solve, 5000-pose grid with loop closures 94.6 vs 118.9 ms
 * (profiles/r02_knobs.txt).  slack = 1 unless there is a reason. */
int32_t iifb200_elimination_order_is(int32_t nvars, int32_t nfactors, const int32_t* fac_off, const int32_t* fac_vars,
                                     int32_t slack, int32_t* order_out);

/* ---- instrumentation ---------------------------------------------------------------- */
/* kernels launched by this ctx since init (for bench.py "gpu_launches") */
int64_t iifb200_launch_count(const iifb200_ctx* ctx);
/* Make the ctx launch on a caller-owned cudaStream_t (e.g. the torch stream NCCL collectives are
 * ordered on), so separator-message exchanges need no host synchronisation.  NULL restores the
 * ctx's own stream.  Captured CUDA graphs stay valid (they are launched into the new stream). */
int32_t iifb200_set_stream(iifb200_ctx* ctx, void* stream);
/* cudaStream_t used by the ctx (as void*), so callers can record CUDA events on it */
void* iifb200_stream(iifb200_ctx* ctx);
/* Runs waves [first,last) WITHOUT the CUDA graph, bracketing every kernel launch with CUDA events on
 * the ctx stream, and returns per-kernel totals: ms[0..2] / launches[0..2] / blocks[0..2] for the
 * convolution, product and copy kernels.  Used by bench.py for the live roofline figure. */
int32_t iifb200_schedule_profile(iifb200_ctx* ctx, int32_t schedule_id, int32_t first_wave,
                                 int32_t last_wave, float* ms, int32_t* launches, int64_t* blocks);
/* time the last schedule_run / *_batch call took on the device, in ms (CUDA events on the
 * ctx stream; valid after iifb200_sync) */
float iifb200_last_elapsed_ms(iifb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* IIFB200_H */
