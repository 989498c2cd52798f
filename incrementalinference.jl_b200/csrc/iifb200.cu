// iifb200.cu — C-ABI host side of libiifb200.so (see include/iifb200.h for the contract and
// the reference interfaces each entry point replaces).  No torch, no Julia, no CPU fallback:
// every hot-path entry point launches the sm_100a kernels in iif_conv.cuh / iif_product.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "iif_deconv.cuh"
#include "iif_ppe.cuh"
#include "iif_product.cuh"

#define IIFB200_VERSION 200

// iif_plan.cpp (host planner)
const std::vector<iif_slot_desc>& iif_plan_slots(const iifb200_plan* p);
const std::vector<iif_factor_desc>& iif_plan_factors(const iifb200_plan* p);
const std::vector<iif_dist_desc>& iif_plan_dists(const iifb200_plan* p);
const std::vector<double>& iif_plan_dparams(const iifb200_plan* p);
const std::vector<iif_prop_op>& iif_plan_props(const iifb200_plan* p);
const std::vector<iif_sched_op>& iif_plan_ops(const iifb200_plan* p);
const std::vector<int32_t>& iif_plan_wave_off(const iifb200_plan* p);
const std::vector<iif_deconv_op>& iif_plan_deconvs(const iifb200_plan* p);

// upper bound of any kernel's static shared memory (the 2 KB Gaussian table appears twice in the product kernel:
// its own and the bandwidth search's; ptxas reports 4.8 KB for the convolution kernel)
#define IIF_STATIC_SMEM_RESERVE 8192
static std::string g_init_error;

struct TreeHost {
  int L = 0, nn = 0;
  int16_t* d_blob = nullptr;
};

#define IIF_MAX_LANES 8
// One wave = the independent ops of one dependency level.  A wave whose ops all carry a lane id (iif_sched_op.lane
// != 0) is split into per-lane segments that are launched on per-lane streams inside the captured CUDA graph: lanes
// are independent sub-trees, so one lane's kernels need not wait for another lane's wave to drain.  A wave that holds
// any lane-0 op is a full barrier (one segment, main stream).
struct Seg {
  int lane = 0;
  int conv0 = 0, nconv = 0, prod0 = 0, nprod = 0, copy0 = 0, ncopy = 0, dcv0 = 0, ndcv = 0;
  int push0 = 0, npush = 0, wait0 = 0, nwait = 0;   // peer messages: waits run first, pushes last
};
struct Wave {
  bool conv_rare = false;  // some convolution of the wave needs the kernel variant with the SO(3) / Nelder-Mead paths
  std::vector<Seg> segs;
  int nconv = 0, nprod = 0, ncopy = 0, ndcv = 0;  // wave totals: CTA size / cluster choice follow the whole wave's width
  size_t prod_smem = 0, conv_smem = 0, dcv_smem = 0;
  int dcv_maxN = 2;
  int maxN = 2;
};

struct Schedule {
  bool pooled = false;  // device arrays live in the ctx pool (propagate_batch): not freed with the schedule
  std::vector<Wave> waves;
  ConvTask* d_conv = nullptr;
  ProdTask* d_prod = nullptr;
  int32_t* d_copy = nullptr;
  DeconvSlotTask* d_dcv = nullptr;
  PushTask* d_push = nullptr;     // multi-GPU schedules (iifb200_schedule_build_dist)
  int32_t* d_wait = nullptr;
  bool dist = false;
  double* d_scratch = nullptr;
  int32_t* d_status = nullptr;  // nconv + nprod + ndcv device-side status codes
  int nconv = 0, nprod = 0, ndcv = 0;
  int nstatus() const { return nconv + nprod + ndcv; }
  std::map<std::pair<int, int>, std::pair<cudaGraphExec_t, int>> graphs;  // (exec, kernel nodes)
  std::vector<cudaEvent_t> events;  // fork / join events of the captured graphs
};

struct iifb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t lane_stream[IIF_MAX_LANES + 1] = {};  // created on first use (schedules with lanes)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  std::string err;
  // graph tables
  std::vector<iif_slot_desc> slots;
  std::vector<iif_factor_desc> factors;
  std::vector<iif_dist_desc> dists;
  iif_solver_params sp{};
  DeviceGraph dg{};
  void* arena = nullptr;
  bool arena_owned = false;
  int64_t total_doubles = 0;
  void* d_tables = nullptr;
  int32_t* d_err = nullptr;
  iif_solver_params* d_sp = nullptr;
  std::vector<void*> pinned;
  // ball-tree structures per N
  TreeHost trees[IIF_MAX_POINTS + 1];
  TreeStruct h_trees[IIF_MAX_POINTS + 1];
  TreeStruct* d_trees = nullptr;
  std::vector<Schedule*> schedules;
  // grow-only device scratch of the *_batch entry points and of propagate_batch's one-wave schedule: a call never
  // pays cudaMalloc / cudaFree once the pools have reached their working size (boundary B3 is one call per belief)
  void* pool[2] = {nullptr, nullptr};
  size_t pool_cap[2] = {0, 0};
  size_t arena_cap = 0, tables_cap = 0;
  // peer memory (one process per GPU; CUDA IPC): arenas and message flags of every rank, this rank's own included
  int world = 1, rank = 0;
  std::vector<void*> peer_arena;
  std::vector<int32_t*> peer_flags;
  int32_t* d_flags = nullptr;
  int32_t* d_epoch = nullptr;
  int nflags = 0;
  int64_t launches = 0;
  int max_smem_optin = 0;  // dynamic budget of a kernel = this - IIF_STATIC_SMEM_RESERVE
  bool defer_sync = false; // inside iifb200_propagate_once: set_graph leaves the stream unsynchronised
  int num_sms = 148;
};

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      char b_[512];                                                                    \
      snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
               __FILE__, __LINE__);                                                    \
      ctx->err = b_;                                                                   \
      return IIF_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

static int32_t fail(iifb200_ctx* ctx, int32_t code, const std::string& msg) {
  ctx->err = msg;
  return code;
}

struct Carver {   // first pass (base == nullptr) sizes, second pass hands out 256-byte aligned pieces
  char* base = nullptr;
  size_t off = 0;
  template <typename T> T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

static const char* status_name(int st) {
  switch (st) {
    case IIF_ERR_ARG: return "bad argument / descriptor";
    case IIF_ERR_CUDA: return "CUDA error";
    case IIF_ERR_UNSUPPORTED: return "unsupported factor / distribution kind on the device";
    case IIF_ERR_STATE: return "state error (uninitialised source belief?)";
    default: return "unknown";
  }
}

// ---- balanced median-split tree structure for N points (depends on N only) ----------------
static int32_t ensure_tree(iifb200_ctx* ctx, int N) {
  if (N < 2 || N > IIF_MAX_POINTS) return fail(ctx, IIF_ERR_ARG, "N out of range [2, IIF_MAX_POINTS]");
  if (ctx->trees[N].d_blob) return IIF_OK;
  const int L = (int)std::floor(std::log((double)N) / std::log(2.0) + 1.0);
  std::vector<int16_t> lev_off(L + 2), lo, hi, child, node_at((size_t)L * N);
  lo.push_back(0);
  hi.push_back((int16_t)(N - 1));
  lev_off[0] = 0;
  lev_off[1] = 1;
  for (int l = 0; l < L; ++l) {
    for (int z = lev_off[l]; z < lev_off[l + 1]; ++z) {
      int a = lo[z], b = hi[z];
      child.push_back((int16_t)((int)lo.size() - lev_off[l + 1]));
      for (int p = a; p <= b; ++p) node_at[(size_t)l * N + p] = (int16_t)(z - lev_off[l]);
      if (a == b) { lo.push_back((int16_t)a); hi.push_back((int16_t)b); continue; }
      int mid = (a + b) / 2;
      lo.push_back((int16_t)a); hi.push_back((int16_t)mid);
      lo.push_back((int16_t)(mid + 1)); hi.push_back((int16_t)b);
    }
    lev_off[l + 2] = (int16_t)lo.size();
  }
  for (int z = lev_off[L]; z < lev_off[L + 1]; ++z) child.push_back((int16_t)(z - lev_off[L]));
  const int nn = (int)lo.size();
  // blob: lev_off | lo | hi | child | node_at
  std::vector<int16_t> blob;
  auto push = [&](const std::vector<int16_t>& v) { size_t o = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); return o; };
  size_t o0 = push(lev_off), o1 = push(lo), o2 = push(hi), o3 = push(child), o4 = push(node_at);
  int16_t* d = nullptr;
  CK(cudaMalloc(&d, blob.size() * sizeof(int16_t)));
  CK(cudaMemcpyAsync(d, blob.data(), blob.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->trees[N].L = L;
  ctx->trees[N].nn = nn;
  ctx->trees[N].d_blob = d;
  TreeStruct ts;
  ts.L = L; ts.nn = nn;
  ts.lev_off = d + o0; ts.lo = d + o1; ts.hi = d + o2; ts.child = d + o3; ts.node_at = d + o4;
  ctx->h_trees[N] = ts;
  CK(cudaMemcpyAsync(ctx->d_trees + N, &ts, sizeof(ts), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
}

// launch with an optional thread-block cluster dimension (cluster == 1: a plain launch)
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), int tasks, int cluster, int threads, size_t smem, cudaStream_t st,
                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tasks * cluster));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (cluster > 1) {
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);  // errors surface through cudaGetLastError at the call site
}

// =============================================================================================
extern "C" {

int32_t iifb200_version(void) { return IIFB200_VERSION; }

const char* iifb200_last_error(const iifb200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int32_t iifb200_init(int32_t device_ordinal, iifb200_ctx** ctx_out) {
  if (!ctx_out) return IIF_ERR_ARG;
  *ctx_out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_init_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libiifb200 has no CPU fallback)";
    return IIF_ERR_CUDA;
  }
  if (device_ordinal < 0 || device_ordinal >= ndev) { g_init_error = "device ordinal out of range"; return IIF_ERR_ARG; }
  iifb200_ctx* ctx = new iifb200_ctx();
  ctx->device = device_ordinal;
  ctx->sp.spreadNH = 3.0; ctx->sp.nullSurplusAdd = 0.3; ctx->sp.inflateCycles = 3; ctx->sp.gibbsNiter = 1; ctx->sp.seed = 42;
  auto bail = [&](const char* what, cudaError_t ee) {
    g_init_error = std::string(what) + ": " + cudaGetErrorString(ee);
    delete ctx;
    return IIF_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaMalloc(&ctx->d_trees, sizeof(TreeStruct) * (IIF_MAX_POINTS + 1))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemset(ctx->d_trees, 0, sizeof(TreeStruct) * (IIF_MAX_POINTS + 1))) != cudaSuccess) return bail("cudaMemset", e);
  if ((e = cudaMalloc(&ctx->d_err, sizeof(int32_t))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemset(ctx->d_err, 0, sizeof(int32_t))) != cudaSuccess) return bail("cudaMemset", e);
  if ((e = cudaMalloc(&ctx->d_sp, sizeof(iif_solver_params))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemcpy(ctx->d_sp, &ctx->sp, sizeof(iif_solver_params), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", e);
  ctx->dg.sp = ctx->d_sp;
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_ordinal);
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device_ordinal);
  if ((e = cudaFuncSetAttribute(iif_product_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE)) != cudaSuccess)
    return bail("cudaFuncSetAttribute(product smem)", e);
  if ((e = cudaFuncSetAttribute(iif_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE)) != cudaSuccess)
    return bail("cudaFuncSetAttribute(conv smem)", e);
  if ((e = cudaFuncSetAttribute(iif_conv_kernel_rare, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE)) != cudaSuccess)
    return bail("cudaFuncSetAttribute(conv smem, rare variant)", e);
  if ((e = cudaFuncSetAttribute(iif_bandwidth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE)) != cudaSuccess)
    return bail("cudaFuncSetAttribute(bandwidth smem)", e);
  if ((e = cudaFuncSetAttribute(iif_deconv_slot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE)) != cudaSuccess)
    return bail("cudaFuncSetAttribute(deconv smem)", e);
  *ctx_out = ctx;
  return IIF_OK;
}

static void free_schedule(Schedule* s) {
  if (!s) return;
  for (auto& kv : s->graphs) cudaGraphExecDestroy(kv.second.first);
  for (auto e : s->events) cudaEventDestroy(e);
  cudaFree(s->d_push); cudaFree(s->d_wait);
  if (!s->pooled) { cudaFree(s->d_conv); cudaFree(s->d_prod); cudaFree(s->d_copy); cudaFree(s->d_dcv); cudaFree(s->d_scratch); cudaFree(s->d_status); }
  delete s;
}

// schedules go; the arena and the descriptor tables stay allocated (grow-only) unless `release`
static void free_graph(iifb200_ctx* ctx, bool release) {
  for (auto*& s : ctx->schedules) { free_schedule(s); s = nullptr; }
  ctx->schedules.clear();
  if (!ctx->arena_owned) { ctx->arena = nullptr; ctx->arena_cap = 0; }
  if (release) {
    if (ctx->arena_owned && ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr; ctx->arena_cap = 0;
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    ctx->d_tables = nullptr; ctx->tables_cap = 0;
    for (int k = 0; k < 2; ++k) { if (ctx->pool[k]) cudaFree(ctx->pool[k]); ctx->pool[k] = nullptr; ctx->pool_cap[k] = 0; }
  }
}

// grow-only pool k; the previous contents are not preserved.  Carve with `Carver`.
static cudaError_t pool_reserve(iifb200_ctx* ctx, int k, size_t bytes, char** base) {
  if (bytes > ctx->pool_cap[k]) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return e;
    if (ctx->pool[k]) cudaFree(ctx->pool[k]);
    ctx->pool[k] = nullptr; ctx->pool_cap[k] = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    e = cudaMalloc(&ctx->pool[k], want);
    if (e != cudaSuccess) return e;
    ctx->pool_cap[k] = want;
  }
  *base = (char*)ctx->pool[k];
  return cudaSuccess;
}

void iifb200_free(iifb200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < (int)ctx->peer_arena.size(); ++r) {
    if (r == ctx->rank) continue;
    if (ctx->peer_arena[r]) cudaIpcCloseMemHandle(ctx->peer_arena[r]);
    if (ctx->peer_flags[r]) cudaIpcCloseMemHandle(ctx->peer_flags[r]);
  }
  free_graph(ctx, true);
  cudaFree(ctx->d_flags);
  cudaFree(ctx->d_epoch);
  for (auto& t : ctx->trees) if (t.d_blob) cudaFree(t.d_blob);
  cudaFree(ctx->d_trees);
  cudaFree(ctx->d_err);
  cudaFree(ctx->d_sp);
  for (void* p : ctx->pinned) cudaFreeHost(p);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->own_stream);
  for (auto st : ctx->lane_stream) if (st) cudaStreamDestroy(st);
  delete ctx;
}

static int64_t layout_slots(int32_t nslots, iif_slot_desc* slots) {
  int64_t off = 0;
  for (int s = 0; s < nslots; ++s) {
    slots[s].pts_off = (int32_t)off;
    off += (int64_t)slots[s].cap * slots[s].dim;
  }
  return off;
}

int64_t iifb200_arena_bytes(int32_t nslots, const iif_slot_desc* slots) {
  int64_t tot = 0;
  for (int s = 0; s < nslots; ++s) tot += (int64_t)slots[s].cap * slots[s].dim;
  int64_t ns = nslots > 0 ? nslots : 1;
  return 8 * (tot + 2 * ns * IIF_MAX_DIM) + 4 * 2 * ns + 64;
}

int32_t iifb200_set_graph(iifb200_ctx* ctx, int32_t nslots, iif_slot_desc* slots, int32_t nfactors,
                          const iif_factor_desc* factors, int32_t ndists, const iif_dist_desc* dists,
                          int32_t nparams, const double* dparams, const iif_solver_params* sp, void* ext_arena) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  if (nslots < 1 || !slots || nfactors < 0 || ndists < 0 || !sp) return fail(ctx, IIF_ERR_ARG, "set_graph: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->defer_sync) CK(cudaStreamSynchronize(ctx->stream));
  free_graph(ctx, false);
  for (int s = 0; s < nslots; ++s) {
    if (slots[s].dim < 1 || slots[s].dim > IIF_MAX_DIM) return fail(ctx, IIF_ERR_ARG, "slot dim out of range");
    if (slots[s].cap < 1 || slots[s].cap > IIF_MAX_POINTS) return fail(ctx, IIF_ERR_ARG, "slot capacity out of range");
  }
  for (int f = 0; f < nfactors; ++f) {
    const iif_factor_desc& F = factors[f];
    if (F.kind < IIF_F_PRIOR || F.kind > IIF_F_SO3_RELATIVE)
      return fail(ctx, IIF_ERR_UNSUPPORTED, "factor kind has no device residual (no CPU fallback)");
    if (F.arity < 1 || F.arity > IIF_MAX_ARITY) return fail(ctx, IIF_ERR_ARG, "factor arity out of range");
    if (F.dist < 0 || F.dist >= ndists) return fail(ctx, IIF_ERR_ARG, "factor distribution index out of range");
    if (F.nmh != 0 && F.nmh != F.arity) return fail(ctx, IIF_ERR_ARG, "multihypo length must equal arity");
    for (int v = 0; v < F.arity; ++v)
      if (F.slot[v] < 0 || F.slot[v] >= nslots) return fail(ctx, IIF_ERR_ARG, "factor slot out of range");
    if (F.kind == IIF_F_SO3_PRIOR || F.kind == IIF_F_SO3_RELATIVE) {  // rotation-vector coordinates
      if (F.zdim != 3) return fail(ctx, IIF_ERR_ARG, "SO(3) factors need zdim == 3");
      for (int v = 0; v < F.arity; ++v)
        if (slots[F.slot[v]].dim != 3 || !(slots[F.slot[v]].circ_mask & IIF_MANI_SO3))
          return fail(ctx, IIF_ERR_ARG, "SO(3) factors need (dim 3, IIF_MANI_SO3) variables");
    }
    if (F.kind == IIF_F_SE2_RELATIVE) {  // SpecialEuclidean(2) coordinates are (x, y, theta)
      if (F.zdim != 3 || F.arity < 2) return fail(ctx, IIF_ERR_ARG, "SE2 relative factor needs zdim == 3 and arity >= 2");
      for (int v = 0; v < F.arity; ++v)
        if (slots[F.slot[v]].dim != 3 || slots[F.slot[v]].circ_mask != 4)
          return fail(ctx, IIF_ERR_ARG, "SE2 relative factor needs (dim 3, circ_mask 0b100) variables");
    }
  }
  for (int k = 0; k < ndists; ++k) {
    if (dists[k].kind < IIF_D_NORMAL || dists[k].kind > IIF_D_SAMPLES)
      return fail(ctx, IIF_ERR_UNSUPPORTED, "distribution kind has no device sampler");
    if (dists[k].kind == IIF_D_KDE && (dists[k].slot < 0 || dists[k].slot >= nslots))
      return fail(ctx, IIF_ERR_ARG, "KDE distribution slot out of range");
    if (dists[k].kind == IIF_D_SAMPLES && (dists[k].ncomp < 1 || dists[k].poff < 0 ||
                                           (int64_t)dists[k].poff + (int64_t)dists[k].ncomp * dists[k].dim > nparams))
      return fail(ctx, IIF_ERR_ARG, "sample-table distribution exceeds the parameter array");
  }
  ctx->total_doubles = layout_slots(nslots, slots);
  ctx->slots.assign(slots, slots + nslots);
  ctx->factors.assign(factors, factors + nfactors);
  ctx->dists.assign(dists, dists + ndists);
  ctx->sp = *sp;
  const int64_t bytes = iifb200_arena_bytes(nslots, slots);
  if (ext_arena) {
    if (ctx->arena_owned && ctx->arena) cudaFree(ctx->arena);
    ctx->arena = ext_arena; ctx->arena_owned = false; ctx->arena_cap = 0;
  } else if (!ctx->arena_owned || !ctx->arena || ctx->arena_cap < (size_t)bytes) {   // grow-only: re-used across set_graph calls
    if (ctx->arena_owned && ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr;
    const size_t want = (size_t)bytes + (size_t)bytes / 4;
    CK(cudaMalloc(&ctx->arena, want));
    ctx->arena_owned = true; ctx->arena_cap = want;
  }
  CK(cudaMemsetAsync(ctx->arena, 0, bytes, ctx->stream));
  // descriptor tables in one allocation
  const size_t b_slots = sizeof(iif_slot_desc) * nslots, b_fac = sizeof(iif_factor_desc) * std::max(nfactors, 1),
               b_dist = sizeof(iif_dist_desc) * std::max(ndists, 1), b_par = sizeof(double) * std::max(nparams, 1);
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t tb = al(b_slots) + al(b_fac) + al(b_dist) + al(b_par);
  if (!ctx->d_tables || ctx->tables_cap < tb) {
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    ctx->d_tables = nullptr;
    CK(cudaMalloc(&ctx->d_tables, tb + tb / 4));
    ctx->tables_cap = tb + tb / 4;
  }
  char* p = (char*)ctx->d_tables;
  DeviceGraph& dg = ctx->dg;
  dg.slots = (iif_slot_desc*)p; p += al(b_slots);
  dg.factors = (iif_factor_desc*)p; p += al(b_fac);
  dg.dists = (iif_dist_desc*)p; p += al(b_dist);
  dg.dparams = (double*)p;
  CK(cudaMemcpyAsync((void*)dg.slots, slots, b_slots, cudaMemcpyHostToDevice, ctx->stream));
  if (nfactors) CK(cudaMemcpyAsync((void*)dg.factors, factors, sizeof(iif_factor_desc) * nfactors, cudaMemcpyHostToDevice, ctx->stream));
  if (ndists) CK(cudaMemcpyAsync((void*)dg.dists, dists, sizeof(iif_dist_desc) * ndists, cudaMemcpyHostToDevice, ctx->stream));
  if (nparams) CK(cudaMemcpyAsync((void*)dg.dparams, dparams, sizeof(double) * nparams, cudaMemcpyHostToDevice, ctx->stream));
  double* a = (double*)ctx->arena;
  dg.pts = a;
  dg.bw = a + ctx->total_doubles;
  dg.ipc = dg.bw + (int64_t)nslots * IIF_MAX_DIM;
  dg.npts = (int32_t*)(dg.ipc + (int64_t)nslots * IIF_MAX_DIM);
  dg.flags = dg.npts + nslots;
  dg.sp = ctx->d_sp;
  CK(cudaMemcpyAsync(ctx->d_sp, sp, sizeof(iif_solver_params), cudaMemcpyHostToDevice, ctx->stream));
  dg.nslots = nslots; dg.nfactors = nfactors; dg.ndists = ndists;
  if (!ctx->defer_sync) CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_set_graph: ") + e.what());
  }
}

int32_t iifb200_set_solver_params(iifb200_ctx* ctx, const iif_solver_params* sp) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx || !sp) return IIF_ERR_ARG;
  // Kernels read the params through a device pointer, so captured CUDA graphs stay valid; the copy
  // is stream-ordered after earlier launches.  nullSurplusAdd is baked into schedules at build time.
  CK(cudaSetDevice(ctx->device));
  ctx->sp = *sp;
  CK(cudaMemcpyAsync(ctx->d_sp, &ctx->sp, sizeof(iif_solver_params), cudaMemcpyHostToDevice, ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_set_solver_params: ") + e.what());
  }
}

// ---- belief I/O --------------------------------------------------------------------------
#define NEED_GRAPH() do { if (!ctx) return IIF_ERR_ARG; if (!ctx->arena) return fail(ctx, IIF_ERR_STATE, "no graph uploaded (call iifb200_set_graph first)"); } while (0)

int32_t iifb200_upload_belief(iifb200_ctx* ctx, int32_t slot, int32_t npts, const double* pts, const double* bw,
                              int32_t initialized) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, IIF_ERR_ARG, "slot out of range");
  const iif_slot_desc& S = ctx->slots[slot];
  if (npts < 0 || npts > S.cap) return fail(ctx, IIF_ERR_ARG, "npts exceeds slot capacity");
  if (npts > 0) CK(cudaMemcpyAsync(ctx->dg.pts + S.pts_off, pts, sizeof(double) * npts * S.dim, cudaMemcpyHostToDevice, ctx->stream));
  double b[IIF_MAX_DIM] = {0, 0, 0, 0};
  if (bw) for (int c = 0; c < S.dim; ++c) b[c] = bw[c];
  CK(cudaMemcpyAsync(ctx->dg.bw + (int64_t)slot * IIF_MAX_DIM, b, sizeof(b), cudaMemcpyHostToDevice, ctx->stream));
  int32_t fl = initialized ? 1 : 0;
  CK(cudaMemcpyAsync(ctx->dg.npts + slot, &npts, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->dg.flags + slot, &fl, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_upload_belief: ") + e.what());
  }
}

int32_t iifb200_download_belief(iifb200_ctx* ctx, int32_t slot, int32_t* npts, double* pts, double* bw, double* ipc) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, IIF_ERR_ARG, "slot out of range");
  const iif_slot_desc& S = ctx->slots[slot];
  int32_t n = 0;
  CK(cudaMemcpyAsync(&n, ctx->dg.npts + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (npts) *npts = n;
  if (pts && n > 0) CK(cudaMemcpyAsync(pts, ctx->dg.pts + S.pts_off, sizeof(double) * n * S.dim, cudaMemcpyDeviceToHost, ctx->stream));
  double b[IIF_MAX_DIM], q[IIF_MAX_DIM];
  CK(cudaMemcpyAsync(b, ctx->dg.bw + (int64_t)slot * IIF_MAX_DIM, sizeof(b), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(q, ctx->dg.ipc + (int64_t)slot * IIF_MAX_DIM, sizeof(q), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bw) for (int c = 0; c < S.dim; ++c) bw[c] = b[c];
  if (ipc) for (int c = 0; c < S.dim; ++c) ipc[c] = q[c];
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_download_belief: ") + e.what());
  }
}

int32_t iifb200_upload_all(iifb200_ctx* ctx, const double* pts, const double* bw, const int32_t* npts, const int32_t* flags) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  const int64_t ns = (int64_t)ctx->slots.size();
  if (pts) CK(cudaMemcpyAsync(ctx->dg.pts, pts, sizeof(double) * ctx->total_doubles, cudaMemcpyHostToDevice, ctx->stream));
  if (bw) CK(cudaMemcpyAsync(ctx->dg.bw, bw, sizeof(double) * ns * IIF_MAX_DIM, cudaMemcpyHostToDevice, ctx->stream));
  if (npts) CK(cudaMemcpyAsync(ctx->dg.npts, npts, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, ctx->stream));
  if (flags) CK(cudaMemcpyAsync(ctx->dg.flags, flags, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, ctx->stream));
  return IIF_OK;  // stream-ordered: later launches on the ctx stream see the data
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_upload_all: ") + e.what());
  }
}

int32_t iifb200_download_all(iifb200_ctx* ctx, double* pts, double* bw, double* ipc, int32_t* npts) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  const int64_t ns = (int64_t)ctx->slots.size();
  if (pts) CK(cudaMemcpyAsync(pts, ctx->dg.pts, sizeof(double) * ctx->total_doubles, cudaMemcpyDeviceToHost, ctx->stream));
  if (bw) CK(cudaMemcpyAsync(bw, ctx->dg.bw, sizeof(double) * ns * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  if (ipc) CK(cudaMemcpyAsync(ipc, ctx->dg.ipc, sizeof(double) * ns * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  if (npts) CK(cudaMemcpyAsync(npts, ctx->dg.npts, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_download_all: ") + e.what());
  }
}

int32_t iifb200_upload_slots(iifb200_ctx* ctx, int32_t first, int32_t count, const double* pts, const double* bw,
                             const int32_t* npts, const int32_t* flags) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  const int ns = (int)ctx->slots.size();
  if (first < 0 || count < 1 || first + count > ns) return fail(ctx, IIF_ERR_ARG, "upload_slots: slot range out of bounds");
  const int64_t o0 = ctx->slots[first].pts_off;
  const int64_t o1 = (first + count < ns) ? ctx->slots[first + count].pts_off : ctx->total_doubles;
  if (pts) CK(cudaMemcpyAsync(ctx->dg.pts + o0, pts, sizeof(double) * (o1 - o0), cudaMemcpyHostToDevice, ctx->stream));
  if (bw) CK(cudaMemcpyAsync(ctx->dg.bw + (int64_t)first * IIF_MAX_DIM, bw, sizeof(double) * count * IIF_MAX_DIM, cudaMemcpyHostToDevice, ctx->stream));
  if (npts) CK(cudaMemcpyAsync(ctx->dg.npts + first, npts, sizeof(int32_t) * count, cudaMemcpyHostToDevice, ctx->stream));
  if (flags) CK(cudaMemcpyAsync(ctx->dg.flags + first, flags, sizeof(int32_t) * count, cudaMemcpyHostToDevice, ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_upload_slots: ") + e.what());
  }
}

int32_t iifb200_download_slots(iifb200_ctx* ctx, int32_t first, int32_t count, double* pts, double* bw, double* ipc,
                               int32_t* npts) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  const int ns = (int)ctx->slots.size();
  if (first < 0 || count < 1 || first + count > ns) return fail(ctx, IIF_ERR_ARG, "download_slots: slot range out of bounds");
  const int64_t o0 = ctx->slots[first].pts_off;
  const int64_t o1 = (first + count < ns) ? ctx->slots[first + count].pts_off : ctx->total_doubles;
  if (pts) CK(cudaMemcpyAsync(pts, ctx->dg.pts + o0, sizeof(double) * (o1 - o0), cudaMemcpyDeviceToHost, ctx->stream));
  if (bw) CK(cudaMemcpyAsync(bw, ctx->dg.bw + (int64_t)first * IIF_MAX_DIM, sizeof(double) * count * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  if (ipc) CK(cudaMemcpyAsync(ipc, ctx->dg.ipc + (int64_t)first * IIF_MAX_DIM, sizeof(double) * count * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  if (npts) CK(cudaMemcpyAsync(npts, ctx->dg.npts + first, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
  return IIF_OK;  // asynchronous: iifb200_sync before reading
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_download_slots: ") + e.what());
  }
}

int32_t iifb200_host_alloc(iifb200_ctx* ctx, int64_t bytes, void** ptr_out) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx || !ptr_out || bytes < 1) return IIF_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  void* p = nullptr;
  CK(cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault));
  ctx->pinned.push_back(p);
  *ptr_out = p;
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_host_alloc: ") + e.what());
  }
}

int32_t iifb200_host_free(iifb200_ctx* ctx, void* ptr) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx || !ptr) return IIF_ERR_ARG;
  auto it = std::find(ctx->pinned.begin(), ctx->pinned.end(), ptr);
  if (it == ctx->pinned.end()) return fail(ctx, IIF_ERR_ARG, "host_free: pointer not from iifb200_host_alloc");
  ctx->pinned.erase(it);
  CK(cudaFreeHost(ptr));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_host_free: ") + e.what());
  }
}

int32_t iifb200_slot_device_ptr(iifb200_ctx* ctx, int32_t slot, void** pts_ptr, void** bw_ptr) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, IIF_ERR_ARG, "slot out of range");
  if (pts_ptr) *pts_ptr = ctx->dg.pts + ctx->slots[slot].pts_off;
  if (bw_ptr) *bw_ptr = ctx->dg.bw + (int64_t)slot * IIF_MAX_DIM;
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_slot_device_ptr: ") + e.what());
  }
}

// ---- helpers -------------------------------------------------------------------------------
// CTA size of a launch: wide (throughput-bound) launches use the smallest CTA that still gives every
// particle its own thread, so several CTAs share an SM and no warp idles in the per-particle stages;
// narrow (latency-bound) launches use 512 threads to split each belief's work further.
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
// CTA size of a launch.  A belief's work can be split over 128..512 threads; registers (126 per thread)
// allow 65536 / (126 * threads) CTAs per SM, shared memory caps the convolution at 4 and the product at 2.
// The largest CTA whose launch still fits in ONE round of resident CTAs wins (per-belief latency is what a
// narrow wave costs); launches beyond that use the smallest CTA (most beliefs in flight per SM).
static int pick_threads(iifb200_ctx* ctx, int grid, int maxN, int small_min = 128) {
  static const int wide = env_int("IIFB200_WIDE_THREADS", 0);  // tuning knob: CTA size of wide launches
  const int fit = (maxN + 31) / 32 * 32;
  int small = std::min(std::max(wide > 0 ? wide : small_min, fit), IIF_MAX_THREADS);
  if (grid <= ctx->num_sms) return IIF_MAX_THREADS;
  if (grid <= 2 * ctx->num_sms) return std::min(IIF_MAX_THREADS, std::max(256, fit));
  return small;
}
static int pick_threads_prod(iifb200_ctx* ctx, int grid, int maxN) {
  static const int wide = env_int("IIFB200_PROD_WIDE_THREADS", 256);
  static const int force = env_int("IIFB200_FORCE_PROD_THREADS", 0);  // development: probe a lone product in its wide-wave shape
  if (force > 0) return std::min(IIF_MAX_THREADS, std::max(force, (maxN + 31) / 32 * 32));
  if (wide <= 0 || grid <= ctx->num_sms) return IIF_MAX_THREADS;
  return std::min(IIF_MAX_THREADS, std::max(wide, (maxN + 31) / 32 * 32));
}

// Narrow launches (fewer beliefs than SMs / 3, i.e. the top of the tree) give every belief a thread-block
// cluster of IIF_SPEC_CLUSTER CTAs that share the golden-section bandwidth search speculatively
// (iif_device.cuh); everything else launches plain CTAs.
static int pick_cluster(iifb200_ctx* ctx, int grid) {
  static const int want = env_int("IIFB200_CLUSTER", IIF_SPEC_CLUSTER);
  return (want == IIF_SPEC_CLUSTER && grid * IIF_SPEC_CLUSTER <= ctx->num_sms) ? IIF_SPEC_CLUSTER : 1;
}
// Development knob: products at the very top of the tree (at most 18 per launch) can take a larger cluster, dealing the
// output samples — hence the Gibbs weight builds and picks — over up to eight SMs (the extra ranks ride along in the
// three-way speculative bandwidth search).  Measured: no gain (profiles/r02_summary.md: a lone product is bound by its
// sequence of per-level barriers, not by per-sample work), so the default stays at the three-CTA cluster.
static int pick_cluster_prod(iifb200_ctx* ctx, int grid) {
  static const int big = env_int("IIFB200_PROD_CLUSTER_BIG", 0);
  if (big > IIF_SPEC_CLUSTER && big <= 8 && grid * big <= ctx->num_sms && pick_cluster(ctx, grid) == IIF_SPEC_CLUSTER) return big;
  return pick_cluster(ctx, grid);
}
static bool is_prior_kind_h(int k) {
  return k == IIF_F_PRIOR || k == IIF_F_PRIOR_CIRCULAR || k == IIF_F_MSG_PRIOR || k == IIF_F_PARTIAL_PRIOR ||
         k == IIF_F_MANIFOLD_PRIOR || k == IIF_F_SO3_PRIOR;
}

static int32_t validate_conv(iifb200_ctx* ctx, const iif_conv_op& op) {
  if (op.factor < 0 || op.factor >= (int)ctx->factors.size()) return fail(ctx, IIF_ERR_ARG, "conv op: factor index out of range");
  const iif_factor_desc& F = ctx->factors[op.factor];
  if (op.sfidx < 1 || op.sfidx > F.arity) return fail(ctx, IIF_ERR_ARG, "conv op: sfidx out of range");
  if (op.N < 2 || op.N > IIF_MAX_POINTS) return fail(ctx, IIF_ERR_ARG, "conv op: N out of range");
  if (ctx->slots[F.slot[op.sfidx - 1]].cap < 1) return fail(ctx, IIF_ERR_ARG, "conv op: bad slot");
  return ensure_tree(ctx, op.N);
}

// ---- hot path: convolution batch -------------------------------------------------------------
int32_t iifb200_conv_batch(iifb200_ctx* ctx, int32_t K, const iif_conv_op* ops, const double* meas,
                           const int32_t* mhidx, const double* uinf, double* out_pts, double* out_bw,
                           double* out_ipc, int32_t* out_mhidx, int32_t* out_nan) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (K < 1 || !ops || !out_pts || !out_bw || !out_ipc) return fail(ctx, IIF_ERR_ARG, "conv_batch: bad arguments");
  CK(cudaSetDevice(ctx->device));
  std::vector<ConvTask> tasks(K);
  std::vector<int64_t> poff(K + 1, 0), noff(K + 1, 0);
  int64_t n_meas = 0, n_lab = 0, n_uinf = 0;
  size_t csmem = 0;
  int cmaxN = 2;
  for (int k = 0; k < K; ++k) {
    int32_t st = validate_conv(ctx, ops[k]);
    if (st != IIF_OK) return st;
    csmem = std::max(csmem, conv_smem_bytes(ops[k].N));
    cmaxN = std::max(cmaxN, (int)ops[k].N);
    const iif_factor_desc& F = ctx->factors[ops[k].factor];
    const int d = ctx->slots[F.slot[ops[k].sfidx - 1]].dim;
    poff[k + 1] = poff[k] + (int64_t)ops[k].N * d;
    noff[k + 1] = noff[k] + ops[k].N;
    if (ops[k].meas_off >= 0) n_meas = std::max<int64_t>(n_meas, ops[k].meas_off + (int64_t)ops[k].N * F.zdim);
    if (ops[k].mhidx_off >= 0) n_lab = std::max<int64_t>(n_lab, ops[k].mhidx_off + (int64_t)ops[k].N);
    if (ops[k].uinf_off >= 0) n_uinf = std::max<int64_t>(n_uinf, ops[k].uinf_off + (int64_t)(ctx->sp.inflateCycles + 1) * ops[k].N * d);
  }
  if ((n_meas && !meas) || (n_lab && !mhidx) || (n_uinf && !uinf)) return fail(ctx, IIF_ERR_ARG, "conv_batch: explicit stream offset given but array is NULL");
  double *d_pts = nullptr, *d_bw = nullptr, *d_meas = nullptr, *d_uinf = nullptr;
  int32_t *d_lab = nullptr, *d_misc = nullptr, *d_labin = nullptr;
  ConvTask* d_tasks = nullptr;
  auto cleanup = [&]() {};   // scratch comes from the grow-only pool
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return IIF_ERR_CUDA; } } while (0)
  {
    Carver cv;
    for (int pass = 0; pass < 2; ++pass) {
      cv.off = 0;
      d_pts = cv.take<double>(poff[K]);
      d_bw = cv.take<double>(2 * (size_t)K * IIF_MAX_DIM);
      d_lab = cv.take<int32_t>(noff[K]);
      d_misc = cv.take<int32_t>(2 * (size_t)K);
      d_tasks = cv.take<ConvTask>(K);
      d_meas = n_meas ? cv.take<double>(n_meas) : nullptr;
      d_labin = n_lab ? cv.take<int32_t>(n_lab) : nullptr;
      d_uinf = n_uinf ? cv.take<double>(n_uinf) : nullptr;
      if (pass == 0) CKC(pool_reserve(ctx, 0, cv.off + 256, &cv.base));
    }
  }
  CKC(cudaMemsetAsync(d_misc, 0, sizeof(int32_t) * 2 * K, ctx->stream));
  if (n_meas) CKC(cudaMemcpyAsync(d_meas, meas, sizeof(double) * n_meas, cudaMemcpyHostToDevice, ctx->stream));
  if (n_lab) CKC(cudaMemcpyAsync(d_labin, mhidx, sizeof(int32_t) * n_lab, cudaMemcpyHostToDevice, ctx->stream));
  if (n_uinf) CKC(cudaMemcpyAsync(d_uinf, uinf, sizeof(double) * n_uinf, cudaMemcpyHostToDevice, ctx->stream));
  for (int k = 0; k < K; ++k) {
    ConvTask& t = tasks[k];
    t.op = ops[k];
    t.out_pts = d_pts + poff[k];
    t.out_bw = d_bw + (int64_t)k * IIF_MAX_DIM;
    t.out_ipc = d_bw + (int64_t)(K + k) * IIF_MAX_DIM;
    t.out_mhidx = d_lab + noff[k];
    t.out_nan = d_misc + k;
    t.out_status = d_misc + K + k;
    t.out_slot = -1;
  }
  CKC(cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(ConvTask) * K, cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaEventRecord(ctx->ev0, ctx->stream));
  bool rare = false;
  for (int k = 0; k < K; ++k) rare |= conv_needs_rare(ctx->factors[ops[k].factor], ctx->slots.data(), ops[k].sfidx);
  launch_k(rare ? iif_conv_kernel_rare : iif_conv_kernel, K, pick_cluster(ctx, K), pick_threads(ctx, K, cmaxN), csmem, ctx->stream, ctx->dg, d_tasks, d_meas, d_labin, d_uinf, ctx->d_trees);
  CKC(cudaGetLastError());
  CKC(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += 1;
  std::vector<int32_t> misc(2 * K);
  CKC(cudaMemcpyAsync(out_pts, d_pts, sizeof(double) * poff[K], cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaMemcpyAsync(out_bw, d_bw, sizeof(double) * K * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaMemcpyAsync(out_ipc, d_bw + (int64_t)K * IIF_MAX_DIM, sizeof(double) * K * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_mhidx) CKC(cudaMemcpyAsync(out_mhidx, d_lab, sizeof(int32_t) * noff[K], cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaMemcpyAsync(misc.data(), d_misc, sizeof(int32_t) * 2 * K, cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  cleanup();
#undef CKC
  if (out_nan) for (int k = 0; k < K; ++k) out_nan[k] = misc[k];
  for (int k = 0; k < K; ++k)
    if (misc[K + k] != IIF_OK) {
      char b[160];
      snprintf(b, sizeof(b), "conv_batch: op %d failed on device: %s", k, status_name(misc[K + k]));
      return fail(ctx, misc[K + k], b);
    }
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_conv_batch: ") + e.what());
  }
}

// ---- hot path: product batch -------------------------------------------------------------------
int32_t iifb200_product_batch(iifb200_ctx* ctx, int32_t V, const iif_product_op* ops, const double* dens_pts,
                              const double* dens_bw, const int32_t* dens_mask, const double* old_pts,
                              const double* randU, const double* randN, double* out_pts, double* out_bw,
                              int32_t* out_labels) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  if (V < 1 || !ops || !dens_pts || !dens_bw || !out_pts || !out_bw) return fail(ctx, IIF_ERR_ARG, "product_batch: bad arguments");
  CK(cudaSetDevice(ctx->device));
  std::vector<int64_t> doff(V + 1, 0), foff(V + 1, 0), ooff(V + 1, 0), loff(V + 1, 0);
  int64_t n_u = 0, n_n = 0;
  size_t smem = 0;
  int pmaxN = 2;
  for (int v = 0; v < V; ++v) {
    const iif_product_op& o = ops[v];
    if (o.dim < 1 || o.dim > IIF_MAX_DIM || o.nfactors < 1 || o.nfactors > IIF_MAX_FACTORS)
      return fail(ctx, IIF_ERR_ARG, "product op: dim / nfactors out of range");
    int32_t st = ensure_tree(ctx, o.N);
    if (st != IIF_OK) return st;
    doff[v + 1] = doff[v] + (int64_t)o.nfactors * o.N * o.dim;
    foff[v + 1] = foff[v] + o.nfactors;
    ooff[v + 1] = ooff[v] + (int64_t)o.N * o.dim;
    loff[v + 1] = loff[v] + (int64_t)o.N * o.nfactors;
    const int L = ctx->trees[o.N].L;
    if (o.randu_off >= 0) n_u = std::max<int64_t>(n_u, o.randu_off + (int64_t)o.N * o.nfactors * (1 + (int64_t)L * (ctx->sp.gibbsNiter + 1)));
    if (o.randn_off >= 0) n_n = std::max<int64_t>(n_n, o.randn_off + (int64_t)o.N * o.dim * (L + 1));
    smem = std::max(smem, prod_smem_bytes(o.nfactors, o.N, o.dim, ctx->trees[o.N].nn, ctx->trees[o.N].L));
    pmaxN = std::max(pmaxN, (int)o.N);
  }
  if ((n_u && !randU) || (n_n && !randN)) return fail(ctx, IIF_ERR_ARG, "product_batch: explicit stream offset given but array is NULL");
  if ((int)smem > ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE) return fail(ctx, IIF_ERR_ARG, "product op exceeds the shared-memory budget (F*N*d too large)");
  double *d_in = nullptr, *d_bw = nullptr, *d_old = nullptr, *d_u = nullptr, *d_n = nullptr, *d_out = nullptr, *d_obw = nullptr;
  int32_t *d_lab = nullptr, *d_st = nullptr;
  ProdTask* d_tasks = nullptr;
  auto cleanup = [&]() {};   // scratch comes from the grow-only pool
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return IIF_ERR_CUDA; } } while (0)
  {
    Carver cv;
    for (int pass = 0; pass < 2; ++pass) {
      cv.off = 0;
      d_in = cv.take<double>(doff[V]);
      d_bw = cv.take<double>((size_t)foff[V] * IIF_MAX_DIM);
      d_out = cv.take<double>(ooff[V]);
      d_obw = cv.take<double>((size_t)V * IIF_MAX_DIM);
      d_lab = cv.take<int32_t>(loff[V]);
      d_st = cv.take<int32_t>(V);
      d_tasks = cv.take<ProdTask>(V);
      d_old = old_pts ? cv.take<double>(ooff[V]) : nullptr;
      d_u = n_u ? cv.take<double>(n_u) : nullptr;
      d_n = n_n ? cv.take<double>(n_n) : nullptr;
      if (pass == 0) CKC(pool_reserve(ctx, 0, cv.off + 256, &cv.base));
    }
  }
  CKC(cudaMemcpyAsync(d_in, dens_pts, sizeof(double) * doff[V], cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaMemcpyAsync(d_bw, dens_bw, sizeof(double) * foff[V] * IIF_MAX_DIM, cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaMemsetAsync(d_st, 0, sizeof(int32_t) * V, ctx->stream));
  if (old_pts) CKC(cudaMemcpyAsync(d_old, old_pts, sizeof(double) * ooff[V], cudaMemcpyHostToDevice, ctx->stream));
  if (n_u) CKC(cudaMemcpyAsync(d_u, randU, sizeof(double) * n_u, cudaMemcpyHostToDevice, ctx->stream));
  if (n_n) CKC(cudaMemcpyAsync(d_n, randN, sizeof(double) * n_n, cudaMemcpyHostToDevice, ctx->stream));
  std::vector<ProdTask> tasks(V);
  for (int v = 0; v < V; ++v) {
    ProdTask& t = tasks[v];
    memset(&t, 0, sizeof(t));
    const iif_product_op& o = ops[v];
    t.dim = o.dim; t.circ_mask = o.circ_mask; t.F = o.nfactors; t.N = o.N;
    t.call_id = o.call_id; t.randu_off = o.randu_off; t.randn_off = o.randn_off;
    t.target_slot = -1; t.out_slot = -1;
    for (int j = 0; j < o.nfactors; ++j) t.mask[j] = dens_mask ? dens_mask[foff[v] + j] : 0;
    t.dens_pts = d_in + doff[v];
    t.dens_bw = d_bw + foff[v] * IIF_MAX_DIM;
    t.old_pts = d_old ? d_old + ooff[v] : nullptr;
    t.out_pts = d_out + ooff[v];
    t.out_bw = d_obw + (int64_t)v * IIF_MAX_DIM;
    t.out_labels = d_lab + loff[v];
    t.conv_status = nullptr;
    t.out_status = d_st + v;
  }
  CKC(cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(ProdTask) * V, cudaMemcpyHostToDevice, ctx->stream));
  CKC(cudaEventRecord(ctx->ev0, ctx->stream));
  launch_k(iif_product_kernel, V, pick_cluster_prod(ctx, V), pick_threads_prod(ctx, V, pmaxN), smem, ctx->stream, ctx->dg, d_tasks, d_u, d_n, ctx->d_trees);
  CKC(cudaGetLastError());
  CKC(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += 1;
  CKC(cudaMemcpyAsync(out_pts, d_out, sizeof(double) * ooff[V], cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaMemcpyAsync(out_bw, d_obw, sizeof(double) * V * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_labels) CKC(cudaMemcpyAsync(out_labels, d_lab, sizeof(int32_t) * loff[V], cudaMemcpyDeviceToHost, ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
  cleanup();
#undef CKC
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_product_batch: ") + e.what());
  }
}

// ---- KDE bandwidth of K point sets ---------------------------------------------------------------
int32_t iifb200_kde_bandwidth(iifb200_ctx* ctx, int32_t K, const int32_t* N, const int32_t* dim,
                              const int32_t* circ_mask, const double* pts, double* out_bw) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  if (K < 1 || !N || !dim || !circ_mask || !pts || !out_bw) return fail(ctx, IIF_ERR_ARG, "kde_bandwidth: bad arguments");
  CK(cudaSetDevice(ctx->device));
  std::vector<int64_t> off(K + 1, 0);
  size_t bsmem = 0;
  int bmaxN = 2;
  for (int k = 0; k < K; ++k) {
    if (dim[k] < 1 || dim[k] > IIF_MAX_DIM) return fail(ctx, IIF_ERR_ARG, "kde_bandwidth: dim out of range");
    int32_t st = ensure_tree(ctx, N[k]);
    if (st != IIF_OK) return st;
    bsmem = std::max(bsmem, conv_smem_bytes(N[k]));
    bmaxN = std::max(bmaxN, (int)N[k]);
    off[k + 1] = off[k] + (int64_t)N[k] * dim[k];
  }
  double *d_pts = nullptr, *d_bw = nullptr;
  BwTask* d_t = nullptr;
  {
    Carver cv;
    for (int pass = 0; pass < 2; ++pass) {
      cv.off = 0;
      d_pts = cv.take<double>(off[K]);
      d_bw = cv.take<double>((size_t)K * IIF_MAX_DIM);
      d_t = cv.take<BwTask>(K);
      if (pass == 0) CK(pool_reserve(ctx, 0, cv.off + 256, &cv.base));
    }
  }
  std::vector<BwTask> t(K);
  for (int k = 0; k < K; ++k) { t[k].pts = d_pts + off[k]; t[k].out_bw = d_bw + (int64_t)k * IIF_MAX_DIM; t[k].N = N[k]; t[k].dim = dim[k]; t[k].circ_mask = circ_mask[k]; t[k]._pad = 0; }
  CK(cudaMemcpyAsync(d_pts, pts, sizeof(double) * off[K], cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_t, t.data(), sizeof(BwTask) * K, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  launch_k(iif_bandwidth_kernel, K, pick_cluster(ctx, K), pick_threads(ctx, K, bmaxN), bsmem, ctx->stream, d_t, ctx->d_trees);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += 1;
  CK(cudaMemcpyAsync(out_bw, d_bw, sizeof(double) * K * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_kde_bandwidth: ") + e.what());
  }
}

// ---- point estimates of V device-resident beliefs (calcPPE, FGOSUtils.jl:237-278) ------------------------
int32_t iifb200_ppe_batch(iifb200_ctx* ctx, int32_t V, const int32_t* slots, double* out_mean, double* out_max) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (V < 1 || !slots || !out_mean || !out_max) return fail(ctx, IIF_ERR_ARG, "ppe_batch: bad arguments");
  for (int v = 0; v < V; ++v)
    if (slots[v] < 0 || slots[v] >= (int)ctx->slots.size()) return fail(ctx, IIF_ERR_ARG, "ppe_batch: slot out of range");
  CK(cudaSetDevice(ctx->device));
  double* d_out = nullptr;
  PpeTask* d_t = nullptr;
  {
    Carver cv;
    for (int pass = 0; pass < 2; ++pass) {
      cv.off = 0;
      d_out = cv.take<double>(2 * (size_t)V * IIF_MAX_DIM);
      d_t = cv.take<PpeTask>(V);
      if (pass == 0) CK(pool_reserve(ctx, 0, cv.off + 256, &cv.base));
    }
  }
  std::vector<PpeTask> t(V);
  for (int v = 0; v < V; ++v) {
    t[v].slot = slots[v]; t[v]._pad = 0;
    t[v].out_mean = d_out + (int64_t)v * IIF_MAX_DIM;
    t[v].out_max = d_out + (int64_t)(V + v) * IIF_MAX_DIM;
  }
  CK(cudaMemcpyAsync(d_t, t.data(), sizeof(PpeTask) * V, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  iif_ppe_kernel<<<V, IIF_PPE_THREADS, 0, ctx->stream>>>(ctx->dg, d_t);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += 1;
  CK(cudaMemcpyAsync(out_mean, d_out, sizeof(double) * V * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_max, d_out + (int64_t)V * IIF_MAX_DIM, sizeof(double) * V * IIF_MAX_DIM, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_ppe_batch: ") + e.what());
  }
}

// ---- approxDeconv of K factors on device-resident beliefs (DeconvUtils.jl:32-162) ----------------------------
int32_t iifb200_deconv_batch(iifb200_ctx* ctx, int32_t K, const int32_t* factors, const int32_t* N,
                             const int32_t* call_ids, double* out_pred, double* out_meas) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (K < 1 || !factors || !N || !call_ids || !out_pred || !out_meas) return fail(ctx, IIF_ERR_ARG, "deconv_batch: bad arguments");
  CK(cudaSetDevice(ctx->device));
  std::vector<int64_t> off(K + 1, 0);
  for (int k = 0; k < K; ++k) {
    if (factors[k] < 0 || factors[k] >= (int)ctx->factors.size()) return fail(ctx, IIF_ERR_ARG, "deconv_batch: factor index out of range");
    if (N[k] < 1 || N[k] > IIF_MAX_POINTS) return fail(ctx, IIF_ERR_ARG, "deconv_batch: N out of range");
    off[k + 1] = off[k] + (int64_t)N[k] * ctx->factors[factors[k]].zdim;
  }
  double* d_out = nullptr;
  int32_t* d_st = nullptr;
  DeconvTask* d_t = nullptr;
  {
    Carver cv;
    for (int pass = 0; pass < 2; ++pass) {
      cv.off = 0;
      d_out = cv.take<double>(2 * (size_t)off[K]);
      d_st = cv.take<int32_t>(K);
      d_t = cv.take<DeconvTask>(K);
      if (pass == 0) CK(pool_reserve(ctx, 0, cv.off + 256, &cv.base));
    }
  }
  std::vector<DeconvTask> t(K);
  for (int k = 0; k < K; ++k) {
    t[k].factor = factors[k]; t[k].N = N[k]; t[k].call_id = call_ids[k]; t[k]._pad = 0;
    t[k].out_pred = d_out + off[k];
    t[k].out_meas = d_out + off[K] + off[k];
    t[k].out_status = d_st + k;
  }
  CK(cudaMemcpyAsync(d_t, t.data(), sizeof(DeconvTask) * K, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  iif_deconv_kernel<<<K, 128, 0, ctx->stream>>>(ctx->dg, d_t);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += 1;
  std::vector<int32_t> st(K);
  CK(cudaMemcpyAsync(out_pred, d_out, sizeof(double) * off[K], cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_meas, d_out + off[K], sizeof(double) * off[K], cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(st.data(), d_st, sizeof(int32_t) * K, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < K; ++k)
    if (st[k] != IIF_OK) return fail(ctx, st[k], std::string("deconv_batch: device reported '") + status_name(st[k]) + "'");
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_deconv_batch: ") + e.what());
  }
}

// ---- mmd kernel-embedding distance of K pairs of point sets (SolverUtilities.jl:25-47 / AMP.mmd!) ---------------
int32_t iifb200_mmd(iifb200_ctx* ctx, int32_t K, const int32_t* na, const int32_t* nb, const int32_t* dim,
                    const int32_t* circ_mask, const double* a, const double* b, double bw, double* out) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  if (K < 1 || !na || !nb || !dim || !circ_mask || !a || !b || !out) return fail(ctx, IIF_ERR_ARG, "mmd: bad arguments");
  CK(cudaSetDevice(ctx->device));
  std::vector<int64_t> oa(K + 1, 0), ob(K + 1, 0);
  size_t smem = 0;
  for (int k = 0; k < K; ++k) {
    if (dim[k] < 1 || dim[k] > IIF_MAX_DIM || na[k] < 1 || nb[k] < 1) return fail(ctx, IIF_ERR_ARG, "mmd: bad sizes");
    oa[k + 1] = oa[k] + (int64_t)na[k] * dim[k];
    ob[k + 1] = ob[k] + (int64_t)nb[k] * dim[k];
    smem = std::max(smem, sizeof(double) * (size_t)(na[k] + nb[k]) * dim[k]);
  }
  if ((int)smem > ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE) return fail(ctx, IIF_ERR_ARG, "mmd: point sets exceed the shared-memory budget");
  double *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
  MmdTask* d_t = nullptr;
  {
    Carver cv;
    for (int pass = 0; pass < 2; ++pass) {
      cv.off = 0;
      d_a = cv.take<double>(oa[K]);
      d_b = cv.take<double>(ob[K]);
      d_o = cv.take<double>(K);
      d_t = cv.take<MmdTask>(K);
      if (pass == 0) CK(pool_reserve(ctx, 0, cv.off + 256, &cv.base));
    }
  }
  std::vector<MmdTask> t(K);
  for (int k = 0; k < K; ++k) {
    t[k].a = d_a + oa[k]; t[k].b = d_b + ob[k];
    t[k].na = na[k]; t[k].nb = nb[k]; t[k].dim = dim[k]; t[k].circ_mask = circ_mask[k];
    t[k].bw = bw; t[k].out = d_o + k;
  }
  CK(cudaMemcpyAsync(d_a, a, sizeof(double) * oa[K], cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_b, b, sizeof(double) * ob[K], cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_t, t.data(), sizeof(MmdTask) * K, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaFuncSetAttribute(iif_mmd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE));
  iif_mmd_kernel<<<K, 256, smem, ctx->stream>>>(d_t);
  CK(cudaGetLastError());
  ctx->launches += 1;
  CK(cudaMemcpyAsync(out, d_o, sizeof(double) * K, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_mmd: ") + e.what());
  }
}

// ---- schedules: propagateBelief waves captured as a CUDA graph -----------------------------------
static int32_t build_schedule(iifb200_ctx* ctx, int32_t nwaves, const int32_t* wave_off, int32_t nops,
                              const iif_sched_op* ops, int32_t nprops, const iif_prop_op* props, int32_t ndeconvs,
                              const iif_deconv_op* deconvs, Schedule** out, bool pooled = false, int32_t nxfers = 0,
                              const iif_xfer_op* xfers = nullptr) {
  for (int k = 0; k < ndeconvs; ++k) {
    const iif_deconv_op& D = deconvs[k];
    if (D.factor < 0 || D.factor >= (int)ctx->factors.size() || D.out_slot < 0 || D.out_slot >= (int)ctx->slots.size())
      return fail(ctx, IIF_ERR_ARG, "deconv op: factor / slot out of range");
    const iif_factor_desc& F = ctx->factors[D.factor];
    if (F.arity != 2 || F.nmh != 0 || is_prior_kind_h(F.kind))
      return fail(ctx, IIF_ERR_UNSUPPORTED, "deconv op: needs a binary relative factor without multihypo");
    if (ctx->slots[D.out_slot].dim != F.zdim || ctx->slots[D.out_slot].cap < D.N)
      return fail(ctx, IIF_ERR_ARG, "deconv op: out slot must hold N points of the factor's measurement dimension");
    int32_t st = ensure_tree(ctx, D.N);
    if (st != IIF_OK) return st;
  }
  // per-prop scratch layout: proposals F*N*d doubles, then bw F*4, ipc F*4 (ipc unused downstream)
  std::vector<int64_t> soff(nprops + 1, 0);
  std::vector<int> cidx(nprops + 1, 0);
  for (int p = 0; p < nprops; ++p) {
    const iif_prop_op& P = props[p];
    if (P.nfactors < 1 || P.nfactors > IIF_MAX_FACTORS) return fail(ctx, IIF_ERR_ARG, "prop op: nfactors out of range");
    if (P.target_slot < 0 || P.target_slot >= (int)ctx->slots.size() || P.out_slot < 0 || P.out_slot >= (int)ctx->slots.size())
      return fail(ctx, IIF_ERR_ARG, "prop op: slot out of range");
    const iif_slot_desc& S = ctx->slots[P.target_slot];
    const iif_slot_desc& O = ctx->slots[P.out_slot];
    if (O.dim != S.dim || O.cap < P.N) return fail(ctx, IIF_ERR_ARG, "prop op: out slot incompatible with target");
    for (int f = 0; f < P.nfactors; ++f) {
      iif_conv_op c{};
      c.factor = P.factor[f]; c.sfidx = P.sfidx[f]; c.N = P.N;
      int32_t st = validate_conv(ctx, c);
      if (st != IIF_OK) return st;
      if (ctx->factors[c.factor].slot[c.sfidx - 1] != P.target_slot)
        return fail(ctx, IIF_ERR_ARG, "prop op: factor's solve-for variable is not the target slot");
    }
    soff[p + 1] = soff[p] + (int64_t)P.nfactors * P.N * S.dim + 2 * (int64_t)P.nfactors * IIF_MAX_DIM;
    cidx[p + 1] = cidx[p] + P.nfactors;
  }
  Schedule* s = new Schedule();
  *out = s;
  s->nconv = cidx[nprops];
  s->nprod = nprops;
  s->ndcv = ndeconvs;
  std::vector<ConvTask> ct;
  std::vector<ProdTask> pt;
  std::vector<int32_t> cp;
  std::vector<DeconvSlotTask> dt;
  std::vector<PushTask> pu;
  std::vector<int32_t> wa;
  std::vector<char> dused(std::max(ndeconvs, 1), 0);
  s->pooled = pooled;
  if (pooled) {   // one-shot schedule (propagate_batch): every device array comes from the ctx pool
    Carver cv;
    for (int pass = 0; pass < 2; ++pass) {
      cv.off = 0;
      s->d_scratch = cv.take<double>(std::max<int64_t>(soff[nprops], 1));
      s->d_status = cv.take<int32_t>(std::max(s->nstatus(), 1));
      s->d_conv = cv.take<ConvTask>(std::max(s->nconv, 1));
      s->d_prod = cv.take<ProdTask>(std::max(nprops, 1));
      s->d_copy = cv.take<int32_t>(2 * (size_t)std::max(nops, 1));
      s->d_dcv = cv.take<DeconvSlotTask>(std::max(ndeconvs, 1));
      if (pass == 0) CK(pool_reserve(ctx, 1, cv.off + 256, &cv.base));
    }
  } else {
    CK(cudaMalloc(&s->d_scratch, sizeof(double) * std::max<int64_t>(soff[nprops], 1)));
    CK(cudaMalloc(&s->d_status, sizeof(int32_t) * std::max(s->nstatus(), 1)));
  }
  CK(cudaMemsetAsync(s->d_status, 0, sizeof(int32_t) * std::max(s->nstatus(), 1), ctx->stream));
  std::vector<char> used(nprops, 0);
  for (int w = 0; w < nwaves; ++w) {
    Wave W;
    if (wave_off[w] < 0 || wave_off[w + 1] > nops || wave_off[w] > wave_off[w + 1]) return fail(ctx, IIF_ERR_ARG, "schedule: bad wave offsets");
    // lanes present in this wave; any lane-0 op makes the wave a barrier (single segment)
    std::vector<int> lanes;
    bool barrier = false;
    for (int k = wave_off[w]; k < wave_off[w + 1]; ++k) {
      const int ln = ops[k].lane;
      if (ln < 0 || ln > IIF_MAX_LANES) return fail(ctx, IIF_ERR_ARG, "schedule: lane out of range");
      if (ln == 0) barrier = true;
      if (std::find(lanes.begin(), lanes.end(), ln) == lanes.end()) lanes.push_back(ln);
    }
    if (barrier || lanes.empty()) lanes.assign(1, 0);
    std::sort(lanes.begin(), lanes.end());
   for (size_t li = 0; li < lanes.size(); ++li) {
    Seg G;
    G.lane = lanes[li];
    G.conv0 = (int)ct.size(); G.prod0 = (int)pt.size(); G.copy0 = (int)cp.size() / 2; G.dcv0 = (int)dt.size();
    G.push0 = (int)pu.size(); G.wait0 = (int)wa.size();
    for (int k = wave_off[w]; k < wave_off[w + 1]; ++k) {
      const iif_sched_op& o = ops[k];
      if (!barrier && o.lane != G.lane) continue;
      if (o.kind == IIF_S_COPY) {
        if (o.a < 0 || o.a >= (int)ctx->slots.size() || o.b < 0 || o.b >= (int)ctx->slots.size()) return fail(ctx, IIF_ERR_ARG, "schedule: copy slot out of range");
        if (ctx->slots[o.a].dim != ctx->slots[o.b].dim || ctx->slots[o.b].cap < ctx->slots[o.a].cap) return fail(ctx, IIF_ERR_ARG, "schedule: copy slots incompatible");
        cp.push_back(o.a); cp.push_back(o.b);
      } else if (o.kind == IIF_S_PUSH || o.kind == IIF_S_WAIT) {
        if (o.a < 0 || o.a >= nxfers || !xfers) return fail(ctx, IIF_ERR_ARG, "schedule: transfer index out of range");
        const iif_xfer_op& X = xfers[o.a];
        if (X.msg < 0 || X.msg >= ctx->nflags) return fail(ctx, IIF_ERR_ARG, "schedule: message id exceeds the exported flags (iifb200_ipc_export)");
        if (o.kind == IIF_S_WAIT) { wa.push_back(X.msg); continue; }
        if (X.peer < 0 || X.peer >= ctx->world || X.peer == ctx->rank || (int)ctx->peer_arena.size() != ctx->world)
          return fail(ctx, IIF_ERR_STATE, "schedule: PUSH to a peer that is not attached (iifb200_ipc_attach)");
        if (X.slot < 0 || X.slot >= (int)ctx->slots.size()) return fail(ctx, IIF_ERR_ARG, "schedule: transfer slot out of range");
        const int64_t ns = (int64_t)ctx->slots.size();
        double* rb = (double*)ctx->peer_arena[X.peer];
        PushTask t;
        t.slot = X.slot; t._pad = 0;
        t.r_pts = rb + ctx->slots[X.slot].pts_off;
        t.r_bw = rb + ctx->total_doubles + (int64_t)X.slot * IIF_MAX_DIM;
        t.r_ipc = rb + ctx->total_doubles + ns * IIF_MAX_DIM + (int64_t)X.slot * IIF_MAX_DIM;
        t.r_npts = (int32_t*)(rb + ctx->total_doubles + 2 * ns * IIF_MAX_DIM) + X.slot;
        t.r_flags = t.r_npts + ns;
        t.r_msgflag = ctx->peer_flags[X.peer] + X.msg;
        pu.push_back(t);
      } else if (o.kind == IIF_S_DECONV) {
        if (o.a < 0 || o.a >= ndeconvs) return fail(ctx, IIF_ERR_ARG, "schedule: deconv index out of range");
        if (dused[o.a]) return fail(ctx, IIF_ERR_ARG, "schedule: a deconv op may appear once");
        dused[o.a] = 1;
        DeconvSlotTask t;
        t.op = deconvs[o.a];
        t.out_status = s->d_status + s->nconv + s->nprod + o.a;
        dt.push_back(t);
        W.dcv_smem = std::max(W.dcv_smem, conv_smem_bytes(t.op.N));
        W.dcv_maxN = std::max(W.dcv_maxN, (int)t.op.N);
      } else if (o.kind == IIF_S_PROPAGATE) {
        if (o.a < 0 || o.a >= nprops) return fail(ctx, IIF_ERR_ARG, "schedule: prop index out of range");
        if (used[o.a]) return fail(ctx, IIF_ERR_ARG, "schedule: a prop op may appear once (its Philox call ids are unique)");
        used[o.a] = 1;
        const iif_prop_op& P = props[o.a];
        const iif_slot_desc& S = ctx->slots[P.target_slot];
        double* base = s->d_scratch + soff[o.a];
        double* bwbase = base + (int64_t)P.nfactors * P.N * S.dim;
        ProdTask t;
        memset(&t, 0, sizeof(t));
        for (int f = 0; f < P.nfactors; ++f) {
          const iif_factor_desc& F = ctx->factors[P.factor[f]];
          ConvTask c;
          memset(&c, 0, sizeof(c));
          c.op.factor = P.factor[f]; c.op.sfidx = P.sfidx[f]; c.op.N = P.N;
          c.op.call_id = P.call_id + 1 + f;
          // proposalbeliefs!: relative non-multihypo siblings of a multihypo factor (ApproxConv.jl:256-265)
          c.op.nullSurplus = (P.any_multihypo && !is_prior_kind_h(F.kind) && !F.nmh) ? ctx->sp.nullSurplusAdd : 0.0;
          c.op.meas_off = c.op.mhidx_off = c.op.uinf_off = -1;
          c.out_pts = base + (int64_t)f * P.N * S.dim;
          c.out_bw = bwbase + (int64_t)f * IIF_MAX_DIM;
          c.out_ipc = bwbase + (int64_t)(P.nfactors + f) * IIF_MAX_DIM;
          c.out_mhidx = nullptr; c.out_nan = nullptr;
          c.out_status = s->d_status + cidx[o.a] + f;
          // one full-dimension factor: the convolution writes the posterior itself, no product task
          c.out_slot = (P.nfactors == 1 && F.partial_mask == 0) ? P.out_slot : -1;
          W.conv_rare |= conv_needs_rare(F, ctx->slots.data(), c.op.sfidx);
          ct.push_back(c);
          t.mask[f] = F.partial_mask;
        }
        W.conv_smem = std::max(W.conv_smem, conv_smem_bytes(P.N));
        W.maxN = std::max(W.maxN, P.N);
        if (P.nfactors == 1 && ctx->factors[P.factor[0]].partial_mask == 0) continue;
        t.dim = S.dim; t.circ_mask = S.circ_mask; t.F = P.nfactors; t.N = P.N;
        t.call_id = P.call_id; t.randu_off = t.randn_off = -1;
        t.target_slot = P.target_slot; t.out_slot = P.out_slot;
        t.dens_pts = base; t.dens_bw = bwbase; t.old_pts = nullptr;
        t.out_pts = nullptr; t.out_bw = nullptr; t.out_labels = nullptr;
        t.conv_status = s->d_status + cidx[o.a];
        t.out_status = s->d_status + s->nconv + o.a;
        pt.push_back(t);
        W.prod_smem = std::max(W.prod_smem, prod_smem_bytes(P.nfactors, P.N, S.dim, ctx->trees[P.N].nn, ctx->trees[P.N].L));
      } else return fail(ctx, IIF_ERR_ARG, "schedule: unknown op kind");
    }
    G.nconv = (int)ct.size() - G.conv0; G.nprod = (int)pt.size() - G.prod0; G.ncopy = (int)cp.size() / 2 - G.copy0;
    G.ndcv = (int)dt.size() - G.dcv0;
    G.npush = (int)pu.size() - G.push0; G.nwait = (int)wa.size() - G.wait0;
    W.nconv += G.nconv; W.nprod += G.nprod; W.ncopy += G.ncopy; W.ndcv += G.ndcv;
    W.segs.push_back(G);
   }
    if ((int)W.prod_smem > ctx->max_smem_optin - IIF_STATIC_SMEM_RESERVE) return fail(ctx, IIF_ERR_ARG, "schedule: product exceeds the shared-memory budget");
    s->waves.push_back(W);
  }
  if (!pooled) {
    CK(cudaMalloc(&s->d_conv, sizeof(ConvTask) * std::max<size_t>(ct.size(), 1)));
    CK(cudaMalloc(&s->d_prod, sizeof(ProdTask) * std::max<size_t>(pt.size(), 1)));
    CK(cudaMalloc(&s->d_copy, sizeof(int32_t) * std::max<size_t>(cp.size(), 2)));
    CK(cudaMalloc(&s->d_dcv, sizeof(DeconvSlotTask) * std::max<size_t>(dt.size(), 1)));
  }
  if (!ct.empty()) CK(cudaMemcpyAsync(s->d_conv, ct.data(), sizeof(ConvTask) * ct.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!pt.empty()) CK(cudaMemcpyAsync(s->d_prod, pt.data(), sizeof(ProdTask) * pt.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!cp.empty()) CK(cudaMemcpyAsync(s->d_copy, cp.data(), sizeof(int32_t) * cp.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!dt.empty()) CK(cudaMemcpyAsync(s->d_dcv, dt.data(), sizeof(DeconvSlotTask) * dt.size(), cudaMemcpyHostToDevice, ctx->stream));
  if (!pu.empty()) {
    CK(cudaMalloc(&s->d_push, sizeof(PushTask) * pu.size()));
    CK(cudaMemcpyAsync(s->d_push, pu.data(), sizeof(PushTask) * pu.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!wa.empty()) {
    CK(cudaMalloc(&s->d_wait, sizeof(int32_t) * wa.size()));
    CK(cudaMemcpyAsync(s->d_wait, wa.data(), sizeof(int32_t) * wa.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  s->dist = !pu.empty() || !wa.empty();
  CK(cudaStreamSynchronize(ctx->stream));
  return IIF_OK;
}

int32_t iifb200_schedule_build(iifb200_ctx* ctx, int32_t nwaves, const int32_t* wave_off, int32_t nops,
                               const iif_sched_op* ops, int32_t nprops, const iif_prop_op* props,
                               int32_t* schedule_id_out) {
  try {   // nothing but status codes crosses the C-ABI
  return iifb200_schedule_build_ex(ctx, nwaves, wave_off, nops, ops, nprops, props, 0, nullptr, schedule_id_out);
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_schedule_build: ") + e.what());
  }
}

int32_t iifb200_schedule_build_ex(iifb200_ctx* ctx, int32_t nwaves, const int32_t* wave_off, int32_t nops,
                                  const iif_sched_op* ops, int32_t nprops, const iif_prop_op* props, int32_t ndeconvs,
                                  const iif_deconv_op* deconvs, int32_t* schedule_id_out) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (nwaves < 1 || !wave_off || nops < 0 || !ops || nprops < 0 || ndeconvs < 0 || (ndeconvs > 0 && !deconvs) || !schedule_id_out)
    return fail(ctx, IIF_ERR_ARG, "schedule_build: bad arguments");
  CK(cudaSetDevice(ctx->device));
  Schedule* s = nullptr;
  int32_t st = build_schedule(ctx, nwaves, wave_off, nops, ops, nprops, props, ndeconvs, deconvs, &s);
  if (st != IIF_OK) { free_schedule(s); return st; }
  ctx->schedules.push_back(s);
  *schedule_id_out = (int32_t)ctx->schedules.size() - 1;
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_schedule_build_ex: ") + e.what());
  }
}

// launches of one segment of a wave on stream `st`; CTA size and cluster choice follow the whole wave's width
static int enqueue_seg(iifb200_ctx* ctx, Schedule* s, const Wave& W, const Seg& G, cudaStream_t st) {
  int k = 0;
  if (G.nwait) {
    iif_wait_kernel<<<(G.nwait + 127) / 128, 128, 0, st>>>(ctx->d_flags, s->d_wait + G.wait0, G.nwait, ctx->d_epoch);
    ++k;
  }
  if (G.ncopy) {
    iif_copy_kernel<<<G.ncopy, 128, 0, st>>>(ctx->dg, s->d_copy + 2 * G.copy0, G.ncopy);
    ++k;
  }
  if (G.ndcv) {
    launch_k(iif_deconv_slot_kernel, G.ndcv, pick_cluster(ctx, W.ndcv), pick_threads(ctx, W.ndcv, W.dcv_maxN), W.dcv_smem, st, ctx->dg, s->d_dcv + G.dcv0, ctx->d_trees);
    ++k;
  }
  if (G.nconv) {
    launch_k(W.conv_rare ? iif_conv_kernel_rare : iif_conv_kernel, G.nconv, pick_cluster(ctx, W.nconv), pick_threads(ctx, W.nconv, W.maxN), W.conv_smem, st, ctx->dg, s->d_conv + G.conv0, nullptr, nullptr, nullptr, ctx->d_trees);
    ++k;
  }
  if (G.nprod) {
    launch_k(iif_product_kernel, G.nprod, pick_cluster_prod(ctx, W.nprod), pick_threads_prod(ctx, W.nprod, W.maxN), W.prod_smem, st, ctx->dg, s->d_prod + G.prod0, nullptr, nullptr, ctx->d_trees);
    ++k;
  }
  if (G.npush) {
    iif_push_kernel<<<G.npush, 128, 0, st>>>(ctx->dg, s->d_push + G.push0, G.npush, ctx->d_epoch);
    ++k;
  }
  return k;
}

// enqueue waves [w0, w1); returns the number of kernels launched.  `lanes` (only inside a stream capture): lane
// segments go to per-lane streams forked from / joined into the ctx stream with events, so the captured graph
// carries the lanes as parallel branches.  Without `lanes` everything is issued on the ctx stream in wave order.
static int32_t enqueue_waves(iifb200_ctx* ctx, Schedule* s, int w0, int w1, int* nk, bool lanes,
                             const std::vector<cudaEvent_t>* pool = nullptr) {
  int k = 0;
  size_t used = 0;
  bool active[IIF_MAX_LANES + 1] = {};
  auto new_event = [&](cudaEvent_t* e) {  // events come from a pool created BEFORE the capture began
    if (!pool || used >= pool->size()) return cudaErrorInvalidValue;
    *e = (*pool)[used++];
    return cudaSuccess;
  };
  auto join_all = [&]() -> cudaError_t {
    for (int l = 1; l <= IIF_MAX_LANES; ++l) {
      if (!active[l]) continue;
      cudaEvent_t e;
      cudaError_t err = new_event(&e);
      if (err == cudaSuccess) err = cudaEventRecord(e, ctx->lane_stream[l]);
      if (err == cudaSuccess) err = cudaStreamWaitEvent(ctx->stream, e, 0);
      if (err != cudaSuccess) return err;
      active[l] = false;
    }
    return cudaSuccess;
  };
  for (int w = w0; w < w1; ++w) {
    const Wave& W = s->waves[w];
    for (const Seg& G : W.segs) {
      if (!lanes || G.lane == 0) {
        if (lanes) CK(join_all());
        k += enqueue_seg(ctx, s, W, G, ctx->stream);
        continue;
      }
      if (!active[G.lane]) {  // fork: the lane continues from the ctx stream's current point
        cudaEvent_t e;
        CK(new_event(&e));
        CK(cudaEventRecord(e, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->lane_stream[G.lane], e, 0));
        active[G.lane] = true;
      }
      k += enqueue_seg(ctx, s, W, G, ctx->lane_stream[G.lane]);
    }
  }
  if (lanes) CK(join_all());
  CK(cudaGetLastError());
  *nk = k;
  return IIF_OK;
}

int32_t iifb200_schedule_run(iifb200_ctx* ctx, int32_t schedule_id, int32_t first_wave, int32_t last_wave) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (schedule_id < 0 || schedule_id >= (int)ctx->schedules.size() || !ctx->schedules[schedule_id]) return fail(ctx, IIF_ERR_ARG, "schedule_run: bad schedule id");
  Schedule* s = ctx->schedules[schedule_id];
  const int nw = (int)s->waves.size();
  if (first_wave < 0) first_wave = 0;
  if (last_wave < 0 || last_wave > nw) last_wave = nw;
  if (first_wave >= last_wave) return IIF_OK;
  CK(cudaSetDevice(ctx->device));
  auto key = std::make_pair(first_wave, last_wave);
  auto it = s->graphs.find(key);
  if (it == s->graphs.end()) {
    cudaGraph_t graph = nullptr;
    int nk = 0;
    // lane streams and the fork / join events are created before the capture begins (no resource creation inside it)
    std::vector<cudaEvent_t> pool;
    {
      size_t need = 0;
      for (int w = first_wave; w < last_wave; ++w)
        for (const Seg& G : s->waves[w].segs) {
          if (G.lane == 0) continue;
          if (!ctx->lane_stream[G.lane]) CK(cudaStreamCreateWithFlags(&ctx->lane_stream[G.lane], cudaStreamNonBlocking));
          need += 2;  // at most one fork and one join per lane segment
        }
      pool.resize(need);
      for (auto& e : pool) {
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->events.push_back(e);
      }
    }
    if (s->dist && (first_wave != 0 || last_wave != nw)) return fail(ctx, IIF_ERR_ARG, "schedule_run: a schedule with peer messages must be run whole");
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    if (s->dist) iif_epoch_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_epoch);   // this replay's flag value
    int32_t st = enqueue_waves(ctx, s, first_wave, last_wave, &nk, true, &pool);
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (st != IIF_OK) { if (graph) cudaGraphDestroy(graph); return st; }
    if (e != cudaSuccess) return fail(ctx, IIF_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(ctx, IIF_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    it = s->graphs.emplace(key, std::make_pair(exec, nk)).first;
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(cudaGraphLaunch(it->second.first, ctx->stream));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  ctx->launches += it->second.second;
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_schedule_run: ") + e.what());
  }
}

int32_t iifb200_schedule_profile(iifb200_ctx* ctx, int32_t schedule_id, int32_t first_wave, int32_t last_wave,
                                 float* ms, int32_t* launches, int64_t* blocks) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (schedule_id < 0 || schedule_id >= (int)ctx->schedules.size() || !ctx->schedules[schedule_id] || !ms || !launches || !blocks)
    return fail(ctx, IIF_ERR_ARG, "schedule_profile: bad arguments");
  Schedule* s = ctx->schedules[schedule_id];
  const int nw = (int)s->waves.size();
  if (first_wave < 0) first_wave = 0;
  if (last_wave < 0 || last_wave > nw) last_wave = nw;
  CK(cudaSetDevice(ctx->device));
  for (int k = 0; k < 3; ++k) { ms[k] = 0.f; launches[k] = 0; blocks[k] = 0; }
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
  auto mark = [&]() { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, ctx->stream); ev.push_back(e); };
  for (int w = first_wave; w < last_wave; ++w) {
    const Wave& W = s->waves[w];
    if (W.ncopy) {
      mark();
      iif_copy_kernel<<<W.ncopy, 128, 0, ctx->stream>>>(ctx->dg, s->d_copy + 2 * W.segs[0].copy0, W.ncopy);
      mark(); kind.push_back(2); blocks[2] += W.ncopy;
    }
    if (W.ndcv) {  // differential-likelihood construction is accounted with the copy (message) kernels
      mark();
      launch_k(iif_deconv_slot_kernel, W.ndcv, pick_cluster(ctx, W.ndcv), pick_threads(ctx, W.ndcv, W.dcv_maxN), W.dcv_smem, ctx->stream, ctx->dg, s->d_dcv + W.segs[0].dcv0, ctx->d_trees);
      mark(); kind.push_back(2); blocks[2] += W.ndcv;
    }
    if (W.nconv) {
      mark();
      launch_k(W.conv_rare ? iif_conv_kernel_rare : iif_conv_kernel, W.nconv, pick_cluster(ctx, W.nconv), pick_threads(ctx, W.nconv, W.maxN), W.conv_smem, ctx->stream, ctx->dg, s->d_conv + W.segs[0].conv0, nullptr, nullptr, nullptr, ctx->d_trees);
      mark(); kind.push_back(0); blocks[0] += W.nconv;
    }
    if (W.nprod) {
      mark();
      launch_k(iif_product_kernel, W.nprod, pick_cluster_prod(ctx, W.nprod), pick_threads_prod(ctx, W.nprod, W.maxN), W.prod_smem, ctx->stream, ctx->dg, s->d_prod + W.segs[0].prod0, nullptr, nullptr, ctx->d_trees);
      mark(); kind.push_back(1); blocks[1] += W.nprod;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  for (size_t i = 0; i < kind.size(); ++i) {
    float t = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
    ms[kind[i]] += t;
    launches[kind[i]] += 1;
  }
  for (auto x : ev) cudaEventDestroy(x);
  ctx->launches += (int64_t)kind.size();
  if (e != cudaSuccess) return fail(ctx, IIF_ERR_CUDA, std::string("schedule_profile: ") + cudaGetErrorString(e));
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_schedule_profile: ") + e.what());
  }
}

int32_t iifb200_schedule_build_dist(iifb200_ctx* ctx, int32_t nwaves, const int32_t* wave_off, int32_t nops,
                                    const iif_sched_op* ops, int32_t nprops, const iif_prop_op* props, int32_t ndeconvs,
                                    const iif_deconv_op* deconvs, int32_t nxfers, const iif_xfer_op* xfers,
                                    int32_t* schedule_id_out) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (nwaves < 1 || !wave_off || nops < 0 || !ops || nprops < 0 || ndeconvs < 0 || (ndeconvs > 0 && !deconvs) || nxfers < 0 ||
      (nxfers > 0 && !xfers) || !schedule_id_out)
    return fail(ctx, IIF_ERR_ARG, "schedule_build_dist: bad arguments");
  CK(cudaSetDevice(ctx->device));
  Schedule* s = nullptr;
  int32_t st = build_schedule(ctx, nwaves, wave_off, nops, ops, nprops, props, ndeconvs, deconvs, &s, false, nxfers, xfers);
  if (st != IIF_OK) { free_schedule(s); return st; }
  ctx->schedules.push_back(s);
  *schedule_id_out = (int32_t)ctx->schedules.size() - 1;
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_schedule_build_dist: ") + e.what());
  }
}

int32_t iifb200_ipc_export(iifb200_ctx* ctx, int32_t nflags, void* arena_handle64, void* flags_handle64) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (nflags < 1 || !arena_handle64 || !flags_handle64) return fail(ctx, IIF_ERR_ARG, "ipc_export: bad arguments");
  if (!ctx->arena_owned) return fail(ctx, IIF_ERR_STATE, "ipc_export: needs a library-owned arena (set_graph with ext_arena == NULL)");
  CK(cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  if (ctx->d_flags && ctx->nflags < nflags) { cudaFree(ctx->d_flags); ctx->d_flags = nullptr; }
  if (!ctx->d_flags) {
    CK(cudaMalloc(&ctx->d_flags, sizeof(int32_t) * (size_t)nflags));
    ctx->nflags = nflags;
  }
  CK(cudaMemset(ctx->d_flags, 0, sizeof(int32_t) * (size_t)ctx->nflags));
  if (!ctx->d_epoch) CK(cudaMalloc(&ctx->d_epoch, sizeof(int32_t)));
  CK(cudaMemset(ctx->d_epoch, 0, sizeof(int32_t)));
  cudaIpcMemHandle_t ha, hf;
  CK(cudaIpcGetMemHandle(&ha, ctx->arena));
  CK(cudaIpcGetMemHandle(&hf, ctx->d_flags));
  memcpy(arena_handle64, &ha, 64);
  memcpy(flags_handle64, &hf, 64);
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_ipc_export: ") + e.what());
  }
}

int32_t iifb200_ipc_attach(iifb200_ctx* ctx, int32_t world, int32_t rank, const void* arena_handles, const void* flags_handles) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (world < 1 || rank < 0 || rank >= world || !arena_handles || !flags_handles) return fail(ctx, IIF_ERR_ARG, "ipc_attach: bad arguments");
  if (!ctx->d_flags) return fail(ctx, IIF_ERR_STATE, "ipc_attach: call iifb200_ipc_export first");
  CK(cudaSetDevice(ctx->device));
  ctx->world = world; ctx->rank = rank;
  ctx->peer_arena.assign(world, nullptr);
  ctx->peer_flags.assign(world, nullptr);
  for (int r = 0; r < world; ++r) {
    if (r == rank) { ctx->peer_arena[r] = ctx->arena; ctx->peer_flags[r] = ctx->d_flags; continue; }
    cudaIpcMemHandle_t ha, hf;
    memcpy(&ha, (const char*)arena_handles + 64 * r, 64);
    memcpy(&hf, (const char*)flags_handles + 64 * r, 64);
    void *pa = nullptr, *pf = nullptr;
    CK(cudaIpcOpenMemHandle(&pa, ha, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_arena[r] = pa;
    ctx->peer_flags[r] = (int32_t*)pf;
  }
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_ipc_attach: ") + e.what());
  }
}

int32_t iifb200_schedule_free(iifb200_ctx* ctx, int32_t schedule_id) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  if (schedule_id < 0 || schedule_id >= (int)ctx->schedules.size()) return fail(ctx, IIF_ERR_ARG, "schedule_free: bad id");
  CK(cudaStreamSynchronize(ctx->stream));
  free_schedule(ctx->schedules[schedule_id]);
  ctx->schedules[schedule_id] = nullptr;
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_schedule_free: ") + e.what());
  }
}

// V independent propagateBelief calls: a one-wave schedule, not cached
int32_t iifb200_propagate_batch(iifb200_ctx* ctx, int32_t V, const iif_prop_op* ops) {
  try {   // nothing but status codes crosses the C-ABI
  NEED_GRAPH();
  if (V < 1 || !ops) return fail(ctx, IIF_ERR_ARG, "propagate_batch: bad arguments");
  CK(cudaSetDevice(ctx->device));
  std::vector<iif_sched_op> so(V);
  for (int v = 0; v < V; ++v) { so[v].kind = IIF_S_PROPAGATE; so[v].a = v; so[v].b = 0; so[v].lane = 0; }
  int32_t wo[2] = {0, V};
  Schedule* s = nullptr;
  int32_t st = build_schedule(ctx, 1, wo, V, so.data(), V, ops, 0, nullptr, &s, true);
  if (st != IIF_OK) { free_schedule(s); return st; }
  int nk = 0;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  st = enqueue_waves(ctx, s, 0, 1, &nk, false);
  if (st == IIF_OK) {
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->timed = true;
    ctx->launches += nk;
    std::vector<int32_t> stat(s->nstatus());
    cudaError_t e = cudaMemcpyAsync(stat.data(), s->d_status, sizeof(int32_t) * stat.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free_schedule(s); return fail(ctx, IIF_ERR_CUDA, std::string("propagate_batch: ") + cudaGetErrorString(e)); }
    for (size_t i = 0; i < stat.size(); ++i)
      if (stat[i] != IIF_OK) { st = fail(ctx, stat[i], std::string("propagate_batch: device reported '") + status_name(stat[i]) + "'"); break; }
  }
  free_schedule(s);
  return st;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_propagate_batch: ") + e.what());
  }
}

int32_t iifb200_propagate_once(iifb200_ctx* ctx, int32_t nslots, iif_slot_desc* slots, int32_t nfactors,
                               const iif_factor_desc* factors, int32_t ndists, const iif_dist_desc* dists, int32_t nparams,
                               const double* dparams, const iif_solver_params* sp, const double* pts, const double* bw,
                               const int32_t* npts, const int32_t* flags, const iif_prop_op* op, int32_t* out_npts,
                               double* out_pts, double* out_bw, double* out_ipc) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  if (!op || !pts || !bw || !npts || !flags || !out_pts) return fail(ctx, IIF_ERR_ARG, "propagate_once: bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));   // earlier work of this context is complete (normally a no-op)
  ctx->defer_sync = true;
  int32_t st = iifb200_set_graph(ctx, nslots, slots, nfactors, factors, ndists, dists, nparams, dparams, sp, nullptr);
  ctx->defer_sync = false;
  if (st != IIF_OK) return st;
  st = iifb200_upload_slots(ctx, 0, nslots, pts, bw, npts, flags);
  if (st != IIF_OK) return st;
  if (op->out_slot < 0 || op->out_slot >= nslots) return fail(ctx, IIF_ERR_ARG, "propagate_once: out_slot out of range");
  iif_sched_op so;
  so.kind = IIF_S_PROPAGATE; so.a = 0; so.b = 0; so.lane = 0;
  int32_t wo[2] = {0, 1};
  Schedule* s = nullptr;
  st = build_schedule(ctx, 1, wo, 1, &so, 1, op, 0, nullptr, &s, true);
  if (st != IIF_OK) { free_schedule(s); return st; }
  int nk = 0;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  st = enqueue_waves(ctx, s, 0, 1, &nk, false);
  if (st == IIF_OK) {
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->timed = true;
    ctx->launches += nk;
    const iif_slot_desc& S = ctx->slots[op->out_slot];
    const int n = op->N;                      // a propagateBelief always leaves N points in its destination
    std::vector<int32_t> stat(s->nstatus());
    double b[IIF_MAX_DIM], q[IIF_MAX_DIM];
    cudaError_t e = cudaMemcpyAsync(stat.data(), s->d_status, sizeof(int32_t) * stat.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_pts, ctx->dg.pts + S.pts_off, sizeof(double) * n * S.dim, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b, ctx->dg.bw + (int64_t)op->out_slot * IIF_MAX_DIM, sizeof(b), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q, ctx->dg.ipc + (int64_t)op->out_slot * IIF_MAX_DIM, sizeof(q), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free_schedule(s); return fail(ctx, IIF_ERR_CUDA, std::string("propagate_once: ") + cudaGetErrorString(e)); }
    for (size_t i = 0; i < stat.size(); ++i)
      if (stat[i] != IIF_OK) { st = fail(ctx, stat[i], std::string("propagate_once: device reported '") + status_name(stat[i]) + "'"); break; }
    if (st == IIF_OK) {
      if (out_npts) *out_npts = n;
      if (out_bw) for (int c = 0; c < S.dim; ++c) out_bw[c] = b[c];
      if (out_ipc) for (int c = 0; c < S.dim; ++c) out_ipc[c] = q[c];
    }
  }
  free_schedule(s);
  return st;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_propagate_once: ") + e.what());
  }
}

int32_t iifb200_plan_upload(iifb200_ctx* ctx, const iifb200_plan* plan, const iif_solver_params* sp, void* ext_arena,
                            int32_t* schedule_id_out) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  if (!plan || !sp || !schedule_id_out) return fail(ctx, IIF_ERR_ARG, "plan_upload: bad arguments");
  std::vector<iif_slot_desc> slots = iif_plan_slots(plan);   // set_graph fills pts_off in place
  const auto& F = iif_plan_factors(plan);
  const auto& D = iif_plan_dists(plan);
  const auto& prm = iif_plan_dparams(plan);
  int32_t st = iifb200_set_graph(ctx, (int32_t)slots.size(), slots.data(), (int32_t)F.size(), F.data(), (int32_t)D.size(),
                                 D.data(), (int32_t)prm.size(), prm.data(), sp, ext_arena);
  if (st != IIF_OK) return st;
  const auto& wo = iif_plan_wave_off(plan);
  const auto& ops = iif_plan_ops(plan);
  const auto& props = iif_plan_props(plan);
  const auto& dcv = iif_plan_deconvs(plan);
  return iifb200_schedule_build_ex(ctx, (int32_t)wo.size() - 1, wo.data(), (int32_t)ops.size(), ops.data(),
                                   (int32_t)props.size(), props.data(), (int32_t)dcv.size(), dcv.empty() ? nullptr : dcv.data(),
                                   schedule_id_out);
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_plan_upload: ") + e.what());
  }
}

int32_t iifb200_sync(iifb200_ctx* ctx) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  // surface device-side failures of schedule runs
  for (auto* s : ctx->schedules) {
    if (!s) continue;
    std::vector<int32_t> stat(s->nstatus());
    if (stat.empty()) continue;
    CK(cudaMemcpy(stat.data(), s->d_status, sizeof(int32_t) * stat.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < stat.size(); ++i)
      if (stat[i] != IIF_OK) {
        cudaMemset(s->d_status, 0, sizeof(int32_t) * stat.size());
        return fail(ctx, stat[i], std::string("schedule: device reported '") + status_name(stat[i]) + "'");
      }
  }
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_sync: ") + e.what());
  }
}

int64_t iifb200_launch_count(const iifb200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int32_t iifb200_set_stream(iifb200_ctx* ctx, void* stream) {
  try {   // nothing but status codes crosses the C-ABI
  if (!ctx) return IIF_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
  return IIF_OK;
  } catch (const std::exception& e) {
    if (!ctx) return IIF_ERR_STATE;
    return fail(ctx, IIF_ERR_STATE, std::string("iifb200_set_stream: ") + e.what());
  }
}
void* iifb200_stream(iifb200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
#ifdef IIF_PHASES
// development only (profiles/phase_probe.py): read and optionally reset the phase clocks
void iifb200_debug_phases(long long* out32, int reset) {
  if (out32) cudaMemcpyFromSymbol(out32, g_iif_phase, sizeof(long long) * 32);
  if (reset) { long long z[32] = {0}; cudaMemcpyToSymbol(g_iif_phase, z, sizeof(z)); }
}
#endif
float iifb200_last_elapsed_ms(iifb200_ctx* ctx) {
  if (!ctx || !ctx->timed) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(ctx->ev1) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

}  // extern "C"
