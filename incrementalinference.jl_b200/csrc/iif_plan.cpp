// iif_plan.cpp — host-side planner of libiifb200.so: lowers ONE solveTree! pass (up + down clique solves) from the
// data a Julia caller already holds (the factor graph's descriptor tables and the Bayes tree `tree.bt` with every
// clique's frontals, separators, potentials and Gibbs variable classes) to belief slots, propagateBelief ops and
// waves of independent ops, including copy forwarding and lane assignment.  No CUDA here: the planner is a pure
// function of its inputs (unit-tested without a GPU); iifb200_plan_upload (iifb200.cu) hands the result to the device.
//
// Reference call sites this replaces on the caller's side (it does what the CSM would do clique by clique):
//   buildCliqSubgraph! deep copies                      CliqueStateMachine.jl step 0b
//   addMsgFactors! (one MsgPrior per separator)         TreeMessageUtils.jl:566-575
//   upGibbsCliqueDensity / fmcmc! sweeps                SolveTree.jl:89-239
//   updateSubFgFromDownMsgs!, addDownVariableFactors!   TreeMessageUtils.jl:66, CliqStateMachineUtils.jl:479-571
//   determineCliqVariableDownSequence / downGibbs       CliqStateMachineUtils.jl:438-571
//   updateFromSubgraph (frontals back to the graph)     CliqueStateMachine.jl:928-966
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/iifb200.h"

struct iifb200_plan {
  std::vector<iif_slot_desc> slots;
  std::vector<iif_factor_desc> factors;
  std::vector<iif_dist_desc> dists;
  std::vector<double> dparams;
  std::vector<iif_prop_op> props;
  std::vector<iif_deconv_op> deconvs; // IIF_S_DECONV ops (useMsgLikelihoods)
  std::vector<iif_sched_op> ops;      // sorted by wave
  std::vector<int32_t> wave_off;
  int32_t n_conv = 0, n_prod = 0, n_msgs = 0, up_last_wave = 0, nvars = 0;
};

static thread_local std::string g_plan_error;

namespace {

bool has(const std::vector<int>& v, int x);

struct Op {
  int kind, a, b, clique;
  std::vector<int> rd, wr;
  double weight;
};

struct Inst {  // one factor instance of a clique sub-graph
  std::vector<int> vars;
  std::vector<int> rd;
  int fi;
  bool mh;
  int type = -1;        // caller's factor-type id (joint-message rules compare types for equality only)
  bool prior = false;
  int tag = 0;          // 0 potential, 1 UPWARD_COMMON message prior, 2 differential
};
struct Relative {       // one differential of a joint up message (addLikelihoodsDifferentialCHILD!, TreeMessageUtils.jl:279-335)
  int v1, v2, kind, type, slot, zdim;
};
struct UpMsg {
  std::vector<Relative> relatives;
  std::vector<std::pair<int, int>> priors;   // (variable, source slot)
  bool hasPriors = false;
};

// findShortestPathDijkstra on the bipartite variable / factor graph of a clique sub-graph (unit weights => BFS, neighbours
// in insertion order): the factor types along the path; `found` false when the variables are not connected
std::vector<int> path_types(const std::vector<Inst>& inst, int src, int dst, int only_type, bool& found) {
  found = true;
  if (src == dst) return {};
  std::unordered_map<int, std::pair<int, int>> prev;   // variable -> (previous variable, instance)
  prev[src] = {-1, -1};
  std::vector<int> frontier{src};
  while (!frontier.empty()) {
    std::vector<int> nxt;
    for (int u : frontier)
      for (size_t k = 0; k < inst.size(); ++k) {
        const Inst& e = inst[k];
        if (!has(e.vars, u) || (only_type >= 0 && e.type != only_type)) continue;
        for (int w : e.vars)
          if (!prev.count(w)) { prev[w] = {u, (int)k}; nxt.push_back(w); }
      }
    if (prev.count(dst)) break;
    frontier = nxt;
  }
  if (!prev.count(dst)) { found = false; return {}; }
  std::vector<int> types;
  for (int w = dst; prev[w].first >= 0; w = prev[w].first) types.push_back(inst[prev[w].second].type);
  std::reverse(types.begin(), types.end());
  return types;
}

std::vector<int> csr(const int32_t* off, const int32_t* val, int i) {
  return std::vector<int>(val + off[i], val + off[i + 1]);
}
bool has(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// wave(op) = 1 + max wave of earlier ops it conflicts with (RAW / WAR / WAW on slots)
std::vector<int> levelize(const std::vector<Op>& ops) {
  std::unordered_map<int, int> last_w, last_r;
  std::vector<int> waves(ops.size());
  for (size_t i = 0; i < ops.size(); ++i) {
    int w = 0;
    for (int s : ops[i].rd) { auto it = last_w.find(s); if (it != last_w.end()) w = std::max(w, it->second + 1); }
    for (int s : ops[i].wr) {
      auto it = last_w.find(s); if (it != last_w.end()) w = std::max(w, it->second + 1);
      auto ir = last_r.find(s); if (ir != last_r.end()) w = std::max(w, ir->second + 1);
    }
    waves[i] = w;
    for (int s : ops[i].rd) { auto ir = last_r.find(s); last_r[s] = (ir == last_r.end()) ? w : std::max(ir->second, w); }
    for (int s : ops[i].wr) last_w[s] = w;
  }
  return waves;
}

// Lanes: groups of disjoint sub-trees (independent until their common ancestor); the tree top stays lane 0 and any
// op that conflicts with another lane without a barrier wave in between is demoted to lane 0.
std::vector<int> assign_lanes(int ncl, const int32_t* parent, const std::vector<std::vector<int>>& children,
                              const std::vector<Op>& ops, const std::vector<int>& waves, int nlanes) {
  const int n = (int)ops.size();
  std::vector<int> lane(n, 0);
  if (nlanes < 2 || n == 0) return lane;
  std::vector<double> sub(ncl, 0.0);
  for (int i = 0; i < n; ++i) if (ops[i].clique >= 0) sub[ops[i].clique] += ops[i].weight;
  for (int c = ncl - 1; c >= 0; --c) if (parent[c] >= 0) sub[parent[c]] += sub[c];   // children have larger ids
  std::vector<int> frontier;
  for (int c = 0; c < ncl; ++c) if (parent[c] < 0) frontier.push_back(c);
  while ((int)frontier.size() < 2 * nlanes) {
    int big = -1;
    for (int c : frontier) if (!children[c].empty() && (big < 0 || sub[c] > sub[big])) big = c;   // first maximum
    if (big < 0) break;
    frontier.erase(std::find(frontier.begin(), frontier.end(), big));
    for (int ch : children[big]) frontier.push_back(ch);
  }
  std::vector<double> load(nlanes + 1, 0.0);
  std::vector<int> lane_of_clique(ncl, 0);
  std::vector<int> order_f = frontier;
  std::stable_sort(order_f.begin(), order_f.end(), [&](int a, int b) { return sub[a] > sub[b]; });
  for (int r : order_f) {
    int ln = 1;
    for (int k = 2; k <= nlanes; ++k) if (load[k] < load[ln]) ln = k;
    load[ln] += sub[r];
    std::vector<int> stack{r};
    while (!stack.empty()) {
      int c = stack.back(); stack.pop_back();
      lane_of_clique[c] = ln;
      for (int ch : children[c]) stack.push_back(ch);
    }
  }
  for (int i = 0; i < n; ++i) lane[i] = ops[i].clique >= 0 ? lane_of_clique[ops[i].clique] : 0;
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return waves[a] < waves[b]; });
  for (int pass = 0; pass < n; ++pass) {   // demotions create new barrier waves: iterate to a fixed point
    std::vector<int> barrier;
    for (int i = 0; i < n; ++i) if (lane[i] == 0) barrier.push_back(waves[i]);
    std::sort(barrier.begin(), barrier.end());
    barrier.erase(std::unique(barrier.begin(), barrier.end()), barrier.end());
    auto separated = [&](int wa, int wb) {   // a barrier wave in [wa, wb]
      auto it = std::lower_bound(barrier.begin(), barrier.end(), wa);
      return it != barrier.end() && *it <= wb;
    };
    std::unordered_map<int, std::pair<int, int>> last_w;
    std::unordered_map<int, std::vector<std::pair<int, int>>> last_r;
    bool changed = false;
    for (int i : order) {
      const int li = lane[i], wi = waves[i];
      bool conflict = false;
      if (li != 0) {
        for (int s : ops[i].rd) {
          auto a = last_w.find(s);
          if (a != last_w.end() && a->second.first != li && a->second.first != 0 && !separated(a->second.second, wi)) conflict = true;
        }
        for (int s : ops[i].wr) {
          auto a = last_w.find(s);
          if (a != last_w.end() && a->second.first != li && a->second.first != 0 && !separated(a->second.second, wi)) conflict = true;
          auto r = last_r.find(s);
          if (r != last_r.end())
            for (auto& lw : r->second)
              if (lw.first != li && lw.first != 0 && !separated(lw.second, wi)) conflict = true;
        }
      }
      if (conflict) { lane[i] = 0; changed = true; break; }
      for (int s : ops[i].wr) { last_w[s] = {li, wi}; last_r[s].clear(); }
      for (int s : ops[i].rd) last_r[s].push_back({li, wi});
    }
    if (!changed) break;
  }
  return lane;
}

}  // namespace

extern "C" {

const char* iifb200_plan_error(void) { return g_plan_error.c_str(); }

void iifb200_plan_free(iifb200_plan* p) { delete p; }

static int32_t plan_tree_impl(const iif_graph_desc* g, const iif_tree_desc* t, const iif_plan_opts* o, iifb200_plan** out);
int32_t iifb200_plan_tree(const iif_graph_desc* g, const iif_tree_desc* t, const iif_plan_opts* o, iifb200_plan** out) {
  try {   // nothing crosses the C-ABI but status codes
    return plan_tree_impl(g, t, o, out);
  } catch (const std::exception& e) {
    g_plan_error = std::string("plan_tree: ") + e.what();
    if (out) *out = nullptr;
    return IIF_ERR_STATE;
  }
}
static int32_t plan_tree_impl(const iif_graph_desc* g, const iif_tree_desc* t, const iif_plan_opts* o, iifb200_plan** out) {
  auto fail = [&](int32_t code, const std::string& msg) { g_plan_error = msg; return code; };
  if (!g || !t || !o || !out) return fail(IIF_ERR_ARG, "plan_tree: null argument");
  *out = nullptr;
  const bool uml = o->useMsgLikelihoods != 0;
  if (uml && (!g->factor_type || !g->var_type || !g->var_relative_kind || !g->var_relative_type))
    return fail(IIF_ERR_ARG, "plan_tree: useMsgLikelihoods needs the type tables of iif_graph_desc");
  const int nv = g->nvars, nf = g->nfactors, ncl = t->ncliques, N = o->N;
  if (nv < 1 || ncl < 1 || N < 2 || N > IIF_MAX_POINTS) return fail(IIF_ERR_ARG, "plan_tree: bad sizes");
  for (int f = 0; f < nf; ++f) {
    const iif_factor_desc& F = g->factors[f];
    if (F.arity < 1 || F.arity > IIF_MAX_ARITY) return fail(IIF_ERR_ARG, "plan_tree: factor arity out of range");
    for (int k = 0; k < F.arity; ++k)
      if (F.slot[k] < 0 || F.slot[k] >= nv) return fail(IIF_ERR_ARG, "plan_tree: factor variable index out of range");
  }
  for (int c = 0; c < ncl; ++c)
    if (t->parent[c] >= c) return fail(IIF_ERR_ARG, "plan_tree: cliques must be numbered parents first (parent id < child id)");
  iifb200_plan* P = new iifb200_plan();
  P->nvars = nv;
  P->dists.assign(g->dists, g->dists + g->ndists);
  P->dparams.assign(g->dparams, g->dparams + g->nparams);

  auto add_slot = [&](int var) {
    iif_slot_desc s = g->vars[var];
    s.cap = std::max(std::max(N, s.cap), 1);
    s.pts_off = 0;
    P->slots.push_back(s);
    return (int)P->slots.size() - 1;
  };
  for (int v = 0; v < nv; ++v) add_slot(v);   // main graph: slot v == variable v
  std::unordered_map<int, UpMsg> upmsg;                 // child clique -> joint up message
  std::unordered_map<int, std::vector<Inst>> kept_diffs; // clique -> differentials kept for the down solve
  // selectFactorType(T1, T2): the default relative factor between two variables of one type (DefaultNodeTypes.jl:12-31)
  auto sel_kind = [&](int a, int b) { return (g->var_type[a] == g->var_type[b]) ? g->var_relative_kind[a] : 0; };
  auto sel_type = [&](int a, int b) { return (g->var_type[a] == g->var_type[b] && g->var_relative_kind[a]) ? g->var_relative_type[a] : -1; };
  // clique-local copies
  std::vector<std::vector<int>> fr(ncl), sp(ncl), allv(ncl), children(ncl);
  std::vector<std::unordered_map<int, int>> cslot(ncl);
  for (int c = 0; c < ncl; ++c) {
    fr[c] = csr(t->frontal_off, t->frontals, c);
    sp[c] = csr(t->separator_off, t->separators, c);
    allv[c] = fr[c];
    allv[c].insert(allv[c].end(), sp[c].begin(), sp[c].end());
    for (int v : allv[c]) {
      if (v < 0 || v >= nv) { delete P; return fail(IIF_ERR_ARG, "plan_tree: clique variable index out of range"); }
      cslot[c][v] = add_slot(v);
    }
    if (t->parent[c] >= ncl || t->parent[c] == c) { delete P; return fail(IIF_ERR_ARG, "plan_tree: clique parent index out of range"); }
    if (t->parent[c] >= 0) children[t->parent[c]].push_back(c);
  }
  // running-intersection property: a clique's separators live in its parent (messages are addressed through them)
  for (int c = 0; c < ncl; ++c)
    if (t->parent[c] >= 0)
      for (int v : sp[c])
        if (!cslot[t->parent[c]].count(v)) { delete P; return fail(IIF_ERR_ARG, "plan_tree: separator missing from the parent clique"); }
  // variable -> factors (graph order)
  std::vector<std::vector<int>> by_var(nv);
  for (int f = 0; f < nf; ++f)
    for (int k = 0; k < g->factors[f].arity; ++k) {
      auto& l = by_var[g->factors[f].slot[k]];
      if (l.empty() || l.back() != f) l.push_back(f);
    }

  std::vector<Op> ops;
  int cur = -1, nconv = 0, n_msgs = 0;
  std::string err;
  // copy forwarding (see tree.compile_solve): a copy whose source was itself filled by a copy from X, X unchanged
  // since, reads X directly
  std::unordered_map<int, int> version;
  std::unordered_map<int, std::pair<int, int>> prov;
  auto ver = [&](int s) { auto it = version.find(s); return it == version.end() ? 0 : it->second; };
  auto wrote = [&](int s) { version[s] = ver(s) + 1; prov.erase(s); };
  auto add_copy = [&](int a, int b) {
    if (o->forward_copies) {
      auto it = prov.find(a);
      if (it != prov.end() && ver(it->second.first) == it->second.second) a = it->second.first;
    }
    Op op{IIF_S_COPY, a, b, cur, {a}, {b}, 0.05};
    ops.push_back(op);
    wrote(b);
    prov[b] = {a, ver(a)};
  };
  auto fac_instance = [&](int f, const std::vector<int>& slots) {
    iif_factor_desc F = g->factors[f];
    for (int k = 0; k < F.arity; ++k) F.slot[k] = slots[k];
    P->factors.push_back(F);
    return (int)P->factors.size() - 1;
  };
  auto add_prop = [&](int target_slot, const std::vector<Inst*>& use, int var) {
    if ((int)use.size() > IIF_MAX_FACTORS) {
      err = "plan_tree: " + std::to_string(use.size()) + " factors and messages on variable " + std::to_string(var) +
            " in clique " + std::to_string(cur) + " exceed IIF_MAX_FACTORS";
      return;
    }
    iif_prop_op p;
    memset(&p, 0, sizeof(p));
    p.target_slot = p.out_slot = target_slot;
    p.nfactors = (int)use.size();
    p.N = N;
    std::vector<int> rd{target_slot};
    bool anymh = false;
    for (size_t k = 0; k < use.size(); ++k) {
      const Inst& e = *use[k];
      p.factor[k] = e.fi;
      p.sfidx[k] = (int)(std::find(e.vars.begin(), e.vars.end(), var) - e.vars.begin()) + 1;
      anymh |= e.mh;
      rd.insert(rd.end(), e.rd.begin(), e.rd.end());
    }
    p.call_id = o->call_base + 16 * (int)P->props.size();
    p.any_multihypo = anymh ? 1 : 0;
    std::sort(rd.begin(), rd.end());
    rd.erase(std::unique(rd.begin(), rd.end()), rd.end());
    P->props.push_back(p);
    Op op{IIF_S_PROPAGATE, (int)P->props.size() - 1, 0, cur, rd, {target_slot}, (double)use.size() + 1.0};
    ops.push_back(op);
    nconv += (int)use.size();
    wrote(target_slot);
  };

  // ---- step 0: clique sub-graphs start from the main graph's beliefs
  for (int c = 0; c < ncl; ++c) {
    cur = c;
    for (int v : allv[c]) add_copy(v, cslot[c][v]);
  }
  // post-order (children in id order)
  std::vector<int> post;
  {
    std::vector<std::pair<int, bool>> stack;
    std::vector<int> roots;
    for (int c = 0; c < ncl; ++c) if (t->parent[c] < 0) roots.push_back(c);
    for (auto it = roots.rbegin(); it != roots.rend(); ++it) stack.push_back({*it, false});
    while (!stack.empty()) {
      auto [c, done] = stack.back();
      stack.pop_back();
      if (done) { post.push_back(c); continue; }
      stack.push_back({c, true});
      for (auto it = children[c].rbegin(); it != children[c].rend(); ++it) stack.push_back({*it, false});
    }
  }
  // ---- up pass
  for (int cid : post) {
    cur = cid;
    std::vector<Inst> inst;
    for (int f : csr(t->potential_off, t->potentials, cid)) {
      if (f < 0 || f >= nf) { delete P; return fail(IIF_ERR_ARG, "plan_tree: potential index out of range"); }
      const iif_factor_desc& F = g->factors[f];
      Inst e;
      for (int k = 0; k < F.arity; ++k) {
        auto it = cslot[cid].find(F.slot[k]);
        if (it == cslot[cid].end()) { delete P; return fail(IIF_ERR_ARG, "plan_tree: potential touches a variable outside its clique"); }
        e.vars.push_back(F.slot[k]);
        e.rd.push_back(it->second);
      }
      e.fi = fac_instance(f, e.rd);
      e.mh = F.nmh != 0;
      e.type = uml ? g->factor_type[f] : -1;
      e.prior = F.arity == 1;
      inst.push_back(e);
    }
    auto add_msg_prior = [&](int s, int src) {
      auto it = cslot[cid].find(s);
      iif_dist_desc D;
      memset(&D, 0, sizeof(D));
      D.kind = IIF_D_KDE; D.dim = g->vars[s].dim; D.slot = src; D.poff = (int)P->dparams.size();
      P->dists.push_back(D);
      iif_factor_desc F;
      memset(&F, 0, sizeof(F));
      F.kind = IIF_F_MSG_PRIOR; F.arity = 1; F.zdim = D.dim; F.dist = (int)P->dists.size() - 1;
      F.slot[0] = it->second; F.nullhypo = 0.0; F.inflation = o->inflation;
      P->factors.push_back(F);
      Inst e;
      e.vars = {s}; e.rd = {it->second, src}; e.fi = (int)P->factors.size() - 1; e.mh = false;
      e.type = uml ? g->msgprior_type : -1; e.prior = true; e.tag = 1;
      inst.push_back(e);
      n_msgs++;
    };
    if (uml)
      for (int ch : children[cid]) {
        const UpMsg& msg = upmsg[ch];
        // addLikelihoodsDifferential! (TreeMessageUtils.jl:225-233): the relatives of the joint message, each sampling its
        // measurements from the belief slot its IIF_S_DECONV op filled (`_sft(newBel)`, :321)
        for (const Relative& rel : msg.relatives) {
          iif_dist_desc D;
          memset(&D, 0, sizeof(D));
          D.kind = IIF_D_KDE; D.dim = rel.zdim; D.slot = rel.slot; D.poff = (int)P->dparams.size();
          P->dists.push_back(D);
          iif_factor_desc F;
          memset(&F, 0, sizeof(F));
          F.kind = rel.kind; F.arity = 2; F.zdim = rel.zdim; F.dist = (int)P->dists.size() - 1;
          F.slot[0] = cslot[cid][rel.v1]; F.slot[1] = cslot[cid][rel.v2]; F.nullhypo = 0.0; F.inflation = o->inflation;
          P->factors.push_back(F);
          Inst e;
          e.vars = {rel.v1, rel.v2}; e.rd = {F.slot[0], F.slot[1], rel.slot}; e.fi = (int)P->factors.size() - 1;
          e.mh = false; e.type = rel.type; e.prior = false; e.tag = 2;
          inst.push_back(e);
          n_msgs++;
        }
        // addLikelihoodPriorCommon! (TreeMessageUtils.jl:454-469)
        for (auto& pr : msg.priors) {
          bool touched = false;
          for (auto& e : inst) touched |= has(e.vars, pr.first);
          if (msg.hasPriors || !touched) add_msg_prior(pr.first, pr.second);
        }
      }
    for (int ch : children[cid])
      for (int s : sp[ch]) {   // addMsgFactors!: one MsgPrior per separator variable of the child
        if (uml) break;
        auto it = cslot[cid].find(s);
        if (it == cslot[cid].end()) continue;
        const int src = cslot[ch][s];
        iif_dist_desc D;
        memset(&D, 0, sizeof(D));
        D.kind = IIF_D_KDE; D.dim = g->vars[s].dim; D.slot = src; D.poff = (int)P->dparams.size();
        P->dists.push_back(D);
        iif_factor_desc F;
        memset(&F, 0, sizeof(F));
        F.kind = IIF_F_MSG_PRIOR; F.arity = 1; F.zdim = D.dim; F.dist = (int)P->dists.size() - 1;
        F.slot[0] = it->second; F.nullhypo = 0.0; F.inflation = o->inflation;
        P->factors.push_back(F);
        Inst e;
        e.vars = {s}; e.rd = {it->second, src}; e.fi = (int)P->factors.size() - 1; e.mh = false;
        inst.push_back(e);
        n_msgs++;
      }
    auto propagate = [&](int v) {
      std::vector<Inst*> use;
      for (auto& e : inst) if (has(e.vars, v)) use.push_back(&e);
      if (!use.empty()) add_prop(cslot[cid][v], use, v);
    };
    auto fmcmc = [&](const std::vector<int>& lbls, int iters) {
      if (lbls.size() == 1) iters = 1;
      for (int it = 0; it < iters; ++it) for (int v : lbls) propagate(v);
    };
    const std::vector<int> dfm = csr(t->directFrtlMsg_off, t->directFrtlMsg, cid), skip = csr(t->msgskip_off, t->msgskip, cid),
                           iter = csr(t->itervar_off, t->itervar, cid), dpm = csr(t->directPriorMsg_off, t->directPriorMsg, cid);
    fmcmc(dfm, 1);
    if (!skip.empty()) fmcmc(skip, 1);
    if (!iter.empty()) fmcmc(iter, o->gibbsIters);
    if (!dpm.empty()) {
      std::vector<int> l;
      for (int v : dpm) if (!has(skip, v)) l.push_back(v);
      fmcmc(l, 1);
    }
    if (!err.empty()) { delete P; return fail(IIF_ERR_ARG, err); }
    if (uml) {
      // only UPWARD_COMMON is deleted after the up solve (CSM :559-563): the differentials stay for the down solve
      for (auto& e : inst) if (e.tag == 2) kept_diffs[cid].push_back(e);
      if (t->parent[cid] >= 0) {
        // prepCliqueMsgUp (TreeMessageUtils.jl:667-703) -> _generateMsgJointRelativesPriors (:417-446)
        UpMsg msg;
        const std::vector<int>& seps = sp[cid];
        std::vector<int> idx(seps.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return g->vars[seps[a]].dim > g->vars[seps[b]].dim; });
        std::vector<int> dec, acc;
        for (int i : idx) dec.push_back(seps[i]);
        acc.assign(dec.rbegin(), dec.rend());
        std::vector<int> already;
        for (int s1 : dec) {
          already.push_back(s1);
          for (int s2 : acc) {
            if (has(already, s2)) continue;
            bool found;
            std::vector<int> types = path_types(inst, s1, s2, -1, found);           // isPathFactorsHomogeneous (DFG)
            if (!found || types.empty()) continue;
            bool homog = true;
            for (int ty : types) homog &= (ty == types[0]);
            if (!homog) continue;
            const int kind = sel_kind(s1, s2), type = sel_type(s1, s2);
            if (!kind || type != types[0]) continue;
            // newBel = manikde!(sft, approxDeconv(dummy factor)): a fresh measurement slot + one IIF_S_DECONV op
            iif_slot_desc ms = g->vars[s1];
            ms.cap = N; ms.pts_off = 0;
            P->slots.push_back(ms);
            const int mslot = (int)P->slots.size() - 1;
            const int d = g->vars[s1].dim;
            iif_dist_desc D;                     // the dummy's own measurement model: LinearRelative{N}() / CircularCircular(Normal(0, 0.1))
            memset(&D, 0, sizeof(D));
            D.poff = (int)P->dparams.size(); D.dim = d; D.slot = -1;
            if (kind == IIF_F_CIRCULAR_CIRCULAR) { D.kind = IIF_D_NORMAL; P->dparams.push_back(0.0); P->dparams.push_back(0.1); }
            else if (d == 1) { D.kind = IIF_D_NORMAL; P->dparams.push_back(0.0); P->dparams.push_back(1.0); }
            else {
              D.kind = IIF_D_MVNORMAL;
              for (int c = 0; c < d; ++c) P->dparams.push_back(0.0);
              for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) P->dparams.push_back(r == c ? 1.0 : 0.0);
            }
            P->dists.push_back(D);
            iif_factor_desc F;
            memset(&F, 0, sizeof(F));
            F.kind = kind; F.arity = 2; F.zdim = d; F.dist = (int)P->dists.size() - 1;
            F.slot[0] = cslot[cid][s1]; F.slot[1] = cslot[cid][s2]; F.nullhypo = 0.0; F.inflation = 5.0;
            P->factors.push_back(F);
            iif_deconv_op dc;
            dc.factor = (int)P->factors.size() - 1; dc.out_slot = mslot; dc.N = N; dc.call_id = -1;
            P->deconvs.push_back(dc);
            std::vector<int> rd{F.slot[0], F.slot[1]};
            std::sort(rd.begin(), rd.end());
            rd.erase(std::unique(rd.begin(), rd.end()), rd.end());
            Op op{IIF_S_DECONV, (int)P->deconvs.size() - 1, 0, cur, rd, {mslot}, 1.0};
            ops.push_back(op);
            msg.relatives.push_back({s1, s2, kind, type, mslot, d});
          }
        }
        // _findSubgraphsFactorType (:118-205): classes of separators connected through the default relative type
        std::unordered_map<int, int> count, cls;
        for (int s_ : seps) count[s_] = 0;
        for (auto& r : msg.relatives) { count[r.v1]++; count[r.v2]++; }
        int ncls = 0;
        for (int s_ : seps) if (count[s_] == 0) cls[s_] = ++ncls;
        std::vector<int> rest;
        for (int s_ : seps) if (!cls.count(s_)) rest.push_back(s_);
        for (int k1 : rest) {
          if (!cls.count(k1)) cls[k1] = ++ncls;
          std::vector<int> rest2;
          for (int s_ : seps) if (!cls.count(s_)) rest2.push_back(s_);
          for (int k2 : rest2) {
            const int ty = sel_type(k1, k2);
            bool found = false;
            std::vector<int> pth;
            if (ty >= 0) pth = path_types(inst, k1, k2, ty, found);
            if (ty < 0 || !found || pth.empty()) cls[k2] = ++ncls;
            else cls[k2] = cls[k1];
          }
        }
        std::vector<int> class_order;
        std::unordered_map<int, std::vector<int>> classes;
        for (int s_ : seps) {
          if (!classes.count(cls[s_])) class_order.push_back(cls[s_]);
          classes[cls[s_]].push_back(s_);
        }
        bool pot_has_prior = false, any_prior = false;                                  // :431, :681
        for (auto& e : inst) { pot_has_prior |= (e.tag == 0 && e.prior); any_prior |= e.prior; }
        for (int c_ : class_order) {
          const std::vector<int>& syms = classes[c_];
          if (syms.size() == 1 || pot_has_prior) {                                      // :402-408
            int md = 0;
            for (int v : syms) md = std::max(md, (int)g->vars[v].dim);
            std::vector<int> cand;
            for (int v : syms) if (g->vars[v].dim == md) cand.push_back(v);
            int best = 0, best_adj = -1;
            for (size_t i = 0; i < cand.size(); ++i) {     // sortperm(mdAdj; rev = true)[1]: first maximum
              int adj = 0;
              for (auto& e : inst) adj += has(e.vars, cand[i]) ? 1 : 0;
              if (adj > best_adj) { best_adj = adj; best = (int)i; }
            }
            msg.priors.push_back({cand[best], cslot[cid][cand[best]]});
          }
        }
        msg.hasPriors = any_prior;
        upmsg[cid] = msg;
      }
    }
  }
  const size_t n_up_ops = ops.size();
  // ---- down pass (parents before children); the root keeps its up-solve result
  if (o->downsolve) {
    for (auto itc = post.rbegin(); itc != post.rend(); ++itc) {
      const int cid = *itc, p = t->parent[cid];
      if (p < 0) continue;
      cur = cid;
      for (int s : sp[cid]) {
        auto ps = cslot[p].find(s);
        if (ps == cslot[p].end()) { delete P; return fail(IIF_ERR_ARG, "plan_tree: separator missing from the parent clique"); }
        add_copy(ps->second, cslot[cid][s]);
        n_msgs++;
      }
      // addDownVariableFactors!: every factor touching a frontal; outside variables are read from the main graph
      std::vector<int> touching;
      for (int v : fr[cid]) for (int f : by_var[v]) touching.push_back(f);
      std::sort(touching.begin(), touching.end());
      touching.erase(std::unique(touching.begin(), touching.end()), touching.end());
      std::vector<Inst> inst;
      if (uml) {
        // the clique sub-graph as the up solve left it: potentials + the children's differentials, no
        // addDownVariableFactors! (CliqueStateMachine.jl:825-834)
        touching = csr(t->potential_off, t->potentials, cid);
      }
      for (int f : touching) {
        const iif_factor_desc& F = g->factors[f];
        Inst e;
        for (int k = 0; k < F.arity; ++k) {
          auto it = cslot[cid].find(F.slot[k]);
          e.vars.push_back(F.slot[k]);
          e.rd.push_back(it != cslot[cid].end() ? it->second : F.slot[k]);
        }
        e.fi = fac_instance(f, e.rd);
        e.mh = F.nmh != 0;
        inst.push_back(e);
      }
      if (uml)
        for (auto& e : kept_diffs[cid]) inst.push_back(e);
      auto local_product = [&](int v) {
        std::vector<Inst*> use;
        for (auto& e : inst) if (has(e.vars, v)) use.push_back(&e);
        if (!use.empty()) add_prop(cslot[cid][v], use, v);
      };
      std::vector<int> iterF;   // frontals sharing a factor iterate downIters times
      for (auto& e : inst) {
        std::vector<int> f2;
        for (int v : e.vars) if (has(fr[cid], v)) f2.push_back(v);
        if (f2.size() > 1) for (int v : f2) if (!has(iterF, v)) iterF.push_back(v);
      }
      std::vector<int> iterO;
      for (int v : fr[cid]) if (has(iterF, v)) iterO.push_back(v);
      for (int v : fr[cid]) if (!has(iterO, v)) local_product(v);
      for (int it = 0; it < o->downIters; ++it) for (int v : iterO) local_product(v);
      if (!err.empty()) { delete P; return fail(IIF_ERR_ARG, err); }
    }
  }
  // ---- step 5: frontal beliefs go back to the main graph
  for (int c = 0; c < ncl; ++c) {
    cur = c;
    for (int v : fr[c]) add_copy(cslot[c][v], v);
  }
  for (size_t k = 0; k < P->deconvs.size(); ++k)      // Philox call ids: props first (16 apart), then the deconvolutions
    P->deconvs[k].call_id = o->call_base + 16 * ((int)P->props.size() + (int)k);
  // ---- waves, lanes
  std::vector<int> waves = levelize(ops);
  int nw = 0;
  for (int w : waves) nw = std::max(nw, w + 1);
  std::vector<int> lane = assign_lanes(ncl, t->parent, children, ops, waves, o->lanes);
  std::vector<int> order(ops.size());
  for (size_t i = 0; i < ops.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return waves[a] < waves[b]; });
  P->wave_off.assign(nw + 1, 0);
  for (int i : order) P->wave_off[waves[i] + 1]++;
  for (int w = 0; w < nw; ++w) P->wave_off[w + 1] += P->wave_off[w];
  for (int i : order) {
    iif_sched_op so;
    so.kind = ops[i].kind; so.a = ops[i].a; so.b = ops[i].b; so.lane = lane[i];
    P->ops.push_back(so);
  }
  int up_last = 0;
  for (size_t i = 0; i < n_up_ops; ++i) up_last = std::max(up_last, waves[i] + 1);
  P->up_last_wave = up_last;
  P->n_conv = nconv;
  P->n_prod = (int)P->props.size();
  P->n_msgs = n_msgs;
  int64_t off = 0;
  for (auto& s : P->slots) { s.pts_off = (int32_t)off; off += (int64_t)s.cap * s.dim; }
  *out = P;
  return IIF_OK;
}

int32_t iifb200_plan_counts(const iifb200_plan* p, int32_t* c) {
  if (!p || !c) return IIF_ERR_ARG;
  c[0] = (int32_t)p->slots.size(); c[1] = (int32_t)p->factors.size(); c[2] = (int32_t)p->dists.size();
  c[3] = (int32_t)p->dparams.size(); c[4] = (int32_t)p->props.size(); c[5] = (int32_t)p->ops.size();
  c[6] = (int32_t)p->wave_off.size() - 1; c[7] = p->n_conv; c[8] = p->n_prod; c[9] = p->n_msgs;
  c[10] = p->up_last_wave; c[11] = p->nvars; c[12] = (int32_t)p->deconvs.size();
  for (int k = 13; k < 16; ++k) c[k] = 0;
  return IIF_OK;
}

int32_t iifb200_plan_export(const iifb200_plan* p, iif_slot_desc* slots, iif_factor_desc* factors, iif_dist_desc* dists,
                            double* dparams, iif_prop_op* props, iif_sched_op* ops, int32_t* wave_off) {
  if (!p) return IIF_ERR_ARG;
  if (slots) std::copy(p->slots.begin(), p->slots.end(), slots);
  if (factors) std::copy(p->factors.begin(), p->factors.end(), factors);
  if (dists) std::copy(p->dists.begin(), p->dists.end(), dists);
  if (dparams) std::copy(p->dparams.begin(), p->dparams.end(), dparams);
  if (props) std::copy(p->props.begin(), p->props.end(), props);
  if (ops) std::copy(p->ops.begin(), p->ops.end(), ops);
  if (wave_off) std::copy(p->wave_off.begin(), p->wave_off.end(), wave_off);
  return IIF_OK;
}

int32_t iifb200_plan_export_deconvs(const iifb200_plan* p, iif_deconv_op* deconvs) {
  if (!p || (!deconvs && !p->deconvs.empty())) return IIF_ERR_ARG;
  std::copy(p->deconvs.begin(), p->deconvs.end(), deconvs);
  return IIF_OK;
}

int32_t iifb200_elimination_order_nd(int32_t nvars, int32_t nfactors, const int32_t* fac_off, const int32_t* fac_vars,
                                     int32_t* order_out) {
  auto fail = [&](int32_t code, const std::string& msg) { g_plan_error = msg; return code; };
  if (nvars < 0 || nfactors < 0 || (nvars && !order_out) || (nfactors && (!fac_off || !fac_vars))) return fail(IIF_ERR_ARG, "elimination_order_nd: null argument");
  // variable adjacency, neighbours in variable order (deterministic)
  std::vector<std::vector<int>> adj(nvars);
  for (int f = 0; f < nfactors; ++f)
    for (int a = fac_off[f]; a < fac_off[f + 1]; ++a)
      for (int b = fac_off[f]; b < fac_off[f + 1]; ++b) {
        const int u = fac_vars[a], w = fac_vars[b];
        if (u < 0 || u >= nvars || w < 0 || w >= nvars) return fail(IIF_ERR_ARG, "elimination_order_nd: variable id out of range");
        if (u != w) adj[u].push_back(w);
      }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  std::vector<int> order;
  order.reserve(nvars);
  std::vector<int> in(nvars, 0), seen(nvars, 0);   // stamps: in[v] == tag <=> v in the current node set
  int tag = 0, stag = 0;
  // BFS level sets of `nodes` (marked in[] == t) from `start`; returns the levels
  auto levels = [&](int t, int start, std::vector<std::vector<int>>& lev) {
    lev.clear();
    ++stag;
    seen[start] = stag;
    std::vector<int> frontier{start};
    while (!frontier.empty()) {
      lev.push_back(frontier);
      std::vector<int> nxt;
      for (int u : frontier)
        for (int w : adj[u])
          if (in[w] == t && seen[w] != stag) { seen[w] = stag; nxt.push_back(w); }
      frontier.swap(nxt);
    }
  };
  // recursive bisection of a sorted node list
  struct Rec {
    static void run(std::vector<int> nodes, std::vector<int>& order, std::vector<int>& in, int& tag,
                    const decltype(levels)& levels) {
      while (!nodes.empty()) {
        if (nodes.size() <= 2) { order.insert(order.end(), nodes.begin(), nodes.end()); return; }
        const int t = ++tag;
        for (int v : nodes) in[v] = t;
        std::vector<std::vector<int>> lev;
        levels(t, nodes.front(), lev);                   // start = smallest variable index
        size_t reached = 0;
        for (auto& l : lev) reached += l.size();
        if (reached < nodes.size()) {                    // disconnected: this component first, then the rest
          std::vector<int> comp, rest;
          for (auto& l : lev) comp.insert(comp.end(), l.begin(), l.end());
          std::sort(comp.begin(), comp.end());
          std::set_difference(nodes.begin(), nodes.end(), comp.begin(), comp.end(), std::back_inserter(rest));
          run(comp, order, in, tag, levels);
          nodes.swap(rest);
          continue;
        }
        const int far = lev.back().back();               // last variable the search reached: a peripheral one
        const int t2 = ++tag;
        for (int v : nodes) in[v] = t2;
        levels(t2, far, lev);
        if (lev.size() < 3) { order.insert(order.end(), nodes.begin(), nodes.end()); return; }
        const double half = (double)nodes.size() / 2.0;
        size_t acc = 0;
        int cut = (int)lev.size() / 2;
        for (size_t i = 0; i < lev.size(); ++i) {
          acc += lev[i].size();
          if ((double)acc >= half) { cut = std::min(std::max((int)i, 1), (int)lev.size() - 2); break; }
        }
        std::vector<int> left, right, sep = lev[cut];
        for (int i = 0; i < cut; ++i) left.insert(left.end(), lev[i].begin(), lev[i].end());
        for (size_t i = cut + 1; i < lev.size(); ++i) right.insert(right.end(), lev[i].begin(), lev[i].end());
        std::sort(left.begin(), left.end());
        std::sort(right.begin(), right.end());
        std::sort(sep.begin(), sep.end());
        run(left, order, in, tag, levels);
        run(right, order, in, tag, levels);
        order.insert(order.end(), sep.begin(), sep.end());
        return;
      }
    }
  };
  std::vector<int> all(nvars);
  for (int v = 0; v < nvars; ++v) all[v] = v;
  Rec::run(all, order, in, tag, levels);
  if ((int)order.size() != nvars) return fail(IIF_ERR_STATE, "elimination_order_nd: internal error");
  std::copy(order.begin(), order.end(), order_out);
  return IIF_OK;
}

int32_t iifb200_elimination_order_is(int32_t nvars, int32_t nfactors, const int32_t* fac_off, const int32_t* fac_vars,
                                     int32_t slack, int32_t* order_out) {
  auto fail = [&](int32_t code, const std::string& msg) { g_plan_error = msg; return code; };
  if (nvars < 0 || nfactors < 0 || slack < 0 || (nvars && !order_out) || (nfactors && (!fac_off || !fac_vars)))
    return fail(IIF_ERR_ARG, "elimination_order_is: bad argument");
  std::vector<std::set<int>> adj(nvars);
  for (int f = 0; f < nfactors; ++f)
    for (int a = fac_off[f]; a < fac_off[f + 1]; ++a)
      for (int b = fac_off[f]; b < fac_off[f + 1]; ++b) {
        const int u = fac_vars[a], w = fac_vars[b];
        if (u < 0 || u >= nvars || w < 0 || w >= nvars) return fail(IIF_ERR_ARG, "elimination_order_is: variable id out of range");
        if (u != w) adj[u].insert(w);
      }
  std::vector<char> alive(nvars, 1), blocked(nvars, 0);
  std::vector<int> order;
  order.reserve(nvars);
  int remaining = nvars;
  while (remaining > 0) {
    if (remaining <= 3) {
      for (int v = 0; v < nvars; ++v) if (alive[v]) order.push_back(v);
      break;
    }
    size_t mind = SIZE_MAX;
    for (int v = 0; v < nvars; ++v) if (alive[v]) mind = std::min(mind, adj[v].size());
    std::fill(blocked.begin(), blocked.end(), 0);
    std::vector<int> chosen;
    for (int v = 0; v < nvars; ++v) {              // greedy maximal independent set in variable order
      if (!alive[v] || blocked[v] || adj[v].size() > mind + (size_t)slack) continue;
      chosen.push_back(v);
      blocked[v] = 1;
      for (int w : adj[v]) blocked[w] = 1;
    }
    for (int v : chosen) {                         // eliminate: the neighbours become a clique
      const std::vector<int> nb(adj[v].begin(), adj[v].end());
      for (int a : nb) {
        adj[a].erase(v);
        for (int b : nb) if (b != a) adj[a].insert(b);
      }
      adj[v].clear();
      alive[v] = 0;
      --remaining;
    }
    order.insert(order.end(), chosen.begin(), chosen.end());
  }
  std::copy(order.begin(), order.end(), order_out);
  return IIF_OK;
}

}  // extern "C"

// accessors for iifb200.cu (iifb200_plan_upload)
const std::vector<iif_deconv_op>& iif_plan_deconvs(const iifb200_plan* p) { return p->deconvs; }
const std::vector<iif_slot_desc>& iif_plan_slots(const iifb200_plan* p) { return p->slots; }
const std::vector<iif_factor_desc>& iif_plan_factors(const iifb200_plan* p) { return p->factors; }
const std::vector<iif_dist_desc>& iif_plan_dists(const iifb200_plan* p) { return p->dists; }
const std::vector<double>& iif_plan_dparams(const iifb200_plan* p) { return p->dparams; }
const std::vector<iif_prop_op>& iif_plan_props(const iifb200_plan* p) { return p->props; }
const std::vector<iif_sched_op>& iif_plan_ops(const iifb200_plan* p) { return p->ops; }
const std::vector<int32_t>& iif_plan_wave_off(const iifb200_plan* p) { return p->wave_off; }
