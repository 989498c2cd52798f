"""Bayes-tree construction and lowering of a `solveTree!` pass to a wave schedule.

Host orchestration is OUT OF SCOPE of the B200 hot path (the factor graph, tree
build and clique state machine stay in Julia).  The GPU batcher, however, consumes the tree's
*output* — per-clique variable / factor lists, Gibbs variable classes and tree edges — so this
module mirrors just enough of the reference to produce that descriptor here (Julia is absent):

  getEliminationOrder      src/services/BayesNet.jl:19-65
  buildBayesNet!           src/services/BayesNet.jl:139-187
  buildTree!/newPotential  src/services/JunctionTreeUtils.jl:435-496
  setCliqPotentials!       src/services/JunctionTreeUtils.jl:1045-1082
  compCliqAssocMatrices!   src/services/JunctionTreeUtils.jl:1294-1341
  setCliqMCIDs! and friends  src/services/JunctionTreeUtils.jl:1352-1523
  upGibbsCliqueDensity     src/services/SolveTree.jl:164-239
  solveCliqDownFrontalProducts!  src/CliqueStateMachine/services/CliqStateMachineUtils.jl:479-571
  CSM up/down message flow src/CliqueStateMachine/services/CliqueStateMachine.jl:212-966

The lowering turns every `propagateBelief` of the pass into an iif_prop_op on clique-local
belief slots and levelises them into waves of mutually independent ops (slot hazards), which
libiifb200 captures as one CUDA graph.
"""
import bisect
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import _abi as A
from . import compile as CP
from . import graph as G


# ------------------------------------------------------------------ elimination order
def getEliminationOrder(fg: G.FactorGraph, ordering: str = "qr") -> List[str]:
    """BayesNet.jl:19-65.  `qr`: column-pivoted QR of the biadjacency matrix, reversed.
    `natural`: insertion order.  `nd`: nested dissection on the variable adjacency graph
    (a user-supplied `eliminationOrder`, SolverAPI.jl:338, that makes chain trees bushy)."""
    labels = list(fg.variables)
    if ordering == "natural":
        return labels
    if ordering == "nd":
        return nested_dissection_order(fg)
    if ordering == "is":
        return independent_set_order(fg)
    if ordering == "qr":
        import scipy.linalg
        facs = list(fg.factors.values())
        Amat = np.zeros((len(facs), len(labels)))
        col = {l: i for i, l in enumerate(labels)}
        for r, f in enumerate(facs):
            for v in f.variables:
                Amat[r, col[v]] = 1.0
        _, _, p = scipy.linalg.qr(Amat, pivoting=True, mode="economic")
        return [labels[i] for i in p[::-1]]
    raise ValueError(f"unknown ordering {ordering}")


def nested_dissection_order(fg: G.FactorGraph) -> List[str]:
    """Recursive bisection by BFS level sets; separators are eliminated last."""
    adj: Dict[str, set] = {l: set() for l in fg.variables}
    for f in fg.factors.values():
        for a in f.variables:
            for b in f.variables:
                if a != b:
                    adj[a].add(b)
    # neighbour lists in variable order: set iteration order depends on the process' string-hash seed and would make
    # the elimination order (hence the tree and the plan) differ from run to run
    adj = {l: sorted(ns, key=lambda x: fg.variables[x].index) for l, ns in adj.items()}
    order: List[str] = []

    def bfs_far(nodes, start):
        seen, frontier, last = {start}, [start], start
        while frontier:
            nxt = []
            for u in frontier:
                for w in adj[u]:
                    if w in nodes and w not in seen:
                        seen.add(w)
                        nxt.append(w)
            if nxt:
                last = nxt[-1]
            frontier = nxt
        return last, seen

    def levels(nodes, start):
        lev, seen, frontier = [], {start}, [start]
        while frontier:
            lev.append(frontier)
            nxt = []
            for u in frontier:
                for w in adj[u]:
                    if w in nodes and w not in seen:
                        seen.add(w)
                        nxt.append(w)
            frontier = nxt
        return lev, seen

    def rec(nodes: set):
        # iterative over connected components / recursion depth O(log n)
        if len(nodes) <= 2:
            order.extend(sorted(nodes, key=lambda l: fg.variables[l].index))
            return
        start = min(nodes, key=lambda l: fg.variables[l].index)
        far, comp = bfs_far(nodes, start)
        if len(comp) < len(nodes):           # disconnected: handle components separately
            rec(comp)
            rec(nodes - comp)
            return
        lev, _ = levels(nodes, far)
        if len(lev) < 3:
            order.extend(sorted(nodes, key=lambda l: fg.variables[l].index))
            return
        half, acc, cut = len(nodes) / 2, 0, len(lev) // 2
        for i, l in enumerate(lev):
            acc += len(l)
            if acc >= half:
                cut = min(max(i, 1), len(lev) - 2)
                break
        sep = set(lev[cut])
        left = set().union(*lev[:cut])
        right = set().union(*lev[cut + 1:])
        rec(left)
        rec(right)
        order.extend(sorted(sep, key=lambda l: fg.variables[l].index))

    rec(set(fg.variables))
    return order


def independent_set_order(fg: G.FactorGraph, slack: int = 1) -> List[str]:
    """Generalised odd-even (cyclic) reduction: rounds of eliminating a maximal independent set of low-degree variables
    (degree <= minimum + slack in the current elimination graph, greedy in variable order), fill-in edges added between
    the neighbours of every eliminated variable.  Variables of one round are pairwise non-adjacent, so their cliques are
    siblings in the Bayes tree (one wide wave), the number of rounds is O(log n) for chains and grids, and the low degree
    bound keeps the cliques small (few frontals => no long in-clique Gibbs iterations).  On a chain this is odd-even
    reduction; on grids with loop closures it gives smaller separator cliques than level-set bisection."""
    adj: Dict[str, set] = {l: set() for l in fg.variables}
    for f in fg.factors.values():
        for a in f.variables:
            for b in f.variables:
                if a != b:
                    adj[a].add(b)
    idx = {l: v.index for l, v in fg.variables.items()}
    remaining = set(fg.variables)
    order: List[str] = []
    while remaining:
        if len(remaining) <= 3:
            order.extend(sorted(remaining, key=lambda l: idx[l]))
            break
        mind = min(len(adj[l]) for l in remaining)
        cand = sorted((l for l in remaining if len(adj[l]) <= mind + slack), key=lambda l: idx[l])
        chosen, blocked = [], set()
        for v in cand:
            if v in blocked:
                continue
            chosen.append(v)
            blocked.add(v)
            blocked |= adj[v]
        for v in chosen:
            nb = adj[v]
            for a in nb:
                adj[a].discard(v)
                adj[a] |= (nb - {a})
            remaining.discard(v)
            del adj[v]
        order.extend(chosen)
    return order


# ------------------------------------------------------------------ Bayes tree
@dataclass
class TreeClique:
    id: int
    frontals: List[str]
    separators: List[str]
    parent: Optional[int] = None
    children: List[int] = field(default_factory=list)
    potentials: List[str] = field(default_factory=list)      # factor labels (up solve)
    inmsgIDs: List[str] = field(default_factory=list)        # separator vars of the children (with repeats)
    directFrtlMsgIDs: List[str] = field(default_factory=list)
    msgskipIDs: List[str] = field(default_factory=list)
    itervarIDs: List[str] = field(default_factory=list)
    directPriorMsgIDs: List[str] = field(default_factory=list)

    @property
    def allvars(self):
        return self.frontals + self.separators


@dataclass
class BayesTree:
    cliques: List[TreeClique]
    frontal_of: Dict[str, int]
    eliminationOrder: List[str]

    @property
    def roots(self):
        return [c.id for c in self.cliques if c.parent is None]

    def depth(self):
        d = {}
        for c in self.cliques:   # parents are created before children (reverse elimination order)
            d[c.id] = 0 if c.parent is None else d[c.parent] + 1
        return d


def buildBayesNet(fg: G.FactorGraph, order: List[str]) -> Dict[str, List[str]]:
    """Symbolic elimination -> conditional p(v | separator)  (BayesNet.jl:139-187)."""
    pos = {v: i for i, v in enumerate(order)}
    # active "factors" as variable sets; chain-rule marginals are added as eliminated vars go
    live = [set(f.variables) for f in fg.factors.values()]
    sep: Dict[str, List[str]] = {}
    for v in order:
        touching = [s for s in live if v in s]
        Si: List[str] = []
        for s in touching:
            for w in sorted(s, key=lambda x: fg.variables[x].index):
                if w != v and w not in Si:
                    Si.append(w)
        live = [s for s in live if v not in s]
        sep[v] = Si if v != order[-1] else []
        if Si:
            live.append(set(Si))
    del pos
    return sep


def buildTree(fg: G.FactorGraph, order: List[str]) -> BayesTree:
    """buildTree!/newPotential (JunctionTreeUtils.jl:435-496; Kaess et al. Bayes tree Alg. 2)."""
    sep = buildBayesNet(fg, order)
    pos = {v: i for i, v in enumerate(order)}
    cliques: List[TreeClique] = []
    frontal_of: Dict[str, int] = {}
    for var in reversed(order):
        Sj = sep[var]
        if not Sj:
            c = TreeClique(len(cliques), [var], [])
            cliques.append(c)
            frontal_of[var] = c.id
            continue
        first = min(Sj, key=lambda s: pos[s])          # first eliminated separator variable
        cp = cliques[frontal_of[first]]
        if sorted(cp.frontals + cp.separators) == sorted(Sj):
            cp.frontals.append(var)                    # appendClique!: add as a frontal
            frontal_of[var] = cp.id
        else:
            c = TreeClique(len(cliques), [var], list(Sj), parent=cp.id)   # newChildClique!
            cp.children.append(c.id)
            cliques.append(c)
            frontal_of[var] = c.id
    tree = BayesTree(cliques, frontal_of, list(order))
    _buildCliquePotentials(fg, tree)
    return tree


def _postorder(tree: BayesTree) -> List[int]:
    out, stack = [], [(r, False) for r in reversed(tree.roots)]
    while stack:
        cid, done = stack.pop()
        if done:
            out.append(cid)
            continue
        stack.append((cid, True))
        for ch in reversed(tree.cliques[cid].children):
            stack.append((ch, False))
    return out


def _buildCliquePotentials(fg: G.FactorGraph, tree: BayesTree):
    """buildCliquePotentials (post-order): setCliqPotentials! + assoc matrices + setCliqMCIDs!."""
    used = set()
    by_var = _factors_by_variable(fg)
    for cid in _postorder(tree):
        c = tree.cliques[cid]
        allv = set(c.allvars)
        pots = []
        for f in _factors_touching(fg, by_var, c.frontals):   # getFactorsAmongVariablesOnly(unused) ∩ frontal factors
            if f.label in used:
                continue
            if set(f.variables) <= allv:
                pots.append(f.label)
        c.potentials = pots
        used.update(pots)
        c.inmsgIDs = [s for ch in c.children for s in tree.cliques[ch].separators]   # collectSeparators
        _setCliqMCIDs(fg, c)


def _factors_by_variable(fg: G.FactorGraph) -> Dict[str, list]:
    by_var: Dict[str, list] = {l: [] for l in fg.variables}
    for f in fg.factors.values():
        for v in f.variables:
            by_var[v].append(f)
    return by_var


def _factors_touching(fg: G.FactorGraph, by_var, variables):
    """factors with at least one of `variables`, in graph (insertion) order, each once"""
    seen, out = set(), []
    for v in variables:
        for f in by_var[v]:
            if f.label not in seen:
                seen.add(f.label)
                out.append(f)
    out.sort(key=lambda f: f.index)
    return out


def _setCliqMCIDs(fg: G.FactorGraph, c: TreeClique):
    """JunctionTreeUtils.jl:1352-1523."""
    cols = c.allvars
    nf = len(c.frontals)
    assoc = np.zeros((len(c.potentials), len(cols)), dtype=int)
    for i, fl in enumerate(c.potentials):
        for v in fg.factors[fl].variables:
            if v in cols:
                assoc[i, cols.index(v)] = 1
    msg = np.zeros((len(c.inmsgIDs), len(cols)), dtype=int)
    for i, v in enumerate(c.inmsgIDs):
        if v in cols:
            msg[i, cols.index(v)] = 1
    mat = np.vstack([assoc, msg]) if len(cols) else np.zeros((0, 0), dtype=int)
    colsum = mat.sum(axis=0) if mat.size else np.zeros(len(cols), dtype=int)
    # directPriorMsgIDs :1368-1382: columns whose every row is a singleton row
    sing_rows = mat.sum(axis=1) == 1 if mat.size else np.zeros(0, dtype=bool)
    sums_sing = mat[sing_rows].sum(axis=0) if mat.size else colsum
    c.directPriorMsgIDs = [cols[j] for j in range(len(cols)) if sums_sing[j] - colsum[j] == 0]
    # directAssignmentIDs :1394-1409
    asum, msum = assoc.sum(axis=0), msg.sum(axis=0)
    directvars = [cols[j] for j in range(len(cols)) if colsum[j] == 1 and asum[j] == 1]
    # mcmcIterationIDs :1411-1431
    multi = [cols[j] for j in range(len(cols)) if colsum[j] > 1]
    alliter = [v for v in _union(directvars, multi) if v not in c.directPriorMsgIDs]
    # ordering :1448-1484: non-singleton vars first (ascending factor count), then singleton vars
    upmsg = [cols[j] for j in range(len(cols)) if msum[j] >= 1]
    prior_rows = assoc.sum(axis=1) == 1
    prior_vars = [cols[j] for j in range(len(cols)) if assoc[prior_rows][:, j].sum() > 0] if len(c.potentials) else []
    # NB getCliqVarIdsPriors(cliq, allids, partials=true) masks prior rows with `partialpotential`, i.e.
    # only *partial* priors count as singletons there; full priors are not in `allsings`.
    partial_prior_vars = []
    for i, fl in enumerate(c.potentials):
        if prior_rows[i] and isinstance(fg.factors[fl].fnc, (G.PartialPrior, G.ManifoldPriorPartial)):
            partial_prior_vars += [cols[j] for j in range(len(cols)) if assoc[i, j]]
    del prior_vars
    allsings = _union(upmsg, partial_prior_vars)
    singl = [v for v in alliter if v in allsings]
    nons = [v for v in alliter if v not in singl]
    key = lambda v: colsum[cols.index(v)]  # noqa: E731
    c.itervarIDs = sorted(nons, key=key) + sorted(singl, key=key)     # sortperm is stable
    # skipThroughMsgsIDs :1343-1355 (separator columns with exactly one row, and it is a message)
    c.msgskipIDs = [cols[j] for j in range(nf, len(cols)) if colsum[j] == 1 and msum[j] == 1]
    # directFrtlMsgIDs :1384-1392
    c.directFrtlMsgIDs = [cols[j] for j in range(nf) if colsum[j] == 1 and msum[j] == 1]


def _union(a, b):
    out = list(a)
    for x in b:
        if x not in out:
            out.append(x)
    return out


# ------------------------------------------------------------------ lowering to a wave schedule
@dataclass
class SolvePlan:
    tables: CP.Tables
    frozen: dict
    props: list                       # prop op spec dicts
    sched: list                       # (kind, a, b) in program order
    wave_off: List[int]
    sched_waved: list                 # (kind, a, b) sorted by wave
    var_slot: Dict[str, int]          # global (main-graph) slot of each variable
    n_conv: int
    n_prod: int
    n_msgs: int
    up_last_wave: int = 0             # waves [0, up_last_wave) cover init copies + the up pass
    op_clique: Optional[List[int]] = None    # clique id of every op of `sched_waved`
    op_wave: Optional[List[int]] = None      # wave of every op of `sched_waved`
    op_reads: Optional[List[List[int]]] = None   # slots read by every op of `sched_waved`
    op_writes: Optional[List[List[int]]] = None
    slot_clique: Optional[Dict[int, int]] = None  # clique-local slot -> clique id (main-graph slots absent)
    deconvs: Optional[list] = None    # IIF_S_DECONV specs (useMsgLikelihoods=true): factor, out_slot, N, call_id
    op_lane: Optional[List[int]] = None   # lane of every op of `sched_waved` (0 = none), see assign_lanes
    c_plan: object = None             # planner.CPlan when the plan was made by iifb200_plan_tree (libiifb200.so)
    up_messages: Optional[dict] = None  # useMsgLikelihoods: clique id -> content of its joint up message (LikelihoodMessage
    #                                     .jointmsg: relatives as variable pairs, prior variables, hasPriors) for inspection


def _levelize(ops, reads, writes):
    """wave(op) = 1 + max wave of earlier ops it conflicts with (RAW / WAR / WAW on slots)."""
    last_w: Dict[int, int] = {}
    last_r: Dict[int, int] = {}
    waves = []
    for i in range(len(ops)):
        w = 0
        for s in reads[i]:
            if s in last_w:
                w = max(w, last_w[s] + 1)
        for s in writes[i]:
            if s in last_w:
                w = max(w, last_w[s] + 1)
            if s in last_r:
                w = max(w, last_r[s] + 1)
        waves.append(w)
        for s in reads[i]:
            last_r[s] = max(last_r.get(s, -1), w)
        for s in writes[i]:
            last_w[s] = w
    return waves


def selectFactorType(t1: G.InferenceVariable, t2: G.InferenceVariable):
    """selectFactorType — src/services/DefaultNodeTypes.jl:12-31: the default relative factor between two
    variable types (Position{N} -> LinearRelative{N}; otherwise getfield(Module, Symbol(T1, T2)), which exists
    for Circular -> CircularCircular).  Returns (type name, default-constructed factor) or None."""
    if t1.circ_mask == 0 and t2.circ_mask == 0 and t1.dim == t2.dim and t1.name.startswith("Position"):
        d = t1.dim
        Z = G.Normal(0.0, 1.0) if d == 1 else G.MvNormal(np.zeros(d), np.eye(d))   # LinearRelative{N}() default
        return "LinearRelative", G.LinearRelative(Z)
    if t1.name == "Circular" and t2.name == "Circular":
        return "CircularCircular", G.CircularCircular(G.Normal(0.0, 0.1))
    return None


def _shortest_path_factor_types(inst, src: str, dst: str, only_type: Optional[str] = None):
    """findShortestPathDijkstra(dfg, from, to; typeFactors) on the bipartite variable/factor graph of a clique
    sub-graph (DFG, un-vendored; unit edge weights => BFS, neighbours in insertion order): the factor type names
    along the path, or None when the two variables are not connected."""
    if src == dst:
        return []
    prev = {src: None}
    frontier = [src]
    while frontier:
        nxt = []
        for u in frontier:
            for k, e in enumerate(inst):
                if u not in e["variables"] or (only_type is not None and e["type"] != only_type):
                    continue
                for w in e["variables"]:
                    if w not in prev:
                        prev[w] = (u, k)
                        nxt.append(w)
        if dst in prev:
            break
        frontier = nxt
    if dst not in prev:
        return None
    types, w = [], dst
    while prev[w] is not None:
        u, k = prev[w]
        types.append(inst[k]["type"])
        w = u
    return types[::-1]


def assign_lanes(tree: BayesTree, op_clique, op_weight, waves, reads, writes, nlanes: int) -> List[int]:
    """Lane (1..nlanes) of every op, 0 = none.  Lanes are groups of disjoint sub-trees of the Bayes tree: cliques in
    disjoint sub-trees are independent until their common ancestor (CliqueStateMachine.jl:221-234 waits only on own
    children), so the device may run one lane's next wave while another lane's wave is still draining.  The top of the
    tree (the "cap") stays lane 0; a wave that holds a lane-0 op is a barrier.

    1. split the heaviest sub-tree at its root until there are ~2*nlanes sub-trees, pack them into lanes (LPT);
    2. hazard check on the slots: an op that conflicts (RAW / WAR / WAW) with an earlier op of a DIFFERENT lane with no
       barrier wave in between is demoted to lane 0 (e.g. the write-back copies into the main graph)."""
    n = len(op_clique)
    if nlanes < 2 or n == 0:
        return [0] * n
    cl = tree.cliques
    w_cl = [0.0] * len(cl)
    for i in range(n):
        if op_clique[i] >= 0:
            w_cl[op_clique[i]] += op_weight[i]
    sub = list(w_cl)                                     # sub-tree weights (children are created after parents)
    for c in reversed(cl):
        if c.parent is not None:
            sub[c.parent] += sub[c.id]
    frontier = list(tree.roots)
    cap = set()
    while len(frontier) < 2 * nlanes:
        cand = [c for c in frontier if cl[c].children]
        if not cand:
            break
        big = max(cand, key=lambda c: sub[c])
        frontier.remove(big)
        cap.add(big)
        frontier += cl[big].children
    load = [0.0] * (nlanes + 1)
    lane_of_clique = [0] * len(cl)
    for r in sorted(frontier, key=lambda c: -sub[c]):
        ln = min(range(1, nlanes + 1), key=lambda k: load[k])
        load[ln] += sub[r]
        stack = [r]
        while stack:
            c = stack.pop()
            lane_of_clique[c] = ln
            stack += cl[c].children
    lane = [lane_of_clique[op_clique[i]] if op_clique[i] >= 0 else 0 for i in range(n)]
    # hazard check in wave order; barrier waves (any lane-0 op) separate everything before from everything after
    order = sorted(range(n), key=lambda i: (waves[i], i))
    for _ in range(n):                                   # demotions create new barrier waves: iterate to a fixed point
        barrier = sorted({waves[i] for i in range(n) if lane[i] == 0})

        def separated(wa, wb):                          # a barrier wave in (wa, wb], or wa itself a barrier
            k = bisect.bisect_left(barrier, wa)
            return k < len(barrier) and barrier[k] <= wb
        last_w: Dict[int, tuple] = {}
        last_r: Dict[int, list] = {}
        changed = False
        for i in order:
            li, wi = lane[i], waves[i]
            conflict = False
            if li != 0:
                for s_ in reads[i]:
                    a = last_w.get(s_)
                    if a and a[0] != li and a[0] != 0 and not separated(a[1], wi):
                        conflict = True
                for s_ in writes[i]:
                    a = last_w.get(s_)
                    if a and a[0] != li and a[0] != 0 and not separated(a[1], wi):
                        conflict = True
                    for (lr, wr) in last_r.get(s_, []):
                        if lr != li and lr != 0 and not separated(wr, wi):
                            conflict = True
            if conflict:
                lane[i] = 0
                changed = True
                break
            for s_ in writes[i]:
                last_w[s_] = (li, wi)
                last_r[s_] = []
            for s_ in reads[i]:
                last_r.setdefault(s_, []).append((li, wi))
        if not changed:
            break
    return lane


def compile_solve(fg: G.FactorGraph, tree: BayesTree, N: Optional[int] = None, downsolve: bool = True,
                  gibbsIters: Optional[int] = None, downIters: int = 3,
                  useMsgLikelihoods: Optional[bool] = None, lanes: int = 0,
                  forward_copies: bool = True) -> SolvePlan:
    """Lower one solveTree! (up + down pass) to slots, props and waves.

    useMsgLikelihoods=false (SolverParams default): up messages are one MsgPrior per separator variable.
    useMsgLikelihoods=true (test/fourdoortest.jl:19, testCircular.jl:12): a child sends the joint of
    its separators as *differential* relative factors + at most one MsgPrior per connected class
    (prepCliqueMsgUp -> _generateMsgJointRelativesPriors, TreeMessageUtils.jl:417-446); every differential is one
    IIF_S_DECONV op (approxDeconv + manikde! on the device) whose belief slot is the measurement density of the
    relative factor the parent adds; differentials stay in the sub-graph for the down solve, which then does not
    merge extra graph factors (CliqueStateMachine.jl:825-834)."""
    sp = fg.solverParams
    N = N or sp.N
    iters = gibbsIters or sp.gibbsIters
    uml = sp.useMsgLikelihoods if useMsgLikelihoods is None else bool(useMsgLikelihoods)
    T = CP.Tables()
    # global slots: the main graph's variables (VariableNodeData)
    var_slot = {l: T.add_slot(v.vartype, max(N, v.val.shape[0], 1)) for l, v in fg.variables.items()}
    # clique-local copies (buildCliqSubgraph! deep-copies the clique's variables, CSM step 0b)
    cslot: Dict[tuple, int] = {}
    for c in tree.cliques:
        for v in c.allvars:
            cslot[(c.id, v)] = T.add_slot(fg.variables[v].vartype, max(N, fg.variables[v].val.shape[0], 1))

    props, sched, reads, writes, opc = [], [], [], [], []
    deconvs = []                     # IIF_S_DECONV specs (useMsgLikelihoods)
    upmsg: Dict[int, dict] = {}      # child clique id -> joint up message (relatives, priors, hasPriors)
    kept_diffs: Dict[int, list] = {}  # clique id -> differential factor instances kept for the down solve
    nconv = [0]
    cur = [-1]     # clique whose ops are being emitted

    # copy forwarding: a separator value travels down the tree through a chain of slot copies (parent -> child ->
    # grandchild ...).  A copy whose source was itself filled by a copy from X, with X unchanged since, reads X
    # directly: same value, but the chain no longer serialises one wave per tree level.
    version: Dict[int, int] = {}
    prov: Dict[int, tuple] = {}

    def wrote(slot):
        version[slot] = version.get(slot, 0) + 1
        prov.pop(slot, None)

    def add_copy(a, b):
        if forward_copies and a in prov and version.get(prov[a][0], 0) == prov[a][1]:
            a = prov[a][0]
        sched.append((A.S_COPY, a, b))
        reads.append([a])
        writes.append([b])
        opc.append(cur[0])
        wrote(b)
        prov[b] = (a, version.get(a, 0))

    def fac_instance(f: G.DFGFactor, slot_of):
        return T.add_factor(f.fnc, [slot_of(v) for v in f.variables], f.multihypo, f.nullhypo, f.inflation)

    def add_prop(target_label, target_slot, fac_list, rd_slots):
        """fac_list: [(factor_table_idx, sfidx, is_multihypo)]"""
        if len(fac_list) > A.IIF_MAX_FACTORS:     # never truncate: dropped likelihoods / messages lose whole sub-trees
            raise A.IIFB200Error(f"solve plan: {len(fac_list)} factors and messages on {target_label} in clique "
                                 f"{cur[0]} exceed IIF_MAX_FACTORS = {A.IIF_MAX_FACTORS}")
        spec = dict(target_slot=target_slot, out_slot=target_slot, factors=[(fi, sf) for fi, sf, _ in fac_list],
                    N=N, call_id=16 * len(props), any_multihypo=int(any(m for _, _, m in fac_list)))
        props.append(spec)
        sched.append((A.S_PROPAGATE, len(props) - 1, 0))
        reads.append(sorted(set(rd_slots) | {target_slot}))
        writes.append([target_slot])
        opc.append(cur[0])
        nconv[0] += len(fac_list)
        wrote(target_slot)

    # ---- step 0: clique sub-graphs start from the main graph's (graph-init) beliefs
    for c in tree.cliques:
        cur[0] = c.id
        for v in c.allvars:
            add_copy(var_slot[v], cslot[(c.id, v)])

    post = _postorder(tree)
    by_var = _factors_by_variable(fg)
    n_msgs = 0
    # ---- up pass (children before parents)
    for cid in post:
        c = tree.cliques[cid]
        cur[0] = cid
        slot_of = lambda v, cid=cid: cslot[(cid, v)]  # noqa: E731
        # factors of the clique sub-graph: potentials + MsgPrior per child up-message belief
        inst = []   # factor instances of the clique sub-graph
        for fl in c.potentials:
            f = fg.factors[fl]
            inst.append(dict(variables=f.variables, fi=fac_instance(f, slot_of), mh=G.isMultihypo(f),
                             rd=[slot_of(v) for v in f.variables], type=type(f.fnc).__name__, prior=f.is_prior,
                             tag="pot"))

        def add_msg_prior(s, src):
            mp = G.MsgPrior(G.SlotRef(src, fg.variables[s].vartype.dim))
            inst.append(dict(variables=[s], fi=T.add_factor(mp, [slot_of(s)], None, 0.0, sp.inflation), mh=False,
                             rd=[slot_of(s), src], type="MsgPrior", prior=True, tag="common"))

        for ch in c.children:
            if not uml:
                for s in tree.cliques[ch].separators:       # addMsgFactors! (TreeMessageUtils.jl:566-575)
                    if s in c.allvars:
                        add_msg_prior(s, cslot[(ch, s)])     # the child's updated separator belief
                        n_msgs += 1
                continue
            msg = upmsg[ch]
            # addLikelihoodsDifferential! (TreeMessageUtils.jl:225-233): the relatives of the joint message
            for rel in msg["relatives"]:
                fnc = type(rel["sft"])(G.SlotRef(rel["slot"], rel["zdim"]))            # _sft(newBel), :321
                inst.append(dict(variables=list(rel["variables"]),
                                 fi=T.add_factor(fnc, [slot_of(v) for v in rel["variables"]], None, 0.0, sp.inflation),
                                 mh=False, rd=[slot_of(v) for v in rel["variables"]] + [rel["slot"]],
                                 type=rel["type"], prior=False, tag="diff"))
                n_msgs += 1
            # addLikelihoodPriorCommon! (TreeMessageUtils.jl:454-469)
            for lbl, src in msg["priors"]:
                if msg["hasPriors"] or not any(lbl in e["variables"] for e in inst):
                    add_msg_prior(lbl, src)
                    n_msgs += 1

        def propagate(v):
            fl, rd = [], []
            for e in inst:
                if v in e["variables"]:
                    fl.append((e["fi"], e["variables"].index(v) + 1, e["mh"]))
                    rd += e["rd"]
            if fl:
                add_prop(v, slot_of(v), fl, rd)

        def fmcmc(lbls, mciter):                              # SolveTree.jl:89-142
            if len(lbls) == 1:
                mciter = 1
            for _ in range(mciter):
                for v in lbls:
                    propagate(v)

        # upGibbsCliqueDensity (SolveTree.jl:193-235)
        fmcmc(c.directFrtlMsgIDs, 1)
        if c.msgskipIDs:
            fmcmc(c.msgskipIDs, 1)
        if c.itervarIDs:
            fmcmc(c.itervarIDs, iters)
        if c.directPriorMsgIDs:
            fmcmc([v for v in c.directPriorMsgIDs if v not in c.msgskipIDs], 1)
        if uml:
            kept_diffs[cid] = [e for e in inst if e["tag"] == "diff"]       # only UPWARD_COMMON is deleted (CSM :559-563)
            if c.parent is not None:
                upmsg[cid] = _joint_up_message(fg, c, inst, slot_of, T, N, deconvs, sched, reads, writes, opc, cur)
    n_up_ops = len(sched)

    # ---- down pass (parents before children); the root keeps its up-solve result
    if downsolve:
        for cid in reversed(post):
            c = tree.cliques[cid]
            if c.parent is None:
                continue
            cur[0] = cid
            p = c.parent
            # updateSubFgFromDownMsgs!: separators adopt the parent's down-message values
            for s in c.separators:
                add_copy(cslot[(p, s)], cslot[(cid, s)])
                n_msgs += 1

            # addDownVariableFactors!: every factor touching a frontal, with outside variables read
            # from the main graph (their graph-init beliefs, see DESIGN.md "down pass")
            def slot_dn(v, cid=cid, c=c):
                return cslot[(cid, v)] if v in c.allvars else var_slot[v]

            inst = []
            if uml:
                # the clique sub-graph as the up solve left it: potentials + the children's differentials; no
                # addDownVariableFactors! (CliqueStateMachine.jl:825-834).  The DOWNWARD_COMMON MsgPriors sit on
                # separators only and never enter a frontal's product.
                for fl in c.potentials:
                    f = fg.factors[fl]
                    inst.append((f.variables, fac_instance(f, slot_dn), G.isMultihypo(f),
                                 [slot_dn(v) for v in f.variables]))
                for e in kept_diffs.get(cid, []):
                    inst.append((e["variables"], e["fi"], False, e["rd"]))
            else:
                for f in _factors_touching(fg, by_var, c.frontals):
                    inst.append((f.variables, fac_instance(f, slot_dn), G.isMultihypo(f),
                                 [slot_dn(v) for v in f.variables]))

            def local_product(v):
                fl, rd = [], []
                for variables, fi, ismh, rds in inst:
                    if v in variables:
                        fl.append((fi, variables.index(v) + 1, ismh))
                        rd += rds
                if fl:
                    add_prop(v, cslot[(cid, v)], fl, rd)

            # determineCliqVariableDownSequence: frontals sharing a factor iterate MCIters times
            iterF = []
            for variables, _, _, _ in inst:
                fr = [v for v in variables if v in c.frontals]
                if len(fr) > 1:
                    iterF = _union(iterF, fr)
            iterF = [v for v in c.frontals if v in iterF]
            for v in c.frontals:
                if v not in iterF:
                    local_product(v)
            for _ in range(downIters):
                for v in iterF:
                    local_product(v)
    # ---- step 5: updateFromSubgraph — frontal beliefs go back to the main graph
    for c in tree.cliques:
        cur[0] = c.id
        for v in c.frontals:
            add_copy(cslot[(c.id, v)], var_slot[v])

    for k, dc in enumerate(deconvs):               # Philox call ids: props first (16 apart), then the deconvolutions
        dc["call_id"] = 16 * (len(props) + k)
    waves = _levelize(sched, reads, writes)
    nw = max(waves) + 1 if waves else 0
    order = sorted(range(len(sched)), key=lambda i: (waves[i], i))
    sched_waved = [sched[i] for i in order]
    wave_off = [0] * (nw + 1)
    for i in order:
        wave_off[waves[i] + 1] += 1
    for w in range(nw):
        wave_off[w + 1] += wave_off[w]
    up_last = max(waves[:n_up_ops]) + 1 if n_up_ops else 0
    wt = [float(len(props[a]["factors"]) + 1) if k == A.S_PROPAGATE else (1.0 if k == A.S_DECONV else 0.05)
          for (k, a, _) in sched]
    lane = assign_lanes(tree, opc, wt, waves, reads, writes, lanes)
    frozen = T.freeze()
    msgs = {cid: dict(relatives=[tuple(r["variables"]) for r in m["relatives"]], priors=[l for l, _ in m["priors"]],
                      hasPriors=m["hasPriors"]) for cid, m in upmsg.items()} if uml else None
    return SolvePlan(T, frozen, props, sched, wave_off, sched_waved, var_slot, nconv[0], len(props), n_msgs,
                     up_last, [opc[i] for i in order], [waves[i] for i in order], [reads[i] for i in order],
                     [writes[i] for i in order],
                     {**{s: cid for (cid, _), s in cslot.items()}, **{dc["out_slot"]: dc["clique"] for dc in deconvs}},
                     deconvs, [lane[i] for i in order], None, msgs)


def plan_call_span(plan: SolvePlan) -> int:
    """number of Philox call ids a plan uses (16 per propagateBelief / deconvolution)"""
    return 16 * (len(plan.props) + len(plan.deconvs or []))


def rebase_calls(plan: SolvePlan, base: int) -> None:
    """shift every call id of a freshly compiled plan (ids start at 0) by `base`"""
    if base:
        for p in plan.props:
            p["call_id"] += base
        for dc in plan.deconvs or []:
            dc["call_id"] += base


def _joint_up_message(fg, c: TreeClique, inst, slot_of, T, N, deconvs, sched, reads, writes, opc, cur):
    """prepCliqueMsgUp with useMsgLikelihoods (TreeMessageUtils.jl:667-703) ->
    _generateMsgJointRelativesPriors (:417-446): differentials between separator pairs
    (addLikelihoodsDifferentialCHILD!, :279-335), connected classes (_findSubgraphsFactorType, :118-205) and one
    candidate MsgPrior per class (_generateSubgraphMsgPriors / _calcCandidatePriorBest, :339-412).

    Each differential becomes one IIF_S_DECONV schedule op writing a fresh measurement slot."""
    seps = list(c.separators)
    dims = [fg.variables[s].vartype.dim for s in seps]
    dec = [seps[i] for i in sorted(range(len(seps)), key=lambda i: -dims[i])]      # sortperm(listDims; rev=true), stable
    acc = dec[::-1]
    relatives, already = [], []
    for s1 in dec:
        already.append(s1)
        for s2 in [v for v in acc if v not in already]:
            types = _shortest_path_factor_types(inst, s1, s2)                       # isPathFactorsHomogeneous (DFG)
            if not types or len(set(types)) != 1:
                continue
            sel = selectFactorType(fg.variables[s1].vartype, fg.variables[s2].vartype)
            if sel is None or sel[0] != types[0]:
                continue
            name, sft = sel
            vt = fg.variables[s1].vartype
            mslot = T.add_slot(vt, N)                                               # newBel = manikde!(sft, pts)
            dummy = T.add_factor(sft, [slot_of(s1), slot_of(s2)], None, 0.0, 5.0)   # tfg dummy factor, :314
            deconvs.append(dict(factor=dummy, out_slot=mslot, N=N, call_id=-1, clique=cur[0]))   # ids follow the props'
            sched.append((A.S_DECONV, len(deconvs) - 1, 0))
            reads.append(sorted({slot_of(s1), slot_of(s2)}))
            writes.append([mslot])
            opc.append(cur[0])
            relatives.append(dict(variables=[s1, s2], type=name, sft=sft, slot=mslot, zdim=vt.dim))
    # _findSubgraphsFactorType: classes of separators connected through the default relative type
    count = {s: 0 for s in seps}
    for r in relatives:
        for v in r["variables"]:
            count[v] += 1
    cls: Dict[str, int] = {}
    ncls = 0
    for s in seps:
        if count[s] == 0:
            ncls += 1
            cls[s] = ncls
    for k1 in [s for s in seps if s not in cls]:
        if k1 not in cls:
            ncls += 1
            cls[k1] = ncls
        for k2 in [s for s in seps if s not in cls]:
            sel = selectFactorType(fg.variables[k1].vartype, fg.variables[k2].vartype)
            pth = _shortest_path_factor_types(inst, k1, k2, sel[0]) if sel is not None else None
            if not pth:
                ncls += 1
                cls[k2] = ncls
            else:
                cls[k2] = cls[k1]
    classes: Dict[int, List[str]] = {}
    for s in seps:
        classes.setdefault(cls[s], []).append(s)
    pot_has_prior = any(e["prior"] for e in inst if e["tag"] == "pot")              # :431
    priors = []
    for syms in classes.values():
        if len(syms) == 1 or pot_has_prior:                                         # :402-408
            md = max(fg.variables[v].vartype.dim for v in syms)
            cand = [v for v in syms if fg.variables[v].vartype.dim == md]
            adj = [sum(1 for e in inst if v in e["variables"]) for v in cand]
            best = cand[max(range(len(cand)), key=lambda i: (adj[i], -i))]          # sortperm(mdAdj; rev=true)[1]
            priors.append((best, slot_of(best)))
    return dict(relatives=relatives, priors=priors, hasPriors=any(e["prior"] for e in inst))   # :681
