"""Reference-facing operator API of the hot path, mirrored in Python above the C-ABI.

Names, argument meaning and error behaviour follow the reference (`!` dropped):
  approxConvBelief / approxConv   src/services/ApproxConv.jl:4-47
  propagateBelief                 src/services/GraphProductOperations.jl:16-81
  localProduct / localProductAndUpdate   src/services/GraphProductOperations.jl:93-155
  doautoinit / initAll            src/services/GraphInit.jl:132-199, 495-556
  solveTree / solveGraph          src/services/SolverAPI.jl:326-446
Every numeric step is a kernel launch in libiifb200.so (Engine); nothing here computes beliefs.
"""
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _abi as A
from . import compile as CP
from . import graph as G
from . import tree as TR
from .engine import Engine


# --------------------------------------------------------------------------- per-graph engine
class GraphEngine:
    """Device mirror of one FactorGraph: slot i == variable i, factor table == graph factors.  ONE iifb200 context for
    the graph's lifetime: structural changes (addVariable / addFactor) re-issue iifb200_set_graph on it (the library's
    arena, tables and scratch are grow-only) instead of tearing the context down, and beliefs are re-uploaded in one
    transfer (incremental construction with graphinit used to re-create the CUDA context per factor)."""

    def __init__(self, fg: G.FactorGraph, device: int = 0, cap: int = 0):
        self.fg = fg
        self.device = device
        self.eng = None
        self.cap = 0
        self._build(cap)

    def _build(self, cap: int = 0):
        fg = self.fg
        self.version = fg._version
        T = CP.Tables()
        self.N = fg.solverParams.N
        # slot capacity is the engine's own business: a call that asks for more points than SolverParams.N grows it
        # without touching the user's solver parameters
        self.cap = max(self.N, int(cap), self.cap)
        self.var_slot = {}
        for l, v in fg.variables.items():
            self.var_slot[l] = T.add_slot(v.vartype, max(self.cap, v.val.shape[0], 1))
        self.fac_idx = {}
        for l, f in fg.factors.items():
            self.fac_idx[l] = T.add_factor(f.fnc, [self.var_slot[v] for v in f.variables], f.multihypo,
                                           f.nullhypo, f.inflation)
        self.frozen = T.freeze()
        self.sp_c = CP.solver_params_c(fg.solverParams)
        if self.eng is None:
            self.eng = Engine(self.frozen, self.sp_c, self.device)
        else:
            self.eng.reset_graph(self.frozen, self.sp_c)      # same context, grow-only device memory
        self.dirty = set(fg.variables)

    def mark_dirty(self, lbl):
        self.dirty.add(lbl)

    def flush(self):
        if not self.dirty:
            return
        if len(self.dirty) > 4:                                # many beliefs: one transfer of the whole arena
            ar = CP.HostArena(self.frozen)
            for l, v in self.fg.variables.items():
                ar.set(self.var_slot[l], v.val, v.bw, v.initialized, v.infoPerCoord)
            self.eng.upload_arena(ar)
        else:
            for l in list(self.dirty):
                v = self.fg.variables[l]
                self.eng.upload_belief(self.var_slot[l], v.val, v.bw, v.initialized)
        self.dirty.clear()

    def next_call(self, n=16):
        return self.fg.next_call(n)       # graph-wide counter: no two numeric calls on a graph share Philox streams

    def close(self):
        self.eng.close()


def _engine(fg: G.FactorGraph, cap: int = 0) -> GraphEngine:
    ge = fg._engine
    if ge is None:
        ge = GraphEngine(fg, cap=cap)
        fg._engine = ge
    elif ge.version != fg._version or ge.N != fg.solverParams.N or ge.cap < cap:
        ge._build(cap)
    sp_c = CP.solver_params_c(fg.solverParams)
    if bytes(sp_c) != bytes(ge.sp_c):
        ge.sp_c = sp_c
        ge.eng.set_solver_params(sp_c)
    ge.flush()
    return ge


def kde_bandwidth(fg: G.FactorGraph, vartype: G.InferenceVariable, pts) -> np.ndarray:
    """manikde!(M, pts) bandwidth selection (AMP; call sites ApproxConv.jl:38-41)."""
    return _engine(fg).eng.kde_bandwidth(np.asarray(pts, dtype=np.float64).reshape(-1, vartype.dim),
                                         vartype.circ_mask)


def manikde(fg: G.FactorGraph, vartype: G.InferenceVariable, pts, bw=None) -> G.ManifoldKernelDensity:
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(-1, vartype.dim))
    if bw is None:
        bw = kde_bandwidth(fg, vartype, pts)
    return G.ManifoldKernelDensity(vartype, pts, np.asarray(bw, dtype=np.float64))


# --------------------------------------------------------------------------- a3: approxConvBelief
def approxConvBelief(fg: G.FactorGraph, fct: str, target: str, measurement=None, N: Optional[int] = None,
                     nullSurplus: float = 0.0, mhidx=None, uinf=None, return_labels: bool = False):
    """approxConvBelief(dfg, fc, target; N, nullSurplus) — ApproxConv.jl:4-45.

    The target variable is NOT modified (ApproxConv.jl:17).  `measurement`, `mhidx`, `uinf` inject
    host-drawn random streams (Julia-drawn labels stay bit-exact); by default the device draws them.
    """
    f = fg.factors[fct]
    if target not in f.variables:
        raise KeyError(f"{target} is not a variable of factor {fct}")
    v = fg.variables[target]
    if N is None:
        N = len(measurement) if measurement is not None and len(measurement) else 0
    N = N if N else (v.val.shape[0] or fg.solverParams.N)     # ApproxConv.jl:15 picks a local N
    ge = _engine(fg, cap=N)                                    # grows the engine's slots, not SolverParams.N
    spec = dict(factor=ge.fac_idx[fct], sfidx=f.variables.index(target) + 1, N=N, call_id=ge.next_call(),
                nullSurplus=nullSurplus)
    meas = None
    if measurement is not None and len(measurement):
        meas = np.ascontiguousarray(np.asarray(measurement, dtype=np.float64).reshape(N, -1))
        spec["meas_off"] = 0
    if mhidx is not None:
        spec["mhidx_off"] = 0
    if uinf is not None:
        spec["uinf_off"] = 0
    ops = CP.make_conv_ops([spec])
    pts, bw, ipc, lab, nnan = ge.eng.conv_batch(ops, 1, meas, mhidx, uinf)[0]
    partial = None
    if np.any(np.abs(ipc) <= 1e-14):                            # ApproxConv.jl:31-42
        partial = [int(i) + 1 for i in np.nonzero(np.abs(ipc) > 1e-14)[0]]
    mkd = G.ManifoldKernelDensity(v.vartype, pts, bw, partial, ipc)
    return (mkd, lab) if return_labels else mkd


def approxConv(fg, fct, target, *a, **kw):
    """approxConv(w...) = getPoints(approxConvBelief(w...), false) — ApproxConv.jl:47."""
    return G.getPoints(approxConvBelief(fg, fct, target, *a, **kw))


# --------------------------------------------------------------------------- a1: propagateBelief
def propagateBelief(fg: G.FactorGraph, destlbl: str, factors=":", N: Optional[int] = None):
    """propagateBelief(dfg, destlbl, factors; N) -> (ManifoldKernelDensity, ipc) — GraphProductOperations.jl:16-81.
    `factors` is ':' (all neighbours) or a list of factor labels."""
    ge = _engine(fg)
    N = N or fg.solverParams.N
    flist = fg.listNeighbors(destlbl) if isinstance(factors, str) and factors == ":" else list(factors)
    if not flist:
        raise A.IIFB200Error(f"propagateBelief: no factors for {destlbl}")
    if len(flist) > A.IIF_MAX_FACTORS:
        raise A.IIFB200Error(f"propagateBelief: {len(flist)} factors exceed IIF_MAX_FACTORS")
    T = ge.frozen
    slot = ge.var_slot[destlbl]
    # posterior lands in the destination slot on the device, then is read back; the live variable
    # keeps its value on the host (setBelief! is the caller's decision, SolveTree.jl:74)
    spec = dict(target_slot=slot, out_slot=slot,
                factors=[(ge.fac_idx[fl], fg.factors[fl].variables.index(destlbl) + 1) for fl in flist],
                N=N, call_id=ge.next_call(), any_multihypo=int(any(G.isMultihypo(fg.factors[fl]) for fl in flist)))
    ge.eng.propagate_batch(CP.make_prop_ops([spec]), 1)
    pts, bw, ipc = ge.eng.download_belief(slot)
    ge.mark_dirty(destlbl)   # device slot now differs from the host variable: re-upload before next use
    del T
    v = fg.variables[destlbl]
    return G.ManifoldKernelDensity(v.vartype, pts, bw, None, ipc), ipc


def localProduct(fg: G.FactorGraph, sym: str, N: Optional[int] = None):
    """localProduct — GraphProductOperations.jl:93-120 -> (mkd, dens=None, lbls, ipc)."""
    lb = fg.listNeighbors(sym)
    mkd, ipc = propagateBelief(fg, sym, lb, N=N)
    return mkd, None, lb, ipc


def localProductAndUpdate(fg: G.FactorGraph, sym: str, setkde: bool = True):
    """localProductAndUpdate! — GraphProductOperations.jl:136-155."""
    mkd, _, lbl, ipc = localProduct(fg, sym)
    if setkde and G.Npts(mkd) > 0:
        G.setValKDE(fg, sym, mkd, False, ipc)
    return mkd, ipc, lbl


# --------------------------------------------------------------------------- §8f-2: deconvolution
def approxDeconv(fg: G.FactorGraph, fctsym: str, N: Optional[int] = None):
    """approxDeconv(dfg, fctsym) — DeconvUtils.jl:32-202: (predicted, sampled) measurement values of a factor,
    N = number of points of its first variable (:194-196)."""
    ge = _engine(fg)
    f = fg.factors[fctsym]
    if G.isMultihypo(f):
        raise A.IIFB200Error("approxDeconv: multihypo factors are not supported (as in the reference, #1096)")
    N = N or fg.variables[f.variables[0]].val.shape[0] or fg.solverParams.N
    return ge.eng.deconv(ge.fac_idx[fctsym], N, ge.next_call())


def mmd(fg: G.FactorGraph, p1, p2, vartype: G.InferenceVariable = None, bw: float = 0.001) -> float:
    """mmd(p1, p2, varType; bw=[0.001]) — SolverUtilities.jl:25-47 (AMP.mmd!)."""
    return _engine(fg).eng.mmd(p1, p2, vartype.circ_mask if vartype is not None else 0, bw)


# --------------------------------------------------------------------------- §8f-3: point estimates
def calcPPE(fg: G.FactorGraph, label: str, solveKey: str = "default") -> G.MeanMaxPPE:
    """calcPPE(dfg, label) — FGOSUtils.jl:237-296: suggested = mean = calcMean(P), max = getKDEMax(P)."""
    return calcPPEs(fg, [label], solveKey)[0]


def calcPPEs(fg: G.FactorGraph, labels: Sequence[str], solveKey: str = "default") -> List[G.MeanMaxPPE]:
    """calcPPE of several variables in one kernel launch (one CTA per belief)."""
    ge = _engine(fg)
    for l in labels:
        if fg.variables[l].val.shape[0] < 1:
            raise A.IIFB200Error(f"calcPPE: variable {l} has no belief points")
    mean, mx = ge.eng.ppe_batch([ge.var_slot[l] for l in labels])
    out = []
    for k, l in enumerate(labels):
        d = fg.variables[l].vartype.dim
        out.append(G.MeanMaxPPE(solveKey, mean[k, :d].copy(), mx[k, :d].copy(), mean[k, :d].copy()))
    return out


def setPPE(fg: G.FactorGraph, label: str, solveKey: str = "default") -> G.MeanMaxPPE:
    """setPPE!(dfg, label) — FGOSUtils.jl:546-570: calcPPE and store it in the variable's ppeDict."""
    ppe = calcPPE(fg, label, solveKey)
    fg.variables[label].ppeDict[solveKey] = ppe
    return ppe


def getPPE(fg: G.FactorGraph, label: str, solveKey: str = "default") -> G.MeanMaxPPE:
    return fg.variables[label].ppeDict[solveKey]


# --------------------------------------------------------------------------- §8f-1: graph init
factorCanInitFromOtherVars = G.factorCanInitFromOtherVars      # GraphInit.jl:62-116 incl. the multihypo carve-out


def doautoinit(fg: G.FactorGraph, lbl: str, singles: bool = True) -> bool:
    """doautoinit! — GraphInit.jl:132-199."""
    v = fg.variables[lbl]
    if v.initialized:
        return False
    nei = fg.listNeighbors(lbl)
    if not (singles or len(nei) > 1):
        return False
    use = [f for f in nei if factorCanInitFromOtherVars(fg, f, lbl)]
    if not use:
        return False
    mkd, ipc = propagateBelief(fg, lbl, use)                   # raises beyond IIF_MAX_FACTORS, never truncates
    G.setValKDE(fg, lbl, mkd, True, ipc)
    return True


def initAll(fg: G.FactorGraph, batched: bool = True) -> None:
    """initAll! — GraphInit.jl:495-556: sweep until nothing new can be initialised.

    `batched`: the sequential sweep of doautoinit! calls is replayed on the
    initialisation flags only, every propagateBelief it would run is recorded in order, the list is
    levelised by slot hazards into waves of independent initialisations and executed as ONE schedule
    (CUDA graph) on the device-resident graph; beliefs come back to the host once at the end.  Order,
    factor selection and Philox call ids are those of the sequential sweep."""
    if not batched:
        for _ in range(len(fg.variables) + 1):
            did = False
            for l in fg.variables:
                did |= doautoinit(fg, l)
            if not did:
                break
        return
    ge = _engine(fg)
    N = fg.solverParams.N
    init = {l: v.initialized for l, v in fg.variables.items()}
    specs, order, reads, writes = [], [], [], []
    for _ in range(len(fg.variables) + 1):
        did = False
        for l in fg.variables:
            if init[l]:
                continue
            use = [f for f in fg.listNeighbors(l) if G.factorCanInitFromOtherVars(fg, f, l, init)]
            if not use:
                continue
            if len(use) > A.IIF_MAX_FACTORS:
                raise A.IIFB200Error(f"initAll: {len(use)} factors on {l} exceed IIF_MAX_FACTORS = {A.IIF_MAX_FACTORS}")
            slot = ge.var_slot[l]
            specs.append(dict(target_slot=slot, out_slot=slot,
                              factors=[(ge.fac_idx[f], fg.factors[f].variables.index(l) + 1) for f in use],
                              N=N, call_id=ge.next_call(),
                              any_multihypo=int(any(G.isMultihypo(fg.factors[f]) for f in use))))
            order.append(l)
            reads.append(sorted({ge.var_slot[v] for f in use for v in fg.factors[f].variables}))
            writes.append([slot])
            init[l] = True
            did = True
        if not did:
            break
    if not specs:
        return
    sched = [(A.S_PROPAGATE, k, 0) for k in range(len(specs))]
    waves = TR._levelize(sched, reads, writes)
    idx = sorted(range(len(sched)), key=lambda i: (waves[i], i))
    nw = max(waves) + 1
    wave_off = [0] * (nw + 1)
    for i in idx:
        wave_off[waves[i] + 1] += 1
    for w in range(nw):
        wave_off[w + 1] += wave_off[w]
    sid = ge.eng.schedule_build(wave_off, CP.make_sched_ops([sched[i] for i in idx]), len(sched),
                                CP.make_prop_ops(specs), len(specs))
    ge.eng.schedule_run(sid)
    ge.eng.sync()
    for l in order:
        pts, bw, ipc = ge.eng.download_belief(ge.var_slot[l])
        v = fg.variables[l]
        v.val, v.bw, v.infoPerCoord, v.initialized = pts, bw, ipc, True      # setValKDE!; the device copy is current
    ge.eng.lib.iifb200_schedule_free(ge.eng.ctx, sid)


# --------------------------------------------------------------------------- solveTree!
class TreeSolver:
    """Compiled solveTree!: one device arena with clique-local slots + one CUDA-graph schedule."""

    def __init__(self, fg: G.FactorGraph, eliminationOrder: Optional[Sequence[str]] = None, ordering: str = "qr",
                 device: int = 0, ext_arena_ptr=None, downsolve: Optional[bool] = None, lanes: Optional[int] = None,
                 forward_copies: bool = True, call_base=0, planner: Optional[str] = None):
        """`call_base`: first Philox call id of the plan.  0 (default) gives the same streams for the same seed —
        what parity tests and the bench want; "auto" reserves a fresh range from the graph's counter, so consecutive
        solveTree calls on one graph see independent noise."""
        self.fg = fg
        order = list(eliminationOrder) if eliminationOrder is not None else TR.getEliminationOrder(fg, ordering)
        self.tree = TR.buildTree(fg, order)
        ds = fg.solverParams.downsolve if downsolve is None else downsolve
        # independent sub-trees become parallel branches ("lanes") of the captured CUDA graph (tree.assign_lanes)
        lanes = int(os.environ.get("IIFB200_LANES", "4")) if lanes is None else lanes
        # `planner`: "c" = iifb200_plan_tree inside libiifb200.so (what a Julia caller uses: the clique table goes in,
        # the schedule comes out), "python" = the mirror in tree.py.  Both give the same plan bit for bit, differential
        # (useMsgLikelihoods = true) messages included (tests/test_plan_abi.py); default: the library's planner.
        uml = bool(fg.solverParams.useMsgLikelihoods)
        planner = planner or os.environ.get("IIFB200_PLANNER") or "c"
        self.sp_c = CP.solver_params_c(fg.solverParams)
        self.plan_handle = None
        if planner == "c":
            from . import planner as PL
            self.plan = PL.plan_tree(fg, self.tree, downsolve=ds, lanes=lanes, forward_copies=forward_copies,
                                     useMsgLikelihoods=uml)
            if call_base == "auto":
                call_base = fg.next_call(TR.plan_call_span(self.plan))
            if call_base:       # ids are baked into the library's plan: plan again with the base
                self.plan = PL.plan_tree(fg, self.tree, downsolve=ds, lanes=lanes, forward_copies=forward_copies,
                                         useMsgLikelihoods=uml, call_base=int(call_base))
            self.plan_handle = self.plan.c_plan
            self.eng, self.sid = Engine.from_plan(self.plan.frozen, self.plan_handle.handle, self.sp_c, device, ext_arena_ptr)
        else:
            self.plan = TR.compile_solve(fg, self.tree, downsolve=ds, lanes=lanes, forward_copies=forward_copies)
            if call_base == "auto":
                call_base = fg.next_call(TR.plan_call_span(self.plan))
            TR.rebase_calls(self.plan, int(call_base))
            self.eng = Engine(self.plan.frozen, self.sp_c, device, ext_arena_ptr)
            self.props_c = CP.make_prop_ops(self.plan.props)
            self.sched_c = CP.make_sched_ops(self.plan.sched_waved, self.plan.op_lane)
            self.deconvs_c = CP.make_deconv_ops(self.plan.deconvs or [])
            self.sid = self.eng.schedule_build(self.plan.wave_off, self.sched_c, len(self.plan.sched_waved),
                                               self.props_c, len(self.plan.props), self.deconvs_c,
                                               len(self.plan.deconvs or []))
        self.arena = CP.HostArena(self.plan.frozen)

    def load_from_graph(self):
        for l, v in self.fg.variables.items():
            self.arena.set(self.plan.var_slot[l], v.val, v.bw, v.initialized, v.infoPerCoord)

    def upload(self):
        self.eng.upload_arena(self.arena)

    def run(self, first=0, last=-1):
        self.eng.schedule_run(self.sid, first, last)

    def download(self):
        self.eng.sync()
        self.eng.download_arena(self.arena)

    def store_to_graph(self, ppe: bool = True):
        """updateFromSubgraph (CSM step 5, CliqueStateMachine.jl:928-966): beliefs and point estimates go back
        to the graph; the PPEs of all variables come from one iifb200_ppe_batch launch on the device-resident
        posteriors."""
        labels = list(self.fg.variables)
        for l in labels:
            pts, bw, ipc = self.arena.get(self.plan.var_slot[l])
            G.setValKDE(self.fg, l, G.ManifoldKernelDensity(self.fg.variables[l].vartype, pts, bw), True, ipc)
        if ppe:
            mean, mx = self.eng.ppe_batch([self.plan.var_slot[l] for l in labels])
            for k, l in enumerate(labels):
                d = self.fg.variables[l].vartype.dim
                self.fg.variables[l].ppeDict["default"] = G.MeanMaxPPE("default", mean[k, :d].copy(), mx[k, :d].copy(),
                                                                       mean[k, :d].copy())

    def close(self):
        self.eng.close()


def solveTree(fg: G.FactorGraph, eliminationOrder: Optional[Sequence[str]] = None, ordering: str = "qr",
              device: int = 0):
    """solveTree!(fg; eliminationOrder) — SolverAPI.jl:326-446 (hot-path subset: graph init, tree build,
    up + down clique solves, write-back).  Returns the TreeSolver (holds the tree and the plan)."""
    if fg.solverParams.graphinit:
        initAll(fg)
    for l, v in fg.variables.items():
        if not v.initialized:
            raise A.IIFB200Error(f"solveTree: variable {l} could not be initialised")
    ts = TreeSolver(fg, eliminationOrder, ordering, device, call_base="auto")
    ts.load_from_graph()
    ts.upload()
    ts.run()
    ts.download()
    ts.store_to_graph()
    return ts


solveGraph = solveTree


# --------------------------------------------------------------------------- boundary B3 driver
class B3Driver:
    """Drives a compiled plan ONE propagateBelief at a time, with the call sequence julia/IIFB200.jl's
    `propagateBelief` makes per belief update (boundary B3): iifb200_propagate_once on the mini graph {destination, its
    factors' variables, message beliefs} — or, with fused = False, its four constituents: iifb200_set_graph, ONE
    iifb200_upload_slots of their beliefs, iifb200_propagate_batch(1) and iifb200_download_belief of the posterior.  Beliefs live on the host between calls (as they do in the DFG's
    VariableNodeData); separator copies are host copies.  Same Philox call ids as the plan => same posteriors as the
    one-schedule (B4) run; what differs is the cost of 7 000 round trips instead of one.

    `contexts` > 1 mirrors the shim's context pool: the reference solves sibling cliques as concurrent Tasks
    (SolverAPI.jl:59-96), so independent propagateBelief calls — here: the ops of one wave — are in flight at once,
    each on its own library context (own stream, arena and scratch) from its own host thread."""

    def __init__(self, plan: TR.SolvePlan, sp_c, device: int = 0, contexts: int = 1, fused: bool = True):
        self.plan, self.sp_c, self.fused = plan, sp_c, fused
        fz = plan.frozen
        self.calls = []
        for spec in plan.props:
            order, loc = [], {}

            def use(s):
                if s not in loc:
                    loc[s] = len(order)
                    order.append(s)
                return loc[s]
            use(spec["target_slot"])
            T = CP.Tables()
            facs = []
            for fi, sf in spec["factors"]:
                f = fz["factors"][fi]
                D = fz["dists"][f.dist]
                sl = [use(f.slot[k]) for k in range(f.arity)]
                ks = use(D.slot) if D.kind == A.D_KDE else -1
                facs.append((f, D, sl, ks, sf))
            ns = len(order)
            slots = (A.SlotDesc * ns)()
            off = 0
            for i, s in enumerate(order):
                g = fz["slots"][s]
                slots[i].dim, slots[i].circ_mask, slots[i].cap, slots[i].pts_off = g.dim, g.circ_mask, g.cap, off
                off += g.dim * g.cap
            factors = (A.FactorDesc * len(facs))()
            dists = (A.DistDesc * len(facs))()
            for k, (f, D, sl, ks, sf) in enumerate(facs):
                C_ = factors[k]
                C_.kind, C_.arity, C_.zdim, C_.dist, C_.nmh, C_.partial_mask = f.kind, f.arity, f.zdim, k, f.nmh, f.partial_mask
                C_.solver = f.solver
                for i in range(A.IIF_MAX_DIM):
                    C_.aux[i] = f.aux[i]
                C_.nullhypo, C_.inflation = f.nullhypo, f.inflation
                for i, s in enumerate(sl):
                    C_.slot[i] = s
                for i in range(f.nmh):
                    C_.mh[i] = f.mh[i]
                dists[k].kind, dists[k].dim, dists[k].ncomp, dists[k].comp_kind = D.kind, D.dim, D.ncomp, D.comp_kind
                dists[k].slot, dists[k].poff = ks, D.poff                 # parameter block shared with the full plan
            mini = dict(nslots=ns, slots=slots, nfactors=len(facs), factors=factors, ndists=len(facs), dists=dists,
                        nparams=fz["nparams"], dparams=fz["dparams"], total_doubles=off)
            op = CP.make_prop_ops([dict(target_slot=0, out_slot=0, factors=[(k, sf) for k, (_, _, _, _, sf) in enumerate(facs)],
                                        N=spec["N"], call_id=spec["call_id"], any_multihypo=spec["any_multihypo"])])
            stage = (np.zeros(off), np.zeros(ns * A.IIF_MAX_DIM), np.zeros(ns, dtype=np.int32), np.ones(ns, dtype=np.int32))
            self.calls.append((order, mini, op, stage))
        self.engs = [Engine(self.calls[0][1], sp_c, device) for _ in range(max(1, contexts))] if self.calls else []
        self.eng = self.engs[0] if self.engs else None
        self.pool = None
        if len(self.engs) > 1:
            import concurrent.futures
            import queue
            self.pool = concurrent.futures.ThreadPoolExecutor(len(self.engs))
            self.free = queue.SimpleQueue()
            for e in self.engs:
                self.free.put(e)

    def run(self, arena: CP.HostArena):
        """one pass over the plan's ops in wave order, beliefs in `arena` (host)"""
        fz = self.plan.frozen
        sched, wo = self.plan.sched_waved, self.plan.wave_off
        for w in range(len(wo) - 1):
            props = []
            for kind, a, b in sched[wo[w]:wo[w + 1]]:
                if kind == A.S_COPY:
                    sa, sb = fz["slots"][a], fz["slots"][b]
                    n = int(arena.npts[a])
                    arena.pts[sb.pts_off:sb.pts_off + n * sa.dim] = arena.pts[sa.pts_off:sa.pts_off + n * sa.dim]
                    arena.bw[b * 4:b * 4 + 4] = arena.bw[a * 4:a * 4 + 4]
                    arena.ipc[b * 4:b * 4 + 4] = arena.ipc[a * 4:a * 4 + 4]
                    arena.npts[b], arena.flags[b] = n, arena.flags[a]
                elif kind == A.S_PROPAGATE:
                    props.append(a)
                else:
                    raise A.IIFB200Error("B3Driver: only PROPAGATE / COPY ops (useMsgLikelihoods = false plans)")
            # the ops of one wave touch disjoint destinations and read nothing another op of the wave writes
            if self.pool is None:
                for a in props:
                    self._call(self.eng, arena, a)
            else:
                def job(a):
                    eng = self.free.get()
                    try:
                        self._call(eng, arena, a)
                    finally:
                        self.free.put(eng)
                for f in [self.pool.submit(job, a) for a in props]:
                    f.result()

    def _call(self, eng, arena, a):
        fz = self.plan.frozen
        order, mini, op, (pts, bw, npts, flags) = self.calls[a]
        for i, s in enumerate(order):
            g, m = fz["slots"][s], mini["slots"][i]
            pts[m.pts_off:m.pts_off + g.cap * g.dim] = arena.pts[g.pts_off:g.pts_off + g.cap * g.dim]
            bw[i * 4:i * 4 + 4] = arena.bw[s * 4:s * 4 + 4]
            npts[i], flags[i] = arena.npts[s], arena.flags[s]
        if self.fused:
            p, w, ipc = eng.propagate_once(mini, pts, bw, npts, flags, op)   # all four steps behind one C-ABI call
        else:
            eng.reset_graph(mini)                                          # iifb200_set_graph
            eng.upload_slots(0, len(order), pts, bw, npts, flags)         # ONE transfer
            eng.propagate_batch(op, 1)
            p, w, ipc = eng.download_belief(0)
        t = order[0]
        g = fz["slots"][t]
        arena.pts[g.pts_off:g.pts_off + p.size] = p.reshape(-1)
        arena.bw[t * 4:t * 4 + g.dim], arena.ipc[t * 4:t * 4 + g.dim] = w, ipc
        arena.npts[t], arena.flags[t] = p.shape[0], 1

    def close(self):
        if self.pool is not None:
            self.pool.shutdown()
        for e in self.engs:
            e.close()
