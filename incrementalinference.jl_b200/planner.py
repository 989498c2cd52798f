"""Python binding of the tree planner in libiifb200.so (iifb200_plan_tree): packs the factor graph's descriptor
tables and the Bayes tree's per-clique lists into the C structs of include/iifb200.h — exactly what the Julia shim
does with `getCliqueData(cliq)` — and wraps the returned plan as a tree.SolvePlan."""
import ctypes as C

import numpy as np

from . import _abi as A
from . import compile as CP
from . import graph as G
from . import tree as TR


def _csr(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int32)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    flat = np.asarray([x for l in lists for x in l] or [0], dtype=np.int32)
    return off, flat


def graph_tables(fg: G.FactorGraph, N: int):
    """iif_graph_desc content: slot v == variable v, factor table in graph order"""
    T = CP.Tables()
    var_idx = {}
    for l, v in fg.variables.items():
        var_idx[l] = T.add_slot(v.vartype, max(N, v.val.shape[0], 1))
    fac_idx = {}
    for l, f in fg.factors.items():
        fac_idx[l] = T.add_factor(f.fnc, [var_idx[v] for v in f.variables], f.multihypo, f.nullhypo, f.inflation)
    return T.freeze(), var_idx, fac_idx


class CPlan:
    """owner of an iifb200_plan handle"""

    def __init__(self, handle, lib):
        self.handle, self.lib = handle, lib

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.iifb200_plan_free(self.handle)
            self.handle = None


def plan_tree(fg: G.FactorGraph, tree: TR.BayesTree, N=None, downsolve=True, gibbsIters=None, downIters=3,
              useMsgLikelihoods=False, lanes=0, forward_copies=True, call_base=0) -> TR.SolvePlan:
    lib = A.load_library()
    sp = fg.solverParams
    N = N or sp.N
    frozen, var_idx, fac_idx = graph_tables(fg, min(N, A.IIF_MAX_POINTS))
    # type tables of iif_graph_desc: ids stand for Julia types; the joint-message rules compare them for equality only
    tid = {}

    def type_id(name):
        return tid.setdefault(name, len(tid))
    ftype = np.asarray([type_id(type(f.fnc).__name__) for f in fg.factors.values()] or [0], dtype=np.int32)
    vtype = np.asarray([type_id("var:" + v.vartype.name) for v in fg.variables.values()], dtype=np.int32)
    vrelkind = np.zeros(len(vtype), dtype=np.int32)
    vreltype = np.full(len(vtype), -1, dtype=np.int32)
    for i, v in enumerate(fg.variables.values()):
        sel = TR.selectFactorType(v.vartype, v.vartype)      # DefaultNodeTypes.jl:12-31
        if sel is not None:
            vrelkind[i], vreltype[i] = sel[1].kind, type_id(sel[0])
    keep = [ftype, vtype, vrelkind, vreltype]
    gd = A.GraphDesc(frozen["nslots"], frozen["slots"], frozen["nfactors"], frozen["factors"], frozen["ndists"],
                     frozen["dists"], frozen["nparams"], A.as_dp(frozen["dparams"]), A.as_ip(ftype), A.as_ip(vtype),
                     A.as_ip(vrelkind), A.as_ip(vreltype), type_id("MsgPrior"), 0)
    cl = tree.cliques

    def lists(get, idx):
        off, flat = _csr([[idx[x] for x in get(c)] for c in cl])
        keep.extend((off, flat))
        return A.as_ip(off), A.as_ip(flat)
    parent = np.asarray([-1 if c.parent is None else c.parent for c in cl], dtype=np.int32)
    td = A.TreeDesc()
    td.ncliques, td.parent = len(cl), A.as_ip(parent)
    td.frontal_off, td.frontals = lists(lambda c: c.frontals, var_idx)
    td.separator_off, td.separators = lists(lambda c: c.separators, var_idx)
    td.potential_off, td.potentials = lists(lambda c: c.potentials, fac_idx)
    td.directFrtlMsg_off, td.directFrtlMsg = lists(lambda c: c.directFrtlMsgIDs, var_idx)
    td.msgskip_off, td.msgskip = lists(lambda c: c.msgskipIDs, var_idx)
    td.itervar_off, td.itervar = lists(lambda c: c.itervarIDs, var_idx)
    td.directPriorMsg_off, td.directPriorMsg = lists(lambda c: c.directPriorMsgIDs, var_idx)
    po = A.PlanOpts(int(N), int(gibbsIters or sp.gibbsIters), int(downIters), int(bool(downsolve)), int(lanes),
                    int(bool(forward_copies)), int(bool(useMsgLikelihoods)), int(call_base), float(sp.inflation))
    h = C.c_void_p()
    st = lib.iifb200_plan_tree(C.byref(gd), C.byref(td), C.byref(po), C.byref(h))
    if st != A.IIF_OK:
        raise A.IIFB200Error(f"iifb200_plan_tree failed ({st}): {lib.iifb200_plan_error().decode()}")
    cnt = np.zeros(16, dtype=np.int32)
    lib.iifb200_plan_counts(h, A.as_ip(cnt))
    ns, nf, nd, npar, nprops, nops, nw, n_conv, n_prod, n_msgs, up_last, nvars, ndec = (int(x) for x in cnt[:13])
    slots = (A.SlotDesc * max(ns, 1))()
    factors = (A.FactorDesc * max(nf, 1))()
    dists = (A.DistDesc * max(nd, 1))()
    dparams = np.zeros(max(npar, 1))
    props = (A.PropOp * max(nprops, 1))()
    ops = (A.SchedOp * max(nops, 1))()
    wave_off = np.zeros(nw + 1, dtype=np.int32)
    lib.iifb200_plan_export(h, slots, factors, dists, A.as_dp(dparams), props, ops, A.as_ip(wave_off))
    dcv = (A.DeconvOp * max(ndec, 1))()
    if ndec:
        lib.iifb200_plan_export_deconvs(h, dcv)
    deconvs = [dict(factor=d.factor, out_slot=d.out_slot, N=d.N, call_id=d.call_id) for d in dcv[:ndec]]
    total = slots[ns - 1].pts_off + slots[ns - 1].cap * slots[ns - 1].dim
    fz = dict(nslots=ns, slots=slots, nfactors=nf, factors=factors, ndists=nd, dists=dists, nparams=npar,
              dparams=dparams, total_doubles=total)
    pspecs = [dict(target_slot=p.target_slot, out_slot=p.out_slot,
                   factors=[(p.factor[k], p.sfidx[k]) for k in range(p.nfactors)], N=p.N, call_id=p.call_id,
                   any_multihypo=p.any_multihypo) for p in props[:nprops]]
    sched = [(o.kind, o.a, o.b) for o in ops[:nops]]
    plan = TR.SolvePlan(None, fz, pspecs, sched, [int(x) for x in wave_off], sched, dict(var_idx), n_conv, n_prod,
                        n_msgs, up_last, None, None, None, None, None, deconvs, [o.lane for o in ops[:nops]])
    plan.c_plan = CPlan(h, lib)
    return plan


def _order(fg: G.FactorGraph, call):
    lib = A.load_library()
    labels = list(fg.variables)
    idx = {l: i for i, l in enumerate(labels)}
    off, flat = _csr([[idx[v] for v in f.variables] for f in fg.factors.values()])
    out = np.zeros(max(len(labels), 1), dtype=np.int32)
    st = call(lib, len(labels), len(fg.factors), A.as_ip(off), A.as_ip(flat), A.as_ip(out))
    if st != A.IIF_OK:
        raise A.IIFB200Error(f"elimination order failed ({st}): {lib.iifb200_plan_error().decode()}")
    return [labels[i] for i in out[:len(labels)]]


def elimination_order_nd(fg: G.FactorGraph):
    """level-set bisection order from the library (iifb200_elimination_order_nd)"""
    return _order(fg, lambda lib, nv, nf, off, flat, out: lib.iifb200_elimination_order_nd(nv, nf, off, flat, out))


def elimination_order_is(fg: G.FactorGraph, slack: int = 1):
    """independent-set (generalised odd-even reduction) order from the library (iifb200_elimination_order_is): what a
    Julia caller passes as `solveTree!(fg; eliminationOrder = ...)` to get a bushy Bayes tree instead of the default
    order's path"""
    return _order(fg, lambda lib, nv, nf, off, flat, out: lib.iifb200_elimination_order_is(nv, nf, off, flat, slack, out))
