"""The tree planner inside libiifb200.so (iifb200_plan_tree, include/iifb200.h): a raw clique table — what a Julia
caller reads off `tree.bt` — goes through the C-ABI and must reproduce the Python mirror's plan (tree.compile_solve)
bit for bit: slots, factor instances, props (targets, factor lists, solve-for indices, Philox call ids, multihypo
flags), schedule ops, waves and lanes.  No GPU needed: the planner is host code."""
import ctypes as C

import numpy as np
import pytest

import parity_cases as PC
from iifb200 import _abi as A
from iifb200 import compile as CP
from iifb200 import planner as PL
from iifb200 import tree as TR
from iifb200 import workloads as W
from test_tree_parity import three_door_graph


def graphs():
    yield "chain_nd", W.scalar_chain(37, N=32, seed=1), W.chain_nd_order(37), {}
    yield "chain_natural", W.scalar_chain(9, N=16, seed=1), [f"x{k}" for k in range(9)], {}
    fg = W.euclid2_grid(rows=4, cols=7, N=24, seed=2, closure_every=2)
    yield "grid_nd", fg, TR.getEliminationOrder(fg, "nd"), {}
    fg = W.generateGraph_Kaess(N=20)
    yield "kaess_qr", fg, TR.getEliminationOrder(fg, "qr"), {}
    fg = W.generateGraph_CaesarRing1D(N=20)
    yield "caesar_ring_no_down", fg, TR.getEliminationOrder(fg, "qr"), dict(downsolve=False)
    fg = W.scalar_chain_sessions(3, 12, N=16)
    yield "forest", fg, W.sessions_nd_order(3, 12), {}
    fg = three_door_graph(1, poses=3, N=50)
    yield "multihypo", fg, TR.getEliminationOrder(fg, "qr"), {}
    fg = W.circular_chain(n=15, N=40)
    yield "circular_no_forwarding", fg, W.chain_nd_order(15), dict(forward_copies=False)
    fg = W.four_door(N=64)
    yield "mixture_priors", fg, TR.getEliminationOrder(fg, "qr"), dict(gibbsIters=5)
    # useMsgLikelihoods = true (fourdoortest.jl:19, testCircular.jl:12): differential separator messages
    yield "uml_chain", W.scalar_chain(37, N=32, seed=1), W.chain_nd_order(37), dict(useMsgLikelihoods=True)
    fg = W.four_door(N=64)
    yield "uml_four_door", fg, TR.getEliminationOrder(fg, "qr"), dict(useMsgLikelihoods=True)
    fg = W.circular_chain(n=15, N=40)
    yield "uml_circular", fg, W.chain_nd_order(15), dict(useMsgLikelihoods=True)
    fg = W.euclid2_grid(rows=4, cols=7, N=24, seed=2, closure_every=2)
    yield "uml_grid", fg, TR.getEliminationOrder(fg, "nd"), dict(useMsgLikelihoods=True)
    fg = W.generateGraph_Kaess(N=20)
    yield "uml_kaess", fg, TR.getEliminationOrder(fg, "qr"), dict(useMsgLikelihoods=True, downsolve=False)
    fg = three_door_graph(1, poses=3, N=50)
    yield "uml_multihypo", fg, TR.getEliminationOrder(fg, "qr"), dict(useMsgLikelihoods=True)
    fg = W.generateGraph_CaesarRing1D(N=20)
    yield "uml_caesar_ring", fg, TR.getEliminationOrder(fg, "qr"), dict(useMsgLikelihoods=True)


def _factor_tuple(frozen, i):
    f = frozen["factors"][i]
    D = frozen["dists"][f.dist]
    prm = frozen["dparams"]
    if D.kind == A.D_KDE:
        dist = ("kde", D.dim, D.slot)
    elif D.kind == A.D_MIXTURE:
        blk = 2 if D.comp_kind != A.D_MVNORMAL else D.dim + D.dim * D.dim
        dist = ("mix", D.dim, D.ncomp, D.comp_kind, tuple(prm[D.poff:D.poff + D.ncomp * (1 + blk)]))
    else:
        n = {A.D_NORMAL: 2, A.D_UNIFORM: 2}.get(D.kind, D.dim + D.dim * D.dim)
        dist = (D.kind, D.dim, tuple(prm[D.poff:D.poff + n]))
    return (f.kind, f.arity, f.zdim, tuple(f.slot[k] for k in range(f.arity)), f.nmh, f.partial_mask, f.solver,
            tuple(f.mh[k] for k in range(f.nmh)), f.nullhypo, f.inflation, tuple(f.aux), dist)


@pytest.mark.parametrize("case", list(graphs()), ids=lambda c: c[0])
def test_c_planner_reproduces_python_plan(built, case):
    name, fg, order, kw = case
    tree = TR.buildTree(fg, order)
    for lanes in (0, 4):
        kw = dict(kw)
        kw.setdefault("useMsgLikelihoods", False)
        ref = TR.compile_solve(fg, tree, lanes=lanes, **kw)
        got = PL.plan_tree(fg, tree, lanes=lanes, **kw)
        a, b = ref.frozen, got.frozen
        assert a["nslots"] == b["nslots"] and a["nfactors"] == b["nfactors"]
        for i in range(a["nslots"]):
            sa, sb = a["slots"][i], b["slots"][i]
            assert (sa.dim, sa.circ_mask, sa.cap, sa.pts_off) == (sb.dim, sb.circ_mask, sb.cap, sb.pts_off), i
        for i in range(a["nfactors"]):
            assert _factor_tuple(a, i) == _factor_tuple(b, i), (name, i)
        assert ref.props == got.props
        assert [{k: d[k] for k in ("factor", "out_slot", "N", "call_id")} for d in ref.deconvs or []] == got.deconvs
        if name in ("uml_grid", "uml_caesar_ring"):      # multi-variable separators: differentials exist
            assert got.deconvs, name
        if not kw["useMsgLikelihoods"]:
            assert not got.deconvs
        assert ref.sched_waved == got.sched_waved
        assert list(ref.wave_off) == list(got.wave_off)
        assert list(ref.op_lane) == list(got.op_lane)
        assert (ref.n_conv, ref.n_prod, ref.n_msgs, ref.up_last_wave) == (got.n_conv, got.n_prod, got.n_msgs, got.up_last_wave)
        assert ref.var_slot == got.var_slot


def test_c_planner_runs_on_the_oracle(built):
    """a plan made by the library drives the oracle to the same posteriors as the Python-made plan"""
    import oracle as O
    fg, order = W.scalar_chain(21, N=32, seed=4), W.chain_nd_order(21)
    tree = TR.buildTree(fg, order)
    res = []
    for plan in (TR.compile_solve(fg, tree, lanes=4), PL.plan_tree(fg, tree, lanes=4)):
        ar = CP.HostArena(plan.frozen)
        for l, v in fg.variables.items():
            ar.set(plan.var_slot[l], v.val, v.bw, True)
        orc = O.Oracle(plan.frozen, ar, CP.solver_params_c(fg.solverParams))
        orc.schedule_run(plan.wave_off, CP.make_sched_ops(plan.sched_waved), CP.make_prop_ops(plan.props))
        res.append(ar)
    assert np.array_equal(res[0].pts, res[1].pts) and np.array_equal(res[0].bw, res[1].bw)


def test_c_planner_errors(built):
    fg, order = W.scalar_chain(5, N=16), W.chain_nd_order(5)
    tree = TR.buildTree(fg, order)
    with pytest.raises(A.IIFB200Error):
        PL.plan_tree(fg, tree, N=100000)
    # a malformed clique table (a separator its parent does not hold) is refused, not lowered
    bad = TR.buildTree(fg, order)
    child = next(c for c in bad.cliques if c.parent is not None)
    stranger = next(v for v in fg.variables if v not in bad.cliques[child.parent].allvars and v not in child.allvars)
    child.separators = child.separators + [stranger]
    with pytest.raises(A.IIFB200Error, match="separator missing"):
        PL.plan_tree(fg, bad)


@pytest.mark.gpu
def test_plan_upload_solves_like_the_python_path(built):
    """iifb200_plan_upload (set_graph + schedule_build inside the library) followed by upload_slots / schedule_run /
    download_slots — the whole B4 sequence a reference-side caller makes — gives the TreeSolver's posteriors."""
    from iifb200 import solver as SV
    fg, order = W.scalar_chain(64, N=100, seed=3), W.chain_nd_order(64)
    ts = SV.TreeSolver(fg, order, planner="python")
    ts.load_from_graph(); ts.upload(); ts.run(); ts.download()
    tc = SV.TreeSolver(fg, order, planner="c")
    assert tc.plan_handle is not None
    tc.load_from_graph(); tc.upload(); tc.run(); tc.download()
    for l in fg.variables:
        for x, y in zip(ts.arena.get(ts.plan.var_slot[l]), tc.arena.get(tc.plan.var_slot[l])):
            assert np.array_equal(x, y), l
    ts.close(); tc.close()


@pytest.mark.gpu
def test_b3_per_call_path_equals_the_one_schedule_run(built):
    """boundary B3 (one propagateBelief per C-ABI round trip, as julia/IIFB200.jl's propagateBelief does) gives the
    posteriors of boundary B4 (one schedule): same ops, same Philox call ids; only the number of round trips differs."""
    from iifb200 import solver as SV
    fg, order = W.scalar_chain(24, N=64, seed=3), W.chain_nd_order(24)
    ts = SV.TreeSolver(fg, order)
    ts.load_from_graph(); ts.upload(); ts.run(); ts.download()
    for fused in (True, False):     # iifb200_propagate_once, and its four constituent calls
        b3 = SV.B3Driver(ts.plan, ts.sp_c, fused=fused)
        ar = CP.HostArena(ts.plan.frozen)
        for l, v in fg.variables.items():
            ar.set(ts.plan.var_slot[l], v.val, v.bw, True)
        b3.run(ar)
        n0 = b3.eng.launch_count()
        assert n0 >= len(ts.plan.props)
        for l in fg.variables:
            a, b = ts.arena.get(ts.plan.var_slot[l]), ar.get(ts.plan.var_slot[l])
            assert np.allclose(a[0], b[0], rtol=0, atol=1e-9) and np.allclose(a[1], b[1], rtol=1e-7), (fused, l)
        b3.close()
    ts.close()


@pytest.mark.gpu
def test_b3_context_pool_equals_the_serial_calls(built):
    """the shim's context pool (independent propagateBelief calls in flight at once, one library context each, from
    several host threads) returns the serial per-call result bit for bit: contexts share nothing but the device"""
    from iifb200 import solver as SV
    fg, order = W.scalar_chain(40, N=64, seed=5), W.chain_nd_order(40)
    ts = SV.TreeSolver(fg, order)
    res = []
    for k in (1, 4):
        b3 = SV.B3Driver(ts.plan, ts.sp_c, contexts=k)
        ar = CP.HostArena(ts.plan.frozen)
        for l, v in fg.variables.items():
            ar.set(ts.plan.var_slot[l], v.val, v.bw, True)
        b3.run(ar)
        res.append(ar)
        b3.close()
    ts.close()
    assert np.array_equal(res[0].pts, res[1].pts) and np.array_equal(res[0].bw, res[1].bw)


def test_library_elimination_orders(built):
    """iifb200_elimination_order_nd / _is reproduce the Python mirror's order on chains, grids with loop closures, forests
    (disconnected graphs), rings and the reference's small canonical graphs; on a chain the resulting tree has
    logarithmic depth (the default order's tree is a path)."""
    cases = [W.scalar_chain(37, N=8, seed=1), W.scalar_chain(1000, N=8, seed=1),
             W.euclid2_grid(rows=5, cols=9, N=8, seed=2, closure_every=2), W.scalar_chain_sessions(3, 12, N=8),
             W.circular_chain(n=15, N=8), W.generateGraph_Kaess(N=8), W.generateGraph_CaesarRing1D(N=8), W.four_door(N=8)]
    for fg in cases:
        ref = TR.nested_dissection_order(fg)
        got = PL.elimination_order_nd(fg)
        assert got == ref
        assert sorted(got) == sorted(fg.variables)
        for slack in (0, 1, 2):
            assert PL.elimination_order_is(fg, slack) == TR.independent_set_order(fg, slack)
    fg = cases[1]
    for order in (PL.elimination_order_nd(fg), PL.elimination_order_is(fg)):
        tree = TR.buildTree(fg, order)
        depth = {}
        for c in tree.cliques:                      # parents are created before their children
            depth[c.id] = 0 if c.parent is None else depth[c.parent] + 1
        assert max(depth.values()) <= 2 * int(np.ceil(np.log2(1000))) + 2
        assert max(len(c.frontals) for c in tree.cliques) <= 3
    tree = TR.buildTree(fg, PL.elimination_order_nd(fg))
    depth = {}
    for c in tree.cliques:                      # parents are created before their children
        depth[c.id] = 0 if c.parent is None else depth[c.parent] + 1
    assert max(depth.values()) <= 2 * int(np.ceil(np.log2(1000))) + 2
    tree_nat = TR.buildTree(fg, TR.getEliminationOrder(fg, "natural"))
    dn = {}
    for c in tree_nat.cliques:
        dn[c.id] = 0 if c.parent is None else dn[c.parent] + 1
    assert max(dn.values()) > 900
