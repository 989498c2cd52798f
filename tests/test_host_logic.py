"""Host logic (no GPU): descriptor lowering, parseusermultihypo, and the Bayes-tree / schedule
descriptor the GPU batcher consumes, pinned by the reference's tree known-answer tests."""
import numpy as np
import pytest

import iifb200  # noqa: F401
from iifb200 import _abi as A
from iifb200 import compile as CP
from iifb200 import graph as G
from iifb200 import tree as TR
from iifb200 import workloads as W


def test_parseusermultihypo():
    """FactorGraph.jl:633-654."""
    assert G.parseusermultihypo(None, 0.1) == (None, 0.1)
    mh, _ = G.parseusermultihypo([1.0, 0.5, 0.5], 0.0)
    assert np.allclose(mh, [0.0, 0.5, 0.5])
    mh, _ = G.parseusermultihypo([1, 0.25, 0.25, 0.25, 0.25], 0.0)
    assert np.allclose(mh, [0, 0.25, 0.25, 0.25, 0.25])
    with pytest.raises(AssertionError):
        G.parseusermultihypo([1.0, 0.5, 0.4], 0.0)


def test_tables_lowering():
    T = CP.Tables()
    a = T.add_slot(G.Position(2), 100)
    b = T.add_slot(G.Circular, 150)
    f1 = T.add_factor(G.LinearRelative(G.MvNormal([1.0, 2.0], np.diag([4.0, 9.0]))), [a, a])
    f2 = T.add_factor(G.Mixture(G.Prior, [G.Normal(0, 1), G.Normal(5, 2)], [1, 3]), [b])
    f3 = T.add_factor(G.PartialPrior(G.Normal(0, 1), (2,)), [a])
    fz = T.freeze()
    assert fz["slots"][1].pts_off == 200 and fz["slots"][1].circ_mask == 1 and fz["total_doubles"] == 350
    d1 = fz["dists"][fz["factors"][f1].dist]
    assert d1.kind == A.D_MVNORMAL and d1.dim == 2
    prm = fz["dparams"][d1.poff:d1.poff + 6]
    assert np.allclose(prm, [1, 2, 2, 0, 0, 3])     # mu, then row-major lower Cholesky
    d2 = fz["dists"][fz["factors"][f2].dist]
    assert d2.kind == A.D_MIXTURE and d2.ncomp == 2 and d2.comp_kind == A.D_NORMAL
    assert np.allclose(fz["dparams"][d2.poff:d2.poff + 6], [0.25, 0.75, 0, 1, 5, 2])
    assert fz["factors"][f3].partial_mask == 0b10 and fz["factors"][f3].kind == A.F_PARTIAL_PRIOR
    with pytest.raises(A.IIFB200Error):
        T.add_dist(object())


def test_unsupported_factor_fails_loudly():
    fg = G.initfg(G.SolverParams(graphinit=False))
    G.addVariable(fg, "x0", G.ContinuousScalar)

    class UserFactor:
        Z = G.Normal()
    with pytest.raises(A.IIFB200Error):
        G.addFactor(fg, ["x0"], UserFactor())


def test_kaess_tree_known_answer():
    """testBayesTreeiSAM2Example.jl:41-52 — order [:l1,:l2,:x1,:x2,:x3] gives 3 cliques."""
    fg = W.generateGraph_Kaess()
    tree = TR.buildTree(fg, ["l1", "l2", "x1", "x2", "x3"])
    assert len(tree.cliques) == 3 and tree.eliminationOrder == ["l1", "l2", "x1", "x2", "x3"]
    root = tree.cliques[tree.roots[0]]
    assert sorted(root.frontals) == ["x2", "x3"] and root.separators == []
    kids = {tuple(sorted(tree.cliques[c].frontals)): tree.cliques[c] for c in root.children}
    assert set(kids) == {("l1", "x1"), ("l2",)}
    assert kids[("l1", "x1")].separators == ["x2"] and kids[("l2",)].separators == ["x3"]
    # every factor is a potential of exactly one clique
    pots = [f for c in tree.cliques for f in c.potentials]
    assert sorted(pots) == sorted(fg.factors)


def test_caesar_ring_tree_known_answer():
    """testJunctionTreeConstruction.jl:19-64."""
    fg = W.generateGraph_CaesarRing1D()
    tree = TR.buildTree(fg, ["x0", "x2", "x4", "x6", "x1", "l1", "x5", "x3"])
    assert len(tree.cliques) == 6
    cl = lambda v: tree.cliques[tree.frontal_of[v]]  # noqa: E731
    C0 = cl("x3")
    assert sorted(C0.frontals) == ["l1", "x3", "x5"] and C0.separators == [] and len(C0.children) == 3
    C1 = cl("x1")
    assert C1.id in C0.children and C1.frontals == ["x1"] and sorted(C1.separators) == ["l1", "x3"]
    assert len(C1.children) == 2
    for v, sep in (("x2", ["x1", "x3"]), ("x0", ["l1", "x1"])):
        c = cl(v)
        assert c.id in C1.children and c.frontals == [v] and sorted(c.separators) == sep
    for v, sep in (("x6", ["l1", "x5"]), ("x4", ["x3", "x5"])):
        c = cl(v)
        assert c.id in C0.children and c.frontals == [v] and sorted(c.separators) == sep


def test_chain_tree_and_gibbs_classes_match_survey_a9():
    """SURVEY.md A.9 worked schedule for a scalar chain (natural order)."""
    n = 8
    fg = W.scalar_chain(n)
    tree = TR.buildTree(fg, [f"x{k}" for k in range(n)])
    assert len(tree.cliques) == n - 1
    root = tree.cliques[tree.roots[0]]
    assert sorted(root.frontals) == [f"x{n-2}", f"x{n-1}"]
    for c in tree.cliques:
        if c.parent is None:
            continue
        k = int(c.frontals[0][1:])
        assert c.frontals == [f"x{k}"] and c.separators == [f"x{k+1}"]
        assert c.itervarIDs == [f"x{k+1}", f"x{k}"] and not c.directFrtlMsgIDs and not c.msgskipIDs
    plan = TR.compile_solve(fg, tree)
    # 3 Gibbs sweeps x (1 + 2) convolutions per clique up, 2 per non-root clique down
    assert plan.n_conv == 9 * (n - 1) + 2 * (n - 2)
    assert plan.n_msgs == (n - 2) + (n - 2)


def test_waves_respect_slot_hazards():
    fg = W.scalar_chain(33)
    tree = TR.buildTree(fg, W.chain_nd_order(33))
    plan = TR.compile_solve(fg, tree)
    last_w, last_r = {}, {}
    for w, rd, wr in zip(plan.op_wave, plan.op_reads, plan.op_writes):
        for s in rd:
            assert last_w.get(s, -1) < w            # RAW: producer strictly earlier
        for s in wr:
            assert last_w.get(s, -1) < w and last_r.get(s, -1) < w   # WAW / WAR
        for s in rd:
            last_r[s] = max(last_r.get(s, -1), w)
        for s in wr:
            last_w[s] = w
    # within a wave no op reads or writes a slot another op of the wave writes
    for w in range(len(plan.wave_off) - 1):
        ops = range(plan.wave_off[w], plan.wave_off[w + 1])
        writes = [s for i in ops for s in plan.op_writes[i]]
        assert len(writes) == len(set(writes))
        for i in ops:
            assert not (set(plan.op_reads[i]) - set(plan.op_writes[i])) & set(writes)
    assert len(plan.wave_off) - 1 < 40       # nested dissection keeps the pass shallow


def test_nested_dissection_orders_are_permutations():
    for fg in (W.scalar_chain(50), W.euclid2_grid(6, 8, N=10), W.generateGraph_Kaess()):
        o = TR.getEliminationOrder(fg, "nd")
        assert sorted(o) == sorted(fg.variables)
        q = TR.getEliminationOrder(fg, "qr")
        assert sorted(q) == sorted(fg.variables)
    assert sorted(W.chain_nd_order(37)) == sorted(f"x{k}" for k in range(37))


def test_lanes_are_hazard_free():
    """tree.assign_lanes: ops of different lanes never conflict on a slot unless a barrier wave (one holding a
    lane-0 op) separates them; lanes carry the bulk of the work and are balanced."""
    for fg, order, L in ((W.scalar_chain(200, N=8), W.chain_nd_order(200), 4),
                         (W.euclid2_grid(8, 10, N=8), None, 3),
                         (W.scalar_chain(40, N=8), TR.getEliminationOrder(W.scalar_chain(40, N=8), "natural"), 4)):
        order = order or TR.getEliminationOrder(fg, "nd")
        tree = TR.buildTree(fg, order)
        plan = TR.compile_solve(fg, tree, lanes=L)
        ln, wv = plan.op_lane, plan.op_wave
        assert len(ln) == len(plan.sched_waved) and set(ln) <= set(range(L + 1))
        barrier = sorted({w for l, w in zip(ln, wv) if l == 0})

        def separated(wa, wb):
            return any(wa <= b <= wb for b in barrier)
        touched = {}
        for i, (rd, wr) in enumerate(zip(plan.op_reads, plan.op_writes)):
            for s_ in set(rd) | set(wr):
                for j, j_writes in touched.get(s_, []):
                    if (j_writes or s_ in wr) and ln[i] != ln[j] and ln[i] != 0 and ln[j] != 0:
                        assert separated(wv[j], wv[i]), (i, j, s_)
                touched.setdefault(s_, []).append((i, s_ in wr))
        plain = TR.compile_solve(fg, tree)
        assert plain.sched_waved == plan.sched_waved and plain.wave_off == plan.wave_off   # lanes only annotate
        assert set(plain.op_lane) == {0}
    # on the nested-dissection chain most of the work sits in lanes and the lanes are balanced
    fg = W.scalar_chain(200, N=8)
    plan = TR.compile_solve(fg, TR.buildTree(fg, W.chain_nd_order(200)), lanes=4)
    cnt = [sum(1 for l in plan.op_lane if l == k) for k in range(5)]
    assert cnt[0] < 0.2 * sum(cnt) and min(cnt[1:]) > 0.5 * max(cnt[1:])


def test_tree_structure_known_answer_testTreeMessageUtils():
    """test/testTreeMessageUtils.jl:6-39: LineStep(8) ring through lm0 with a fixed elimination order.  The reference
    asserts: up messages exist for cliques 2..8; clique 2's message carries [:x0, :x4]; the variables that appear in
    messages are [:lm0, :x0, :x2, :x4, :x6, :x7]; x0 is in 3 messages; x4 is in the messages of cliques
    4 => depth 2, 6 => 3, 2 => 1, 8 => 1; clique 7 is a child of clique 3.  (Clique ids are 1-based there.)"""
    fg = G.initfg(G.SolverParams(graphinit=False))
    for i in range(9):
        G.addVariable(fg, f"x{i}", G.ContinuousScalar)
        if i == 0:
            G.addFactor(fg, ["x0"], G.Prior(G.Normal(0.0, 0.1)))
            G.addVariable(fg, "lm0", G.ContinuousScalar)
        else:
            G.addFactor(fg, [f"x{i - 1}", f"x{i}"], G.LinearRelative(G.Normal(1.0, 0.1)))
    G.addFactor(fg, ["x0", "lm0"], G.LinearRelative(G.Normal(0.0, 0.1)))
    G.addFactor(fg, ["x8", "lm0"], G.LinearRelative(G.Normal(-8.0, 0.1)))
    tree = TR.buildTree(fg, ["x3", "x8", "x5", "x1", "x6", "lm0", "x7", "x4", "x2", "x0"])
    depth = tree.depth()
    assert len(tree.cliques) == 8 and [c.id + 1 for c in tree.cliques if c.parent is not None] == list(range(2, 9))
    assert set(tree.cliques[1].separators) == {"x0", "x4"}
    assert set().union(*[c.separators for c in tree.cliques]) == {"lm0", "x0", "x2", "x4", "x6", "x7"}
    assert sum(1 for c in tree.cliques if "x0" in c.separators) == 3
    assert {(c.id + 1, depth[c.id]) for c in tree.cliques if "x4" in c.separators} == {(4, 2), (6, 3), (2, 1), (8, 1)}
    assert tree.cliques[6].parent == 2          # clique 7 -> clique 3


def test_factorCanInitFromOtherVars_multihypo_carve_out():
    """GraphInit.jl:62-116 with isLeastOneHypoAvailable (FactorGraph.jl:772-784); the cases listed in the source:
    multihypo=[1;0.5;0.5]: sfidx=1, isinit=[0,1,0] -> true; sfidx=1, isinit=[0,0,1] -> true; sfidx=2|3, isinit=[1,0,0] -> true."""
    fg = G.initfg(G.SolverParams(graphinit=False))
    for l in ("x", "a", "b", "y"):
        G.addVariable(fg, l, G.ContinuousScalar)
    G.addFactor(fg, ["x"], G.Prior(G.Normal()), label="p")
    G.addFactor(fg, ["x", "a", "b"], G.LinearRelative(G.Normal()), multihypo=[1.0, 0.5, 0.5], label="mh")
    G.addFactor(fg, ["x", "y"], G.LinearRelative(G.Normal()), label="xy")
    can = lambda f, l, **init: G.factorCanInitFromOtherVars(fg, f, l, dict(dict(x=False, a=False, b=False, y=False), **init))  # noqa: E731
    assert can("p", "x")                                            # priors always
    assert can("mh", "x", a=True) and can("mh", "x", b=True) and not can("mh", "x")
    assert can("mh", "a", x=True) and can("mh", "b", x=True) and not can("mh", "a", b=True)
    assert can("xy", "y", x=True) and not can("xy", "y") and not can("xy", "y", x=True, y=True)
