"""SURVEY 8f-2 (second half): separator messages with SolverParams.useMsgLikelihoods = true — differential
likelihoods built by approxDeconv + manikde! on the device (IIF_S_DECONV ops) and the joint-message rules of
prepCliqueMsgUp / addLikelihoodsDifferentialCHILD! (src/services/TreeMessageUtils.jl:279-335, 417-469).
CPU tests run the lowered plan through the oracle; `-m gpu` tests run it through libiifb200.so and compare."""
import numpy as np
import pytest

import oracle as O
import parity_cases as PC
from iifb200 import _abi as A
from iifb200 import compile as CP
from iifb200 import graph as G
from iifb200 import tree as TR
from iifb200 import workloads as W


def _oracle_solve(fg, order, **kw):
    tree = TR.buildTree(fg, order)
    plan = TR.compile_solve(fg, tree, **kw)
    arena = CP.HostArena(plan.frozen)
    for l, v in fg.variables.items():
        arena.set(plan.var_slot[l], v.val, v.bw, True, v.infoPerCoord)
    orc = O.Oracle(plan.frozen, arena, CP.solver_params_c(fg.solverParams))
    orc.schedule_run(plan.wave_off, CP.make_sched_ops(plan.sched_waved), CP.make_prop_ops(plan.props),
                     deconvs=CP.make_deconv_ops(plan.deconvs or []))
    return tree, plan, arena


def test_deconv_to_slot_oracle_known_answer():
    """slot := manikde!(exp(M, eps, approxDeconv(LinearRelative dummy))) — TreeMessageUtils.jl:314-321: the points
    are the particle-wise differences x2 - x1 (wrapped on the circle), the bandwidth is manikde!'s."""
    R = np.random.default_rng(3)
    N = 64
    for vt, circ in ((G.ContinuousScalar, False), (G.Circular, True), (G.Position(2), False)):
        P = PC.Problem()
        d = vt.dim
        a = R.normal(3.0, 0.5, (N, d))
        b = R.normal(-2.9, 0.5, (N, d))
        if circ:
            a, b = PC.wrap(a), PC.wrap(b)
        s1, s2 = P.slot(vt, N, a), P.slot(vt, N, b)
        m = P.slot(vt, N, np.zeros((0, d)), initialized=False)
        name, sft = TR.selectFactorType(vt, vt)
        f = P.factor(sft, [s1, s2])
        P.freeze()
        orc = P.oracle()
        orc.deconv_to_slot(CP.make_deconv_ops([dict(factor=f, out_slot=m, N=N, call_id=77)])[0])
        pts, bw, ipc = orc.arena.get(m)
        want = PC.wrap(b - a) if circ else b - a
        assert np.allclose(pts, want, rtol=0, atol=1e-15) and pts.shape == (N, d)
        assert np.allclose(bw, O.kde_bandwidth(want, vt.circ_mask), rtol=1e-12) and np.array_equal(ipc, np.ones(d))
        assert orc.arena.flags[m] == 1 and orc.arena.npts[m] == N
        assert name == ("CircularCircular" if circ else "LinearRelative")
    assert TR.selectFactorType(G.SpecialEuclidean2, G.SpecialEuclidean2) is None
    assert TR.selectFactorType(G.Position(2), G.Position(1)) is None


def test_plan_structure_chain_with_msg_likelihoods():
    """Nested-dissection chain: a clique {x_k | x_a, x_b} marginalises its frontal into ONE differential between its
    two separators (homogeneous LinearRelative path through the frontal); only the branch holding the prior also
    sends a MsgPrior (hasPriors rule, TreeMessageUtils.jl:402-408,464)."""
    n = 17
    fg = W.scalar_chain(n, N=16)
    tree = TR.buildTree(fg, W.chain_nd_order(n))
    p0 = TR.compile_solve(fg, tree, useMsgLikelihoods=False)
    p1 = TR.compile_solve(fg, tree, useMsgLikelihoods=True)
    assert not p0.deconvs and p0.n_conv > 0
    two_sep = [c for c in tree.cliques if c.parent is not None and len(c.separators) == 2]
    assert len(p1.deconvs) == len(two_sep) > 0
    kinds = [k for k, _, _ in p1.sched_waved]
    assert kinds.count(A.S_DECONV) == len(p1.deconvs)
    fz = p1.frozen
    for dc in p1.deconvs:
        f = fz["factors"][dc["factor"]]
        assert f.kind == A.F_LINEAR_RELATIVE and f.arity == 2 and fz["slots"][dc["out_slot"]].dim == 1
    # every differential's measurement slot feeds a relative factor whose distribution is that slot's KDE
    kde_slots = {fz["dists"][fz["factors"][i].dist].slot for i in range(fz["nfactors"])
                 if fz["dists"][fz["factors"][i].dist].kind == A.D_KDE and fz["factors"][i].arity == 2}
    assert kde_slots == {dc["out_slot"] for dc in p1.deconvs}
    # hazards: a DECONV op runs strictly after the writes of the beliefs it reads and before its readers
    last_w = {}
    for w, rd, wr in zip(p1.op_wave, p1.op_reads, p1.op_writes):
        for s in rd:
            assert last_w.get(s, -1) < w
        for s in wr:
            last_w[s] = w
    # with this order the prior's variable x0 sits in the root: no branch holds a prior, two-separator classes send
    # no MsgPrior at all (the information travels in the differentials only, TreeMessageUtils.jl:402-408)
    n_msgprior = lambda p: sum(1 for i in range(p.frozen["nfactors"]) if p.frozen["factors"][i].kind == A.F_MSG_PRIOR)  # noqa: E731
    one_sep = [c for c in tree.cliques if c.parent is not None and len(c.separators) == 1]
    assert n_msgprior(p0) == sum(len(c.separators) for c in tree.cliques if c.parent is not None)
    assert n_msgprior(p1) <= len(one_sep)
    # natural order: a strictly sequential tree of single-separator cliques whose leaf holds the prior: hasPriors
    # propagates up the branch and every message is one MsgPrior, no differentials
    t2 = TR.buildTree(fg, TR.getEliminationOrder(fg, "natural"))
    q1 = TR.compile_solve(fg, t2, useMsgLikelihoods=True)
    q0 = TR.compile_solve(fg, t2, useMsgLikelihoods=False)
    assert not q1.deconvs and n_msgprior(q1) == n_msgprior(q0) == len(t2.cliques) - 1


def test_shortest_path_helper():
    inst = [dict(variables=["a", "b"], type="LinearRelative"), dict(variables=["b", "c"], type="LinearRelative"),
            dict(variables=["c", "d"], type="EuclidDistance"), dict(variables=["a"], type="Prior")]
    assert TR._shortest_path_factor_types(inst, "a", "c") == ["LinearRelative", "LinearRelative"]
    assert TR._shortest_path_factor_types(inst, "a", "d") == ["LinearRelative", "LinearRelative", "EuclidDistance"]
    assert TR._shortest_path_factor_types(inst, "a", "d", "LinearRelative") is None
    assert TR._shortest_path_factor_types(inst, "a", "zz") is None


@pytest.mark.parametrize("circular", [False, True])
def test_oracle_solve_bands_with_msg_likelihoods(circular):
    """The solve with differential messages localises the chain as the plain solve does (test/testCircular.jl:12-29
    runs with useMsgLikelihoods=true and asserts the PPE within 0.1-0.2 rad of rem2pi(k); the scalar line is the
    C2 workload scaled down)."""
    n = 9
    fg = W.circular_chain(n=n, N=64, seed=5) if circular else W.scalar_chain(n, N=64, seed=5)
    fg.solverParams.useMsgLikelihoods = True
    tree, plan, arena = _oracle_solve(fg, W.chain_nd_order(n))
    assert plan.deconvs
    for k in range(n):
        p = arena.get(plan.var_slot[f"x{k}"])[0][:, 0]
        assert p.shape == (64,) and np.isfinite(p).all()
        if circular:
            mu = np.arctan2(np.sin(p).mean(), np.cos(p).mean())
            assert abs(PC.wrap(mu - k)) < 0.5, (k, mu)
        else:
            assert abs(p.mean() - k) < 0.3 + 0.15 * np.sqrt(k + 1.0), (k, p.mean())
    for dc in plan.deconvs:                    # every differential is the odometry its frontal carried: ~ +2
        pts, bw, _ = arena.get(dc["out_slot"])
        m = np.arctan2(np.sin(pts).mean(), np.cos(pts).mean()) if circular else pts.mean()
        assert bw[0] > 0 and pts.shape[0] == 64 and np.isfinite(m)


def _caesar_ring():
    """generateGraph_CaesarRing1D — CanonicalGraphExamples.jl:123-147"""
    fg = G.initfg(G.SolverParams(graphinit=False, N=32))
    for k in range(7):
        G.addVariable(fg, f"x{k}", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal()))
    for k in range(6):
        G.addFactor(fg, [f"x{k}", f"x{k + 1}"], G.LinearRelative(G.Normal()))
    G.addVariable(fg, "l1", G.ContinuousScalar)
    G.addFactor(fg, ["x0", "l1"], G.LinearRelative(G.Normal()))
    G.addFactor(fg, ["x6", "l1"], G.LinearRelative(G.Normal()))
    R = np.random.default_rng(0)
    for l, v in fg.variables.items():
        v.val, v.bw, v.initialized = R.normal(0, 1, (32, 1)), np.array([0.5]), True
    return fg


def test_joint_message_known_answer_testUseMsgLikelihoods():
    """test/testUseMsgLikelihoods.jl:41-78: CaesarRing1D, eliminationOrder [x3,x5,l1,x1,x6,x4,x2,x0].  Clique 2 has
    3 variables and no factor of its own; after addMsgFactors! of the up messages of cliques 4 and 5 it holds exactly
    2 factors, one of them :x0x6f1, a LinearRelative whose Z is a ManifoldKernelDensity (no MsgPrior is added: neither
    child holds a prior)."""
    fg = _caesar_ring()
    tree = TR.buildTree(fg, ["x3", "x5", "l1", "x1", "x6", "x4", "x2", "x0"])
    c2, c4, c5 = tree.cliques[1], tree.cliques[3], tree.cliques[4]
    assert len(c2.allvars) == 3 and not c2.potentials and c2.children == [3, 4]
    assert set(c4.frontals) == {"l1"} and set(c4.separators) == {"x0", "x6"} and set(c5.separators) == {"x4", "x6"}
    plan = TR.compile_solve(fg, tree, useMsgLikelihoods=True, downsolve=False)
    fz = plan.frozen
    slot_var = {}
    idx = len(fg.variables)
    for c in tree.cliques:                       # clique-local slots are allocated in clique order, variable order
        for v in c.allvars:
            slot_var[idx] = (c.id, v)
            idx += 1
    on_c2 = []
    dummies = {dc["factor"] for dc in plan.deconvs}      # the tfg dummy factors of addLikelihoodsDifferentialCHILD!
    for i in range(fz["nfactors"]):
        if i in dummies:
            continue
        f = fz["factors"][i]
        owners = {slot_var.get(f.slot[k], (None, None))[0] for k in range(f.arity)}
        if owners == {1}:
            on_c2.append((f.kind, fz["dists"][f.dist].kind, tuple(slot_var[f.slot[k]][1] for k in range(f.arity))))
    assert len(on_c2) == 2, on_c2
    assert all(k == A.F_LINEAR_RELATIVE and dk == A.D_KDE for k, dk, _ in on_c2)
    assert {frozenset(v) for _, _, v in on_c2} == {frozenset(("x0", "x6")), frozenset(("x4", "x6"))}
    assert len(plan.deconvs) >= 2


def test_hasPriors913_band_oracle():
    """test/testHasPriors913.jl:9-47: line x0..x4 closed through lm0, every belief initialised WRONG (around 5 + i), the
    correct prior Normal(0, 0.01) on x0, useMsgLikelihoods=true; after three consecutive solves every PPE is within
    0.7 of i (the message priors must carry the gauge up and down the tree)."""
    sp = G.SolverParams(graphinit=False, N=64, useMsgLikelihoods=True, seed=7)
    fg = G.initfg(sp)
    for i in range(5):
        G.addVariable(fg, f"x{i}", G.ContinuousScalar)
        if i == 0:
            G.addVariable(fg, "lm0", G.ContinuousScalar)
        else:
            G.addFactor(fg, [f"x{i - 1}", f"x{i}"], G.LinearRelative(G.Normal(1.0, 0.1)))
    G.addFactor(fg, ["x0", "lm0"], G.LinearRelative(G.Normal(0.0, 0.1)))
    G.addFactor(fg, ["x4", "lm0"], G.LinearRelative(G.Normal(-4.0, 0.1)))
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(0.0, 0.01)))
    R = np.random.default_rng(1)
    for l, v in fg.variables.items():
        c = 5.0 + (int(l[1:]) if l.startswith("x") else 0.0)
        v.val, v.initialized = R.normal(c, 0.1, (64, 1)), True
        v.bw = O.kde_bandwidth(v.val)
    order = TR.getEliminationOrder(fg, "qr")
    for it in range(3):
        fg.solverParams.seed = 7 + it
        tree, plan, arena = _oracle_solve(fg, order)
        for l, v in fg.variables.items():
            v.val, v.bw, _ = arena.get(plan.var_slot[l])
    for i in range(5):
        assert abs(fg.variables[f"x{i}"].val.mean() - i) < 0.7, (i, fg.variables[f"x{i}"].val.mean())


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_deconv_slot_op_parity(built):
    """One schedule holding IIF_S_DECONV ops followed by a propagate that samples its measurements from the
    deconvolved belief: device vs oracle."""
    R = np.random.default_rng(9)
    N = 100
    P = PC.Problem(seed=13)
    x0 = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    x1 = P.slot(G.ContinuousScalar, N, R.normal(2, 1, (N, 1)))
    y0 = P.slot(G.ContinuousScalar, N, R.normal(5, 1, (N, 1)))
    y1 = P.slot(G.ContinuousScalar, N, R.normal(0, 3, (N, 1)))
    c0 = P.slot(G.Circular, N, PC.wrap(R.normal(3.0, 0.3, (N, 1))))
    c1 = P.slot(G.Circular, N, PC.wrap(R.normal(-3.0, 0.3, (N, 1))))
    q0 = P.slot(G.Position(2), N, R.normal(0, 1, (N, 2)))
    q1 = P.slot(G.Position(2), 40, R.normal(3, 1, (40, 2)))      # short belief: random partner draws
    m_s = P.slot(G.ContinuousScalar, N, np.zeros((0, 1)), initialized=False)
    m_c = P.slot(G.Circular, N, np.zeros((0, 1)), initialized=False)
    m_q = P.slot(G.Position(2), N, np.zeros((0, 2)), initialized=False)
    d_s = P.factor(TR.selectFactorType(G.ContinuousScalar, G.ContinuousScalar)[1], [x0, x1])
    d_c = P.factor(TR.selectFactorType(G.Circular, G.Circular)[1], [c0, c1])
    d_q = P.factor(TR.selectFactorType(G.Position(2), G.Position(2))[1], [q0, q1])
    rel = P.factor(G.LinearRelative(G.SlotRef(m_s, 1)), [y0, y1])           # _sft(newBel)
    pri = P.factor(G.Prior(G.Normal(7.0, 0.5)), [y1])
    P.freeze()
    dec = [dict(factor=d_s, out_slot=m_s, N=N, call_id=5000), dict(factor=d_c, out_slot=m_c, N=N, call_id=5016),
           dict(factor=d_q, out_slot=m_q, N=N, call_id=5032)]
    props = [dict(target_slot=y1, out_slot=y1, factors=[(rel, 2), (pri, 1)], N=N, call_id=64)]
    sched = [(A.S_DECONV, 0, 0), (A.S_DECONV, 1, 0), (A.S_DECONV, 2, 0), (A.S_PROPAGATE, 0, 0)]
    wave_off = [0, 3, 4]
    dc, pc, sc = CP.make_deconv_ops(dec), CP.make_prop_ops(props), CP.make_sched_ops(sched)
    orc = P.oracle()
    orc.schedule_run(wave_off, sc, pc, deconvs=dc)
    eng = P.engine()
    sid = eng.schedule_build(wave_off, sc, len(sched), pc, len(props), dc, len(dec))
    eng.schedule_run(sid)
    eng.sync()
    ag = P.arena.copy()
    eng.download_arena(ag)
    eng.close()
    PC.assert_arena_equal("deconv_slot", orc.arena, ag, P.frozen, [m_s, m_q, y1])
    PC.assert_arena_equal("deconv_slot_circ", orc.arena, ag, P.frozen, [m_c], circ=True)
    assert ag.flags[m_s] == 1 and ag.npts[m_q] == N


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["chain", "circular", "grid"])
def test_tree_solve_parity_with_msg_likelihoods(built, kind):
    """Whole up + down pass with differential messages: TreeSolver (CUDA graph) vs the oracle on the same plan."""
    from iifb200 import solver as SV
    if kind == "chain":
        fg, order = W.scalar_chain(13, N=64, seed=3), W.chain_nd_order(13)
    elif kind == "circular":
        fg, order = W.circular_chain(n=9, N=64, seed=3), W.chain_nd_order(9)
    else:
        fg = W.euclid2_grid(rows=3, cols=4, N=64, seed=3, closure_every=2)
        order = TR.getEliminationOrder(fg, "nd")
    fg.solverParams.useMsgLikelihoods = True
    tree, plan, ao = _oracle_solve(fg, order)
    assert plan.deconvs
    ts = SV.TreeSolver(fg, order)
    assert len(ts.plan.deconvs) == len(plan.deconvs)
    ts.load_from_graph()
    ts.upload()
    ts.run()
    ts.download()
    slots = [plan.var_slot[l] for l in fg.variables] + [d["out_slot"] for d in plan.deconvs]
    PC.assert_arena_equal(f"uml_{kind}", ao, ts.arena, plan.frozen, slots, circ=(kind == "circular"))
    ts.close()


@pytest.mark.gpu
def test_api_fourdoor_and_circular_as_written(built):
    """BASELINE configs[2] / [3] 'as written in the tests': fourdoortest.jl:19 and testCircular.jl:12 both set
    useMsgLikelihoods=true."""
    from iifb200 import solver as SV
    fg = W.four_door(N=200, seed=42)
    fg.solverParams.useMsgLikelihoods = True
    SV.solveTree(fg)
    doors = np.array([-100.0, 0.0, 100.0, 300.0])
    for l, off in (("x1", 0.0), ("x3", 0.0), ("x4", 0.0), ("x2", 50.0)):
        p = G.getPoints(G.getBelief(fg, l))[:, 0]
        near = np.min(np.abs(p[:, None] - (doors + off)[None, :]), axis=1) < 25
        assert len(p) == 200 and near.mean() > 0.8, (l, near.mean())
    fg = W.circular_chain(n=12, N=150, seed=42)
    fg.solverParams.useMsgLikelihoods = True
    ts = SV.solveTree(fg, eliminationOrder=W.chain_nd_order(12))
    assert ts.plan.deconvs
    for k in range(12):
        p = G.getPoints(G.getBelief(fg, f"x{k}"))[:, 0]
        mu = np.arctan2(np.sin(p).mean(), np.cos(p).mean())
        assert abs(PC.wrap(mu - k)) < 0.35, (k, mu)


@pytest.mark.parametrize("kind", ["disjoint", "homogeneous"])
def test_joint_enforcement_known_answers(built, kind):
    """testJointEnforcement.jl — the reference's known answers for _generateMsgJointRelativesPriors on the clique of x3
    with separators {x0, x2} (eliminationOrder x3, x1, x2, x0):
      :3-77   x3 tied to x0 and x2 by EuclidDistance factors: the path between the separators is homogeneous but NOT
              of the default relative type of ContinuousEuclid{2} (LinearRelative) => 0 relatives, 2 priors;
      :80-150 the same with LinearRelative factors => 1 relative between {x0, x2}, 0 priors.
    Checked on the plan of the Python mirror and on the library's planner (iifb200_plan_tree)."""
    from iifb200 import planner as PL
    rng = np.random.default_rng(1)
    fg = G.initfg(G.SolverParams(N=32, graphinit=False, useMsgLikelihoods=True))
    for k, off in (("x0", 0.0), ("x1", 10.0), ("x2", 20.0)):
        G.addVariable(fg, k, G.Position(2))
        G.initVariable(fg, k, rng.standard_normal((32, 2)) + off, bw=[1.0, 1.0])
    Z = G.MvNormal([10.0, 10.0], np.eye(2))
    G.addFactor(fg, ["x0", "x1"], G.LinearRelative(Z), graphinit=False)
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(Z), graphinit=False)
    G.addVariable(fg, "x3", G.Position(2))
    G.initVariable(fg, "x3", rng.standard_normal((32, 2)) + 30.0, bw=[1.0, 1.0])
    if kind == "disjoint":
        G.addFactor(fg, ["x2", "x3"], G.EuclidDistance(G.Normal(10.0, 1.0)), graphinit=False)
        G.addFactor(fg, ["x0", "x3"], G.EuclidDistance(G.Normal(30.0, 1.0)), graphinit=False)
    else:
        G.addFactor(fg, ["x2", "x3"], G.LinearRelative(Z), graphinit=False)
        G.addFactor(fg, ["x0", "x3"], G.LinearRelative(Z), graphinit=False)
    tree = TR.buildTree(fg, ["x3", "x1", "x2", "x0"])
    c3 = next(c for c in tree.cliques if "x3" in c.frontals)
    assert sorted(c3.separators) == ["x0", "x2"]
    plan = TR.compile_solve(fg, tree, useMsgLikelihoods=True)
    fz = plan.frozen
    mine = [d for d in plan.deconvs if d["clique"] == c3.id]                      # relatives of x3's up message
    slots3 = {s for s, cid in plan.slot_clique.items() if cid == c3.id}
    priors = [i for i in range(fz["nfactors"]) if fz["factors"][i].kind == A.F_MSG_PRIOR
              and fz["dists"][fz["factors"][i].dist].slot in slots3]             # MsgPriors built from x3's beliefs
    msg = plan.up_messages[c3.id]                                                # the message itself (upTx.jointmsg)
    if kind == "disjoint":
        assert len(msg["relatives"]) == 0 and sorted(msg["priors"]) == ["x0", "x2"]          # :72-76
        # the parent adds none of the two priors: the sending sub-graph holds no prior (hasPriors false) and both
        # variables are touched by the parent's own factors (addLikelihoodPriorCommon!, TreeMessageUtils.jl:454-469)
        assert len(mine) == 0 and not msg["hasPriors"] and len(priors) == 0
    else:
        assert [set(r) for r in msg["relatives"]] == [{"x0", "x2"}] and msg["priors"] == []  # :145-149
        assert len(mine) == 1 and len(priors) == 0
        f = fz["factors"][mine[0]["factor"]]
        assert f.kind == A.F_LINEAR_RELATIVE and {plan.slot_clique[f.slot[0]], plan.slot_clique[f.slot[1]]} == {c3.id}
    got = PL.plan_tree(fg, tree, useMsgLikelihoods=True)
    assert [{k: d[k] for k in ("factor", "out_slot", "N", "call_id")} for d in plan.deconvs] == got.deconvs
    assert got.frozen["nfactors"] == fz["nfactors"] and got.props == plan.props
