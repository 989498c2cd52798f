"""End-to-end tests through the reference-facing API mirror (initfg / addVariable / addFactor /
approxConv / propagateBelief / initAll / solveTree) on the GPU, with the reference's own acceptance
bands (statistical, as in the reference's tests).  Every numeric step is a launch in libiifb200.so."""
import numpy as np
import pytest

import iifb200  # noqa: F401
from iifb200 import graph as G
from iifb200 import solver as SV
from iifb200 import tree as TR
from iifb200 import workloads as W

pytestmark = pytest.mark.gpu


def _pts(fg, l):
    return G.getPoints(G.getBelief(fg, l))[:, 0]


def test_single_prior_solveTree_testBasicGraphs(built):
    """testBasicGraphs.jl:19-47 and :59-72."""
    for mu, lo, hi in ((0.0, 0.3, 1.9), (1000.0, 0.4, 1.8)):
        fg = G.initfg(G.SolverParams(seed=11))
        G.addVariable(fg, "x0", G.ContinuousScalar)
        G.addFactor(fg, ["x0"], G.Prior(G.Normal(mu, 1.0)))
        SV.solveTree(fg)
        p = _pts(fg, "x0")
        assert len(p) == 100 and abs(p.mean() - mu) < 0.5 and lo < p.var(ddof=1) < hi


def test_two_and_three_identical_priors(built):
    """testBasicGraphs.jl:77-113."""
    for k, hi in ((2, 1.0), (3, 0.75)):
        fg = G.initfg(G.SolverParams(seed=5))
        G.addVariable(fg, "x0", G.ContinuousScalar)
        for _ in range(k):
            G.addFactor(fg, ["x0"], G.Prior(G.Normal(0.0, 1.0)))
        SV.solveTree(fg)
        p = _pts(fg, "x0")
        assert abs(p.mean()) < 0.4 and 0.1 < p.var(ddof=1) < hi


def test_five_variable_chain_two_priors(built):
    """testBasicGraphs.jl:249-300: priors at -3 / +3 on the ends of a 5-chain."""
    fg = G.initfg(G.SolverParams(seed=7))
    for k in range(5):
        G.addVariable(fg, f"x{k}", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(-3.0, 1.0)))
    G.addFactor(fg, ["x4"], G.Prior(G.Normal(3.0, 1.0)))
    for k in range(4):
        G.addFactor(fg, [f"x{k}", f"x{k+1}"], G.LinearRelative(G.Normal(0.0, 1.0)))
    ts = SV.solveTree(fg)
    X = [_pts(fg, f"x{k}").mean() for k in range(5)]
    assert X[0] < X[1] < X[2] < X[3] < X[4]
    assert abs(X[0] + X[4]) < 2.2 and abs(X[1] + X[3]) < 2.2 and abs(X[2]) < 2.2
    for k, hi in enumerate((2.8, 2.9, 3.0, 3.1, 3.2)):
        assert 0.2 < _pts(fg, f"x{k}").var(ddof=1) < hi
    assert ts.plan.n_conv > 0
    mkd, _, lbls, ipc = SV.localProduct(fg, "x2")       # :303-306
    assert G.Npts(mkd) == 100 and len(lbls) == 2


def test_c1_four_variable_chain(built):
    """BASELINE configs[0]: 4-variable scalar chain, Prior + LinearRelative(Normal(1, 0.01)) (testBasicGraphs.jl:325-343)."""
    fg = G.initfg(G.SolverParams(seed=42))
    G.addVariable(fg, "x0", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(1.0, 0.01)))
    SV.initAll(fg)
    assert abs(_pts(fg, "x0").mean() - 1.0) < 0.1
    for k in range(1, 4):
        G.addVariable(fg, f"x{k}", G.ContinuousScalar)
        G.addFactor(fg, [f"x{k-1}", f"x{k}"], G.LinearRelative(G.Normal(1.0, 0.01)))
    SV.solveTree(fg)
    for k in range(4):
        assert abs(_pts(fg, f"x{k}").mean() - (k + 1.0)) < 0.1


def test_approxConv_kaess_chain(built):
    """testApproxConv.jl:40-60: prior -> neighbour -> neighbour."""
    fg = W.generateGraph_Kaess(seed=3)
    pts = SV.approxConv(fg, "x1f1", "x1", N=100)
    assert pts.shape == (100, 1) and abs(pts.mean()) < 0.4 and 0.5 < pts.std() < 1.5
    G.initVariable(fg, "x1", pts)
    pts = SV.approxConv(fg, "x1x2f1", "x2")
    assert abs(pts.mean()) < 0.7 and 0.7 < pts.std() < 2.0
    # the target variable itself is untouched by approxConv (ApproxConv.jl:17)
    assert fg.variables["x2"].val.shape[0] == 0
    assert SV.approxConv(fg, "x1x2f1", "x2", N=101).shape == (101, 1)        # testVariousNSolveSize.jl:18-25


def test_multihypo_api(built):
    """testmultihypothesisapi.jl style: x0 observes a landmark that is l1 or l2 with probability 1/2."""
    fg = G.initfg(G.SolverParams(seed=13, graphinit=False))
    for l in ("x0", "l1", "l2"):
        G.addVariable(fg, l, G.ContinuousScalar)
    R = np.random.default_rng(0)
    G.initVariable(fg, "x0", R.normal(0, 1, (100, 1)))
    G.initVariable(fg, "l1", R.normal(-30, 1, (100, 1)))
    G.initVariable(fg, "l2", R.normal(40, 1, (100, 1)))
    G.addFactor(fg, ["x0", "l1", "l2"], G.LinearRelative(G.Normal(10.0, 1.0)), multihypo=[1.0, 0.5, 0.5])
    mkd, lab = SV.approxConvBelief(fg, "x0l1l2f1", "x0", return_labels=True)
    p = G.getPoints(mkd)[:, 0]
    assert (np.abs(p + 40) < 6).sum() > 20 and (np.abs(p - 30) < 6).sum() > 20 and set(np.unique(lab)) <= {2, 3}
    # host-drawn labels pass through bit-exact (boundary B2)
    mine = R.integers(2, 4, 100).astype(np.int32)
    _, lab2 = SV.approxConvBelief(fg, "x0l1l2f1", "x0", mhidx=mine, return_labels=True)
    assert np.array_equal(lab2, mine)


def test_c3_four_door_mixture_solve(built):
    """BASELINE configs[2] (test/fourdoortest.jl, useMsgLikelihoods=false): runs and stays multi-modal."""
    fg = W.four_door(N=200, seed=42)
    ts = SV.solveTree(fg)
    for l in ("x1", "x2", "x3", "x4"):
        p = _pts(fg, l)
        assert len(p) == 200 and np.isfinite(p).all()
    # the reference's fourdoortest.jl is a smoke test (no assertions); every belief must stay on the door
    # lattice implied by the mixture priors and the odometry (doors at -100, 0, 100, 300)
    doors = np.array([-100.0, 0.0, 100.0, 300.0])
    for l, off in (("x1", 0.0), ("x3", 0.0), ("x4", 0.0), ("x2", 50.0)):
        p = _pts(fg, l)
        near = np.min(np.abs(p[:, None] - (doors + off)[None, :]), axis=1) < 25
        assert near.mean() > 0.8, (l, near.mean())
    assert ts.plan.n_conv > 10


def test_c4_circular_chain(built):
    """BASELINE configs[3] scaled down (testCircular.jl:14-29): PPE ~ rem2pi(k) within 0.35 rad."""
    fg = W.circular_chain(n=12, N=150, seed=42)
    SV.solveTree(fg, eliminationOrder=W.chain_nd_order(12))
    for k in range(12):
        p = _pts(fg, f"x{k}")
        mu = np.arctan2(np.sin(p).mean(), np.cos(p).mean())
        err = np.abs((mu - k + np.pi) % (2 * np.pi) - np.pi)
        assert err < 0.35, (k, mu)


def test_c5_euclid2_grid_small(built):
    """BASELINE configs[4] scaled down: Position{2} grid with loop closures, nested-dissection order."""
    fg = W.euclid2_grid(rows=4, cols=6, N=100, seed=42, closure_every=2)
    ts = SV.solveTree(fg, ordering="nd")
    assert max(len(c.separators) for c in ts.tree.cliques) >= 2
    for k, v in enumerate(fg.variables.values()):
        assert v.val.shape == (100, 2) and np.isfinite(v.val).all()
    # the corner far from the prior is still located to within the accumulated odometry noise
    last = fg.variables[f"x{len(fg.variables)-1}"].val.mean(axis=0)
    assert np.abs(last - np.array([0.0, 3.0])).max() < 1.5


def test_unsupported_variable_dim_fails_loudly():
    fg = G.initfg(G.SolverParams(graphinit=False))
    with pytest.raises(AssertionError):
        G.addVariable(fg, "p", G.Position(7))


def test_calcPPE_after_solveTree(built):
    """setPPE! at CSM step 5 (CliqueStateMachine.jl:933-939): after solveTree every variable carries a
    MeanMaxPPE; on the 4-variable chain of testBasicGraphs.jl:325-343 the estimates sit at 0, 1, 2, 3."""
    fg = G.initfg(G.SolverParams(seed=3))
    for k in range(4):
        G.addVariable(fg, f"x{k}", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(0.0, 1.0)))
    for k in range(3):
        G.addFactor(fg, [f"x{k}", f"x{k + 1}"], G.LinearRelative(G.Normal(1.0, 0.01)))
    SV.solveTree(fg)
    for k in range(4):
        ppe = SV.getPPE(fg, f"x{k}")
        assert abs(ppe.suggested[0] - _pts(fg, f"x{k}").mean()) < 1e-9
        assert abs(ppe.suggested[0] - k) < 0.75 and abs(ppe.max[0] - k) < 1.0
    one = SV.calcPPE(fg, "x2")
    assert np.allclose(one.mean, SV.getPPE(fg, "x2").mean) and np.allclose(one.max, SV.getPPE(fg, "x2").max)


def test_initAll_batched_matches_sequential(built):
    """the wavefront (one CUDA-graph schedule) form of initAll! reproduces the sequential
    doautoinit! sweep (same order, factors and Philox call ids) on a chain and on a grid with loop closures."""
    def chain(seed):
        fg = G.initfg(G.SolverParams(seed=seed))
        for k in range(12):
            G.addVariable(fg, f"x{k}", G.ContinuousScalar)
        G.addFactor(fg, ["x0"], G.Prior(G.Normal(0.0, 0.1)))
        for k in range(11):
            G.addFactor(fg, [f"x{k}", f"x{k + 1}"], G.LinearRelative(G.Normal(1.0, 0.1)))
        return fg

    def kaess(seed):
        return W.generateGraph_Kaess(N=100, seed=seed, graphinit=False)

    for make in (chain, kaess):
        a, b = make(21), make(21)
        SV.initAll(a, batched=False)
        SV.initAll(b, batched=True)
        for l in a.variables:
            assert a.variables[l].initialized and b.variables[l].initialized
            pa, pb = a.variables[l].val, b.variables[l].val
            assert pa.shape == pb.shape
            # identical streams; only the CTA size of a launch (hence a summation order) may differ
            assert np.allclose(pa, pb, rtol=0, atol=1e-9), l
            assert np.allclose(a.variables[l].bw, b.variables[l].bw, rtol=1e-7)
    # the batched form leaves the graph solvable
    c = chain(5)
    SV.solveTree(c)
    assert abs(_pts(c, "x11").mean() - 11.0) < 1.5


def test_c2_full_size_chain_properties(built):
    """BASELINE configs[1] at FULL size (1000 poses, N=100, nested-dissection order): size-independent
    properties of the posterior — every belief complete and finite, located at its pose index within the
    accumulated odometry noise, bandwidths positive, same seed => bit-identical run, new seed => new samples,
    convolution count of the plan as derived in SURVEY.md A.9 (~10 per pose with this order)."""
    n = 1000
    fg = W.scalar_chain(n, N=100, seed=42)
    ts = SV.TreeSolver(fg, W.chain_nd_order(n))
    assert 9 * n < ts.plan.n_conv < 13 * n and ts.plan.n_prod > 5 * n

    def run(seed):
        from iifb200 import compile as CP
        ts.eng.set_solver_params(CP.solver_params_c(fg.solverParams, seed))
        ts.load_from_graph()
        ts.upload()
        ts.run()
        ts.download()
        nv = len(fg.variables)
        pts = np.stack([ts.arena.get(ts.plan.var_slot[f"x{k}"])[0][:, 0] for k in range(nv)])
        bw = np.array([ts.arena.get(ts.plan.var_slot[f"x{k}"])[1][0] for k in range(nv)])
        return pts, bw

    p1, b1 = run(7)
    p2, b2 = run(7)
    p3, _ = run(8)
    assert p1.shape == (n, 100) and np.isfinite(p1).all() and (b1 > 0).all()
    assert np.array_equal(p1, p2) and np.array_equal(b1, b2)          # deterministic replay of the CUDA graph
    assert not np.array_equal(p1, p3)
    mean_err = np.abs(p1.mean(axis=1) - np.arange(n))
    sigma = 0.1 * np.sqrt(np.arange(n) + 1.0)                          # analytic marginal: prior 0.1, odometry 0.1 per step
    # Band derived from the ORACLE on this plan over 20 seeds (round 2): max |mean error| = 0.68 sigma (2.1 absolute,
    # at the chain end), posterior std between 0.30 and 1.31 sigma.  (Round 1 accepted 6 sigma + 0.5 because its
    # product sampler carried an upward bias of ~3.5 sigma at x997..x999)
    assert (mean_err < 1.0 * sigma + 0.3).all(), (int(np.argmax(mean_err / sigma)), float((mean_err / sigma).max()))
    spread = p1.std(axis=1)
    assert (spread < 1.7 * sigma + 0.1).all() and (spread > 0.15 * sigma).all(), (float((spread / sigma).max()), float((spread / sigma).min()))
    ts.close()


def test_approxDeconv_reference_bands(built):
    """test/testDefaultDeconv.jl:10-31: after solving a 3-pose line, the predicted measurements of the prior and
    of an odometry factor agree with fresh factor samples in the mmd sense (< 1e-8 resp. < 1e-3 there; the prior
    band is taken at 1e-6 because x0 here is a solved posterior, not the prior's own samples)."""
    fg = W.scalar_chain(3, N=100, seed=9)
    SV.solveTree(fg)
    pred, meas = SV.approxDeconv(fg, "x0f1")
    assert pred.shape == (100, 1) and SV.mmd(fg, pred, meas, G.ContinuousScalar) < 1e-6
    pred, meas = SV.approxDeconv(fg, "x0x1f1")
    assert SV.mmd(fg, pred, meas, G.ContinuousScalar) < 1e-3
    assert abs(pred.mean() - 1.0) < 0.2 and abs(meas.mean() - 1.0) < 0.05


def test_se2_solveTree_testSpecialEuclidean2Mani(built):
    """SURVEY 8f-4, test/testSpecialEuclidean2Mani.jl:35-88: ManifoldPrior at the identity + ManifoldFactor
    MvNormal([1, 2, pi/4], 0.01) chain on SpecialEuclidean(2); means within atol 0.1 after doautoinit! and after
    solveTree!; :113-123 a PartialPrior on the translation gives a partial belief with 3 infoPerCoord."""
    se = G.SpecialEuclidean2
    fg = G.initfg(G.SolverParams(seed=17))
    G.addVariable(fg, "x0", se)
    G.addFactor(fg, ["x0"], G.ManifoldPrior(se, G.se2_point_to_coords([0.0, 0.0], np.eye(2)),
                                            G.MvNormal([0, 0, 0], np.diag([1e-4] * 3))))
    SV.doautoinit(fg, "x0")
    assert np.abs(fg.variables["x0"].val.mean(axis=0)).max() < 0.1
    G.addVariable(fg, "x1", se)
    G.addFactor(fg, ["x0", "x1"], G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, np.pi / 4], np.diag([0.01] * 3))))
    SV.doautoinit(fg, "x1")

    def mean(l):
        p = fg.variables[l].val
        return np.array([p[:, 0].mean(), p[:, 1].mean(), np.arctan2(np.sin(p[:, 2]).mean(), np.cos(p[:, 2]).mean())])

    assert np.abs(mean("x1") - [1.0, 2.0, np.pi / 4]).max() < 0.1
    SV.solveTree(fg)
    assert np.abs(mean("x0")).max() < 0.1 and np.abs(mean("x1") - [1.0, 2.0, np.pi / 4]).max() < 0.1
    t, Rm = G.se2_coords_to_point(mean("x1"))
    assert np.allclose(Rm, [[0.7071, -0.7071], [0.7071, 0.7071]], atol=0.1) and np.allclose(t, [1.0, 2.0], atol=0.1)
    G.addVariable(fg, "x2", se)
    G.addFactor(fg, ["x1", "x2"], G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, np.pi / 4], np.diag([0.01] * 3))))
    SV.solveTree(fg)
    c = np.sqrt(0.5)
    assert np.abs(mean("x2") - [1.0 - c, 2.0 + 3 * c, np.pi / 2]).max() < 0.25
    ppe = SV.getPPE(fg, "x2")
    assert np.abs(ppe.suggested - mean("x2")).max() < 1e-9
    # partial prior on an SE(2) variable
    fg = G.initfg(G.SolverParams(seed=3, graphinit=False))
    G.addVariable(fg, "x0", se)
    G.addFactor(fg, ["x0"], G.PartialPrior(G.MvNormal([0.01, 0.01], np.eye(2) * 1e-4), (1, 2)))
    pbel = SV.approxConvBelief(fg, "x0f1", "x0")
    assert pbel.partial == [1, 2] and len(pbel.infoPerCoord) == 3


def _solve_twice(ts, fg, seed):
    from iifb200 import compile as CP
    out = []
    for _ in range(2):
        ts.eng.set_solver_params(CP.solver_params_c(fg.solverParams, seed))
        ts.load_from_graph()
        ts.upload()
        ts.run()
        ts.download()
        out.append({l: ts.arena.get(ts.plan.var_slot[l]) for l in fg.variables})
    return out


def test_c4_full_size_circular_chain_properties(built):
    """BASELINE configs[3] at FULL size on one GPU (500 Circular poses, N=150, testCircular.jl:14-16 scaled):
    every belief complete, in [-pi, pi), located at rem2pi(k) within the accumulated odometry noise (prior 0.1,
    0.1 rad per step), bandwidths positive, deterministic replay."""
    n = 500
    fg = W.circular_chain(n=n, N=150, seed=42)
    ts = SV.TreeSolver(fg, W.chain_nd_order(n))
    assert 9 * n < ts.plan.n_conv < 13 * n
    a, b = _solve_twice(ts, fg, 11)
    for k in range(n):
        p, bw, ipc = a[f"x{k}"]
        assert p.shape == (150, 1) and np.isfinite(p).all() and (-np.pi <= p).all() and (p < np.pi).all() and bw[0] > 0
        assert np.array_equal(p, b[f"x{k}"][0])
        mu = np.arctan2(np.sin(p).mean(), np.cos(p).mean())
        err = abs((mu - k + np.pi) % (2 * np.pi) - np.pi)
        assert err < 0.15 + 0.1 * np.sqrt(k + 1.0), (k, mu, err)
    ts.close()


def test_c5_full_size_grid_properties(built):
    """BASELINE configs[4] at FULL size on one GPU (5000 Position{2} poses on a 50 x 100 boustrophedon grid with
    loop closures every 5th column, N=100, nested-dissection order: ~4200 cliques, separators up to ~100
    variables, ~79 k convolutions): beliefs complete and finite, every pose located on its grid node (mean error
    < 0.4 on average, < 2.5 anywhere - the loop closures bound the drift), deterministic replay."""
    rows, cols = 50, 100
    fg = W.euclid2_grid(rows=rows, cols=cols, N=100, seed=42, closure_every=5)
    ts = SV.TreeSolver(fg, TR.getEliminationOrder(fg, "nd"))
    assert ts.plan.n_conv > 10 * rows * cols and max(len(c.separators) for c in ts.tree.cliques) >= 50
    a, b = _solve_twice(ts, fg, 5)
    pos = []
    for r in range(rows):
        for c in (range(cols) if r % 2 == 0 else range(cols - 1, -1, -1)):
            pos.append((float(c), float(r)))
    pos = np.array(pos)
    mean = np.stack([a[f"x{k}"][0].mean(axis=0) for k in range(rows * cols)])
    for k in range(rows * cols):
        p, bw, _ = a[f"x{k}"]
        assert p.shape == (100, 2) and np.isfinite(p).all() and (bw > 0).all()
    assert all(np.array_equal(a[l][0], b[l][0]) for l in fg.variables)
    err = np.abs(mean - pos).max(axis=1)
    assert err.mean() < 0.4 and err.max() < 2.5, (float(err.mean()), float(err.max()))
    ts.close()


def test_lanes_do_not_change_results(built):
    """Lanes only add parallel branches to the captured CUDA graph: the posterior of every variable is bit-identical
    with 0, 2 and 4 lanes (chain with differential messages off and on, and a grid)."""
    cases = [(W.scalar_chain(120, N=100, seed=4), W.chain_nd_order(120), False),
             (W.scalar_chain(60, N=64, seed=4), W.chain_nd_order(60), True),
             (W.euclid2_grid(rows=5, cols=8, N=64, seed=4, closure_every=2), None, False)]
    for fg, order, uml in cases:
        fg.solverParams.useMsgLikelihoods = uml
        order = order or TR.getEliminationOrder(fg, "nd")
        res = []
        for lanes in (0, 2, 4):
            ts = SV.TreeSolver(fg, order, lanes=lanes)
            assert (max(ts.plan.op_lane) > 0) == (lanes > 0)
            ts.load_from_graph()
            ts.upload()
            ts.run()
            ts.run()                                    # replay of the cached graph
            ts.download()
            res.append({l: ts.arena.get(ts.plan.var_slot[l]) for l in fg.variables})
            ts.close()
        for l in fg.variables:
            for r in res[1:]:
                assert np.array_equal(res[0][l][0], r[l][0]) and np.array_equal(res[0][l][1], r[l][1]), l


def test_independent_sessions_in_one_pass(built):
    """Several independent graphs in one factor graph (a forest: one Bayes-tree root per session) are solved by one
    pass; every session is localised as when solved alone (priors at 0, 10, 20; unit odometry)."""
    B, n = 3, 24
    fg = W.scalar_chain_sessions(B, n, N=100, seed=6)
    ts = SV.solveTree(fg, eliminationOrder=W.sessions_nd_order(B, n))
    assert len(ts.tree.roots) == B
    for b in range(B):
        for k in (0, n // 2, n - 1):
            p = _pts(fg, f"s{b}x{k}")
            assert abs(p.mean() - (10.0 * b + k)) < 0.35 + 0.15 * np.sqrt(k + 1.0), (b, k, p.mean())


def test_so3_solveTree_testSpecialOrthogonalMani(built):
    """SURVEY 8f-4, test/testSpecialOrthogonalMani.jl:75-140: SpecialOrthogonal(3) ManifoldPrior at the identity
    (MvNormal([0.01, 0.01, 0.01])) and ManifoldFactor MvNormal([0.01, 0.01, 0.01], [0.01, 0.01, 0.01]): mean(M, pts)
    within atol 0.01 of I resp. Exp([0.01, 0.01, 0.01]) after doautoinit! and after solveTree!; all points valid."""
    so3 = G.SpecialOrthogonal3
    fg = G.initfg(G.SolverParams(seed=23))
    G.addVariable(fg, "x0", so3)
    G.addFactor(fg, ["x0"], G.ManifoldPrior(so3, np.eye(3), G.MvNormal([0, 0, 0], np.diag([1e-4] * 3))))
    SV.doautoinit(fg, "x0")
    assert np.abs(G.so3_mean(fg.variables["x0"].val) - np.eye(3)).max() < 0.01
    G.addVariable(fg, "x1", so3)
    G.addFactor(fg, ["x0", "x1"], G.ManifoldFactor(so3, G.MvNormal([0.01, 0.01, 0.01], np.diag([1e-4] * 3))))
    SV.doautoinit(fg, "x1")
    want = np.array([[0.9999, -0.00995, 0.01005], [0.01005, 0.9999, -0.00995], [-0.00995, 0.01005, 0.9999]])
    assert np.abs(G.so3_mean(fg.variables["x1"].val) - want).max() < 0.01
    SV.solveTree(fg)
    assert np.abs(G.so3_mean(fg.variables["x0"].val) - np.eye(3)).max() < 0.01
    assert np.abs(G.so3_mean(fg.variables["x1"].val) - want).max() < 0.01
    for l in ("x0", "x1"):
        p = fg.variables[l].val
        assert p.shape == (100, 3) and (np.linalg.norm(p, axis=1) <= np.pi).all()
    # a larger rotation chain: x2 = x1 Exp((0.8, -0.5, 1.1)) keeps composing on the group, not in coordinates
    G.addVariable(fg, "x2", so3)
    G.addFactor(fg, ["x1", "x2"], G.ManifoldFactor(so3, G.MvNormal([0.8, -0.5, 1.1], np.diag([1e-4] * 3))))
    SV.solveTree(fg)
    want2 = want @ G.so3_coords_to_point([0.8, -0.5, 1.1])
    assert np.abs(G.so3_mean(fg.variables["x2"].val) - want2).max() < 0.03
