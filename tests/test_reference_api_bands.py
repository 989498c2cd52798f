"""More of the reference's own API-level tests, mirrored on the GPU path with the reference's acceptance bands:
priorusetest.jl, testlocalconstraintexamples.jl, testVariousNSolveSize.jl, testMultihypoAndChain.jl,
testSolveSetPPE.jl.  Graph construction and solveTree follow the Julia scripts line by line (Python mirror of the
operator API); every numeric step is a launch in libiifb200.so.  The reference seeds its RNG where it is sensitive; here
every test fixes SolverParams.seed and, where the reference's band is a statement about a random outcome, is repeated
over several seeds with the band required of all of them."""
import numpy as np
import pytest

import iifb200  # noqa: F401
from iifb200 import graph as G
from iifb200 import solver as SV
from iifb200 import workloads as W

pytestmark = pytest.mark.gpu


def _mean(fg, l):
    return float(G.getPoints(G.getBelief(fg, l))[:, 0].mean())


@pytest.mark.parametrize("seed", [3, 11, 42])
def test_priorusetest_two_priors_on_a_rigid_chain(built, seed):
    """priorusetest.jl:10-50 (graphinit = true): priors N(-1, 1) on x0 and N(+1, 1) on x2, rigid links (sigma 0.01):
    every mean within 1.0 of 0 and within 0.4 of their common mean."""
    fg = G.initfg(G.SolverParams(N=100, seed=seed))
    G.addVariable(fg, "x0", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(-1.0, 1.0)))
    G.addVariable(fg, "x1", G.ContinuousScalar)
    G.addVariable(fg, "x2", G.ContinuousScalar)
    G.addFactor(fg, ["x2"], G.Prior(G.Normal(+1.0, 1.0)))
    G.addFactor(fg, ["x0", "x1"], G.LinearRelative(G.Normal(0.0, 0.01)))
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(G.Normal(0.0, 0.01)))
    SV.solveTree(fg)
    m = [_mean(fg, l) for l in ("x0", "x1", "x2")]
    assert all(abs(v) < 1.0 for v in m), m
    assert all(abs(v - np.mean(m)) < 0.4 for v in m), m


@pytest.mark.parametrize("seed", [3, 11, 42])
def test_priorusetest_two_priors_with_landmarks(built, seed):
    """priorusetest.jl:54-118: the same two priors through a loop of rigid links over two landmarks."""
    fg = G.initfg(G.SolverParams(N=100, seed=seed))
    G.addVariable(fg, "x0", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(-1.0, 1.0)))
    G.addVariable(fg, "l0", G.ContinuousScalar)
    G.addFactor(fg, ["l0"], G.Prior(G.Normal(+1.0, 1.0)))
    G.addVariable(fg, "l1", G.ContinuousScalar)
    G.addFactor(fg, ["x0", "l0"], G.LinearRelative(G.Normal(0.0, 0.01)))
    G.addFactor(fg, ["x0", "l1"], G.LinearRelative(G.Normal(0.0, 0.01)))
    G.addVariable(fg, "x1", G.ContinuousScalar)
    G.addFactor(fg, ["x0", "x1"], G.LinearRelative(G.Normal(0.0, 0.01)))
    G.addVariable(fg, "x2", G.ContinuousScalar)
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(G.Normal(0.0, 0.01)))
    G.addFactor(fg, ["x2", "l0"], G.LinearRelative(G.Normal(0.0, 0.01)))
    G.addFactor(fg, ["x2", "l1"], G.LinearRelative(G.Normal(0.0, 0.01)))
    SV.solveTree(fg)
    m = {l: _mean(fg, l) for l in ("x0", "x1", "x2", "l0", "l1")}
    assert all(abs(m[l]) < 1.0 for l in ("x0", "x1", "x2")), m
    assert all(abs(m[l]) < 1.2 for l in ("l0", "l1")), m
    assert all(abs(v - np.mean(list(m.values()))) < 0.3 for v in m.values()), m


def test_localconstraint_pose_pose(built):
    """testlocalconstraintexamples.jl:8-47: a belief prior (one door at 0 with bandwidth 3, resampled to N) on x1 and
    LinearRelative(Normal(50, 2)): the convolution onto x2 and the solved x2 have their mean within 15 of 50."""
    N = 100
    rng = np.random.default_rng(5)
    doors2 = 0.0 + 3.0 * rng.standard_normal((N, 1))       # resample(manikde!([[0.0]]; bw = [3.0]), N)
    fg = G.initfg(G.SolverParams(N=N, seed=9))
    G.addVariable(fg, "x1", G.ContinuousScalar)
    G.addFactor(fg, ["x1"], G.Prior(G.SampledBelief(doors2)))
    G.addVariable(fg, "x2", G.ContinuousScalar)
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(G.Normal(50.0, 2.0)))
    SV.initAll(fg)
    pts = SV.approxConv(fg, "x1x2f1", "x2")
    assert abs(np.asarray(pts).mean() - 50.0) < 15.0
    SV.solveTree(fg)
    assert abs(_mean(fg, "x2") - 50.0) < 15.0


def test_various_N_convolution_size(built):
    """testVariousNSolveSize.jl:16-25: approxConv with N = 101 on a graph initialised with N = 100 returns 101 points
    (the destination is resized, short partners are indexed at random: _getindex_anyn)."""
    fg = W.generateGraph_CaesarRing1D(N=100)
    SV.initAll(fg)
    pts = SV.approxConv(fg, "x0x1f1", "x1", N=101)
    assert len(pts) == 101 and np.isfinite(np.asarray(pts)).all()


@pytest.mark.parametrize("seed", [42, 7, 19])
def test_multihypo_basic_single_clique(built, seed):
    """testMultihypoAndChain.jl:7-92: x0 at 0, x1 at 1, l1 at 1, l2 at 2; three sightings with
    multihypo = [1, 0.99, 0.01]; one clique (eliminationOrder l2, x1, x0, l1), gibbsIters = 5, spreadNH = 5.
    PPEs of x0, x1, l1 within 0.2; l2 keeps a mode at 2 (mmd against N(2, 0.1) samples < 1e-3 in the reference — the
    check here is the fraction of l2's points near 2)."""
    rng = np.random.default_rng(seed)
    sp = G.SolverParams(graphinit=True, gibbsIters=5, spreadNH=5.0, N=100, seed=seed)
    fg = G.initfg(sp)
    pr_noise, od_noise, lm_noise = 0.01, 0.1, 0.01
    G.addVariable(fg, "x0", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(rng.normal(0.0, pr_noise), pr_noise)))
    G.addVariable(fg, "l1", G.ContinuousScalar)
    G.addVariable(fg, "l2", G.ContinuousScalar)
    G.addFactor(fg, ["x0", "l1", "l2"], G.LinearRelative(G.Normal(rng.normal(1.0, lm_noise), lm_noise)), multihypo=[1, 0.99, 0.01])
    G.addVariable(fg, "x1", G.ContinuousScalar)
    G.addFactor(fg, ["x0", "x1"], G.LinearRelative(G.Normal(rng.normal(1.0, od_noise), od_noise)))
    G.addFactor(fg, ["x1", "l1", "l2"], G.LinearRelative(G.Normal(rng.normal(0.0, lm_noise), lm_noise)), multihypo=[1, 0.99, 0.01])
    G.addFactor(fg, ["x1", "l2", "l1"], G.LinearRelative(G.Normal(rng.normal(1.0, lm_noise), lm_noise)), multihypo=[1, 0.99, 0.01])
    SV.solveTree(fg, eliminationOrder=["l2", "x1", "x0", "l1"])
    for l, want in (("x0", 0.0), ("x1", 1.0), ("l1", 1.0)):
        assert abs(SV.getPPE(fg, l).suggested[0] - want) < 0.2, (l, SV.getPPE(fg, l).suggested)
    p = G.getPoints(G.getBelief(fg, "l2"))[:, 0]
    assert (np.abs(p - 2.0) < 0.5).mean() > 0.5, np.sort(p)[::10]


@pytest.mark.parametrize("seed", [42, 7, 19])
def test_multihypo_chain_462(built, seed):
    """testMultihypoAndChain.jl:94-136: landmarks at -10 / +10 seen from x1 / x2 with multihypo = [1, 1/2, 1/2]
    against unanchored twins, a tight link between x1 and x2."""
    l1, l2, lnoise, onoise = -10.0, 10.0, 1.0, 0.1
    fg = G.initfg(G.SolverParams(N=100, seed=seed))
    for l in ("x1", "x2", "l1", "l1_0", "l2", "l2_0"):
        G.addVariable(fg, l, G.ContinuousScalar)
    G.addFactor(fg, ["l1"], G.Prior(G.Normal(l1, lnoise)))
    G.addFactor(fg, ["l2"], G.Prior(G.Normal(l2, lnoise)))
    G.addFactor(fg, ["x1", "l1", "l1_0"], G.LinearRelative(G.Normal(l1 - 0.0, lnoise)), multihypo=[1, 0.5, 0.5])
    G.addFactor(fg, ["x2", "l2", "l2_0"], G.LinearRelative(G.Normal(l2 - 0.0, lnoise)), multihypo=[1, 0.5, 0.5])
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(G.Normal(0.0, onoise)))
    SV.solveTree(fg)
    ppe = {l: SV.getPPE(fg, l).suggested[0] for l in fg.variables}
    assert abs(ppe["x1"]) < 1.2 and abs(ppe["x2"]) < 1.2, ppe
    assert abs(ppe["l1"] - l1) < 1.2 and abs(ppe["l2"] - l2) < 1.2, ppe
    assert abs(ppe["l1_0"] - l1) < 10 and abs(ppe["l2_0"] - l2) < 10, ppe


def test_solve_sets_ppe(built):
    """testSolveSetPPE.jl: after solveTree! every variable carries a PPE whose `suggested` is the belief's mean and whose
    `max` lies inside the belief's support; a second solve refreshes it."""
    fg = W.generateGraph_Kaess(N=100)
    fg.solverParams.graphinit = True
    SV.solveTree(fg)
    first = {}
    for l in fg.variables:
        ppe = SV.getPPE(fg, l)
        p = G.getPoints(G.getBelief(fg, l))[:, 0]
        assert abs(ppe.suggested[0] - p.mean()) < 1e-9 and abs(ppe.mean[0] - p.mean()) < 1e-9
        assert p.min() - 1.0 <= ppe.max[0] <= p.max() + 1.0
        first[l] = ppe.suggested[0]
    SV.solveTree(fg)
    assert any(abs(SV.getPPE(fg, l).suggested[0] - first[l]) > 0 for l in fg.variables)   # fresh noise, fresh estimate


@pytest.mark.parametrize("seed", [42, 7, 19])
def test_multimodal_1d(built, seed):
    """testMultimodal1D.jl:34-101: landmark priors at -30 / +30, two sightings from x1 (-20 and +20, sigma 1) with
    multihypo = [1, 0.4, 0.6] between a free landmark and the prior one; gibbsIters = 6, spreadNH = 0.3, given
    elimination order.  x1 sits on one of the two consistent modes, the prior landmarks stay, the free ones keep mass
    near the sighted positions."""
    l1, l2, p_meas = -30.0, 30.0, 0.4
    fg = G.initfg(G.SolverParams(N=100, gibbsIters=6, spreadNH=0.3, seed=seed))
    G.addVariable(fg, "lp1", G.ContinuousScalar)
    G.addFactor(fg, ["lp1"], G.Prior(G.Normal(l1, 1.0)), graphinit=False)
    G.addVariable(fg, "lp2", G.ContinuousScalar)
    G.addFactor(fg, ["lp2"], G.Prior(G.Normal(l2, 1.0)), graphinit=False)
    G.addVariable(fg, "x1", G.ContinuousScalar)
    G.addVariable(fg, "lm2", G.ContinuousScalar)
    G.addFactor(fg, ["x1", "lm2", "lp2"], G.LinearRelative(G.Normal(20.0, 1.0)), multihypo=[1.0, p_meas, 1 - p_meas], graphinit=False)
    G.addVariable(fg, "lm1", G.ContinuousScalar)
    G.addFactor(fg, ["x1", "lm1", "lp1"], G.LinearRelative(G.Normal(-20.0, 1.0)), multihypo=[1.0, p_meas, 1 - p_meas], graphinit=False)
    SV.solveTree(fg, eliminationOrder=["x1", "lm1", "lm2", "lp1", "lp2"])
    n = fg.solverParams.N
    p = {l: G.getPoints(G.getBelief(fg, l))[:, 0] for l in fg.variables}
    assert 0.7 * n < ((-20 < p["x1"]) & (p["x1"] < 0)).sum() + ((0 < p["x1"]) & (p["x1"] < 20)).sum(), np.sort(p["x1"])[::10]
    assert 0.7 * n < ((-38 < p["lp1"]) & (p["lp1"] < -28)).sum()
    assert 0.7 * n < ((28 < p["lp2"]) & (p["lp2"] < 38)).sum()
    assert 0.1 * n < ((-38 < p["lm1"]) & (p["lm1"] < -25)).sum(), np.sort(p["lm1"])[::10]
    assert 0.1 * n < ((25 < p["lm2"]) & (p["lm2"] < 38)).sum(), np.sort(p["lm2"])[::10]


def test_forest_of_orphaned_graphs(built):
    """testSolveOrphanedFG.jl: two disconnected chains in one graph, given elimination order: a forest with two roots
    (the cliques of x1 and x10), one child each, and the reference's mean bands."""
    from iifb200 import tree as TR
    fg = G.initfg(G.SolverParams(N=100, seed=4))
    G.addVariable(fg, "x0", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(0.0, 0.1)))
    G.addVariable(fg, "x1", G.ContinuousScalar)
    G.addFactor(fg, ["x0", "x1"], G.LinearRelative(G.Normal(10.0, 0.1)))
    G.addVariable(fg, "x2", G.ContinuousScalar)
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(G.Normal(10.0, 0.1)))
    G.addVariable(fg, "x10", G.ContinuousScalar)
    G.addFactor(fg, ["x10"], G.Prior(G.Normal(0.0, 1.0)))
    G.addVariable(fg, "x11", G.ContinuousScalar)
    G.addFactor(fg, ["x10", "x11"], G.LinearRelative(G.Normal(-10.0, 1.0)))
    G.addVariable(fg, "x12", G.ContinuousScalar)
    G.addFactor(fg, ["x11", "x12"], G.LinearRelative(G.Normal(-10.0, 1.0)))
    vo = ["x12", "x2", "x0", "x11", "x1", "x10"]
    ts = SV.solveTree(fg, eliminationOrder=vo)
    cl = ts.tree.cliques
    of = lambda v: next(c for c in cl if v in c.frontals)  # noqa: E731
    assert of("x1").parent is None and of("x10").parent is None
    assert len(of("x1").children) == 1 and len(of("x10").children) == 1
    assert len(of("x2").children) == 0 and len(of("x12").children) == 0
    for l, want, band in (("x0", 0, 1.0), ("x1", 10, 2.0), ("x2", 20, 3.0), ("x10", 0, 2.0), ("x11", -10, 4.0), ("x12", -20, 5.0)):
        assert abs(_mean(fg, l) - want) < band, (l, _mean(fg, l))


def test_skip_downsolve_with_msg_likelihoods(built):
    """testSkipUpDown.jl:8-32 (first half): generateGraph_LineStep(6; poseEvery = 1) with the landmark lm0 sighted from
    x0 and x6 only, useMsgLikelihoods = true, downsolve = false: every PPE within 0.2 of the variable's index."""
    fg = W.generateGraph_LineStep(6, poseEvery=1, landmarkEvery=7, posePriorsAt=(0,), landmarkPriorsAt=(), sightDistance=1,
                                  solverParams=G.SolverParams(N=100, seed=6, graphinit=False))
    G.addFactor(fg, ["x6", "lm0"], G.LinearRelative(G.Normal(-6.0, 0.1)), graphinit=False)    # x0lm0f1 came with sightDistance 1
    assert sorted(fg.factors) == sorted(["x0f1", "x0lm0f1", "x6lm0f1"] + [f"x{k}x{k+1}f1" for k in range(6)])
    fg.solverParams.graphinit = True
    fg.solverParams.useMsgLikelihoods = True
    fg.solverParams.downsolve = False
    ts = SV.solveTree(fg)
    assert ts.plan.up_last_wave == len(ts.plan.wave_off) - 1 or all(k == 2 for k, _, _ in ts.plan.sched_waved[ts.plan.wave_off[ts.plan.up_last_wave]:])
    for l in fg.variables:
        want = float(l.lstrip("xlm"))
        assert abs(SV.getPPE(fg, l).suggested[0] - want) < 0.2, (l, SV.getPPE(fg, l).suggested)
