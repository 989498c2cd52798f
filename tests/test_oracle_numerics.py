"""CPU tests that pin the oracle (test infrastructure) before it is trusted as the GPU's checker:
known-answer vectors, the reference's own deterministic tests and analytic roots."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
import parity_cases as PC
from iifb200 import _abi as A
from iifb200 import compile as CP
from iifb200 import graph as G


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    out = (C.c_uint32 * 4)()
    for ctr, key, exp in kat:
        O.lib().iifo_philox4x32(*ctr, *key, out)
        assert tuple(out) == exp


def test_stream_uniform_and_normal_moments():
    u = np.array([O.uniform(42, 7, 3, i) for i in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
    z = np.array([O.normal(42, 7, 1, i) for i in range(20000)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    # streams are independent of each other and reproducible
    assert O.uniform(42, 7, 3, 5) == O.uniform(42, 7, 3, 5) != O.uniform(42, 8, 3, 5)


def test_residual_known_answer_testApproxConv():
    """calcFactorResidualTemporary(lr3, ..., [0;0;0.5], (zeros(3), [0;0;1.0])): sum(abs(res)) ≈ 0.5
    (test/testApproxConv.jl:26-32) and _getZDim == 3 (:24)."""
    res = O.residual(A.F_LINEAR_RELATIVE, [0, 0, 0.5], [np.zeros(3), np.array([0, 0, 1.0])])
    assert len(res) == 3 and abs(np.abs(res).sum() - 0.5) < 1e-10


def test_residual_library_values():
    assert np.allclose(O.residual(A.F_PRIOR, [2.0], [np.array([0.5])]), [1.5])
    assert np.allclose(O.residual(A.F_EUCLID_DISTANCE, [5.0], [np.zeros(2), np.array([3.0, 4.0])]), [0.0])
    # circular: z=0.2, p=3.1, q=-3.1 => qhat = wrap(3.3) = -2.983..; log_q(qhat) = 0.1168..
    r = O.residual(A.F_CIRCULAR_CIRCULAR, [0.2], [np.array([3.1]), np.array([-3.1])], circ_mask=1)
    assert np.allclose(r, [(3.1 + 0.2 - 2 * np.pi) + 3.1])
    r = O.residual(A.F_PRIOR_CIRCULAR, [3.1], [np.array([-3.1])], circ_mask=1)
    assert np.allclose(r, [6.2 - 2 * np.pi])


def test_std_basic_spread():
    R = np.random.default_rng(0)
    x = R.normal(3, 2, (100, 1))
    assert np.isclose(O.std_basic_spread(x), x.std(ddof=1))                    # VariableStatistics.jl:30-31
    x2 = R.normal(0, 1, (50, 2))
    mu = x2.mean(axis=0)
    assert np.isclose(O.std_basic_spread(x2), np.sqrt(((x2 - mu) ** 2).sum() / 49))  # distance-based std
    assert O.std_basic_spread(np.full((10, 1), 4.0)) == 1.0                    # :34 sigma < 1e-10 -> 1.0
    assert O.std_basic_spread(np.zeros((1, 1))) == 1.0
    th = PC.wrap(R.normal(np.pi, 0.2, (80, 1)))                               # straddles the +-pi cut
    assert abs(O.std_basic_spread(th, 1) - 0.2) < 0.05


def test_loo_objective_matches_brute_force_and_golden_is_a_minimum():
    R = np.random.default_rng(1)
    x = R.normal(0, 1, 60)
    for h in (0.05, 0.3, 2.0):
        d = x[:, None] - x[None, :]
        K = np.exp(-d * d / (2 * h * h)) / (np.sqrt(2 * np.pi) * h)
        np.fill_diagonal(K, 0.0)
        ref = -np.mean(np.log(K.sum(axis=1) / (len(x) - 1)))
        assert np.isclose(O.loo_nll(x, h), ref, rtol=1e-12)
    bw = O.kde_bandwidth(x)[0]
    f0 = O.loo_nll(x, bw)
    assert f0 <= O.loo_nll(x, bw * 1.1) and f0 <= O.loo_nll(x, bw / 1.1)
    assert 0.5 < bw / (1.06 * x.std() * len(x) ** -0.2) < 2.0                  # same ballpark as Silverman
    # per-dimension independence and scale equivariance
    x2 = np.stack([x, 10 * x], axis=1)
    b2 = O.kde_bandwidth(x2)
    assert np.isclose(b2[0], bw) and np.isclose(b2[1], 10 * bw, rtol=1e-9)


def _two_var(N=100, nullhypo=0.0, seed=3):
    R = np.random.default_rng(seed)
    P = PC.Problem(seed=seed)
    x0 = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    x1 = P.slot(G.ContinuousScalar, N, R.normal(5, 3, (N, 1)))
    f = P.factor(G.LinearRelative(G.Normal(10.0, 1.0)), [x0, x1], nullhypo=nullhypo)
    return P.freeze(), x0, x1, f


def test_unique_root_is_analytic_and_independent_of_the_inflated_start():
    """for LinearRelative the proposal is x_s[n] + z[n] exactly."""
    P, x0, x1, f = _two_var()
    meas = np.random.default_rng(5).normal(10, 1, 100)
    op = CP.make_conv_ops([dict(factor=f, sfidx=2, N=100, call_id=1, meas_off=0)])[0]
    pts, bw, ipc, lab, nan = P.oracle().conv(op, meas=meas)
    src = P.arena.get(x0)[0][:, 0]
    assert np.abs(pts[:, 0] - (src + meas)).max() < 1e-12 and nan == 0 and np.all(ipc == 1.0)
    op = CP.make_conv_ops([dict(factor=f, sfidx=1, N=100, call_id=2, meas_off=0)])[0]
    pts, *_ = P.oracle().conv(op, meas=meas)
    assert np.abs(pts[:, 0] - (P.arena.get(x1)[0][:, 0] - meas)).max() < 1e-12


def test_point_count_contract_and_no_aliasing():
    """approxConv(...; N=101) returns 101 points (testVariousNSolveSize.jl:18-25) and never mutates
    the target variable (testMultiHypo3Door.jl:59-90)."""
    R = np.random.default_rng(2)
    P = PC.Problem()
    x0 = P.slot(G.ContinuousScalar, 128, R.normal(0, 1, (100, 1)))
    x1 = P.slot(G.ContinuousScalar, 128, R.normal(0, 1, (100, 1)))
    f = P.factor(G.LinearRelative(G.Normal(1.0, 0.1)), [x0, x1])
    P.freeze()
    before = P.arena.copy()
    orc = P.oracle(P.arena)
    pts, *_ = orc.conv(CP.make_conv_ops([dict(factor=f, sfidx=2, N=101, call_id=1)])[0])
    assert pts.shape == (101, 1)
    assert np.array_equal(before.pts, P.arena.pts) and np.array_equal(before.bw, P.arena.bw)


def test_nan_solve_leaves_particle_unchanged():
    """NumericalCalculations.jl:348-351."""
    P, x0, x1, f = _two_var()
    s = P.frozen["slots"][x0]
    P.arena.pts[s.pts_off + 7] = np.nan
    op = CP.make_conv_ops([dict(factor=f, sfidx=2, N=100, call_id=1)])[0]
    pts, bw, ipc, lab, nan = P.oracle().conv(op)
    assert nan == 3  # inflateCycles solves on that particle
    assert np.isfinite(pts).all()


# ---- statistical acceptance bands ported from the reference's tests (>= 20 seeds each) -------------
def _posterior_of_priors(mus, seed):
    R = np.random.default_rng(seed)
    P = PC.Problem(seed=seed)
    x0 = P.slot(G.ContinuousScalar, 100, R.normal(np.mean(mus), 1.0, (100, 1)))
    fs = [P.factor(G.Prior(G.Normal(m, 1.0)), [x0]) for m in mus]
    P.freeze()
    orc = P.oracle()
    orc.propagate(CP.make_prop_ops([dict(target_slot=x0, factors=[(f, 1) for f in fs], N=100, call_id=16)])[0])
    return orc.arena.get(x0)[0][:, 0]


@pytest.mark.parametrize("mus,mean_tol,lo,hi", [
    ([0.0], 0.5, 0.3, 1.9),              # testBasicGraphs.jl:44-47
    ([0.0, 0.0], 0.4, 0.3, 1.0),         # :86-92  (two identical priors)
    ([0.0, 0.0, 0.0], 0.4, 0.1, 0.75),   # :107-113 (three identical priors)
    ([-1.0, 1.0], 0.8, 0.2, 1.5),        # :128-134
    ([-1001.0, -999.0], 0.6, 0.2, 1.1),  # :149-155 (offset by -1000)
])
def test_prior_product_bands_testBasicGraphs(mus, mean_tol, lo, hi):
    ok = 0
    for seed in range(20):
        p = _posterior_of_priors(mus, seed)
        ok += (abs(p.mean() - np.mean(mus)) < mean_tol) and (lo < p.var(ddof=1) < hi)
    assert ok >= 19   # the reference allows a retry for stochastic flakiness (testVariousNSolveSize.jl:17)


def test_nullhypo_prior_band_testnullhypothesis():
    """Prior(Normal(10,1)), nullhypo=0.5 (test/testnullhypothesis.jl:27-41): about half the mass stays
    near the previous belief, half moves to the prior."""
    ok = 0
    for seed in range(20):
        R = np.random.default_rng(seed)
        P = PC.Problem(seed=seed)
        x0 = P.slot(G.ContinuousScalar, 100, R.normal(0, 1, (100, 1)))
        f = P.factor(G.Prior(G.Normal(10.0, 1.0)), [x0], nullhypo=0.5)
        P.freeze()
        pts, bw, ipc, lab, _ = P.oracle().conv(CP.make_conv_ops([dict(factor=f, sfidx=1, N=100, call_id=3)])[0])
        a = ((-15 < pts) & (pts < 4)).sum()
        b = ((4 < pts) & (pts < 16)).sum()
        ok += (10 < a < 60) and (30 < b < 85)
    assert ok >= 19


def test_nullhypo_relative_band():
    """LinearRelative(Normal(10,1)), nullhypo=0.5 (testnullhypothesis.jl:60-66)."""
    ok = 0
    for seed in range(20):
        R = np.random.default_rng(seed)
        P = PC.Problem(seed=seed)
        x0 = P.slot(G.ContinuousScalar, 100, R.normal(0, 1, (100, 1)))
        x1 = P.slot(G.ContinuousScalar, 100, R.normal(0, 1, (100, 1)))
        f = P.factor(G.LinearRelative(G.Normal(10.0, 1.0)), [x0, x1], nullhypo=0.5)
        P.freeze()
        pts, *_ = P.oracle().conv(CP.make_conv_ops([dict(factor=f, sfidx=2, N=100, call_id=3)])[0])
        ok += (20 < (pts < 5).sum()) and (20 < ((5 < pts) & (pts < 15)).sum())
    assert ok >= 19


def test_multihypo_bimodal_band_testmultihypothesisapi():
    """x --(10)--> {la @ -30 | lb @ 40} with multihypo [1, .5, .5]: solving for x gives two modes near
    la-10 and lb-10 (test/testmultihypothesisapi.jl:87-102 analogue)."""
    ok = 0
    for seed in range(20):
        R = np.random.default_rng(seed)
        P = PC.Problem(seed=seed)
        x = P.slot(G.ContinuousScalar, 100, R.normal(0, 1, (100, 1)))
        la = P.slot(G.ContinuousScalar, 100, R.normal(-30, 1, (100, 1)))
        lb = P.slot(G.ContinuousScalar, 100, R.normal(40, 1, (100, 1)))
        f = P.factor(G.LinearRelative(G.Normal(10.0, 1.0)), [x, la, lb], mh=[1.0, 0.5, 0.5])
        P.freeze()
        pts, bw, ipc, lab, _ = P.oracle().conv(CP.make_conv_ops([dict(factor=f, sfidx=1, N=100, call_id=9)])[0])
        m1, m2 = (np.abs(pts + 40) < 6).sum(), (np.abs(pts - 30) < 6).sum()
        ok += (m1 > 20) and (m2 > 20) and (m1 + m2 >= 95) and set(np.unique(lab)) <= {2, 3}
    assert ok >= 19


def test_mixture_prior_band_testMixtureLinearConditional():
    """Mixture prior: >= 20 % of the points in each mode, < 10 % elsewhere
    (test/testMixtureLinearConditional.jl:15-67 style)."""
    R = np.random.default_rng(0)
    P = PC.Problem()
    x0 = P.slot(G.ContinuousScalar, 200, R.normal(0, 1, (200, 1)))
    doors = G.Mixture(G.Prior, [G.Normal(-5, 1), G.Normal(5, 1)], [0.5, 0.5])
    f = P.factor(doors, [x0])
    P.freeze()
    pts, *_ = P.oracle().conv(CP.make_conv_ops([dict(factor=f, sfidx=1, N=200, call_id=2)])[0])
    lo, hi = (np.abs(pts + 5) < 3).mean(), (np.abs(pts - 5) < 3).mean()
    assert lo > 0.2 and hi > 0.2 and 1 - lo - hi < 0.1


def test_ppe_oracle_known_answers():
    """calcPPE restatement (FGOSUtils.jl:237-278; getKDEMax from KDE.jl, parity unpinned): a hand-computable
    two-kernel case, a bimodal case and the circular wrap."""
    # kernels at 0, 0 and 1, h = 0.1: grid [-0.1, 1.1] with 200 points, the density peaks at the grid point
    # nearest to 0
    mean, mx = O.ppe(np.array([[0.0], [0.0], [1.0]]), [0.1])
    step = 1.2 / 199
    assert abs(mean[0] - 1.0 / 3.0) < 1e-15
    assert abs(mx[0] - (-0.1 + round(0.1 / step) * step)) < 1e-12
    R = np.random.default_rng(3)
    x = np.concatenate([R.normal(-2, 0.3, (30, 1)), R.normal(4, 0.3, (70, 1))])
    mean, mx = O.ppe(x, [0.2])
    assert abs(mean[0] - x.mean()) < 1e-12 and abs(mx[0] - 4.0) < 0.3      # the heavier mode
    # circular coordinate: points around +-pi; the mean is the extrinsic (atan2) mean, the max stays in [-pi, pi)
    th = PC.wrap(R.normal(np.pi, 0.2, (80, 1)))
    mean, mx = O.ppe(th, [0.1], circ_mask=1)
    assert abs(abs(mean[0]) - np.pi) < 0.1 and -np.pi <= mx[0] < np.pi
    # two coordinates are treated independently (marginals)
    y = R.normal([1.0, -5.0], [0.5, 2.0], (100, 2))
    mean, mx = O.ppe(y, [0.2, 0.8])
    assert np.allclose(mean, y.mean(axis=0)) and abs(mx[0] - 1.0) < 0.5 and abs(mx[1] + 5.0) < 2.0


def test_deconv_and_mmd_oracle_known_answers():
    """approxDeconv restatement (DeconvUtils.jl:32-162) and mmd (SolverUtilities.jl:25-47 / AMP.mmd!)."""
    # mmd: identical sets -> 0; two single points at distance D -> 2 (1 - exp(-bw D^2))
    R = np.random.default_rng(2)
    a = R.normal(0, 1, (50, 1))
    assert abs(O.mmd(a, a)) < 1e-15
    assert abs(O.mmd(np.array([[0.0]]), np.array([[3.0]]), bw=0.5) - 2 * (1 - np.exp(-0.5 * 9.0))) < 1e-15
    assert O.mmd(a, a + 100.0) > O.mmd(a, a + 1.0) > 0
    # circular distance: points at +-(pi - 0.01) are 0.02 apart on the circle
    assert O.mmd(np.array([[np.pi - 0.01]]), np.array([[-np.pi + 0.01]]), circ_mask=1, bw=1.0) < 1e-3
    # deconv: prior -> the variable's own particles; relative -> particle differences; sampled = factor samples
    P, xs, fs = PC.chain_problem(n=3, N=100, seed=4)
    orc = P.oracle()
    pred, meas = orc.deconv(fs[0], 100, 77)
    x0 = P.arena.get(xs[0])[0]
    assert np.array_equal(pred, x0) and meas.shape == (100, 1) and abs(meas.std() - 0.1) < 0.05
    pred, meas = orc.deconv(fs[1], 100, 78)
    x1 = P.arena.get(xs[1])[0]
    assert np.allclose(pred, x1 - x0, rtol=0, atol=1e-15) and abs(meas.mean() - 1.0) < 0.05
    # the reference's own bands (test/testDefaultDeconv.jl:18-31) on beliefs consistent with the factors
    x0s = R.normal(0.0, 0.1, (100, 1))
    Q = PC.Problem()
    from iifb200 import graph as G
    a0 = Q.slot(G.ContinuousScalar, 100, x0s)
    a1 = Q.slot(G.ContinuousScalar, 100, x0s + R.normal(1.0, 0.1, (100, 1)))
    fp = Q.factor(G.Prior(G.Normal(0.0, 0.1)), [a0])
    fr = Q.factor(G.LinearRelative(G.Normal(1.0, 0.1)), [a0, a1])
    orc = Q.freeze().oracle()
    pred, meas = orc.deconv(fp, 100, 5)
    assert O.mmd(pred, meas) < 1e-6
    pred, meas = orc.deconv(fr, 100, 6)
    assert O.mmd(pred, meas) < 1e-3


def test_se2_residual_and_root_known_answers():
    """SURVEY 8f-4.  ManifoldFactor{SpecialEuclidean(2)} residual (GenericFunctions.jl:39-44, hybrid tangent
    representation as in test/testSpecialEuclidean2Mani.jl:14): p = identity, X = (1, 2, pi/4) puts q at
    t = (1, 2), R(pi/4) (the values the reference test asserts, :62-63); composing once more from there gives
    t = (1,2) + R(pi/4)(1,2), theta = pi/2."""
    X = [1.0, 2.0, np.pi / 4]
    e = np.zeros(3)
    q1 = np.array([1.0, 2.0, np.pi / 4])
    assert np.allclose(O.residual(A.F_SE2_RELATIVE, X, [e, q1], circ_mask=4), 0, atol=1e-15)
    c = np.sqrt(0.5)
    q2 = np.array([1.0 + c * 1 - c * 2, 2.0 + c * 1 + c * 2, np.pi / 2])
    assert np.allclose(O.residual(A.F_SE2_RELATIVE, X, [q1, q2], circ_mask=4), 0, atol=1e-15)
    # a displaced q gives (t_qhat - t_q, wrapped angle difference)
    r = O.residual(A.F_SE2_RELATIVE, X, [e, q1 + [0.1, -0.2, 0.3]], circ_mask=4)
    assert np.allclose(r, [-0.1, 0.2, -0.3])
    # the angle residual wraps across the seam
    r = O.residual(A.F_SE2_RELATIVE, [0, 0, 0.2], [np.array([0, 0, 3.1]), np.array([0, 0, -3.1])], circ_mask=4)
    assert np.allclose(r, [0, 0, (3.3 - 2 * np.pi) + 3.1])
    # ManifoldPrior residual: log(p, m) in coordinates
    r = O.residual(A.F_MANIFOLD_PRIOR, [1.0, 2.0, 3.1], [np.array([0.5, 0.5, -3.1])], circ_mask=4)
    assert np.allclose(r, [0.5, 1.5, 6.2 - 2 * np.pi])


def test_se2_proposals_are_roots_and_deconv_inverts():
    """Every proposal of the SE(2) convolution is a root of the residual for the sample's own measurement (both
    solve directions), and approxDeconv recovers the measurement that was used."""
    R = np.random.default_rng(8)
    N = 64
    se = G.SpecialEuclidean2
    P = PC.Problem()
    a = np.column_stack([R.normal(0, 2, N), R.normal(0, 2, N), PC.wrap(R.normal(3.0, 1.0, N))])
    b = np.column_stack([R.normal(0, 2, N), R.normal(0, 2, N), PC.wrap(R.normal(-1.0, 1.0, N))])
    s0, s1 = P.slot(se, N, a), P.slot(se, N, b)
    f = P.factor(G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, np.pi / 4], np.diag([0.04, 0.04, 0.01]))), [s0, s1])
    P.freeze()
    meas = R.normal([1.0, 2.0, 0.8], [0.2, 0.2, 0.1], (N, 3))
    orc = P.oracle()
    op = CP.make_conv_ops([dict(factor=f, sfidx=2, N=N, call_id=3, meas_off=0),
                           dict(factor=f, sfidx=1, N=N, call_id=4, meas_off=0)])
    q, *_ = orc.conv(op[0], meas.reshape(-1))
    p, *_ = orc.conv(op[1], meas.reshape(-1))
    for n in range(N):
        assert np.allclose(O.residual(A.F_SE2_RELATIVE, meas[n], [a[n], q[n]], circ_mask=4), 0, atol=1e-12)
        assert np.allclose(O.residual(A.F_SE2_RELATIVE, meas[n], [p[n], b[n]], circ_mask=4), 0, atol=1e-12)
    assert np.all(np.abs(q[:, 2]) <= np.pi) and np.all(np.abs(p[:, 2]) <= np.pi)
    # deconv on (a, q): predicted == the measurements (angles modulo 2 pi)
    orc.arena.set(s1, q, None, True)
    pred, _ = orc.deconv(f, N, 5)
    d = pred - meas
    d[:, 2] = PC.wrap(d[:, 2])
    assert np.abs(d).max() < 1e-12


def test_se2_reference_bands_testSpecialEuclidean2Mani():
    """test/testSpecialEuclidean2Mani.jl:35-63 on the oracle: ManifoldPrior at the identity (sigma 0.01) initialises
    x0 near the identity; ManifoldFactor MvNormal([1,2,pi/4], 0.01) initialises x1 near (1, 2, R(pi/4)); atol 0.1.
    :113-123: a PartialPrior on dims (1,2) of an SE(2) variable gives a partial belief with 3 infoPerCoord."""
    se = G.SpecialEuclidean2
    N = 100
    ok = 0
    for seed in range(20):
        P = PC.Problem(seed=seed)
        x0 = P.slot(se, N, np.zeros((0, 3)), initialized=False)
        x1 = P.slot(se, N, np.zeros((0, 3)), initialized=False)
        fp = P.factor(G.ManifoldPrior(se, [0.0, 0.0, 0.0], G.MvNormal([0, 0, 0], np.diag([1e-4] * 3))), [x0])
        ff = P.factor(G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, np.pi / 4], np.diag([0.01] * 3))), [x0, x1])
        P.freeze()
        orc = P.oracle()
        pr = CP.make_prop_ops([dict(target_slot=x0, out_slot=x0, factors=[(fp, 1)], N=N, call_id=16),
                               dict(target_slot=x1, out_slot=x1, factors=[(ff, 2)], N=N, call_id=32)])
        orc.propagate(pr[0])
        orc.propagate(pr[1])
        a, _, _ = orc.arena.get(x0)
        b, bw, ipc = orc.arena.get(x1)
        m1 = np.array([b[:, 0].mean(), b[:, 1].mean(), np.arctan2(np.sin(b[:, 2]).mean(), np.cos(b[:, 2]).mean())])
        ok += (np.abs(a.mean(axis=0)).max() < 0.1) and (np.abs(m1 - [1.0, 2.0, np.pi / 4]).max() < 0.1) \
            and np.all(bw > 0) and np.array_equal(ipc, [1.0, 1.0, 1.0])
    assert ok == 20
    P = PC.Problem()
    x0 = P.slot(se, N, np.zeros((0, 3)), initialized=False)
    fpp = P.factor(G.PartialPrior(G.MvNormal([0.01, 0.01], np.eye(2) * 1e-4), (1, 2)), [x0])
    P.freeze()
    pts, bw, ipc, _, _ = P.oracle().conv(CP.make_conv_ops([dict(factor=fpp, sfidx=1, N=N, call_id=1)])[0])
    assert np.array_equal(ipc, [1.0, 1.0, 0.0]) and np.all(pts[:, 2] == 0.0) and abs(pts[:, 0].mean() - 0.01) < 0.01


def test_chain_solve_bands_testProductReproducable():
    """test/testProductReproducable.jl:10-41: a..e with Prior(Normal()) and LinearRelative(Normal(10, 1)); after
    initAll! + solveTree! the means sit at 0, 10, .., 40 (|err| < 3, 4, 4, 5, 5) and the spreads stay inside
    (0.3, 2), (0.5, 4), (0.9, 6), (1.2, 7), (1.5, 8).  >= 19 of 20 seeds."""
    ok = 0
    for seed in range(20):
        fg = G.initfg(G.SolverParams(graphinit=False, seed=seed))
        for l in "abcde":
            G.addVariable(fg, l, G.ContinuousScalar)
        G.addFactor(fg, ["a"], G.Prior(G.Normal()))
        for u, v in zip("abcd", "bcde"):
            G.addFactor(fg, [u, v], G.LinearRelative(G.Normal(10.0, 1.0)))
        PC.oracle_initAll(fg)
        PC.oracle_solveTree(fg)
        good = True
        for k, (l, em, lo, hi) in enumerate(zip("abcde", (3, 4, 4, 5, 5), (0.3, 0.5, 0.9, 1.2, 1.5), (2, 4, 6, 7, 8))):
            p = fg.variables[l].val[:, 0]
            good &= abs(p.mean() - 10.0 * k) < em and lo < p.std(ddof=1) < hi
        ok += good
    assert ok >= 19, ok


def test_back_and_forth_convolution_spreads_testProductReproducable():
    """test/testProductReproducable.jl:52-98: ten rounds of approxConv a -> b -> a over LinearRelative(Normal(10, 1)):
    the means stay put (|a| < 2, |b - 10| < 2) and the spreads grow beyond 3."""
    R = np.random.default_rng(4)
    P = PC.Problem()
    a = P.slot(G.ContinuousScalar, 100, R.normal(0, 1, (100, 1)))
    b = P.slot(G.ContinuousScalar, 100, R.normal(10, 1, (100, 1)))
    f = P.factor(G.LinearRelative(G.Normal(10.0, 1.0)), [a, b])
    P.freeze()
    orc = P.oracle()
    A0, B0 = orc.arena.get(a)[0].copy(), orc.arena.get(b)[0].copy()
    for i in range(10):
        for sf, slot in ((2, b), (1, a)):
            pts, bw, *_ = orc.conv(CP.make_conv_ops([dict(factor=f, sfidx=sf, N=100, call_id=100 + 2 * i + sf)])[0])
            orc.arena.set(slot, pts, bw, True)               # initVariable!(fg, :b, manikde!(pts))
    A1, B1 = orc.arena.get(a)[0], orc.arena.get(b)[0]
    assert abs(A0.mean()) < 1 and abs(A1.mean()) < 2 and abs(B0.mean() - 10) < 1 and abs(B1.mean() - 10) < 2
    assert A0.std(ddof=1) < 2 and 3 < A1.std(ddof=1) and B0.std(ddof=1) < 2 and 3 < B1.std(ddof=1)


def test_forward_convolve_testBasicForwardConvolve():
    """test/testBasicForwardConvolve.jl:13-66 (#477): X0 ~ N(0, 0.1) -> LinearRelative(Normal(11, 1)) -> product with a
    measurement belief N(9.5, 0.75) -> LinearRelative(Normal(8, 2)); 100 points, 15 < mean < 25."""
    ok = 0
    for seed in range(20):
        R = np.random.default_rng(seed)
        P = PC.Problem(seed=seed)
        x0 = P.slot(G.ContinuousScalar, 100, R.normal(0, 0.1, (100, 1)))
        x1 = P.slot(G.ContinuousScalar, 100, np.zeros((0, 1)), initialized=False)
        x2 = P.slot(G.ContinuousScalar, 100, np.zeros((0, 1)), initialized=False)
        f1 = P.factor(G.LinearRelative(G.Normal(11.0, 1.0)), [x0, x1])
        f2 = P.factor(G.LinearRelative(G.Normal(8.0, 2.0)), [x1, x2])
        P.freeze()
        orc = P.oracle()
        X1_, bw1, *_ = orc.conv(CP.make_conv_ops([dict(factor=f1, sfidx=2, N=100, call_id=1)])[0])
        meas = R.normal(9.5, 0.75, (100, 1))
        X1, bwp, _ = O.product(np.stack([X1_, meas]), np.stack([bw1, O.kde_bandwidth(meas)]), 1, seed=seed, call_id=2)
        orc.arena.set(x1, X1, bwp, True)
        X2, *_ = orc.conv(CP.make_conv_ops([dict(factor=f2, sfidx=2, N=100, call_id=3)])[0])
        ok += X2.shape == (100, 1) and 15 < X2.mean() < 25
    assert ok == 20


def test_euclid_distance_bands_testEuclidDistance():
    """test/testEuclidDistance.jl:10-77: x0 with a unit prior, x1 tied by EuclidDistance(Normal(10, 1)).
    1-D: the solve-for belief is two-sided (> 30 % beyond +5, > 30 % below -5, < 10 % in between);
    2-D: more than half of the points on the ring 7 < |x1| < 13; x0's estimate stays within 1 of the origin."""
    ok1 = ok2 = 0
    for seed in range(20):
        fg = G.initfg(G.SolverParams(graphinit=False, seed=seed))
        G.addVariable(fg, "x0", G.ContinuousScalar)
        G.addFactor(fg, ["x0"], G.Prior(G.Normal()))
        G.addVariable(fg, "x1", G.ContinuousScalar)
        G.addFactor(fg, ["x0", "x1"], G.EuclidDistance(G.Normal(10.0, 1.0)))
        PC.oracle_initAll(fg)
        PC.oracle_solveTree(fg)
        p = fg.variables["x1"].val[:, 0]
        n = len(p)
        ok1 += abs(fg.variables["x0"].val.mean()) < 1 and (p > 5).sum() > 0.3 * n and (p < -5).sum() > 0.3 * n \
            and ((p > -5) & (p < 5)).sum() < 0.1 * n
        fg = G.initfg(G.SolverParams(graphinit=False, seed=seed))
        G.addVariable(fg, "x0", G.Position(2))
        G.addFactor(fg, ["x0"], G.Prior(G.MvNormal(np.zeros(2), np.eye(2))))
        G.addVariable(fg, "x1", G.Position(2))
        G.addFactor(fg, ["x0", "x1"], G.EuclidDistance(G.Normal(10.0, 1.0)))
        PC.oracle_initAll(fg)
        PC.oracle_solveTree(fg)
        r = np.linalg.norm(fg.variables["x1"].val, axis=1)
        ok2 += np.abs(fg.variables["x0"].val.mean(axis=0)).max() < 1 and ((r > 7) & (r < 13)).sum() > 0.5 * len(r)
    assert ok1 >= 18 and ok2 >= 18, (ok1, ok2)


def test_partial_prior_bands_testpartialconstraint():
    """test/testpartialconstraint.jl:53-124: x1 in R^2 with a full prior N(0, 0.01 I) and a partial prior on
    coordinate 1 ~ N(2, 1).  The full prior's proposal has mean[1] within 0.3 of 0; the partial proposal moves
    coordinate 1 (mean within 0.75 of 2, far from the current values) and leaves coordinate 2 EXACTLY at the current
    values, without touching the variable; the belief is partial; solving the graph keeps both coordinates within 0.4
    of 0."""
    ok = 0
    for seed in range(20):
        fg = G.initfg(G.SolverParams(graphinit=False, seed=seed))
        G.addVariable(fg, "x1", G.Position(2))
        G.addFactor(fg, ["x1"], G.Prior(G.MvNormal([0.0, 0.0], np.diag([0.01, 0.01]))), label="x1f1")
        PC.oracle_initAll(fg)                                     # doautoinit! before the partial factor exists (:63-65)
        G.addFactor(fg, ["x1"], G.PartialPrior(G.Normal(2.0, 1.0), (1,)), label="x1f2")
        P = PC.Problem(sp=fg.solverParams, seed=seed)
        X1 = fg.variables["x1"].val.copy()
        s1 = P.slot(G.Position(2), 100, X1, fg.variables["x1"].bw)
        f1 = P.factor(fg.factors["x1f1"].fnc, [s1])
        f2 = P.factor(fg.factors["x1f2"].fnc, [s1])
        P.freeze()
        orc = P.oracle()
        ops = CP.make_conv_ops([dict(factor=f1, sfidx=1, N=100, call_id=1), dict(factor=f2, sfidx=1, N=100, call_id=2)])
        full, *_ = orc.conv(ops[0])
        part, bw, ipc, _, _ = orc.conv(ops[1])
        good = full.shape == (100, 2) and abs(full[:, 0].mean()) < 0.3
        good &= abs(part[:, 0].mean() - 2.0) < 0.75 and np.linalg.norm(X1[:, 0] - part[:, 0]) > 2.0
        good &= np.linalg.norm(X1[:, 1] - part[:, 1]) < 1e-10                 # untouched coordinate
        good &= np.array_equal(orc.arena.get(s1)[0], X1)                      # the variable itself is not modified
        good &= np.array_equal(ipc, [1.0, 0.0])                               # isPartial(X1_)
        PC.oracle_solveTree(fg)
        m = fg.variables["x1"].val.mean(axis=0)
        good &= abs(m[0]) < 0.4 and abs(m[1]) < 0.4
        ok += bool(good)
    assert ok >= 19, ok


def test_mixture_prior_balance_testMixturePrior():
    """test/testMixturePrior.jl:11-68 (#605): Mixture(Prior, [Normal(-5, 1), Normal(0, 1)], [0.5; 0.5]) — the proposal and
    the solved marginal keep both modes in balance: |#(x < -2.5) - #(x > -2.5)| < 0.35 N.  (The reference's last variant
    swaps the second component for an AliasingScalarSampler, which has no device sampler.)"""
    ok = 0
    for seed in range(20):
        fg = G.initfg(G.SolverParams(graphinit=False, seed=seed, N=100))
        G.addVariable(fg, "x0", G.ContinuousScalar)
        G.addFactor(fg, ["x0"], G.Mixture(G.Prior, [G.Normal(-5.0, 1.0), G.Normal(0.0, 1.0)], [0.5, 0.5]))
        PC.oracle_initAll(fg)                                              # approxConv(fg, :x0f1, :x0)
        p = fg.variables["x0"].val[:, 0]
        good = abs(int((p < -2.5).sum()) - int((p > -2.5).sum())) < 35
        PC.oracle_solveTree(fg)
        q = fg.variables["x0"].val[:, 0]
        good &= abs(int((q < -2.5).sum()) - int((q > -2.5).sum())) < 35
        ok += bool(good)
    assert ok >= 19, ok


@pytest.mark.parametrize("nullhypo", [0.0, 0.2])
def test_n_dimensional_partials_testPartialNH(nullhypo):
    """test/testPartialNH.jl:9-90: x0, x1 in R^3; PartialPrior on coordinates (2, 3) of x0 ~ N(0, I), PartialPrior on
    coordinate 1 of x1 ~ N(10, 1), LinearRelative(MvNormal([10, 0, 0], I)); with and without nullhypo = 0.2 on the x0
    prior and the relative.  After initAll! + solveTree!: x0 within 1 of [0, 0, 0], x1 within 1 (2 with nullhypo) of
    [10, 0, 0].  The products mix full and partial proposals (coordinates nobody informs keep their old values)."""
    ok = 0
    for seed in range(20):
        fg = G.initfg(G.SolverParams(graphinit=False, seed=seed, N=100))
        G.addVariable(fg, "x0", G.Position(3))
        G.addFactor(fg, ["x0"], G.PartialPrior(G.MvNormal(np.zeros(2), np.eye(2)), (2, 3)), nullhypo=nullhypo)
        G.addVariable(fg, "x1", G.Position(3))
        G.addFactor(fg, ["x1"], G.PartialPrior(G.Normal(10.0, 1.0), (1,)))
        G.addFactor(fg, ["x0", "x1"], G.LinearRelative(G.MvNormal([10.0, 0.0, 0.0], np.eye(3))), nullhypo=nullhypo)
        PC.oracle_initAll(fg)
        PC.oracle_solveTree(fg)
        m0, m1 = fg.variables["x0"].val.mean(axis=0), fg.variables["x1"].val.mean(axis=0)
        ok += np.abs(m0).max() < 1 and np.abs(m1 - [10.0, 0.0, 0.0]).max() < (1 if nullhypo == 0 else 2)
    assert ok >= 18, ok


def test_translation_group_manifold_factors_testTranslationMani():
    """test/testTranslationMani.jl:7-38 (a smoke test there): ManifoldPrior(TranslationGroup(2), [10, 20], MvNormal([1, 1]))
    and ManifoldFactor(TranslationGroup(2), MvNormal([1, 2], [0.1, 0.1])) initialise and solve; here also located:
    x0 near [10, 20], x1 near [11, 22]."""
    M = G.TranslationGroup(2)
    fg = G.initfg(G.SolverParams(graphinit=False, seed=2))
    G.addVariable(fg, "x0", M)
    G.addVariable(fg, "x1", M)
    G.addFactor(fg, ["x0"], G.ManifoldPrior(M, [10.0, 20.0], G.MvNormal([0.0, 0.0], np.eye(2))))
    f = G.addFactor(fg, ["x0", "x1"], G.ManifoldFactor(M, G.MvNormal([1.0, 2.0], np.diag([0.1, 0.1]))))
    assert f.fnc.kind == A.F_LINEAR_RELATIVE and fg.factors["x0f1"].fnc.kind == A.F_MANIFOLD_PRIOR
    PC.oracle_initAll(fg)
    PC.oracle_solveTree(fg)
    assert np.abs(fg.variables["x0"].val.mean(axis=0) - [10.0, 20.0]).max() < 0.5
    assert np.abs(fg.variables["x1"].val.mean(axis=0) - [11.0, 22.0]).max() < 0.7


def test_special_orthogonal2_testSpecialOrthogonalMani():
    """test/testSpecialOrthogonalMani.jl:12-58: ManifoldPrior(SpecialOrthogonal(2), I, MvNormal([0.01])) initialises x0
    at the identity (atol 0.1); ManifoldFactor(SpecialOrthogonal(2), MvNormal([pi], [0.01])) puts x1 half a turn away;
    the tree solve runs.  SO(2) points are carried as their angle."""
    M = G.SpecialOrthogonal2
    fg = G.initfg(G.SolverParams(graphinit=False, seed=4))
    G.addVariable(fg, "x0", M)
    G.addFactor(fg, ["x0"], G.ManifoldPrior(M, G.so2_point_to_coords(np.eye(2)), G.Normal(0.0, 0.01)))
    PC.oracle_initAll(fg)
    th = fg.variables["x0"].val[:, 0]
    Rm = np.array([[np.cos(th).mean(), -np.sin(th).mean()], [np.sin(th).mean(), np.cos(th).mean()]])
    assert np.allclose(Rm, np.eye(2), atol=0.1)
    G.addVariable(fg, "x1", M)
    f = G.addFactor(fg, ["x0", "x1"], G.ManifoldFactor(M, G.Normal(np.pi, 0.1)))
    assert f.fnc.kind == A.F_CIRCULAR_CIRCULAR
    PC.oracle_initAll(fg)
    PC.oracle_solveTree(fg)
    t0, t1 = fg.variables["x0"].val[:, 0], fg.variables["x1"].val[:, 0]
    assert np.all(np.abs(t1) <= np.pi) and abs(np.arctan2(np.sin(t0).mean(), np.cos(t0).mean())) < 0.1
    assert np.cos(t1).mean() < -0.9                                   # x1 sits at +-pi (the seam)


def test_so3_known_answers_and_bands_testSpecialOrthogonalMani():
    """SURVEY 8f-4, test/testSpecialOrthogonalMani.jl:75-140 (SpecialOrthogonal(3)): proposals of the relative factor are
    roots of vee(log(q, p Exp(X))) (checked with rotation MATRICES built by numpy, independent of the oracle's
    quaternions); ManifoldPrior at the identity with MvNormal([0.01, 0.01, 0.01]) gives mean(M, pts) ~ I (atol 0.01);
    ManifoldFactor MvNormal([0.01, 0.01, 0.01], [0.01, 0.01, 0.01]) puts x1 at Exp([0.01, 0.01, 0.01]) before and after
    solveTree!; every point is a valid rotation (|omega| <= pi)."""
    name, P, specs, _ = [c for c in PC.conv_cases() if c[0] == "so3"][0]
    orc = P.oracle()
    ops = CP.make_conv_ops(specs)
    w0, w1 = P.arena.get(0)[0], P.arena.get(1)[0]
    e2m = G.so3_coords_to_point
    # forward: q = p Exp(z) => Exp(w0)^T q is a rotation of angle |z| ~ |(0.3, 0.1, -0.2)|; backward likewise
    q = orc.conv(ops[1])[0]
    z = np.array([G.so3_point_to_coords(e2m(w0[n]).T @ e2m(q[n])) for n in range(len(q))])
    assert np.abs(z.mean(axis=0) - [0.3, 0.1, -0.2]).max() < 0.05 and np.abs(z.std(axis=0) - [0.1, 0.1, 0.1414]).max() < 0.04
    p = orc.conv(ops[2])[0]
    z = np.array([G.so3_point_to_coords(e2m(p[n]).T @ e2m(w1[n])) for n in range(len(p))])
    assert np.abs(z.mean(axis=0) - [0.3, 0.1, -0.2]).max() < 0.05
    # prior: p Exp(z) around the point with rotation vector (0.2, -0.1, 0.4)
    pr = orc.conv(ops[0])[0]
    z = np.array([G.so3_point_to_coords(e2m([0.2, -0.1, 0.4]).T @ e2m(pr[n])) for n in range(len(pr))])
    assert np.abs(z.mean(axis=0)).max() < 0.05 and np.abs(z.std(axis=0) - [0.1, 0.1414, 0.1]).max() < 0.04
    # null hypothesis: ~30 % of the particles keep their (entropy-spread) value, the rest are roots
    qn, _, _, lab, _ = orc.conv(ops[3])
    assert 10 < (lab == 0).sum() < 55
    # numeric solve lands on the closed-form root within Nelder-Mead's accuracy
    qm = orc.conv(ops[4])[0]
    ref = P.oracle().conv(CP.make_conv_ops([dict(specs[4], factor=specs[1]["factor"])])[0])[0]
    assert np.abs(qm - ref).max() < 5e-3 and np.abs(qm - ref).max() > 0
    assert (np.linalg.norm(q, axis=1) <= np.pi + 1e-12).all()
    # the reference's test graph, 6 seeds
    so3 = G.SpecialOrthogonal3
    for seed in range(6):
        fg = G.initfg(G.SolverParams(seed=seed, graphinit=False))
        G.addVariable(fg, "x0", so3)
        G.addFactor(fg, ["x0"], G.ManifoldPrior(so3, np.eye(3), G.MvNormal([0, 0, 0], np.diag([1e-4] * 3))))
        PC.oracle_initAll(fg)
        assert np.abs(G.so3_mean(fg.variables["x0"].val) - np.eye(3)).max() < 0.01
        G.addVariable(fg, "x1", so3)
        G.addFactor(fg, ["x0", "x1"], G.ManifoldFactor(so3, G.MvNormal([0.01, 0.01, 0.01], np.diag([1e-4] * 3))))
        PC.oracle_initAll(fg)
        want = np.array([[0.9999, -0.00995, 0.01005], [0.01005, 0.9999, -0.00995], [-0.00995, 0.01005, 0.9999]])
        assert np.abs(G.so3_mean(fg.variables["x1"].val) - want).max() < 0.01
        PC.oracle_solveTree(fg)
        assert np.abs(G.so3_mean(fg.variables["x0"].val) - np.eye(3)).max() < 0.01
        assert np.abs(G.so3_mean(fg.variables["x1"].val) - want).max() < 0.01
        assert (np.linalg.norm(fg.variables["x1"].val, axis=1) <= np.pi).all()


def test_sample_table_distribution():
    """any host-side distribution as a measurement (sampleFactor! = rand(Z), SolverUtilities.jl:50-76): Rayleigh(2)
    range noise through a prior and a relative factor keeps its mean sigma sqrt(pi/2) and its skew"""
    name, P, specs, _ = [c for c in PC.conv_cases() if c[0] == "sample_table"][0]
    orc = P.oracle()
    ops = CP.make_conv_ops(specs)
    pri = orc.conv(ops[1])[0][:, 0]
    assert abs(pri.mean() - 2.0 * np.sqrt(np.pi / 2)) < 0.5 and pri.min() > 0 and abs(pri.std() - 2.0 * np.sqrt(2 - np.pi / 2)) < 0.45
    rel = orc.conv(ops[0])[0][:, 0] - P.arena.get(0)[0][:, 0]          # x1 - x0 = z for every particle (label 1)
    assert rel.min() > 0 and abs(rel.mean() - 2.0 * np.sqrt(np.pi / 2)) < 0.5
    g2 = P.arena.get(3)[0] - orc.conv(ops[2])[0]                        # y1 - y0' = z (solving the first variable)
    assert (g2 > 0).all() and np.abs(g2.mean(axis=0) - 1.0).max() < 0.3


def test_numeric_solver_nelder_mead():
    """SURVEY a10b (_solveLambdaNumeric, NumericalCalculations.jl:49-133): the restated Optim.NelderMead.  Unique-root
    factors forced through it land on the analytic root within the optimiser's own accuracy (its stopping rule
    sqrt(var f) < 1e-8 leaves residuals of ~1e-4, which is why the reference's tests use loose bands); EuclidDistance
    in 2-D / 3-D lands ON the ring / sphere z = |x2 - x1| at the point the simplex walks to from the inflated start
    (the radial projection is only its 1-D special case)."""
    name, P, specs, _ = [c for c in PC.conv_cases() if c[0] == "numeric_solve"][0]
    orc = P.oracle()
    ops = CP.make_conv_ops(specs)
    fz = P.frozen
    res = [orc.conv(ops[k]) for k in range(len(specs))]
    # EuclidDistance: every proposal sits on the ring around its partner particle (labels all 1 => partner n)
    for k in (0, 1, 2):
        f = fz["factors"][specs[k]["factor"]]
        d = fz["slots"][f.slot[0]].dim
        other = f.slot[0] if specs[k]["sfidx"] == 2 else f.slot[1]
        po = P.arena.get(other)[0]
        dist = np.linalg.norm(res[k][0] - po, axis=1)
        mu = 6.0 if d == 2 else 5.0
        assert np.all(np.abs(dist - mu) < 5 * (0.2 if d == 2 else 0.1) + 1e-2), k        # z ~ Normal(mu, sigma)
        assert dist.std() > 0.01 and res[k][4] == 0
    # unique roots: the same convolutions with the closed form (solver = 0) agree within Nelder-Mead's accuracy
    P2 = PC.Problem()
    import copy
    T2 = copy.deepcopy(P.T)
    for f in T2.factors:
        f["solver"] = 0
    fz2 = T2.freeze()
    orc2 = O.Oracle(fz2, P.arena.copy(), P.sp_c)
    for k in (3, 4, 5):
        ref = orc2.conv(ops[k])[0]
        dlt = np.abs(res[k][0] - ref)
        if k == 5:
            dlt[:, 2] = np.minimum(dlt[:, 2], 2 * np.pi - dlt[:, 2])
        assert dlt.max() < 2e-3 and dlt.max() > 0, (k, dlt.max())                          # close, but not the closed form
    del P2


def test_four_equal_peaks_testMultiHypo3Door():
    """test/testMultiHypo3Door.jl:40-99: x0 sees one of four doors (landmarks at 0, 10, 20, 40, sigma 0.01) through a
    5-ary LinearRelative(Normal(0, 0.25)) with multihypo = [1, 1/4, 1/4, 1/4, 1/4], N = 200: the proposal has four
    peaks, its KDE exceeds 0.1 at every door (:96-99), and the convolution leaves x0 itself untouched (:80-90)."""
    def dens(pts, bw, x):
        return np.mean(np.exp(-0.5 * ((x - pts) / bw) ** 2) / (np.sqrt(2 * np.pi) * bw))
    ok = 0
    for seed in range(20):
        R = np.random.default_rng(seed)
        P = PC.Problem(sp=G.SolverParams(N=200), seed=seed)
        ls = [P.slot(G.ContinuousScalar, 200, R.normal(m, 0.01, (200, 1))) for m in (0.0, 10.0, 20.0, 40.0)]
        x0 = P.slot(G.ContinuousScalar, 200, np.zeros((0, 1)), initialized=False)
        f = P.factor(G.LinearRelative(G.Normal(0.0, 0.25)), [x0] + ls, mh=[1.0, 0.25, 0.25, 0.25, 0.25])
        P.freeze()
        orc = P.oracle()
        pts, bw, ipc, lab, _ = orc.conv(CP.make_conv_ops([dict(factor=f, sfidx=1, N=200, call_id=3)])[0])
        ok += all(dens(pts[:, 0], bw[0], l) > 0.1 for l in (0.0, 10.0, 20.0, 40.0)) and orc.arena.npts[x0] == 0 \
            and set(np.unique(lab)) == {2, 3, 4, 5}
    assert ok >= 19, ok
