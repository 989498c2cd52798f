"""Shared parity cases: the same seeded problem is run through the CPU oracle and through the
C-ABI of libiifb200.so; results must agree (labels bit-exact, points/bandwidths within TOL).

Used by tests/test_gpu_parity.py (`-m gpu`) and by __graft_entry__.smoke().
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import iifb200  # noqa: E402
from iifb200 import _abi as A  # noqa: E402
from iifb200 import compile as CP  # noqa: E402
from iifb200 import graph as G  # noqa: E402

# FP tolerance of GPU vs oracle for closed-form (unique-root) proposals and for posterior points
# whose Gibbs labels agree: the two sides differ only by FMA contraction / libm ulps / reduction
# order, amplified at most by the bandwidth-search conditioning.
TOL_PTS = 1e-9
TOL_BW = 1e-7


class Problem:
    """tables + host arena (+ engine / oracle built lazily on the same data)."""

    def __init__(self, sp=None, seed=42):
        self.T = CP.Tables()
        self.sp = sp or G.SolverParams()
        self.seed = seed
        self.frozen = None
        self.arena = None
        self._init = []

    def slot(self, vartype, cap, pts=None, bw=None, initialized=True):
        s = self.T.add_slot(vartype, cap)
        if pts is not None:
            self._init.append((s, np.asarray(pts, dtype=np.float64), bw, initialized))
        return s

    def factor(self, fnc, slots, mh=None, nullhypo=0.0, inflation=5.0):
        mhp = None
        if mh is not None:
            mhp, _ = G.parseusermultihypo(list(mh), nullhypo)
        return self.T.add_factor(fnc, slots, mhp, nullhypo, inflation)

    def freeze(self):
        self.frozen = self.T.freeze()
        self.arena = CP.HostArena(self.frozen)
        import oracle as O
        for s, pts, bw, init in self._init:
            sd = self.frozen["slots"][s]
            if bw is None and pts.size:
                bw = O.kde_bandwidth(pts.reshape(-1, sd.dim), sd.circ_mask)
            self.arena.set(s, pts, bw, init)
        self.sp_c = CP.solver_params_c(self.sp, self.seed)
        return self

    def oracle(self, arena=None):
        import oracle as O
        return O.Oracle(self.frozen, arena if arena is not None else self.arena.copy(), self.sp_c)

    def engine(self, device=0):
        from iifb200.engine import Engine
        e = Engine(self.frozen, self.sp_c, device)
        e.upload_arena(self.arena)
        return e


def rng(seed):
    return np.random.default_rng(seed)


# ------------------------------------------------------------------------------ convolution cases
def conv_cases():
    """yield (name, Problem, [conv op spec dicts], explicit streams dict or None)"""
    R = rng(7)
    N = 100
    # 1. scalar chain: Prior + LinearRelative (C1/C2 building block, testBasicGraphs.jl:325-343)
    P = Problem()
    x0 = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    x1 = P.slot(G.ContinuousScalar, N, R.normal(0.5, 2, (N, 1)))
    fp = P.factor(G.Prior(G.Normal(0.0, 1.0)), [x0])
    fr = P.factor(G.LinearRelative(G.Normal(1.0, 0.1)), [x0, x1])
    yield "scalar_prior_and_relative", P.freeze(), [
        dict(factor=fp, sfidx=1, N=N, call_id=10), dict(factor=fr, sfidx=2, N=N, call_id=20),
        dict(factor=fr, sfidx=1, N=N, call_id=30)], None

    # 2. nullhypo on prior and relative (testnullhypothesis.jl)
    P = Problem()
    x0 = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    x1 = P.slot(G.ContinuousScalar, N, R.normal(10, 1, (N, 1)))
    fp = P.factor(G.Prior(G.Normal(0.0, 1.0)), [x0], nullhypo=0.5)
    fr = P.factor(G.LinearRelative(G.Normal(10.0, 1.0)), [x0, x1], nullhypo=0.5)
    yield "nullhypo", P.freeze(), [dict(factor=fp, sfidx=1, N=N, call_id=11),
                                   dict(factor=fr, sfidx=2, N=N, call_id=12),
                                   dict(factor=fr, sfidx=2, N=N, call_id=13, nullSurplus=0.3)], None

    # 3. bi-modal multihypo [1, .5, .5], solving for each variable (testmultihypothesisapi.jl)
    P = Problem()
    x = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    la = P.slot(G.ContinuousScalar, N, R.normal(-30, 1, (N, 1)))
    lb = P.slot(G.ContinuousScalar, N, R.normal(40, 1, (N, 1)))
    fm = P.factor(G.LinearRelative(G.Normal(10.0, 1.0)), [x, la, lb], mh=[1.0, 0.5, 0.5])
    yield "multihypo_bimodal", P.freeze(), [dict(factor=fm, sfidx=k, N=N, call_id=40 + k) for k in (1, 2, 3)], None

    # 4. four-door style 5-ary multihypo [1, 1/4 x4], N=200 (testMultiHypo3Door.jl:40-57)
    N2 = 200
    P = Problem()
    xs = [P.slot(G.ContinuousScalar, N2, R.normal(0, 1, (N2, 1)))]
    for m in (0.0, 10.0, 20.0, 40.0):
        xs.append(P.slot(G.ContinuousScalar, N2, R.normal(m, 0.01, (N2, 1))))
    f5 = P.factor(G.LinearRelative(G.Normal(0.0, 0.25)), xs, mh=[1.0, 0.25, 0.25, 0.25, 0.25])
    yield "multihypo_fourdoor", P.freeze(), [dict(factor=f5, sfidx=k, N=N2, call_id=50 + k) for k in (1, 2, 5)], None

    # 5. circular manifold (testCircular.jl)
    Nc = 150
    P = Problem()
    c0 = P.slot(G.Circular, Nc, wrap(R.normal(3.0, 0.3, (Nc, 1))))
    c1 = P.slot(G.Circular, Nc, wrap(R.normal(-2.5, 0.5, (Nc, 1))))
    fpc = P.factor(G.PriorCircular(G.Normal(3.0, 0.1)), [c0])
    fcc = P.factor(G.CircularCircular(G.Normal(1.0, 0.1)), [c0, c1], nullhypo=0.1)
    yield "circular", P.freeze(), [dict(factor=fpc, sfidx=1, N=Nc, call_id=60),
                                   dict(factor=fcc, sfidx=2, N=Nc, call_id=61),
                                   dict(factor=fcc, sfidx=1, N=Nc, call_id=62)], None

    # 6. Euclid(2) with MvNormal, and EuclidDistance on 2-D points (non-unique root)
    P = Problem()
    p0 = P.slot(G.Position(2), N, R.normal(0, 1, (N, 2)))
    p1 = P.slot(G.Position(2), N, R.normal(0, 1, (N, 2)) + [5.0, -3.0])
    cov = np.array([[0.04, 0.01], [0.01, 0.09]])
    f2 = P.factor(G.LinearRelative(G.MvNormal([5.0, -3.0], cov)), [p0, p1])
    fpr = P.factor(G.Prior(G.MvNormal([0.0, 0.0], np.eye(2) * 0.01)), [p0])
    fd = P.factor(G.EuclidDistance(G.Normal(6.0, 0.2)), [p0, p1])
    yield "euclid2", P.freeze(), [dict(factor=f2, sfidx=2, N=N, call_id=70), dict(factor=f2, sfidx=1, N=N, call_id=71),
                                  dict(factor=fpr, sfidx=1, N=N, call_id=72), dict(factor=fd, sfidx=2, N=N, call_id=73)], None

    # 7. Mixture prior (fourdoortest.jl:9-54) and Mixture relative (testMixtureLinearConditional.jl)
    P = Problem()
    x0 = P.slot(G.ContinuousScalar, N2, R.normal(0, 1, (N2, 1)))
    x1 = P.slot(G.ContinuousScalar, N2, R.normal(0, 1, (N2, 1)))
    doors = G.Mixture(G.Prior, [G.Normal(-100, 3), G.Normal(0, 3), G.Normal(100, 3), G.Normal(300, 3)], [0.25] * 4)
    fmx = P.factor(doors, [x0])
    mr = G.Mixture(G.LinearRelative, [G.Normal(-5, 0.5), G.Normal(5, 0.5)], [0.5, 0.5])
    fmr = P.factor(mr, [x0, x1])
    yield "mixture", P.freeze(), [dict(factor=fmx, sfidx=1, N=N2, call_id=80), dict(factor=fmr, sfidx=2, N=N2, call_id=81)], None

    # 8. MsgPrior on a device-resident belief (separator message), incl. a circular one
    P = Problem()
    m0 = P.slot(G.ContinuousScalar, N, R.normal(4, 1, (N, 1)))
    t0 = P.slot(G.ContinuousScalar, N, R.normal(0, 3, (N, 1)))
    fmsg = P.factor(G.MsgPrior(G.SlotRef(m0, 1)), [t0])
    mc = P.slot(G.Circular, N, wrap(R.normal(3.1, 0.2, (N, 1))))
    tc = P.slot(G.Circular, N, wrap(R.normal(0, 1, (N, 1))))
    fmsgc = P.factor(G.MsgPrior(G.SlotRef(mc, 1)), [tc])
    yield "msgprior", P.freeze(), [dict(factor=fmsg, sfidx=1, N=N, call_id=90), dict(factor=fmsgc, sfidx=1, N=N, call_id=91)], None

    # 9. explicit host streams (Julia-drawn labels / measurements / inflation uniforms)
    P = Problem()
    x = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    la = P.slot(G.ContinuousScalar, N, R.normal(-30, 1, (N, 1)))
    lb = P.slot(G.ContinuousScalar, N, R.normal(40, 1, (N, 1)))
    fm = P.factor(G.LinearRelative(G.Normal(10.0, 1.0)), [x, la, lb], mh=[1.0, 0.5, 0.5])
    meas = R.normal(10.0, 1.0, N)
    lab = R.integers(2, 4, N).astype(np.int32)
    uinf = R.random(4 * N)
    yield "explicit_streams", P.freeze(), [dict(factor=fm, sfidx=1, N=N, call_id=95, meas_off=0, mhidx_off=0, uinf_off=0)], \
        dict(meas=meas, mhidx=lab, uinf=uinf)

    # 10. uninitialised / short target and source with fewer points (resize + _getindex_anyn)
    P = Problem()
    s0 = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (60, 1)))
    s1 = P.slot(G.ContinuousScalar, N, np.zeros((0, 1)), initialized=False)
    frs = P.factor(G.LinearRelative(G.Normal(2.0, 0.1)), [s0, s1])
    yield "ragged", P.freeze(), [dict(factor=frs, sfidx=2, N=N, call_id=97)], None

    # 11. partial prior on dim 2 of a 2-D variable (testPartialPrior.jl)
    P = Problem()
    q0 = P.slot(G.Position(2), N, R.normal(0, 1, (N, 2)))
    fpp = P.factor(G.PartialPrior(G.Normal(-20.0, 1.0), (2,)), [q0])
    yield "partial_prior", P.freeze(), [dict(factor=fpp, sfidx=1, N=N, call_id=98)], None

    # 12. extreme sizes: N=2 (minimum) and N=256 (IIF_MAX_POINTS)
    for Nx in (2, 3, 256):
        P = Problem()
        a = P.slot(G.ContinuousScalar, Nx, R.normal(0, 1, (Nx, 1)), bw=[0.5])
        b = P.slot(G.ContinuousScalar, Nx, R.normal(0, 1, (Nx, 1)), bw=[0.5])
        fr = P.factor(G.LinearRelative(G.Normal(1.0, 0.1)), [a, b])
        yield f"size_{Nx}", P.freeze(), [dict(factor=fr, sfidx=2, N=Nx, call_id=99)], None


    # 13. SpecialEuclidean(2) (testSpecialEuclidean2Mani.jl): ManifoldPrior, ManifoldFactor solved both
    # ways, a partial prior on the translation and on the angle, nullhypo entropy on the manifold
    P = Problem()
    se = G.SpecialEuclidean2
    p0 = np.column_stack([R.normal(0, 0.5, N), R.normal(0, 0.5, N), wrap(R.normal(3.0, 0.4, N))])
    p1 = np.column_stack([R.normal(1, 0.5, N), R.normal(2, 0.5, N), wrap(R.normal(-2.6, 0.4, N))])
    s0, s1 = P.slot(se, N, p0), P.slot(se, N, p1)
    fmp = P.factor(G.ManifoldPrior(se, [0.5, -0.5, 3.1], G.MvNormal([0, 0, 0], np.diag([0.01, 0.02, 0.09]))), [s0])
    fmf = P.factor(G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, np.pi / 4], np.diag([0.01, 0.01, 0.01]))), [s0, s1])
    fmn = P.factor(G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, np.pi / 4], np.diag([0.01, 0.01, 0.01]))), [s0, s1],
                   nullhypo=0.3)
    fpt = P.factor(G.PartialPrior(G.MvNormal([0.01, 0.01], np.eye(2) * 0.01), (1, 2)), [s0])
    fpa = P.factor(G.ManifoldPriorPartial(se, G.Normal(3.3, 0.1), (3,)), [s0])
    yield "se2_circ", P.freeze(), [dict(factor=fmp, sfidx=1, N=N, call_id=110), dict(factor=fmf, sfidx=2, N=N, call_id=111),
                                   dict(factor=fmf, sfidx=1, N=N, call_id=112), dict(factor=fmn, sfidx=2, N=N, call_id=113),
                                   dict(factor=fpt, sfidx=1, N=N, call_id=114), dict(factor=fpa, sfidx=1, N=N, call_id=115)], None


    # 14. numeric per-sample solve: Nelder-Mead restated from Optim as _solveLambdaNumeric configures it —
    # EuclidDistance in 2-D and 3-D (ring / sphere of roots: the result is the optimiser's path from the inflated
    # start), and unique-root factors forced through the numeric solve (solver = 1)
    P = Problem()
    p0 = P.slot(G.Position(2), N, R.normal(0, 1, (N, 2)))
    p1 = P.slot(G.Position(2), N, R.normal(0, 1, (N, 2)) + [5.0, -3.0])
    q0 = P.slot(G.Position(3), N, R.normal(0, 1, (N, 3)))
    q1 = P.slot(G.Position(3), N, R.normal(0, 1, (N, 3)) + [2.0, 2.0, -4.0])
    fd2 = P.factor(G.EuclidDistance(G.Normal(6.0, 0.2)), [p0, p1])
    fd3 = P.factor(G.EuclidDistance(G.Normal(5.0, 0.1)), [q0, q1])
    lr = G.LinearRelative(G.MvNormal([5.0, -3.0], np.diag([0.04, 0.09])))
    lr.numeric = True
    fnm = P.factor(lr, [p0, p1])
    se = G.SpecialEuclidean2
    e0 = P.slot(se, N, np.column_stack([R.normal(0, 0.5, N), R.normal(0, 0.5, N), wrap(R.normal(0.3, 0.2, N))]))
    e1 = P.slot(se, N, np.column_stack([R.normal(1, 0.5, N), R.normal(2, 0.5, N), wrap(R.normal(1.0, 0.2, N))]))
    mf = G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, np.pi / 4], np.diag([0.01, 0.01, 0.01])))
    mf.numeric = True
    fse = P.factor(mf, [e0, e1])
    yield "numeric_solve", P.freeze(), [dict(factor=fd2, sfidx=2, N=N, call_id=120), dict(factor=fd2, sfidx=1, N=N, call_id=121),
                                        dict(factor=fd3, sfidx=2, N=N, call_id=122), dict(factor=fnm, sfidx=2, N=N, call_id=123),
                                        dict(factor=fnm, sfidx=1, N=N, call_id=124), dict(factor=fse, sfidx=2, N=N, call_id=125)], None


    # 15. measurement distributions beyond the parametric ones: a host-drawn sample table (here Rayleigh range noise
    # and a skewed 2-D Gamma offset), resampled on the device
    P = Problem()
    x0 = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    x1 = P.slot(G.ContinuousScalar, N, R.normal(3, 1, (N, 1)))
    y0 = P.slot(G.Position(2), N, R.normal(0, 1, (N, 2)))
    y1 = P.slot(G.Position(2), N, R.normal(2, 1, (N, 2)))
    ray = G.SampledBelief(R.rayleigh(2.0, 2048))
    gam = G.SampledBelief(R.gamma(2.0, 0.5, (1024, 2)))
    fr1 = P.factor(G.LinearRelative(ray), [x0, x1])
    fp1 = P.factor(G.Prior(ray), [x0])
    fr2 = P.factor(G.LinearRelative(gam), [y0, y1])
    yield "sample_table", P.freeze(), [dict(factor=fr1, sfidx=2, N=N, call_id=130), dict(factor=fp1, sfidx=1, N=N, call_id=131),
                                       dict(factor=fr2, sfidx=1, N=N, call_id=132)], None


    # 16. SpecialOrthogonal(3) (test/testSpecialOrthogonalMani.jl:75-140): rotation-vector coordinates; prior
    # p Exp(z), relative q = p Exp(X) solved both ways, null-hypothesis entropy on the manifold, and the relative factor
    # through the numeric solve
    P = Problem()
    so3 = G.SpecialOrthogonal3
    w0 = R.normal(0, 0.4, (N, 3)) + [0.3, -0.2, 0.5]
    w1 = R.normal(0, 0.4, (N, 3)) + [-0.4, 0.6, 0.1]
    r0, r1 = P.slot(so3, N, w0), P.slot(so3, N, w1)
    fp3 = P.factor(G.ManifoldPrior(so3, [0.2, -0.1, 0.4], G.MvNormal([0, 0, 0], np.diag([0.01, 0.02, 0.01]))), [r0])
    fr3 = P.factor(G.ManifoldFactor(so3, G.MvNormal([0.3, 0.1, -0.2], np.diag([0.01, 0.01, 0.02]))), [r0, r1])
    fn3 = P.factor(G.ManifoldFactor(so3, G.MvNormal([0.3, 0.1, -0.2], np.diag([0.01, 0.01, 0.02]))), [r0, r1], nullhypo=0.3)
    mfn = G.ManifoldFactor(so3, G.MvNormal([0.3, 0.1, -0.2], np.diag([0.01, 0.01, 0.02])))
    mfn.numeric = True
    fnm3 = P.factor(mfn, [r0, r1])
    yield "so3", P.freeze(), [dict(factor=fp3, sfidx=1, N=N, call_id=140), dict(factor=fr3, sfidx=2, N=N, call_id=141),
                              dict(factor=fr3, sfidx=1, N=N, call_id=142), dict(factor=fn3, sfidx=2, N=N, call_id=143),
                              dict(factor=fnm3, sfidx=2, N=N, call_id=144)], None


def wrap(a):
    return (np.asarray(a) + np.pi) % (2 * np.pi) - np.pi


def run_conv_case(P, specs, streams, engine=None):
    """returns list of (oracle result, gpu result) tuples"""
    ops = CP.make_conv_ops(specs)
    orc = P.oracle()
    st = streams or {}
    o_res = [orc.conv(ops[k], st.get("meas"), st.get("mhidx"), st.get("uinf")) for k in range(len(specs))]
    eng = engine or P.engine()
    g_res = eng.conv_batch(ops, len(specs), st.get("meas"), st.get("mhidx"), st.get("uinf"))
    if engine is None:
        eng.close()
    return list(zip(o_res, g_res))


def assert_conv_equal(name, pairs, circ=False):
    for k, (o, g) in enumerate(pairs):
        op, obw, oipc, olab, onan = o
        gp, gbw, gipc, glab, gnan = g
        assert np.array_equal(olab, glab), f"{name}[{k}]: hypothesis labels differ"
        dp = np.abs(op - gp)
        if circ:
            dp = np.minimum(dp, 2 * np.pi - dp)
        assert dp.max() <= TOL_PTS * max(1.0, np.abs(op).max()), f"{name}[{k}]: points differ by {dp.max()}"
        assert np.allclose(obw, gbw, rtol=TOL_BW, atol=0), f"{name}[{k}]: bandwidth {obw} vs {gbw}"
        assert np.array_equal(oipc, gipc), f"{name}[{k}]: ipc differ"
        assert onan == gnan


# ------------------------------------------------------------------------------ product cases
def product_cases():
    """yield (name, dict(dens_pts F x N x d, dens_bw F x d, dim, circ_mask, dens_mask, old_pts, call_id))"""
    import oracle as O
    R = rng(11)
    N = 100

    def bws(pts, cm=0):
        return np.stack([O.kde_bandwidth(p, cm) for p in pts])

    a = np.stack([R.normal(0, 1, (N, 1)), R.normal(0.5, 1, (N, 1))])
    yield "two_gaussians_1d", dict(dens_pts=a, dens_bw=bws(a), dim=1, call_id=5)
    a = np.stack([R.normal(0, 1, (N, 1)), R.normal(1, 2, (N, 1)), R.normal(-1, 0.5, (N, 1))])
    yield "three_gaussians_1d", dict(dens_pts=a, dens_bw=bws(a), dim=1, call_id=6)
    # bimodal x unimodal (a multihypo proposal times a prior)
    bi = np.concatenate([R.normal(-30, 1, (N // 2, 1)), R.normal(40, 1, (N // 2, 1))])
    a = np.stack([bi, R.normal(38, 3, (N, 1))])
    yield "bimodal_times_prior", dict(dens_pts=a, dens_bw=bws(a), dim=1, call_id=7)
    a = np.stack([R.normal(0, 1, (N, 2)), R.normal(0, 1, (N, 2)) * [0.3, 2.0] + [0.5, -0.5]])
    yield "two_gaussians_2d", dict(dens_pts=a, dens_bw=bws(a), dim=2, call_id=8)
    a = np.stack([R.normal(0, 1, (N, 2)) for _ in range(5)])
    yield "five_gaussians_2d", dict(dens_pts=a, dens_bw=bws(a), dim=2, call_id=9)
    c = np.stack([wrap(R.normal(3.0, 0.3, (150, 1))), wrap(R.normal(-3.0, 0.4, (150, 1)))])
    yield "circular_wraparound", dict(dens_pts=c, dens_bw=bws(c, 1), dim=1, circ_mask=1, call_id=10)
    # partial: density 1 informs only dim 2; dims nobody informs keep oldPoints
    a = np.stack([R.normal(0, 1, (N, 2)), R.normal(5, 1, (N, 2))])
    yield "partial_mask", dict(dens_pts=a, dens_bw=bws(a), dim=2, dens_mask=[0, 2], old_pts=R.normal(9, 1, (N, 2)), call_id=11)
    a = np.stack([R.normal(0, 1, (N, 2)), R.normal(5, 1, (N, 2))])
    yield "partial_uncovered_dim", dict(dens_pts=a, dens_bw=bws(a), dim=2, dens_mask=[2, 2], old_pts=R.normal(9, 1, (N, 2)), call_id=12)
    a = np.stack([R.normal(0, 1, (N, 1))])
    yield "single_passthrough", dict(dens_pts=a, dens_bw=bws(a), dim=1, call_id=13)
    a = np.stack([R.normal(0, 1, (256, 1)), R.normal(0.2, 1, (256, 1))])
    yield "max_points_256", dict(dens_pts=a, dens_bw=bws(a), dim=1, call_id=14)
    a = np.stack([R.normal(0, 1, (7, 1)), R.normal(0.2, 1, (7, 1))])
    yield "tiny_7", dict(dens_pts=a, dens_bw=bws(a), dim=1, call_id=15)
    # explicit Gibbs streams (AMP _randU / _randN keyword vectors)
    a = np.stack([R.normal(0, 1, (N, 1)), R.normal(0.5, 1, (N, 1))])
    yield "explicit_gibbs_streams", dict(dens_pts=a, dens_bw=bws(a), dim=1, call_id=16,
                                         randU=R.random(N * 2 * (1 + 7 * 2)), randN=R.normal(0, 1, N * 8))   # F*(1 + L*(Niter+1)) and d*(L+1) per sample, L = 7

    # SpecialEuclidean(2) coordinates (x, y, theta): two Euclid + one circular coordinate across the +-pi seam
    a = np.stack([np.column_stack([R.normal(1, 0.3, N), R.normal(2, 0.3, N), wrap(R.normal(3.1, 0.2, N))]),
                  np.column_stack([R.normal(1.1, 0.4, N), R.normal(1.9, 0.2, N), wrap(R.normal(-3.1, 0.3, N))])])
    yield "se2_circ_product", dict(dens_pts=a, dens_bw=bws(a, 4), dim=3, circ_mask=4, call_id=17)


def run_product_case(case, engine):
    import oracle as O
    seed = int(engine.sp_c.seed)  # the engine's Philox key drives the Gibbs streams
    kw = dict(case)
    dim = kw.pop("dim")
    o = O.product(kw["dens_pts"], kw["dens_bw"], dim, kw.get("circ_mask", 0), kw.get("dens_mask"),
                  kw.get("old_pts"), seed, kw.get("call_id", 0), 1, kw.get("randU"), kw.get("randN"))
    g = engine.product(kw["dens_pts"], kw["dens_bw"], dim, kw.get("circ_mask", 0), kw.get("dens_mask"),
                       kw.get("old_pts"), kw.get("call_id", 0), kw.get("randU"), kw.get("randN"))
    return o, g


def assert_product_equal(name, o, g, circ=False):
    op, obw, olab = o
    gp, gbw, glab = g
    same = np.all(olab == glab, axis=1)
    # Gibbs labels are picked by inverse CDF on identical uniforms; a flip needs u within ~1e-15 of a
    # boundary, so demand (essentially) all of them and exact agreement where they match
    assert same.mean() >= 0.99, f"{name}: only {same.mean():.3f} of Gibbs label rows agree"
    dp = np.abs(op - gp)[same]
    if circ:
        dp = np.minimum(dp, 2 * np.pi - dp)
    assert dp.max() <= TOL_PTS * max(1.0, np.abs(op).max()), f"{name}: points differ by {dp.max()}"
    if same.all():
        assert np.allclose(obw, gbw, rtol=TOL_BW, atol=0), f"{name}: bandwidth {obw} vs {gbw}"


# ------------------------------------------------------------------------------ propagate / schedule
def chain_problem(n=4, N=100, seed=3, circular=False):
    """x0..x_{n-1}, prior on x0, relative between neighbours; every variable initialised."""
    R = rng(seed)
    P = Problem(seed=seed)
    vt = G.Circular if circular else G.ContinuousScalar
    xs = []
    for k in range(n):
        pts = R.normal(k, 0.5 + 0.1 * k, (N, 1))
        xs.append(P.slot(vt, N, wrap(pts) if circular else pts))
    fs = [P.factor((G.PriorCircular if circular else G.Prior)(G.Normal(0.0, 0.1)), [xs[0]])]
    for k in range(n - 1):
        rel = G.CircularCircular if circular else G.LinearRelative
        fs.append(P.factor(rel(G.Normal(1.0, 0.1)), [xs[k], xs[k + 1]]))
    return P.freeze(), xs, fs


def chain_prop_specs(xs, fs, N, call0=1000):
    """one propagateBelief per variable (all its factors), each into its own slot"""
    specs = []
    n = len(xs)
    for k in range(n):
        fl = []
        if k == 0:
            fl.append((fs[0], 1))
        if k > 0:
            fl.append((fs[k], 2))
        if k < n - 1:
            fl.append((fs[k + 1], 1))
        specs.append(dict(target_slot=xs[k], out_slot=xs[k], factors=fl, N=N, call_id=call0 + 16 * k))
    return specs


def assert_arena_equal(name, ao, ag, frozen, slots, circ=False):
    for s in slots:
        po, bo, io = ao.get(s)
        pg, bg, ig = ag.get(s)
        assert po.shape == pg.shape, f"{name}: slot {s} point count {po.shape} vs {pg.shape}"
        dp = np.abs(po - pg)
        if circ:
            dp = np.minimum(dp, 2 * np.pi - dp)
        frac = (dp.max(axis=1) <= TOL_PTS * max(1.0, np.abs(po).max())).mean() if po.size else 1.0
        assert frac >= 0.99, f"{name}: slot {s}: only {frac:.3f} of posterior points agree (max diff {dp.max()})"
        if frac == 1.0:
            assert np.allclose(bo, bg, rtol=TOL_BW), f"{name}: slot {s} bw {bo} vs {bg}"
        assert np.array_equal(io, ig), f"{name}: slot {s} ipc {io} vs {ig}"


def run_smoke():
    """one small propagateBelief (prior + relative -> product) on cuda:0 vs the oracle"""
    P, xs, fs = chain_problem(n=3, N=64, seed=5)
    specs = chain_prop_specs(xs, fs, 64)[:1]  # x0: prior + relative => 2 convolutions + product
    props = CP.make_prop_ops(specs)
    orc = P.oracle()
    orc.propagate(props[0])
    eng = P.engine()
    eng.propagate_batch(props, 1)
    ag = P.arena.copy()
    eng.download_arena(ag)
    assert eng.launch_count() >= 2
    eng.close()
    assert_arena_equal("smoke", orc.arena, ag, P.frozen, [xs[0]])


# ------------------------------------------------------------------------------ oracle-backed host flows (CPU tests)
def oracle_initAll(fg):
    """initAll! (GraphInit.jl:495-556) with the ORACLE as compute: the sequential doautoinit! sweep of
    iifb200.solver.initAll(batched=False), for CPU tests of the reference's acceptance bands."""
    import oracle as O
    T = CP.Tables()
    N = fg.solverParams.N
    var_slot = {l: T.add_slot(v.vartype, max(N, v.val.shape[0], 1)) for l, v in fg.variables.items()}
    fac_idx = {l: T.add_factor(f.fnc, [var_slot[v] for v in f.variables], f.multihypo, f.nullhypo, f.inflation)
               for l, f in fg.factors.items()}
    frozen = T.freeze()
    arena = CP.HostArena(frozen)
    for l, v in fg.variables.items():
        if v.initialized:
            arena.set(var_slot[l], v.val, v.bw, True, v.infoPerCoord)
    orc = O.Oracle(frozen, arena, CP.solver_params_c(fg.solverParams))
    for _ in range(len(fg.variables) + 1):
        did = False
        for l, v in fg.variables.items():
            if v.initialized:
                continue
            use = [f for f in fg.listNeighbors(l) if G.factorCanInitFromOtherVars(fg, f, l)]
            if not use:
                continue
            spec = dict(target_slot=var_slot[l], out_slot=var_slot[l], N=N, call_id=fg.next_call(),
                        factors=[(fac_idx[f], fg.factors[f].variables.index(l) + 1) for f in use],
                        any_multihypo=int(any(G.isMultihypo(fg.factors[f]) for f in use)))
            orc.propagate(CP.make_prop_ops([spec])[0])
            v.val, v.bw, v.infoPerCoord = arena.get(var_slot[l])
            v.initialized = True
            did = True
        if not did:
            break


def oracle_solveTree(fg, order=None, ordering="qr", **kw):
    """solveTree! with the ORACLE as compute (same plan the device runs); posteriors are written back to `fg`."""
    import oracle as O
    from iifb200 import tree as TR
    order = order or TR.getEliminationOrder(fg, ordering)
    tree = TR.buildTree(fg, order)
    plan = TR.compile_solve(fg, tree, **kw)
    TR.rebase_calls(plan, fg.next_call(TR.plan_call_span(plan)))     # as solver.solveTree does (call_base="auto")
    arena = CP.HostArena(plan.frozen)
    for l, v in fg.variables.items():
        arena.set(plan.var_slot[l], v.val, v.bw, True, v.infoPerCoord)
    orc = O.Oracle(plan.frozen, arena, CP.solver_params_c(fg.solverParams))
    orc.schedule_run(plan.wave_off, CP.make_sched_ops(plan.sched_waved), CP.make_prop_ops(plan.props),
                     deconvs=CP.make_deconv_ops(plan.deconvs or []))
    for l, v in fg.variables.items():
        v.val, v.bw, v.infoPerCoord = arena.get(plan.var_slot[l])
    return tree, plan, arena
