"""GPU parity tests proper: the CUDA path (through the C-ABI of libiifb200.so) against the CPU
oracle on identical seeded inputs.  Labels bit-exact; points within parity_cases.TOL_PTS."""
import numpy as np
import pytest

import parity_cases as PC
from iifb200 import compile as CP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", list(PC.conv_cases()), ids=lambda c: c[0])
def test_conv_parity(built, case):
    name, P, specs, streams = case
    pairs = PC.run_conv_case(P, specs, streams)
    PC.assert_conv_equal(name, pairs, circ=("circ" in name or "msgprior" in name))


@pytest.fixture(scope="module")
def bare_engine(built):
    P, xs, fs = PC.chain_problem(n=2, N=8)
    eng = P.engine()
    yield eng
    eng.close()


@pytest.mark.parametrize("case", list(PC.product_cases()), ids=lambda c: c[0])
def test_product_parity(bare_engine, case):
    name, kw = case
    o, g = PC.run_product_case(kw, bare_engine)
    PC.assert_product_equal(name, o, g, circ="circ" in name)


def test_kde_bandwidth_parity(bare_engine):
    import oracle as O
    R = np.random.default_rng(5)
    for n, d, cm in [(100, 1, 0), (100, 2, 0), (150, 1, 1), (200, 3, 0b010), (256, 1, 0), (2, 1, 0), (17, 4, 0)]:
        pts = R.normal(0, 1, (n, d)) * R.uniform(0.1, 10, d)
        for c in range(d):
            if (cm >> c) & 1:
                pts[:, c] = PC.wrap(pts[:, c])
        assert np.allclose(O.kde_bandwidth(pts, cm), bare_engine.kde_bandwidth(pts, cm), rtol=PC.TOL_BW)


@pytest.mark.parametrize("circular", [False, True])
def test_propagate_batch_parity(built, circular):
    P, xs, fs = PC.chain_problem(n=5, N=100, seed=9, circular=circular)
    specs = PC.chain_prop_specs(xs, fs, 100)
    # independent ops only: even variables (their factors read odd variables) into their own slots
    specs = [s for k, s in enumerate(specs) if k % 2 == 0]
    props = CP.make_prop_ops(specs)
    orc = P.oracle()
    for k in range(len(specs)):
        orc.propagate(props[k])
    eng = P.engine()
    eng.propagate_batch(props, len(specs))
    ag = P.arena.copy()
    eng.download_arena(ag)
    eng.close()
    PC.assert_arena_equal("propagate", orc.arena, ag, P.frozen, [s["out_slot"] for s in specs], circ=circular)


def test_schedule_parity_gauss_seidel(built):
    """3 Gibbs sweeps over a 4-variable chain (fmcmc!, SolveTree.jl:112-134): each propagate sees the
    previous one's write (Gauss-Seidel), captured as one CUDA graph."""
    P, xs, fs = PC.chain_problem(n=4, N=100, seed=21)
    specs, sched, wave_off = [], [], [0]
    for sweep in range(3):
        for s in PC.chain_prop_specs(xs, fs, 100, call0=5000 + 1000 * sweep):
            specs.append(s)
            sched.append((CP.A.S_PROPAGATE, len(specs) - 1, 0))
            wave_off.append(len(sched))
    props = CP.make_prop_ops(specs)
    ops = CP.make_sched_ops(sched)
    orc = P.oracle()
    orc.schedule_run(wave_off, ops, props)
    eng = P.engine()
    sid = eng.schedule_build(wave_off, ops, len(sched), props, len(specs))
    eng.schedule_run(sid)
    eng.sync()
    ag = P.arena.copy()
    eng.download_arena(ag)
    n_launch = eng.launch_count()
    eng.close()
    # one conv kernel per wave + one product kernel per wave that holds a multi-factor propagate
    assert len(sched) <= n_launch <= 2 * len(sched)
    PC.assert_arena_equal("schedule", orc.arena, ag, P.frozen, xs)


def test_conv_does_not_mutate_target(built):
    """approxConv must not change the target variable (testMultiHypo3Door.jl:59-90, ApproxConv.jl:17)."""
    name, P, specs, streams = next(PC.conv_cases())
    eng = P.engine()
    eng.conv_batch(CP.make_conv_ops(specs), len(specs))
    after = P.arena.copy()
    eng.download_arena(after)
    eng.close()
    assert np.array_equal(after.pts, P.arena.pts) and np.array_equal(after.bw, P.arena.bw)


def test_unsupported_and_bad_arguments_fail_loudly(built):
    from iifb200._abi import IIFB200Error
    P, xs, fs = PC.chain_problem(n=2, N=16)
    eng = P.engine()
    with pytest.raises(IIFB200Error):
        eng.conv_batch(CP.make_conv_ops([dict(factor=99, sfidx=1, N=16, call_id=1)]), 1)
    with pytest.raises(IIFB200Error):
        eng.conv_batch(CP.make_conv_ops([dict(factor=fs[1], sfidx=3, N=16, call_id=1)]), 1)
    with pytest.raises(IIFB200Error):
        eng.conv_batch(CP.make_conv_ops([dict(factor=fs[1], sfidx=2, N=1000, call_id=1)]), 1)
    eng.close()


def test_ppe_parity(built):
    """calcPPE: device mean / KDE-max of resident beliefs against the oracle.  The grid is
    identical on both sides, so the max agrees exactly unless two grid points tie to the last ulp."""
    import oracle as O
    from iifb200 import graph as G
    R = np.random.default_rng(12)
    P = PC.Problem()
    cases = [(G.ContinuousScalar, R.normal(3, 1, (100, 1))),
             (G.ContinuousScalar, np.concatenate([R.normal(-2, 0.3, (30, 1)), R.normal(4, 0.3, (70, 1))])),
             (G.Position(2), R.normal([1.0, -5.0], [0.5, 2.0], (150, 2))),
             (G.Circular, PC.wrap(R.normal(3.0, 0.3, (200, 1)))),
             (G.ContinuousScalar, R.normal(0, 1e-3, (17, 1)) + 1000.0)]
    slots = [P.slot(vt, len(x), x) for vt, x in cases]
    P.freeze()
    eng = P.engine()
    mean, mx = eng.ppe_batch(slots)
    eng.close()
    step_tol = 1.2 / 199
    for k, (vt, x) in enumerate(cases):
        s = slots[k]
        bw = P.arena.bw[4 * s:4 * s + vt.dim]
        om, ox = O.ppe(x, bw, vt.circ_mask)
        assert np.allclose(mean[k, :vt.dim], om, rtol=0, atol=1e-11 * max(1.0, np.abs(x).max())), (k, mean[k], om)
        rng_ = x.max(axis=0) - x.min(axis=0)
        assert np.all(np.abs(mx[k, :vt.dim] - ox) <= 1e-12 + 0 * rng_) or \
            np.all(np.abs(mx[k, :vt.dim] - ox) <= step_tol * rng_ * 1.0001), (k, mx[k], ox)


def test_deconv_and_mmd_parity(built):
    """approxDeconv and mmd on the device against the oracle (same Philox streams)."""
    import oracle as O
    from iifb200 import graph as G
    R = np.random.default_rng(31)
    P = PC.Problem()
    N = 100
    x0 = P.slot(G.ContinuousScalar, N, R.normal(0, 1, (N, 1)))
    x1 = P.slot(G.ContinuousScalar, N, R.normal(1, 1, (N, 1)))
    c0 = P.slot(G.Circular, N, PC.wrap(R.normal(3.0, 0.5, (N, 1))))
    c1 = P.slot(G.Circular, N, PC.wrap(R.normal(-2.5, 0.5, (N, 1))))
    p0 = P.slot(G.Position(2), 60, R.normal(0, 1, (60, 2)))          # shorter than N: random partner (_getindex_anyn)
    p1 = P.slot(G.Position(2), N, R.normal(5, 1, (N, 2)))
    se = G.SpecialEuclidean2
    e0 = P.slot(se, N, np.column_stack([R.normal(0, 1, (N, 2)), PC.wrap(R.normal(3.0, 0.5, N))]))
    e1 = P.slot(se, N, np.column_stack([R.normal(2, 1, (N, 2)), PC.wrap(R.normal(-2.0, 0.5, N))]))
    fs = [P.factor(G.Prior(G.Normal(0.0, 1.0)), [x0]),
          P.factor(G.LinearRelative(G.Normal(1.0, 0.1)), [x0, x1]),
          P.factor(G.CircularCircular(G.Normal(0.5, 0.1)), [c0, c1]),
          P.factor(G.PriorCircular(G.Normal(3.0, 0.1)), [c0]),
          P.factor(G.LinearRelative(G.MvNormal([5.0, 5.0], np.diag([0.1, 0.2]))), [p0, p1]),
          P.factor(G.EuclidDistance(G.Normal(7.0, 0.1)), [p0, p1]),
          P.factor(G.ManifoldFactor(se, G.MvNormal([1.0, 2.0, 0.7], np.diag([0.01] * 3))), [e0, e1]),
          P.factor(G.ManifoldPrior(se, [0.1, 0.2, 3.0], G.MvNormal([0, 0, 0], np.diag([0.01] * 3))), [e0])]
    P.freeze()
    orc, eng = P.oracle(), P.engine()
    for k, f in enumerate(fs):
        po, mo = orc.deconv(f, N, 900 + k)
        pg, mg = eng.deconv(f, N, 900 + k)
        assert np.allclose(pg, po, rtol=0, atol=1e-12), k
        assert np.allclose(mg, mo, rtol=0, atol=1e-12), k
        cm = 1 if k in (2, 3) else (4 if k in (6, 7) else 0)
        assert abs(eng.mmd(pg, mg, cm) - O.mmd(po, mo, cm)) < 1e-12
    a, b = R.normal(0, 2, (200, 3)), R.normal(0.5, 2, (150, 3))
    assert abs(eng.mmd(a, b, 0b010, 0.01) - O.mmd(a, b, 0b010, 0.01)) < 1e-12
    eng.close()
