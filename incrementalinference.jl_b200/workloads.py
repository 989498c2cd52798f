"""Synthetic workloads = the five BASELINE.json configs, built with the mirrored
graph API.  Initial beliefs are generated on the host the way graph-init would leave them
(x_k ~ N(k·step, sigma·sqrt(k+1))), so that the timed region is the tree solve itself."""
import numpy as np

from . import graph as G


def _init(fg, lbl, pts, bw_rule=1.06):
    v = fg.variables[lbl]
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, v.vartype.dim)
    n = pts.shape[0]
    bw = np.maximum(bw_rule * pts.std(axis=0) * n ** (-0.2), 1e-3)  # Silverman: only seeds VND.bw
    v.val, v.bw, v.initialized = pts, bw, True
    v.infoPerCoord = np.ones(v.vartype.dim)


def _wrap(a):
    return (np.asarray(a) + np.pi) % (2 * np.pi) - np.pi


def scalar_chain(n=4, N=100, seed=42, prior_sigma=0.1, odo=1.0, odo_sigma=0.1):
    """C1 (n=4) / C2 (n=1000): generateGraph_LineStep-style ContinuousScalar odometry chain
    (CanonicalGraphExamples.jl:154-240; testBasicGraphs.jl:325-343)."""
    sp = G.SolverParams(N=N, seed=seed, graphinit=False)
    fg = G.initfg(sp)
    R = np.random.default_rng(seed)
    for k in range(n):
        G.addVariable(fg, f"x{k}", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal(0.0, prior_sigma)))
    for k in range(n - 1):
        G.addFactor(fg, [f"x{k}", f"x{k+1}"], G.LinearRelative(G.Normal(odo, odo_sigma)))
    for k in range(n):
        _init(fg, f"x{k}", R.normal(odo * k, odo_sigma * np.sqrt(k + 1.0), (N, 1)))
    return fg


def scalar_chain_sessions(sessions=4, n=1000, N=100, seed=42, prior_sigma=0.1, odo=1.0, odo_sigma=0.1):
    """`sessions` independent C2-style chains s{b}x0..s{b}x{n-1} in ONE factor graph (a forest): several robots /
    sessions solved by one solveTree pass.  The Bayes tree has one root per session, so every wave of the pass holds
    the independent work of all sessions."""
    sp = G.SolverParams(N=N, seed=seed, graphinit=False)
    fg = G.initfg(sp)
    R = np.random.default_rng(seed)
    for b in range(sessions):
        for k in range(n):
            G.addVariable(fg, f"s{b}x{k}", G.ContinuousScalar)
        G.addFactor(fg, [f"s{b}x0"], G.Prior(G.Normal(10.0 * b, prior_sigma)))
        for k in range(n - 1):
            G.addFactor(fg, [f"s{b}x{k}", f"s{b}x{k+1}"], G.LinearRelative(G.Normal(odo, odo_sigma)))
        for k in range(n):
            _init(fg, f"s{b}x{k}", R.normal(10.0 * b + odo * k, odo_sigma * np.sqrt(k + 1.0), (N, 1)))
    return fg


def sessions_nd_order(sessions, n):
    """nested-dissection order of every session's chain, one after the other"""
    return [f"s{b}{x}" for b in range(sessions) for x in chain_nd_order(n)]


def four_door(N=200, seed=42):
    """C3: test/fourdoortest.jl:9-54 (Mixture priors = four doors) with useMsgLikelihoods=false."""
    sp = G.SolverParams(N=N, seed=seed, graphinit=False)
    fg = G.initfg(sp)
    R = np.random.default_rng(seed)
    doors = lambda: G.Mixture(G.Prior, [G.Normal(-100, 3), G.Normal(0, 3), G.Normal(100, 3), G.Normal(300, 3)],  # noqa: E731
                              [0.25] * 4)
    for k in range(1, 5):
        G.addVariable(fg, f"x{k}", G.ContinuousScalar)
    G.addFactor(fg, ["x1"], doors())
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(G.Normal(50.0, 2.0)))
    G.addFactor(fg, ["x2", "x3"], G.LinearRelative(G.Normal(50.0, 4.0)))
    G.addFactor(fg, ["x3"], doors())
    G.addFactor(fg, ["x3", "x4"], G.LinearRelative(G.Normal(200.0, 4.0)))
    G.addFactor(fg, ["x4"], doors())
    cents = np.array([-100.0, 0.0, 100.0, 300.0])
    for k, off in zip(range(1, 5), (0.0, 50.0, 100.0, 300.0)):
        c = R.choice(cents, N) if k != 2 else R.choice(cents, N) + 50.0
        _init(fg, f"x{k}", (c + R.normal(0, 3, N)).reshape(N, 1))
    return fg


def multihypo_doors(N=200, seed=42):
    """C3 companion: true `multihypo=` data association (testMultiHypo3Door.jl:40-57)."""
    sp = G.SolverParams(N=N, seed=seed, graphinit=False)
    fg = G.initfg(sp)
    R = np.random.default_rng(seed)
    G.addVariable(fg, "x0", G.ContinuousScalar)
    for i, m in enumerate((0.0, 10.0, 20.0, 40.0)):
        G.addVariable(fg, f"l{i}", G.ContinuousScalar)
        G.addFactor(fg, [f"l{i}"], G.Prior(G.Normal(m, 0.01)))
        _init(fg, f"l{i}", R.normal(m, 0.01, (N, 1)))
    G.addFactor(fg, ["x0", "l0", "l1", "l2", "l3"], G.LinearRelative(G.Normal(0.0, 0.25)),
                multihypo=[1.0, 0.25, 0.25, 0.25, 0.25])
    _init(fg, "x0", R.choice([0.0, 10.0, 20.0, 40.0], N).reshape(N, 1) + R.normal(0, 0.25, (N, 1)))
    return fg


def circular_chain(n=500, N=150, seed=42):
    """C4: testCircular.jl:14-16 scaled to n poses."""
    sp = G.SolverParams(N=N, seed=seed, graphinit=False)
    fg = G.initfg(sp)
    R = np.random.default_rng(seed)
    for k in range(n):
        G.addVariable(fg, f"x{k}", G.Circular)
    G.addFactor(fg, ["x0"], G.PriorCircular(G.Normal(0.0, 0.1)))
    for k in range(n - 1):
        G.addFactor(fg, [f"x{k}", f"x{k+1}"], G.CircularCircular(G.Normal(1.0, 0.1)))
    for k in range(n):
        _init(fg, f"x{k}", _wrap(R.normal(1.0 * k, min(0.1 * np.sqrt(k + 1.0), 1.0), (N, 1))))
    return fg


def euclid2_grid(rows=50, cols=100, N=100, seed=42, closure_every=5):
    """C5: rows x cols boustrophedon Position{2} grid with vertical loop closures."""
    sp = G.SolverParams(N=N, seed=seed, graphinit=False)
    fg = G.initfg(sp)
    R = np.random.default_rng(seed)
    pos, idx = [], {}
    for r in range(rows):
        cs = range(cols) if r % 2 == 0 else range(cols - 1, -1, -1)
        for c in cs:
            idx[(r, c)] = len(pos)
            pos.append((float(c), float(r)))
    n = len(pos)
    for k in range(n):
        G.addVariable(fg, f"x{k}", G.Position(2))
    cov = np.eye(2) * 0.01
    G.addFactor(fg, ["x0"], G.Prior(G.MvNormal([0.0, 0.0], cov)))
    for k in range(n - 1):
        d = np.subtract(pos[k + 1], pos[k])
        G.addFactor(fg, [f"x{k}", f"x{k+1}"], G.LinearRelative(G.MvNormal(d, cov)))
    for r in range(rows - 1):
        for c in range(0, cols, closure_every):
            a, b = idx[(r, c)], idx[(r + 1, c)]
            if abs(a - b) > 1:
                G.addFactor(fg, [f"x{a}", f"x{b}"], G.LinearRelative(G.MvNormal([0.0, 1.0], cov)))
    for k in range(n):
        _init(fg, f"x{k}", np.asarray(pos[k]) + R.normal(0, min(0.1 * np.sqrt(k + 1.0), 0.5), (N, 2)))
    return fg


def chain_nd_order(n):
    """Nested-dissection elimination order of a chain x0..x_{n-1} (odd-even reduction): eliminates every
    other remaining variable per round, so the Bayes tree has depth O(log n) instead of n."""
    remaining = list(range(n))
    order = []
    while len(remaining) > 2:
        elim = remaining[1::2] if len(remaining) % 2 == 1 else remaining[1:-1:2]
        if not elim:
            break
        order += elim
        es = set(elim)
        remaining = [k for k in remaining if k not in es]
    order += remaining
    return [f"x{k}" for k in order]


def generateGraph_Kaess(N=100, seed=42, graphinit=False):
    """CanonicalGraphExamples.jl:15-37 (Kaess et al. example; all Normal())."""
    fg = G.initfg(G.SolverParams(N=N, seed=seed, graphinit=graphinit))
    G.addVariable(fg, "x1", G.ContinuousScalar)
    G.addFactor(fg, ["x1"], G.Prior(G.Normal()))
    G.addVariable(fg, "x2", G.ContinuousScalar)
    G.addFactor(fg, ["x1", "x2"], G.LinearRelative(G.Normal()))
    G.addVariable(fg, "x3", G.ContinuousScalar)
    G.addFactor(fg, ["x2", "x3"], G.LinearRelative(G.Normal()))
    G.addVariable(fg, "l1", G.ContinuousScalar)
    G.addFactor(fg, ["x1", "l1"], G.LinearRelative(G.Normal()))
    G.addFactor(fg, ["x2", "l1"], G.LinearRelative(G.Normal()))
    G.addVariable(fg, "l2", G.ContinuousScalar)
    G.addFactor(fg, ["x3", "l2"], G.LinearRelative(G.Normal()))
    return fg


def generateGraph_CaesarRing1D(N=100, seed=42, graphinit=False):
    """CanonicalGraphExamples.jl:123-150."""
    fg = G.initfg(G.SolverParams(N=N, seed=seed, graphinit=graphinit))
    for k in range(7):
        G.addVariable(fg, f"x{k}", G.ContinuousScalar)
    G.addFactor(fg, ["x0"], G.Prior(G.Normal()))
    for k in range(6):
        G.addFactor(fg, [f"x{k}", f"x{k+1}"], G.LinearRelative(G.Normal()))
    G.addVariable(fg, "l1", G.ContinuousScalar)
    G.addFactor(fg, ["x0", "l1"], G.LinearRelative(G.Normal()))
    G.addFactor(fg, ["x6", "l1"], G.LinearRelative(G.Normal()))
    return fg


def generateGraph_LineStep(lineLength, poseEvery=2, landmarkEvery=4, posePriorsAt=(0,), landmarkPriorsAt=(), sightDistance=4,
                           graphinit=False, sigma_pose_prior=0.1, sigma_lm_prior=0.1, sigma_pose_pose=0.1, sigma_pose_lm=0.1,
                           solverParams=None):
    """CanonicalGraphExamples.jl:154-238 (scalar, noise-free means): poses every `poseEvery` steps along a line, landmarks
    every `landmarkEvery`, odometry between consecutive poses, sightings within `sightDistance`."""
    fg = G.initfg(solverParams or G.SolverParams(graphinit=graphinit))
    xs, lms = [], []
    for i in range(lineLength + 1):
        if i % poseEvery == 0:
            xs.append(i)
            G.addVariable(fg, f"x{i}", G.ContinuousScalar)
            if i in posePriorsAt:
                G.addFactor(fg, [f"x{i}"], G.Prior(G.Normal(float(i), sigma_pose_prior)), graphinit=graphinit)
            if i > 0:
                G.addFactor(fg, [f"x{i - poseEvery}", f"x{i}"], G.LinearRelative(G.Normal(float(poseEvery), sigma_pose_pose)),
                            graphinit=graphinit)
        if landmarkEvery != 0 and i % landmarkEvery == 0:
            lms.append(i)
            G.addVariable(fg, f"lm{i}", G.ContinuousScalar)
            if i in landmarkPriorsAt:
                G.addFactor(fg, [f"lm{i}"], G.Prior(G.Normal(float(i), sigma_lm_prior)), graphinit=graphinit)
    for xi in xs:
        for lmi in lms:
            dist = lmi - xi
            if abs(dist) < sightDistance:
                G.addFactor(fg, [f"x{xi}", f"lm{lmi}"], G.LinearRelative(G.Normal(float(dist), sigma_pose_lm)), graphinit=graphinit)
    return fg
