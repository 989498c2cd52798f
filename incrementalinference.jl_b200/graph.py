"""Host-side mirror of the reference's factor-graph / plugin surface for the hot path.

Julia is the reference's host language and is not installed in this image, so the host
orchestration above the C-ABI is mirrored in Python with the reference's names and argument
meaning (`initfg`, `addVariable!` -> `addVariable`, `addFactor!` -> `addFactor`, ...).
Only what the clique belief-convolution path needs is mirrored; the production host remains
the Julia package calling `libiifb200.so` through `ccall`.

Reference: src/services/FactorGraph.jl (addVariable! :573-631, addFactor! :806-861,
parseusermultihypo :633-654), src/Variables/DefaultVariables.jl, src/Factors/*.jl,
src/entities/SolverParams.jl.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _abi as A


# ---------------------------------------------------------------- variable types
@dataclass(frozen=True)
class InferenceVariable:
    """@defVariable Name Manifold identity (src/Variables/DefaultVariables.jl:9-52)."""
    name: str
    dim: int
    circ_mask: int = 0


def Position(n: int) -> InferenceVariable:  # DefaultVariables.jl:9-24
    return InferenceVariable(f"Position{{{n}}}", n, 0)


ContinuousEuclid = Position
ContinuousScalar = Position(1)              # DefaultVariables.jl:32
Circular = InferenceVariable("Circular", 1, 1)  # DefaultVariables.jl:52  RealCircleGroup
# @defVariable SpecialEuclidean2 SpecialEuclidean(2; vectors=HybridTangentRepresentation()) (test/testSpecialEuclidean2Mani.jl:14):
# device points are the coordinates (x, y, theta) = vee(log(M, eps, p)); theta is circular
SpecialEuclidean2 = InferenceVariable("SpecialEuclidean2", 3, 0b100)


# @defVariable SpecialOrthogonal2 SpecialOrthogonal(2) (test/testSpecialOrthogonalMani.jl:19): a rotation matrix is its
# angle atan2(R21, R11) on the device — the same single circular coordinate as Circular / RealCircleGroup
SpecialOrthogonal2 = InferenceVariable("SpecialOrthogonal2", 1, 1)


# @defVariable SO3 SpecialOrthogonal(3) (test/testSpecialOrthogonalMani.jl:84): device points are rotation vectors
# omega = vee(log(M, eps, R)); AMP's KDE treats them as (:Euclid, :Euclid, :Euclid) (same test, :80-81), the factors,
# entropy and statistics compose on the group (slot flag IIF_MANI_SO3)
SpecialOrthogonal3 = InferenceVariable("SpecialOrthogonal3", 3, A.MANI_SO3)


def so3_point_to_coords(R) -> np.ndarray:
    """rotation matrix -> rotation vector (log at the identity, DefaultOrthogonalBasis: [X32, X13, X21])"""
    R = np.asarray(R, dtype=np.float64)
    c = np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-9:
        return 0.5 * v
    if np.pi - th < 1e-6:                                        # near pi: axis from the symmetric part
        A_ = (R + np.eye(3)) / 2.0
        ax = np.sqrt(np.maximum(np.diag(A_), 0.0))
        k = int(np.argmax(ax))
        ax = A_[:, k] / ax[k]
        return th * ax / np.linalg.norm(ax)
    return th / (2.0 * np.sin(th)) * v


def so3_coords_to_point(w) -> np.ndarray:
    """rotation vector -> rotation matrix (Rodrigues)"""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def so3_mean(pts) -> np.ndarray:
    """mean(SpecialOrthogonal(3), pts) as a rotation matrix (iterated Karcher mean on rotation vectors)"""
    Rm = so3_coords_to_point(pts[0])
    for _ in range(20):
        d = np.mean([so3_point_to_coords(Rm.T @ so3_coords_to_point(p)) for p in pts], axis=0)
        Rm = Rm @ so3_coords_to_point(d)
        if np.linalg.norm(d) < 1e-12:
            break
    return Rm


def so2_point_to_coords(R) -> np.ndarray:
    R = np.asarray(R, dtype=np.float64)
    return np.array([np.arctan2(R[1, 0], R[0, 0])])


def TranslationGroup(n: int) -> InferenceVariable:
    return Position(n)


RealCircleGroup = Circular


def se2_point_to_coords(t, R) -> np.ndarray:
    """ArrayPartition(t, R) -> (x, y, theta): AMP.makeCoordsFromPoint for SpecialEuclidean(2)."""
    R = np.asarray(R, dtype=np.float64)
    return np.array([t[0], t[1], np.arctan2(R[1, 0], R[0, 0])])


def se2_coords_to_point(c):
    """(x, y, theta) -> (t, R): AMP.makePointFromCoords for SpecialEuclidean(2)."""
    cs, sn = np.cos(c[2]), np.sin(c[2])
    return np.array([c[0], c[1]]), np.array([[cs, -sn], [sn, cs]])


# ---------------------------------------------------------------- SamplableBelief
@dataclass
class Normal:
    mu: float = 0.0
    sigma: float = 1.0
    dim: int = 1


@dataclass
class Uniform:
    a: float = 0.0
    b: float = 1.0
    dim: int = 1


class SampledBelief:
    """Any SamplableBelief (`rand(Z)` is all sampleFactor! needs, SolverUtilities.jl:50-76): a table of host-drawn
    samples that the device resamples with replacement — Rayleigh, Gamma, user-defined distributions ..."""

    def __init__(self, samples):
        s = np.ascontiguousarray(np.asarray(samples, dtype=np.float64))
        self.samples = s.reshape(len(s), -1)
        self.dim = self.samples.shape[1]

    @classmethod
    def from_sampler(cls, draw, count: int = 4096):
        """`draw(count)` returns count samples (count x dim or count,), e.g. a scipy.stats frozen distribution's rvs"""
        return cls(draw(count))


class MvNormal:
    def __init__(self, mu, cov):
        self.mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        cov = np.asarray(cov, dtype=np.float64)
        if cov.ndim == 1:
            cov = np.diag(cov)
        self.cov = cov
        self.L = np.linalg.cholesky(cov)
        self.dim = self.mu.shape[0]


@dataclass
class ManifoldKernelDensity:
    """AMP.ManifoldKernelDensity as IIF consumes it: points, per-dim bandwidth, partial, ipc."""
    vartype: InferenceVariable
    pts: np.ndarray                      # N x d
    bw: np.ndarray                       # d
    partial: Optional[List[int]] = None  # 1-based coordinate list
    infoPerCoord: Optional[np.ndarray] = None

    @property
    def dim(self):
        return self.vartype.dim


def getPoints(mkd: ManifoldKernelDensity, _=False):
    return mkd.pts


def getBW(mkd: ManifoldKernelDensity):
    return mkd.bw


def Npts(mkd: ManifoldKernelDensity):
    return mkd.pts.shape[0]


# ---------------------------------------------------------------- factors (src/Factors)
@dataclass
class Prior:               # DefaultPrior.jl:8
    Z: object
    kind = A.F_PRIOR
    is_prior = True


@dataclass
class PriorCircular:       # Circular.jl:54
    Z: object
    kind = A.F_PRIOR_CIRCULAR
    is_prior = True


@dataclass
class PartialPrior:        # PartialPrior.jl:11-21
    Z: object
    partial: Sequence[int]  # 1-based coordinates
    kind = A.F_PARTIAL_PRIOR
    is_prior = True


@dataclass
class MsgPrior:            # MsgPrior.jl:10
    Z: object              # ManifoldKernelDensity or SlotRef (device-resident belief)
    infoPerCoord: Optional[np.ndarray] = None
    kind = A.F_MSG_PRIOR
    is_prior = True


@dataclass
class LinearRelative:      # LinearRelative.jl:13
    Z: object
    kind = A.F_LINEAR_RELATIVE
    is_prior = False


@dataclass
class CircularCircular:    # Circular.jl:13
    Z: object
    kind = A.F_CIRCULAR_CIRCULAR
    is_prior = False


@dataclass
class EuclidDistance:      # EuclidDistance.jl:9
    Z: object
    kind = A.F_EUCLID_DISTANCE
    is_prior = False


class ManifoldPrior:
    """ManifoldPrior(M, p, Z) — GenericFunctions.jl:163-214: Z is a tangent-coordinate distribution at the point
    `p`; a sample is retract(M, p, hat(M, p, rand(Z))), which for TranslationGroup / RealCircleGroup /
    SpecialEuclidean(2) (hybrid representation) is p (+) Z coordinate-wise, so `p` (coordinates) is folded
    into Z's mean when the factor is lowered."""
    kind = A.F_MANIFOLD_PRIOR
    is_prior = True

    def __init__(self, M: InferenceVariable, p, Z):
        self.M, self.p, self.Z0 = M, np.atleast_1d(np.asarray(p, dtype=np.float64)), Z
        if M.circ_mask & A.MANI_SO3:
            # SpecialOrthogonal(3): retract(M, p, hat(Z)) = p Exp(z) is not coordinate-additive: p travels with the factor
            if self.p.shape == (3, 3):
                self.p = so3_point_to_coords(self.p)
            self.kind, self.aux, self.Z = A.F_SO3_PRIOR, self.p, Z
            return
        assert self.p.shape[0] == M.dim, "p must be given in coordinates (see se2_point_to_coords)"
        if isinstance(Z, MvNormal):
            self.Z = MvNormal(Z.mu + self.p, Z.cov)
        elif isinstance(Z, Normal):
            self.Z = Normal(Z.mu + float(self.p[0]), Z.sigma)
        else:
            raise A.IIFB200Error(f"ManifoldPrior: distribution {type(Z).__name__} has no device sampler")


class ManifoldPriorPartial:
    """ManifoldPriorPartial(M, Z, partial) — GenericFunctions.jl:288-303: Xc[partial] = rand(Z) at the identity."""
    kind = A.F_MANIFOLD_PRIOR
    is_prior = True

    def __init__(self, M: InferenceVariable, Z, partial: Sequence[int]):
        self.M, self.Z, self.partial = M, Z, tuple(int(c) for c in partial)


class ManifoldFactor:
    """ManifoldFactor(M, Z) — GenericFunctions.jl:64-100: relative factor with the measurement a tangent at the
    identity; residual distanceTangent2Point (:39-52).  Lowered to the residual of the group M."""
    is_prior = False

    def __init__(self, M: InferenceVariable, Z):
        self.M, self.Z = M, Z
        if M.name == "SpecialEuclidean2":
            self.kind = A.F_SE2_RELATIVE
        elif M.circ_mask & A.MANI_SO3:
            self.kind = A.F_SO3_RELATIVE              # q = p Exp(X) on SpecialOrthogonal(3)
        elif M.circ_mask == 0:
            self.kind = A.F_LINEAR_RELATIVE           # TranslationGroup: exp(p, X) = p + X
        elif M.dim == 1 and M.circ_mask == 1:
            self.kind = A.F_CIRCULAR_CIRCULAR         # RealCircleGroup
        else:
            raise A.IIFB200Error(f"ManifoldFactor on {M.name} has no device residual (no CPU fallback)")


class Mixture:
    """Mixture(mechanics, components, diversity) — src/Factors/Mixture.jl:37-45."""

    def __init__(self, mechanics, components, diversity):
        self.mechanics = mechanics if isinstance(mechanics, type) else type(mechanics)
        self.components = list(components)
        w = np.asarray(diversity, dtype=np.float64)
        self.diversity = w / w.sum()
        self.kind = self.mechanics.kind
        self.is_prior = self.mechanics.is_prior
        self.Z = self


@dataclass
class SlotRef:
    """A density that already lives in a device belief slot (separator message)."""
    slot: int
    dim: int


# ---------------------------------------------------------------- SolverParams
@dataclass
class SolverParams:
    """src/entities/SolverParams.jl:12-75 (hot-path subset, same defaults)."""
    N: int = 100
    spreadNH: float = 3.0
    inflation: float = 5.0
    nullSurplusAdd: float = 0.3
    inflateCycles: int = 3
    gibbsIters: int = 3
    graphinit: bool = True
    useMsgLikelihoods: bool = False
    alwaysFreshMeasurements: bool = True
    downsolve: bool = True
    seed: int = 42            # Random.seed! analogue: Philox key of every device stream
    devParams: Dict[str, str] = field(default_factory=lambda: {"backend": "b200"})


# ---------------------------------------------------------------- graph
@dataclass
class DFGVariable:
    label: str
    vartype: InferenceVariable
    val: np.ndarray                 # npts x d (VariableNodeData.val)
    bw: np.ndarray                  # d       (VariableNodeData.bw[:,1])
    infoPerCoord: np.ndarray
    initialized: bool = False
    index: int = -1
    ppeDict: dict = field(default_factory=dict)   # solveKey -> MeanMaxPPE (getPPEDict)


@dataclass
class MeanMaxPPE:
    """MeanMaxPPE (DFG): suggested / max / mean coordinates of a belief — calcPPE, FGOSUtils.jl:237-278"""
    solveKey: str
    suggested: np.ndarray
    max: np.ndarray
    mean: np.ndarray


@dataclass
class DFGFactor:
    label: str
    variables: List[str]            # getVariableOrder
    fnc: object                     # getFactorType
    multihypo: Optional[np.ndarray]  # parseusermultihypo output (Categorical.p) or None
    nullhypo: float
    inflation: float
    index: int = -1

    @property
    def is_prior(self):
        return self.fnc.is_prior


def isMultihypo(f: DFGFactor) -> bool:
    return f.multihypo is not None


def parseusermultihypo(multihypo, nullhypo):
    """FactorGraph.jl:633-654."""
    if multihypo is None or len(multihypo) == 0:
        return None, float(nullhypo)
    mh = np.array(multihypo, dtype=np.float64)
    mh[mh > 1 - 1e-10] = 0.0
    s = mh.sum()
    assert abs(s % 1) < 1e-10 or 1 - 1e-10 < s % 1, "ensure multihypo sums to an integer, see #1086"
    assert abs(mh[mh > 1e-10].sum() - 1) < 1e-8
    mh /= mh.sum()
    return mh, float(nullhypo)


class FactorGraph:
    def __init__(self, solverParams: Optional[SolverParams] = None):
        self.variables: Dict[str, DFGVariable] = {}
        self.factors: Dict[str, DFGFactor] = {}
        self.solverParams = solverParams or SolverParams()
        self._engine = None          # lazily-built device engine (solver.Engine)
        self._version = 0            # bumped on structural change
        self._by_var: Dict[str, List[str]] = {}   # variable -> factor labels (insertion order)
        # Philox call ids handed out so far.  Every numeric call on this graph (approxConv, propagateBelief, initAll,
        # solveTree) draws its ids from here, so consecutive calls and consecutive solves see fresh noise, as the
        # reference's global RNG gives them (a per-engine counter restarted at 0 replayed the same streams).
        self._call_counter = 0

    def next_call(self, n: int = 16) -> int:
        """reserve `n` consecutive Philox call ids; returns the first"""
        c = self._call_counter
        self._call_counter = (c + int(n)) % (1 << 31)
        return c

    # DFG accessors used by the mirrored API
    def getVariable(self, lbl):
        return self.variables[lbl]

    def getFactor(self, lbl):
        return self.factors[lbl]

    def ls(self, lbl=None):
        return list(self.variables) if lbl is None else self.listNeighbors(lbl)

    def lsf(self, lbl=None):
        return list(self.factors) if lbl is None else self.listNeighbors(lbl)

    def listNeighbors(self, lbl):
        if lbl in self.variables:
            return list(self._by_var.get(lbl, ()))
        return list(self.factors[lbl].variables)


def initfg(solverParams: Optional[SolverParams] = None) -> FactorGraph:
    return FactorGraph(solverParams)


def getSolverParams(fg: FactorGraph) -> SolverParams:
    return fg.solverParams


def addVariable(fg: FactorGraph, label: str, vartype: InferenceVariable) -> DFGVariable:
    """addVariable! — FactorGraph.jl:573-631 (uninitialised: no points yet)."""
    assert label not in fg.variables
    d = vartype.dim
    assert d <= A.IIF_MAX_DIM, f"variable dimension {d} > IIF_MAX_DIM"
    v = DFGVariable(label, vartype, np.zeros((0, d)), np.zeros(d), np.zeros(d), False,
                    len(fg.variables))
    fg.variables[label] = v
    fg._version += 1
    return v


def addFactor(fg: FactorGraph, variables: Sequence[str], fnc, multihypo=None, nullhypo: float = 0.0,
              graphinit: Optional[bool] = None, inflation: Optional[float] = None,
              label: Optional[str] = None) -> DFGFactor:
    """addFactor! — FactorGraph.jl:806-861."""
    variables = list(variables)
    for v in variables:
        assert v in fg.variables, f"unknown variable {v}"
    assert len(variables) <= A.IIF_MAX_ARITY
    if not hasattr(fnc, "kind"):
        raise A.IIFB200Error(
            f"factor type {type(fnc).__name__} has no device residual: only the built-in residual "
            "library is supported on the b200 backend (no CPU fallback)")
    if label is None:
        base = "".join(variables) + "f"
        k = 1
        while f"{base}{k}" in fg.factors:
            k += 1
        label = f"{base}{k}"
    mh, nh = parseusermultihypo(multihypo, nullhypo)
    if mh is not None:
        assert len(mh) == len(variables)
    f = DFGFactor(label, variables, fnc, mh, nh,
                  fg.solverParams.inflation if inflation is None else float(inflation),
                  len(fg.factors))
    fg.factors[label] = f
    for v in dict.fromkeys(variables):
        fg._by_var.setdefault(v, []).append(label)
    fg._version += 1
    do_init = fg.solverParams.graphinit if graphinit is None else graphinit
    if do_init:
        from .solver import doautoinit
        for v in variables:
            doautoinit(fg, v)
    return f


def factorCanInitFromOtherVars(fg: "FactorGraph", fct: str, loovar: str, isinit=None) -> bool:
    """factorCanInitFromOtherVars — GraphInit.jl:62-116: priors always; n-ary factors when `loovar` is the only
    uninitialised variable; multihypo carve-out (#427, isLeastOneHypoAvailable FactorGraph.jl:772-784): solving a
    certain variable needs at least one initialised hypothesis, solving an uncertain one needs all certain ones.
    `isinit` (label -> bool) overrides the variables' own flags (used when a sweep is replayed on flags only)."""
    f = fg.factors[fct]
    init = [(fg.variables[v].initialized if isinit is None else isinit[v]) for v in f.variables]
    fail = [v for v, i in zip(f.variables, init) if not i]
    canuse = len(f.variables) == 1 or (len(fail) == 1 and loovar in fail)
    if not canuse and isMultihypo(f):
        sfidx = f.variables.index(loovar)
        certain = [i for i, p in enumerate(f.multihypo) if p == 0.0]     # parseusermultihypo: certain -> 0.0
        uncertn = [i for i, p in enumerate(f.multihypo) if p > 0.0]
        canuse = (sfidx in certain and any(init[i] for i in uncertn)) or \
                 (sfidx in uncertn and all(init[i] for i in certain))
    return canuse


def isInitialized(fg_or_var, lbl=None) -> bool:
    v = fg_or_var if lbl is None else fg_or_var.variables[lbl]
    return v.initialized


def getVal(v: DFGVariable):
    return v.val


def getBelief(fg: FactorGraph, lbl: str) -> ManifoldKernelDensity:
    """FactorGraph.jl:335 — rebuild an MKD from VND (val, bw)."""
    v = fg.variables[lbl]
    return ManifoldKernelDensity(v.vartype, v.val.copy(), v.bw.copy(), None, v.infoPerCoord.copy())


def setValKDE(fg: FactorGraph, lbl: str, mkd: ManifoldKernelDensity, setinit=True, ipc=None):
    """setValKDE! — FactorGraph.jl:237-286."""
    v = fg.variables[lbl]
    v.val = np.ascontiguousarray(mkd.pts, dtype=np.float64).reshape(-1, v.vartype.dim).copy()
    v.bw = np.asarray(mkd.bw, dtype=np.float64).copy()
    if ipc is not None:
        v.infoPerCoord = np.asarray(ipc, dtype=np.float64).copy()
    if setinit:
        v.initialized = True
    if fg._engine is not None:
        fg._engine.mark_dirty(lbl)


def initVariable(fg: FactorGraph, lbl: str, pts, bw=None):
    """initVariable!(fg, lbl, points) — GraphInit.jl:288-330: manual initialisation."""
    from .solver import kde_bandwidth
    v = fg.variables[lbl]
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(-1, v.vartype.dim))
    if bw is None:
        bw = kde_bandwidth(fg, v.vartype, pts)
    setValKDE(fg, lbl, ManifoldKernelDensity(v.vartype, pts, np.asarray(bw, dtype=np.float64)), True,
              np.zeros(v.vartype.dim))
