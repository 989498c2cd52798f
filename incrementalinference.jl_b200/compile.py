"""Lower host graph objects to the flat descriptor tables of include/iifb200.h.

The tables are the GPU analogue of CommonConvWrapper (src/entities/FactorOperationalMemory.jl:21-70):
an immutable factor descriptor bound to device belief slots instead of aliased Julia vectors.
"""
import ctypes as C

import numpy as np

from . import _abi as A
from . import graph as G


class Tables:
    """Accumulates slots / distributions / factors and freezes them into ctypes arrays."""

    def __init__(self):
        self.slots = []      # (dim, circ_mask, cap)
        self.dists = []      # DistDesc field tuples
        self.dparams = []    # flat doubles
        self.factors = []    # FactorDesc field dicts
        self._frozen = None

    # ---- slots
    def add_slot(self, vartype: G.InferenceVariable, cap: int) -> int:
        assert cap <= A.IIF_MAX_POINTS, f"N={cap} exceeds IIF_MAX_POINTS={A.IIF_MAX_POINTS}"
        self.slots.append((vartype.dim, vartype.circ_mask, int(cap)))
        return len(self.slots) - 1

    # ---- distributions
    def _simple_block(self, Z):
        if isinstance(Z, G.Normal):
            return A.D_NORMAL, 1, [float(Z.mu), float(Z.sigma)]
        if isinstance(Z, G.Uniform):
            return A.D_UNIFORM, 1, [float(Z.a), float(Z.b)]
        if isinstance(Z, G.MvNormal):
            return A.D_MVNORMAL, Z.dim, list(Z.mu) + list(np.tril(Z.L).reshape(-1))
        raise A.IIFB200Error(f"distribution {type(Z).__name__} has no device sampler")

    def add_dist(self, Z) -> int:
        poff = len(self.dparams)
        if isinstance(Z, G.Mixture):
            blocks = [self._simple_block(c) for c in Z.components]
            kinds = {b[0] for b in blocks}
            dims = {b[1] for b in blocks}
            if len(kinds) != 1 or len(dims) != 1:
                raise A.IIFB200Error("Mixture components must share one distribution kind and dimension")
            self.dparams += list(Z.diversity)
            for b in blocks:
                self.dparams += b[2]
            self.dists.append((A.D_MIXTURE, dims.pop(), len(blocks), kinds.pop(), -1, poff))
        elif isinstance(Z, G.SlotRef):
            self.dists.append((A.D_KDE, Z.dim, 0, 0, Z.slot, poff))
        elif isinstance(Z, G.SampledBelief):
            self.dparams += list(Z.samples.reshape(-1))
            self.dists.append((A.D_SAMPLES, Z.dim, Z.samples.shape[0], 0, -1, poff))
        else:
            kind, dim, prm = self._simple_block(Z)
            self.dparams += prm
            self.dists.append((kind, dim, 0, 0, -1, poff))
        return len(self.dists) - 1

    # ---- factors
    def add_factor(self, fnc, slots, mh=None, nullhypo=0.0, inflation=5.0, vartype_sf=None) -> int:
        Z = fnc.Z
        d = self.add_dist(Z)
        zdim = self.dists[d][1]
        partial_mask = 0
        if isinstance(fnc, (G.PartialPrior, G.ManifoldPriorPartial)):
            for c in fnc.partial:
                partial_mask |= 1 << (int(c) - 1)
        self.factors.append(dict(kind=fnc.kind, arity=len(slots), zdim=zdim, dist=d, slot=list(slots),
                                 nmh=0 if mh is None else len(mh), partial_mask=partial_mask,
                                 solver=int(bool(getattr(fnc, "numeric", False))),
                                 aux=[float(x) for x in getattr(fnc, "aux", [])],
                                 mh=[] if mh is None else list(mh), nullhypo=float(nullhypo),
                                 inflation=float(inflation)))
        return len(self.factors) - 1

    # ---- freeze
    def freeze(self):
        ns, nf, nd = len(self.slots), len(self.factors), len(self.dists)
        slots = (A.SlotDesc * max(ns, 1))()
        off = 0
        for i, (dim, cm, cap) in enumerate(self.slots):
            slots[i].dim, slots[i].circ_mask, slots[i].cap, slots[i].pts_off = dim, cm, cap, off
            off += dim * cap
        dists = (A.DistDesc * max(nd, 1))()
        for i, t in enumerate(self.dists):
            (dists[i].kind, dists[i].dim, dists[i].ncomp, dists[i].comp_kind, dists[i].slot,
             dists[i].poff) = t
        factors = (A.FactorDesc * max(nf, 1))()
        for i, f in enumerate(self.factors):
            fd = factors[i]
            fd.kind, fd.arity, fd.zdim, fd.dist = f["kind"], f["arity"], f["zdim"], f["dist"]
            for k, s in enumerate(f["slot"]):
                fd.slot[k] = s
            fd.nmh, fd.partial_mask, fd.solver = f["nmh"], f["partial_mask"], f.get("solver", 0)
            for k, p in enumerate(f["mh"]):
                fd.mh[k] = p
            fd.nullhypo, fd.inflation = f["nullhypo"], f["inflation"]
            for k, x in enumerate(f.get("aux", [])):
                fd.aux[k] = x
        dparams = np.asarray(self.dparams if self.dparams else [0.0], dtype=np.float64)
        self._frozen = dict(nslots=ns, slots=slots, nfactors=nf, factors=factors, ndists=nd,
                            dists=dists, nparams=len(self.dparams), dparams=dparams,
                            total_doubles=off)
        return self._frozen


def solver_params_c(sp: G.SolverParams, seed=None) -> A.SolverParamsC:
    c = A.SolverParamsC()
    c.spreadNH, c.nullSurplusAdd = sp.spreadNH, sp.nullSurplusAdd
    c.inflateCycles, c.gibbsNiter = sp.inflateCycles, 1   # GraphProductOperations.jl:56 Niter=1
    c.seed = int(sp.seed if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
    return c


class HostArena:
    """Host mirror of the device arena (same packing as iifb200_upload_all / download_all)."""

    def __init__(self, frozen):
        self.frozen = frozen
        ns = frozen["nslots"]
        self.pts = np.zeros(max(frozen["total_doubles"], 1), dtype=np.float64)
        self.bw = np.zeros(max(ns, 1) * A.IIF_MAX_DIM, dtype=np.float64)
        self.ipc = np.zeros(max(ns, 1) * A.IIF_MAX_DIM, dtype=np.float64)
        self.npts = np.zeros(max(ns, 1), dtype=np.int32)
        self.flags = np.zeros(max(ns, 1), dtype=np.int32)

    def set(self, slot, pts, bw=None, initialized=True, ipc=None):
        s = self.frozen["slots"][slot]
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, s.dim)
        n = pts.shape[0]
        assert n <= s.cap
        self.pts[s.pts_off:s.pts_off + n * s.dim] = pts.reshape(-1)
        self.npts[slot] = n
        self.flags[slot] = 1 if initialized else 0
        if bw is not None:
            self.bw[slot * A.IIF_MAX_DIM:slot * A.IIF_MAX_DIM + s.dim] = np.asarray(bw)[:s.dim]
        if ipc is not None:
            self.ipc[slot * A.IIF_MAX_DIM:slot * A.IIF_MAX_DIM + s.dim] = np.asarray(ipc)[:s.dim]

    def get(self, slot):
        s = self.frozen["slots"][slot]
        n = int(self.npts[slot])
        pts = self.pts[s.pts_off:s.pts_off + n * s.dim].reshape(n, s.dim).copy()
        bw = self.bw[slot * A.IIF_MAX_DIM:slot * A.IIF_MAX_DIM + s.dim].copy()
        ipc = self.ipc[slot * A.IIF_MAX_DIM:slot * A.IIF_MAX_DIM + s.dim].copy()
        return pts, bw, ipc

    def copy(self):
        h = HostArena(self.frozen)
        h.pts[:], h.bw[:], h.ipc[:] = self.pts, self.bw, self.ipc
        h.npts[:], h.flags[:] = self.npts, self.flags
        return h


def make_conv_ops(specs):
    """specs: list of dicts(factor, sfidx, N, call_id, nullSurplus, meas_off, mhidx_off, uinf_off)."""
    ops = (A.ConvOp * max(len(specs), 1))()
    for i, s in enumerate(specs):
        o = ops[i]
        o.factor, o.sfidx, o.N, o.call_id = s["factor"], s["sfidx"], s["N"], s["call_id"]
        o.nullSurplus = s.get("nullSurplus", 0.0)
        o.meas_off = s.get("meas_off", -1)
        o.mhidx_off = s.get("mhidx_off", -1)
        o.uinf_off = s.get("uinf_off", -1)
    return ops


def make_prop_ops(specs):
    """specs: list of dicts(target_slot, out_slot, factors=[(idx, sfidx)...], N, call_id, any_multihypo)."""
    ops = (A.PropOp * max(len(specs), 1))()
    for i, s in enumerate(specs):
        o = ops[i]
        o.target_slot, o.out_slot = s["target_slot"], s.get("out_slot", s["target_slot"])
        fl = s["factors"]
        assert 1 <= len(fl) <= A.IIF_MAX_FACTORS
        o.nfactors, o.N = len(fl), s["N"]
        for k, (fi, sf) in enumerate(fl):
            o.factor[k], o.sfidx[k] = fi, sf
        o.call_id, o.any_multihypo = s["call_id"], int(s.get("any_multihypo", 0))
    return ops


def make_deconv_ops(specs):
    """specs: list of dicts(factor, out_slot, N, call_id) -> iif_deconv_op array (IIF_S_DECONV)."""
    ops = (A.DeconvOp * max(len(specs), 1))()
    for i, s in enumerate(specs):
        ops[i].factor, ops[i].out_slot, ops[i].N, ops[i].call_id = s["factor"], s["out_slot"], s["N"], s["call_id"]
    return ops


def make_sched_ops(ops_list, lanes=None):
    """ops_list: list of (kind, a, b); lanes: optional lane id per op (0 = none / barrier)."""
    ops = (A.SchedOp * max(len(ops_list), 1))()
    for i, (k, a, b) in enumerate(ops_list):
        ops[i].kind, ops[i].a, ops[i].b = k, a, b
        ops[i].lane = 0 if lanes is None else int(lanes[i])
    return ops
