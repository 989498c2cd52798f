"""ctypes wrapper of the CPU ORACLE (oracle/iif_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package never imports this module.  KDE bandwidth (a14) and KDE product
(a15) are PARITY UNPINNED restatements (see oracle/iif_oracle.h).
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "_build", "libiif_oracle.so")


def _load_pkg():
    if "iifb200" in sys.modules:
        return sys.modules["iifb200"]
    sys.path.insert(0, _ROOT)
    import iifb200  # noqa: F401  (root-level shim that loads incrementalinference.jl_b200/)
    return sys.modules["iifb200"]


_pkg = _load_pkg()
A = _pkg._abi


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in ("iif_oracle.c", "iif_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


class OracleGraph(C.Structure):
    _fields_ = [("nslots", C.c_int32), ("slots", C.POINTER(A.SlotDesc)), ("pts", A._dp), ("bw", A._dp),
                ("ipc", A._dp), ("npts", A._ip), ("flags", A._ip), ("nfactors", C.c_int32),
                ("factors", C.POINTER(A.FactorDesc)), ("ndists", C.c_int32),
                ("dists", C.POINTER(A.DistDesc)), ("dparams", A._dp), ("sp", A.SolverParamsC)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, ip = A._dp, A._ip
        L.iifo_uniform.restype = C.c_double
        L.iifo_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.iifo_normal.restype = C.c_double
        L.iifo_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.iifo_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.iifo_hypo_recipe.restype = C.c_int32
        L.iifo_hypo_recipe.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, ip, C.c_double, dp, ip, ip, ip,
                                       ip, ip, ip, ip, ip]
        L.iifo_residual.restype = C.c_int32
        L.iifo_residual.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, C.c_int32, dp, dp]
        L.iifo_std_basic_spread.restype = C.c_double
        L.iifo_std_basic_spread.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32]
        L.iifo_kde_bandwidth.restype = C.c_int32
        L.iifo_kde_bandwidth.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, dp]
        L.iifo_loo_nll.restype = C.c_double
        L.iifo_loo_nll.argtypes = [dp, C.c_int32, C.c_int32, C.c_double]
        L.iifo_conv.restype = C.c_int32
        L.iifo_conv.argtypes = [C.POINTER(OracleGraph), C.POINTER(A.ConvOp), dp, ip, dp, dp, dp, dp, ip, ip]
        L.iifo_product.restype = C.c_int32
        L.iifo_product.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, dp, dp, ip, dp, C.c_uint64,
                                   C.c_uint32, C.c_int32, dp, dp, dp, dp, ip]
        L.iifo_propagate.restype = C.c_int32
        L.iifo_propagate.argtypes = [C.POINTER(OracleGraph), C.POINTER(A.PropOp)]
        L.iifo_schedule_run.restype = C.c_int32
        L.iifo_schedule_run.argtypes = [C.POINTER(OracleGraph), C.c_int32, ip, C.POINTER(A.SchedOp),
                                        C.POINTER(A.PropOp), C.c_int32, C.c_int32]
        L.iifo_schedule_run_ex.restype = C.c_int32
        L.iifo_schedule_run_ex.argtypes = [C.POINTER(OracleGraph), C.c_int32, ip, C.POINTER(A.SchedOp),
                                           C.POINTER(A.PropOp), C.POINTER(A.DeconvOp), C.c_int32, C.c_int32]
        L.iifo_deconv_to_slot.restype = C.c_int32
        L.iifo_deconv_to_slot.argtypes = [C.POINTER(OracleGraph), C.POINTER(A.DeconvOp)]
        L.iifo_conv_count.restype = C.c_int64
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(A._dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(A._ip) if a is not None else None


def _check(st, what):
    if st != 0:
        raise RuntimeError(f"oracle {what} failed with status {st}")


class Oracle:
    """Oracle bound to frozen tables + a HostArena (both from iifb200.compile)."""

    def __init__(self, frozen, arena, sp_c):
        self.frozen, self.arena, self.sp_c = frozen, arena, sp_c
        g = OracleGraph()
        g.nslots, g.slots = frozen["nslots"], frozen["slots"]
        g.pts, g.bw, g.ipc = _dp(arena.pts), _dp(arena.bw), _dp(arena.ipc)
        g.npts, g.flags = _ip(arena.npts), _ip(arena.flags)
        g.nfactors, g.factors = frozen["nfactors"], frozen["factors"]
        g.ndists, g.dists = frozen["ndists"], frozen["dists"]
        g.dparams = _dp(frozen["dparams"])
        g.sp = sp_c
        self.g = g

    def conv(self, op, meas=None, mhidx=None, uinf=None):
        """one iif_conv_op -> (pts N x d, bw d, ipc d, mhidx N, nan_count)"""
        f = self.frozen["factors"][op.factor]
        d = self.frozen["slots"][f.slot[op.sfidx - 1]].dim
        pts = np.zeros((op.N, d))
        bw, ipc = np.zeros(A.IIF_MAX_DIM), np.zeros(A.IIF_MAX_DIM)
        lab = np.zeros(op.N, dtype=np.int32)
        nan = C.c_int32(0)
        st = lib().iifo_conv(C.byref(self.g), C.byref(op), _dp(meas), _ip(mhidx), _dp(uinf), _dp(pts),
                             _dp(bw), _dp(ipc), _ip(lab), C.cast(C.byref(nan), A._ip))
        _check(st, "conv")
        return pts, bw[:d], ipc[:d], lab, nan.value

    def deconv(self, factor, N, call_id):
        """approxDeconv of one factor -> (predicted N x z, sampled N x z)"""
        zd = self.frozen["factors"][factor].zdim
        pred, meas = np.zeros((N, zd)), np.zeros((N, zd))
        L = lib()
        L.iifo_deconv.restype = C.c_int32
        _check(L.iifo_deconv(C.byref(self.g), C.c_int32(factor), C.c_int32(N), C.c_int32(call_id), _dp(pred), _dp(meas)),
               "deconv")
        return pred, meas

    def propagate(self, op):
        _check(lib().iifo_propagate(C.byref(self.g), C.byref(op)), "propagate")

    def schedule_run(self, wave_off, ops, props, first=0, last=None, deconvs=None):
        nw = len(wave_off) - 1
        wo = np.asarray(wave_off, dtype=np.int32)
        _check(lib().iifo_schedule_run_ex(C.byref(self.g), nw, _ip(wo), ops, props, deconvs, first,
                                          nw if last is None else last), "schedule_run")

    def deconv_to_slot(self, op):
        """IIF_S_DECONV: differential likelihood of an up message into a belief slot"""
        _check(lib().iifo_deconv_to_slot(C.byref(self.g), C.byref(op)), "deconv_to_slot")


def uniform(seed, call, stream, idx):
    return lib().iifo_uniform(seed, call, stream, idx)


def normal(seed, call, stream, idx):
    return lib().iifo_normal(seed, call, stream, idx)


def kde_bandwidth(pts, circ_mask=0):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    n, d = pts.shape
    bw = np.zeros(A.IIF_MAX_DIM)
    _check(lib().iifo_kde_bandwidth(_dp(pts), n, d, circ_mask, _dp(bw)), "kde_bandwidth")
    return bw[:d]


def ppe(pts, bw, circ_mask=0):
    """calcPPE: (mean, max) coordinates of one belief"""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    n, d = pts.shape
    b = np.zeros(A.IIF_MAX_DIM)
    b[:d] = np.asarray(bw, dtype=np.float64)[:d]
    mean, mx = np.zeros(A.IIF_MAX_DIM), np.zeros(A.IIF_MAX_DIM)
    L = lib()
    L.iifo_ppe.restype = C.c_int32
    L.iifo_ppe.argtypes = [_dp(pts).__class__, C.c_int32, C.c_int32, C.c_int32, _dp(b).__class__, _dp(mean).__class__,
                           _dp(mx).__class__]
    _check(L.iifo_ppe(_dp(pts), n, d, circ_mask, _dp(b), _dp(mean), _dp(mx)), "ppe")
    return mean[:d], mx[:d]


def mmd(a, b, circ_mask=0, bw=0.001):
    """AMP.mmd kernel-embedding distance between two point sets (N x d each)"""
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(len(a), -1)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(len(b), -1)
    L = lib()
    L.iifo_mmd.restype = C.c_double
    return float(L.iifo_mmd(_dp(a), C.c_int32(a.shape[0]), _dp(b), C.c_int32(b.shape[0]), C.c_int32(a.shape[1]),
                            C.c_int32(circ_mask), C.c_double(bw)))


def loo_nll(x, h, circular=0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().iifo_loo_nll(_dp(x), len(x), circular, h)


def std_basic_spread(pts, circ_mask=0):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    if pts.ndim == 1:
        pts = pts.reshape(-1, 1)
    return lib().iifo_std_basic_spread(_dp(pts), pts.shape[0], pts.shape[1], circ_mask)


def residual(kind, z, xs, d=None, circ_mask=0):
    z = np.ascontiguousarray(z, dtype=np.float64)
    xs = [np.ascontiguousarray(x, dtype=np.float64) for x in xs]
    d = d or len(xs[0])
    x = np.concatenate(xs)
    res = np.zeros(A.IIF_MAX_DIM)
    _check(lib().iifo_residual(kind, d, circ_mask, len(z), _dp(z), len(xs), _dp(x), _dp(res)), "residual")
    return res[:len(z)]


def hypo_recipe(mh, maxlen, sfidx, lenXi, isinit=None, nullhypo=0.0, u=None, mhidx_in=None):
    """_prepareHypoRecipe!(mh, maxlen, sfidx, lenXi, isinit, nullhypo) -> dict like HypoRecipe."""
    mh_a = None if mh is None else np.ascontiguousarray(mh, dtype=np.float64)
    isin = None if isinit is None else np.ascontiguousarray(isinit, dtype=np.int32)
    if u is None:
        u = np.random.default_rng(0).random(maxlen)
    u = np.ascontiguousarray(u, dtype=np.float64)
    mi = None if mhidx_in is None else np.ascontiguousarray(mhidx_in, dtype=np.int32)
    mhidx = np.zeros(maxlen, dtype=np.int32)
    nb = C.c_int32(0)
    ncer = C.c_int32(0)
    MA = A.IIF_MAX_ARITY
    bh = np.zeros(MA + 2, dtype=np.int32)
    bnv = np.zeros(MA + 2, dtype=np.int32)
    bv = np.zeros((MA + 2) * MA, dtype=np.int32)
    cer = np.zeros(MA, dtype=np.int32)
    st = lib().iifo_hypo_recipe(_dp(mh_a), lenXi, maxlen, sfidx, _ip(isin), nullhypo, _dp(u), _ip(mi),
                                _ip(mhidx), C.cast(C.byref(nb), A._ip), _ip(bh), _ip(bnv), _ip(bv), _ip(cer),
                                C.cast(C.byref(ncer), A._ip))
    _check(st, "hypo_recipe")
    activehypo = [(int(bh[b]), [int(v) for v in bv[b * MA:b * MA + bnv[b]]]) for b in range(nb.value)]
    allelements = [[int(n) + 1 for n in np.nonzero(mhidx == bh[b])[0]] if (bnv[b] > 0 or mh is None) else []
                   for b in range(nb.value)]
    # the `Nothing` method puts every non-null element in bucket 1 and leaves the rest empty
    return dict(certainidx=[int(c) for c in cer[:ncer.value]], allelements=allelements,
                activehypo=activehypo, mhidx=mhidx.copy())


def product(dens_pts, dens_bw, vartype_dim, circ_mask=0, dens_mask=None, old_pts=None, seed=42, call_id=0,
            niter=1, randU=None, randN=None):
    """dens_pts F x N x d, dens_bw F x d -> (pts N x d, bw d, labels N x F)"""
    dens_pts = np.ascontiguousarray(dens_pts, dtype=np.float64)
    F, N = dens_pts.shape[0], dens_pts.shape[1]
    d = vartype_dim
    bwp = np.zeros((F, A.IIF_MAX_DIM))
    bwp[:, :d] = np.asarray(dens_bw, dtype=np.float64).reshape(F, d)
    mask = None if dens_mask is None else np.ascontiguousarray(dens_mask, dtype=np.int32)
    old = None if old_pts is None else np.ascontiguousarray(old_pts, dtype=np.float64)
    out = np.zeros((N, d))
    obw = np.zeros(A.IIF_MAX_DIM)
    lab = np.zeros((N, F), dtype=np.int32)
    ru = None if randU is None else np.ascontiguousarray(randU, dtype=np.float64)
    rn = None if randN is None else np.ascontiguousarray(randN, dtype=np.float64)
    st = lib().iifo_product(d, circ_mask, F, N, _dp(dens_pts), _dp(bwp), _ip(mask), _dp(old), seed, call_id,
                            niter, _dp(ru), _dp(rn), _dp(out), _dp(obw), _ip(lab))
    _check(st, "product")
    return out, obw[:d], lab
