"""ctypes mirror of include/iifb200.h and loader for libiifb200.so.

The product path FAILS LOUDLY when the CUDA library is missing: there is no CPU fallback
(BASELINE.json north_star).  Struct layouts here are shared (as data layout only) with the
test oracle wrapper under oracle/.
"""
import ctypes as C
import os

IIF_MAX_DIM = 4
IIF_MAX_ARITY = 6
IIF_MAX_FACTORS = 16
IIF_MAX_POINTS = 256

IIF_OK, IIF_ERR_ARG, IIF_ERR_CUDA, IIF_ERR_UNSUPPORTED, IIF_ERR_STATE = 0, -1, -2, -3, -4

# iif_factor_kind
F_PRIOR, F_LINEAR_RELATIVE, F_PRIOR_CIRCULAR, F_CIRCULAR_CIRCULAR = 1, 2, 3, 4
F_EUCLID_DISTANCE, F_MSG_PRIOR, F_PARTIAL_PRIOR = 5, 6, 7
F_MANIFOLD_PRIOR, F_SE2_RELATIVE = 8, 9
F_SO3_PRIOR, F_SO3_RELATIVE = 10, 11
MANI_SO3 = 0x100
# iif_dist_kind
D_NORMAL, D_MVNORMAL, D_MIXTURE, D_KDE, D_UNIFORM, D_SAMPLES = 1, 2, 3, 4, 5, 6
# iif_sched_kind
S_PROPAGATE, S_COPY, S_DECONV, S_PUSH, S_WAIT = 1, 2, 3, 4, 5


class DistDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("ncomp", C.c_int32),
                ("comp_kind", C.c_int32), ("slot", C.c_int32), ("poff", C.c_int32)]


class SlotDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("circ_mask", C.c_int32), ("cap", C.c_int32),
                ("pts_off", C.c_int32)]


class FactorDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("arity", C.c_int32), ("zdim", C.c_int32),
                ("dist", C.c_int32), ("slot", C.c_int32 * IIF_MAX_ARITY), ("nmh", C.c_int32),
                ("partial_mask", C.c_int32), ("solver", C.c_int32), ("_pad", C.c_int32),
                ("mh", C.c_double * IIF_MAX_ARITY),
                ("nullhypo", C.c_double), ("inflation", C.c_double), ("aux", C.c_double * IIF_MAX_DIM)]


class SolverParamsC(C.Structure):
    _fields_ = [("spreadNH", C.c_double), ("nullSurplusAdd", C.c_double),
                ("inflateCycles", C.c_int32), ("gibbsNiter", C.c_int32), ("seed", C.c_uint64)]


class ConvOp(C.Structure):
    _fields_ = [("factor", C.c_int32), ("sfidx", C.c_int32), ("N", C.c_int32),
                ("call_id", C.c_int32), ("nullSurplus", C.c_double), ("meas_off", C.c_int32),
                ("mhidx_off", C.c_int32), ("uinf_off", C.c_int32), ("_pad", C.c_int32)]


class PropOp(C.Structure):
    _fields_ = [("target_slot", C.c_int32), ("out_slot", C.c_int32), ("nfactors", C.c_int32),
                ("N", C.c_int32), ("factor", C.c_int32 * IIF_MAX_FACTORS),
                ("sfidx", C.c_int32 * IIF_MAX_FACTORS), ("call_id", C.c_int32),
                ("any_multihypo", C.c_int32)]


class ProductOp(C.Structure):
    _fields_ = [("dim", C.c_int32), ("circ_mask", C.c_int32), ("nfactors", C.c_int32),
                ("N", C.c_int32), ("call_id", C.c_int32), ("randu_off", C.c_int32),
                ("randn_off", C.c_int32), ("_pad", C.c_int32)]


class DeconvOp(C.Structure):
    _fields_ = [("factor", C.c_int32), ("out_slot", C.c_int32), ("N", C.c_int32), ("call_id", C.c_int32)]


class SchedOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("lane", C.c_int32)]


class XferOp(C.Structure):
    _fields_ = [("slot", C.c_int32), ("peer", C.c_int32), ("msg", C.c_int32), ("_pad", C.c_int32)]


class GraphDesc(C.Structure):
    _fields_ = [("nvars", C.c_int32), ("vars", C.POINTER(SlotDesc)), ("nfactors", C.c_int32),
                ("factors", C.POINTER(FactorDesc)), ("ndists", C.c_int32), ("dists", C.POINTER(DistDesc)),
                ("nparams", C.c_int32), ("dparams", C.POINTER(C.c_double)),
                ("factor_type", C.POINTER(C.c_int32)), ("var_type", C.POINTER(C.c_int32)),
                ("var_relative_kind", C.POINTER(C.c_int32)), ("var_relative_type", C.POINTER(C.c_int32)),
                ("msgprior_type", C.c_int32), ("_pad", C.c_int32)]


_ipt = C.POINTER(C.c_int32)


class TreeDesc(C.Structure):
    _fields_ = [("ncliques", C.c_int32), ("parent", _ipt)] + [
        (f"{name}{suffix}", _ipt)
        for name in ("frontal", "separator", "potential", "directFrtlMsg", "msgskip", "itervar", "directPriorMsg")
        for suffix in ("_off", "s" if name in ("frontal", "separator", "potential") else "")]


class PlanOpts(C.Structure):
    _fields_ = [("N", C.c_int32), ("gibbsIters", C.c_int32), ("downIters", C.c_int32), ("downsolve", C.c_int32),
                ("lanes", C.c_int32), ("forward_copies", C.c_int32), ("useMsgLikelihoods", C.c_int32),
                ("call_base", C.c_int32), ("inflation", C.c_double)]


P = C.POINTER
_dp, _ip, _vp = P(C.c_double), P(C.c_int32), C.c_void_p

# every symbol include/iifb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "iifb200_init": (C.c_int32, [C.c_int32, P(_vp)]),
    "iifb200_free": (None, [_vp]),
    "iifb200_last_error": (C.c_char_p, [_vp]),
    "iifb200_version": (C.c_int32, []),
    "iifb200_arena_bytes": (C.c_int64, [C.c_int32, P(SlotDesc)]),
    "iifb200_set_graph": (C.c_int32, [_vp, C.c_int32, P(SlotDesc), C.c_int32, P(FactorDesc),
                                      C.c_int32, P(DistDesc), C.c_int32, _dp, P(SolverParamsC), _vp]),
    "iifb200_set_solver_params": (C.c_int32, [_vp, P(SolverParamsC)]),
    "iifb200_upload_belief": (C.c_int32, [_vp, C.c_int32, C.c_int32, _dp, _dp, C.c_int32]),
    "iifb200_download_belief": (C.c_int32, [_vp, C.c_int32, _ip, _dp, _dp, _dp]),
    "iifb200_upload_all": (C.c_int32, [_vp, _dp, _dp, _ip, _ip]),
    "iifb200_download_all": (C.c_int32, [_vp, _dp, _dp, _dp, _ip]),
    "iifb200_upload_slots": (C.c_int32, [_vp, C.c_int32, C.c_int32, _dp, _dp, _ip, _ip]),
    "iifb200_download_slots": (C.c_int32, [_vp, C.c_int32, C.c_int32, _dp, _dp, _dp, _ip]),
    "iifb200_host_alloc": (C.c_int32, [_vp, C.c_int64, P(_vp)]),
    "iifb200_host_free": (C.c_int32, [_vp, _vp]),
    "iifb200_slot_device_ptr": (C.c_int32, [_vp, C.c_int32, P(_vp), P(_vp)]),
    "iifb200_conv_batch": (C.c_int32, [_vp, C.c_int32, P(ConvOp), _dp, _ip, _dp, _dp, _dp, _dp, _ip, _ip]),
    "iifb200_product_batch": (C.c_int32, [_vp, C.c_int32, P(ProductOp), _dp, _dp, _ip, _dp, _dp, _dp,
                                          _dp, _dp, _ip]),
    "iifb200_kde_bandwidth": (C.c_int32, [_vp, C.c_int32, _ip, _ip, _ip, _dp, _dp]),
    "iifb200_ppe_batch": (C.c_int32, [_vp, C.c_int32, _ip, _dp, _dp]),
    "iifb200_deconv_batch": (C.c_int32, [_vp, C.c_int32, _ip, _ip, _ip, _dp, _dp]),
    "iifb200_mmd": (C.c_int32, [_vp, C.c_int32, _ip, _ip, _ip, _ip, _dp, _dp, C.c_double, _dp]),
    "iifb200_propagate_batch": (C.c_int32, [_vp, C.c_int32, P(PropOp)]),
    "iifb200_propagate_once": (C.c_int32, [_vp, C.c_int32, P(SlotDesc), C.c_int32, P(FactorDesc), C.c_int32, P(DistDesc),
                                           C.c_int32, _dp, P(SolverParamsC), _dp, _dp, _ip, _ip, P(PropOp), _ip, _dp, _dp, _dp]),
    "iifb200_schedule_build": (C.c_int32, [_vp, C.c_int32, _ip, C.c_int32, P(SchedOp), C.c_int32,
                                           P(PropOp), _ip]),
    "iifb200_schedule_build_ex": (C.c_int32, [_vp, C.c_int32, _ip, C.c_int32, P(SchedOp), C.c_int32,
                                              P(PropOp), C.c_int32, P(DeconvOp), _ip]),
    "iifb200_schedule_build_dist": (C.c_int32, [_vp, C.c_int32, _ip, C.c_int32, P(SchedOp), C.c_int32, P(PropOp),
                                                C.c_int32, P(DeconvOp), C.c_int32, P(XferOp), _ip]),
    "iifb200_ipc_export": (C.c_int32, [_vp, C.c_int32, _vp, _vp]),
    "iifb200_ipc_attach": (C.c_int32, [_vp, C.c_int32, C.c_int32, _vp, _vp]),
    "iifb200_schedule_run": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32]),
    "iifb200_schedule_free": (C.c_int32, [_vp, C.c_int32]),
    "iifb200_schedule_profile": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, P(C.c_float), _ip,
                                             P(C.c_int64)]),
    "iifb200_plan_tree": (C.c_int32, [P(GraphDesc), P(TreeDesc), P(PlanOpts), P(_vp)]),
    "iifb200_plan_error": (C.c_char_p, []),
    "iifb200_plan_free": (None, [_vp]),
    "iifb200_plan_counts": (C.c_int32, [_vp, _ip]),
    "iifb200_plan_export": (C.c_int32, [_vp, P(SlotDesc), P(FactorDesc), P(DistDesc), _dp, P(PropOp), P(SchedOp), _ip]),
    "iifb200_plan_export_deconvs": (C.c_int32, [_vp, P(DeconvOp)]),
    "iifb200_plan_upload": (C.c_int32, [_vp, _vp, P(SolverParamsC), _vp, _ip]),
    "iifb200_elimination_order_nd": (C.c_int32, [C.c_int32, C.c_int32, _ip, _ip, _ip]),
    "iifb200_elimination_order_is": (C.c_int32, [C.c_int32, C.c_int32, _ip, _ip, C.c_int32, _ip]),
    "iifb200_sync": (C.c_int32, [_vp]),
    "iifb200_launch_count": (C.c_int64, [_vp]),
    "iifb200_set_stream": (C.c_int32, [_vp, _vp]),
    "iifb200_stream": (_vp, [_vp]),
    "iifb200_last_elapsed_ms": (C.c_float, [_vp]),
}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libiifb200.so")
_lib = None


class IIFB200Error(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libiifb200.so and bind every declared symbol.  Raises if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise IIFB200Error(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the product path.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_dp(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def as_ip(a):
    return a.ctypes.data_as(_ip) if a is not None else None
