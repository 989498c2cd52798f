"""Thin Python binding of the C-ABI (include/iifb200.h): one Engine == one iifb200_ctx.

This is the analogue of the Julia `ccall` shim in INTEGRATION.md; it holds no numerics.  Every
method ends in a kernel launch inside libiifb200.so or raises IIFB200Error — there is no CPU
fallback.
"""
import ctypes as C

import numpy as np

from . import _abi as A


class Engine:
    def __init__(self, frozen, sp_c, device=0, ext_arena_ptr=None, _plan_handle=None):
        self.lib = A.load_library()
        self.frozen = frozen
        self.sp_c = sp_c
        ctx = C.c_void_p()
        st = self.lib.iifb200_init(device, C.byref(ctx))
        if st != A.IIF_OK:
            raise A.IIFB200Error(f"iifb200_init failed ({st}): {self.lib.iifb200_last_error(None).decode()}")
        self.ctx = ctx
        if _plan_handle is not None:      # iifb200_plan_upload: set_graph + schedule_build inside the library
            sid = C.c_int32(-1)
            self._check(self.lib.iifb200_plan_upload(ctx, _plan_handle, C.byref(sp_c),
                                                     C.c_void_p(ext_arena_ptr) if ext_arena_ptr else None,
                                                     C.cast(C.byref(sid), A._ip)), "plan_upload")
            self.plan_sid = sid.value
            return
        self._check(self.lib.iifb200_set_graph(
            ctx, frozen["nslots"], frozen["slots"], frozen["nfactors"], frozen["factors"], frozen["ndists"],
            frozen["dists"], frozen["nparams"], A.as_dp(frozen["dparams"]), C.byref(sp_c),
            C.c_void_p(ext_arena_ptr) if ext_arena_ptr else None), "set_graph")

    @classmethod
    def from_plan(cls, frozen, plan_handle, sp_c, device=0, ext_arena_ptr=None):
        """engine + schedule from a plan made by iifb200_plan_tree -> (Engine, schedule id)"""
        e = cls(frozen, sp_c, device, ext_arena_ptr, _plan_handle=plan_handle)
        return e, e.plan_sid

    def reset_graph(self, frozen, sp_c=None):
        """iifb200_set_graph again on the same context (arena, tables and scratch are re-used, grow-only)"""
        self.frozen = frozen
        if sp_c is not None:
            self.sp_c = sp_c
        self._check(self.lib.iifb200_set_graph(
            self.ctx, frozen["nslots"], frozen["slots"], frozen["nfactors"], frozen["factors"], frozen["ndists"],
            frozen["dists"], frozen["nparams"], A.as_dp(frozen["dparams"]), C.byref(self.sp_c), None), "set_graph")

    def propagate_once(self, frozen, pts, bw, npts, flags, prop_op, sp_c=None):
        """iifb200_propagate_once: set_graph + upload of all slots + ONE propagateBelief + download of its destination in a
        single C-ABI call (boundary B3); returns (points, bw, ipc) of prop_op's out_slot"""
        self.frozen = frozen
        if sp_c is not None:
            self.sp_c = sp_c
        out = prop_op[0].out_slot
        s = frozen["slots"][out]
        n = C.c_int32(0)
        opts = np.zeros((max(prop_op[0].N, 1), s.dim))
        obw, oipc = np.zeros(A.IIF_MAX_DIM), np.zeros(A.IIF_MAX_DIM)
        self._check(self.lib.iifb200_propagate_once(
            self.ctx, frozen["nslots"], frozen["slots"], frozen["nfactors"], frozen["factors"], frozen["ndists"],
            frozen["dists"], frozen["nparams"], A.as_dp(frozen["dparams"]), C.byref(self.sp_c), A.as_dp(pts), A.as_dp(bw),
            A.as_ip(npts), A.as_ip(flags), prop_op, C.cast(C.byref(n), A._ip), A.as_dp(opts), A.as_dp(obw), A.as_dp(oipc)),
            "propagate_once")
        return opts[:n.value], obw[:s.dim].copy(), oipc[:s.dim].copy()

    # ---- plumbing
    def _check(self, st, what):
        if st != A.IIF_OK:
            msg = self.lib.iifb200_last_error(self.ctx).decode()
            raise A.IIFB200Error(f"{what} failed ({st}): {msg}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.iifb200_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_solver_params(self, sp_c):
        self.sp_c = sp_c
        self._check(self.lib.iifb200_set_solver_params(self.ctx, C.byref(sp_c)), "set_solver_params")

    # ---- beliefs
    def upload_arena(self, arena):
        self._check(self.lib.iifb200_upload_all(self.ctx, A.as_dp(arena.pts), A.as_dp(arena.bw),
                                                A.as_ip(arena.npts), A.as_ip(arena.flags)), "upload_all")

    def download_arena(self, arena):
        self._check(self.lib.iifb200_download_all(self.ctx, A.as_dp(arena.pts), A.as_dp(arena.bw),
                                                  A.as_dp(arena.ipc), A.as_ip(arena.npts)), "download_all")
        arena.flags[:] = (arena.npts > 0).astype(np.int32) | arena.flags

    def upload_belief(self, slot, pts, bw=None, initialized=True):
        s = self.frozen["slots"][slot]
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, s.dim)
        bwa = None if bw is None else np.ascontiguousarray(bw, dtype=np.float64)
        self._check(self.lib.iifb200_upload_belief(self.ctx, slot, pts.shape[0], A.as_dp(pts), A.as_dp(bwa),
                                                   1 if initialized else 0), "upload_belief")

    def download_belief(self, slot):
        s = self.frozen["slots"][slot]
        n = C.c_int32(0)
        pts = np.zeros((s.cap, s.dim))
        bw, ipc = np.zeros(A.IIF_MAX_DIM), np.zeros(A.IIF_MAX_DIM)
        self._check(self.lib.iifb200_download_belief(self.ctx, slot, C.cast(C.byref(n), A._ip), A.as_dp(pts),
                                                     A.as_dp(bw), A.as_dp(ipc)), "download_belief")
        return pts[:n.value].copy(), bw[:s.dim].copy(), ipc[:s.dim].copy()

    def host_alloc(self, nbytes):
        """pinned host buffer (numpy uint8 view) for upload_slots / download_slots"""
        p = C.c_void_p()
        self._check(self.lib.iifb200_host_alloc(self.ctx, nbytes, C.byref(p)), "host_alloc")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def upload_slots(self, first, count, pts, bw, npts, flags):
        self._check(self.lib.iifb200_upload_slots(self.ctx, first, count, A.as_dp(pts), A.as_dp(bw), A.as_ip(npts),
                                                  A.as_ip(flags)), "upload_slots")

    def download_slots(self, first, count, pts, bw, ipc, npts):
        self._check(self.lib.iifb200_download_slots(self.ctx, first, count, A.as_dp(pts), A.as_dp(bw), A.as_dp(ipc),
                                                    A.as_ip(npts)), "download_slots")

    def slot_device_ptr(self, slot):
        p, b = C.c_void_p(), C.c_void_p()
        self._check(self.lib.iifb200_slot_device_ptr(self.ctx, slot, C.byref(p), C.byref(b)), "slot_device_ptr")
        return p.value, b.value

    # ---- hot path
    def conv_batch(self, ops, K, meas=None, mhidx=None, uinf=None):
        """ops: ctypes array of ConvOp.  Returns list of (pts N x d, bw, ipc, mhidx, nan)."""
        dims, Ns = [], []
        for k in range(K):
            # invalid descriptors are sized conservatively here and rejected by the C side
            ok = 0 <= ops[k].factor < self.frozen["nfactors"]
            f = self.frozen["factors"][ops[k].factor] if ok else None
            ok = ok and 1 <= ops[k].sfidx <= f.arity
            dims.append(self.frozen["slots"][f.slot[ops[k].sfidx - 1]].dim if ok else A.IIF_MAX_DIM)
            Ns.append(max(int(ops[k].N), 1))
        tot = sum(n * d for n, d in zip(Ns, dims))
        out_pts = np.zeros(tot)
        out_bw = np.zeros(K * A.IIF_MAX_DIM)
        out_ipc = np.zeros(K * A.IIF_MAX_DIM)
        out_lab = np.zeros(sum(Ns), dtype=np.int32)
        out_nan = np.zeros(K, dtype=np.int32)
        meas = None if meas is None else np.ascontiguousarray(meas, dtype=np.float64)
        mhidx = None if mhidx is None else np.ascontiguousarray(mhidx, dtype=np.int32)
        uinf = None if uinf is None else np.ascontiguousarray(uinf, dtype=np.float64)
        self._check(self.lib.iifb200_conv_batch(self.ctx, K, ops, A.as_dp(meas), A.as_ip(mhidx), A.as_dp(uinf),
                                                A.as_dp(out_pts), A.as_dp(out_bw), A.as_dp(out_ipc),
                                                A.as_ip(out_lab), A.as_ip(out_nan)), "conv_batch")
        res, po, no = [], 0, 0
        for k in range(K):
            n, d = Ns[k], dims[k]
            res.append((out_pts[po:po + n * d].reshape(n, d).copy(),
                        out_bw[k * A.IIF_MAX_DIM:k * A.IIF_MAX_DIM + d].copy(),
                        out_ipc[k * A.IIF_MAX_DIM:k * A.IIF_MAX_DIM + d].copy(),
                        out_lab[no:no + n].copy(), int(out_nan[k])))
            po += n * d
            no += n
        return res

    def product(self, dens_pts, dens_bw, dim, circ_mask=0, dens_mask=None, old_pts=None, call_id=0,
                randU=None, randN=None):
        """One AMP.manifoldProduct: dens_pts F x N x d, dens_bw F x d -> (pts, bw, labels N x F)."""
        dens_pts = np.ascontiguousarray(dens_pts, dtype=np.float64)
        F, N = dens_pts.shape[0], dens_pts.shape[1]
        op = (A.ProductOp * 1)()
        op[0].dim, op[0].circ_mask, op[0].nfactors, op[0].N = dim, circ_mask, F, N
        op[0].call_id = call_id
        op[0].randu_off = 0 if randU is not None else -1
        op[0].randn_off = 0 if randN is not None else -1
        bwp = np.zeros((F, A.IIF_MAX_DIM))
        bwp[:, :dim] = np.asarray(dens_bw, dtype=np.float64).reshape(F, dim)
        mask = None if dens_mask is None else np.ascontiguousarray(dens_mask, dtype=np.int32)
        old = None if old_pts is None else np.ascontiguousarray(old_pts, dtype=np.float64)
        ru = None if randU is None else np.ascontiguousarray(randU, dtype=np.float64)
        rn = None if randN is None else np.ascontiguousarray(randN, dtype=np.float64)
        out = np.zeros((N, dim))
        obw = np.zeros(A.IIF_MAX_DIM)
        lab = np.zeros((N, F), dtype=np.int32)
        self._check(self.lib.iifb200_product_batch(self.ctx, 1, op, A.as_dp(dens_pts), A.as_dp(bwp), A.as_ip(mask),
                                                   A.as_dp(old), A.as_dp(ru), A.as_dp(rn), A.as_dp(out),
                                                   A.as_dp(obw), A.as_ip(lab)), "product_batch")
        return out, obw[:dim].copy(), lab

    def kde_bandwidth(self, pts, circ_mask=0):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        if pts.ndim == 1:
            pts = pts.reshape(-1, 1)
        n, d = pts.shape
        N = np.array([n], dtype=np.int32)
        D = np.array([d], dtype=np.int32)
        M = np.array([circ_mask], dtype=np.int32)
        bw = np.zeros(A.IIF_MAX_DIM)
        self._check(self.lib.iifb200_kde_bandwidth(self.ctx, 1, A.as_ip(N), A.as_ip(D), A.as_ip(M), A.as_dp(pts),
                                                   A.as_dp(bw)), "kde_bandwidth")
        return bw[:d].copy()

    def ppe_batch(self, slots):
        """calcPPE of device-resident beliefs -> (mean, max) arrays of shape (V, IIF_MAX_DIM)"""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        V = len(sl)
        mean, mx = np.zeros((V, A.IIF_MAX_DIM)), np.zeros((V, A.IIF_MAX_DIM))
        self._check(self.lib.iifb200_ppe_batch(self.ctx, V, A.as_ip(sl), A.as_dp(mean), A.as_dp(mx)), "ppe_batch")
        return mean, mx

    def deconv(self, factor, N, call_id):
        """approxDeconv of one factor -> (predicted N x z, sampled N x z)"""
        zd = self.frozen["factors"][factor].zdim
        pred, meas = np.zeros((N, zd)), np.zeros((N, zd))
        f, n, c = (np.array([x], dtype=np.int32) for x in (factor, N, call_id))
        self._check(self.lib.iifb200_deconv_batch(self.ctx, 1, A.as_ip(f), A.as_ip(n), A.as_ip(c), A.as_dp(pred),
                                                  A.as_dp(meas)), "deconv_batch")
        return pred, meas

    def mmd(self, a, b, circ_mask=0, bw=0.001):
        """AMP.mmd kernel-embedding distance between two point sets"""
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(len(a), -1)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(len(b), -1)
        na, nb, d, cm = (np.array([x], dtype=np.int32) for x in (a.shape[0], b.shape[0], a.shape[1], circ_mask))
        out = np.zeros(1)
        self._check(self.lib.iifb200_mmd(self.ctx, 1, A.as_ip(na), A.as_ip(nb), A.as_ip(d), A.as_ip(cm), A.as_dp(a),
                                         A.as_dp(b), float(bw), A.as_dp(out)), "mmd")
        return float(out[0])

    def propagate_batch(self, prop_ops, V):
        self._check(self.lib.iifb200_propagate_batch(self.ctx, V, prop_ops), "propagate_batch")

    # ---- schedules
    def schedule_build(self, wave_off, sched_ops, nops, prop_ops, nprops, deconv_ops=None, ndeconvs=0):
        wo = np.ascontiguousarray(wave_off, dtype=np.int32)
        sid = C.c_int32(-1)
        if ndeconvs:
            self._check(self.lib.iifb200_schedule_build_ex(self.ctx, len(wo) - 1, A.as_ip(wo), nops, sched_ops, nprops,
                                                           prop_ops, ndeconvs, deconv_ops,
                                                           C.cast(C.byref(sid), A._ip)), "schedule_build_ex")
        else:
            self._check(self.lib.iifb200_schedule_build(self.ctx, len(wo) - 1, A.as_ip(wo), nops, sched_ops, nprops,
                                                        prop_ops, C.cast(C.byref(sid), A._ip)), "schedule_build")
        return sid.value

    def schedule_build_dist(self, wave_off, sched_ops, nops, prop_ops, nprops, deconv_ops, ndeconvs, xfer_ops, nxfers):
        wo = np.ascontiguousarray(wave_off, dtype=np.int32)
        sid = C.c_int32(-1)
        self._check(self.lib.iifb200_schedule_build_dist(self.ctx, len(wo) - 1, A.as_ip(wo), nops, sched_ops, nprops,
                                                         prop_ops, ndeconvs, deconv_ops if ndeconvs else None, nxfers,
                                                         xfer_ops if nxfers else None, C.cast(C.byref(sid), A._ip)),
                    "schedule_build_dist")
        return sid.value

    def ipc_export(self, nflags):
        """-> (arena handle, flags handle): 64-byte CUDA IPC handles as bytes"""
        ha, hf = C.create_string_buffer(64), C.create_string_buffer(64)
        self._check(self.lib.iifb200_ipc_export(self.ctx, int(nflags), C.cast(ha, C.c_void_p), C.cast(hf, C.c_void_p)), "ipc_export")
        return ha.raw, hf.raw

    def ipc_attach(self, world, rank, arena_handles: bytes, flags_handles: bytes):
        a, f = C.create_string_buffer(arena_handles, len(arena_handles)), C.create_string_buffer(flags_handles, len(flags_handles))
        self._check(self.lib.iifb200_ipc_attach(self.ctx, world, rank, C.cast(a, C.c_void_p), C.cast(f, C.c_void_p)), "ipc_attach")

    def schedule_run(self, sid, first=0, last=-1):
        self._check(self.lib.iifb200_schedule_run(self.ctx, sid, first, last), "schedule_run")

    def schedule_profile(self, sid, first=0, last=-1):
        """per-kernel CUDA-event totals: dict(conv|product|copy -> (ms, launches, blocks))"""
        ms = (C.c_float * 3)()
        ln = (C.c_int32 * 3)()
        bl = (C.c_int64 * 3)()
        self._check(self.lib.iifb200_schedule_profile(self.ctx, sid, first, last, ms, ln, bl), "schedule_profile")
        return {k: (float(ms[i]), int(ln[i]), int(bl[i])) for i, k in enumerate(("conv", "product", "copy"))}

    def sync(self):
        self._check(self.lib.iifb200_sync(self.ctx), "sync")

    def launch_count(self):
        return int(self.lib.iifb200_launch_count(self.ctx))

    def last_elapsed_ms(self):
        return float(self.lib.iifb200_last_elapsed_ms(self.ctx))

    def set_stream(self, stream_ptr):
        self._check(self.lib.iifb200_set_stream(self.ctx, C.c_void_p(stream_ptr) if stream_ptr else None), "set_stream")

    def stream(self):
        return self.lib.iifb200_stream(self.ctx)
