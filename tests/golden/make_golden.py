"""Generates tests/golden/hotpath_golden.npz: input-independent regression vectors of the hot path.

The reference (Julia) cannot be executed in the build image and ships no golden vectors for this
path, so these fixtures are produced by the CPU oracle on the seeded cases of
tests/parity_cases.py.  They freeze the oracle's behaviour (any change to the restated algorithm
shows up as a diff here) and give the `-m gpu` tests a committed target that does not need the
oracle at run time.  Re-generate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity_cases as PC  # noqa: E402
from iifb200 import compile as CP  # noqa: E402


def uml_case():
    """13-pose scalar chain, N=64, nested-dissection order, useMsgLikelihoods=true, solved by the oracle"""
    import oracle as O
    from iifb200 import tree as TR
    from iifb200 import workloads as W
    fg, order = W.scalar_chain(13, N=64, seed=3), W.chain_nd_order(13)
    fg.solverParams.useMsgLikelihoods = True
    plan = TR.compile_solve(fg, TR.buildTree(fg, order))
    arena = CP.HostArena(plan.frozen)
    for l, v in fg.variables.items():
        arena.set(plan.var_slot[l], v.val, v.bw, True, v.infoPerCoord)
    orc = O.Oracle(plan.frozen, arena, CP.solver_params_c(fg.solverParams))
    orc.schedule_run(plan.wave_off, CP.make_sched_ops(plan.sched_waved), CP.make_prop_ops(plan.props),
                     deconvs=CP.make_deconv_ops(plan.deconvs or []))
    return fg, order, plan, arena


def main():
    out = {}
    for name, P, specs, streams in PC.conv_cases():
        ops = CP.make_conv_ops(specs)
        st = streams or {}
        orc = P.oracle()
        for k in range(len(specs)):
            pts, bw, ipc, lab, nan = orc.conv(ops[k], st.get("meas"), st.get("mhidx"), st.get("uinf"))
            out[f"conv/{name}/{k}/pts"], out[f"conv/{name}/{k}/bw"] = pts, bw
            out[f"conv/{name}/{k}/mhidx"] = lab
    import oracle as O
    for name, kw in PC.product_cases():
        kw = dict(kw)
        dim = kw.pop("dim")
        pts, bw, lab = O.product(kw["dens_pts"], kw["dens_bw"], dim, kw.get("circ_mask", 0), kw.get("dens_mask"),
                                 kw.get("old_pts"), 3, kw.get("call_id", 0), 1, kw.get("randU"), kw.get("randN"))
        out[f"product/{name}/pts"], out[f"product/{name}/bw"], out[f"product/{name}/labels"] = pts, bw, lab
    P, xs, fs = PC.chain_problem(n=4, N=100, seed=21)
    specs, sched, wave_off = [], [], [0]
    for sweep in range(3):
        for s in PC.chain_prop_specs(xs, fs, 100, call0=5000 + 1000 * sweep):
            specs.append(s)
            sched.append((CP.A.S_PROPAGATE, len(specs) - 1, 0))
            wave_off.append(len(sched))
    orc = P.oracle()
    orc.schedule_run(wave_off, CP.make_sched_ops(sched), CP.make_prop_ops(specs))
    for k, x in enumerate(xs):
        pts, bw, ipc = orc.arena.get(x)
        out[f"schedule/chain4/x{k}/pts"], out[f"schedule/chain4/x{k}/bw"] = pts, bw
    # a whole tree solve with differential messages (useMsgLikelihoods=true), incl. the IIF_S_DECONV slots
    fg, order, plan, arena = uml_case()
    for l in fg.variables:
        pts, bw, _ = arena.get(plan.var_slot[l])
        out[f"uml/chain13/{l}/pts"], out[f"uml/chain13/{l}/bw"] = pts, bw
    for k, dc in enumerate(plan.deconvs):
        pts, bw, _ = arena.get(dc["out_slot"])
        out[f"uml/chain13/diff{k}/pts"], out[f"uml/chain13/diff{k}/bw"] = pts, bw
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
