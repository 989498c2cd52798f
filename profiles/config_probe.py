"""Solve time of the other BASELINE configs on ONE B200 (full sizes; C4 / C5 are quoted on 4 / 8 GPUs in BASELINE.json,
here everything is batched on one device): C3 four-door (N=200, Mixture priors), C4 500-pose Circular chain (N=150),
C5 5000-pose Euclid(2) grid with loop closures (N=100).  Device time of the whole solveTree pass (CUDA events around the
graph launch), best of 5 after 2 warm-up solves.   usage (GPU box): python profiles/config_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iifb200  # noqa: E402,F401
from iifb200 import compile as CP  # noqa: E402
from iifb200 import planner as PL  # noqa: E402
from iifb200 import solver as SV  # noqa: E402
from iifb200 import tree as TR  # noqa: E402
from iifb200 import workloads as W  # noqa: E402


def probe(name, fg, order):
    t0 = time.time()
    ts = SV.TreeSolver(fg, order)
    t_plan = time.time() - t0
    ts.load_from_graph()
    ts.upload()
    best = None
    for it in range(7):
        ts.eng.set_solver_params(CP.solver_params_c(fg.solverParams, 100 + it))
        ts.upload()
        ts.run()
        ts.eng.sync()
        ms = ts.eng.last_elapsed_ms()
        if it >= 2:
            best = ms if best is None else min(best, ms)
    p = ts.plan
    print(f"{name}: {len(fg.variables)} variables, {len(ts.tree.cliques)} cliques, {len(p.wave_off) - 1} waves, "
          f"{p.n_conv} convolutions + {p.n_prod} propagateBeliefs per solve: {best:.3f} ms  ->  "
          f"{p.n_conv / best / 1e3:.3f} M conv/s   (plan lowering on the host: {t_plan:.1f} s)")
    ts.close()


fg = W.four_door(N=200, seed=42)
probe("C3 four-door N=200", fg, TR.getEliminationOrder(fg, "qr"))
fg = W.circular_chain(n=500, N=150, seed=42)
probe("C4 circular chain 500 poses N=150", fg, PL.elimination_order_is(fg))
fg = W.euclid2_grid(rows=50, cols=100, N=100, seed=42, closure_every=5)
probe("C5 Euclid(2) grid 5000 poses N=100", fg, PL.elimination_order_is(fg) if os.environ.get("IIFB200_PROBE_ORDER", "is") == "is" else TR.getEliminationOrder(fg, "nd"))
