"""Throughput when several independent graphs are solved by ONE pass: B independent 1000-pose scalar chains (N=100)
in one factor graph (a forest; one Bayes-tree root per session).  Every wave of the pass then holds B times the work at
the same depth, so the latency of the narrow tree-top waves is shared by B solves.  Device time per pass (CUDA events
around the graph launch), best of 5 after 2 warm-up passes.   usage (GPU box): python profiles/sessions_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iifb200  # noqa: E402,F401
from iifb200 import compile as CP  # noqa: E402
from iifb200 import solver as SV  # noqa: E402
from iifb200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
print(f"{'sessions':>8} {'waves':>6} {'conv/pass':>10} {'ms/pass':>9} {'ms/session':>11} {'M conv/s':>9}")
for B in (1, 2, 4, 8):
    fg = W.scalar_chain_sessions(B, n, N=100, seed=42)
    ts = SV.TreeSolver(fg, W.sessions_nd_order(B, n))
    ts.load_from_graph()
    best = None
    for it in range(7):
        ts.eng.set_solver_params(CP.solver_params_c(fg.solverParams, 100 + it))
        ts.upload()
        ts.run()
        ts.eng.sync()
        ms = ts.eng.last_elapsed_ms()
        if it >= 2:
            best = ms if best is None else min(best, ms)
    ts.download()
    err = max(abs(ts.arena.get(ts.plan.var_slot[f"s{b}x{n - 1}"])[0].mean() - (10.0 * b + n - 1)) for b in range(B))
    p = ts.plan
    print(f"{B:8d} {len(p.wave_off) - 1:6d} {p.n_conv:10d} {best:9.3f} {best / B:11.3f} {p.n_conv / best / 1e3:9.3f}"
          f"   (last-pose mean error max {err:.2f})")
    ts.close()
