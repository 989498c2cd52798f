"""Per-wave timing of the C2 solve (1000-pose scalar chain, N=100): CUDA-event time of the copy / convolution /
product kernel of every wave (iifb200_schedule_profile on single-wave ranges, warm), next to the wave's width.
Shows where the pass is throughput-bound (wide waves at the leaves) and where it is latency-bound (tree top).
usage (GPU box): python profiles/wave_probe.py [poses]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iifb200  # noqa: E402,F401
from iifb200 import _abi as A  # noqa: E402
from iifb200 import solver as SV  # noqa: E402
from iifb200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
fg = W.scalar_chain(n, N=100, seed=42)
from iifb200 import planner as PL  # noqa: E402
ts = SV.TreeSolver(fg, PL.elimination_order_is(fg))   # the bench's order
ts.load_from_graph()
ts.upload()
for _ in range(3):
    ts.run()
ts.eng.sync()
plan = ts.plan
nw = len(plan.wave_off) - 1
tot = {"conv": 0.0, "product": 0.0, "copy": 0.0}
print(f"{'wave':>4} {'copies':>6} {'convs':>6} {'prods':>6} | {'copy us':>8} {'conv us':>8} {'prod us':>8}")
for w in range(nw):
    best = None
    for rep in range(3):
        r = ts.eng.schedule_profile(ts.sid, w, w + 1)
        if best is None or sum(v[0] for v in r.values()) < sum(v[0] for v in best.values()):
            best = r
    for k in tot:
        tot[k] += best[k][0]
    print(f"{w:4d} {best['copy'][2]:6d} {best['conv'][2]:6d} {best['product'][2]:6d} | {best['copy'][0]*1e3:8.1f} "
          f"{best['conv'][0]*1e3:8.1f} {best['product'][0]*1e3:8.1f}")
print("totals (ms):", {k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
ts.run()
ts.eng.sync()
print("whole pass as one CUDA graph (ms):", round(ts.eng.last_elapsed_ms(), 3))
ts.close()
