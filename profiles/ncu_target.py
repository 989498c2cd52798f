"""Small target for ncu captures: runs one tree solve of a scalar chain WITHOUT the CUDA graph
(every kernel is a plain launch).  usage: python profiles/ncu_target.py <poses> <nd|natural> [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iifb200  # noqa: E402
from iifb200 import solver as SV  # noqa: E402
from iifb200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "nd"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
fg = W.scalar_chain(n, N=100, seed=42)
order = W.chain_nd_order(n) if kind == "nd" else [f"x{k}" for k in range(n)]
ts = SV.TreeSolver(fg, order)
ts.load_from_graph()
ts.upload()
for _ in range(reps):
    prof = ts.eng.schedule_profile(ts.sid)
print({k: (round(v[0], 3), v[1], v[2]) for k, v in prof.items()})
