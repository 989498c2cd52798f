"""cluster size of the tree-top products (development).  usage: python profiles/knob_sweep2.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in (0, 4, 6, 8):
    e = dict(os.environ, IIFB200_PROD_CLUSTER_BIG=str(v))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline",
                          "--no-b3"], capture_output=True, text=True, env=e).stdout
    d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    print(f"IIFB200_PROD_CLUSTER_BIG={v}: {d['ms_per_step']:.3f} ms/solve", {a: round(b, 2) for a, b in d["roofline"]["kernel_ms"].items()}, flush=True)
