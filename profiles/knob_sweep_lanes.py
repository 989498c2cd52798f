import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in (1, 2, 3, 4, 5, 6, 8):
    e = dict(os.environ, IIFB200_LANES=str(v))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--no-b3"], capture_output=True, text=True, env=e).stdout
    d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    print(f"IIFB200_LANES={v}: {d['ms_per_step']:.3f} ms/solve", {a: round(b, 2) for a, b in d["roofline"]["kernel_ms"].items()}, flush=True)
