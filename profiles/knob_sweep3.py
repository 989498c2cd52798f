"""CTA size of the wide product launches after the 8-candidate pieces (3 CTAs per SM fit in shared memory).
usage: python profiles/knob_sweep3.py [values...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vals = [int(x) for x in sys.argv[1:]] or [128, 160, 192, 224, 256]
for v in vals:
    e = dict(os.environ, IIFB200_PROD_WIDE_THREADS=str(v))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline",
                          "--no-b3"], capture_output=True, text=True, env=e).stdout
    d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    print(f"IIFB200_PROD_WIDE_THREADS={v}: {d['ms_per_step']:.3f} ms/solve", {a: round(b, 2) for a, b in d["roofline"]["kernel_ms"].items()},
          "oracle-equal poses", d.get("posterior_check", {}).get("vs_oracle", {}).get("poses_all_points_equal"), flush=True)
