"""Launch-shape knob sweep on one GPU (development): CTA size of the wide product / convolution launches and the number
of lanes, each through a short bench.py run.  usage: python profiles/knob_sweep.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env):
    e = dict(os.environ, **{k: str(v) for k, v in env.items()})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline",
                          "--no-b3"], capture_output=True, text=True, env=e).stdout
    d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    return d["ms_per_step"], d["roofline"]["kernel_ms"]


for name, vals in (("IIFB200_PROD_WIDE_THREADS", (128, 160, 256, 384, 512)), ("IIFB200_WIDE_THREADS", (128, 160, 256)),
                   ("IIFB200_LANES", (2, 4, 6, 8))):
    for v in vals:
        ms, k = run({name: v})
        print(f"{name}={v}: {ms:.3f} ms/solve  kernels {({a: round(b, 2) for a, b in k.items()})}", flush=True)
