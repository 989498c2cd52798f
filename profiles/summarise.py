"""Turn the raw ncu artefacts a `gpurun` call left in gpurun_out/ into the tracked summaries under
profiles/:  r<NN>_launches_summary.csv (per-kernel totals of the `--metrics gpu__time_duration.sum`
launch list), r<NN>_kernels.json (key metrics of the `--set full` captures, incl. DRAM traffic per
launch) and r<NN>_summary.md.   usage: python profiles/summarise.py r01"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max']

lines = [f"# ncu summary {tag}", ""]
ll = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        k = r[ik].split('(')[0]
        agg[k][0] += 1
        agg[k][1] += float(r[iv].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(dst, f"{tag}_launches_summary.csv"), "w") as f:
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{v[0]},{v[1]/1e3:.1f},{100*v[1]/tot:.2f},{v[1]/v[0]/1e3:.2f}\n")
    lines += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised)", "",
              "| kernel | launches | total µs | share | avg µs |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:60]}` | {v[0]} | {v[1]/1e3:.1f} | {100*v[1]/tot:.1f}% | {v[1]/v[0]/1e3:.1f} |")
    lines.append("")
kern = {}
for fn in sorted(os.listdir(src)):
    if fn.startswith(tag) and fn.endswith(".ncu-rep"):
        raw = subprocess.run(["ncu", "-i", os.path.join(src, fn), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        if len(rr) < 3:
            continue
        h, units, r = rr[0], rr[1], rr[2]
        d = {"kernel": r[h.index("Kernel Name")] if "Kernel Name" in h else fn}
        for k in KEYS:
            if k in h:
                d[k] = [r[h.index(k)], units[h.index(k)]]
        kern[fn[:-8]] = d
        lines += [f"## `{fn}` — {d['kernel'].split('(')[0]}", ""] + [f"* `{k}` = {v[0]} {v[1]}" for k, v in d.items() if k != "kernel"] + [""]
json.dump(kern, open(os.path.join(dst, f"{tag}_kernels.json"), "w"), indent=1)
open(os.path.join(dst, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
