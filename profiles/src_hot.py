"""Per-source-line cost of an ncu report (needs -lineinfo + --import-source on): instructions executed,
stall samples and the FP64 share per CUDA source line, top lines first.
usage: python profiles/src_hot.py <report.ncu-rep> [ntop]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fpath, hdr, lines = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ism = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":                     # a CUDA source line with its aggregated metrics
        try:
            key = (fpath, int(r[0]))
            lines[key] = [r[1].strip(), int(r[ie]), int(r[ism]), 0]
            cur = key
        except ValueError:
            cur = None
    elif r[2] not in ("...", "-") and cur is not None:   # a SASS row under the current line
        t = r[3].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        if op in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
            try:
                lines[cur][3] += int(r[ie])
            except ValueError:
                pass
tot = sum(v[1] for v in lines.values()) or 1
smp = sum(v[2] for v in lines.values()) or 1
print(f"total instructions {tot}  samples {smp}  fp64 instr {sum(v[3] for v in lines.values())}")
print(f"{'file:line':28s} {'instr%':>7s} {'smp%':>6s} {'fp64%':>6s}  source")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:ntop]:
    print(f"{f + ':' + str(ln):28s} {100 * v[1] / tot:7.2f} {100 * v[2] / smp:6.2f} {100 * v[3] / max(v[1], 1):6.1f}  {v[0][:90]}")
