"""Summarise an ncu report's SASS page: opcode mix, stall mix and the hottest instruction windows.
usage: python profiles/sass_hot.py <report.ncu-rep> [window] [ntop]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]; W = int(sys.argv[2]) if len(sys.argv) > 2 else 40; NT = int(sys.argv[3]) if len(sys.argv) > 3 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
for r in rows[2:3]:
    for w in ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
              'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active']:
        if w in hdr: print(w, r[hdr.index(w)])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]
ia, isrc, ismp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
inst = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ('Kernel Name', 'Address'):
        if len(inst) > 100: break
        continue
    inst.append(r)
def op(r):
    t = r[isrc].split(); o = t[1] if t[0].startswith('@') else t[0]; return o.split('.')[0]
tot = sum(int(r[ia]) for r in inst); nsmp = sum(int(r[ismp]) for r in inst)
print('sass instrs', len(inst), 'executed', tot, 'samples', nsmp)
c = Counter()
for r in inst: c[op(r)] += int(r[ia])
print('opcodes', c.most_common(14))
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = Counter()
for r in inst:
    for i in stall: agg[hdr[i]] += int(r[i] or 0)
print('stalls', [(k, v) for k, v in agg.most_common(9)])
b = [(sum(int(r[ismp]) for r in inst[i:i + W]), sum(int(r[ia]) for r in inst[i:i + W]), i) for i in range(0, len(inst), W)]
for sm, n, i in sorted(b, reverse=True)[:NT]:
    st = Counter()
    for r in inst[i:i + W]:
        for k in stall: st[hdr[k]] += int(r[k] or 0)
    print(f'win {i:6d} samples {sm:6d} ({100*sm/max(nsmp,1):4.1f}%) exec {n:8d} ({100*n/tot:4.1f}%)', Counter(op(r) for r in inst[i:i+W]).most_common(4), st.most_common(3))
