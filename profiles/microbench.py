"""Micro-benchmarks of the building blocks (device time from CUDA events inside the library)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import parity_cases as PC
from iifb200 import _abi as A, compile as CP, graph as G

P, xs, fs = PC.chain_problem(n=3, N=100, seed=1)
eng = P.engine()
R = np.random.default_rng(0)

def bw_time(K, N, d=1, reps=5):
    pts = R.normal(0, 1, (K, N, d))
    Ns = np.full(K, N, dtype=np.int32); Ds = np.full(K, d, dtype=np.int32); Ms = np.zeros(K, dtype=np.int32)
    out = np.zeros(K * 4)
    best = 1e9
    for _ in range(reps):
        eng._check(eng.lib.iifb200_kde_bandwidth(eng.ctx, K, A.as_ip(Ns), A.as_ip(Ds), A.as_ip(Ms), A.as_dp(pts), A.as_dp(out)), "bw")
        best = min(best, eng.last_elapsed_ms())
    return best * 1e3

for K in (1, 148, 296, 1184, 4736):
    t = bw_time(K, 100)
    print(f"bandwidth kernel K={K:5d} N=100: {t:9.1f} us  -> {t/K*148 if K>=148 else t:8.1f} us*SM per belief")
for N in (50, 100, 150, 200, 256):
    print(f"bandwidth kernel K=1 N={N}: {bw_time(1, N):8.1f} us")

def conv_time(K, reps=5):
    specs = [dict(factor=fs[1], sfidx=2, N=100, call_id=16 * k) for k in range(K)]
    ops = CP.make_conv_ops(specs)
    best = 1e9
    for _ in range(reps):
        eng.conv_batch(ops, K)
        best = min(best, eng.last_elapsed_ms())
    return best * 1e3

for K in (1, 148, 1184, 4736):
    t = conv_time(K)
    print(f"conv kernel K={K:5d}: {t:9.1f} us -> {t/K*148 if K>=148 else t:8.1f} us*SM per conv")

def prod_time(V, F=2, N=100, reps=3):
    a = R.normal(0, 1, (V, F, N, 1))
    import oracle as O
    bws = np.zeros((V * F, 4)); bws[:, 0] = 0.4
    ops = (A.ProductOp * V)()
    for v in range(V):
        ops[v].dim, ops[v].circ_mask, ops[v].nfactors, ops[v].N, ops[v].call_id = 1, 0, F, N, 16 * v
        ops[v].randu_off = ops[v].randn_off = -1
    out = np.zeros((V, N, 1)); obw = np.zeros(V * 4)
    best = 1e9
    for _ in range(reps):
        eng._check(eng.lib.iifb200_product_batch(eng.ctx, V, ops, A.as_dp(a), A.as_dp(bws), None, None, None, None,
                                                 A.as_dp(out), A.as_dp(obw), None), "prod")
        best = min(best, eng.last_elapsed_ms())
    return best * 1e3

sys.path.insert(0, os.path.join(ROOT, "oracle"))
for V in (1, 148, 592, 2368):
    t = prod_time(V)
    print(f"product kernel V={V:5d} F=2: {t:9.1f} us -> {t/V*148 if V>=148 else t:8.1f} us*SM per product")
print(f"product kernel V=1 F=3: {prod_time(1,3):8.1f} us ; F=5: {prod_time(1,5):8.1f} us")
