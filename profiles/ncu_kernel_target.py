"""Target for single-kernel ncu captures: one WIDE launch of the product, convolution or bandwidth kernel
(N=100 scalar beliefs).  usage: python profiles/ncu_kernel_target.py <prod|conv|bw> [count] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import parity_cases as PC  # noqa: E402
from iifb200 import _abi as A, compile as CP  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "prod"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
N = 100
R = np.random.default_rng(0)
if what == "conv":
    P, xs, fs = PC.chain_problem(n=3, N=N, seed=1)
    eng = P.engine()
    ops = CP.make_conv_ops([dict(factor=fs[1], sfidx=2, N=N, call_id=16 * k) for k in range(V)])
    for _ in range(reps):
        eng.conv_batch(ops, V)
        print("conv", V, eng.last_elapsed_ms() * 1e3, "us")
else:
    P, xs, fs = PC.chain_problem(n=3, N=N, seed=1)
    eng = P.engine()
    if what == "bw":
        pts = R.normal(0, 1, (V, N, 1))
        Ns = np.full(V, N, dtype=np.int32); Ds = np.ones(V, dtype=np.int32); Ms = np.zeros(V, dtype=np.int32)
        out = np.zeros(V * 4)
        for _ in range(reps):
            eng._check(eng.lib.iifb200_kde_bandwidth(eng.ctx, V, A.as_ip(Ns), A.as_ip(Ds), A.as_ip(Ms), A.as_dp(pts), A.as_dp(out)), "bw")
            print("bw", V, eng.last_elapsed_ms() * 1e3, "us")
    else:
        F = 2
        a = R.normal(0, 1, (V, F, N, 1))
        bws = np.zeros((V * F, 4)); bws[:, 0] = 0.4
        ops = (A.ProductOp * V)()
        for v in range(V):
            ops[v].dim, ops[v].circ_mask, ops[v].nfactors, ops[v].N, ops[v].call_id = 1, 0, F, N, 16 * v
            ops[v].randu_off = ops[v].randn_off = -1
        out = np.zeros((V, N, 1)); obw = np.zeros(V * 4)
        for _ in range(reps):
            eng._check(eng.lib.iifb200_product_batch(eng.ctx, V, ops, A.as_dp(a), A.as_dp(bws), None, None, None, None,
                                                     A.as_dp(out), A.as_dp(obw), None), "prod")
            print("prod", V, eng.last_elapsed_ms() * 1e3, "us")
eng.close()
