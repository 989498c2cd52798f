"""Kernel-development probe: builds a copy of the library with -DIIF_PHASES (clock64 marks inside the
kernels), runs ONE belief through the bandwidth, convolution and product kernels and prints the cycles
per phase seen by one thread.  usage (GPU box): python profiles/phase_probe.py [N] [tid ...]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
CSRC = os.path.join(ROOT, "incrementalinference.jl_b200", "csrc")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
TIDS = [int(a) for a in sys.argv[2:]] or [0, 160]
VWIDE = int(os.environ.get("IIFB200_PROBE_V", "1"))     # beliefs per launch: > 148 probes block 0 of a WIDE launch
LOO = ["loo rows (hot loop)", "loo barrier1", "loo gather", "loo part+barrier2", "loo partsum+log+warpsum",
       "loo barrier3", "loo tot", "-"]
KER = {"bandwidth": {8: "load", 13: "search+write"},
       "conv": {8: "setup/load", 9: "measurements+recipe+labels", 10: "inflate/solve", 11: "write proposal",
                12: "bandwidth (all)", 13: "write bw/slot"},
       "product": {8: "load", 9: "ball trees", 10: "node stats + rsqrt", 20: "gibbs: samplePoint per level",
                   16: "gibbs: conditional setup", 17: "gibbs: barrier before build", 7: "gibbs: flat weight build",
                   18: "gibbs: barrier after build", 15: "gibbs: owner pick (rest)", 21: "  pick: row sums", 22: "  pick: walk to the piece",
                   23: "  pick: piece re-evaluation", 24: "  pick: candidate", 11: "gibbs: rest",
                   12: "final sample", 13: "bandwidth (all)", 14: "write"}}
from iifb200 import _abi as A, compile as CP  # noqa: E402
import parity_cases as PC  # noqa: E402

for tid in TIDS:
    lib = os.path.join(CSRC, f"libiifb200_ph{tid}.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                           "-shared", "-DIIF_PHASES", f"-DIIF_PHASE_TID={tid}", "-o", lib, os.path.join(CSRC, "iifb200.cu"), os.path.join(CSRC, "iif_plan.cpp")])
    A._lib = None
    L = A.load_library(lib)
    L.iifb200_debug_phases.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    P, xs, fs = PC.chain_problem(n=3, N=N, seed=1)
    eng = P.engine()
    R = np.random.default_rng(0)
    buf = (C.c_longlong * 32)()

    def report(name, run):
        for rep in range(3):
            L.iifb200_debug_phases(buf, 1)
            run()
            ms = eng.last_elapsed_ms()
        L.iifb200_debug_phases(buf, 0)
        v = np.array(list(buf), dtype=np.float64)
        print(f"--- {name} kernel, {VWIDE} belief(s) per launch, block 0, N={N}, tid {tid}: {ms*1e3:.1f} us (events around the launch)")
        for k, nm in KER[name].items():
            print(f"   {nm:28s} {v[k]:9.0f} cyc  {v[k]/1.965e3:6.1f} us")
        print(f"   [cost of one phase mark: {v[31] / max(v[30] and 18, 1):.0f} cycles if 18 evaluations; raw {v[30]:.0f} {v[31]:.0f}]")
        print("   inside the bandwidth search:")
        for k in range(7):
            if name == "product" and k == 7:
                continue
            print(f"      {LOO[k]:25s} {v[k]:9.0f} cyc  {v[k]/1.965e3:6.1f} us")

    V = VWIDE
    pts = R.normal(0, 1, (V, N, 1))
    Ns = np.full(V, N, dtype=np.int32); Ds = np.ones(V, dtype=np.int32); Ms = np.zeros(V, dtype=np.int32)
    out = np.zeros(4 * V)
    report("bandwidth", lambda: eng._check(L.iifb200_kde_bandwidth(eng.ctx, V, A.as_ip(Ns), A.as_ip(Ds), A.as_ip(Ms),
                                                                    A.as_dp(pts), A.as_dp(out)), "bw"))
    ops = CP.make_conv_ops([dict(factor=fs[1], sfidx=2, N=N, call_id=16 * (k + 1)) for k in range(V)])
    report("conv", lambda: eng.conv_batch(ops, V))
    a = R.normal(0, 1, (V, 2, N, 1))
    bws = np.zeros((2 * V, 4)); bws[:, 0] = 0.4
    pop = (A.ProductOp * V)()
    for v in range(V):
        pop[v].dim, pop[v].circ_mask, pop[v].nfactors, pop[v].N, pop[v].call_id = 1, 0, 2, N, 16 * (v + 1)
        pop[v].randu_off = pop[v].randn_off = -1
    o2 = np.zeros((V, N, 1)); obw = np.zeros(4 * V)
    report("product", lambda: eng._check(L.iifb200_product_batch(eng.ctx, V, pop, A.as_dp(a), A.as_dp(bws), None, None, None,
                                                                  None, A.as_dp(o2), A.as_dp(obw), None), "prod"))
    eng.close()
    os.remove(lib)
