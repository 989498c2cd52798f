"""Kernel-development probe: builds a copy of the library with -DIIF_PHASES (clock64 marks inside the
leave-one-out objective), runs single-belief bandwidth searches and prints cycles per phase and per
objective evaluation for an on-critical-path thread (tid 0) and an off-path one.
usage (GPU box): python profiles/phase_probe.py [N] [threads-independent]"""
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
NAMES = ["rows(hot loop)", "barrier1", "gather", "part+barrier2", "partsum+log+warpsum", "barrier3", "tot",
         "-", "-", "-"]
for tid in (0, 160, 480):
    lib = os.path.join(CSRC, f"libiifb200_ph{tid}.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                           "-shared", "-DIIF_PHASES", f"-DIIF_PHASE_TID={tid}", "-o", lib, os.path.join(CSRC, "iifb200.cu")])
    import importlib
    from iifb200 import _abi as A
    A._lib = None
    L = A.load_library(lib)
    L.iifb200_debug_phases.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    import parity_cases as PC
    P, xs, fs = PC.chain_problem(n=3, N=100, seed=1)
    eng = P.engine()
    R = np.random.default_rng(0)
    pts = R.normal(0, 1, (1, N, 1))
    Ns = np.full(1, N, dtype=np.int32); Ds = np.ones(1, dtype=np.int32); Ms = np.zeros(1, dtype=np.int32)
    out = np.zeros(4)
    buf = (C.c_longlong * 16)()
    for rep in range(3):
        L.iifb200_debug_phases(buf, 1)
        eng._check(L.iifb200_kde_bandwidth(eng.ctx, 1, A.as_ip(Ns), A.as_ip(Ds), A.as_ip(Ms), A.as_dp(pts), A.as_dp(out)), "bw")
        ms = eng.last_elapsed_ms()
    L.iifb200_debug_phases(buf, 0)
    v = np.array(list(buf), dtype=np.float64)
    print(f"tid {tid}: kernel {ms*1e3:.1f} us, bw {out[0]:.6f}; cycles per phase (total over the search): ")
    for k in range(7):
        print(f"   {NAMES[k]:22s} {v[k]:10.0f}")
    print(f"   sum {v[:7].sum():.0f} cycles = {v[:7].sum()/1.965e3:.1f} us at 1965 MHz")
    eng.close()
    os.remove(lib)
