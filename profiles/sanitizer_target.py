"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): a few products (fused two-density path, generic
path with three densities, N = 150 beyond the fused limit), convolutions of every variant (lean and rare kernel: SO(3),
Nelder-Mead, sample table) and one small tree solve with lanes.  usage: compute-sanitizer --tool memcheck python profiles/sanitizer_target.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import parity_cases as PC  # noqa: E402
from iifb200 import compile as CP, solver as SV, workloads as W  # noqa: E402

P0, xs, fs = PC.chain_problem(n=2, N=8)
eng = P0.engine()
for name, kw in PC.product_cases():
    if name in ("two_gaussians_1d", "three_gaussians_1d", "two_gaussians_2d", "circular_wraparound", "partial_mask", "tiny_7"):
        PC.run_product_case(kw, eng)
        print("product", name, "ok")
eng.close()
for name, P, specs, streams in PC.conv_cases():
    if name in ("scalar_prior_and_relative", "multihypo_bimodal", "numeric_solve", "so3", "sample_table", "ragged"):
        PC.run_conv_case(P, specs, streams)
        print("conv", name, "ok")
fg = W.scalar_chain(24, N=64, seed=3)
ts = SV.TreeSolver(fg, W.chain_nd_order(24))
ts.load_from_graph(); ts.upload(); ts.run(); ts.run(); ts.download(); ts.close()
print("tree solve ok")
# boundary B3: one propagateBelief per C-ABI call (iifb200_propagate_once), serial and through a pool of contexts
from iifb200 import compile as CP  # noqa: E402
fg = W.scalar_chain(12, N=64, seed=3)
ts = SV.TreeSolver(fg, W.chain_nd_order(12))
for k in (1, 3):
    b3 = SV.B3Driver(ts.plan, ts.sp_c, contexts=k)
    ar = CP.HostArena(ts.plan.frozen)
    for l, v in fg.variables.items():
        ar.set(ts.plan.var_slot[l], v.val, v.bw, True)
    b3.run(ar)
    b3.close()
ts.close()
print("per-call path ok")
