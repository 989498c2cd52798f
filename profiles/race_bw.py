import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import parity_cases as PC
P, xs, fs = PC.chain_problem(n=2, N=8)
eng = P.engine()
R = np.random.default_rng(5)
n, d, cm = int(sys.argv[1]), int(sys.argv[2]), 0
pts = R.normal(0, 1, (n, d))
print("case", n, d, eng.kde_bandwidth(pts, cm), flush=True)
