"""BASELINE configs at their stated GPU counts (VERDICT r1, row g): launched one process per GPU,
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 profiles/config_probe_multi.py c4
   ... --nproc-per-node 8 ... profiles/config_probe_multi.py c5          (c2 = the 1000-pose chain, strong scaling)
Every rank runs its share of ONE tree solve (iifb200.multigpu.ShardedTreeSolver: in-graph NVLink pushes), rank 0 holds
the gathered posteriors, checks the size-independent properties the single-GPU tests check and prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import iifb200  # noqa: E402,F401
from iifb200 import compile as CP  # noqa: E402
from iifb200 import planner as PL  # noqa: E402
from iifb200 import tree as TR  # noqa: E402
from iifb200 import workloads as W  # noqa: E402
from iifb200.multigpu import ShardedTreeSolver  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "c4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
uml = len(sys.argv) > 3 and sys.argv[3] == "uml"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if what == "c4":
    n = 500
    fg = W.circular_chain(n=n, N=150, seed=42)
    order = PL.elimination_order_is(fg)
elif what == "c5":
    rows, cols = 50, 100
    fg = W.euclid2_grid(rows=rows, cols=cols, N=100, seed=42, closure_every=5)
    order = PL.elimination_order_is(fg)
else:
    n = 1000
    fg = W.scalar_chain(n, N=100, seed=42)
    order = PL.elimination_order_is(fg)
fg.solverParams.useMsgLikelihoods = uml
sv = ShardedTreeSolver(fg, order, rank, world, local, dist, gather="root")
nv = len(fg.variables)
ms = []
for it in range(steps + 2):
    sv.eng.set_solver_params(CP.solver_params_c(fg.solverParams, 100 + it))
    sv.eng.upload_arena(sv.arena)
    sv.eng.sync()
    dist.barrier()
    torch.cuda.synchronize()
    t = sv.run_timed()
    if it >= 2:
        ms.append(t)
tt = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device="cuda")
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
sv.eng.sync()
dist.barrier()
if rank == 0:
    got = CP.HostArena(sv.plan.frozen)
    sv.eng.download_arena(got)
    out = {"config": what, "n_gpus": world, "useMsgLikelihoods": uml, "variables": nv, "convolutions": sv.plan.n_conv,
           "products": sv.plan.n_prod, "messages_between_ranks": sv.n_msgs, "waves": sv.nw,
           "ms_per_solve_max_over_ranks": float(tt[0]), "conv_per_s": sv.plan.n_conv / (float(tt[0]) * 1e-3),
           "convolutions_per_rank": None}
    pts = [got.get(sv.plan.var_slot[l])[0] for l in fg.variables]
    assert all(np.isfinite(p).all() and p.shape[0] == fg.solverParams.N for p in pts)
    if what == "c4":
        err = [abs((np.arctan2(np.sin(p).mean(), np.cos(p).mean()) - k + np.pi) % (2 * np.pi) - np.pi) for k, p in enumerate(pts)]
        out["max_mean_err_rad"] = float(max(err))
        # with differential messages the posteriors are honest: sigma = 0.1 sqrt(k + 1) rad, i.e. nearly uniform on the
        # circle beyond a few dozen poses, where a circular mean says nothing — check the poses that are still localised
        chk = [(k, e) for k, e in enumerate(err) if (not uml) or 0.1 * np.sqrt(k + 1.0) < 0.6]
        out["poses_checked"] = len(chk)
        out["located"] = bool(all(e < 0.15 + (0.1 if not uml else 0.3) * np.sqrt(k + 1.0) for k, e in chk))
    elif what == "c5":
        pos = []
        for r in range(rows):
            for c in (range(cols) if r % 2 == 0 else range(cols - 1, -1, -1)):
                pos.append((float(c), float(r)))
        err = np.abs(np.stack([p.mean(axis=0) for p in pts]) - np.array(pos)).max(axis=1)
        out["mean_err"], out["max_err"] = float(err.mean()), float(err.max())
        out["located"] = bool(err.mean() < 0.4 and err.max() < 2.5)
    else:
        err = np.abs(np.array([p.mean() for p in pts]) - np.arange(nv))
        sig = 0.1 * np.sqrt(np.arange(nv) + 1.0)
        out["max_err_over_sigma"] = float((err / sig).max())
        out["located"] = bool((err < 1.0 * sig + 0.3).all())
    print(json.dumps(out))
conv = torch.tensor([float(sv.my_conv)], dtype=torch.float64, device="cuda")
allc = [torch.zeros_like(conv) for _ in range(world)]
dist.all_gather(allc, conv)
if rank == 0:
    print(json.dumps({"convolutions_per_rank": [int(c.item()) for c in allc]}))
sv.close()
dist.destroy_process_group()
