"""World-size-2 `gloo` test (CPU) of the multi-GPU sharding logic (iifb200.multigpu): clique
ownership, the separator-message transfer list and the per-rank wave schedules.  Each rank runs
only its own cliques (through the CPU oracle standing in for the kernels) and exchanges exactly the
listed separator beliefs; the merged result must equal the single-process solve bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import iifb200  # noqa: F401
    import oracle as O
    from iifb200 import compile as CP
    from iifb200 import multigpu as MG
    from iifb200 import tree as TR
    from iifb200 import workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fg = W.scalar_chain(n, N=32, seed=7)
    tree = TR.buildTree(fg, W.chain_nd_order(n))
    plan = TR.compile_solve(fg, tree)
    owner = MG.clique_owner(fg, tree, world)
    op_rank, transfers = MG.partition_plan(plan, owner, world)
    my_ops, my_wave_off = MG.rank_schedule(plan, op_rank, rank)
    arena = CP.HostArena(plan.frozen)
    for l, v in fg.variables.items():
        arena.set(plan.var_slot[l], v.val, v.bw, True)
    nv = len(fg.variables)
    arena.npts[nv:] = 32
    arena.flags[nv:] = 1
    sp = CP.solver_params_c(fg.solverParams)
    orc = O.Oracle(plan.frozen, arena, sp)
    props, ops = CP.make_prop_ops(plan.props), CP.make_sched_ops(my_ops)
    nw = len(plan.wave_off) - 1
    comm = sorted({t[0] for t in transfers if rank in (t[2], t[3])})     # as ShardedTreeSolver: own exchange waves only
    slots = plan.frozen["slots"]

    def exchange(w):
        for (ww, s, a, b) in transfers:
            if ww != w or rank not in (a, b):
                continue
            sd = slots[s]
            pts = torch.from_numpy(arena.pts[sd.pts_off:sd.pts_off + sd.cap * sd.dim])
            bw = torch.from_numpy(arena.bw[s * 4:(s + 1) * 4])
            for t in (pts, bw):
                if rank == a:
                    dist.send(t, b)
                else:
                    dist.recv(t, a)

    prev = 0
    for w in comm:
        if w > prev:
            orc.schedule_run(my_wave_off, ops, props, prev, w)
        exchange(w)
        prev = w
    orc.schedule_run(my_wave_off, ops, props, prev, nw)
    # gather the posteriors of the variables whose frontal clique this rank owns
    mine = {l: arena.get(plan.var_slot[l])[0] for l in fg.variables if owner[tree.frontal_of[l]] == rank}
    out = [None] * world
    dist.all_gather_object(out, mine)
    if rank == 0:
        merged = {}
        for d in out:
            merged.update(d)
        # single-process reference on the same plan
        ar1 = CP.HostArena(plan.frozen)
        for l, v in fg.variables.items():
            ar1.set(plan.var_slot[l], v.val, v.bw, True)
        o1 = O.Oracle(plan.frozen, ar1, sp)
        o1.schedule_run(plan.wave_off, CP.make_sched_ops(plan.sched_waved), props)
        ok = all(np.array_equal(merged[l], ar1.get(plan.var_slot[l])[0]) for l in fg.variables)
        ret["ok"] = bool(ok)
        ret["ntransfers"] = len(transfers)
        ret["nvars"] = len(merged)
        ret["split"] = [sum(1 for r in op_rank if r == k) for k in range(world)]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [24, 41])
def test_sharded_solve_matches_single_process(built, n):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + n) % 1000
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
        assert ret["ok"], "sharded solve differs from the single-process solve"
        assert ret["nvars"] == n
        assert 0 < ret["ntransfers"] <= 16      # only the cut's separator messages cross ranks
        assert min(ret["split"]) > 0.3 * max(ret["split"])   # both ranks carry a real share


def test_sharded_solve_world4_matches_single_process(built):
    """four ranks: some exchange waves involve only two of them, the others must not wait there"""
    import torch.multiprocessing as mp
    n = 96
    port = 29500 + (os.getpid() + 7 * n) % 1000
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_worker, args=(4, port, n, ret), nprocs=4, join=True)
        assert ret["ok"], "sharded solve differs from the single-process solve"
        assert ret["nvars"] == n and min(ret["split"]) > 0


def test_partition_properties():
    import iifb200  # noqa: F401
    from iifb200 import multigpu as MG
    from iifb200 import tree as TR
    from iifb200 import workloads as W
    fg = W.scalar_chain(64, N=16)
    tree = TR.buildTree(fg, W.chain_nd_order(64))
    plan = TR.compile_solve(fg, tree)
    for world in (2, 4, 8):
        owner = MG.clique_owner(fg, tree, world)
        op_rank, transfers = MG.partition_plan(plan, owner, world)
        assert set(op_rank) == set(range(world))
        # every transfer crosses ranks, moves a clique-local slot, and precedes its reading wave
        for (w, s, a, b) in transfers:
            assert a != b and s in plan.slot_clique and owner[plan.slot_clique[s]] == a and 0 < w
        # per-rank schedules partition the global one
        tot = 0
        for r in range(world):
            ops, woff = MG.rank_schedule(plan, op_rank, r)
            assert woff[-1] == len(ops) and len(woff) == len(plan.wave_off)
            tot += len(ops)
        assert tot == len(plan.sched_waved)


def test_rank_lanes_are_hazard_free():
    """Per-rank lanes (multigpu.rank_lanes): among one rank's ops, conflicting ops of different lanes are separated
    by one of THAT rank's barrier waves; beliefs received from other ranks land between graph launches."""
    import iifb200  # noqa: F401
    from iifb200 import multigpu as MG
    from iifb200 import tree as TR
    from iifb200 import workloads as W
    fg = W.scalar_chain(128, N=16)
    tree = TR.buildTree(fg, W.chain_nd_order(128))
    plan = TR.compile_solve(fg, tree, useMsgLikelihoods=False)
    for world in (2, 4):
        owner = MG.clique_owner(fg, tree, world)
        op_rank, transfers = MG.partition_plan(plan, owner, world)
        for r in range(world):
            idx = [i for i, q in enumerate(op_rank) if q == r]
            ln = MG.rank_lanes(plan, tree, op_rank, r, 4)
            assert len(ln) == len(idx) and max(ln) > 0
            wv = [plan.op_wave[i] for i in idx]
            barrier = sorted({w for l, w in zip(ln, wv) if l == 0})
            touched = {}
            for k, i in enumerate(idx):
                rd, wr = plan.op_reads[i], plan.op_writes[i]
                for s_ in set(rd) | set(wr):
                    for j, j_writes in touched.get(s_, []):
                        if (j_writes or s_ in wr) and ln[k] != ln[j] and ln[k] != 0 and ln[j] != 0:
                            assert any(wv[j] <= b <= wv[k] for b in barrier), (r, k, j, s_)
                    touched.setdefault(s_, []).append((k, s_ in wr))
