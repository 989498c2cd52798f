"""World-size-2 `gloo` test (CPU) of the multi-GPU sharding logic (iifb200.multigpu): clique
ownership, the separator-message transfer list and the per-rank wave schedules.  Each rank runs
only its own cliques (through the CPU oracle standing in for the kernels) and exchanges exactly the
listed separator beliefs; the merged result must equal the single-process solve bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, ret, uml=False, kind="chain"):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import iifb200  # noqa: F401
    import oracle as O
    from iifb200 import _abi as A
    from iifb200 import compile as CP
    from iifb200 import multigpu as MG
    from iifb200 import tree as TR
    from iifb200 import workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if kind == "chain":
        fg = W.scalar_chain(n, N=32, seed=7)
        order = W.chain_nd_order(n)
    else:
        fg = W.euclid2_grid(rows=4, cols=n // 4, N=24, seed=7, closure_every=2)
        order = TR.getEliminationOrder(fg, "nd")
    fg.solverParams.useMsgLikelihoods = uml
    tree = TR.buildTree(fg, order)
    plan = TR.compile_solve(fg, tree)
    owner = MG.clique_owner_balanced(plan, tree, world)
    sched = MG.dist_schedule(plan, tree, owner, world, rank, nlanes=4, gather="root")
    arena = CP.HostArena(plan.frozen)
    for l, v in fg.variables.items():
        arena.set(plan.var_slot[l], v.val, v.bw, True)
    nv = len(fg.variables)
    arena.npts[nv:] = fg.solverParams.N
    arena.flags[nv:] = 1
    sp = CP.solver_params_c(fg.solverParams)
    orc = O.Oracle(plan.frozen, arena, sp)
    props, deconvs = CP.make_prop_ops(plan.props), CP.make_deconv_ops(plan.deconvs or [])
    slots = plan.frozen["slots"]
    xf = sched["xfers"]

    def view(s):      # one message = the slot's points, bandwidths, ipc, point count and flags (what iif_push_kernel writes)
        sd = slots[s]
        return [arena.pts[sd.pts_off:sd.pts_off + sd.cap * sd.dim], arena.bw[s * 4:(s + 1) * 4], arena.ipc[s * 4:(s + 1) * 4]]

    pending, received = [], set()
    nw = len(sched["wave_off"]) - 1
    for w in range(nw):
        ops_w = sched["ops"][sched["wave_off"][w]:sched["wave_off"][w + 1]]
        for (k, a, _) in ops_w:                       # WAIT nodes run in front of the segment's kernels
            if k == A.S_WAIT:
                s, src, m = xf[a]
                if m in received:                     # flags are level-triggered: later waits on a raised flag pass
                    continue
                received.add(m)
                parts = view(s)
                buf = torch.zeros(sum(len(x) for x in parts) + 2, dtype=torch.float64)
                dist.recv(buf, src, tag=m)
                o = 0
                for x in parts:
                    x[:] = buf[o:o + len(x)].numpy()
                    o += len(x)
                arena.npts[s], arena.flags[s] = int(buf[o]), int(buf[o + 1])
        comp = [(k, a, b) for (k, a, b) in ops_w if k in (A.S_PROPAGATE, A.S_COPY, A.S_DECONV)]
        if comp:
            orc.schedule_run([0, len(comp)], CP.make_sched_ops(comp), props, deconvs=deconvs)
        for (k, a, _) in ops_w:                       # PUSH nodes run behind them
            if k == A.S_PUSH:
                s, dst, m = xf[a]
                parts = view(s)
                buf = torch.from_numpy(np.concatenate(parts + [np.array([arena.npts[s], arena.flags[s]], dtype=np.float64)]))
                pending.append(dist.isend(buf, dst, tag=m))
    for h in pending:
        h.wait()
    if rank == 0:
        # rank 0 holds the whole solution after the gather wave; single-process reference on the same plan
        ar1 = CP.HostArena(plan.frozen)
        for l, v in fg.variables.items():
            ar1.set(plan.var_slot[l], v.val, v.bw, True)
        o1 = O.Oracle(plan.frozen, ar1, sp)
        o1.schedule_run(plan.wave_off, CP.make_sched_ops(plan.sched_waved), props, deconvs=deconvs)
        ok = all(np.array_equal(arena.get(plan.var_slot[l])[0], ar1.get(plan.var_slot[l])[0]) and
                 np.array_equal(arena.get(plan.var_slot[l])[1], ar1.get(plan.var_slot[l])[1]) for l in fg.variables)
        ret["ok"] = bool(ok)
        ret["nmsgs"] = len(sched["msgs"])
        ret["nvars"] = nv
        ret["split"] = [sum(1 for r in sched["op_rank"] if r == k) for k in range(world)]
        ret["ndeconv"] = len(plan.deconvs or [])
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
        assert 0 < ret["nmsgs"] <= 16           # only the cut's separator messages cross ranks
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


@pytest.mark.parametrize("uml,kind,n", [(True, "chain", 40), (False, "grid", 32), (True, "grid", 32)])
def test_sharded_solve_with_differential_messages_and_grid(built, uml, kind, n):
    """useMsgLikelihoods = true sharded (the differentials' measurement beliefs travel too) and a grid whose separators
    hold several variables: rank 0's gathered solution equals the single-process solve bit for bit"""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + 13 * n + 3 * uml) % 1000
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_worker, args=(2, port, n, ret, uml, kind), nprocs=2, join=True)
        assert ret["ok"], "sharded solve differs from the single-process solve"
        assert ret["nmsgs"] > 0 and (ret["ndeconv"] > 0) == uml


def test_dist_schedule_properties():
    """messages: numbered alike on every rank, pushed from the producer's wave, awaited in every reader's wave"""
    import iifb200  # noqa: F401
    from iifb200 import _abi as A
    from iifb200 import multigpu as MG
    from iifb200 import tree as TR
    from iifb200 import workloads as W
    fg = W.scalar_chain(96, N=16)
    tree = TR.buildTree(fg, W.chain_nd_order(96))
    plan = TR.compile_solve(fg, tree)
    for world in (2, 4, 8):
        owner = MG.clique_owner_balanced(plan, tree, world)
        load = [0.0] * world
        for (k, a, _), c in zip(plan.sched_waved, plan.op_clique):
            if k == A.S_PROPAGATE:
                load[owner[c]] += len(plan.props[a]["factors"]) + 1
        assert min(load) > 0.6 * max(load), load                       # balanced by convolution weight
        sch = [MG.dist_schedule(plan, tree, owner, world, r, 4, "root") for r in range(world)]
        assert all(s["msgs"] == sch[0]["msgs"] and s["nflags"] == sch[0]["nflags"] for s in sch)
        pushes, waits = {}, {}
        for r, s in enumerate(sch):
            assert s["wave_off"][-1] == len(s["ops"]) == len(s["lanes"])
            for w in range(len(s["wave_off"]) - 1):
                for (k, a, _) in s["ops"][s["wave_off"][w]:s["wave_off"][w + 1]]:
                    if k == A.S_PUSH:
                        slot, dst, m = s["xfers"][a]
                        assert m not in pushes and dst != r
                        pushes[m] = (r, dst, w)
                    elif k == A.S_WAIT:
                        slot, src, m = s["xfers"][a]
                        waits.setdefault(m, []).append((r, src, w))
        assert set(pushes) == set(waits) == set(range(sch[0]["nflags"]))
        for m, (src, dst, wp) in pushes.items():
            for (r, s_, ww) in waits[m]:
                assert r == dst and s_ == src and wp <= ww           # produced no later than it is awaited
        assert sum(s["my_conv"] for s in sch) == plan.n_conv


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
