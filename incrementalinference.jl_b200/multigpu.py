"""Sharding one tree solve over the GPUs of a box: one process per GPU (torch.distributed).

Cliques run on the rank that owns their (first) frontal variable — contiguous variable ranges,
i.e. contiguous chain segments — and the ONLY data-path exchange is the separator message of a
tree edge whose two cliques live on different ranks: the child's updated
separator belief going up (prepCliqueMsgUp, TreeMessageUtils.jl:667-703) and the parent's
belief coming down (CliqDownMessage, CliqueStateMachine.jl:672-691).  Payload per message =
N*d doubles of points + IIF_MAX_DIM bandwidths.  Messages are sent with NCCL point-to-point ops
(`batch_isend_irecv`) straight out of / into the device arena, on the same CUDA stream the
kernels are launched on, so no host synchronisation separates compute from exchange.

The partitioning logic (ownership, transfer list, per-rank wave schedule) is pure host code and
is covered by world_size-2 `gloo` tests on CPU (tests/test_multigpu_gloo.py).
"""
from collections import defaultdict
from typing import Dict, List, Tuple

import numpy as np

from . import _abi as A
from . import compile as CP
from . import graph as G
from . import tree as TR


def clique_owner(fg: G.FactorGraph, tree: TR.BayesTree, world: int) -> Dict[int, int]:
    """rank of every clique: contiguous ranges of the variable index of its first frontal"""
    nv = len(fg.variables)
    return {c.id: min(fg.variables[c.frontals[0]].index * world // nv, world - 1) for c in tree.cliques}


def partition_plan(plan: TR.SolvePlan, owner: Dict[int, int], world: int):
    """-> (rank_of_op, transfers) where transfers = sorted list of (before_wave, slot, src, dst).

    A transfer is needed whenever an op on rank b reads a clique-local slot whose home clique lives on
    rank a != b.  It is placed immediately before the reading wave; levelisation guarantees the slot's
    last write (on a) happened in an earlier wave.  Main-graph slots are replicated on every rank (they
    are read-only until the final write-back, which is rank-local)."""
    op_rank = [owner[c] for c in plan.op_clique]
    last_write: Dict[int, int] = {}
    seen = set()
    transfers: List[Tuple[int, int, int, int]] = []
    # ops are sorted by wave; process wave by wave so that "last write before the read" is well defined
    for i, (w, rd, wr) in enumerate(zip(plan.op_wave, plan.op_reads, plan.op_writes)):
        b = op_rank[i]
        for s in rd:
            if s not in plan.slot_clique:
                continue
            a = owner[plan.slot_clique[s]]
            if a != b:
                key = (s, b, last_write.get(s, -1))
                if key not in seen:
                    seen.add(key)
                    transfers.append((w, s, a, b))
        for s in wr:
            last_write[s] = w
    transfers.sort()
    return op_rank, transfers


def rank_schedule(plan: TR.SolvePlan, op_rank: List[int], rank: int):
    """the ops of `rank`, with the plan's GLOBAL wave indices (other ranks' ops leave empty waves)"""
    nw = len(plan.wave_off) - 1
    ops = [(o, w) for o, w, r in zip(plan.sched_waved, plan.op_wave, op_rank) if r == rank]
    wave_off = [0] * (nw + 1)
    for _, w in ops:
        wave_off[w + 1] += 1
    for w in range(nw):
        wave_off[w + 1] += wave_off[w]
    return [o for o, _ in ops], wave_off


def rank_lanes(plan: TR.SolvePlan, tree: TR.BayesTree, op_rank: List[int], rank: int, nlanes: int) -> List[int]:
    """Lane of every op of `rank` (same order as rank_schedule): tree.assign_lanes restricted to the rank's own
    ops, so barrier waves and hazards are those this device sees.  Beliefs arriving from other ranks are written
    between two graph launches (wave-range boundaries), where every lane has joined."""
    idx = [i for i, r in enumerate(op_rank) if r == rank]
    wt = [float(len(plan.props[plan.sched_waved[i][1]]["factors"]) + 1) if plan.sched_waved[i][0] == A.S_PROPAGATE else 0.05
          for i in idx]
    return TR.assign_lanes(tree, [plan.op_clique[i] for i in idx], wt, [plan.op_wave[i] for i in idx],
                           [plan.op_reads[i] for i in idx], [plan.op_writes[i] for i in idx], nlanes)


class ShardedTreeSolver:
    """One rank's share of a tree solve.  `dist` is torch.distributed (nccl on GPUs)."""

    def __init__(self, fg, order, rank, world, local_rank, dist, device_engine=True):
        import torch
        from .engine import Engine
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.fg = fg
        self.tree = TR.buildTree(fg, list(order))
        self.plan = TR.compile_solve(fg, self.tree, useMsgLikelihoods=False)   # sharded solves exchange plain separator beliefs
        self.owner = clique_owner(fg, self.tree, world)
        self.op_rank, self.transfers = partition_plan(self.plan, self.owner, world)
        my_ops, my_wave_off = rank_schedule(self.plan, self.op_rank, rank)
        self.my_conv = sum(len(self.plan.props[a]["factors"]) for k, a, _ in my_ops if k == A.S_PROPAGATE)
        # this rank splits its graph only at the waves where IT sends or receives (point-to-point exchanges involve
        # nobody else); waves where other pairs exchange do not interrupt its graph
        self.comm_waves = sorted({t[0] for t in self.transfers if rank in (t[2], t[3])})
        self.by_wave = defaultdict(list)
        for w, s, a, b in self.transfers:
            if rank in (a, b):
                self.by_wave[w].append((s, a, b))
        # device arena owned by torch so NCCL can address slots directly
        fz = self.plan.frozen
        lib = A.load_library()
        nbytes = int(lib.iifb200_arena_bytes(fz["nslots"], fz["slots"]))
        self.dev = torch.device("cuda", local_rank)
        self.arena_t = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.sp_c = CP.solver_params_c(fg.solverParams)
        self.eng = Engine(fz, self.sp_c, local_rank, self.arena_t.data_ptr())
        self.eng.set_stream(self.stream.cuda_stream)
        self.props_c = CP.make_prop_ops(self.plan.props)
        import os
        self.lanes = rank_lanes(self.plan, self.tree, self.op_rank, rank, int(os.environ.get("IIFB200_LANES", "4")))
        self.sched_c = CP.make_sched_ops(my_ops, self.lanes)
        self.sid = self.eng.schedule_build(my_wave_off, self.sched_c, len(my_ops), self.props_c, len(self.plan.props))
        self.arena = CP.HostArena(fz)
        # every rank knows the size of every clique-local belief (remote replicas are receive buffers)
        N = fg.solverParams.N
        self.arena.npts[len(fg.variables):] = N
        self.arena.flags[len(fg.variables):] = 1
        self.eng.upload_arena(self.arena)
        self.nw = len(self.plan.wave_off) - 1
        self.total = fz["total_doubles"]
        # TreeSolver-compatible facade for bench.py
        self.ts = self

    # facade
    def load_from_graph(self):
        for l, v in self.fg.variables.items():
            self.arena.set(self.plan.var_slot[l], v.val, v.bw, v.initialized, v.infoPerCoord)

    def _slot_views(self, s):
        sd = self.plan.frozen["slots"][s]
        pts = self.arena_t[sd.pts_off:sd.pts_off + sd.cap * sd.dim]
        bw = self.arena_t[self.total + s * A.IIF_MAX_DIM:self.total + (s + 1) * A.IIF_MAX_DIM]
        return pts, bw

    def _exchange(self, w):
        items = self.by_wave.get(w)
        if not items:
            return
        dist, ops = self.dist, []
        for s, a, b in items:
            pts, bw = self._slot_views(s)
            for t in (pts, bw):
                if self.rank == a:
                    ops.append(dist.P2POp(dist.isend, t, b))
                else:
                    ops.append(dist.P2POp(dist.irecv, t, a))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def run(self):
        """all waves, with the separator-message exchanges at their wave boundaries"""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            prev = 0
            for w in self.comm_waves:
                if w > prev:
                    self.eng.schedule_run(self.sid, prev, w)
                self._exchange(w)
                prev = w
            if prev < self.nw:
                self.eng.schedule_run(self.sid, prev, self.nw)

    def run_timed(self):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        self.run()
        e1.record(self.stream)
        e1.synchronize()
        return e0.elapsed_time(e1)

    def profile(self):
        return self.eng.schedule_profile(self.sid)

    def close(self):
        self.eng.close()
