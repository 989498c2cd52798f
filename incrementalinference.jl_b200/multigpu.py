"""Sharding one tree solve over the GPUs of a box: one process per GPU (torch.distributed for the plumbing).

Cliques are dealt to ranks in contiguous post-order runs of equal convolution weight (whole sub-trees);
the ONLY data-path exchange is the belief of a slot that an op on one rank reads and a clique on another rank owns —
the separator message of a tree edge that crosses a GPU boundary: the child's updated separator belief going up
(prepCliqueMsgUp, TreeMessageUtils.jl:667-703; with useMsgLikelihoods also the differential's measurement belief) and
the parent's belief coming down (CliqDownMessage, CliqueStateMachine.jl:672-691).  Payload = N*d doubles + bandwidths.

Messages travel INSIDE the captured CUDA graphs: every rank maps its peers' arenas through CUDA IPC
(iifb200_ipc_export / iifb200_ipc_attach); the sender's graph holds a PUSH node (copy kernel writing the slot into the
peer's arena over NVLink, then raising a flag there) right behind the kernels that produced the belief, the receiver's
graph a WAIT node in front of the kernels that read it.  A rank's whole pass is ONE graph launch — no host call, no
NCCL call and no graph split per message (round 1 issued `batch_isend_irecv` from Python between graph segments).

The partitioning logic (ownership, messages, per-rank schedules) is pure host code and is covered by world_size 2 / 4
`gloo` tests on CPU (tests/test_multigpu_gloo.py), which replay the very same PUSH / WAIT ops as sends and receives.
"""
import ctypes as C
from collections import defaultdict
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _abi as A
from . import compile as CP
from . import graph as G
from . import tree as TR


# ------------------------------------------------------------------------------------------ ownership
def clique_owner(fg: G.FactorGraph, tree: TR.BayesTree, world: int) -> Dict[int, int]:
    """round-1 rule (kept for comparison): contiguous ranges of the variable index of the first frontal"""
    nv = len(fg.variables)
    return {c.id: min(fg.variables[c.frontals[0]].index * world // nv, world - 1) for c in tree.cliques}


def clique_owner_balanced(plan: TR.SolvePlan, tree: TR.BayesTree, world: int) -> Dict[int, int]:
    """Balanced sub-tree partition: cliques in post-order (every sub-tree is a contiguous run), cut into `world` runs
    of equal convolution weight (sum over the clique's ops of factors + 1, the product)."""
    w = [0.0] * len(tree.cliques)
    for (k, a, _), c in zip(plan.sched_waved, plan.op_clique):
        if c >= 0 and k == A.S_PROPAGATE:
            w[c] += len(plan.props[a]["factors"]) + 1.0
        elif c >= 0 and k == A.S_DECONV:
            w[c] += 1.0
    post = TR._postorder(tree)
    total = sum(w) or 1.0
    owner, acc = {}, 0.0
    for cid in post:
        owner[cid] = min(int(world * (acc + 0.5 * w[cid]) / total), world - 1)
        acc += w[cid]
    return owner


# ------------------------------------------------------------------------------------------ messages
def partition_plan(plan: TR.SolvePlan, owner: Dict[int, int], world: int):
    """-> (rank_of_op, transfers) where transfers = sorted list of (before_wave, slot, src, dst): one entry per belief
    version that an op on rank dst reads from a slot whose home clique lives on rank src != dst.  Main-graph slots are
    replicated on every rank (read-only until the final write-back, which is rank-local)."""
    op_rank = [owner[c] for c in plan.op_clique]
    last_write: Dict[int, int] = {}
    seen = set()
    transfers: List[Tuple[int, int, int, int]] = []
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
    ops, so barrier waves and hazards are those this device sees."""
    idx = [i for i, r in enumerate(op_rank) if r == rank]
    wt = [float(len(plan.props[plan.sched_waved[i][1]]["factors"]) + 1) if plan.sched_waved[i][0] == A.S_PROPAGATE else 0.05
          for i in idx]
    return TR.assign_lanes(tree, [plan.op_clique[i] for i in idx], wt, [plan.op_wave[i] for i in idx],
                           [plan.op_reads[i] for i in idx], [plan.op_writes[i] for i in idx], nlanes)


def dist_schedule(plan: TR.SolvePlan, tree: TR.BayesTree, owner: Dict[int, int], world: int, rank: int, nlanes: int = 4,
                  gather: Optional[str] = "root"):
    """One rank's pass as a single wave schedule with in-graph messages.

    Every belief version that crosses ranks gets a message id (the same numbering on every rank).  The sender's PUSH op
    sits in the wave and lane of the op that produced the version (it runs behind that segment's kernels), the
    receiver's WAIT ops in the wave and lane of every op that reads it (in front of that segment's kernels).
    `gather`: "root" appends a last wave in which every rank pushes the posteriors of the variables it owns to rank 0
    (the solution then sits in rank 0's main-graph slots), "all" to every rank, None: nothing.

    -> dict(ops, lanes, wave_off, xfers=[(slot, peer, msg)], nflags, my_conv, msgs=[(slot, src, dst, producer_wave)])"""
    n = len(plan.sched_waved)
    op_rank = [owner[c] for c in plan.op_clique]
    last_writer: Dict[int, int] = {}
    msg_of: Dict[tuple, int] = {}
    msgs: List[tuple] = []                      # (slot, src, dst, producer op index)
    waits: Dict[int, List[int]] = defaultdict(list)   # reader op index -> message ids
    for i in range(n):
        b = op_rank[i]
        for s in plan.op_reads[i]:
            if s not in plan.slot_clique:
                continue
            a = owner[plan.slot_clique[s]]
            if a == b:
                continue
            p = last_writer.get(s, -1)
            if p < 0 or op_rank[p] != a:
                raise A.IIFB200Error(f"dist_schedule: slot {s} read on rank {b} before its home rank {a} wrote it")
            key = (s, b, p)
            if key not in msg_of:
                msg_of[key] = len(msgs)
                msgs.append((s, a, b, p))
            if msg_of[key] not in waits[i]:
                waits[i].append(msg_of[key])
        for s in plan.op_writes[i]:
            last_writer[s] = i
    dup = defaultdict(int)
    for (s, a, b, p) in msgs:
        dup[(s, b)] += 1
    if any(v > 1 for v in dup.values()):
        # the receive buffer of a slot on a rank is the slot's replica there: a second version could overwrite the
        # first while it is still being read.  Does not occur for up / down separator messages (one version each).
        raise A.IIFB200Error("dist_schedule: a slot is sent twice to the same rank in one pass")
    mine = [i for i in range(n) if op_rank[i] == rank]
    lanes_own = rank_lanes(plan, tree, op_rank, rank, nlanes)
    lane_of = dict(zip(mine, lanes_own))
    nw = len(plan.wave_off) - 1
    xfers: List[tuple] = []
    per_wave: Dict[int, list] = defaultdict(list)       # wave -> [(kind, a, b, lane)]
    for i in mine:
        k, a, b = plan.sched_waved[i]
        per_wave[plan.op_wave[i]].append((k, a, b, lane_of[i]))
        seen_here = set()
        for m in waits.get(i, ()):
            if (m, lane_of[i]) in seen_here:
                continue
            seen_here.add((m, lane_of[i]))
            xfers.append((msgs[m][0], msgs[m][1], m))
            per_wave[plan.op_wave[i]].append((A.S_WAIT, len(xfers) - 1, 0, lane_of[i]))
    for m, (s, a, b, p) in enumerate(msgs):
        if a == rank:
            xfers.append((s, b, m))
            per_wave[plan.op_wave[p]].append((A.S_PUSH, len(xfers) - 1, 0, lane_of[p]))
    nflags = len(msgs)
    if gather:
        dsts = [0] if gather == "root" else list(range(world))
        for c in tree.cliques:
            for v in c.frontals:
                s = plan.var_slot[v]
                for dst in dsts:
                    if dst == owner[c.id]:
                        continue
                    if owner[c.id] == rank:
                        xfers.append((s, dst, nflags))
                        per_wave[nw].append((A.S_PUSH, len(xfers) - 1, 0, 0))
                    if dst == rank:
                        xfers.append((s, owner[c.id], nflags))
                        per_wave[nw].append((A.S_WAIT, len(xfers) - 1, 0, 0))
                    nflags += 1
        nw += 1
    ops, lanes, wave_off = [], [], [0]
    for w in range(nw):
        for (k, a, b, ln) in per_wave.get(w, ()):
            ops.append((k, a, b))
            lanes.append(ln)
        wave_off.append(len(ops))
    my_conv = sum(len(plan.props[a]["factors"]) for k, a, _ in ops if k == A.S_PROPAGATE)
    return dict(ops=ops, lanes=lanes, wave_off=wave_off, xfers=xfers, nflags=max(nflags, 1), my_conv=my_conv, msgs=msgs,
                op_rank=op_rank)


def make_xfer_ops(xfers):
    arr = (A.XferOp * max(len(xfers), 1))()
    for i, (s, peer, m) in enumerate(xfers):
        arr[i].slot, arr[i].peer, arr[i].msg = s, peer, m
    return arr


# ------------------------------------------------------------------------------------------ device solver
class ShardedTreeSolver:
    """One rank's share of a tree solve.  `dist` is torch.distributed (used for the handle exchange and barriers only)."""

    def __init__(self, fg, order, rank, world, local_rank, dist, gather: Optional[str] = "root", balanced: bool = True,
                 lanes: Optional[int] = None):
        import os

        import torch
        from .engine import Engine
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.fg = fg
        self.tree = TR.buildTree(fg, list(order))
        self.plan = TR.compile_solve(fg, self.tree)        # SolverParams.useMsgLikelihoods is honoured (differentials travel too)
        self.owner = clique_owner_balanced(self.plan, self.tree, world) if balanced else clique_owner(fg, self.tree, world)
        nl = int(os.environ.get("IIFB200_LANES", "4")) if lanes is None else lanes
        self.sched = dist_schedule(self.plan, self.tree, self.owner, world, rank, nl, gather)
        self.lanes = self.sched["lanes"]
        self.my_conv = self.sched["my_conv"]
        self.n_msgs = len(self.sched["msgs"])
        fz = self.plan.frozen
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.sp_c = CP.solver_params_c(fg.solverParams)
        self.eng = Engine(fz, self.sp_c, local_rank)        # library-owned arena: exported to the peers through CUDA IPC
        self.eng.set_stream(self.stream.cuda_stream)
        ha, hf = self.eng.ipc_export(self.sched["nflags"])
        allh = [None] * world
        dist.all_gather_object(allh, (ha, hf))
        self.eng.ipc_attach(world, rank, b"".join(h[0] for h in allh), b"".join(h[1] for h in allh))
        self.props_c = CP.make_prop_ops(self.plan.props)
        self.sched_c = CP.make_sched_ops(self.sched["ops"], self.lanes)
        self.deconvs_c = CP.make_deconv_ops(self.plan.deconvs or [])
        self.xfers_c = make_xfer_ops(self.sched["xfers"])
        self.sid = self.eng.schedule_build_dist(self.sched["wave_off"], self.sched_c, len(self.sched["ops"]), self.props_c,
                                                len(self.plan.props), self.deconvs_c, len(self.plan.deconvs or []),
                                                self.xfers_c, len(self.sched["xfers"]))
        self.arena = CP.HostArena(fz)
        # every rank knows the size of every clique-local belief (remote replicas are receive buffers)
        N = fg.solverParams.N
        self.arena.npts[len(fg.variables):] = N
        self.arena.flags[len(fg.variables):] = 1
        self.load_from_graph()
        self.eng.upload_arena(self.arena)
        self.eng.sync()
        dist.barrier()
        self.nw = len(self.sched["wave_off"]) - 1
        self.ts = self          # TreeSolver-compatible facade for bench.py

    def load_from_graph(self):
        for l, v in self.fg.variables.items():
            self.arena.set(self.plan.var_slot[l], v.val, v.bw, v.initialized, v.infoPerCoord)

    def run(self):
        """the rank's whole pass: ONE graph launch.  Callers separate consecutive passes by a barrier (a push of pass
        k+1 must not land in a replica that pass k still reads)."""
        with self.torch.cuda.stream(self.stream):
            self.eng.schedule_run(self.sid)

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
        self.eng.sync()
        self.dist.barrier()     # nobody unmaps memory a peer may still write
        self.eng.close()
