"""One process per GPU over torch.distributed (NCCL over NVLink/NVSwitch on the B200 box, gloo in CPU tests).

Two multi-GPU modes, both replacing the reference's single-process thread-per-GPU ``parallel.py``
(DataParallelModel / DataParallelCriterion2):

* ``ReplicaTrainer`` -- the reference's own ``--parallel True`` semantics (train_eval_Gowalla.py:97-104,137;
  SURVEY.md 5.8): every GPU runs the full-graph propagation with its own dropout stream and scores a
  different ``batch_size`` slice of the train rows, so a step consumes ``batch_size * world`` rows; the
  gradient is the SUM of the per-GPU mean-loss gradients (``loss.backward(ones(ndev))``).  Instead of
  broadcasting 18 MB of parameters before every forward and reduce-adding gradients to GPU 0
  (parallel.py:19,127-130), the replicas stay bit-identical: one NCCL all-reduce of the flat 18 MB gradient
  buffer per step and the same Adam update everywhere.  Weak scaling.
* ``shard_eval_users`` / ``allreduce_sums`` -- AllNeg evaluation shards the test users over the ranks;
  only 16 metric sums (and optionally the top-20 lists) are merged.

The host-side pieces (row slicing, user sharding, flat-gradient all-reduce, metric merge) are plain
functions so that the world_size-2 gloo tests in tests/test_dist_cpu.py exercise them without a GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .train import FusedTrainer


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def rank_rows(cursor: int, batch: int, rank: int, world: int, n_rows: int):
    """Train rows [lo,hi) of `rank` for the step starting at global `cursor` (a step consumes batch*world rows,
    train_eval_Gowalla.py:104).  Rows past n_rows are clipped (tail step)."""
    lo = min(n_rows, cursor + rank * batch)
    hi = min(n_rows, lo + batch)
    return lo, hi


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo,hi) split of n items."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flat_views(params):
    """One contiguous gradient buffer with a view per parameter (a single all-reduce per step)."""
    total = sum(p.numel() for p in params)
    flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
    views, off = [], 0
    for p in params:
        views.append(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    return flat, views


def allreduce_sums(t: torch.Tensor):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_eval_users(eval_users: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(int(eval_users.numel()), rank, world)
    return eval_users[lo:hi].contiguous()


class ReplicaTrainer(FusedTrainer):
    """Data-parallel replicas of the fused step (see module docstring)."""

    def __init__(self, model, inter, graph, batch_size, optim, sample_seed, use_cuda_graph=True, two_streams=True):
        self.rank, self.world = world_info()
        super().__init__(model, inter, graph, batch_size, optim, sample_seed, use_cuda_graph, two_streams)
        self.weak = True

    def _setup_params(self):
        m = self.model
        params = [m.uEmbd.weight, m.iEmbd.weight] + m._flat_stage_params()
        self.flat_grad, views = flat_views(params)
        for p, v in zip(params, views):
            p.grad = v
        super()._setup_params()

    # hooks used by FusedTrainer._step_body ---------------------------------------------------------
    def _row_offset(self):
        return self.rank * self.B

    def _row_stride(self):
        return self.B * self.world

    def _dropout_seed(self, seed):
        return (seed ^ (0x9E3779B97F4A7C15 * (self.rank + 1))) & 0xFFFFFFFFFFFFFFFF if self.world > 1 else seed

    def _reduce_grads(self):
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)      # sum of per-GPU mean-loss gradients

    def _collective_between(self):
        return self.world > 1

    def _epoch_loss(self, n):
        t = self.total.clone()
        allreduce_sums(t)          # the reference adds the per-GPU losses (loss.sum(), train_eval_Gowalla.py:139)
        return float(t.item()) / n

    def units_per_step(self):
        return 2 * self.g.E * self.world

    def parallelism(self):
        return "dp%d replicas (reference --parallel semantics): batch %d x %d rows/step, one NCCL all-reduce of %.1f MB grads" % (
            self.world, self.B, self.world, self.flat_grad.numel() * 4 / 2 ** 20)


# ------------------------------------------------------------------------------------------------
# north-star partition: users (and their edge rows) range-partitioned over the GPUs
# ------------------------------------------------------------------------------------------------
class UserShard:
    """Contiguous user range of this rank, balanced by EDGE count, and the rank-local graph: the CSR rows of the own
    users and the CSC restricted to them (items x own users).  Item-side state is replicated: the item-side partial sums
    of every stage are summed across ranks with ONE all-reduce (= reduce-scatter + all-gather in a single latency-bound
    NCCL call).  Local CSR edge ids are the global ones minus `edge_offset`, so the dropout streams (keyed
    by global edge id) are identical to the single-GPU run."""

    def __init__(self, edge_u, edge_i, U, I, rank, world, device):
        import numpy as np
        from .graph import BipartiteGraph
        key = np.unique(np.asarray(edge_u, np.int64) * np.int64(I) + np.asarray(edge_i, np.int64))     # global coalesce (host, once)
        eu, ei = key // I, key % I
        rowptr = np.zeros(U + 1, np.int64)
        np.add.at(rowptr, eu + 1, 1)
        rowptr = np.cumsum(rowptr)
        E = int(key.shape[0])
        # boundaries: first user whose cumulative edge count reaches r*E/world
        bounds = [int(np.searchsorted(rowptr, (E * r) // world, side="left")) for r in range(world)] + [U]
        bounds[0] = 0
        self.bounds = bounds
        self.rank, self.world, self.U, self.I, self.E = rank, world, U, I, E
        self.u_lo, self.u_hi = bounds[rank], bounds[rank + 1]
        self.edge_offset = int(rowptr[self.u_lo])
        sel = slice(int(rowptr[self.u_lo]), int(rowptr[self.u_hi]))
        idx = torch.from_numpy(np.stack([eu[sel], ei[sel]])).to(device)
        self.graph = BipartiteGraph(idx, U, I, check_users=False)
        self.local_edges = self.graph.E


class ShardedPropagation:
    """Propagation over a UserShard: own users complete locally, item rows via partial sums + all-reduce."""

    def __init__(self, shard: UserShard, stages=None):
        from .ops import D, STAGES
        from .propagation import Propagation
        self.sh = shard
        g = shard.graph
        self.p = Propagation(g, stages or STAGES)
        for t in self.p.h + self.p.Z + self.p.s + self.p.norm:      # rows of other ranks' users are never written: keep them finite
            t.zero_()
        self.p._bwd_buffers()
        for t in (self.p._bwd["G"] + [self.p._bwd["Ghat"], self.p._bwd["dh"], self.p._bwd["dN"], self.p._bwd["dS"]]):
            t.zero_()
        dev = g.device
        S = len(self.p.stages)
        self._fm = [torch.empty(g.N, dtype=torch.int64, device=dev) for _ in range(S)]
        self._em = [torch.empty(max(shard.E, 1), dtype=torch.uint8, device=dev) for _ in range(S)]     # GLOBAL edge ids
        self.featmask = [None] * S
        self.edgemask = [None] * S
        self.scale = 1.0

    def set_dropout(self, droprate, seed=0, call=0, call_dev=None):
        S = len(self.p.stages)
        if droprate <= 0:
            self.featmask, self.edgemask, self.scale = [None] * S, [None] * S, 1.0
            return
        self.scale = 1.0 / (1.0 - droprate)
        ops.dropout_masks(self._fm, self._em, [H for H, _ in self.p.stages], self.sh.graph.N, self.sh.E, seed, call, droprate, call_dev)
        self.featmask = list(self._fm)
        self.edgemask = [m[self.sh.edge_offset:] for m in self._em]       # local edge id + offset = global edge id

    def forward_plan(self, uEmbd, iEmbd, wtabs):
        """list of ('k', fn) compute items and ('c', fn) collectives."""
        p, sh, g = self.p, self.sh, self.sh.graph
        U, I, lo, hi = sh.U, sh.I, sh.u_lo, sh.u_hi
        plan = []
        for k, (H, _) in enumerate(p.stages):
            Xu = uEmbd if k == 0 else p.Z[k - 1]
            Xi = iEmbd if k == 0 else p.Z[k - 1][U:]
            act = 0 if k == 0 else 1

            def compute(k=k, H=H, Xu=Xu, Xi=Xi, act=act):
                fm = self.featmask[k]
                ops.transform_fwd(Xu[lo:hi], None, act, None if fm is None else fm[lo:], self.scale, wtabs[k], H, hi - lo, 0,
                                  p.h[k][lo:hi], p.s[k][lo:hi])
                ops.transform_fwd(None, Xi, act, None if fm is None else fm[U:], self.scale, wtabs[k], H, 0, I, p.h[k][U:], p.s[k][U:])
                ops.aggregate_fwd(g, p.scratch, p.counter, p.h[k], p.s[k], H, self.edgemask[k], self.scale, p.Z[k], p.norm[k],
                                  partial_from=g.T_users)
            plan.append(("k", compute))
            plan.append(("c", lambda k=k: (dist.all_reduce(p.Z[k][U:]), dist.all_reduce(p.norm[k][U:]))))
            plan.append(("k", lambda k=k, H=H: ops.aggregate_finalize(p.Z[k][U:], p.h[k][U:], p.norm[k][U:], H)))
        return plan

    def backward_plan(self, uEmbd, iEmbd, wtabs, gtabs, dU, dI, accumulate):
        """G_last = self.p.grad_in(): own-user rows and ALL item rows valid (item rows already summed across ranks)."""
        p, sh, g = self.p, self.sh, self.sh.graph
        U, I, lo, hi = sh.U, sh.I, sh.u_lo, sh.u_hi
        b = p._bwd
        plan = []
        G = b["G"][0]
        for k in range(len(p.stages) - 1, -1, -1):
            H, _ = p.stages[k]

            def edges(k=k, H=H, G=G):
                ops.stage_bwd_prep(G[lo:hi], p.Z[k][lo:hi], p.h[k][lo:hi], p.norm[k][lo:hi], H, b["Ghat"][lo:hi], self._dN(H)[lo:hi])
                ops.stage_bwd_prep(G[U:], p.Z[k][U:], p.h[k][U:], p.norm[k][U:], H, b["Ghat"][U:], self._dN(H)[U:])
                for mode in (0, 1):
                    ops.stage_bwd_edges(mode, g, p.scratch, p.counter, G, b["Ghat"], self._dN(H), p.h[k], p.s[k], H, self.edgemask[k],
                                        self.scale, wtabs[k], b["ds"], b["dh"], self._dS(H), partial=mode)
            plan.append(("k", edges))
            plan.append(("c", lambda H=H: (dist.all_reduce(b["dh"][U:]), dist.all_reduce(self._dS(H)[U:]))))
            Gprev = b["G"][1] if G is b["G"][0] else b["G"][0]

            def dense(k=k, H=H, G=G, Gprev=Gprev):
                ops.stage_bwd_finalize(b["dh"][U:], self._dS(H)[U:], G[U:], wtabs[k], H, 1)
                fm = self.featmask[k]
                if k > 0:
                    Zp = p.Z[k - 1]
                    ops.transform_bwd(b["dh"][lo:hi], self._dS(H)[lo:hi], p.h[k][lo:hi], Zp[lo:hi], None, 1, None if fm is None else fm[lo:],
                                      self.scale, wtabs[k], gtabs[k], H, hi - lo, 0, Gprev[lo:hi], None, 0, int(accumulate), b["ws"])
                    ops.transform_bwd(b["dh"][U:], self._dS(H)[U:], p.h[k][U:], None, Zp[U:], 1, None if fm is None else fm[U:], self.scale,
                                      wtabs[k], gtabs[k], H, 0, I, None, Gprev[U:], 0, int(accumulate), b["ws"])
                else:
                    ops.transform_bwd(b["dh"][lo:hi], self._dS(H)[lo:hi], p.h[k][lo:hi], uEmbd[lo:hi], None, 0, None if fm is None else fm[lo:],
                                      self.scale, wtabs[k], gtabs[k], H, hi - lo, 0, dU[lo:hi], None, int(accumulate), int(accumulate), b["ws"])
                    ops.transform_bwd(b["dh"][U:], self._dS(H)[U:], p.h[k][U:], None, iEmbd, 0, None if fm is None else fm[U:], self.scale,
                                      wtabs[k], gtabs[k], H, 0, I, None, dI, int(accumulate), int(accumulate), b["ws"])
            plan.append(("k", dense))
            G = Gprev
        return plan

    # dN/dS are allocated (N,8); the kernels index them as (N,H): use a compact (N,H) view of the same storage
    def _dN(self, H):
        return self.p._bwd["dN"].view(-1)[: self.sh.graph.N * H].view(self.sh.graph.N, H)

    def _dS(self, H):
        return self.p._bwd["dS"].view(-1)[: self.sh.graph.N * H].view(self.sh.graph.N, H)


class ShardedTrainer:
    """PairSampling step with the propagation sharded by user range (one process per GPU).  Same maths as the single-GPU
    FusedTrainer step on the same batch (strong scaling): every rank samples the same batch, scores the pairs whose user
    it owns, and the item-side partial sums are exchanged with NCCL all-reduces between the compute segments; each
    compute segment between two collectives is a captured CUDA graph."""

    def __init__(self, model, inter, edge_u, edge_i, batch_size, optim, sample_seed, use_cuda_graph=True):
        from .train import FusedTrainer   # noqa: F401  (shares the optimizer-state conventions)
        self.rank, self.world = world_info()
        self.model, self.inter, self.B, self.optim = model, inter, int(batch_size), optim
        self.sample_seed = int(sample_seed)
        dev = model.uEmbd.weight.device
        self.dev = dev
        self.shard = UserShard(edge_u, edge_i, model.userNum, model.itemNum, self.rank, self.world, dev)
        self.g = self.shard.graph
        self.props = [ShardedPropagation(self.shard, model.stages), ShardedPropagation(self.shard, model.stages)]
        self.use_cuda_graph = use_cuda_graph
        i64, f32 = dict(dtype=torch.int64, device=dev), dict(dtype=torch.float32, device=dev)
        self.users, self.pos, self.neg = torch.zeros(self.B, **i64), torch.zeros(self.B, **i64), torch.zeros(self.B, **i64)
        self.sc = [torch.zeros(self.B, **f32), torch.zeros(self.B, **f32)]
        self.dsc = [torch.zeros(self.B, **f32), torch.zeros(self.B, **f32)]
        self.loss = torch.zeros((), **f32)
        self.total = torch.zeros((), dtype=torch.float64, device=dev)
        self.row_dev = torch.zeros(2, **i64)
        self.call_dev = torch.zeros(1, **i64)
        # parameters / gradients (flat buffer -> one all-reduce) / Adam state
        m = model
        self.params = [m.uEmbd.weight, m.iEmbd.weight] + m._flat_stage_params()
        self.flat_grad, views = flat_views(self.params)
        for p, v in zip(self.params, views):
            p.grad = v
        stage_params = m.gat.stage_parameters()
        self.wtabs = [ops.pointer_table([p.detach() for p in st]) for st in stage_params]
        self.gtabs = [ops.pointer_table([p.grad for p in st]) for st in stage_params]
        group = optim.param_groups[0]
        self.hyper = dict(lr=float(group["lr"]), b1=float(group["betas"][0]), b2=float(group["betas"][1]), eps=float(group["eps"]),
                          wd=float(group["weight_decay"]))
        rows = []
        for p in self.params:
            st = optim.state[p]
            if len(st) == 0:
                st["step"], st["exp_avg"], st["exp_avg_sq"] = torch.tensor(0.0), torch.zeros_like(p), torch.zeros_like(p)
            rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()])
        self.adam_tab = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.adam_total = sum(p.numel() for p in self.params)
        self.adam_state = torch.tensor([float(optim.state[self.params[0]]["step"]), 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
        # gradients that every rank computes in full (item side) are summed `world` times by the all-reduce: pre-scale them
        self._replicated = [m.iEmbd.weight.grad]
        for (H, DH), st in zip(self.props[0].p.stages, stage_params):
            self._replicated += [p.grad for p in st[H:2 * H]]                 # W_i heads
            self._replicated += [p.grad[:, DH:] for p in st[2 * H:3 * H]]     # a_i halves
        self._segments = None
        self._cursor = 0

    # ------------------------------------------------------------------------------------------
    def _plan(self, droprate, seed):
        m, sh = self.model, self.shard
        U = sh.U
        uE, iE = m.uEmbd.weight.detach(), m.iEmbd.weight.detach()
        items = (self.pos, self.neg)
        B = self.B
        plan = []

        def head():
            ops.sample_pairs(self.inter, 0, B, self.sample_seed, 0, self.users, self.pos, self.neg, self.row_dev)
            for k in (0, 1):
                self.props[k].set_dropout(droprate, seed, k, self.call_dev)
        plan.append(("k", head))
        for k in (0, 1):
            plan += self.props[k].forward_plan(uE, iE, self.wtabs)

        def score_and_scatter():
            for k in (0, 1):
                ops.score_pairs(self.props[k].p.Z[-1], U, self.users, items[k], self.sc[k])
            ops.bpr_loss_owned(self.sc[0], self.sc[1], 1.0, self.loss, self.dsc[0], self.dsc[1], self.users, sh.u_lo, sh.u_hi)
            for k in (0, 1):
                G = self.props[k].p.grad_in()
                G.zero_()
                ops.score_pairs_bwd(self.props[k].p.Z[-1], U, self.users, items[k], self.dsc[k], G)
            m.uEmbd.weight.grad.zero_()          # rows of other ranks' users stay zero for the gradient all-reduce
        plan.append(("k", score_and_scatter))
        plan.append(("c", lambda: (dist.all_reduce(self.props[0].p.grad_in()[U:]), dist.all_reduce(self.props[1].p.grad_in()[U:]),
                                   dist.all_reduce(self.loss))))
        for k in (0, 1):
            plan += self.props[k].backward_plan(uE, iE, self.wtabs, self.gtabs, m.uEmbd.weight.grad, m.iEmbd.weight.grad, accumulate=(k == 1))

        def prescale():
            inv = 1.0 / self.world
            for gr in self._replicated:
                gr.mul_(inv)
        plan.append(("k", prescale))
        plan.append(("c", lambda: dist.all_reduce(self.flat_grad)))

        def update():
            h = self.hyper
            ops.adam_step_dev(self.adam_tab, len(self.params), self.adam_total, h["lr"], h["b1"], h["b2"], h["eps"], h["wd"], self.adam_state)
            self.total.add_(self.loss.double())
            ops.counter_add(self.row_dev, B)
            ops.counter_add(self.call_dev, 2)
        plan.append(("k", update))
        return plan

    def _compile(self, droprate, seed):
        """Merge consecutive compute items and capture each run as a CUDA graph; collectives stay eager between them."""
        plan = self._plan(droprate, seed)
        groups, cur = [], []
        for kind, fn in plan:
            if kind == "k":
                cur.append(fn)
            else:
                if cur:
                    groups.append(("k", cur))
                    cur = []
                groups.append(("c", fn))
        if cur:
            groups.append(("k", cur))
        if not self.use_cuda_graph:
            self._segments = [(kind, (lambda fs=f: [x() for x in fs]) if kind == "k" else f) for kind, f in groups]
            return
        # warm-up pass (eager) with state restored afterwards, then capture every compute group
        snap = [p.detach().clone() for p in self.params]
        opt_snap = [(self.optim.state[p]["exp_avg"].clone(), self.optim.state[p]["exp_avg_sq"].clone()) for p in self.params]
        misc = [t.clone() for t in (self.adam_state, self.total, self.row_dev, self.call_dev)]
        for kind, f in groups:
            if kind == "k":
                for x in f:
                    x()
            else:
                f()
        torch.cuda.synchronize(self.dev)
        segs = []
        for kind, f in groups:
            if kind == "c":
                segs.append(("c", f))
                continue
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for x in f:
                    x()
            segs.append(("k", gr.replay))
        with torch.no_grad():
            for p, sp, (a, b) in zip(self.params, snap, opt_snap):
                p.copy_(sp)
                self.optim.state[p]["exp_avg"].copy_(a)
                self.optim.state[p]["exp_avg_sq"].copy_(b)
            for t, s in zip((self.adam_state, self.total, self.row_dev, self.call_dev), misc):
                t.copy_(s)
        self._segments = segs
        self._seg_key = (droprate, seed)

    # ------------------------------------------------------------------------------------------
    def run_steps(self, n_steps, host_rows=None, read_loss=False):
        m = self.model
        m.train()
        droprate = m.droprate if m.droprate > 0 else 0.0
        seed = m._seed()
        n = len(self.inter)
        if self._segments is None or getattr(self, "_seg_key", None) != (droprate, seed):
            self._compile(droprate, seed)
            self._seg_key = (droprate, seed)
        self.row_dev.copy_(torch.tensor([self._cursor, 0], dtype=torch.int64))
        self.call_dev.fill_(m._call)
        losses = []
        for _ in range(n_steps):
            if self._cursor + self.B > n:
                self._cursor = 0
                self.row_dev.zero_()
            if host_rows is not None:
                lo = self._cursor
                self.inter.train_rows_user[lo:lo + self.B].copy_(host_rows[lo:lo + self.B], non_blocking=True)
            for kind, f in self._segments:
                f()
            self._cursor += self.B
            if read_loss:
                losses.append(float(self.loss.item()))
        m._call += 2 * n_steps
        return losses

    def launches_per_step(self, droprate):
        S = len(self.props[0].p.stages)
        per_prop = (2 * S if droprate > 0 else 0) + S * 4 + 1 + 1 + S * (2 + 2 + 1 + 4)
        return 1 + 2 * per_prop + 1 + 2 + 2 + len(self._replicated)

    def units_per_step(self):
        return 2 * self.shard.E          # strong scaling: the same global step on every world size

    def parallelism(self):
        return "users range-partitioned over %d GPUs by edge count (rank 0: users [%d,%d), %d of %d edges); item partials: NCCL all-reduce per stage" % (
            self.world, self.shard.u_lo, self.shard.u_hi, self.shard.local_edges, self.shard.E)

    def working_set_bytes(self):
        N, E = self.g.N, self.shard.local_edges
        return 2 * (8 * N * 256 + 8 * E * 4) + 7 * self.adam_total * 4

    def profile_kernels(self, n_steps=3):
        return []
