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
