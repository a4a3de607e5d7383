"""Diagnostic: where does the pruned user pass spend its time?  Times ngacf_stage_bwd_edges_active(mode 0) and the list-based
forward on sub-ranges of the task list (tasks are ordered longest first: the head of each side is the 128-edge chunks of the
long rows).  Run on a GPU box: python scripts/probe/time_active_ranges.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ngacf_b200 import _lib, hostdata, ops  # noqa: E402
from ngacf_b200.data import Interactions  # noqa: E402
from ngacf_b200.model import SPUIGACF  # noqa: E402
from ngacf_b200.optim import FusedAdam  # noqa: E402
from ngacf_b200.train import FusedTrainer  # noqa: E402

DEV = "cuda:0"
U, I, E = 29858, 40981, 1027370
u, i = hostdata.synth_bipartite(U, I, E, 0)
(tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
torch.manual_seed(2019)
model = SPUIGACF(U, I, 64, [64, 64], 0.2).to(DEV).train()
inter = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
g = model.graph_for(torch.from_numpy(np.stack([u, i])).to(DEV))
tr = FusedTrainer(model, inter, g, 2048, FusedAdam(model.parameters(), lr=0.01, weight_decay=1e-6), sample_seed=0, two_streams=False)
row0 = (len(inter) // 2 // 2048) * 2048
tr._step_body(2048, 0, 0.2, model._seed(), row0, 0, False, part="compute")
torch.cuda.synchronize()
p = tr.props[0]
act = tr.active[0]
b = p._bwd
k = len(p.stages) - 1
H = 1
print("T", g.T, "T_users", g.T_users, "active tasks (users, items)", act.task_count.tolist())
tasks = g.tasks.cpu().numpy()
stamp = act.stamp.cpu().numpy()
val = act.val
is_act = stamp[tasks[:, 0]] == val
print("active user tasks", int(is_act[:g.T_users].sum()), "active item tasks", int(is_act[g.T_users:].sum()))
ln = tasks[:, 2] - tasks[:, 1]
print("user tasks with 128 edges:", int((ln[:g.T_users] == 128).sum()), " item tasks with 128:", int((ln[g.T_users:] == 128).sum()))


def timed(fn, reps=20):
    """GPU time per call: `reps` calls captured in one CUDA graph (no launch gaps), replayed until the clocks are up"""
    fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    for _ in range(50):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (10 * reps) * 1000.0


def users_range(t0, t1, mode=0):
    _lib.call("ngacf_stage_bwd_edges_active", mode, ops._p(g.tasks), t0, t1, ops._p(act.task_list), ops._p(act.task_count), ops._p(g.adj_ptr), ops._p(g.adj_idx), ops._p(g.adj_eid),
              ops._p(g.long_first_slot), ops._p(p.counter), ops._p(p.scratch), ops._p(b["G"][0]), ops._p(b["Ghat"]), ops._p(b["dN"]), ops._p(p.h[k]),
              ops._p(p.s[k]), H, ops._p(p.edgemask[k]), float(p.scale), ops._p(tr.wtabs[k]), g.U, ops._p(act.stamp), act.val, ops._p(act.val_dev),
              ops._p(act.edge_bits), ops._p(b["ds"]), ops._p(b["dh"]), ops._p(b["dS"]), ops._s())


Tu = g.T_users
n128 = int((ln[:Tu] == 128).sum())
for (a, c) in ((0, Tu), (0, n128), (n128, Tu), (n128, n128 + 4096), (Tu - 8192, Tu), (0, 16), (0, 256)):
    # note: partial ranges of long rows leave the arrival counters non-zero; re-arm them afterwards
    us = timed(lambda: users_range(a, c))
    torch.cuda.synchronize()
    p.counter.zero_()
    print("users_active tasks [%d,%d): %.1f us" % (a, c, us))
ni128 = int((ln[Tu:] == 128).sum())
for (a, c) in ((Tu, g.T), (Tu, Tu + ni128), (Tu + ni128, g.T)):
    us = timed(lambda: users_range(a, c, 1))
    p.counter.zero_()
    print("items_active tasks [%d,%d): %.1f us" % (a, c, us))
us = timed(lambda: ops.aggregate_fwd_active(g, p.scratch, p.counter, p.h[k], p.s[k], H, p.edgemask[k], p.scale, p.Z[k], p.norm[k], act))
print("aggregate_fwd_active: %.1f us" % us)
us = timed(lambda: act.mark(tr.users, tr.pos))
print("mark + plan: %.1f us" % us)
us = timed(lambda: ops.stage_bwd_prep_active(g, b["G"][0], p.Z[k], p.h[k], p.norm[k], H, b["Ghat"], b["dN"], act))
print("prep_active: %.1f us" % us)

