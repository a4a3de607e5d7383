"""Diagnostic: CUDA-event time of every C-ABI call inside one AllNeg evaluation (propagate + rank + metrics).
    python scripts/probe/profile_eval_calls.py [workload]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from ngacf_b200 import _lib, hostdata  # noqa: E402
from ngacf_b200.data import Interactions  # noqa: E402
from ngacf_b200.evaluate import AllNegEvaluator  # noqa: E402
from ngacf_b200.model import SPUIGACF  # noqa: E402

DEV = "cuda:0"
U, I, E = bench.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "amazon-book"][:3]
u, i = hostdata.synth_bipartite(U, I, E, 0)
(tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
torch.manual_seed(2019)
model = SPUIGACF(U, I, 64, [64, 64], 0.2).to(DEV).eval()
inter = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
adj = torch.from_numpy(np.stack([u, i])).to(DEV)
ev = AllNegEvaluator(inter, "tc")
with torch.no_grad():
    for _ in range(3):
        model._eval_key = None
        ev(model.propagate(adj))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model._eval_key = None
        ev(model.propagate(adj))
    e1.record()
    torch.cuda.synchronize()
    print("evaluation: %.3f ms" % (e0.elapsed_time(e1) / 5))
    _lib.PROFILE = []
    model._eval_key = None
    ev(model.propagate(adj))
    torch.cuda.synchronize()
    for name, _, a, b in _lib.PROFILE:
        print("  %-34s %8.3f ms" % (name, a.elapsed_time(b)))
    _lib.PROFILE = None
    print("fallback rows", ev.n_fallback)

# cost of the exact fallback for a handful of flagged rows: one CTA per 16 users over all items vs the item-split entry point
from ngacf_b200 import ops  # noqa: E402
for n in (1, 16, 64):
    users = inter.eval_users[:n].contiguous()
    ids = torch.empty((n, 20), dtype=torch.int32, device=DEV)
    sc = torch.empty((n, 20), dtype=torch.float32, device=DEV)
    for name, fn in (("score_topk_exact", ops.score_topk_exact), ("score_topk_exact_split", ops.score_topk_exact_split)):
        fn(ev.F, U, I, users, inter, ids, sc)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn(ev.F, U, I, users, inter, ids, sc)
        e1.record()
        torch.cuda.synchronize()
        print("%-24s %3d users: %.3f ms" % (name, n, e0.elapsed_time(e1) / 5))
