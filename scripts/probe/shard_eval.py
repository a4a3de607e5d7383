"""Diagnostic: the AllNeg scoring of ONE RANK's user shard of an N-GPU evaluation, on one GPU, for several segment counts S
(NGACF_TOPK_SEGMENTS; read per call).  python scripts/probe/shard_eval.py [world] [workload]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from ngacf_b200 import _lib, hostdata  # noqa: E402
from ngacf_b200.data import Interactions  # noqa: E402
from ngacf_b200.dist import shard_eval_users  # noqa: E402
from ngacf_b200.evaluate import AllNegEvaluator  # noqa: E402
from ngacf_b200.model import SPUIGACF  # noqa: E402

DEV = "cuda:0"
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
U, I, E = bench.SHAPES[sys.argv[2] if len(sys.argv) > 2 else "gowalla"][:3]
u, i = hostdata.synth_bipartite(U, I, E, 0)
(tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
torch.manual_seed(2019)
model = SPUIGACF(U, I, 64, [64, 64], 0.2).to(DEV).eval()
inter = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
adj = torch.from_numpy(np.stack([u, i])).to(DEV)
users = shard_eval_users(inter.eval_users, 0, world)
print("rank 0 of %d: %d users = %d blocks of 128" % (world, users.numel(), (users.numel() + 127) // 128))
with torch.no_grad():
    Z = model.propagate(adj)
    for S in ("", "1", "2", "3", "4", "6", "9", "12"):
        os.environ["NGACF_TOPK_SEGMENTS"] = S
        ev = AllNegEvaluator(inter, "tc", users=users)
        for _ in range(3):
            ev.rank(Z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ev.rank(Z)
        e1.record()
        torch.cuda.synchronize()
        _lib.PROFILE = []
        ev.rank(Z)
        torch.cuda.synchronize()
        sc = [a.elapsed_time(b) for n_, _, a, b in _lib.PROFILE if n_ == "ngacf_score_topk_tc"]
        _lib.PROFILE = None
        ev.resolve()
        print("S=%-7s rank(): %.3f ms   (score_topk_tc entry point %.3f ms, fallback rows %d)" % (S or "default", e0.elapsed_time(e1) / 10, sum(sc), ev.n_fallback))
