"""Tiny-graph training steps (eager, no CUDA graph) for compute-sanitizer:
    compute-sanitizer --tool racecheck python scripts/sanitize_step.py
Runs the fused PairSampling step with the pruned output stage and with the full one, long rows included, then an AllNeg evaluation."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngacf_b200 import hostdata  # noqa: E402
from ngacf_b200.data import Interactions  # noqa: E402
from ngacf_b200.evaluate import AllNegEvaluator  # noqa: E402
from ngacf_b200.model import SPUIGACF  # noqa: E402
from ngacf_b200.optim import FusedAdam  # noqa: E402
from ngacf_b200.train import FusedTrainer  # noqa: E402

DEV = "cuda:0"
U, I, E, B = 300, 200, 9000, 256          # items average 45 edges, the popular ones exceed the 128-edge chunk
u, i = hostdata.synth_bipartite(U, I, E, 3)
(tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
for prune in ("1", "0"):
    os.environ["NGACF_PRUNE"] = prune
    torch.manual_seed(1)
    model = SPUIGACF(U, I, 64, [64, 64], 0.2).to(DEV).train()
    inter = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
    g = model.graph_for(torch.from_numpy(np.stack([u, i])).to(DEV))
    tr = FusedTrainer(model, inter, g, B, FusedAdam(model.parameters(), lr=0.01, weight_decay=1e-6), sample_seed=0, use_cuda_graph=False)
    loss = tr.train_epoch(0, max_steps=2)
    torch.cuda.synchronize()
    print("prune", prune, "long rows", g.L, "epoch loss", loss)
model.eval()
with torch.no_grad():
    res = AllNegEvaluator(inter, "tc")(model.propagate(g))
torch.cuda.synchronize()
print("eval recall@20", float(res["recall"][3]))
