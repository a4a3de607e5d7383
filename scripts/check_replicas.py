"""torchrun check (N GPUs): ReplicaTrainer steps == single-GPU reference computation of the summed gradients."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from ngacf_b200.data import Interactions
from ngacf_b200.dist import ReplicaTrainer
from ngacf_b200.graph import BipartiteGraph
from ngacf_b200.model import SPUIGACF
from ngacf_b200.optim import FusedAdam

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
U, I, E, B = 600, 900, 30000, 256
u, i = port.synth_bipartite(U, I, E, 2)
(tu, ti), (su, si) = port.split_train_test(u, i, U, 3)
p = port.init_params(U, I, 7)
p["uEmbd"] *= 20; p["iEmbd"] *= 20
model = SPUIGACF(U, I, 64, [64, 64], 0.0)
model.load_state_dict(port.state_dict_from_params(p))
model = model.to(dev)
inter = Interactions.from_arrays(U, I, tu, ti, su, si, device=dev)
graph = BipartiteGraph(torch.from_numpy(np.stack([u, i])).to(dev), U, I)
optim = FusedAdam(model.parameters(), lr=0.01, weight_decay=1e-6)
def say(*a):
    print("[rank %d]" % rank, *a, flush=True)
say("pg up")
t = torch.ones(4, device=dev); dist.all_reduce(t); torch.cuda.synchronize(); say("eager all_reduce ok", t[0].item())
use_graph = os.environ.get("NGACF_DIST_GRAPH", "1") != "0"
tr = ReplicaTrainer(model, inter, graph, B, optim, sample_seed=5, use_cuda_graph=use_graph)
steps = 3
if use_graph:
    tr.run_steps(steps)
else:
    for sidx in range(steps):
        tr._step_body(B, 0, 0.0, tr._dropout_seed(model._seed()), sidx * B * world + rank * B, 2 * sidx, False)
torch.cuda.synchronize()
say("steps done, graph=%s" % use_graph)
# oracle: same rows, summed per-replica mean-loss gradients, Adam
g = port.build_graph(np.stack([u, i]), U, I)
it = port.build_interactions(U, I, tu, ti, su, si)
st = port.adam_init(p)
for s in range(steps):
    tot = None
    for r in range(world):
        lo = s * B * world + r * B
        users, pos, neg = port.sample_pairs(it, lo, lo + B, 5, 0)
        _, grads, _, _ = port.train_step_grads(p, g, users, pos, neg)
        tot = grads if tot is None else port.add_grads(tot, grads)
    port.adam_step(p, tot, st, 0.01, 1e-6)
ref = port.state_dict_from_params(p)
worst = 0.0
for k, v in model.state_dict().items():
    a, b = v.cpu().numpy().astype(np.float64), ref[k].numpy().astype(np.float64)
    worst = max(worst, float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)))
# replicas identical?
flat = torch.cat([q.detach().reshape(-1) for q in model.parameters()])
mx = flat.clone(); mn = flat.clone()
dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
same = bool(torch.equal(mx, mn))
if rank == 0:
    print("replica check: world=%d worst rel err vs oracle %.3e, replicas bit-identical: %s" % (world, worst, same))
    assert worst < 2e-3 and same
dist.destroy_process_group()
