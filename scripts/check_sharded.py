"""torchrun check (N GPUs): user-range sharded training steps == the single-process oracle step (same batch, same dropout streams)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from ngacf_b200.data import Interactions
from ngacf_b200.dist import ShardedTrainer
from ngacf_b200.model import SPUIGACF
from ngacf_b200.optim import FusedAdam

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def say(*a): print("[rank %d]" % rank, *a, flush=True)
U, I, E, B = 600, 900, 30000, 256
u, i = port.synth_bipartite(U, I, E, 2)
(tu, ti), (su, si) = port.split_train_test(u, i, U, 3)
p = port.init_params(U, I, 7)
p["uEmbd"] *= 20; p["iEmbd"] *= 20
drop = float(os.environ.get("NGACF_DROP", "0.2"))
model = SPUIGACF(U, I, 64, [64, 64], drop)
model.load_state_dict(port.state_dict_from_params(p))
model = model.to(dev)
model.drop_seed = 99
inter = Interactions.from_arrays(U, I, tu, ti, su, si, device=dev)
optim = FusedAdam(model.parameters(), lr=0.01, weight_decay=1e-6)
tr = ShardedTrainer(model, inter, u, i, B, optim, sample_seed=5, use_cuda_graph=os.environ.get("NGACF_DIST_GRAPH", "1") != "0")
say("shard users [%d,%d) local edges %d of %d" % (tr.shard.u_lo, tr.shard.u_hi, tr.shard.local_edges, tr.shard.E))
steps = 3
losses = tr.run_steps(steps, read_loss=True)
torch.cuda.synchronize()
say("steps done", losses)
g = port.build_graph(np.stack([u, i]), U, I)
it = port.build_interactions(U, I, tu, ti, su, si)
st = port.adam_init(p)
ref_losses = []
for s in range(steps):
    users, pos, neg = port.sample_pairs(it, s * B, (s + 1) * B, 5, 0)
    mp = port.dropout_masks(g, 99, 2 * s, drop) if drop > 0 else None
    mn = port.dropout_masks(g, 99, 2 * s + 1, drop) if drop > 0 else None
    loss, grads, _, _ = port.train_step_grads(p, g, users, pos, neg, mp, mn, drop)
    port.adam_step(p, grads, st, 0.01, 1e-6)
    ref_losses.append(float(loss))
ref = port.state_dict_from_params(p)
worst = 0.0
for k, v in model.state_dict().items():
    a, b = v.cpu().numpy().astype(np.float64), ref[k].numpy().astype(np.float64)
    worst = max(worst, float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)))
flat = torch.cat([q.detach().reshape(-1) for q in model.parameters()])
mx, mn_ = flat.clone(), flat.clone()
dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn_, op=dist.ReduceOp.MIN)
same = bool(torch.equal(mx, mn_))
if rank == 0:
    print("sharded check: world=%d losses %s vs oracle %s; worst param rel err %.3e; replicas bit-identical: %s" % (world, losses, ref_losses, worst, same))
    assert np.allclose(losses, ref_losses, rtol=1e-4) and worst < 2e-3 and same
dist.destroy_process_group()
