"""torchrun check (N GPUs): the user-range partition (ngacf_b200.dist.ShardedTrainer over NCCL) == the single-GPU fused step on
rank 0 (same batches, same Philox dropout streams).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngacf_b200 import hostdata  # noqa: E402
from ngacf_b200.data import Interactions  # noqa: E402
from ngacf_b200.dist import ShardedTrainer  # noqa: E402
from ngacf_b200.model import SPUIGACF  # noqa: E402
from ngacf_b200.optim import FusedAdam  # noqa: E402
from ngacf_b200.train import FusedTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    print("[rank %d]" % rank, *a, flush=True)


U, I, E, B = 3000, 5001, 200000, 1024
drop = float(os.environ.get("NGACF_DROP", "0.2"))
u, i = hostdata.synth_bipartite(U, I, E, 2)
(tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 3)
torch.manual_seed(7)
ref = SPUIGACF(U, I, 64, [64, 64], drop)
with torch.no_grad():
    ref.uEmbd.weight.mul_(10.0)
    ref.iEmbd.weight.mul_(10.0)
sd = {k: v.clone() for k, v in ref.state_dict().items()}


def make():
    m = SPUIGACF(U, I, 64, [64, 64], drop)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    m.drop_seed, m._call = 99, 0
    return m, Interactions.from_arrays(U, I, tu, ti, su, si, device=dev)


model, inter = make()
tr = ShardedTrainer(model, inter, u, i, B, FusedAdam(model.parameters(), lr=0.01, weight_decay=1e-6), sample_seed=5)
say("users [%d,%d) items [%d,%d) edges [%d,%d)" % (tr.u_lo, tr.u_hi, tr.i_lo, tr.i_hi, tr.e_lo, tr.e_hi))
steps = 6
losses = tr.run_steps(steps, read_loss=True)
torch.cuda.synchronize()
say("capture mode:", tr.capture_mode, "losses", ["%.6f" % x for x in losses])
tr.sync_embeddings()
# single-GPU reference on every rank (cheap at this size)
os.environ["NGACF_PRUNE"] = "0"
m1, it1 = make()
g1 = m1.graph_for(torch.from_numpy(np.stack([u, i])).to(dev))
t1 = FusedTrainer(m1, it1, g1, B, FusedAdam(m1.parameters(), lr=0.01, weight_decay=1e-6), sample_seed=5)
ref_losses = t1.run_steps(steps, read_loss=True)
worst = 0.0
for (k, a), (_, b) in zip(model.state_dict().items(), m1.state_dict().items()):
    a, b = a.double(), b.double()
    worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-30)))
flat = torch.cat([q.detach().reshape(-1) for q in model.parameters()])
mx, mn_ = flat.clone(), flat.clone()
dist.all_reduce(mx, op=dist.ReduceOp.MAX)
dist.all_reduce(mn_, op=dist.ReduceOp.MIN)
same = bool(torch.equal(mx, mn_))
ok = np.allclose(losses, ref_losses, rtol=1e-4) and worst < 5e-3 and same
say("sharded check: world=%d ok=%s | losses vs 1-GPU max rel %.2e | worst param rel err %.2e | tables identical on all ranks after sync: %s"
    % (world, ok, float(np.max(np.abs(np.array(losses) - np.array(ref_losses)) / np.abs(ref_losses))), worst, same))
tr.release()
dist.barrier()
say("leaving")
sys.stdout.flush()
os._exit(0 if ok else 1)       # no destroy_process_group: tearing NCCL down after captured collectives hung on the box
