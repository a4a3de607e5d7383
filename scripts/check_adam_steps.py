"""1-GPU diagnostic: three fused training steps (droprate 0) vs the oracle, per-tensor relative error of the parameters and of the
first step's gradients.  Run with NGACF_DENSE=ffma / tc to compare the dense kernel families."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from ngacf_b200.data import Interactions
from ngacf_b200.graph import BipartiteGraph
from ngacf_b200.model import SPUIGACF
from ngacf_b200.optim import FusedAdam
from ngacf_b200.train import FusedTrainer

dev = torch.device("cuda:0")
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
tr = FusedTrainer(model, inter, graph, B, optim, sample_seed=5, use_cuda_graph=False)
g = port.build_graph(np.stack([u, i]), U, I)
it = port.build_interactions(U, I, tu, ti, su, si)
st = port.adam_init(p)
steps = 3
for s in range(steps):
    users, pos, neg = port.sample_pairs(it, s * B, s * B + B, 5, 0)
    _, grads, _, _ = port.train_step_grads(p, g, users, pos, neg)
    tr._step_body(B, 0, 0.0, tr._dropout_seed(model._seed()), s * B, 2 * s, False, part="compute")
    torch.cuda.synchronize()
    if s == 0:
        gsd = port.state_dict_from_params(grads)
        worst = []
        for k, prm in model.named_parameters():
            a, b = prm.grad.cpu().numpy().astype(np.float64), gsd[k].numpy().astype(np.float64)
            worst.append((float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)), k, float(np.abs(b).max())))
        worst.sort(reverse=True)
        print("step-0 gradient rel err (top 5):", [(round(e, 9), k, "%.2e" % m) for e, k, m in worst[:5]])
    tr._reduce_grads(); tr._step_update(False)
    port.adam_step(p, grads, st, 0.01, 1e-6)
torch.cuda.synchronize()
ref = port.state_dict_from_params(p)
worst = []
for k, v in model.state_dict().items():
    a, b = v.cpu().numpy().astype(np.float64), ref[k].numpy().astype(np.float64)
    worst.append((float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)), k))
worst.sort(reverse=True)
print("NGACF_DENSE=%s  params after %d Adam steps, rel err (top 5):" % (os.environ.get("NGACF_DENSE", "tc"), steps), [(round(e, 7), k) for e, k in worst[:5]])
