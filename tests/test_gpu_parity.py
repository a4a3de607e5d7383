"""GPU parity tests: every CUDA entry point, called through the C-ABI (ctypes), against the oracle
(oracle/port.py) and the committed reference fixtures (tests/golden).  Bit-exact for integer work
(graph, masks, sampler), 1e-4 relative for fp32 (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def sd_from(gz, prefix):
    return {k[len(prefix):]: torch.from_numpy(gz[k]) for k in gz.files if k.startswith(prefix)}


def cuda_graph(u, i, U, I):
    from ngacf_b200.graph import BipartiteGraph
    idx = torch.from_numpy(np.stack([u, i]).astype(np.int64)).to(DEV)
    return BipartiteGraph(idx, U, I)


# ------------------------------------------------------------------------------------------------
# graph builder
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("U,I,E,seed", [(50, 70, 900, 0), (400, 700, 9000, 1), (3000, 5000, 200000, 2)])
def test_graph_build_bit_exact(U, I, E, seed):
    rng = np.random.default_rng(seed)
    u, i = port.synth_bipartite(U, I, E, seed)
    # duplicates + shuffled input
    dup = rng.integers(0, E, E // 10)
    u2, i2 = np.concatenate([u, u[dup]]), np.concatenate([i, i[dup]])
    p = rng.permutation(u2.shape[0])
    g = cuda_graph(u2[p], i2[p], U, I)
    ref = port.build_graph(np.stack([u, i]), U, I)
    assert g.E == ref.E
    assert np.array_equal(g.rowptr.cpu().numpy(), ref.rowptr)
    assert np.array_equal(g.colidx.cpu().numpy(), ref.colidx)
    assert np.array_equal(g.colptr.cpu().numpy(), ref.colptr)
    assert np.array_equal(g.rowidx.cpu().numpy(), ref.rowidx)
    assert np.array_equal(g.perm.cpu().numpy(), ref.perm)
    # unified adjacency
    adj_ptr = g.adj_ptr.cpu().numpy()
    assert np.array_equal(adj_ptr[:U + 1], ref.rowptr) and np.array_equal(adj_ptr[U:], ref.colptr + ref.E)
    adj = g.adj_idx.cpu().numpy()
    assert np.array_equal(adj[:ref.E], ref.colidx + U) and np.array_equal(adj[ref.E:], ref.rowidx)
    eid = g.adj_eid.cpu().numpy()
    assert np.array_equal(eid[:ref.E], np.arange(ref.E)) and np.array_equal(eid[ref.E:], ref.perm)
    # tasks: cover every adjacency range exactly once, <= CHUNK edges unless single-task, users first, longest first
    t = g.tasks.cpu().numpy()
    deg = np.diff(adj_ptr)
    ntasks = np.where(deg <= 128, 1, -(-deg // 128))
    assert t.shape[0] == ntasks.sum() == g.T
    covered = np.zeros(2 * ref.E + 1, np.int32)
    for node, b, e, lid in t:
        covered[b:e] += 1
        assert adj_ptr[node] <= b <= e <= adj_ptr[node + 1]
        assert (lid >= 0) == (deg[node] > 128)
    assert (covered[:2 * ref.E] == 1).all()
    side = (t[:, 0] >= U).astype(int)
    assert (np.diff(side) >= 0).all() and side[:g.T_users].sum() == 0 and side[g.T_users:].all()
    ln = t[:, 2] - t[:, 1]
    assert (np.diff(ln[:g.T_users]) <= 0).all() and (np.diff(ln[g.T_users:]) <= 0).all()
    assert g.L == int((deg > 128).sum()) and g.S == int(ntasks[deg > 128].sum())


def test_graph_build_rejects_bad_input():
    from ngacf_b200.graph import BipartiteGraph
    idx = torch.tensor([[0, 2], [1, 1]], device=DEV)
    with pytest.raises(ValueError):
        BipartiteGraph(idx, 3, 2)            # user 1 has no edge (SPUIGACF.py:368)
    with pytest.raises(ValueError):
        BipartiteGraph(torch.tensor([[0, 5], [1, 1]], device=DEV), 3, 2)   # out of range


# ------------------------------------------------------------------------------------------------
# dropout masks / sampler: bit exact vs the CPU restatement
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [0.1, 0.2, 0.5])
def test_dropout_masks_bit_exact(p):
    from ngacf_b200 import ops
    N, E = 3001, 70001
    for stage, H in ((0, 8), (1, 1)):
        fm = torch.empty(N, dtype=torch.int64, device=DEV)
        em = torch.empty(E, dtype=torch.uint8, device=DEV)
        ops.feature_mask(fm, 0x1234567890ABCDEF, 7, stage, p)
        ops.edge_mask(em, H, 0x1234567890ABCDEF, 7, stage, p)
        assert np.array_equal(fm.cpu().numpy().view(np.uint64), port.feature_mask_bits(N, 0x1234567890ABCDEF, 7, stage, p))
        assert np.array_equal(em.cpu().numpy(), port.edge_mask_bits(E, H, 0x1234567890ABCDEF, 7, stage, p))
    # the single-launch form the propagation uses: all stages at once, same bits (three stages = SPUIMultiGACF)
    heads = (8, 8, 1)
    fms = [torch.empty(N, dtype=torch.int64, device=DEV) for _ in heads]
    ems = [torch.empty(E, dtype=torch.uint8, device=DEV) for _ in heads]
    ops.dropout_masks(fms, ems, heads, N, E, 0x1234567890ABCDEF, 7, p)
    for stage, H in enumerate(heads):
        assert np.array_equal(fms[stage].cpu().numpy().view(np.uint64), port.feature_mask_bits(N, 0x1234567890ABCDEF, 7, stage, p))
        assert np.array_equal(ems[stage].cpu().numpy(), port.edge_mask_bits(E, H, 0x1234567890ABCDEF, 7, stage, p))


def make_interactions(U, I, E, seed):
    from ngacf_b200.data import Interactions
    u, i = port.synth_bipartite(U, I, E, seed)
    (tu, ti), (su, si) = port.split_train_test(u, i, U, seed + 1)
    it = port.build_interactions(U, I, tu, ti, su, si)
    return it, Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV), (u, i)


def test_sampler_bit_exact():
    from ngacf_b200 import ops
    it, dit, _ = make_interactions(300, 500, 6000, 3)
    n = it.train_rows_user.shape[0]
    for (lo, hi, seed, epoch) in ((0, n, 11, 0), (100, 1500, 2 ** 40 + 5, 3)):
        users = torch.empty(hi - lo, dtype=torch.int64, device=DEV)
        pos = torch.empty_like(users)
        neg = torch.empty_like(users)
        ops.sample_pairs(dit, lo, hi, seed, epoch, users, pos, neg)
        ru, rp, rn = port.sample_pairs(it, lo, hi, seed, epoch)
        assert np.array_equal(users.cpu().numpy(), ru)
        assert np.array_equal(pos.cpu().numpy(), rp)
        assert np.array_equal(neg.cpu().numpy(), rn)


# ------------------------------------------------------------------------------------------------
# model forward / backward vs the reference fixtures
# ------------------------------------------------------------------------------------------------
def make_model(gz, prefix, droprate=0.0):
    from ngacf_b200.model import SPUIGACF
    U, I = int(gz["U"]), int(gz["I"])
    m = SPUIGACF(U, I, 64, [64, 64], droprate)
    m.load_state_dict(sd_from(gz, prefix))
    return m.to(DEV)


@pytest.mark.parametrize("name", ["fwd_bwd_small", "fwd_bwd_medium"])
def test_forward_backward_vs_reference(golden, name):
    gz = golden(name)
    model = make_model(gz, "sd/")
    model.train()
    adj = torch.from_numpy(np.stack([gz["edge_u"], gz["edge_i"]])).to(DEV)
    users = torch.from_numpy(gz["users"]).to(DEV)
    items = torch.from_numpy(gz["items"]).to(DEV)
    sc = model(users, items, adj)
    assert rel_err(sc.detach().cpu().numpy(), gz["scores_f64"]) < 1e-4
    (sc * torch.from_numpy(gz["w"]).float().to(DEV)).sum().backward()
    for k, v in model.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), gz["grad_f64/" + k]) < 1e-4, k


@pytest.mark.parametrize("name", ["fwd_bwd_small", "fwd_bwd_medium"])
def test_forward_backward_with_injected_dropout(golden, name):
    """masks generated ON THE GPU by the Philox kernels == the masks that were injected into the reference."""
    gz = golden(name)
    model = make_model(gz, "sd/", float(gz["drop_p"]))
    model.train()
    model.drop_seed = int(gz["drop_seed"])
    model._call = int(gz["drop_call"])
    adj = torch.from_numpy(np.stack([gz["edge_u"], gz["edge_i"]])).to(DEV)
    users = torch.from_numpy(gz["users"]).to(DEV)
    items = torch.from_numpy(gz["items"]).to(DEV)
    sc = model(users, items, adj)
    assert rel_err(sc.detach().cpu().numpy(), gz["scores_drop_f64"]) < 1e-4
    (sc * torch.from_numpy(gz["w"]).float().to(DEV)).sum().backward()
    for k, v in model.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), gz["grad_drop_f64/" + k]) < 1e-4, k


def test_long_rows_and_isolated_items_vs_port():
    """Rows longer than CHUNK (split across tasks, combined by the last arriver) and items with no edge
    (NaN -> 0 path, SPUIGACF.py:389), fp32 GPU vs fp64 port."""
    U, I, E = 600, 900, 60000
    u, i = port.synth_bipartite(U, I - 7, E, 5)       # items I-7.. have no edges
    ref = port.build_graph(np.stack([u, i]), U, I)
    assert (np.diff(ref.rowptr) > 128).any() and (np.diff(ref.colptr) > 128).any() and (np.diff(ref.colptr) == 0).any()
    p64 = port.init_params(U, I, 3, torch.float64)
    p64["uEmbd"] *= 20
    p64["iEmbd"] *= 20
    from ngacf_b200.model import SPUIGACF
    model = SPUIGACF(U, I, 64, [64, 64], 0.25)
    model.load_state_dict({k: v.float() for k, v in port.state_dict_from_params(p64).items()})
    model = model.to(DEV).train()
    model.drop_seed, model._call = 99, 3
    masks = port.dropout_masks(ref, 99, 3, 0.25)
    F, caches = port.propagate(p64, ref, masks, 0.25)
    rng = np.random.default_rng(0)
    users, items, w = rng.integers(0, U, 512), rng.integers(0, I, 512), rng.standard_normal(512)
    sc_ref = port.scores(F, U, users, items)
    ut, itt, wt = torch.from_numpy(users), torch.from_numpy(items) + U, torch.from_numpy(w)
    dF = torch.zeros_like(F)
    dF.index_add_(0, ut, wt[:, None] * F[itt])
    dF.index_add_(0, itt, wt[:, None] * F[ut])
    gref = port.state_dict_from_params(port.propagate_backward(dF, p64, ref, caches))
    adj = torch.from_numpy(np.stack([u, i])).to(DEV)
    sc = model(torch.from_numpy(users).to(DEV), torch.from_numpy(items).to(DEV), adj)
    assert rel_err(sc.detach().cpu().numpy(), sc_ref.numpy()) < 1e-4
    (sc * wt.float().to(DEV)).sum().backward()
    for k, v in model.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), gref[k].numpy()) < 1e-4, k
    # determinism: a second identical call gives bit-identical scores (fixed combine order of long rows)
    model._call = 3
    sc2 = model(torch.from_numpy(users).to(DEV), torch.from_numpy(items).to(DEV), adj)
    assert torch.equal(sc, sc2)


def test_score_tree_bit_exact():
    """scores follow the specified summation tree exactly (given the GPU's own ELU values)."""
    from ngacf_b200 import ops
    rng = np.random.default_rng(4)
    U, I = 37, 53
    Z = torch.from_numpy(rng.standard_normal((U + I, 64)).astype(np.float32)).to(DEV)
    F = torch.empty_like(Z)
    ops.final_features(Z, F)
    users = torch.from_numpy(rng.integers(0, U, 300)).to(DEV)
    items = torch.from_numpy(rng.integers(0, I, 300)).to(DEV)
    out = torch.empty(300, device=DEV)
    ops.score_pairs(Z, U, users, items, out)
    Fn = F.cpu().numpy()
    ref = port.dot64_tree(Fn[users.cpu().numpy()], Fn[items.cpu().numpy() + U])
    assert np.array_equal(out.cpu().numpy(), ref)
    Fref = torch.nn.functional.elu(Z.cpu().double()).numpy()
    assert rel_err(Fn, Fref) < 1e-6


def test_bpr_loss_and_saturation():
    from ngacf_b200.loss import BPRLoss
    pos = torch.tensor([0.3, -2.0, 40.0, -40.0, 120.0, -120.0, 0.0], device=DEV, requires_grad=True)
    neg = torch.tensor([0.1, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0], device=DEV, requires_grad=True)
    loss = BPRLoss()(pos, neg)
    loss.backward()
    p64 = pos.detach().cpu().double().requires_grad_()
    n64 = neg.detach().cpu().double().requires_grad_()
    ref = torch.nn.functional.softplus(-(p64 - n64)).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert np.allclose(pos.grad.cpu().numpy(), p64.grad.numpy(), rtol=1e-5, atol=1e-12)
    assert np.allclose(neg.grad.cpu().numpy(), n64.grad.numpy(), rtol=1e-5, atol=1e-12)
    assert np.isfinite(loss.item())


def test_adam_matches_torch():
    from ngacf_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(1000, 64), (777, 64), (64, 8), (1, 16), (64, 64), (1, 128)]
    ps = [torch.randn(s, device=DEV).requires_grad_() for s in shapes]
    qs = [p.detach().clone().requires_grad_() for p in ps]
    a = FusedAdam(ps, lr=0.01, weight_decay=1e-3)
    b = torch.optim.Adam(qs, lr=0.01, weight_decay=1e-3)
    for step in range(5):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad = g.clone()
            q.grad = g.clone()
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        assert rel_err(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 1e-5
    sd = a.state_dict()
    b.load_state_dict(sd)          # state layout is torch.optim.Adam's (checkpoint compatibility)


@pytest.mark.parametrize("fused", [False, True, "split"])
def test_train_epochs_vs_reference(golden, fused, monkeypatch):
    """Two PairSampling epochs (dropout 0.2, Adam, GPU sampler + GPU Philox masks) vs the reference run that
    had the same samples and masks injected (tests/golden/train_eval_small.npz)."""
    import train_eval_Gowalla as T
    from ngacf_b200.data import Interactions
    from ngacf_b200.loss import BPRLoss
    from ngacf_b200.optim import FusedAdam
    gz = golden("train_eval_small")
    if fused == "split":       # dX / dW split kernels with the weight gradients deferred to a third stream
        monkeypatch.setenv("NGACF_SPLIT_DENSE_BWD", "1")
        fused = True
    model = make_model(gz, "sd0/", float(gz["droprate"]))
    U, I = int(gz["U"]), int(gz["I"])
    dit = Interactions.from_arrays(U, I, gz["train_u"], gz["train_i"], gz["test_u"], gz["test_i"], device=DEV)
    adj = torch.from_numpy(np.stack([np.concatenate([gz["train_u"], gz["test_u"]]), np.concatenate([gz["train_i"], gz["test_i"]])]))
    optim = FusedAdam(model.parameters(), lr=float(gz["lr"]), weight_decay=float(gz["wd"]))
    model.drop_seed = int(gz["drop_seed"])
    losses = []
    lossfn = BPRLoss()
    for ep in range(int(gz["epochs"])):
        losses.append(T.train_bpr(model, int(gz["batch"]), dit, dit, adj, optim, lossfn, False,
                                  epoch=ep, sample_seed=int(gz["sample_seed"]), fused=fused))
    assert rel_err(np.array(losses), gz["epoch_losses"]) < 1e-4
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    for k in sd:
        assert rel_err(sd[k], gz["sd1/" + k]) < 5e-3, k
    # eval on the GPU-trained weights: metrics within tolerance of the reference's eval of ITS weights
    res = T.eval_neg_all(model, 2048, dit, dit, adj, I, False)
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        assert np.allclose(res[k], gz["eval/" + k], rtol=2e-2, atol=2e-3), (k, res[k], gz["eval/" + k])


# ------------------------------------------------------------------------------------------------
# AllNeg evaluation
# ------------------------------------------------------------------------------------------------
def _eval_case(U, I, E, seed, dup_items=0):
    from ngacf_b200 import ops
    it, dit, _ = make_interactions(U, I, E, seed)
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((U + I, 64)).astype(np.float32) * 0.5
    if dup_items:       # identical item rows -> exactly equal scores -> the tie rule (lowest id) decides
        src = rng.integers(0, I, dup_items)
        dst = rng.integers(0, I, dup_items)
        Z[U + dst] = Z[U + src]
    Zt = torch.from_numpy(Z).to(DEV)
    F = torch.empty_like(Zt)
    ops.final_features(Zt, F)
    return it, dit, Zt, F.cpu().numpy()


@pytest.mark.parametrize("mode", ["exact", "tc"])
@pytest.mark.parametrize("U,I,E,seed,dup", [(90, 150, 1500, 1, 40), (300, 1000, 9000, 2, 200), (1000, 3000, 40000, 3, 0)])
def test_eval_topk_ids_bit_exact(mode, U, I, E, seed, dup):
    from ngacf_b200 import _lib
    from ngacf_b200.evaluate import AllNegEvaluator
    if mode == "tc" and not hasattr(_lib.load(), "ngacf_score_topk_tc"):
        pytest.skip("tensor-core path not built")
    it, dit, Zt, Fn = _eval_case(U, I, E, seed, dup)
    ev = AllNegEvaluator(dit, mode)
    ev.rank(Zt)
    users = dit.eval_users.cpu().numpy()
    top = ev.top_ids.cpu().numpy()
    tsc = ev.top_scores.cpu().numpy()
    Fi = Fn[U:]
    for j, u in enumerate(users):
        sc = port.dot64_tree(np.broadcast_to(Fn[u], Fi.shape), Fi)
        ref = port.topk_allneg(sc, it, int(u))
        assert np.array_equal(top[j][:ref.shape[0]], ref), (mode, u)
        assert np.array_equal(tsc[j][:ref.shape[0]], sc[ref]), (mode, u)
        assert (top[j][ref.shape[0]:] == -1).all()
    if mode == "tc":
        # the tensor-core result itself must be the exact one for (nearly) every row: the exact fallback is an escape
        # hatch for rows whose error-bound proof fails, not the path that makes this test pass
        assert ev.n_fallback <= max(2, len(users) // 20), ev.n_fallback


def test_eval_metrics_vs_reference_fixture(golden):
    """hit lists + metric sums from the reference's own top-20 lists (tests/golden/train_eval_small.npz)."""
    from ngacf_b200 import ops
    from ngacf_b200.data import Interactions
    gz = golden("train_eval_small")
    U, I = int(gz["U"]), int(gz["I"])
    dit = Interactions.from_arrays(U, I, gz["train_u"], gz["train_i"], gz["test_u"], gz["test_i"], device=DEV)
    assert np.array_equal(dit.eval_users.cpu().numpy(), gz["eval/users"])
    n = dit.eval_users.numel()
    top = torch.from_numpy(gz["eval/top20"].astype(np.int32)).to(DEV)
    hits = torch.empty((n, 20), dtype=torch.uint8, device=DEV)
    sums = torch.zeros(16, dtype=torch.float64, device=DEV)
    ws = torch.empty(n * 16, dtype=torch.float64, device=DEV)
    ops.eval_metrics(top, dit.eval_users, dit, hits, sums, ws)
    s = sums.cpu().numpy() / dit.n_train_users
    for q, k in enumerate(("precision", "recall", "ndcg", "hit_ratio")):
        assert np.allclose(s[4 * q:4 * q + 4], gz["eval/" + k], rtol=1e-12, atol=1e-15), k


def test_eval_topk_on_reference_scores_fixture(golden):
    """Same fp32 embeddings as the reference model -> same top-20 ids as its ranklist_by_heapq, except where
    two reference scores differ by less than fp32 summation-order noise."""
    import train_eval_Gowalla as T
    from ngacf_b200.data import Interactions
    gz = golden("train_eval_small")
    model = make_model(gz, "sd1/")
    U, I = int(gz["U"]), int(gz["I"])
    dit = Interactions.from_arrays(U, I, gz["train_u"], gz["train_i"], gz["test_u"], gz["test_i"], device=DEV)
    adj = torch.from_numpy(np.stack([np.concatenate([gz["train_u"], gz["test_u"]]), np.concatenate([gz["train_i"], gz["test_i"]])]))
    res = T.eval_neg_all(model, 2048, dit, dit, adj, I, False, mode="exact")
    top = model._evaluator.top_ids.cpu().numpy()
    ref_top, ref_sc = gz["eval/top20"], gz["eval/scores"]           # the reference's own top-20 lists and its fp32 score rows
    assert (top == ref_top).mean() > 0.995
    # every slot that differs must be a near tie IN THE REFERENCE'S OWN SCORES: the two items' reference scores differ by no more
    # than fp32 summation-order noise (64 products: 64 eps * |u| * |i| bounds it; measured with the reference's numbers only)
    eps = np.finfo(np.float32).eps
    with torch.no_grad():
        F = torch.nn.functional.elu(model.propagate(adj.to(DEV))).double().cpu().numpy()
    users = gz["eval/users"]
    n_diff = 0
    for r, uu in enumerate(users):
        for j in np.nonzero(top[r] != ref_top[r])[0]:
            a, b = int(top[r, j]), int(ref_top[r, j])
            bound = 64 * eps * np.linalg.norm(F[uu]) * max(np.linalg.norm(F[U + a]), np.linalg.norm(F[U + b]))
            assert abs(float(ref_sc[r, a]) - float(ref_sc[r, b])) <= bound, (uu, j, a, b, ref_sc[r, a], ref_sc[r, b], bound)
            n_diff += 1
    assert n_diff <= 0.005 * top.size
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        assert np.allclose(res[k], gz["eval/" + k], rtol=1e-3, atol=1e-4), k


def test_config1_ml100k_two_epochs_vs_reference(golden):
    """BASELINE config 1 (the reference README's smoke test): SPUIGACF on the real ml100k split, PairSampling, AllNeg,
    2 epochs, lr 0.002, wd 1e-6, droprate 0.2, B 2048 -- GPU run vs the UNMODIFIED reference's run with the same
    sampler/dropout streams injected (tests/golden/ml100k_2epochs.npz, oracle/make_golden.py:case_ml100k)."""
    import train_eval_Gowalla as T
    from ngacf_b200.data import Interactions
    from ngacf_b200.loss import BPRLoss
    gz = golden("ml100k_2epochs")
    model = make_model(gz, "sd0/", float(gz["droprate"]))
    U, I = int(gz["U"]), int(gz["I"])
    dit = Interactions.from_arrays(U, I, gz["train_u"], gz["train_i"], gz["test_u"], gz["test_i"], device=DEV)
    adj = torch.from_numpy(np.stack([np.concatenate([gz["train_u"], gz["test_u"]]), np.concatenate([gz["train_i"], gz["test_i"]])]).astype(np.int64))
    optim = torch.optim.Adam(model.parameters(), lr=float(gz["lr"]), weight_decay=float(gz["wd"]))   # the reference's own optimizer class
    model.drop_seed = int(gz["drop_seed"])
    lossfn = BPRLoss()
    losses = [T.train_bpr(model, int(gz["batch"]), dit, dit, adj, optim, lossfn, False, epoch=ep, sample_seed=int(gz["sample_seed"]))
              for ep in range(int(gz["epochs"]))]
    assert rel_err(np.array(losses), gz["epoch_losses"]) < 1e-3, (losses, gz["epoch_losses"])
    res = T.eval_neg_all(model, int(gz["batch"]), dit, dit, adj, I, False)
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        assert np.allclose(res[k], gz["eval/" + k], rtol=3e-2, atol=3e-3), (k, res[k], gz["eval/" + k])
    sd = model.state_dict()
    assert rel_err(sd["uEmbd.weight"].cpu().numpy(), gz["sd1/uEmbd.weight"]) < 2e-2


# ------------------------------------------------------------------------------------------------
# BASELINE.json full sizes (configs 2-4): one real training step vs the oracle, and tc == exact evaluation
# ------------------------------------------------------------------------------------------------
FULL = {"gowalla": (29858, 40981, 1027370), "yelp2018": (31668, 38048, 1561406), "amazon-book": (52643, 91599, 2984108)}


@pytest.mark.parametrize("shape", ["gowalla", "yelp2018", "amazon-book"])
def test_full_size_training_step_and_eval(shape):
    """At the BASELINE shapes: (1) one PairSampling step with dropout 0.2 -- loss and every gradient vs the fp32 oracle port
    on the same sampled batch and the same Philox masks; (2) graph invariants; (3) AllNeg: the tensor-core path returns exactly
    the ids/scores of the exact path for EVERY evaluated user, and metric sums agree."""
    from ngacf_b200 import hostdata
    from ngacf_b200.data import Interactions
    from ngacf_b200.evaluate import AllNegEvaluator
    from ngacf_b200.loss import BPRLoss
    from ngacf_b200.model import SPUIGACF
    U, I, E = FULL[shape]
    u, i = hostdata.synth_bipartite(U, I, E, 0)
    (tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
    torch.manual_seed(2019)
    model = SPUIGACF(U, I, 64, [64, 64], 0.2)
    with torch.no_grad():
        model.uEmbd.weight.mul_(10.0)
        model.iEmbd.weight.mul_(10.0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).train()
    model.drop_seed, model._call = 31, 0
    adj = torch.from_numpy(np.stack([u, i])).to(DEV)
    g = model.graph_for(adj)
    # (2) graph invariants at full size
    assert g.E == E
    rp, cp = g.rowptr.cpu().numpy(), g.colptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == E and (np.diff(rp) >= 1).all() and cp[-1] == E and (np.diff(cp) >= 0).all()
    ci = g.colidx.cpu().numpy()
    assert np.array_equal(np.bincount(ci, minlength=I), np.diff(cp))
    assert np.array_equal(np.sort(g.perm.cpu().numpy()), np.arange(E))
    t = g.tasks.cpu().numpy()
    assert int((t[:, 2] - t[:, 1]).sum()) == 2 * E and int((t[:, 2] - t[:, 1]).max()) <= 128
    # (1) one step vs the oracle
    it = port.build_interactions(U, I, tu, ti, su, si)
    B = 2048
    lo = 4096
    users, pos, neg = port.sample_pairs(it, lo, lo + B, 7, 0)
    ut, pt, nt = (torch.from_numpy(x).to(DEV) for x in (users, pos, neg))
    loss = BPRLoss()(model(ut, pt, g), model(ut, nt, g))
    loss.backward()
    pg = port.build_graph(np.stack([u, i]), U, I)
    p = port.params_from_state_dict(sd, torch.float32)
    mp, mn = port.dropout_masks(pg, 31, 0, 0.2), port.dropout_masks(pg, 31, 1, 0.2)
    ref_loss, ref_grads, _, _ = port.train_step_grads(p, pg, users, pos, neg, mp, mn, 0.2)
    assert abs(loss.item() - float(ref_loss)) < 1e-4 * abs(float(ref_loss))
    ref_sd = port.state_dict_from_params(ref_grads)
    for k, v in model.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), ref_sd[k].numpy()) < 2e-4, k
    # (3) evaluation: tensor-core path == exact path, every user
    dit = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
    model.eval()
    with torch.no_grad():
        Z = model.propagate(g)
        ev_tc, ev_ex = AllNegEvaluator(dit, "tc"), AllNegEvaluator(dit, "exact")
        r_tc, r_ex = ev_tc(Z), ev_ex(Z)
    assert torch.equal(ev_tc.top_ids, ev_ex.top_ids) and torch.equal(ev_tc.top_scores, ev_ex.top_scores)
    assert ev_tc.n_fallback <= dit.eval_users.numel() // 100
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        assert np.array_equal(r_tc[k], r_ex[k])


def test_edge_cases_small_batches_and_single_edge_users():
    """B=1, a batch made of one repeated user, users with exactly one edge, an item adjacent to every user."""
    from ngacf_b200.loss import BPRLoss
    from ngacf_b200.model import SPUIGACF
    U, I = 40, 30
    u = np.concatenate([np.arange(U), np.arange(U), np.arange(10)])
    i = np.concatenate([np.zeros(U, np.int64), (np.arange(U) % 7) + 1, np.full(10, 20)])      # item 0 touches every user; users >= 10 have 2 edges
    u, i = np.concatenate([u, [39]]), np.concatenate([i, [29]])
    g = port.build_graph(np.stack([u, i]), U, I)
    p64 = port.init_params(U, I, 11, torch.float64)
    p64["uEmbd"] *= 30
    p64["iEmbd"] *= 30
    model = SPUIGACF(U, I, 64, [64, 64], 0.0)
    model.load_state_dict({k: v.float() for k, v in port.state_dict_from_params(p64).items()})
    model = model.to(DEV).train()
    adj = torch.from_numpy(np.stack([u, i])).to(DEV)
    for users, pos, neg in (([3], [0], [5]), ([7] * 9, [0, 1, 0, 1, 0, 1, 0, 1, 0], [9, 9, 9, 11, 11, 12, 13, 14, 15])):
        model.zero_grad()
        ut, pt, nt = (torch.tensor(x, device=DEV) for x in (users, pos, neg))
        loss = BPRLoss()(model(ut, pt, adj), model(ut, nt, adj))
        loss.backward()
        ref_loss, ref_grads, _, _ = port.train_step_grads(p64, g, np.array(users), np.array(pos), np.array(neg))
        assert abs(loss.item() - float(ref_loss)) < 1e-4 * abs(float(ref_loss))
        ref_sd = port.state_dict_from_params(ref_grads)
        for k, v in model.named_parameters():
            assert rel_err(v.grad.cpu().numpy(), ref_sd[k].numpy()) < 1e-4, k


def test_multi_gacf_three_stage_vs_reference(golden):
    """SPUIMultiGACF (3 stages: 8 heads, 8 heads, out_att) on the same kernels -- GPU vs the reference's fp64 run with the
    Philox masks injected (tests/golden/multi_fwd_bwd_small.npz); also through the fused trainer for one step."""
    from ngacf_b200.model import SPUIMultiGACF
    gz = golden("multi_fwd_bwd_small")
    U, I = int(gz["U"]), int(gz["I"])
    model = SPUIMultiGACF(U, I, 64, [64, 64], float(gz["drop_p"]))
    model.load_state_dict(sd_from(gz, "sd/"))
    model = model.to(DEV).train()
    model.drop_seed, model._call = int(gz["drop_seed"]), int(gz["drop_call"])
    adj = torch.from_numpy(np.stack([gz["edge_u"], gz["edge_i"]])).to(DEV)
    sc = model(torch.from_numpy(gz["users"]).to(DEV), torch.from_numpy(gz["items"]).to(DEV), adj)
    assert rel_err(sc.detach().cpu().numpy(), gz["scores_drop_f64"]) < 1e-4
    (sc * torch.from_numpy(gz["w"]).float().to(DEV)).sum().backward()
    assert len(list(model.parameters())) == 2 + 3 * 17
    for k, v in model.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), gz["grad_drop_f64/" + k]) < 1e-4, k


# ------------------------------------------------------------------------------------------------
# dense stage transforms: tensor-core (tcgen05 3xTF32 / bf16 three-term) and CUDA-core kernels against an fp64 product
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dense", ["tc", "ffma"])
def test_dense_transforms_vs_fp64(dense):
    """h, s, dX, dW, da of ngacf_transform_fwd / _bwd / _bwd_dx for ragged, single-row and multi-tile shapes, both heads layouts,
    with and without ELU / feature dropout; the kernel family is chosen once per process (NGACF_DENSE), hence the subprocess."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NGACF_DENSE=dense)
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "exp_dense_tc.py"), "--check"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


# ------------------------------------------------------------------------------------------------
# NegSampling training / SampledNeg evaluation
# ------------------------------------------------------------------------------------------------
def test_sample_negs_bit_exact():
    """K distinct negatives per row from the specified Philox stream: kernel == oracle, for the train (K=4) and eval (K=99) tags,
    with and without the device-resident row/epoch offsets."""
    from ngacf_b200 import ops
    it, dit, _ = make_interactions(300, 517, 9000, 5)
    ap = port.AllPositives(it)
    assert np.array_equal(dit.all_ptr.cpu().numpy(), ap.ptr) and np.array_equal(dit.all_rank.cpu().numpy(), ap.rank)
    for K, tag, rows_u, rows_i, hu, hi in ((4, ops.NEG_TAG_TRAIN, dit.train_rows_user, dit.train_rows_item, it.train_rows_user, None),
                                            (99, ops.NEG_TAG_EVAL, dit.test_rows_user, dit.test_rows_item, None, None)):
        hu = rows_u.cpu().numpy()
        hi = rows_i.cpu().numpy()
        lo, hi_row = 7, min(7 + 600, hu.shape[0])
        n = hi_row - lo
        pu = torch.zeros(n * (K + 1), dtype=torch.int64, device=DEV)
        pi = torch.zeros_like(pu)
        ops.sample_negs(dit, rows_u, rows_i, lo, hi_row, 0xABCDEF0123, 3, K, tag, pu, pi)
        eu, ei = port.sample_negs(it, ap, hu, hi, lo, hi_row, 0xABCDEF0123, 3, K, tag)
        assert np.array_equal(pu.cpu().numpy().reshape(n, K + 1), eu) and np.array_equal(pi.cpu().numpy().reshape(n, K + 1), ei)
        # device counters: row_begin + row_dev[0], epoch + row_dev[1]
        rd = torch.tensor([lo, 3], dtype=torch.int64, device=DEV)
        pi2 = torch.zeros_like(pi)
        ops.sample_negs(dit, rows_u, rows_i, 0, n, 0xABCDEF0123, 0, K, tag, pu, pi2, rd)
        assert torch.equal(pi, pi2)
        # column-major layout (what the captured NegSampling step uses): element (row, column) at column*stride + row
        pu3, pi3 = torch.full(((K + 1) * (n + 3),), -7, dtype=torch.int64, device=DEV), torch.full(((K + 1) * (n + 3),), -7, dtype=torch.int64, device=DEV)
        ops.sample_negs(dit, rows_u, rows_i, lo, hi_row, 0xABCDEF0123, 3, K, tag, pu3, pi3, col_stride=n + 3)
        assert np.array_equal(pi3.cpu().numpy().reshape(K + 1, n + 3)[:, :n].T, ei) and np.array_equal(pu3.cpu().numpy().reshape(K + 1, n + 3)[:, :n].T, eu)
        assert (pi3.cpu().numpy().reshape(K + 1, n + 3)[:, n:] == -7).all()


def test_bce_loss_and_rank_metrics():
    from ngacf_b200 import ops
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(5 * 777) * 4).astype(np.float32)
    xt = torch.from_numpy(x).to(DEV)
    loss = torch.zeros((), device=DEV)
    dx = torch.zeros_like(xt)
    ops.bce_logits_loss(xt, 5, loss, dx)
    y = torch.zeros(777, 5, dtype=torch.float64)
    y[:, 0] = 1
    l_ref, d_ref = port.bce_logits(torch.from_numpy(x).double(), y.reshape(-1))
    assert abs(loss.item() - float(l_ref)) / float(l_ref) < 1e-5
    assert rel_err(dx.cpu().numpy(), d_ref.numpy()) < 1e-5
    ops.bce_logits_loss(xt.reshape(777, 5).t().contiguous().reshape(-1), -777, loss, None)      # column-major: the first 777 are the positives
    assert abs(loss.item() - float(l_ref)) / float(l_ref) < 1e-5
    sc = rng.standard_normal((333, 100)).astype(np.float32)
    sc[5, 7] = sc[5, 0]            # a tie with the positive does not outrank it
    sums = torch.zeros(2, dtype=torch.float64, device=DEV)
    ops.rank_metrics(torch.from_numpy(sc).to(DEV).reshape(-1), 100, 10, sums)
    hr, nd = port.rank_metrics(sc, 10)
    got = (sums / 333).tolist()
    assert abs(got[0] - hr) < 1e-12 and abs(got[1] - nd) < 1e-12


@pytest.mark.parametrize("fused", [False, True])
def test_neg_sampling_epochs_vs_reference(golden, fused):
    """Two NegSampling epochs (BCE on 1 + 4 sampled items per train row, dropout 0.2, Adam; GPU sampler + GPU Philox masks) vs the
    reference's train_neg_sample with the same samples and masks injected, then SampledNeg HR/NDCG@10 vs its eval_neg_sample
    (tests/golden/neg_sampling_small.npz)."""
    import train_eval_Gowalla as T
    from ngacf_b200.data import Interactions
    from ngacf_b200.optim import FusedAdam
    gz = golden("neg_sampling_small")
    model = make_model(gz, "sd0/", float(gz["droprate"]))
    U, I = int(gz["U"]), int(gz["I"])
    dit = Interactions.from_arrays(U, I, gz["train_u"], gz["train_i"], gz["test_u"], gz["test_i"], device=DEV)
    adj = torch.from_numpy(np.stack([np.concatenate([gz["train_u"], gz["test_u"]]), np.concatenate([gz["train_i"], gz["test_i"]])]))
    optim = FusedAdam(model.parameters(), lr=float(gz["lr"]), weight_decay=float(gz["wd"]))
    model.drop_seed = int(gz["drop_seed"])
    lossfn = torch.nn.BCEWithLogitsLoss()
    losses = [T.train_neg_sample(model, int(gz["batch"]), dit, dit, adj, optim, lossfn, False, epoch=ep, sample_seed=int(gz["sample_seed"]),
                                 fused=fused) for ep in range(int(gz["epochs"]))]
    assert rel_err(np.array(losses), gz["epoch_losses"]) < 1e-4
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    for k in sd:
        assert rel_err(sd[k], gz["sd1/" + k]) < 5e-3, k
    # evaluation on the REFERENCE's trained weights: the same candidates, scores equal to 1e-4 -> identical ranks
    ref = make_model(gz, "sd1/", float(gz["droprate"]))
    hr, nd = T.eval_neg_sample(ref, 512, dit, dit, adj, int(gz["top_k"]), False, seed=int(gz["eval_seed"]))
    assert abs(hr - float(gz["eval/hr"])) < 1e-12 and abs(nd - float(gz["eval/ndcg"])) < 1e-9
    hr2, nd2 = T.eval_neg_sample(model, 512, dit, dit, adj, int(gz["top_k"]), False, seed=int(gz["eval_seed"]))
    assert abs(hr2 - float(gz["eval/hr"])) < 0.05 and abs(nd2 - float(gz["eval/ndcg"])) < 0.05


def test_neg_sampling_reference_frames():
    """train_neg_sample / eval_neg_sample accept the reference's pandas structures (positives_negtives sets + (userId,itemId) rows)."""
    import pandas as pd
    import train_eval_Gowalla as T
    from ngacf_b200.data import Interactions
    U, I, E = 90, 140, 1500
    u, i = port.synth_bipartite(U, I, E, 8)
    order = np.lexsort((np.arange(u.shape[0]), u))
    u, i = u[order], i[order]
    first = np.r_[True, u[1:] != u[:-1]]
    su, si, tu, ti = u[first], i[first], u[~first], i[~first]
    train_df = pd.DataFrame(dict(userId=tu, itemId=ti, rating=1))
    test_df = pd.DataFrame(dict(userId=su, itemId=si, rating=1))
    pool = set(np.unique(i).tolist())
    pos = pd.DataFrame(dict(userId=np.arange(U))).assign(positive_items=[set(i[u == k].tolist()) for k in range(U)])
    pos["negative_items"] = pos["positive_items"].apply(lambda s: pool - s)
    a = Interactions.from_negsampling_frames(U, I, pos, train_df=train_df, device=DEV)
    b = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
    for k in ("train_rows_user", "train_rows_item", "all_ptr", "all_rank", "pool"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    c = Interactions.from_negsampling_frames(U, I, pos, test_df=test_df, device=DEV)
    for k in ("test_rows_user", "test_rows_item", "all_ptr", "all_rank", "pool"):
        assert torch.equal(getattr(c, k), getattr(b, k)), k


def test_score_pairs_bwd_accumulates_over_calls():
    """One call on all pairs == the same pairs split over several accumulating calls (what the NegSampling step does per column),
    up to the fp32 summation order of rows that occur in more than one call."""
    from ngacf_b200 import ops
    rng = np.random.default_rng(4)
    U, I, n = 50, 70, 1500
    Z = torch.from_numpy(rng.standard_normal((U + I, 64)).astype(np.float32)).to(DEV)
    users = torch.from_numpy(rng.integers(0, U, n)).to(DEV)
    items = torch.from_numpy(rng.integers(0, I, n)).to(DEV)
    d = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV)
    G1 = torch.zeros_like(Z)
    ops.score_pairs_bwd(Z, U, users, items, d, G1)
    G2 = torch.zeros_like(Z)
    for j in range(3):
        sl = slice(j * 500, (j + 1) * 500)
        ops.score_pairs_bwd(Z, U, users[sl], items[sl], d[sl], G2, accumulate=j > 0)
    assert rel_err(G2.cpu().numpy(), G1.cpu().numpy()) < 1e-5


# ------------------------------------------------------------------------------------------------
# pruned last stage (round 2): the fused trainer computes the output stage for the batch rows only
# ------------------------------------------------------------------------------------------------
def _fused_trainer(U, I, u, i, tu, ti, su, si, sd, droprate, batch, prune, monkeypatch):
    from ngacf_b200.data import Interactions
    from ngacf_b200.model import SPUIGACF
    from ngacf_b200.optim import FusedAdam
    from ngacf_b200.train import FusedTrainer
    monkeypatch.setenv("NGACF_PRUNE", "1" if prune else "0")
    model = SPUIGACF(U, I, 64, [64, 64], droprate)
    model.load_state_dict(sd)
    model = model.to(DEV).train()
    model.drop_seed, model._call = 77, 0
    dit = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
    g = model.graph_for(torch.from_numpy(np.stack([u, i])).to(DEV))
    optim = FusedAdam(model.parameters(), lr=0.01, weight_decay=1e-6)
    tr = FusedTrainer(model, dit, g, batch, optim, sample_seed=5)
    assert tr.prune_last_stage == prune
    return model, tr


@pytest.mark.parametrize("droprate", [0.0, 0.2])
@pytest.mark.parametrize("U,I,E,batch", [(300, 500, 6000, 256), (3000, 5000, 200000, 2048)])
def test_pruned_last_stage_equals_full(U, I, E, batch, droprate, monkeypatch):
    """The default trainer skips every last-stage row / edge whose contribution to the step is an exact zero (ngacf_*_active).
    Against the same trainer running the full last stage (NGACF_PRUNE=0): identical losses and parameters after captured AND
    eager steps, up to the order in which the surviving terms are added (bound 2e-6 relative; long rows, heavy users in the
    batch and repeated users are all present at the larger size)."""
    from ngacf_b200 import hostdata
    from ngacf_b200.model import SPUIGACF
    u, i = hostdata.synth_bipartite(U, I, E, 3)
    (tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
    torch.manual_seed(11)
    ref = SPUIGACF(U, I, 64, [64, 64], droprate)
    with torch.no_grad():
        ref.uEmbd.weight.mul_(10.0)
        ref.iEmbd.weight.mul_(10.0)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    out = {}
    for prune in (True, False):
        model, tr = _fused_trainer(U, I, u, i, tu, ti, su, si, sd, droprate, batch, prune, monkeypatch)
        # (1) one step without the optimizer: loss and EVERY gradient, batches from the head (heavy users) and the middle of the rows
        grads = []
        for row0 in (0, (len(tu) // 2 // batch) * batch):
            tr._step_body(batch, 0, droprate, model._seed(), row0, 40 + row0 % 7, False, part="compute")
            grads.append([float(tr.loss.item())] + [p.grad.detach().cpu().numpy().copy() for p in tr.params])
        # (2) captured and eager steps with Adam
        losses = tr.run_steps(4, read_loss=True)                       # captured graph, device-resident counters
        ep = tr.train_epoch(epoch=1, max_steps=3)                       # captured steps again + bookkeeping
        tr.use_cuda_graph = False
        ep2 = tr.train_epoch(epoch=2, max_steps=2)                      # eager path (host-side counters)
        out[prune] = (grads, np.array(losses + [ep, ep2]), {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()})
    for ga, gb in zip(out[True][0], out[False][0]):
        assert abs(ga[0] - gb[0]) <= 1e-6 * abs(gb[0])
        for x, y in zip(ga[1:], gb[1:]):
            assert rel_err(x, y) < 1e-5
    assert np.isfinite(out[True][1]).all()
    assert rel_err(out[True][1], out[False][1]) < 1e-4
    for k in out[True][2]:      # nine Adam steps amplify last-bit gradient differences (update = lr * m / (sqrt(v) + eps), |g| ~ eps rows)
        assert rel_err(out[True][2][k], out[False][2][k]) < 5e-3, k


def test_optimizer_hyperparameters_follow_param_groups(monkeypatch):
    """lr / weight_decay are arguments baked into the captured step: changing optim.param_groups between epochs (an LR
    scheduler) must re-capture, not be ignored (round-1 advisor finding)."""
    from ngacf_b200 import hostdata
    from ngacf_b200.model import SPUIGACF
    U, I, E = 300, 500, 6000
    u, i = hostdata.synth_bipartite(U, I, E, 3)
    (tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
    torch.manual_seed(11)
    sd = {k: v.clone() for k, v in SPUIGACF(U, I, 64, [64, 64], 0.0).state_dict().items()}
    res = []
    for change in (False, True):
        model, tr = _fused_trainer(U, I, u, i, tu, ti, su, si, sd, 0.0, 256, True, monkeypatch)
        tr.train_epoch(0, max_steps=2)
        if change:
            tr.optim.param_groups[0]["lr"] = 0.0          # frozen: nothing may move any more
        before = model.uEmbd.weight.detach().clone()
        tr.train_epoch(1, max_steps=2)
        res.append(float((model.uEmbd.weight.detach() - before).abs().max()))
    assert res[0] > 0 and res[1] == 0.0


# ------------------------------------------------------------------------------------------------
# multi-GPU partition (round 2), checked on ONE GPU: every rank of the user-range partition lives in this process and the
# collectives are applied to the ranks' buffers directly (ngacf_b200.dist.LocalCluster); the NCCL run only swaps the transport
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world,droprate,prop_parallel", [(2, 0.2, True), (2, 0.2, False), (3, 0.0, False), (4, 0.2, True), (4, 0.0, False),
                                                          (6, 0.2, True)])
def test_sharded_partition_equals_single_gpu(world, droprate, prop_parallel, monkeypatch):
    """Users range-partitioned by edge count, item rows owned by range, item rows all-gathered and item partials
    reduce-scattered per stage, optionally with the pos and the neg propagation on two groups of GPUs
    (prop_parallel): loss and EVERY gradient of one step equal the single-GPU step's (same batch, same
    Philox dropout streams, 2e-5 relative), and the losses of further Adam steps follow the single-GPU run."""
    from ngacf_b200 import hostdata
    from ngacf_b200.data import Interactions
    from ngacf_b200.dist import LocalCluster, ShardedTrainer, ShardPlan, Topology
    from ngacf_b200.model import SPUIGACF
    from ngacf_b200.optim import FusedAdam
    U, I, E, batch = 1500, 2501, 60000, 512          # I is not a multiple of the world size: the last item range is ragged
    u, i = hostdata.synth_bipartite(U, I, E, 5)
    (tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
    torch.manual_seed(3)
    ref = SPUIGACF(U, I, 64, [64, 64], droprate)
    with torch.no_grad():
        ref.uEmbd.weight.mul_(10.0)
        ref.iEmbd.weight.mul_(10.0)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    model, tr = _fused_trainer(U, I, u, i, tu, ti, su, si, sd, droprate, batch, False, monkeypatch)
    row0 = 0
    tr._step_body(batch, 0, droprate, model._seed(), row0, 0, False, part="compute")
    ref_loss = float(tr.loss.item())
    ref_grads = {n: p.grad.detach().cpu().numpy().copy() for n, p in model.named_parameters()}
    model2, tr2 = _fused_trainer(U, I, u, i, tu, ti, su, si, sd, droprate, batch, False, monkeypatch)
    ref_losses = tr2.run_steps(4, read_loss=True)
    # the partition
    topo = Topology(0, world, prop_parallel)
    assert topo.prop_parallel == prop_parallel and topo.G == (world // 2 if prop_parallel else world)
    plan = ShardPlan(u, i, U, I, topo.G)
    assert plan.ub[0] == 0 and plan.ub[-1] == U and plan.eb[-1] == E and plan.I_pad >= I and plan.chunk * (topo.G - 1) < I
    trainers = []
    for r in range(world):
        mr = SPUIGACF(U, I, 64, [64, 64], droprate)
        mr.load_state_dict(sd)
        mr = mr.to(DEV).train()
        mr.drop_seed, mr._call = 77, 0
        dit = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
        opt = FusedAdam(mr.parameters(), lr=0.01, weight_decay=1e-6)
        trainers.append(ShardedTrainer(mr, dit, u, i, batch, opt, sample_seed=5, use_cuda_graph=False, transport=object(), rank=r, world=world,
                                       plan=plan, prop_parallel=prop_parallel))
    cluster = LocalCluster(trainers)
    losses = cluster.run_steps(1, read_loss=True)
    assert abs(losses[0] - ref_loss) <= 2e-6 * abs(ref_loss)
    got = cluster.gather_grads()
    for n_, gref in ref_grads.items():
        assert rel_err(got[n_].cpu().numpy(), gref) < 2e-5, n_
    losses += cluster.run_steps(3, read_loss=True)
    assert rel_err(np.array(losses), np.array(ref_losses)) < 1e-4
    sd_ref = {k: v.detach().cpu().numpy() for k, v in model2.state_dict().items()}
    sd_got = cluster.gather_state()
    for k in sd_ref:
        assert rel_err(sd_got[k].cpu().numpy(), sd_ref[k]) < 5e-3, k


# ------------------------------------------------------------------------------------------------
# the drop-in CLI (run_Gowalla.py:118-160 of the reference): train, evaluate, save, resume
# ------------------------------------------------------------------------------------------------
def test_cli_run_gowalla_train_eval_save_resume(tmp_path, monkeypatch, capsys):
    """run_Gowalla.main with the reference's flags on a small synthetic dataset: two epochs straight == one epoch + `--resume_from 1`
    (checkpoint naming ckpts/{model}_{dataset}_{epoch:03d}.pkl with the reference's two keys; model, Adam state, sampler epoch and
    dropout streams all resume), the printed lines keep the reference's format, the metrics dict has its keys."""
    import run_Gowalla as R
    monkeypatch.chdir(tmp_path)
    flags = ["--dataset", "synth-tiny", "--model", "SPUIGACF", "--adj_type", "ui_mat", "--train_mode", "PairSampling", "--eval_mode", "AllNeg",
             "--lr", "0.002", "--weight_decay", "0.000001", "--droprate", "0.2", "--batch_size", "2048", "--eval_every", "1", "--save_every", "1",
             "--parallel", "False"]

    def run(extra):
        args = R.build_parser().parse_args(flags + extra)
        torch.manual_seed(args.seed)
        torch.cuda.manual_seed_all(args.seed)
        np.random.seed(args.seed)
        R.main(args)
        return capsys.readouterr().out
    out2 = run(["--epochs", "2"])
    assert "------epoch:1, train_loss:" in out2 and "metrics:" in out2 and "userNum:2000, itemNum:3000" in out2
    ck = torch.load(tmp_path / "ckpts" / "SPUIGACF_synth-tiny_002.pkl", map_location="cpu", weights_only=False)
    assert {"model", "optim"} <= set(ck) and set(ck["model"]) >= {"uEmbd.weight", "iEmbd.weight", "gat.attention_0.W_u", "gat.out_att.a"}
    straight = {k: v.clone() for k, v in ck["model"].items()}
    # second process: one epoch, then resume for the second
    (tmp_path / "ckpts" / "SPUIGACF_synth-tiny_002.pkl").unlink()
    run(["--epochs", "1"])
    out_r = run(["--epochs", "2", "--resume_from", "1"])
    assert "=> loaded checkpoint" in out_r and "------epoch:1, train_loss:" in out_r and "------epoch:0," not in out_r
    resumed = torch.load(tmp_path / "ckpts" / "SPUIGACF_synth-tiny_002.pkl", map_location="cpu", weights_only=False)["model"]
    for k in straight:
        assert rel_err(resumed[k].numpy(), straight[k].numpy()) < 1e-6, k

    def loss_of(text, epoch):
        line = [l for l in text.splitlines() if l.startswith("------epoch:%d," % epoch)][0]
        return float(line.split("train_loss:")[1].split(",")[0])
    assert abs(loss_of(out_r, 1) - loss_of(out2, 1)) < 1e-5


# ------------------------------------------------------------------------------------------------
# parity statements tightened in round 2
# ------------------------------------------------------------------------------------------------
def test_propagated_embeddings_elementwise_vs_reference(golden):
    """north_star: "embeddings ... agree within 1e-4 relative (fp32)".  ELEMENT-WISE against the reference's own fp64 propagation
    (tests/golden/embeddings_medium.npz, written by oracle/make_golden.py:case_embeddings from the unmodified reference):
    |ours - ref| <= 1e-4 |ref| + 1e-6 for every one of the (U+I) x 64 features (the absolute floor covers entries that are
    themselves ~1e-7, below fp32 resolution of the sums that form them); both dense-transform families."""
    gz = golden("embeddings_medium")
    model = make_model(gz, "sd/")
    model.eval()
    adj = torch.from_numpy(np.stack([gz["edge_u"], gz["edge_i"]])).to(DEV)
    with torch.no_grad():
        F = torch.nn.functional.elu(model.propagate(adj)).cpu().numpy().astype(np.float64)
    ref = gz["features_f64"]
    assert F.shape == ref.shape
    err = np.abs(F - ref)
    assert (err <= 1e-4 * np.abs(ref) + 1e-6).all(), float((err / (np.abs(ref) + 1e-30)).max())
    assert np.median(err / (np.abs(ref) + 1e-12)) < 2e-6           # the typical element is at fp32 round-off


def test_long_row_combine_is_deterministic_under_stress():
    """Rows longer than CHUNK are summed by whichever chunk task arrives last (fence + counter protocol, csrc/common.cuh and
    csrc/pruned_stage.cu).  The arrival order changes from launch to launch; the result must not: 2000 forward + backward
    replays of a graph with 600-chunk rows, every output compared bit for bit with the first replay."""
    from ngacf_b200 import ops
    from ngacf_b200.propagation import Propagation
    U, I = 900, 40
    rng = np.random.default_rng(0)
    u = np.repeat(np.arange(U), 30)
    i = np.concatenate([rng.choice(I, 30, replace=False) for _ in range(U)])       # every item has ~675 edges: 6 chunks
    g = cuda_graph(u, i, U, I)
    assert g.L >= I
    torch.manual_seed(0)
    from ngacf_b200.model import SPUIGACF
    model = SPUIGACF(U, I, 64, [64, 64], 0.0).to(DEV)
    with torch.no_grad():
        model.uEmbd.weight.mul_(30.0)
        model.iEmbd.weight.mul_(30.0)
    prop = Propagation(g, model.stages)
    prop.set_dropout(0.3, 5, 0)
    wt = [ops.pointer_table([p.detach() for p in st]) for st in model.gat.stage_parameters()]
    grads = [torch.zeros_like(p) for st in model.gat.stage_parameters() for p in st]
    it = iter(grads)
    gt = [ops.pointer_table([next(it) for _ in st]) for st in model.gat.stage_parameters()]
    dU, dI = torch.zeros_like(model.uEmbd.weight), torch.zeros_like(model.iEmbd.weight)
    Gin = torch.randn(U + I, 64, device=DEV)

    def once():
        Z = prop.forward(model.uEmbd.weight.detach(), model.iEmbd.weight.detach(), wt)
        prop.grad_in().copy_(Gin)
        prop.backward(prop.grad_in(), model.uEmbd.weight.detach(), model.iEmbd.weight.detach(), wt, gt, dU, dI, False)
        return [Z.clone(), dU.clone(), dI.clone()] + [x.clone() for x in grads]
    first = once()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        Z = prop.forward(model.uEmbd.weight.detach(), model.iEmbd.weight.detach(), wt)
        prop.grad_in().copy_(Gin)
        prop.backward(prop.grad_in(), model.uEmbd.weight.detach(), model.iEmbd.weight.detach(), wt, gt, dU, dI, False)
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    for _ in range(2000):
        gr.replay()
        for a, b in zip([Z, dU, dI] + grads, first):
            bad += (a != b).sum()
    assert int(bad.item()) == 0


# ------------------------------------------------------------------------------------------------
# SURVEY 8f-1, second half: SPUIGAGPCF (SpUIGAT + Laplacian-propagation layers)
# ------------------------------------------------------------------------------------------------
def test_spuigagpcf_forward_backward_vs_reference(golden):
    """SPUIGAGPCF.forward / backward against the reference's own fp64 run (tests/golden/gagpcf_small.npz: the reference class with the
    reference's buildLaplacianMat 'norm_adj' Laplacian over integer ratings).  Scores and EVERY gradient -- embeddings, attention
    parameters and the affine layers behind the Laplacian product -- to 1e-4."""
    from graphattention.SPUIGACF import SPUIGAGPCF
    gz = golden("gagpcf_small")
    U, I = int(gz["U"]), int(gz["I"])
    L = torch.sparse_coo_tensor(torch.from_numpy(np.stack([gz["lap_row"], gz["lap_col"]])), torch.from_numpy(gz["lap_val"]), (U + I, U + I))
    model = SPUIGAGPCF(U, I, L, 64, [64, 64], 0.0)
    model.load_state_dict(sd_from(gz, "sd/"))
    model = model.to(DEV).train()
    adj = torch.from_numpy(np.stack([gz["edge_u"], gz["edge_i"]])).to(DEV)
    sc = model(torch.from_numpy(gz["users"]).to(DEV), torch.from_numpy(gz["items"]).to(DEV), adj)
    assert rel_err(sc.detach().cpu().numpy(), gz["scores_f64"]) < 1e-4
    (sc * torch.from_numpy(gz["w"]).float().to(DEV)).sum().backward()
    names = [k for k, _ in model.named_parameters()]
    assert "Affinelayers.1.bias" in names and "gat.out_att.a" in names
    for k, v in model.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), gz["grad_f64/" + k]) < 1e-4, k
    # the operator itself: (L + I) X against the dense product, long rows included
    from ngacf_b200.gp import LaplacianOp
    X = torch.randn(U + I, 64, device=DEV)
    got = LaplacianOp(model.graph_for(adj), L)(X)
    want = (L.to_dense().to(DEV).double() + torch.eye(U + I, device=DEV, dtype=torch.float64)) @ X.double()
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-5


# ------------------------------------------------------------------------------------------------
# SpGraphAttentionLayer / SpGAT / SPGACF (the single-table layer, graphattention/SPGA.py:85-140,330-421)
# ------------------------------------------------------------------------------------------------
def _spgacf_masks(gz, graph):
    """the fixture's keep masks (edge masks in adj.nonzero() order) in the kernels' directed-edge order"""
    nzi = graph.nonzero_index().cpu().numpy()
    return dict(feat=[torch.from_numpy(gz["drop_feat0"].view(np.int64)).to(DEV), torch.from_numpy(gz["drop_feat1"].view(np.int64)).to(DEV)],
                edge=[torch.from_numpy(gz["drop_edge0_nz"][nzi]).to(DEV), torch.from_numpy(gz["drop_edge1_nz"][nzi]).to(DEV)])


@pytest.mark.parametrize("name", ["spgacf_small", "spgacf_selfloops_small"])
def test_spgacf_forward_backward_vs_reference(golden, name):
    """SPGACF (SpGAT over the dense N x N adjacency, without and with the diagonal) against the reference's own fp64 run: scores,
    final features and EVERY gradient to 1e-4, without dropout and with the fixture's injected keep masks (p = 0.3)."""
    from graphattention.SPGA import SPGACF, HomoGraph
    gz = golden(name)
    U, I = int(gz["U"]), int(gz["I"])
    N = U + I
    adj = torch.zeros(N, N)
    adj[torch.from_numpy(gz["row"]), torch.from_numpy(gz["col"])] = 1.0
    adj = adj.to(DEV)
    users, items = torch.from_numpy(gz["users"]).to(DEV), torch.from_numpy(gz["items"]).to(DEV)
    w = torch.from_numpy(gz["w"]).float().to(DEV)
    for tag, p_drop in (("f64", 0.0), ("drop_f64", float(gz["drop_p"]))):
        model = SPGACF(U, I, None, 64, [64, 64], p_drop)
        model.load_state_dict(sd_from(gz, "sd/"))
        model = model.to(DEV).train()
        graph = model.gat.graph_for(adj, U)
        assert graph.self_loops == bool(gz["self_loops"]) and graph.n_edges == gz["row"].shape[0]
        if p_drop > 0:
            model.gat.injected_masks = _spgacf_masks(gz, graph)
        sc = model(users, items, adj)
        assert rel_err(sc.detach().cpu().numpy(), gz["scores_" + tag]) < 1e-4, tag
        (sc * w).sum().backward()
        names = [k for k, _ in model.named_parameters()]
        assert "gat.attention_7.W" in names and "gat.out_att.a" in names and len(names) == 20
        for k, v in model.named_parameters():
            assert rel_err(v.grad.cpu().numpy(), gz["grad_%s/%s" % (tag, k)]) < 1e-4, (tag, k)
    # eval mode: the propagated features, element-wise (sparse adjacency input, inferred user count)
    model.eval()
    with torch.no_grad():
        F = model.gat(model.getFeatureMat()[2], adj.to_sparse())
    ref = gz["features_f64"]
    assert np.abs(F.cpu().numpy() - ref).max() < 1e-4 * np.abs(ref).max()
    # the same graph from (user, item) pairs, no dense matrix
    g2 = HomoGraph.from_pairs(torch.from_numpy(np.stack([gz["edge_u"], gz["edge_i"]])).to(DEV), U, I, bool(gz["self_loops"]))
    assert torch.equal(g2.rev, graph.rev) and torch.equal(g2.g.adj_idx, graph.g.adj_idx)
    # reverse map is an involution that swaps the endpoints
    rev = graph.rev.long()
    assert torch.equal(rev[rev], torch.arange(rev.numel(), device=DEV))


def test_spgat_long_rows_philox_dropout_and_refusals():
    """rows longer than one chunk (the combine path) against the port restatement, Philox dropout determinism, and the loud refusals
    (non-bipartite / asymmetric pattern, node without edges)."""
    from oracle import port
    from graphattention.SPGA import SPGACF, HomoGraph
    rng = np.random.default_rng(5)
    U2, I2 = 400, 500
    u = np.concatenate([rng.integers(0, U2, 4000), np.arange(U2), np.full(300, 3), rng.integers(0, U2, 300), rng.integers(0, U2, I2)])   # user 3 / item 7: > 128 edges
    i = np.concatenate([rng.integers(0, I2, 4000), np.arange(U2) % I2, rng.integers(0, I2, 300), np.full(300, 7), np.arange(I2)])
    g = port.build_graph(np.stack([u, i]), U2, I2)
    assert np.diff(g.rowptr).max() > 128 and np.diff(g.colptr).max() > 128
    for self_loops in (False, True):
        torch.manual_seed(3)
        model = SPGACF(U2, I2, None, 64, [64, 64], 0.0)
        with torch.no_grad():
            model.uEmbd.weight.mul_(20.0)
            model.iEmbd.weight.mul_(20.0)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        model = model.to(DEV).train()
        graph = HomoGraph.from_pairs(torch.from_numpy(np.stack([g.eu, g.ei])).to(DEV), U2, I2, self_loops)
        assert graph.g.L > 0
        users, items = torch.from_numpy(rng.integers(0, U2, 256)).to(DEV), torch.from_numpy(rng.integers(0, I2, 256)).to(DEV)
        w = torch.from_numpy(rng.standard_normal(256)).float().to(DEV)
        sc = model(users, items, graph)
        (sc * w).sum().backward()
        p = port.spgat_params_from_state_dict({k: v.double() for k, v in sd.items()})
        row, col = port.homo_edges(g, self_loops)
        F, caches = port.spgat_propagate(p, row, col)
        uu, ii, ww = users.cpu(), items.cpu(), w.cpu().double()
        want = (F[uu] * F[ii + U2]).sum(1)
        assert rel_err(sc.detach().cpu().numpy(), want.numpy()) < 1e-4
        dF = torch.zeros_like(F).index_add_(0, uu, ww[:, None] * F[ii + U2]).index_add_(0, ii + U2, ww[:, None] * F[uu])
        gr = port.spgat_propagate_backward(dF, p, row, col, caches)
        assert rel_err(model.uEmbd.weight.grad.cpu().numpy(), gr["uEmbd"].numpy()) < 1e-4
        assert rel_err(model.iEmbd.weight.grad.cpu().numpy(), gr["iEmbd"].numpy()) < 1e-4
        assert rel_err(model.gat.out_att.W.grad.cpu().numpy(), gr["stages"][1]["W"][0].numpy()) < 1e-4
        assert rel_err(model.gat.out_att.a.grad.cpu().numpy()[0], gr["stages"][1]["a"][0].numpy()) < 1e-4
        for k in range(8):
            assert rel_err(getattr(model.gat, "attention_%d" % k).W.grad.cpu().numpy(), gr["stages"][0]["W"][k].numpy()) < 1e-4
            assert rel_err(getattr(model.gat, "attention_%d" % k).a.grad.cpu().numpy()[0], gr["stages"][0]["a"][k].numpy()) < 1e-4
    # Philox dropout: same (seed, call) -> same bits; training calls advance the stream
    model.gat.dropout = 0.3
    model.gat.drop_seed, model.gat._call = 11, 0
    a = model(users, items, graph).detach().clone()
    b = model(users, items, graph).detach().clone()
    model.gat._call = 0
    c = model(users, items, graph).detach().clone()
    assert torch.equal(a, c) and not torch.equal(a, b)
    # refusals
    N = 6
    bad = torch.zeros(N, N, device=DEV)
    bad[0, 1] = bad[1, 0] = 1.0            # with userNum = 3 this is a user-user edge
    with pytest.raises(NotImplementedError):
        HomoGraph(bad, 3)
    asym = torch.zeros(N, N, device=DEV)
    asym[0, 4] = 1.0
    with pytest.raises(NotImplementedError):
        HomoGraph(asym, 3)
    lonely = torch.zeros(N, N, device=DEV)
    lonely[0, 4] = lonely[4, 0] = 1.0
    with pytest.raises(ValueError):
        HomoGraph(lonely, 3)


# ------------------------------------------------------------------------------------------------
# AllNeg tensor-core scorer: the state it keeps between calls and its escape hatch
# ------------------------------------------------------------------------------------------------
def test_eval_tc_mask_reuse_rebuild_and_fallback():
    """(1) the allowed-column matrix in the workspace is reused by the second evaluation and REBUILT when the pool changes (in-place
    edit of `in_pool`: the evaluator keys it on the tensors' versions); (2) rows whose proof fails -- here every row: all item
    embeddings equal, so every score ties -- are counted by the library, recomputed through the exact entry point when the result
    is read, and come out identical to the exact evaluator."""
    from ngacf_b200.evaluate import AllNegEvaluator
    U, I = 300, 1000
    it, dit, Zt, Fn = _eval_case(U, I, 9000, 5)
    ev_tc, ev_ex = AllNegEvaluator(dit, "tc"), AllNegEvaluator(dit, "exact")
    r1 = ev_tc(Zt)
    assert ev_tc._mask_key is not None
    key = ev_tc._mask_key
    r2 = ev_tc(Zt)                                   # second call: reuse_mask = 1
    assert ev_tc._mask_key == key
    rx = ev_ex(Zt)
    assert torch.equal(ev_tc.top_ids, ev_ex.top_ids) and ev_tc.n_fallback == 0
    for k in ("precision", "recall", "ndcg", "hit_ratio"):
        assert np.array_equal(r1[k], r2[k]) and np.allclose(r1[k], rx[k], rtol=0, atol=1e-12)
    # take the 50 best-scored items of user 0 out of the pool, in place
    best = torch.topk(torch.from_numpy(Fn[U:] @ Fn[int(dit.eval_users[0])]), 50).indices.to(DEV)
    dit.in_pool[best] = 0
    ev_tc(Zt)
    assert ev_tc._mask_key != key                    # rebuilt
    ev_ex(Zt)
    assert torch.equal(ev_tc.top_ids, ev_ex.top_ids) and torch.equal(ev_tc.top_scores, ev_ex.top_scores)
    assert not bool(torch.isin(ev_tc.top_ids[0].long(), best).any())
    # all item rows equal -> all scores of a user tie -> no margin for the proof -> every row goes through the fallback
    Zt2 = Zt.clone()
    Zt2[U:] = Zt2[U:U + 1]
    ev_tc.rank(Zt2)
    assert ev_tc._pending_fallback
    ids = ev_tc.top_ids                              # reading the result resolves the flagged rows
    assert ev_tc.n_fallback == dit.eval_users.numel() and not ev_tc._pending_fallback
    ev_ex.rank(Zt2)
    assert torch.equal(ids, ev_ex.top_ids) and torch.equal(ev_tc.top_scores, ev_ex.top_scores)


def test_eval_exact_split_equals_exact():
    """ngacf_score_topk_exact_split (the fallback for the few rows the tc path flags: item range split over CTAs + merge of the
    partial lists) returns exactly what ngacf_score_topk_exact returns -- ids, scores, ties by lowest id, -1 padding."""
    from ngacf_b200 import ops
    U, I = 1000, 3000
    it, dit, Zt, Fn = _eval_case(U, I, 40000, 3, 300)          # 300 duplicated item rows: exact score ties
    F = torch.empty_like(Zt)
    ops.final_features(Zt, F)
    for n in (1, 5, 37):
        users = dit.eval_users[torch.randperm(dit.eval_users.numel(), generator=torch.Generator().manual_seed(n))[:n].to(DEV)].contiguous()
        a_ids = torch.empty((n, 20), dtype=torch.int32, device=DEV)
        a_sc = torch.empty((n, 20), dtype=torch.float32, device=DEV)
        b_ids, b_sc = torch.full_like(a_ids, -7), torch.full_like(a_sc, -7.0)
        ops.score_topk_exact(F, U, I, users, dit, a_ids, a_sc)
        ops.score_topk_exact_split(F, U, I, users, dit, b_ids, b_sc)
        assert torch.equal(a_ids, b_ids) and torch.equal(a_sc, b_sc), n
    # a pool of 12 items: fewer than 20 candidates -> -1 padding survives the merge
    dit.in_pool[:] = 0
    dit.in_pool[torch.arange(0, I, I // 12, device=DEV)[:12]] = 1
    users = dit.eval_users[:3].contiguous()
    a_ids = torch.empty((3, 20), dtype=torch.int32, device=DEV)
    a_sc = torch.empty((3, 20), dtype=torch.float32, device=DEV)
    b_ids, b_sc = torch.full_like(a_ids, -7), torch.full_like(a_sc, -7.0)
    ops.score_topk_exact(F, U, I, users, dit, a_ids, a_sc)
    ops.score_topk_exact_split(F, U, I, users, dit, b_ids, b_sc)
    assert torch.equal(a_ids, b_ids) and torch.equal(a_sc, b_sc) and bool((a_ids[:, 12:] == -1).all())


def test_spgraphattentionlayer_standalone_64x64():
    """One SpGraphAttentionLayer(64 -> 64) called on its own (SPGA.py:375-417) against the port restatement: output, d input, dW, da;
    and a fresh dropout stream per training call."""
    from oracle import port
    from graphattention.SPGA import HomoGraph, SpGraphAttentionLayer
    U, I = 70, 110
    u, i = port.synth_bipartite(U, I, 900, 11)
    g = port.build_graph(np.stack([u, i]), U, I)
    row, col = port.homo_edges(g, True)
    graph = HomoGraph.from_pairs(torch.from_numpy(np.stack([g.eu, g.ei])).to(DEV), U, I, True)
    torch.manual_seed(4)
    layer = SpGraphAttentionLayer(64, 64, 0.0, 0.2, concat=True).to(DEV).train()
    x = (torch.randn(U + I, 64) * 0.5)
    w = torch.randn(U + I, 64)
    xd = x.to(DEV).requires_grad_(True)
    out = layer(xd, graph)
    (out * w.to(DEV)).sum().backward()
    st = dict(W=layer.W.detach().cpu().double()[None], a=layer.a.detach().cpu().double())
    c = port.spgat_stage_forward(x.double(), st, row, col)
    ref = torch.where(c["Z"] > 0, c["Z"], torch.expm1(c["Z"]))
    assert rel_err(out.detach().cpu().numpy(), ref.numpy()) < 1e-4
    G = w.double() * torch.where(c["Z"] > 0, torch.ones_like(c["Z"]), torch.exp(c["Z"]))
    dX, gs = port.spgat_stage_backward(G, c, st, row, col)
    assert rel_err(xd.grad.cpu().numpy(), dX.numpy()) < 1e-4
    assert rel_err(layer.W.grad.cpu().numpy(), gs["W"][0].numpy()) < 1e-4
    assert rel_err(layer.a.grad.cpu().numpy(), gs["a"].numpy()) < 1e-4
    layer.p = 0.4
    a, b = layer(xd, graph).detach(), layer(xd, graph).detach()
    assert not torch.equal(a, b) and layer._call == 2


def test_eval_tc_segmented_user_shard_equals_exact():
    """A small user shard of a large item set (a rank of a multi-GPU evaluation): the tensor-core scorer splits the item tiles into
    S = 4 segments per user block, i.e. eight lists per user sharing one threshold -- ids and scores must equal the exact path."""
    from ngacf_b200.evaluate import AllNegEvaluator
    U, I = 2000, 12000                      # 94 item tiles -> S = 4 for 3 user blocks
    it, dit, Zt, Fn = _eval_case(U, I, 60000, 9, 500)
    users = dit.eval_users[:300].contiguous()
    ev_tc, ev_ex = AllNegEvaluator(dit, "tc", users=users), AllNegEvaluator(dit, "exact", users=users)
    ev_tc.rank(Zt)
    ev_ex.rank(Zt)
    assert torch.equal(ev_tc.top_ids, ev_ex.top_ids) and torch.equal(ev_tc.top_scores, ev_ex.top_scores)
    assert ev_tc.n_fallback <= 3


def test_eval_tc_log_overflow_goes_to_exact_fallback():
    """Adversarial score order: every item scores higher than all items before it, so every accumulator passes every threshold and a
    list's survivor log (384 entries) overflows after six tiles.  The overflowing lists must be flagged (never truncated silently)
    and the rows come back exact through the fallback."""
    from ngacf_b200.evaluate import AllNegEvaluator
    U, I = 200, 2500
    it, dit, Zt, Fn = _eval_case(U, I, 6000, 12)
    v = torch.randn(64, device=DEV).abs() + 0.1
    Zt = Zt.clone()
    Zt[:U] = v                                                   # every user = v (ELU keeps positive values)
    Zt[U:] = v[None, :] * torch.linspace(0.5, 3.0, I, device=DEV)[:, None]      # item i = c_i v with c_i increasing
    ev_tc, ev_ex = AllNegEvaluator(dit, "tc"), AllNegEvaluator(dit, "exact")
    ev_tc.rank(Zt)
    ev_ex.rank(Zt)
    assert torch.equal(ev_tc.top_ids, ev_ex.top_ids) and torch.equal(ev_tc.top_scores, ev_ex.top_scores)
    assert ev_tc.n_fallback == dit.eval_users.numel()           # every row overflowed
