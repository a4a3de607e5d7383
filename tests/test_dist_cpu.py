"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in ngacf_b200/dist.py: row slicing of a
data-parallel step, the flat-gradient all-reduce (sum of per-replica mean-loss gradients = the reference's
loss.backward(ones(ndev)), train_eval_Gowalla.py:137) and the sharded AllNeg evaluation merge.  The per-rank
compute is the oracle port (test infrastructure): the CUDA kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import port


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    U, I, E = 80, 120, 1600
    u, i = port.synth_bipartite(U, I, E, 4)
    (tu, ti), (su, si) = port.split_train_test(u, i, U, 5)
    g = port.build_graph(np.stack([u, i]), U, I)
    it = port.build_interactions(U, I, tu, ti, su, si)
    p = port.init_params(U, I, 1, torch.float64)
    p["uEmbd"] *= 20
    p["iEmbd"] *= 20
    return g, it, p


def _worker(rank, world, port_no, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngacf_b200 import dist as nd
    assert nd.world_info() == (rank, world)
    g, it, p = _case()
    B = 64
    n = it.train_rows_user.shape[0]
    lo, hi = nd.rank_rows(128, B, rank, world, n)
    users, pos, neg = port.sample_pairs(it, lo, hi, 3, 0)
    loss, grads, _, _ = port.train_step_grads(p, g, users, pos, neg)
    tensors = port.flat_tensors(grads)
    flat, views = nd.flat_views([torch.zeros_like(t) for t in tensors])
    for v, t in zip(views, tensors):
        v.copy_(t)
    nd.allreduce_sums(flat)
    # sharded evaluation: each rank ranks its slice of the evaluable users, only 16 sums are merged
    F, _ = port.propagate(p, g)
    ev_users, divisor = port.eval_users(it)
    mine = nd.shard_eval_users(torch.from_numpy(ev_users), rank, world).numpy()
    sums = torch.zeros(16, dtype=torch.float64)
    Fn = F.numpy().astype(np.float32)
    for uu in mine:
        sc = port.dot64_tree(np.broadcast_to(Fn[uu], Fn[g.U:].shape), Fn[g.U:])
        top = port.topk_allneg(sc, it, int(uu))
        m = port.metrics_from_hits(port.hits_for(top, it, int(uu)), int(it.test_ptr[uu + 1] - it.test_ptr[uu]))
        sums += torch.from_numpy(np.concatenate([m["precision"], m["recall"], m["ndcg"], m["hit_ratio"]]))
    nd.allreduce_sums(sums)
    if rank == 0:
        torch.save(dict(flat=flat, sums=sums / divisor, loss=float(loss), rows=(lo, hi)), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_step_and_sharded_eval(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    g, it, p = _case()
    B, n = 64, it.train_rows_user.shape[0]
    from ngacf_b200.dist import rank_rows, shard_range
    total = None
    for r in range(2):
        lo, hi = rank_rows(128, B, r, 2, n)
        assert (lo, hi) == (128 + r * B, 128 + (r + 1) * B)
        users, pos, neg = port.sample_pairs(it, lo, hi, 3, 0)
        _, grads, _, _ = port.train_step_grads(p, g, users, pos, neg)
        flat = torch.cat([t.reshape(-1) for t in port.flat_tensors(grads)])
        total = flat if total is None else total + flat
    assert torch.allclose(got["flat"], total, rtol=1e-12, atol=1e-15)
    F, _ = port.propagate(p, g)
    res, _, _ = port.eval_neg_all(F.numpy().astype(np.float32), it)
    ref = np.concatenate([res["precision"], res["recall"], res["ndcg"], res["hit_ratio"]])
    assert np.allclose(got["sums"].numpy(), ref, rtol=1e-12)
    # shard_range covers [0,n) exactly once, balanced
    for n_items in (0, 1, 7, 64, 29854):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n_items, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n_items
            assert all(parts[k][1] == parts[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    # tail step: rows are clipped, a rank may get none
    assert rank_rows(n - 10, B, 0, 2, n) == (n - 10, n) and rank_rows(n - 10, B, 1, 2, n) == (n, n)


# ------------------------------------------------------------------------------------------------
# round 2: the partition plan, the two-level topology and the scoped in-place collectives (gloo, world 4)
# ------------------------------------------------------------------------------------------------
def test_shard_plan_and_topology():
    from ngacf_b200.dist import ShardPlan, Topology
    U, I, E = 300, 501, 9000
    u, i = port.synth_bipartite(U, I, E, 6)
    g = port.build_graph(np.stack([u, i]), U, I)
    for world in (1, 2, 3, 8):
        P = ShardPlan(u, i, U, I, world)
        assert P.E == g.E and P.ub[0] == 0 and P.ub[-1] == U and P.eb[0] == 0 and P.eb[-1] == g.E
        assert all(P.ub[r] <= P.ub[r + 1] for r in range(world))
        # users balanced by EDGE count: no rank holds more than its share plus one user's edges
        deg_max = int(np.diff(g.rowptr).max())
        for r in range(world):
            lo, hi = P.edges(r)
            assert hi - lo <= g.E // world + deg_max + 1
            eu, ei = P.local_edges(r)
            ulo, uhi = P.users(r)
            assert eu.size == hi - lo and (eu.size == 0 or (eu.min() >= ulo and eu.max() < uhi))
            assert np.array_equal(ei, g.colidx[lo:hi])          # the own users' CSR rows = contiguous GLOBAL edge ids
        # item ranges: equal chunks (NCCL counts), the last one ragged, padded rows beyond I
        assert P.chunk * world == P.I_pad >= I and P.items(0) == (0, min(I, P.chunk))
        covered = np.zeros(I, np.int32)
        for r in range(world):
            a, b = P.items(r)
            covered[a:b] += 1
        assert (covered == 1).all()
    # topology: two groups of G with an even world, one group otherwise
    t = Topology(5, 8)
    assert t.prop_parallel and t.G == 4 and t.props == [1] and t.sub_rank == 1
    assert t.members("sub") == [4, 5, 6, 7] and t.members("pair") == [1, 5] and t.members("world") == list(range(8))
    assert t.rank_in("sub") == 1 and t.rank_in("pair") == 1
    t = Topology(2, 3)
    assert not t.prop_parallel and t.G == 3 and t.props == [0, 1] and t.members("sub") == [0, 1, 2] and t.members("pair") == [2]
    t = Topology(1, 2)
    assert t.prop_parallel and t.G == 1 and t.props == [1] and t.members("sub") == [1] and t.members("pair") == [0, 1]
    assert not Topology(1, 2, prop_parallel=False).prop_parallel


def _collective_worker(rank, world, port_no, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngacf_b200.dist import NcclTransport, Topology
    res = {}
    for pp in (True, False):
        topo = Topology(rank, world, pp)
        tp = NcclTransport(topo)
        n, chunk = len(topo.members("sub")), 3
        r = topo.rank_in("sub")
        # all-gather: only the own rows are valid beforehand
        t = torch.full((n * chunk, 2), -1.0)
        t[r * chunk:(r + 1) * chunk] = float(rank)
        tp.all_gather_rows([t], chunk, "sub")
        want = torch.cat([torch.full((chunk, 2), float(m)) for m in topo.members("sub")])
        ok_ag = bool(torch.equal(t, want))
        # reduce-scatter: the own rows end up with the scope's sum
        p = torch.arange(n * chunk * 2, dtype=torch.float64).reshape(n * chunk, 2) * (rank + 1)
        tp.reduce_scatter_rows([p], chunk, "sub")
        base = torch.arange(n * chunk * 2, dtype=torch.float64).reshape(n * chunk, 2)
        ok_rs = bool(torch.equal(p[r * chunk:(r + 1) * chunk], base[r * chunk:(r + 1) * chunk] * sum(m + 1 for m in topo.members("sub"))))
        a = torch.tensor([float(rank + 1)])
        tp.all_reduce([a], "pair")
        ok_pair = float(a) == float(sum(m + 1 for m in topo.members("pair")))
        w = torch.tensor([1.0])
        tp.all_reduce([w], "world")
        res[pp] = (ok_ag, ok_rs, ok_pair, float(w) == world)
    torch.save(res, out + str(rank))
    dist.barrier()
    dist.destroy_process_group()


def test_scoped_collectives_four_ranks(tmp_path):
    out = str(tmp_path / "r")
    mp.spawn(_collective_worker, args=(4, _free_port(), out), nprocs=4, join=True)
    for rank in range(4):
        res = torch.load(out + str(rank), weights_only=False)
        assert all(all(v) for v in res.values()), (rank, res)
