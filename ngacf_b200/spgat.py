"""SpGraphAttentionLayer / SpGAT / SPGACF (graphattention/SPGA.py:85-140,330-421 of the reference): the single-table
cousin of the bipartite layer.  One ``W`` for every node, ``a = [a_src | a_dst]``, DIRECTED logit

    e(n -> m) = exp(-LeakyReLU(a_src . h[n] + a_dst . h[m])),     out[n] = sum_m dropout(e) h[m] / sum_m e        (:386-409)

(row-normalised only, no residual), over the nonzeros of the N x N adjacency the caller passes as ``adj`` / ``mask``.  SPGACF builds
that adjacency from the user-item graph, so its pattern is the symmetric bipartite one -- with or without the diagonal -- which is
exactly the unified adjacency ``ngacf_graph_build`` already lays out; ``HomoGraph`` validates that and adds the reverse-edge map the
backward needs.  A general (non-bipartite) pattern is refused loudly.

Kernels: ``ngacf_transform_fwd`` (with W_u = W_i = W) for h, ``ngacf_node_logits`` for (p, q), ``ngacf_spgat_aggregate_fwd`` for the
aggregation; backward ``ngacf_spgat_bwd`` (row pass + column pass, closed form, no atomics), ``ngacf_node_logits_bwd`` for da and
``ngacf_transform_bwd`` for dW / dX.  The reference's backward materialises a dense N x N gradient
(SpecialSpmmFunction.backward, SPGA.py:436-440).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .graph import BipartiteGraph
from .ops import D
from .propagation import ScoreFn

ALPHA = 0.2
STAGES = ((8, 8), (1, 64))
_DA_BLOCKS = 64


class HomoGraph:
    """The N x N adjacency of SpGAT as {unified bipartite adjacency, self-loop flag, reverse-edge map}.
    adj: dense or sparse (N,N) tensor (only the pattern of its nonzeros is used, like ``adj.nonzero()`` at SPGA.py:379);
    ``from_pairs`` builds it from (user,item) pairs without the dense matrix; userNum: number of user rows (inferred from the pattern when None)."""

    def __init__(self, adj: torch.Tensor, userNum=None, self_loops=None):
        _lib.require_cuda(adj)
        dev = adj.device
        if adj.dim() != 2 or adj.shape[0] != adj.shape[1]:
            raise ValueError("adj must be (N,N); got %s" % (tuple(adj.shape),))
        N = adj.shape[0]
        if adj.is_sparse:
            a = adj.coalesce()
            nzv = a.values() != 0
            r, c = a.indices()[0][nzv], a.indices()[1][nzv]
        else:
            nz = adj.nonzero()
            r, c = nz[:, 0], nz[:, 1]
        on = r == c
        n_diag = int(on.sum())
        if n_diag not in (0, N):
            raise NotImplementedError("the diagonal of adj must be empty or full (%d of %d entries present)" % (n_diag, N))
        ro, co = r[~on], c[~on]
        if userNum is None:
            up = ro < co
            if not bool(up.any()):
                raise ValueError("adj has no off-diagonal entry")
            userNum = int(co[up].min())
        U = int(userNum)
        up = (ro < U) & (co >= U)
        lo = (ro >= U) & (co < U)
        if bool((~(up | lo)).any()):
            raise NotImplementedError("adj has user-user or item-item entries: only the bipartite user-item pattern SPGACF builds is supported")
        ku = torch.sort(ro[up] * N + co[up]).values
        kl = torch.sort(co[lo] * N + ro[lo]).values
        if ku.numel() != kl.numel() or not bool((ku == kl).all()):
            raise NotImplementedError("adj is not symmetric")
        self._init(BipartiteGraph(torch.stack([ro[up], co[up] - U]), U, N - U, check_users=False), n_diag == N if self_loops is None else bool(self_loops))

    @classmethod
    def from_pairs(cls, indices: torch.Tensor, userNum: int, itemNum: int, self_loops: bool):
        self = cls.__new__(cls)
        self._init(BipartiteGraph(indices, userNum, itemNum, check_users=False), self_loops)
        return self

    def _init(self, g: BipartiteGraph, self_loops: bool):
        self.g, self.self_loops = g, bool(self_loops)
        self.U, self.I, self.N, self.E, self.device = g.U, g.I, g.N, g.E, g.device
        E = g.E
        self.n_adj = 2 * E
        self.n_edges = 2 * E + (g.N if self.self_loops else 0)          # directed edges = nonzeros of adj
        # reverse edge: CSR position e <-> CSC position E + c with perm[c] = e
        inv = torch.empty(max(E, 1), dtype=torch.int32, device=g.device)
        inv[g.perm.to(torch.int64)] = torch.arange(E, dtype=torch.int32, device=g.device)
        self.rev = torch.cat([inv[:E] + E, g.adj_eid[E:2 * E]]).contiguous()
        if not self.self_loops:
            deg = torch.diff(g.adj_ptr.to(torch.int64))
            if bool((deg == 0).any()):
                # the reference divides 0 by 0 there and its `assert not torch.isnan(h_prime).any()` fires (SPGA.py:405-407)
                raise ValueError("%d nodes have no edge: the reference asserts on the NaN of their empty attention row (SPGA.py:407)"
                                 % int((deg == 0).sum()))

    def nonzero_index(self) -> torch.Tensor:
        """index into ``adj.nonzero()`` (row-major) of every directed edge in this graph's order: positions 0..2E-1, then the self
        edges.  Used to bring keep masks recorded from / injected into the reference into kernel order."""
        g = self.g
        ptr = g.adj_ptr.to(torch.int64)
        node = torch.repeat_interleave(torch.arange(g.N, device=g.device), torch.diff(ptr))
        pos = torch.arange(2 * g.E, device=g.device)
        if not self.self_loops:
            return pos
        # row n holds its own diagonal entry BEFORE its neighbours if it is a user (neighbours are items, > n), AFTER them if an item
        nb = pos + node + (node < g.U).to(torch.int64)
        n = torch.arange(g.N, device=g.device)
        dg = torch.where(n < g.U, ptr[:-1] + n, ptr[1:] + n)
        return torch.cat([nb, dg])


class SpGraphAttentionLayer(nn.Module):
    """Parameter container with the reference's names and init (SPGA.py:363-374).  All heads of a stage are evaluated batched by
    SpGAT.forward; a single 64 -> 64 layer can also be called on its own."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True):
        super().__init__()
        self.in_features, self.out_features, self.alpha, self.concat, self.p = in_features, out_features, alpha, concat, dropout
        self.W = nn.Parameter(torch.zeros(size=(in_features, out_features)))
        nn.init.xavier_normal_(self.W.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(size=(1, 2 * out_features)))
        nn.init.xavier_normal_(self.a.data, gain=1.414)
        self._call = 0              # one dropout stream per training-mode call of a stand-alone layer

    def forward(self, input, adj, userNum=None):
        if self.in_features != D or self.out_features != D:
            raise NotImplementedError("a stand-alone layer runs for 64 -> 64 only; the 8-wide heads are evaluated batched by SpGAT.forward")
        graph = adj if isinstance(adj, HomoGraph) else HomoGraph(adj, userNum)
        drop = self.p if self.training else 0.0
        call = self._call
        if drop > 0:
            self._call += 1
        Z = SpGATFn.apply(graph, drop, int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF, call, None, ((1, D),), False, input, self.W, self.a)
        return torch.nn.functional.elu(Z) if self.concat else Z

    def __repr__(self):
        return self.__class__.__name__ + " (" + str(self.in_features) + " -> " + str(self.out_features) + ")"


class SpGAT(nn.Module):
    """attention_0..7 (64 -> 8, concat) + out_att (64 -> 64)  (SPGA.py:330-357)"""

    def __init__(self, nfeat, nhid, nclass, dropout, alpha, nheads):
        super().__init__()
        if (nfeat, nhid * nheads, nclass) != (D, D, D) or nheads != 8:
            raise ValueError("the sm_100a kernels are specialised for 64 -> 8 x 8 -> 64")
        self.dropout = dropout
        self.nheads = nheads
        for k in range(nheads):
            self.add_module("attention_{}".format(k), SpGraphAttentionLayer(nfeat, nhid, dropout, alpha, True))
        self.out_att = SpGraphAttentionLayer(nhid * nheads, nclass, dropout, alpha, False)
        self.drop_seed, self._call = None, 0
        self.injected_masks = None      # parity tests: dict(feat=[int64 [N]]*2, edge=[uint8 [n_edges]]*2 in HomoGraph order)
        self._graph_key, self._graph = None, None

    @property
    def attentions(self):
        return [getattr(self, "attention_{}".format(k)) for k in range(self.nheads)]

    def stage_parameters(self):
        att = self.attentions
        return [[l.W for l in att] + [l.a for l in att], [self.out_att.W, self.out_att.a]]

    def graph_for(self, adj, userNum=None) -> HomoGraph:
        if isinstance(adj, HomoGraph):
            return adj
        key = (adj.data_ptr() if not adj.is_sparse else adj._values().data_ptr(), tuple(adj.shape), adj._version, str(adj.device))
        if key != self._graph_key:
            self._graph, self._graph_key = HomoGraph(adj, userNum), key
        return self._graph

    def pre_activation(self, x, adj, userNum=None):
        """out_att output before the final ELU (the scorer applies it on load)"""
        if x.device.type != "cuda":
            raise _lib.NgacfError("SpGAT runs on CUDA only (sm_100a kernels, no CPU fallback)")
        graph = self.graph_for(adj, userNum)
        drop = self.dropout if self.training else 0.0
        if self.drop_seed is None:
            self.drop_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        call = self._call
        if drop > 0:
            self._call += 1
        injected, self.injected_masks = self.injected_masks, None
        params = [p for st in self.stage_parameters() for p in st]
        return SpGATFn.apply(graph, drop, self.drop_seed, call, injected, STAGES, True, x, *params)

    def forward(self, x, adj, userNum=None):
        return torch.nn.functional.elu(self.pre_activation(x, adj, userNum))       # SPGA.py:356


def _stage_lists(params, stages):
    out, i = [], 0
    for H, _ in stages:
        out.append((list(params[i:i + H]), list(params[i + H:i + 2 * H])))
        i += 2 * H
    return out


class SpGATFn(torch.autograd.Function):
    """Z_last = SpGAT stages on x (N,64).  elu_between: the SpGAT composition (dropout on every stage input, ELU on the concatenated
    head outputs, SPGA.py:352-357); False: one layer on its own (edge dropout only)."""

    @staticmethod
    def forward(ctx, graph: HomoGraph, droprate, seed, call, injected, stages, elu_between, x, *params):
        g = graph.g
        dev, N, U, I = g.device, g.N, g.U, g.I
        f32 = dict(dtype=torch.float32, device=dev)
        x = x.detach().contiguous().float()
        if tuple(x.shape) != (N, D):
            raise ValueError("features must be (%d,%d); got %s" % (N, D, tuple(x.shape)))
        S = len(stages)
        per_stage = _stage_lists([p.detach() for p in params], stages)
        wtabs = [ops.pointer_table(Ws + Ws + As) for Ws, As in per_stage]             # [W x H | W x H | a x H]: W_u = W_i = W
        scale, featmask, edgemask = 1.0, [None] * S, [None] * S
        if droprate > 0:
            scale = 1.0 / (1.0 - droprate)
            if injected is not None:
                featmask = [m.contiguous() for m in injected["feat"]]
                edgemask = [m.contiguous() for m in injected["edge"]]
            else:
                edgemask = [torch.empty(graph.n_edges, dtype=torch.uint8, device=dev) for _ in range(S)]
                for k, (H, _) in enumerate(stages):
                    ops.edge_mask(edgemask[k], H, seed, call, k, droprate)
                if elu_between:         # SpGAT.forward drops its stage inputs (SPGA.py:353,355); a layer on its own only its edges (:398)
                    featmask = [torch.empty(N, dtype=torch.int64, device=dev) for _ in range(S)]
                    for k in range(S):
                        ops.feature_mask(featmask[k], seed, call, k, droprate)
        scratch, counter = g.scratch("spgat")
        h = [torch.empty((N, D), **f32) for _ in stages]
        Z = [torch.empty((N, D), **f32) for _ in stages]
        p = [torch.empty((N, H), **f32) for H, _ in stages]
        q = [torch.empty((N, H), **f32) for H, _ in stages]
        norm = [torch.empty((N, H), **f32) for H, _ in stages]
        s_unused = torch.empty((N, 8), **f32)
        Xu, Xi, act = x, x[U:], 0
        for k, (H, _) in enumerate(stages):
            ops.transform_fwd(Xu, Xi, act, featmask[k], scale, wtabs[k], H, U, I, h[k], s_unused)
            ops.node_logits(h[k], wtabs[k], H, N, p[k], q[k])
            ops.spgat_aggregate_fwd(graph, scratch, counter, h[k], p[k], q[k], H, edgemask[k], scale, Z[k], norm[k])
            Xu, Xi, act = Z[k], Z[k][U:], 1
        if not elu_between and S > 1:
            raise NotImplementedError
        ctx.graph, ctx.stages, ctx.wtabs, ctx.scale = graph, stages, wtabs, scale
        ctx.bufs = (x, h, Z, p, q, norm, featmask, edgemask, scratch, counter)
        ctx.shapes = [t.shape for t in params]
        return Z[-1]

    @staticmethod
    def backward(ctx, G):
        graph, stages, wtabs, scale = ctx.graph, ctx.stages, ctx.wtabs, ctx.scale
        g = graph.g
        dev, N, U, I = g.device, g.N, g.U, g.I
        f32 = dict(dtype=torch.float32, device=dev)
        x, h, Z, p, q, norm, featmask, edgemask, scratch, counter = ctx.bufs
        G = G.contiguous().float()
        Ghat, dh = torch.empty((N, D), **f32), torch.empty((N, D), **f32)
        Gbuf = [torch.empty((N, D), **f32), torch.empty((N, D), **f32)]
        dx = torch.empty((N, D), **f32)
        zeros_dS = torch.zeros((N, 8), **f32)
        ws = torch.empty(ops.transform_bwd_workspace_bytes(U, I) // 4, **f32)
        partials = torch.empty((_DA_BLOCKS, 2 * D), **f32)
        grads = [None] * len(ctx.shapes)
        off = [0]
        for H, _ in stages:
            off.append(off[-1] + 2 * H)
        for k in range(len(stages) - 1, -1, -1):
            H, DH = stages[k]
            pairs = torch.empty((graph.n_edges, H, 2), **f32)
            dP, dQ = torch.empty((N, H), **f32), torch.empty((N, H), **f32)
            ops.spgat_bwd(graph, scratch, counter, G, Z[k], norm[k], h[k], p[k], q[k], H, edgemask[k], scale, wtabs[k], Ghat, pairs, dP, dQ, dh)
            ops.node_logits_bwd(h[k], dP, dQ, H, N, partials)
            dWu = [torch.empty((D, DH), **f32) for _ in range(H)]
            dWi = [torch.empty((D, DH), **f32) for _ in range(H)]
            da_unused = [torch.empty((1, 2 * DH), **f32) for _ in range(H)]
            gtab = ops.pointer_table(dWu + dWi + da_unused)
            if k > 0:
                Gprev = Gbuf[k & 1]
                ops.transform_bwd(dh, zeros_dS, None, Z[k - 1], Z[k - 1][U:], 1, featmask[k], scale, wtabs[k], gtab, H, U, I, Gprev, Gprev[U:], 0, 0, ws)
            else:
                ops.transform_bwd(dh, zeros_dS, None, x, x[U:], 0, featmask[k], scale, wtabs[k], gtab, H, U, I, dx, dx[U:], 0, 0, ws)
            da = partials.sum(0)                                   # (128,): [sum dP h | sum dQ h], fixed block order
            for j in range(H):
                grads[off[k] + j] = dWu[j] + dWi[j]                # one W serves both halves of the node table
                grads[off[k] + H + j] = torch.cat([da[j * DH:(j + 1) * DH], da[D + j * DH:D + (j + 1) * DH]]).view(1, 2 * DH)
            if k > 0:
                G = Gprev
        ctx.bufs = None
        return (None, None, None, None, None, None, None, dx, *grads)


class SPGACF(nn.Module):
    """Drop-in for graphattention/SPGA.py:85-140: same constructor, parameter names (uEmbd, iEmbd, gat.attention_k.W/.a,
    gat.out_att.W/.a) and ``forward(userIdx, itemIdx, mask) -> scores[B]`` with ``mask`` the (N,N) adjacency."""

    def __init__(self, userNum, itemNum, adj, embedSize, layers, droprate, useCuda=True):
        super().__init__()
        if embedSize != D:
            raise ValueError("the sm_100a kernels are specialised for embedSize 64")
        self.useCuda, self.userNum, self.itemNum, self.droprate = useCuda, int(userNum), int(itemNum), float(droprate)
        self.uEmbd = nn.Embedding(userNum, embedSize)
        self.iEmbd = nn.Embedding(itemNum, embedSize)
        self.gat = SpGAT(nfeat=embedSize, nhid=8, nclass=embedSize, dropout=droprate, nheads=8, alpha=ALPHA)
        nn.init.normal_(self.uEmbd.weight, std=0.01)
        nn.init.normal_(self.iEmbd.weight, std=0.01)

    def getFeatureMat(self):
        dev = self.uEmbd.weight.device
        uidx = torch.arange(self.userNum, device=dev)
        iidx = torch.arange(self.itemNum, device=dev)
        return uidx, iidx + self.userNum, torch.cat([self.uEmbd.weight, self.iEmbd.weight], dim=0)

    def forward(self, userIdx, itemIdx, mask):
        _, _, features = self.getFeatureMat()
        Z = self.gat.pre_activation(features, mask, self.userNum)
        return ScoreFn.apply(Z, self.userNum, userIdx.to(Z.device), itemIdx.to(Z.device))
