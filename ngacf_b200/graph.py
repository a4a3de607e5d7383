"""BipartiteGraph: the user-item adjacency resident in HBM as CSR + CSC + CSC->CSR permutation, the
unified node adjacency and the degree-bucketed task list the aggregation kernels walk.

Built once from the COO ``mask`` tensor the reference passes to every ``model(u, i, adj)`` call
(run_Gowalla.py:94 -> train_eval_Gowalla.py:106,131) by ``ngacf_graph_build``; the reference instead
rebuilds and re-coalesces a COO tensor four times per head-layer (SPUIGACF.py:365-377).
"""
from __future__ import annotations

import torch

from . import _lib

CHUNK = 128
SCRATCH_STRIDE = 72


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream():
    return torch.cuda.current_stream().cuda_stream


class BipartiteGraph:
    def __init__(self, indices: torch.Tensor, userNum: int, itemNum: int, check_users: bool = True):
        """indices: (2,E) or (1,2,E) integer CUDA tensor of (user,item) pairs, any order, duplicates allowed."""
        if indices.dim() == 3:
            indices = indices.squeeze(0)
        if indices.dim() != 2 or indices.shape[0] != 2:
            raise ValueError("mask must have shape (2,E); got %s" % (tuple(indices.shape),))
        _lib.require_cuda(indices)
        dev = indices.device
        idx = indices.to(torch.int64).contiguous()
        E_in = idx.shape[1]
        U, I = int(userNum), int(itemNum)
        N = U + I
        self.U, self.I, self.N, self.device = U, I, N, dev
        i32 = dict(dtype=torch.int32, device=dev)
        cap_e = max(E_in, 1)
        cap_tasks = N + 2 * E_in // CHUNK + 2
        cap_long = 2 * E_in // CHUNK + 2
        rowptr = torch.empty(U + 1, **i32)
        colidx = torch.empty(cap_e, **i32)
        colptr = torch.empty(I + 1, **i32)
        rowidx = torch.empty(cap_e, **i32)
        perm = torch.empty(cap_e, **i32)
        adj_ptr = torch.empty(N + 1, **i32)
        adj_idx = torch.empty(2 * cap_e, **i32)
        adj_eid = torch.empty(2 * cap_e, **i32)
        tasks = torch.empty((cap_tasks, 4), **i32)
        long_first = torch.zeros(cap_long, **i32)
        long_counter = torch.zeros(cap_long, **i32)
        counts = torch.zeros(8, **i32)
        ws_bytes = int(_lib.load().ngacf_graph_build_workspace_bytes(E_in, U, I))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.call("ngacf_graph_build", _ptr(idx[0]), _ptr(idx[1]), E_in, U, I, _ptr(rowptr), _ptr(colidx), _ptr(colptr),
                      _ptr(rowidx), _ptr(perm), _ptr(adj_ptr), _ptr(adj_idx), _ptr(adj_eid), _ptr(tasks), _ptr(long_first),
                      _ptr(long_counter), _ptr(counts), _ptr(ws), ws_bytes, _stream())
            c = counts.cpu().tolist()       # the one build-time sync: data-dependent sizes
        E, T, L, S, zero_users, bad, Tu = c[0], c[1], c[2], c[3], c[4], c[5], c[6]
        if bad:
            raise ValueError("%d edges have a user/item index outside [0,%d)x[0,%d)" % (bad, U, I))
        if check_users and zero_users:
            # the reference asserts e_rowsum != 0 for every user at every layer (SPUIGACF.py:368)
            raise ValueError("%d users have no edge: the reference asserts a non-zero attention row sum "
                             "(SPUIGACF.py:368)" % zero_users)
        self.E, self.T, self.L, self.S, self.T_users = E, T, L, S, Tu
        # adj_* were laid out by the builder with the CSC half starting at E (not E_in)
        self.rowptr, self.colidx = rowptr, colidx[:E]
        self.colptr, self.rowidx, self.perm = colptr, rowidx[:E], perm[:E]
        self.adj_ptr, self.adj_idx, self.adj_eid = adj_ptr, adj_idx[:2 * E], adj_eid[:2 * E]
        self.tasks = tasks[:T]
        self.long_first_slot = long_first[:L + 1]
        self.long_counter = long_counter[:max(L, 1)]
        self._scratch = {}
        del ws

    def scratch(self, key="default"):
        """Per-stream partial-sum slots for rows longer than CHUNK edges (one buffer per concurrent stream)."""
        if key not in self._scratch:
            self._scratch[key] = (torch.empty(max(self.S, 1) * SCRATCH_STRIDE, dtype=torch.float32, device=self.device),
                                  torch.zeros(max(self.L, 1), dtype=torch.int32, device=self.device))
        return self._scratch[key]

    def structure_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.rowptr, self.colidx, self.colptr, self.rowidx, self.perm,
                                                          self.adj_ptr, self.adj_idx, self.adj_eid, self.tasks))
