"""Diagnostic: phase timeline of transform_tc_kernel (forward dense transform) for two CTAs.
Needs a debug build:  NGACF_NVCC_EXTRA=-DNGACF_DENSE_TRACE python -m ngacf_b200.build --force ; rebuild without it afterwards."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ngacf_b200 import _lib, ops  # noqa: E402
from ngacf_b200.model import SPUIGACF  # noqa: E402

DEV = "cuda:0"
U, I = 29858, 40981
torch.manual_seed(0)
model = SPUIGACF(U, I, 64, [64, 64], 0.2).to(DEV)
sp = model.gat.stage_parameters()
wt = [ops.pointer_table([p.detach() for p in st]) for st in sp]
h = torch.empty((U + I, 64), device=DEV)
s = torch.empty((U + I, 8), device=DEV)
for k, H in ((0, 8), (1, 1)):
    fn = lambda: ops.transform_fwd(model.uEmbd.weight.detach(), model.iEmbd.weight.detach(), 0, None, 1.0, wt[k], H, U, I, h, s)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("transform_fwd H=%d: %.2f us per launch (20 back to back in a graph)" % (H, e0.elapsed_time(e1) / 20 * 1000))
    fn()
    torch.cuda.synchronize()
    buf = np.zeros((2, 32), np.int64)
    assert _lib.load().ngacf_debug_dense_trace(ctypes.c_void_p(buf.ctypes.data)) == 0
    for c in range(2):
        t = buf[c].astype(np.float64)
        t0 = t[0]
        print("  CTA %d: prologue %.0f cycles (tile request issued %.0f, barrier init + TMEM alloc %.0f, W loads issued/arrived %.0f, split + stores %.0f, barrier %.0f)" % (
            c, t[1] - t0, t[27] - t0, t[28] - t[27], t[29] - t[28], t[30] - t[29], t[1] - t[30]))
        k2 = 0
        while 6 + 5 * k2 < 32 and t[6 + 5 * k2] > 0 and t[6 + 5 * k2] > t0:
            b = 2 + 5 * k2
            prev = t[1] if k2 == 0 else t[b - 1]
            print("    tile %d: image %.0f, MMA issue %.0f, MMA wait %.0f, epilogue math %.0f, staging+stores+sync %.0f   (end at %.0f)" % (
                k2, t[b] - prev, t[b + 1] - t[b], t[b + 2] - t[b + 1], t[b + 3] - t[b + 2], t[b + 4] - t[b + 3], t[b + 4] - t0))
            k2 += 1

# ---- fused backward kernel (dX, dW, da): one 512-thread CTA per SM, persistent over ~3.7 tiles ----
from ngacf_b200.graph import BipartiteGraph  # noqa: E402,F401
N = U + I
dh = torch.randn((N, 64), device=DEV)
dS = torch.randn((N, 8), device=DEV)
X = torch.randn((N, 64), device=DEV)
dX = torch.empty((N, 64), device=DEV)
ws = torch.empty(ops.transform_bwd_workspace_bytes(U, I) // 4, device=DEV)
for k, H in ((0, 8), (1, 1)):
    grads = [torch.empty_like(p_) for p_ in sp[k]]
    gt = ops.pointer_table(grads)
    fn = lambda: ops.transform_bwd(dh, dS, None, X, X[U:], 1, None, 1.0, wt[k], gt, H, U, I, dX, dX[U:], 0, 0, ws)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("transform_bwd H=%d: %.2f us per call incl. the partial reduction (20 back to back in a graph)" % (H, e0.elapsed_time(e1) / 20 * 1000))
    fn()
    torch.cuda.synchronize()
    buf = np.zeros((2, 32), np.int64)
    assert _lib.load().ngacf_debug_dense_bwd_trace(ctypes.c_void_p(buf.ctypes.data)) == 0
    for c in range(2):
        t = buf[c].astype(np.float64)
        t0 = t[0]
        print("  CTA %d: prologue %.0f cycles, whole CTA %.0f" % (c, t[1] - t0, t[30] - t0))
        k2 = 0
        while 5 + 4 * k2 < 30 and t[5 + 4 * k2] > t0:
            b = 2 + 4 * k2
            prev = t[1] if k2 == 0 else t[b - 1]
            print("    tile %d: images %.0f, MMA issue (72) %.0f, MMA wait %.0f, epilogue + stores + sync %.0f" % (
                k2, t[b] - prev, t[b + 1] - t[b], t[b + 2] - t[b + 1], t[b + 3] - t[b + 2]))
            k2 += 1
        print("    partials + teardown %.0f" % (t[30] - t[5 + 4 * (k2 - 1)]))
