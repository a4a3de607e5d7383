"""Diagnostic: pipeline timeline of score_topk_tc_kernel (TMA producer / MMA issuer / epilogue) for two CTAs.
Needs a debug build:  NGACF_NVCC_EXTRA=-DNGACF_TOPK_TRACE python -m ngacf_b200.build --force
Run on a GPU box:     python scripts/probe/trace_topk.py        (rebuild WITHOUT the flag afterwards)"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ngacf_b200 import _lib, hostdata  # noqa: E402
from ngacf_b200.data import Interactions  # noqa: E402
from ngacf_b200.evaluate import AllNegEvaluator  # noqa: E402
from ngacf_b200.model import SPUIGACF  # noqa: E402

DEV = "cuda:0"
U, I, E = 29858, 40981, 1027370
u, i = hostdata.synth_bipartite(U, I, E, 0)
(tu, ti), (su, si) = hostdata.split_per_user(u, i, U, 1)
torch.manual_seed(2019)
model = SPUIGACF(U, I, 64, [64, 64], 0.2).to(DEV).eval()
inter = Interactions.from_arrays(U, I, tu, ti, su, si, device=DEV)
adj = torch.from_numpy(np.stack([u, i])).to(DEV)
ev = AllNegEvaluator(inter, "tc")
with torch.no_grad():
    Z = model.propagate(adj)
    for _ in range(3):
        ev.rank(Z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ev.rank(Z)
    e1.record()
    torch.cuda.synchronize()
print("rank(): %.3f ms" % (e0.elapsed_time(e1) / 5))
lib = _lib.load()
buf = np.zeros((2, 8, 512), np.int64)
rc = lib.ngacf_debug_topk_trace(ctypes.c_void_p(buf.ctypes.data))
assert rc == 0, rc
rs = np.zeros(4, np.uint64)
assert lib.ngacf_debug_rescore_stat(ctypes.c_void_p(rs.ctypes.data)) == 0
print("rescore: kept candidates per user %.1f; proof failures over all calls: %d with fewer than K candidates, %d on the margin (of %d user evaluations)" % (rs[0] / max(rs[2], 1), rs[1], rs[3], rs[2]))
cta = np.zeros((4096, 4), np.int64)
assert lib.ngacf_debug_topk_cta(ctypes.c_void_p(cta.ctypes.data)) == 0
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/topk_trace.npy", buf)
np.save("gpurun_out/topk_cta.npy", cta)
c = cta[cta[:, 1] > 0]
nb = c.shape[0]
t0g = c[:, 0].min()
dur = (c[:, 1] - c[:, 0]) / 1000.0
print("CTAs %d: epilogue duration us min %.0f median %.0f max %.0f; last end at %.0f us; starts spread %.0f us" % (nb, dur.min(), np.median(dur), dur.max(), (c[:, 1].max() - t0g) / 1000.0, (c[:, 0].max() - t0g) / 1000.0))
sm_count = np.bincount(c[:, 2].astype(int), minlength=148)
shared = sm_count[c[:, 2].astype(int)] > 1
print("  CTAs sharing an SM: %d, median dur %.0f us; alone: %d, median dur %.0f us" % (shared.sum(), np.median(dur[shared]), (~shared).sum(), np.median(dur[~shared])))
print("  log entries of (warp 0, lane 0) min/median/max", c[:, 3].min(), np.median(c[:, 3]), c[:, 3].max())
print("  kernel span %.0f us; sum of CTA durations / (2 x 148 slots) = %.0f us" % ((c[:, 1].max() - t0g) / 1000.0, dur.sum() / 296))
order = np.argsort(-dur)[:8]
print("  slowest CTAs:", [(int(k), round(float(dur[k])), int(c[k, 2]), int(c[k, 3])) for k in order])
names = ["tma_issue", "mma_full_ok", "mma_tempty_ok", "mma_issued", "epi_tfull_ok", "epi_done"]
for c in range(2):
    x = buf[c].astype(np.float64)
    nt = int(np.count_nonzero(x[5]))
    if nt < 8:
        continue
    x = x[:, :nt]
    epi = x[5] - x[4]
    print("traced CTA %d: %d tiles, %.0f cycles (%.0f/tile)" % (c, nt, x[5, nt - 1] - x[0, 0], (x[5, nt - 1] - x[0, 0]) / nt))
    for lt in range(0, 4):
        print("   tile %d " % lt + "  ".join("%s %7.0f" % (n, x[k, lt] - x[0, 0]) for k, n in enumerate(names)))
    edges = [0, 4, 8, 16, 32, 64, 128, 192, 256, 100000]
    for a, b in zip(edges[:-1], edges[1:]):
        b = min(b, nt)
        if a >= b:
            break
        el = x[5, b - 1] - (x[5, a - 1] if a > 0 else x[0, 0])
        print("   tiles %3d-%3d: %8.0f cycles (%6.0f/tile)  epilogue mean %6.0f max %6.0f   epi_done(t-1)->tfull_ok(t) mean %6.0f" % (
            a, b, el, el / (b - a), epi[a:b].mean(), epi[a:b].max(), np.mean(x[4, max(a, 1):b] - x[5, max(a, 1) - 1:b - 1])))
    gap = x[4, 1:] - x[5, :-1]
    big = np.argsort(-gap)[:8]
    print("   largest epi_done(t-1)->tfull_ok(t):", sorted([(int(k) + 1, int(gap[k])) for k in big]))
    sched = [t for t in range(1, nt) if t & (t - 1) == 0]
    print("   scheduled merges (tile: merge, exchange incl. barrier waits, then wait for the accumulator):",
          [(t, int(x[6, t] - x[5, t - 1]), int(x[7, t] - x[6, t]), int(x[4, t] - x[7, t])) for t in sched])
