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
buf = np.zeros((2, 6, 512), np.int64)
rc = lib.ngacf_debug_topk_trace(ctypes.c_void_p(buf.ctypes.data))
assert rc == 0, rc
cta = np.zeros((1024, 4), np.int64)
assert lib.ngacf_debug_topk_cta(ctypes.c_void_p(cta.ctypes.data)) == 0
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/topk_trace.npy", buf)
np.save("gpurun_out/topk_cta.npy", cta)
nb = (inter.eval_users.numel() + 127) // 128
c = cta[:nb]
t0g = c[:, 0].min()
dur = (c[:, 1] - c[:, 0]) / 1000.0
print("CTAs %d: epilogue duration us min %.0f median %.0f max %.0f; last end at %.0f us; starts spread %.0f us" % (nb, dur.min(), np.median(dur), dur.max(), (c[:, 1].max() - t0g) / 1000.0, (c[:, 0].max() - t0g) / 1000.0))
sm_count = np.bincount(c[:, 2].astype(int), minlength=148)
shared = sm_count[c[:, 2].astype(int)] > 1
print("  CTAs sharing an SM: %d, median dur %.0f us; alone: %d, median dur %.0f us" % (shared.sum(), np.median(dur[shared]), (~shared).sum(), np.median(dur[~shared])))
print("  flushes (warp 0) min/median/max", c[:, 3].min(), np.median(c[:, 3]), c[:, 3].max())
order = np.argsort(-dur)[:8]
print("  slowest CTAs:", [(int(k), round(float(dur[k])), int(c[k, 2]), int(c[k, 3])) for k in order])
names = ["tma_issue", "mma_full_ok", "mma_tempty_ok", "mma_issued", "epi_tfull_ok", "epi_done"]
nt = 321
for c in range(2):
    t = buf[c][:, :nt].astype(np.float64)
    t0 = t[0, 0]
    print("CTA slot %d: total %.0f cycles for %d tiles = %.0f cycles/tile" % (c, t[5, nt - 1] - t0, nt, (t[5, nt - 1] - t0) / nt))
    for lt in list(range(0, 6)) + list(range(150, 156)):
        print("  tile %3d " % lt + "  ".join("%s %7.0f" % (n, t[k, lt] - t0) for k, n in enumerate(names)))
    mid = slice(20, nt - 5)
    print("  steady-state means (cycles):")
    print("    TMA issue -> full seen by MMA      %7.0f" % np.mean(t[1, mid] - t[0, mid]))
    print("    MMA wait for tempty after full     %7.0f" % np.mean(t[2, mid] - t[1, mid]))
    print("    MMA issue (12 MMAs + commits)      %7.0f" % np.mean(t[3, mid] - t[2, mid]))
    print("    MMA issued -> tfull seen by epi    %7.0f" % np.mean(t[4, mid] - t[3, mid]))
    print("    epilogue (tfull -> done)           %7.0f" % np.mean(t[5, mid] - t[4, mid]))
    print("    epi done(t) -> MMA tempty ok(t+2)  %7.0f" % np.mean(t[2, 22:nt - 3] - t[5, 20:nt - 5]))
    print("    MMA issued(t) -> TMA issue(t+2)    %7.0f" % np.mean(t[0, 22:nt - 3] - t[3, 20:nt - 5]))
    print("    tile period (epi done deltas)      %7.0f" % np.mean(np.diff(t[5, mid])))
