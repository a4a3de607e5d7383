"""GPU experiment: dense stage transforms (ngacf_transform_fwd / ngacf_transform_bwd_dx) against an fp64 torch product.
Run twice (NGACF_DENSE=ffma / default tc) to compare the CUDA-core and tcgen05 3xTF32 kernels: max relative error and time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngacf_b200 import ops  # noqa: E402


def run(U, I, H, act, p, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    N, D, DH = U + I, 64, 64 // H
    Xu = torch.randn(U, D, device=dev, generator=g)
    Xi = torch.randn(I, D, device=dev, generator=g)
    Wu = [torch.randn(D, DH, device=dev, generator=g) * 0.3 for _ in range(H)]
    Wi = [torch.randn(D, DH, device=dev, generator=g) * 0.3 for _ in range(H)]
    a = [torch.randn(2 * DH, device=dev, generator=g) for _ in range(H)]
    wtab = torch.tensor([t.data_ptr() for t in Wu + Wi + a], dtype=torch.int64, device=dev)
    fm, scale = None, 1.0
    if p > 0:
        fm = torch.empty(N, dtype=torch.int64, device=dev)
        ops.feature_mask(fm, 1234, 3, 0, p)
        scale = 1.0 / (1.0 - p)
    h = torch.empty(N, D, device=dev)
    s = torch.empty(N, H, device=dev)
    ops.transform_fwd(Xu, Xi, act, fm, scale, wtab, H, U, I, h, s)
    torch.cuda.synchronize()
    # fp64 reference
    X = torch.cat([Xu, Xi]).double()
    if act:
        X = torch.where(X > 0, X, torch.expm1(X))
    if fm is not None:
        bits = ((fm.view(-1, 1) >> torch.arange(64, device=dev).view(1, -1)) & 1).double()
        X = X * bits * scale
    Wcu, Wci = torch.cat(Wu, 1).double(), torch.cat(Wi, 1).double()
    href = torch.cat([X[:U] @ Wcu, X[U:] @ Wci])
    au = torch.cat([t[:DH] for t in a]).double()
    ai = torch.cat([t[DH:] for t in a]).double()
    sref = torch.cat([(href[:U] * au).view(U, H, DH).sum(-1), (href[U:] * ai).view(I, H, DH).sum(-1)])
    eh = ((h.double() - href).abs().max() / href.abs().max()).item()
    es = ((s.double() - sref).abs().max() / sref.abs().max()).item()
    # backward dX
    dh = torch.randn(N, D, device=dev, generator=g)
    Zu, Zi = Xu, Xi
    dXu, dXi = torch.zeros(U, D, device=dev), torch.zeros(I, D, device=dev)
    ops.transform_bwd_dx(dh, Zu if act else None, Zi if act else None, act, fm, scale, wtab, H, U, I, dXu, dXi, 0)
    ops.transform_bwd_dx(dh, Zu if act else None, Zi if act else None, act, fm, scale, wtab, H, U, I, dXu, dXi, 1)   # accumulate: 2x
    torch.cuda.synchronize()
    dref = torch.cat([dh[:U].double() @ Wcu.T, dh[U:].double() @ Wci.T])
    if fm is not None:
        dref = dref * bits * scale
    if act:
        Z = torch.cat([Zu, Zi]).double()
        dref = dref * torch.where(Z > 0, torch.ones_like(Z), torch.exp(Z))
    dref = 2 * dref
    ed = ((torch.cat([dXu, dXi]).double() - dref).abs().max() / dref.abs().max()).item()

    # fused backward: dX, dW, da
    dS = torch.randn(N, H, device=dev, generator=g)
    gWu = [torch.zeros(D, DH, device=dev) for _ in range(H)]
    gWi = [torch.zeros(D, DH, device=dev) for _ in range(H)]
    ga = [torch.zeros(2 * DH, device=dev) for _ in range(H)]
    gtab = torch.tensor([t.data_ptr() for t in gWu + gWi + ga], dtype=torch.int64, device=dev)
    ws = torch.empty(ops.transform_bwd_workspace_bytes(U, I) // 4 + 16, device=dev)
    fXu, fXi = torch.zeros(U, D, device=dev), torch.zeros(I, D, device=dev)
    ops.transform_bwd(dh, dS, h, Xu, Xi, act, fm, scale, wtab, gtab, H, U, I, fXu, fXi, 0, 0, ws)
    ops.transform_bwd(dh, dS, h, Xu, Xi, act, fm, scale, wtab, gtab, H, U, I, fXu, fXi, 1, 1, ws)    # accumulate: 2x
    torch.cuda.synchronize()
    efx = ((torch.cat([fXu, fXi]).double() - dref).abs().max() / dref.abs().max()).item()
    dWu_ref = 2 * X[:U].T @ dh[:U].double()
    dWi_ref = 2 * X[U:].T @ dh[U:].double()
    ew = max(((torch.cat(gWu, 1).double() - dWu_ref).abs().max() / dWu_ref.abs().max()).item(),
             ((torch.cat(gWi, 1).double() - dWi_ref).abs().max() / dWi_ref.abs().max()).item())
    dau = 2 * (dS[:U].double().repeat_interleave(DH, 1) * href[:U]).sum(0)
    dai = 2 * (dS[U:].double().repeat_interleave(DH, 1) * href[U:]).sum(0)
    ga_u = torch.cat([t[:DH] for t in ga]).double()
    ga_i = torch.cat([t[DH:] for t in ga]).double()
    ea = max(((ga_u - dau).abs().max() / dau.abs().max()).item(), ((ga_i - dai).abs().max() / dai.abs().max()).item())

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    tf = timeit(lambda: ops.transform_fwd(Xu, Xi, act, fm, scale, wtab, H, U, I, h, s))
    td = timeit(lambda: ops.transform_bwd_dx(dh, Zu if act else None, Zi if act else None, act, fm, scale, wtab, H, U, I, dXu, dXi, 0))
    tb = timeit(lambda: ops.transform_bwd(dh, dS, h, Xu, Xi, act, fm, scale, wtab, gtab, H, U, I, fXu, fXi, 0, 0, ws))
    print("U=%d I=%d H=%d act=%d p=%.1f  err h %.2e s %.2e dX %.2e | fused dX %.2e dW %.2e da %.2e | fwd %.1f us  dx %.1f us  bwd %.1f us"
          % (U, I, H, act, p, eh, es, ed, efx, ew, ea, tf, td, tb), flush=True)
    return max(eh, es, ed, efx, ew, ea)


if __name__ == "__main__":
    print("NGACF_DENSE =", os.environ.get("NGACF_DENSE", "(tc)"))
    if "--check" in sys.argv:       # tests/test_gpu_parity.py::test_dense_transforms_vs_fp64: ragged, tiny and multi-tile shapes
        worst = 0.0
        for (U, I) in ((1, 1), (127, 129), (300, 517), (2048, 1000)):
            for H in (8, 1):
                for act, p in ((0, 0.0), (1, 0.2), (1, 0.0), (0, 0.5)):
                    worst = max(worst, run(U, I, H, act, p, seed=U + H))
        print("worst relative error", worst)
        sys.exit(0 if worst < 2e-5 else 1)
    for (U, I) in ((300, 517), (29858, 40981)):
        for H in (8, 1):
            for act, p in ((0, 0.0), (1, 0.2)):
                run(U, I, H, act, p)
