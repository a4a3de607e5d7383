import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngacf_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
U, I, H, D = 300, 517, 8, 64
DH = D // H
N = U + I
Xu = torch.randn(U, D, device=dev, generator=g); Xi = torch.randn(I, D, device=dev, generator=g)
Wu = [torch.randn(D, DH, device=dev, generator=g) * 0.3 for _ in range(H)]
Wi = [torch.randn(D, DH, device=dev, generator=g) * 0.3 for _ in range(H)]
a = [torch.randn(2 * DH, device=dev, generator=g) for _ in range(H)]
wtab = torch.tensor([t.data_ptr() for t in Wu + Wi + a], dtype=torch.int64, device=dev)
dh = torch.randn(N, D, device=dev, generator=g); dS = torch.randn(N, H, device=dev, generator=g)
gWu = [torch.zeros(D, DH, device=dev) for _ in range(H)]; gWi = [torch.zeros(D, DH, device=dev) for _ in range(H)]
ga = [torch.zeros(2 * DH, device=dev) for _ in range(H)]
gtab = torch.tensor([t.data_ptr() for t in gWu + gWi + ga], dtype=torch.int64, device=dev)
ws = torch.full((ops.transform_bwd_workspace_bytes(U, I) // 4 + 16,), 7.0, device=dev)
h = torch.empty(N, D, device=dev)
fXu, fXi = torch.zeros(U, D, device=dev), torch.zeros(I, D, device=dev)
ops.transform_bwd(dh, dS, h, Xu, Xi, 0, None, 1.0, wtab, gtab, H, U, I, fXu, fXi, 0, 0, ws)
torch.cuda.synchronize()
ref = Xu.double().T @ dh[:U].double()
got = torch.cat(gWu, 1)
print("ref[0,:8]", ref[0, :8].tolist())
print("got[0,:8]", got[0, :8].tolist())
print("ws part0 [0,:8]", ws[:8].tolist(), "ws[4096:4104]", ws[4096:4104].tolist())
p0 = ws[:4096].view(64, 64)
# partial of CTA 0 = tile 0 of users (rows 0..127)
r0 = Xu[:128].double().T @ dh[:128].double()
print("part0 vs tile-0 ref: maxerr", (p0.double() - r0).abs().max().item(), "ref max", r0.abs().max().item())
print("part0 transposed?", (p0.double().T - r0).abs().max().item())
print("nonzero frac", (p0 != 0).float().mean().item(), "sevens", (p0 == 7).float().mean().item())
