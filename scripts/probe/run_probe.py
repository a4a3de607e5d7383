"""Explores tcgen05 descriptor settings (kind::tf32 and kind::f16, K-major vs MN-major operands over [chunk][row][16 B] images):
D[m][n] = sum_r A[r][m] B[r][n].  Findings: profiles/r1e_umma_probe.txt.  Build the probe first:

    nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Xcompiler -fPIC -shared scripts/probe/umma_probe.cu -o scripts/probe/libprobe.so
"""
import ctypes, os, sys, itertools
import torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "libprobe.so"))
P = ctypes.c_void_p
lib.probe_launch.argtypes = [P, P, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                             ctypes.c_uint32, P]
dev = torch.device("cuda:0")
torch.manual_seed(0)
A = torch.randn(128, 64, device=dev)
B = torch.randn(128, 80, device=dev)
Dout = torch.zeros(128, 96, device=dev)


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(name, M, N, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, expect):
    Dout.fill_(-777.0)
    rc = lib.probe_launch(A.data_ptr(), B.data_ptr(), idesc(M, N, a_mn, b_mn), a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, Dout.data_ptr())
    D = Dout.clone()
    e = expect.float()
    m, n = e.shape
    err = ((D[:m, :n] - e).abs().max() / e.abs().max()).item()
    print("%-34s rc=%d err=%.3e  D[0,:4]=%s  exp[0,:4]=%s  |D|max=%.3g" % (name, rc, err, [round(x, 3) for x in D[0, :4].tolist()],
                                                                         [round(x, 3) for x in e[0, :4].tolist()], D[:m, :n].abs().max().item()), flush=True)
    return D



lib.probe_bf16_launch.argtypes = lib.probe_launch.argtypes
exp_mn = A.double().T @ B.double()              # [64][80]
# tf32: only one operand MN-major (B K-major needs an [n][k] image: use A's image as B, N = 64 rows of A)
Apad = torch.zeros(128, 80, device=dev); Apad[:, :64] = A
def run_b(name, M, N, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, expect, Bbuf):
    Dout.fill_(-777.0)
    idc = (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)
    rc = lib.probe_bf16_launch(A.data_ptr(), Bbuf.data_ptr(), idc, a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, Dout.data_ptr())
    D = Dout.clone(); e = expect.float(); m, n = e.shape
    err = ((D[:m, :n] - e).abs().max() / e.abs().max()).item()
    print("%-40s rc=%d err=%.3e  D[0,:4]=%s  exp[0,:4]=%s" % (name, rc, err, [round(x, 3) for x in D[0, :4].tolist()], [round(x, 3) for x in e[0, :4].tolist()]), flush=True)

# bf16 K-major sanity: D[r][n] = sum_c A[r][c] A[n][c]; K = 16 per MMA = 2 chunks; 4 k-steps
run_b("bf16 K-major both", 128, 64, 0, 0, 2048, 128, 2048, 128, 4, 4096, 4096, A.double() @ A[:64].double().T, Apad)
# bf16 MN-major both: K = rows, 16 per MMA = two 8-row groups (LBO = 128), MN chunks at SBO = 2048; 8 k-steps of 256 B
for (M, N) in ((128, 80), (64, 80), (128, 64)):
    for (lbo, sbo) in ((128, 2048), (2048, 128)):
        run_b("bf16 MN both M=%d N=%d lbo=%d sbo=%d" % (M, N, lbo, sbo), M, N, 1, 1, lbo, sbo, lbo, sbo, 8, 256, 256, exp_mn[:, :N], B)
# tf32 mixed majors
B2 = Apad
Dout.fill_(0)
def run_t(name, M, N, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, expect, Bbuf):
    Dout.fill_(-777.0)
    rc = lib.probe_launch(A.data_ptr(), Bbuf.data_ptr(), idesc(M, N, a_mn, b_mn), a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, Dout.data_ptr())
    D = Dout.clone(); e = expect.float(); m, n = e.shape
    err = ((D[:m, :n] - e).abs().max() / e.abs().max()).item()
    print("%-40s rc=%d err=%.3e  D[0,:4]=%s  exp[0,:4]=%s" % (name, rc, err, [round(x, 3) for x in D[0, :4].tolist()], [round(x, 3) for x in e[0, :4].tolist()]), flush=True)
# A MN-major (M = feature of A, K = rows: 8 per MMA), B K-major: B image rows = n, K along chunks -> needs B[n][k=r]: not available; just see if non-zero
run_t("tf32 A MN, B K (nonzero?)", 128, 64, 1, 0, 128, 2048, 2048, 128, 8, 128, 4096, exp_mn[:, :64], Apad)
run_t("tf32 A K, B MN (nonzero?)", 128, 64, 0, 1, 2048, 128, 128, 2048, 8, 4096, 128, exp_mn[:, :64], B)
