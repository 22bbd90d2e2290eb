"""Ablation timings of the brick-mode 3x3x3 kernel at block-1 size (latency analysis aid, not a test)."""
import sys
import torch
sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L
from tests import engine_helpers as H


def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


adt = H.act_dtype()
for (B, dims) in ((16, (32, 32, 16)), (16, (16, 16, 8)), (16, (8, 8, 4))):
    Dz, Dy, Dx = dims
    M = B * Dz * Dy * Dx
    Cin, N = 128, 32
    bott = torch.randn(M, Cin, device="cuda").to(adt)
    w = torch.randn(N, Cin, 3, 3, 3, device="cuda") * 0.05
    bp = H.pack(w, N, 32, Cin, 64, 27, Cin * 27, 27, 1)
    out = torch.zeros(M, 256, dtype=adt, device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    gamma = torch.ones(Cin, device="cuda"); beta = torch.zeros(Cin, device="cuda")
    s1 = torch.zeros(Cin, dtype=torch.float64, device="cuda"); s2 = torch.full((Cin,), float(M), dtype=torch.float64, device="cuda")
    bn = H.bnsrc(s1, s2, gamma, beta, count=M)
    t_full = timeit(lambda: H.brick(B, dims, Cin, N, bott, Cin, bp, out[:, 64:], 256, bnA=bn, st_sum=st[0], st_sq=st[1]))
    t_raw = timeit(lambda: H.brick(B, dims, Cin, N, bott, Cin, bp, out[:, 64:], 256, grad=2, st_sum=st[0], st_sq=st[1]))
    g = torch.randn(M, 32, device="cuda").to(torch.bfloat16)
    wd = torch.randn(32, 128, 3, 3, 3, device="cuda") * 0.05
    bpd = H.pack(wd, 128, 128, 32, 32, 27, 27, 128 * 27, 1, fwd=False)
    outd = torch.zeros(M, 128, dtype=torch.bfloat16, device="cuda")
    std = torch.zeros(2, 128, dtype=torch.float64, device="cuda")
    xb = torch.randn(M, 128, device="cuda").to(adt)
    bnE = H.bnsrc(torch.zeros(128, dtype=torch.float64, device="cuda"), torch.full((128,), float(M), dtype=torch.float64, device="cuda"), torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"), count=M)
    t_dg = timeit(lambda: H.brick(B, dims, 32, 128, g, 32, bpd, outd, 128, grad=1, tap_sign=-1, st_sum=std[0], st_sq=std[1], e_src=xb, e_pitch=128, bnE=bnE))
    fl = 2.0 * M * 32 * 128 * 27
    print(f"dims {dims} M={M}: fprop {t_full:.1f} us ({fl/t_full/1e6:.0f} TF/s) | fprop no-transform {t_raw:.1f} us | dgrad {t_dg:.1f} us ({fl/t_dg/1e6:.0f} TF/s)")
