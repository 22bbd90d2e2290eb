"""Micro-benchmark of the block-1 / block-2 1x1x1 forward GEMM launches of configs[1] (M = 262144 / 32768): which part of the launch costs
what (raw vs BN+ReLU operand, store vs store + statistics).  Analysis aid, not a test."""
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
for M, ctot in ((262144, 256), (32768, 512)):
    a = torch.randn(M, ctot, device="cuda").to(adt)
    for Cin in (64, 224) if ctot == 256 else (128, 480):
        N = 128
        w = torch.randn(N, Cin, device="cuda") * 0.1
        bp = H.pack(w, N, 128, Cin, 64, 1, Cin, 1, 0)
        out = torch.zeros(M, N, dtype=adt, device="cuda")
        st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
        gamma = torch.ones(Cin, device="cuda"); beta = torch.zeros(Cin, device="cuda")
        s1 = torch.zeros(Cin, dtype=torch.float64, device="cuda"); s2 = torch.full((Cin,), float(M), dtype=torch.float64, device="cuda")
        bn = H.bnsrc(s1, s2, gamma, beta, count=M)
        t0 = timeit(lambda: H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, ctot, bp, out, N))
        t2 = timeit(lambda: H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, ctot, bp, out, N, epi=L.EP_STORE_STATS, st_sum=st[0], st_sq=st[1]))
        t1 = timeit(lambda: H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, ctot, bp, out, N, trans=L.T_BNRELU, epi=L.EP_STORE_STATS, bnA=bn, st_sum=st[0], st_sq=st[1]))
        mb = (M * Cin * 2 + M * N * 2) / 1e6
        print(f"M={M} Cin={Cin}: raw/store {t0:.1f} us | raw/store+stats {t2:.1f} us | bnrelu/store+stats {t1:.1f} us | {mb:.0f} MB = {mb / 6547.5 * 1e3 / 1e3:.1f} us at the HBM peak")
