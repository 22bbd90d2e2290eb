"""Micro-benchmark of small rows-kernel launches (latency analysis aid, not a test)."""
import sys
import torch
sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L
from tests import engine_helpers as H


def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


adt = H.act_dtype()
for M, Cin in ((512, 992), (512, 512), (512, 128), (4096, 992), (4096, 256)):
    N = 128
    a = torch.randn(M, 1024, device="cuda").to(adt)
    w = torch.randn(N, Cin, device="cuda") * 0.1
    bp = H.pack(w, N, 128, Cin, 64, 1, Cin, 1, 0)
    out = torch.zeros(M, N, dtype=adt, device="cuda")
    st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
    gamma = torch.ones(Cin, device="cuda"); beta = torch.zeros(Cin, device="cuda")
    s1 = torch.zeros(Cin, dtype=torch.float64, device="cuda"); s2 = torch.full((Cin,), float(M), dtype=torch.float64, device="cuda")
    bn = H.bnsrc(s1, s2, gamma, beta, count=M)
    t0 = timeit(lambda: H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, 1024, bp, out, N))
    t1 = timeit(lambda: H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, 1024, bp, out, N, trans=L.T_BNRELU, epi=L.EP_STORE_STATS, bnA=bn, st_sum=st[0], st_sq=st[1]))
    t2 = timeit(lambda: H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, 1024, bp, out, N, epi=L.EP_STORE_STATS, st_sum=st[0], st_sq=st[1]))
    print(f"M={M} Cin={Cin}: raw/store {t0:.1f} us | raw/store+stats {t2:.1f} us | bnrelu/store+stats {t1:.1f} us  (incl. ~2-3 us python/ctypes launch)")
empty = timeit(lambda: torch.empty(1, device="cuda").zero_())
print("tiny torch kernel loop:", round(empty, 1), "us")
