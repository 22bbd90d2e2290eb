"""Micro-benchmark of the block-1 / block-2 weight-gradient launches of configs[1] (standalone, warm), with the kernel's
ablation switches (WgradParams.cin_real for the non-stem kinds: 1 = producers skip the A stores, 2 = no B boxes, 3 = both, 4 = no MMA)."""
import sys
import torch
sys.path.insert(0, ".")
from tests import engine_helpers as H


def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


adt = H.act_dtype()
for name, B, dims in (("block1", 16, (32, 32, 16)), ("block2", 16, (16, 16, 8))):
    Dz, Dy, Dx = dims
    M = B * Dz * Dy * Dx
    bott = torch.randn(M, 128, device="cuda").to(adt)
    g = (torch.randn(M, 32, device="cuda") * 0.1).to(torch.bfloat16)
    gamma = torch.ones(128, device="cuda"); beta = torch.zeros(128, device="cuda")
    s1 = torch.zeros(128, dtype=torch.float64, device="cuda"); s2 = torch.full((128,), float(M), dtype=torch.float64, device="cuda")
    bn = H.bnsrc(s1, s2, gamma, beta, count=M)
    dwt = torch.zeros(27, 32, 128, device="cuda")
    for dbg in (0, 1, 2, 3, 4):
        t = timeit(lambda: H.wgrad(1, M, 32, 9, 128, 32, dims, bott, 128, g, 32, dwt, 1, 128, 32 * 128, bnA=bn, cin_real=dbg))
        print(f"{name} conv2 wgrad M={M} dbg={dbg}: {t:.1f} us")
    # 1x1x1: A = block buffer (cin channels of ctot), B = dBott (128)
    for cin, ctot in ((224, 256), (64, 256)):
        buf = torch.randn(M, ctot, device="cuda").to(adt)
        db = (torch.randn(M, 128, device="cuda") * 0.1).to(torch.bfloat16)
        gm = torch.ones(cin, device="cuda"); bt = torch.zeros(cin, device="cuda")
        t1 = torch.zeros(cin, dtype=torch.float64, device="cuda"); t2 = torch.full((cin,), float(M), dtype=torch.float64, device="cuda")
        bn1 = H.bnsrc(t1, t2, gm, bt, count=M)
        dw = torch.zeros(128, cin, device="cuda")
        for dbg in (0, 1, 2, 3, 4):
            t = timeit(lambda: H.wgrad(0, M, 128, 1, cin, 128, dims, buf, ctot, db, 128, dw, 1, cin, bnA=bn1, cin_real=dbg))
            print(f"{name} conv1 wgrad M={M} cin={cin} dbg={dbg}: {t:.1f} us")

# debug build only (MMNN_EXTRA_NVCC_FLAGS=-DMMNN_WGRAD_TIMING python -m mmnn_sts_b200.build --force): per-role cycle counters
import ctypes as C
from mmnn_sts_b200 import _lib as L
if hasattr(L.lib(), "mmnn_wgrad_dbg"):
    fn = L.lib().mmnn_wgrad_dbg
    fn.argtypes = [C.c_void_p, C.c_int]
    B, dims = 16, (32, 32, 16)
    M = B * dims[0] * dims[1] * dims[2]
    bott = torch.randn(M, 128, device="cuda").to(adt)
    g = (torch.randn(M, 32, device="cuda") * 0.1).to(torch.bfloat16)
    gamma = torch.ones(128, device="cuda"); beta = torch.zeros(128, device="cuda")
    s1 = torch.zeros(128, dtype=torch.float64, device="cuda"); s2 = torch.full((128,), float(M), dtype=torch.float64, device="cuda")
    bn = H.bnsrc(s1, s2, gamma, beta, count=M)
    dwt = torch.zeros(27, 32, 128, device="cuda")
    for dbg in (0, 3):
        H.wgrad(1, M, 32, 9, 128, 32, dims, bott, 128, g, 32, dwt, 1, 128, 32 * 128, bnA=bn, cin_real=dbg)
        torch.cuda.synchronize()
        fn(None, 1)
        H.wgrad(1, M, 32, 9, 128, 32, dims, bott, 128, g, 32, dwt, 1, 128, 32 * 128, bnA=bn, cin_real=dbg)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 12)()
        fn(out, 0)
        v = list(out)
        n = max(v[6], 1)
        print(f"dbg={dbg}: per CTA cycles: mma wait full {v[0]/n:.0f} | mma issue {v[1]/n:.0f} | producer wait empty {v[2]/n:.0f} | producer stores "
              f"{v[3]/n:.0f} | fill latency (free -> full) {v[4]/n:.0f} | kernel {v[5]/n:.0f} | tiles/CTA {v[7]/n:.1f} || fence+elect {v[8]/n:.0f} | MMAs {v[9]/n:.0f} | commits {v[10]/n:.0f} | last commit -> accumulator complete {v[11]/n:.0f}")
    # same launch on 15 SMs only (split 5 x 3 tap groups): is the per-MMA issue time a chip-wide effect?
    for dbg in (0, 3):
        fn(None, 1)
        H.wgrad(1, M, 32, 9, 128, 32, dims, bott, 128, g, 32, dwt, 1, 128, 32 * 128, bnA=bn, cin_real=dbg, split=5)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 12)()
        fn(out, 0)
        v = list(out)
        n = max(v[6], 1)
        print(f"split 5, dbg={dbg}: per CTA cycles: mma wait full {v[0]/n:.0f} | mma issue {v[1]/n:.0f} | producer wait empty {v[2]/n:.0f} | producer stores "
              f"{v[3]/n:.0f} | fill latency {v[4]/n:.0f} | kernel {v[5]/n:.0f} | tiles/CTA {v[7]/n:.1f}")
    # ring depth sensitivity (block-1 3x3x3 weight gradient)
    for stages in (1, 2, 3):
        t = timeit(lambda: H.wgrad(1, M, 32, 9, 128, 32, dims, bott, 128, g, 32, dwt, 1, 128, 32 * 128, bnA=bn, stages=stages))
        print(f"block1 conv2 wgrad stages={stages}: {t:.1f} us")
