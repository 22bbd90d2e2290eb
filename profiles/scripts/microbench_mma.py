"""tcgen05.mma issue-rate microbenchmark (analysis aid, not a test): cycles per back-to-back MMA of shape 128 x N x 16
from shared-memory operands, SWIZZLE_NONE vs SWIZZLE_32B, with / without a moving A start address."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L

lib = L.lib()
lib.mmnn_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.mmnn_mma_rate.restype = C.c_int
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for layout in (0, 6):
    for step in (0, 176):
        row = []
        for N in (32, 64, 128, 256):
            best = None
            for rep in range(3):
                rc = lib.mmnn_mma_rate(N, layout, 2000, step, out.data_ptr(), None)
                assert rc == 0, rc
                torch.cuda.synchronize()
                c = int(out.item()) / 2000
                best = c if best is None else min(best, c)
            row.append(f"N={N}: {best:6.1f}")
        print(f"layout {'NONE' if layout == 0 else 'SW32'} a_step {step:4d} B | cycles per MMA  " + "  ".join(row))

# MN-major operands of the weight-gradient kernels (K = voxel rows): a_step = one K step of 16 rows (256 B planes / 1024 B swizzled)
for layout, step, name in ((100, 256, "A planes + B SW64 (wgrad)"), (101, 1024, "A, B SW64 (stem wgrad)")):
    row = []
    for N in (32, 64, 96, 128):
        best = None
        for rep in range(3):
            rc = lib.mmnn_mma_rate(N, layout, 2000, step, out.data_ptr(), None)
            assert rc == 0, rc
            torch.cuda.synchronize()
            c = int(out.item()) / 2000
            best = c if best is None else min(best, c)
        row.append(f"N={N}: {best:6.1f}")
    print(f"MN-major {name:28s} | cycles per MMA  " + "  ".join(row))

# accumulator dependency: back-to-back MMAs into 1, 2 or 4 rotating accumulators (K-major SWIZZLE_NONE and the MN-major wgrad operands)
lib.mmnn_mma_rate_acc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
for layout, step in ((0, 176), (100, 256)):
    for nacc in (1, 2, 4):
        row = []
        for N in (32, 64, 96, 128):
            if N * nacc > 512: continue
            best = None
            for rep in range(3):
                rc = lib.mmnn_mma_rate_acc(N, layout, 2000, step, nacc, out.data_ptr())
                assert rc == 0, rc
                torch.cuda.synchronize()
                c = int(out.item()) / 2000
                best = c if best is None else min(best, c)
            row.append(f"N={N}: {best:6.1f}")
        print(f"layout {layout:3d} accumulators {nacc} | cycles per MMA  " + "  ".join(row))

# the 3x3x3 weight gradient's exact issue pattern (3 accumulators x 8 K steps per tile, N = 96), without / with a commit per tile
for nacc, name in ((2, "no commit"), (1, "commit per tile")):
    best = None
    for rep in range(3):
        rc = lib.mmnn_mma_rate_acc(96, 102, 2400, 0, nacc, out.data_ptr())
        assert rc == 0, rc
        torch.cuda.synchronize()
        c = int(out.item()) / 2400
        best = c if best is None else min(best, c)
    print(f"wgrad 3x3x3 issue pattern, {name}: {best:.1f} cycles per MMA")
