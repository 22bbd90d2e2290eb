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
