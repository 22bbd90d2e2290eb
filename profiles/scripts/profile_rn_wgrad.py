"""Runs the three biggest ResNet weight-gradient launches once (for ncu).  Not a pytest test."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda", 0)
B = 8
st = torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else "wgrad"
for cin, cout, k, s, p, dims in [(64, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1), (258, 128, 32)), (8, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1), (258, 128, 32)),
                                 (1, 64, (1, 7, 7), (1, 2, 2), (1, 3, 3), (256, 256, 64))]:
    Do, Ho, Wo = [(dims[i] + 2 * p[i] - k[i]) // s[i] + 1 for i in range(3)]
    g = L.RnConvGeom(B, dims[0], dims[1], dims[2], cin, Do, Ho, Wo, cout, *k, *s, *p)
    f32 = cin == 1
    x = torch.rand((B,) + dims + (cin,), device=dev, dtype=torch.float32 if f32 else torch.float16)
    w = torch.randn((cout, cin) + k, device=dev) * 0.1
    dy = torch.randn((B, Do, Ho, Wo, cout), device=dev).bfloat16()
    y = torch.empty((B, Do, Ho, Wo, cout), device=dev, dtype=torch.float16)
    dx = torch.empty((B,) + dims + (cin,), device=dev, dtype=torch.bfloat16)
    dw = torch.zeros_like(w)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    if which == "wgrad":
        assert lib.mmnn_rn_conv_wgrad(C.byref(g), int(f32), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), st) == 0
    elif which == "fwd":
        assert lib.mmnn_rn_conv(C.byref(g), 0, int(f32), x.data_ptr(), w.data_ptr(), y.data_ptr(), None, stats.data_ptr(), st) == 0
    elif not f32:
        assert lib.mmnn_rn_conv(C.byref(g), 1, 0, dy.data_ptr(), w.data_ptr(), dx.data_ptr(), None, None, st) == 0
    torch.cuda.synchronize()
