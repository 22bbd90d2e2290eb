"""Times every conv geometry of r3d_18 at configs[3] (B 8, 1x256x256x64): forward, data gradient, weight gradient.  Not a pytest test."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda", 0)
B = 8
GEOMS = [  # name, cin, cout, k, s, p, input dims
    ("stem", 1, 64, (1, 7, 7), (1, 2, 2), (1, 3, 3), (256, 256, 64)),
    ("l1 64->8 3^3", 64, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1), (258, 128, 32)),
    ("l1 64->8 1^3", 64, 8, (1, 1, 1), (1, 1, 1), (0, 0, 0), (258, 128, 32)),
    ("l1 8->8 3^3", 8, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1), (258, 128, 32)),
    ("l2 8->16 3^3 s2", 8, 16, (3, 3, 3), (2, 2, 2), (1, 1, 1), (258, 128, 32)),
    ("l2 8->16 1^3 s2", 8, 16, (1, 1, 1), (2, 2, 2), (0, 0, 0), (258, 128, 32)),
    ("l2 16->16 3^3", 16, 16, (3, 3, 3), (1, 1, 1), (1, 1, 1), (129, 64, 16)),
    ("l3 16->8 3^3 s2", 16, 8, (3, 3, 3), (2, 2, 2), (1, 1, 1), (129, 64, 16)),
    ("l3 8->8 3^3", 8, 8, (3, 3, 3), (1, 1, 1), (1, 1, 1), (65, 32, 8)),
]
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name, cin, cout, k, s, p, dims in GEOMS:
    Do, Ho, Wo = [(dims[i] + 2 * p[i] - k[i]) // s[i] + 1 for i in range(3)]
    g = L.RnConvGeom(B, dims[0], dims[1], dims[2], cin, Do, Ho, Wo, cout, *k, *s, *p)
    f32 = cin == 1
    x = torch.rand((B,) + dims + (cin,), device=dev, dtype=torch.float32 if f32 else torch.float16)
    w = torch.randn((cout, cin) + k, device=dev) * 0.1
    y = torch.empty((B, Do, Ho, Wo, cout), device=dev, dtype=torch.float16)
    dy = torch.randn((B, Do, Ho, Wo, cout), device=dev).bfloat16()
    dx = torch.empty((B,) + dims + (cin,), device=dev, dtype=torch.bfloat16)
    dw = torch.zeros_like(w)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    gf = 2.0 * B * Do * Ho * Wo * cout * cin * k[0] * k[1] * k[2] / 1e9
    tf = timeit(lambda: lib.mmnn_rn_conv(C.byref(g), 0, int(f32), x.data_ptr(), w.data_ptr(), y.data_ptr(), None, stats.data_ptr(), st))
    td = float("nan") if f32 else timeit(lambda: lib.mmnn_rn_conv(C.byref(g), 1, 0, dy.data_ptr(), w.data_ptr(), dx.data_ptr(), None, None, st))
    tw = timeit(lambda: lib.mmnn_rn_conv_wgrad(C.byref(g), int(f32), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), st))
    mb = (x.numel() * x.element_size() + y.numel() * 2) / 1e6
    print(f"{name:18s} {gf:7.1f} GF  {mb:7.0f} MB | fwd {tf:7.3f} ms ({gf / tf:6.1f} TF/s, {mb / tf:6.0f} GB/s) | dgrad {td:7.3f} ms | wgrad {tw:7.3f} ms ({gf / tw:6.1f} TF/s)")
