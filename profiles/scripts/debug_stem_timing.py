import ctypes as C, sys, torch
sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L
from tests import engine_helpers as H
lib = L.lib()
B, cin, X, Y, Z = 16, 2, 128, 128, 64
Dz, Dy, Dx = 64, 64, 32
s2d = torch.rand(B * (Dz + 3) * (Dy + 3) * (Dx + 3) * 16, device="cuda").to(H.act_dtype())
w = torch.randn(64, cin, 7, 7, 7, device="cuda") * 0.05
bp = H.pack(w, 64, 64, 64, 64, 16, 0, 0, 0, mode=L.PACK_STEM_SW32, cin_real=cin)
out = torch.zeros(B * Dz * Dy * Dx, 64, dtype=H.act_dtype(), device="cuda")
st = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
for _ in range(3):
    H.stem_brick(B, (Dz, Dy, Dx), s2d, bp, out, 64, st_sum=st[0], st_sq=st[1])
torch.cuda.synchronize()
buf = (C.c_longlong * 8)()
lib.mmnn_stem_dbg.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.mmnn_stem_dbg(buf))
w_acc, w_brick, total, tiles = buf[0], buf[1], buf[2], buf[3]
print(f"CTA 0: tiles {tiles}, MMA-warp loop {total} cycles ({total/tiles:.0f}/tile), waiting acc_empty {w_acc} ({100*w_acc/total:.1f} %), waiting brick_full {w_brick} ({100*w_brick/total:.1f} %)")
