"""Debug driver of the all-TMA stem weight gradient (engine.cuh tma_a): one small launch, printed error vs the register path."""
import sys
import torch
sys.path.insert(0, ".")
from tests import engine_helpers as H

torch.manual_seed(17)
cin, (X, Y, Z), B = 2, (16, 16, 32), 3
img = torch.rand(B, cin, X, Y, Z, device="cuda")
Dz, Dy, Dx = (X - 1) // 2 + 1, (Y - 1) // 2 + 1, (Z - 1) // 2 + 1
Sz, Sy, Sx = Dz + 3, Dy + 3, Dx + 3
pad = torch.zeros(B, 2, 2 * Sz, 2 * Sy, 2 * Sx, device="cuda")
pad[:, :cin, 3:3 + X, 3:3 + Y, 3:3 + Z] = img
s2d = pad.view(B, 2, Sz, 2, Sy, 2, Sx, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).contiguous().to(H.act_dtype()).to(torch.bfloat16)
s2d = torch.cat([s2d.view(-1), torch.zeros(64, dtype=torch.bfloat16, device="cuda")])
M0 = B * Dz * Dy * Dx
dconv = (torch.randn(M0, 64, device="cuda") * 0.1).to(torch.bfloat16)
got = []
for a_bf16 in (2, 1):
    dw0 = torch.zeros(64, cin, 7, 7, 7, device="cuda")
    H.wgrad(3, M0, 64, 1, 128, 64, (Dz, Dy, Dx), s2d, 16, dconv, 64, dw0, 0, 0, sdims=(Sz, Sy, Sx), cin_real=cin, a_bf16=a_bf16, split=1)
    torch.cuda.synchronize()
    got.append(dw0)
    print("a_bf16", a_bf16, "ok", float(dw0.abs().max()))
print("max diff", float((got[0] - got[1]).abs().max()), "equal", torch.equal(got[0], got[1]))
