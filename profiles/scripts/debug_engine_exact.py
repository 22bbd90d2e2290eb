import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L
from tests import engine_helpers as H
bf = lambda x: x.to(torch.bfloat16).float()
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
torch.manual_seed(0)
# 1. plain GEMM
M, Cin, N = 4096, 128, 128
a = torch.randn(M, Cin, device="cuda").to(torch.bfloat16)
w = torch.randn(N, Cin, device="cuda") * 0.2
bp = H.pack(w, N, 128, Cin, 64, 1, Cin, 1, 0)
out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, Cin, bp, out, N)
torch.cuda.synchronize()
ref = bf(a.double() @ bf(w).double().t())
print("gemm K=128: rel", rel(out.float(), ref), "exact frac", float((out.float() == ref.float()).float().mean()))
# 2. GEMM with BNRELU (batch stats)
x = a.float(); gamma = torch.rand(Cin, device="cuda") + 0.5; beta = torch.randn(Cin, device="cuda") * 0.3
s1 = x.double().sum(0); s2 = (x.double() ** 2).sum(0)
st = torch.zeros(2, N, dtype=torch.float64, device="cuda")
H.rows(M, 128, N, Cin, 64, 1, (1, 1, M), a, Cin, bp, out, N, trans=L.T_BNRELU, epi=L.EP_STORE_STATS, bnA=H.bnsrc(s1, s2, gamma, beta, count=M), st_sum=st[0], st_sq=st[1])
torch.cuda.synchronize()
aa = bf(F.relu(F.batch_norm(x, None, None, gamma, beta, True, 0.0, 1e-5)))
ref = bf(aa.double() @ bf(w).double().t())
print("gemm bnrelu: rel", rel(out.float(), ref), "exact frac", float((out.float() == ref.float()).float().mean()))
# 3. conv3
B, Dz, Dy, Dx, Cin, N = 2, 8, 8, 8, 128, 32
M = B * Dz * Dy * Dx
bott = torch.randn(M, Cin, device="cuda").to(torch.bfloat16)
w = torch.randn(N, Cin, 3, 3, 3, device="cuda") * 0.05
bp = H.pack(w, N, 32, Cin, 64, 27, Cin * 27, 27, 1)
out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
H.rows(M, 32, N, Cin, 64, 27, (Dz, Dy, Dx), bott, Cin, bp, out, N)
torch.cuda.synchronize()
x5 = bott.double().view(B, Dz, Dy, Dx, Cin).permute(0, 4, 1, 2, 3)
ref = bf(F.conv3d(x5, bf(w).double(), padding=1).permute(0, 2, 3, 4, 1).reshape(M, N))
print("conv3 K=3456: rel", rel(out.float(), ref), "exact frac", float((out.float() == ref.float()).float().mean()))
