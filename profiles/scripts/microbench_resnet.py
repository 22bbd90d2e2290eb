"""Times the configs[3] ResNet classification train step (B 8, 1x256x256x64) per kernel class.  Not a pytest test."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L
from mmnn_sts_b200.models.resnet import r3d_18

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dims = (256, 256, 64)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = r3d_18(2).to(dev).train()
x = torch.rand((B, 1) + dims, device=dev)
y = (torch.rand((B, 2), device=dev) < 0.4).float()
pw = torch.tensor([1.5, 2.0], device=dev)


def step():
    out = m(x)
    loss = F.binary_cross_entropy_with_logits(out, y, pos_weight=pw, reduction="sum")
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
n = 3
for _ in range(n):
    step()
b.record()
torch.cuda.synchronize()
print(f"step {a.elapsed_time(b) / n:.2f} ms  -> {B * n / a.elapsed_time(b) * 1e3:.1f} volumes/s; peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
L.lib().mmnn_profile_enable(1)
step()
prof = L.profile_collect()
L.lib().mmnn_profile_enable(0)
for k, (ms, cnt) in prof.items():
    if cnt:
        print(f"  {k:12s} {ms:8.3f} ms  {cnt:4d} launches")
