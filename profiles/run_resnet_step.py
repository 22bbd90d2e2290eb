"""One configs[3] ResNet classification training step bracketed by cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmnn_sts_b200 import ops  # noqa: E402
from mmnn_sts_b200.models.resnet import r3d_18  # noqa: E402
from mmnn_sts_b200.optim import SGD  # noqa: E402

wl = bench.WORKLOADS["cfg4"]
dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = r3d_18(2).to(dev).train()
opt = SGD(m.parameters(), 5e-4, momentum=0.9, nesterov=True, weight_decay=1e-4)
x = torch.rand((wl["batch"], 1) + wl["spatial"], device=dev)
y = (torch.rand((wl["batch"], 2), device=dev) < 0.4).float()
pw = torch.tensor([1.5, 2.0], device=dev)


def step():
    loss = ops.bce_with_logits(m(x), y, pw).sum()
    loss.backward()
    opt.step(); opt.zero_grad(set_to_none=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
