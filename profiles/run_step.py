"""One training step of the bench workload bracketed by cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmnn_sts_b200.losses.GradientBlender import GradientBlender  # noqa: E402
from mmnn_sts_b200.losses.losses import CoxPH  # noqa: E402
from mmnn_sts_b200.optim import SGD  # noqa: E402
from mmnn_sts_b200.utils.utils import surv_criterion  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda", 0)
model = bench.build_model(wl, dev)
opt = SGD(model.parameters(), 5e-4, momentum=0.9, nesterov=True, weight_decay=1e-4)
gb = GradientBlender(CoxPH, survival=True, surv_criterion=surv_criterion)
batch = bench.make_batches(wl, 1, device=dev)[0]


def step():
    out = model({"image": batch[0], "clinical": batch[1]})
    loss, _ = gb.computeLoss(out, batch[2], batch[3])
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
