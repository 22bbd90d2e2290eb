"""Timeline of ONE configs[1] training step with the real stream structure (main + weight-gradient side stream, early start, PDL):
python profiles/timeline.py [cfg2] > gpurun_out/timeline.txt.  Prints every launch's class, start / end (ms from the first launch)
and, at the end, the busy time per class, the span of the step and the gaps on the main chain."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmnn_sts_b200 import _lib as L  # noqa: E402
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


for _ in range(3):
    step()
torch.cuda.synchronize()
L.lib().mmnn_profile_enable(2)
step()
torch.cuda.synchronize()
tl = L.profile_timeline()
L.lib().mmnn_profile_enable(0)
side = {"conv2_wgrad", "conv1_wgrad", "trans_wgrad", "tails"}
for name, a, b in tl:
    print(f"{a:9.4f} {b:9.4f} {b - a:8.4f} {'S' if name in side else 'M'} {name}")
span = max(b for _, _, b in tl)
busy = {}
for name, a, b in tl:
    busy[name] = busy.get(name, 0.0) + (b - a)
print("# span of the step (first launch start -> last launch end): %.3f ms" % span)
main = sorted([(a, b, n) for n, a, b in tl if n not in side])
gap = sum(max(0.0, main[i + 1][0] - main[i][1]) for i in range(len(main) - 1))
print("# main-stream chain: busy %.3f ms, gaps between consecutive launches %.3f ms" % (sum(b - a for a, b, _ in main), gap))
sd = sorted([(a, b, n) for n, a, b in tl if n in side])
print("# side stream: busy %.3f ms, first start %.3f, last end %.3f" % (sum(b - a for a, b, _ in sd), sd[0][0], max(b for _, b, _ in sd)))
for k, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    print("# %-14s %.3f ms" % (k, v))
