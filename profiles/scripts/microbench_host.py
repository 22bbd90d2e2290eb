"""Host-side enqueue time of one training step vs its GPU time (is the step CPU-launch-bound?)."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from mmnn_sts_b200.losses.GradientBlender import GradientBlender
from mmnn_sts_b200.losses.losses import CoxPH
from mmnn_sts_b200.utils.utils import surv_criterion
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda", 0)
model = bench.build_model(wl, dev)
opt = torch.optim.SGD(model.parameters(), 5e-4, momentum=0.9, nesterov=True, weight_decay=1e-4)
gb = GradientBlender(CoxPH, survival=True, surv_criterion=surv_criterion)
b = bench.make_batches(wl, 1, device=dev)[0]
def step():
    t = [time.perf_counter()]
    out = model({"image": b[0], "clinical": b[1]}); t.append(time.perf_counter())
    loss, _ = gb.computeLoss(out, b[2], b[3]); t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    opt.step(); opt.zero_grad(set_to_none=True); t.append(time.perf_counter())
    return [1e3 * (t[i + 1] - t[i]) for i in range(4)]
for _ in range(5): step()
torch.cuda.synchronize()
acc = [0, 0, 0, 0]; n = 10
t0 = time.perf_counter()
for _ in range(n):
    d = step(); acc = [a + x for a, x in zip(acc, d)]
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue ms/step: fwd %.2f loss %.2f bwd %.2f opt %.2f | total %.2f ; wall incl. drain %.2f" % (*[a / n for a in acc], 1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n))
