"""2-GPU check of the overlapped gradient all-reduce (run under torchrun on a GPU box; not collected by pytest):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_overlap_check.py
One training step from identical weights with (a) per-group all-reduce overlapped with backward and (b) one all-reduce
after backward must leave identical gradients (SUM over 2 ranks is order-independent), and the reduced trunk gradient
must equal the sum of the two ranks' local gradients."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mmnn_sts_b200 import distributed as D  # noqa: E402
from mmnn_sts_b200.losses.GradientBlender import GradientBlender  # noqa: E402
from mmnn_sts_b200.losses.losses import CoxPH  # noqa: E402
from mmnn_sts_b200.utils.utils import surv_criterion  # noqa: E402

rank, world, dev = D.init_from_env()
wl = bench.WORKLOADS["cfg1"]
model = bench.build_model(wl, dev)
for m in model.modules():
    if isinstance(m, torch.nn.Dropout) or hasattr(m, "dropout_prob"):
        pass
model.image_model.model.backbone.dropout_prob = 0.0       # deterministic forward: the two passes must see the same masks
model.eval(); model.train()
batch = bench.make_batches(wl, 1, device=dev, seed=1234 + 100 * rank)[0]
gb = GradientBlender(CoxPH, survival=True, surv_criterion=surv_criterion)


def one_pass(overlap):
    model.zero_grad(set_to_none=True)
    for mod in model.modules():
        if hasattr(mod, "grad_group_hook"):
            mod.grad_group_hook = None
    sync = D.GradientAllReducer(model.parameters(), model=model, overlap=overlap)
    torch.manual_seed(5)                                  # same MLP / feature dropout masks in both passes
    out = model({"image": batch[0], "clinical": batch[1]})
    loss, _ = gb.computeLoss(out, batch[2], batch[3])
    local = None
    if not overlap:
        loss.backward()
        local = model.image_model.model.backbone.flat_grad_buffer().clone()
        sync()
    else:
        sync.arm()
        loss.backward()
        sync()
    torch.cuda.synchronize()
    return torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).clone(), local


g_plain, local = one_pass(False)
g_over, _ = one_pass(True)
both = [torch.empty_like(local) for _ in range(world)]
dist.all_gather(both, local)
trunk = model.image_model.model.backbone.flat_grad_buffer()
d1 = float((g_plain - g_over).abs().max())
d2 = float((trunk - sum(both)).abs().max())
groups = model.image_model.model.backbone.grad_groups()
if rank == 0:
    print(f"dp_overlap_check: |plain - overlapped| max {d1:.3e}; |reduced - sum of locals| max {d2:.3e}; groups {groups}; total {trunk.numel()}")
    assert groups[-1][0] == 0 and max(h for _, h in groups) == trunk.numel() and sum(h - l for l, h in groups) == trunk.numel()
assert d1 == 0.0 or d1 < 1e-6 * float(g_plain.abs().max()), d1     # atomics make the local gradients run-dependent in the last bits
assert d2 < 1e-5 * float(trunk.abs().max()) + 1e-6, d2
dist.barrier()
dist.destroy_process_group()
