"""Prints per-parameter gradient errors of the ResNet path against the oracle (debug helper, not a pytest test)."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import resnet as orn
from mmnn_sts_b200.models.resnet import r3d_18

NUM_CLASSES, SPATIAL, BATCH = 5, (12, 32, 24), 3
dev = torch.device("cuda", 0)
sd = orn.make_state_dict(7, NUM_CLASSES)
image, labels = orn.make_batch(11, BATCH, SPATIAL, NUM_CLASSES)
pos_weight = torch.linspace(0.5, 3.0, NUM_CLASSES)
m = r3d_18(NUM_CLASSES)
m.load_state_dict(sd)
m = m.to(dev).train()
m.dropout.p = 0.0
out = m(image.to(dev))
loss = F.binary_cross_entropy_with_logits(out, labels.to(dev), pos_weight=pos_weight.to(dev), reduction="sum")
loss.backward()
p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
ref = orn.resnet_forward(p, image, training=True, masks=None)
ref_loss = orn.train_step_loss(ref, labels, pos_weight)
ref_loss.backward()
print("out err", float((out.detach().cpu() - ref.detach()).abs().max()), "loss", loss.item(), ref_loss.item())
for k, q in m.named_parameters():
    g, r = q.grad.double().cpu(), p[k].grad.double()
    print(f"{k:36s} rel {float((g - r).norm() / (r.norm() + 1e-30)):.3e}  |ref| {float(r.norm()):.3e}  cos {float((g * r).sum() / (g.norm() * r.norm() + 1e-30)):.5f}")
