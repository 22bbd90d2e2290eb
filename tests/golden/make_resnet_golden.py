"""Generates tests/golden/resnet_{eval,train}.npz from the UNCHANGED reference module /root/reference/models/resnet.py
(run in the build container only: the reference does not travel to the GPU box).

    python tests/golden/make_resnet_golden.py

Inputs and weights come from oracle.resnet.make_state_dict / make_batch (seeded CPU generators), so the fixtures hold
only the reference's OUTPUTS: sigmoid scores, the BCE-with-logits 'sum' loss the reference's training loop applies to
them, a few gradients and updated running statistics.  Train mode is run with `model.dropout.p = 0` (an attribute of the
constructed module, no reference code is modified): torch's Dropout draws its own mask, which no other implementation
can reproduce; dropout is tested separately with injected masks against the restatement."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")

from models.resnet import r3d_18  # noqa: E402  (the reference)
from oracle import resnet as orn  # noqa: E402

NUM_CLASSES, SPATIAL, BATCH = 5, (12, 32, 24), 3
GRAD_KEYS = ["stem.0.weight", "stem.1.weight", "layer1.0.conv1.0.weight", "layer1.0.downsample.0.weight", "layer1.1.conv2.1.bias",
             "layer2.0.conv1.0.weight", "layer2.0.downsample.1.weight", "layer3.1.conv1.0.weight", "layer4.0.conv2.0.weight",
             "layer4.1.conv2.1.weight", "fc.weight", "fc.bias"]
STAT_KEYS = ["stem.1.running_mean", "stem.1.running_var", "layer2.0.downsample.1.running_var", "layer4.1.conv2.1.running_mean"]


def main():
    torch.set_num_threads(4)
    sd = orn.make_state_dict(7, NUM_CLASSES)
    image, labels = orn.make_batch(11, BATCH, SPATIAL, NUM_CLASSES)
    pos_weight = torch.linspace(0.5, 3.0, NUM_CLASSES)
    m = r3d_18(NUM_CLASSES)
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        out_eval = m(image)
    np.savez_compressed(os.path.join(HERE, "resnet_eval.npz"), out=out_eval.numpy())
    m.train()
    m.dropout.p = 0.0
    out = m(image)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=pos_weight, reduction="sum")(out, labels)
    loss.backward()
    named = dict(m.named_parameters())
    fx = {"out": out.detach().numpy(), "loss": np.float64(loss.item()), "pos_weight": pos_weight.numpy()}
    for k in GRAD_KEYS:
        fx["grad:" + k] = named[k].grad.numpy()
    msd = m.state_dict()
    for k in STAT_KEYS:
        fx["stat:" + k] = msd[k].numpy()
    fx["grad_norms"] = np.array([float(named[k].grad.norm()) for k in sorted(named)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "resnet_train.npz"), **fx)
    print("eval", out_eval[0], "\ntrain loss", loss.item())


if __name__ == "__main__":
    main()
