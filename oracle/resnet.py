"""CPU restatement of the reference's 3-D ResNet encoder (BASELINE configs[3]).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Functional fp32 restatement of /root/reference/models/resnet.py over a plain state_dict (the reference's key names):
  stem            BasicStem :5-13       Conv3d(1->64, k (1,7,7), s (1,2,2), p (1,3,3), no bias) -> BN -> ReLU
  basic_block     BasicBlock :61-95     conv1(3x3x3, stride) -> BN -> ReLU -> conv2(3x3x3) -> BN; (+ downsample(x) | + x); ReLU
  downsample      Resnet18._make_layer :172-179   Conv3d 1x1x1 (stride) -> BN, present when stride != 1 or planes change
  resnet_forward  Resnet18.forward :154-170       stem -> layer1..4 with nn.Dropout after each -> avgpool -> flatten -> fc -> sigmoid
  state_dict      Resnet18._initialize_weights :189-203 (kaiming_normal_ fan_out/relu convs, BN 1/0, Linear N(0, 0.01), bias 0),
                  drawn key by key from a CPU generator (independent of module construction order)
  train_step_loss criterion(BCEWithLogitsLoss(pos_weight, 'sum'), model(x), labels) as /root/reference/main.py:148-153,208:
                  the loss is applied to the already-sigmoided output (the reference's own quirk, kept).
Pinned against the UNCHANGED reference module (importable without shims) by tests/golden/make_resnet_golden.py ->
tests/golden/resnet_*.npz (outputs, loss, gradients, running statistics in eval / train mode).
Dropout masks are injected (list of 4 keep-masks in NCDHW, one per stage) so a CUDA run can be compared element-wise.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

PLANES = (8, 16, 8, 16)          # /root/reference/models/resnet.py:134-137
STRIDES = (1, 2, 2, 2)
BLOCKS = (2, 2, 2, 2)            # r3d_18 :221


def state_dict_spec(num_classes):
    spec = []

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,), "ones"), (prefix + ".bias", (c,), "zeros"),
                     (prefix + ".running_mean", (c,), "zeros"), (prefix + ".running_var", (c,), "ones"),
                     (prefix + ".num_batches_tracked", (), "long0")])

    spec.append(("stem.0.weight", (64, 1, 1, 7, 7), "conv"))
    bn("stem.1", 64)
    inplanes = 64
    for li, (planes, stride, nb) in enumerate(zip(PLANES, STRIDES, BLOCKS)):
        for b in range(nb):
            p = f"layer{li + 1}.{b}."
            cin = inplanes if b == 0 else planes
            spec.append((p + "conv1.0.weight", (planes, cin, 3, 3, 3), "conv"))
            bn(p + "conv1.1", planes)
            spec.append((p + "conv2.0.weight", (planes, planes, 3, 3, 3), "conv"))
            bn(p + "conv2.1", planes)
            if b == 0 and (stride != 1 or inplanes != planes):
                spec.append((p + "downsample.0.weight", (planes, inplanes, 1, 1, 1), "conv"))
                bn(p + "downsample.1", planes)
        inplanes = planes
    spec.append(("fc.weight", (num_classes, 16), "fc"))
    spec.append(("fc.bias", (num_classes,), "zeros"))
    return spec


def make_state_dict(seed, num_classes, perturb_bn=True):
    """Reference initialisation law; with perturb_bn the BN affine / running statistics are moved off 1/0 so that parity
    tests exercise them (a trained checkpoint looks like that)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in state_dict_spec(num_classes):
        if kind == "conv":
            fan_out = shape[0] * shape[2] * shape[3] * shape[4]
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif kind == "fc":
            sd[key] = torch.randn(shape, generator=g) * (0.01 if not perturb_bn else 0.5)
        elif kind == "ones":
            sd[key] = torch.ones(shape) + (0.2 * torch.randn(shape, generator=g) if perturb_bn else 0.0)
            if key.endswith("running_var"):
                sd[key] = sd[key].abs() + 0.1
        elif kind == "zeros":
            sd[key] = torch.zeros(shape) + (0.1 * torch.randn(shape, generator=g) if perturb_bn and not key.startswith("fc") else 0.0)
        elif kind == "long0":
            sd[key] = torch.zeros(shape, dtype=torch.long)
    return sd


def make_batch(seed, batch, spatial, num_classes):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((batch, 1) + tuple(spatial), generator=g)
    labels = (torch.rand((batch, num_classes), generator=g) < 0.4).float()
    return image, labels


def _bn(p, prefix, x, training):
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".weight"], p[prefix + ".bias"],
                        training, 0.1, 1e-5)


def basic_block(p, prefix, x, stride, training):
    out = F.relu(_bn(p, prefix + "conv1.1", F.conv3d(x, p[prefix + "conv1.0.weight"], None, stride, 1), training))
    out = _bn(p, prefix + "conv2.1", F.conv3d(out, p[prefix + "conv2.0.weight"], None, 1, 1), training)
    residual = x
    if prefix + "downsample.0.weight" in p:
        residual = _bn(p, prefix + "downsample.1", F.conv3d(x, p[prefix + "downsample.0.weight"], None, stride, 0), training)
    return F.relu(out + residual)


def resnet_forward(p, image, training, dropout_p=0.2, masks=None, taps=None):
    """p: dict of tensors (parameters may require grad; running statistics are updated in place in training mode, as
    nn.BatchNorm3d does).  masks: None (no dropout) or 4 keep-masks broadcastable to the stage outputs."""
    x = F.relu(_bn(p, "stem.1", F.conv3d(image, p["stem.0.weight"], None, (1, 2, 2), (1, 3, 3)), training))
    if taps is not None:
        taps["stem"] = x
    for li, (stride, nb) in enumerate(zip(STRIDES, BLOCKS)):
        for b in range(nb):
            x = basic_block(p, f"layer{li + 1}.{b}.", x, stride if b == 0 else 1, training)
        if training and masks is not None:
            x = x * masks[li] / (1.0 - dropout_p)
        if taps is not None:
            taps[f"layer{li + 1}"] = x
    x = F.adaptive_avg_pool3d(x, 1).flatten(1)
    return torch.sigmoid(F.linear(x, p["fc.weight"], p["fc.bias"]))


def train_step_loss(out, labels, pos_weight):
    """/root/reference/main.py:148-153,208 + utils/utils.py criterion: BCEWithLogitsLoss(pos_weight, reduction='sum')(out, labels)."""
    return F.binary_cross_entropy_with_logits(out, labels, pos_weight=pos_weight, reduction="sum")


def stage_shapes(batch, spatial):
    """Shapes [B, C, D, H, W] of the four stage outputs (where Dropout acts) for an input [B, 1, *spatial]."""
    d, h, w = spatial
    d, h, w = d + 2, (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
    shapes = []
    for planes, stride in zip(PLANES, STRIDES):
        if stride != 1:
            d, h, w = (d - 1) // 2 + 1, (h - 1) // 2 + 1, (w - 1) // 2 + 1
        shapes.append((batch, planes, d, h, w))
    return shapes
