"""Deterministic synthetic inputs and weights shared by the oracle, the golden generator, the tests and bench.py.

TEST/BENCH INFRASTRUCTURE.  Everything is drawn from a CPU torch.Generator so the same seed gives the same
tensors in the build container and on the GPU box (same torch build).  Input distributions follow SURVEY.md
section 8d: image U[0,1) (post-ScaleIntensity range, /root/reference/main.py:89), clinical = 10 N(0,1) columns
+ 10 integer category columns 0..4, events Bernoulli(0.5) int64, durations integer days in [1,3650].
Weights follow the reference initialisation LAW (/root/reference/models/densenet.py:258-265: kaiming_normal_
conv weights, BN gamma=1 beta=0, Linear bias 0, Linear weight torch default U(-1/sqrt(fan_in), 1/sqrt(fan_in)))
but are generated key by key so they do not depend on module construction order.
"""
import math
from collections import OrderedDict

import torch

BLOCK_CONFIG = (6, 12, 24, 16)


def state_dict_spec(in_channels=2, num_clinical=20, num_classes=2, num_features=12, init_features=64,
                    growth_rate=32, bn_size=4, block_config=BLOCK_CONFIG):
    """Ordered (key, shape, kind) list reproducing the reference MultiModalModel.state_dict() layout."""
    spec = []

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,), "ones"), (prefix + ".bias", (c,), "zeros"),
                     (prefix + ".running_mean", (c,), "zeros"), (prefix + ".running_var", (c,), "ones"),
                     (prefix + ".num_batches_tracked", (), "long0")])

    def lin(prefix, o, i, zero_bias=False):
        spec.extend([(prefix + ".weight", (o, i), "linear"), (prefix + ".bias", (o,), "zeros" if zero_bias else "linear_bias")])

    p = "image_model.model.backbone."
    spec.append((p + "conv0.weight", (init_features, in_channels, 7, 7, 7), "conv"))
    bn(p + "norm0", init_features)
    c = init_features
    for b, nl in enumerate(block_config):
        for l in range(nl):
            q = f"{p}denseblock{b + 1}.denselayer{l + 1}.layers."
            bn(q + "norm1", c)
            spec.append((q + "conv1.weight", (bn_size * growth_rate, c, 1, 1, 1), "conv"))
            bn(q + "norm2", bn_size * growth_rate)
            spec.append((q + "conv2.weight", (growth_rate, bn_size * growth_rate, 3, 3, 3), "conv"))
            c += growth_rate
        if b == len(block_config) - 1:
            bn(p + "norm5", c)
        else:
            q = f"{p}transition{b + 1}."
            bn(q + "norm", c)
            spec.append((q + "conv.weight", (c // 2, c, 1, 1, 1), "conv"))
            c //= 2
    lin("image_model.model.features.feature_layer", num_features, c, zero_bias=True)
    lin("image_model.model.class_layers.out", num_classes, num_features, zero_bias=True)
    m = "clinical_model.model."
    widths = [num_clinical, 32, 16, 8, 8, 8]
    for i in range(5):
        lin(f"{m}backbone.dense{i}", widths[i + 1], widths[i])
        bn(f"{m}backbone.bn{i}", widths[i + 1])
    lin(m + "features.dense5", num_features, 8)
    bn(m + "features.bn5", num_features)
    lin(m + "output_head.dense6", num_classes, num_features)
    lin("output_head", num_classes, 2 * num_features)
    lin("clinical_output_head", num_classes, num_features)
    lin("image_output_head", num_classes, num_features)
    return spec


def make_state_dict(seed=42, perturb_bn=True, **kw):
    """perturb_bn: draw BN gamma in [0.5,1.5], beta in [-0.3,0.3] and non-trivial running stats so that parity
    tests exercise the affine/eval paths (the reference init gamma=1, beta=0 would hide bugs there)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in state_dict_spec(**kw):
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "linear":
            bound = 1.0 / math.sqrt(shape[1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "linear_bias":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "ones":
            t = torch.ones(shape)
            if perturb_bn:
                if key.endswith("running_var"):
                    t = 0.5 + torch.rand(shape, generator=g)
                else:
                    t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape)
            if perturb_bn and (key.endswith("running_mean") or ".norm" in key or ".bn" in key):
                t = (torch.rand(shape, generator=g) - 0.5) * 0.6
        elif kind == "long0":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise ValueError(kind)
        sd[key] = t
    return sd


def make_batch(seed, batch, in_channels, spatial, num_clinical=20, num_classes=2, tie_free=True):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((batch, in_channels) + tuple(spatial), generator=g)
    cont = torch.randn((batch, num_clinical // 2), generator=g)
    cat = torch.randint(0, 5, (batch, num_clinical - num_clinical // 2), generator=g).float()
    clinical = torch.cat([cont, cat], 1)
    events = torch.randint(0, 2, (batch, num_classes), generator=g)
    if tie_free:
        durations = torch.stack([torch.randperm(3650, generator=g)[:batch] + 1 for _ in range(num_classes)], 1)
    else:
        durations = torch.randint(1, 3651, (batch, num_classes), generator=g)
    return image, clinical, events, durations


def make_masks(seed, batch, p_dense=0.2, p_feat=0.2, p_mlp=0.2, growth_rate=32, num_features=12,
               block_config=BLOCK_CONFIG):
    """Keep-masks already scaled by 1/(1-p) (see oracle/model.py)."""
    g = torch.Generator().manual_seed(seed)

    def keep(shape, p):
        if p <= 0:
            return torch.ones(shape)
        return (torch.rand(shape, generator=g) >= p).float() / (1 - p)

    return {"dense": {(b, l): keep((batch, growth_rate), p_dense) for b, nl in enumerate(block_config) for l in range(nl)},
            "image_features": keep((batch, num_features), p_feat),
            "mlp": [keep((batch,), p_mlp) for _ in range(6)]}
