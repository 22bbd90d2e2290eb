"""Rounding-matched CPU model of the sm_100a trunk: the SAME algorithm as oracle/model.py (i.e. the reference's,
/root/reference/models/densenet.py:46-148,196-231) with rounding inserted at exactly the points where the CUDA
build stores 16-bit values: forward tensors (weights, conv inputs after BN+ReLU, conv outputs, pooled tensors) in the
activation format (fp16), gradient tensors of the backward pass in bf16.  TEST INFRASTRUCTURE.

Why it exists: with reduced-precision storage a ReLU whose pre-activation is within rounding distance of zero flips,
and a flipped element changes its gradient by 100 %; gradient errors against the fp32 oracle therefore scale like
sqrt(forward error) and cannot separate "bf16 noise" from "kernel bug".  Against this model the kernels must agree
tightly (same masks, same roundings); against oracle/model.py (fp32) they must agree within the documented bf16
tolerance.  Rounding is straight-through in the forward direction and explicit in the backward direction."""
import torch
import torch.nn.functional as F


ACT_DTYPE = torch.float16   # storage format of forward activations / forward weights in the build (see common.cuh)


def _bf(x):
    return x.to(torch.bfloat16).float()


def _act(x):
    return x.to(ACT_DTYPE).float()


class _RoundSTE(torch.autograd.Function):
    """forward: round to the activation format; backward: identity."""

    @staticmethod
    def forward(ctx, x):
        return _act(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """forward: identity; backward: round the gradient to bf16 (a gradient tensor the CUDA build stores as bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf(g)


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _act(x)

    @staticmethod
    def backward(ctx, g):
        return _bf(g)


rs, rg, rb = _RoundSTE.apply, _RoundGrad.apply, _RoundBoth.apply


def _bn(x, sd, q, training):
    return F.batch_norm(x, sd[q + ".running_mean"].clone(), sd[q + ".running_var"].clone(), sd[q + ".weight"], sd[q + ".bias"], training, 0.1, 1e-5)


def backbone_bf16(sd, x, training, masks=None, prefix="", block_config=(6, 12, 24, 16), collect=None):
    p = prefix + "backbone."
    # conv0: image and weights rounded; output stored bf16; its gradient (dConv0) is stored bf16
    x = rb(F.conv3d(_act(x), rs(sd[p + "conv0.weight"]), None, stride=2, padding=3))
    if collect is not None: collect["conv0"] = x
    # norm0+relu0+pool0 fused; the masked pool gradient (dR) is stored bf16
    x = rs(F.max_pool3d(rg(F.relu(_bn(x, sd, p + "norm0", training))), 3, 2, 1))
    for b, nl in enumerate(block_config):
        for l in range(nl):
            q = f"{p}denseblock{b + 1}.denselayer{l + 1}.layers."
            a1 = rs(F.relu(_bn(x, sd, q + "norm1", training)))              # dA1 is never stored: the 1x1x1 dgrad epilogue adds it (fp32) to the block's accumulator
            bott = rb(F.conv3d(a1, rs(sd[q + "conv1.weight"])))             # bott bf16; dBott bf16
            a2 = rs(rg(F.relu(_bn(bott, sd, q + "norm2", training))))       # dA2 (masked) stored bf16
            y = F.conv3d(a2, rs(sd[q + "conv2.weight"]), padding=1)
            if masks is not None and masks.get("dense") is not None:
                y = y * masks["dense"][(b, l)][:, :, None, None, None]
            y = rb(y)                                                       # slice stored bf16; gslice bf16
            x = torch.cat([x, y], 1)
        if collect is not None: collect[f"block{b + 1}"] = x
        if b == len(block_config) - 1:
            x = _bn(x, sd, p + "norm5", training)
        else:
            q = f"{p}transition{b + 1}."
            pooled = rb(F.avg_pool3d(F.relu(_bn(x, sd, q + "norm", training)), 2, 2))   # pooled bf16; dpooled bf16
            x = rs(rg(F.conv3d(pooled, rs(sd[q + "conv.weight"]))))                     # next buffer bf16; gout bf16
    return x
