"""3-D ResNet encoder of BASELINE configs[3] on the sm_100a direct-convolution kernels (csrc/resnet.cu).

Mirror of /root/reference/models/resnet.py: `r3d_18(num_classes)` (:205-226) builds `Resnet18(BasicBlock, [Conv3DSimple] * 4,
[2, 2, 2, 2], BasicStem)` (:112-203) with the same module tree, so the 80 706 parameters / buffers carry the reference's
state_dict names (`stem.0.weight`, `layer1.0.conv1.0.weight`, `layer2.0.downsample.1.running_mean`, `fc.bias`, ...) and the
reference's `_initialize_weights` (:189-203).  The nn.Conv3d / nn.BatchNorm3d / nn.Linear children are parameter
containers only: `forward` runs the whole network (stem -> 4 stages of 2 BasicBlocks with element-wise Dropout after each
stage -> AdaptiveAvgPool3d(1) -> fc -> sigmoid, :154-170) as one autograd function over the C-ABI kernels, activations in
channels-last fp16, gradients in channels-last bf16.  No CPU path.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib as L
from ..ops import _p, _require_cuda, _stream


class BasicStem(nn.Sequential):
    """/root/reference/models/resnet.py:5-13 (note the depth padding of 1 under a depth kernel of 1: D grows by 2)."""

    def __init__(self):
        super().__init__(
            nn.Conv3d(1, 64, kernel_size=(1, 7, 7), stride=(1, 2, 2), padding=(1, 3, 3), bias=False),
            nn.BatchNorm3d(64),
            nn.ReLU(inplace=True))


class Conv3DSimple(nn.Conv3d):
    """/root/reference/models/resnet.py:97-114."""

    def __init__(self, in_planes, out_planes, midplanes=None, stride=1, padding=1):
        super().__init__(in_channels=in_planes, out_channels=out_planes, kernel_size=(3, 3, 3), stride=stride,
                         padding=padding, bias=False)

    @staticmethod
    def get_downsample_stride(stride):
        return stride, stride, stride


class BasicBlock(nn.Module):
    """/root/reference/models/resnet.py:61-95 (parameter container; executed by _ResnetFn)."""
    expansion = 1

    def __init__(self, inplanes, planes, conv_builder, stride=1, downsample=None):
        midplanes = (inplanes * planes * 3 * 3 * 3) // (inplanes * 3 * 3 + 3 * planes)
        super().__init__()
        self.conv1 = nn.Sequential(conv_builder(inplanes, planes, midplanes, stride), nn.BatchNorm3d(planes),
                                   nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(conv_builder(planes, planes, midplanes), nn.BatchNorm3d(planes))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


def _geom(conv, N, dims):
    """RnConvGeom for nn.Conv3d `conv` applied to a channels-last [N, D, H, W, C_in] tensor."""
    kd, kh, kw = conv.kernel_size
    sd, sh, sw = conv.stride
    pd, ph, pw = conv.padding
    Di, Hi, Wi = dims
    Do, Ho, Wo = (Di + 2 * pd - kd) // sd + 1, (Hi + 2 * ph - kh) // sh + 1, (Wi + 2 * pw - kw) // sw + 1
    return L.RnConvGeom(N, Di, Hi, Wi, conv.in_channels, Do, Ho, Wo, conv.out_channels, kd, kh, kw, sd, sh, sw, pd, ph, pw)


class _Unit:
    """One conv + BatchNorm pair of the tape: input tensor, geometry, raw conv output, BN coefficient table."""
    __slots__ = ("conv", "bn", "x", "x_f32", "g", "raw", "coef", "count")


def _new_unit(conv, bn, x, x_f32, N, dims):
    u = _Unit()
    u.conv, u.bn, u.x, u.x_f32 = conv, bn, x, x_f32
    u.g = _geom(conv, N, dims)
    g = u.g
    u.raw = torch.empty((N, g.Do, g.Ho, g.Wo, g.Cout), dtype=torch.float16, device=x.device)
    u.count = float(N * g.Do * g.Ho * g.Wo)
    return u, torch.zeros((2 * g.Cout,), dtype=torch.float64, device=x.device)


def _bn_table(u, stats, training, momentum_default=0.1):
    bn, Cout, dev = u.bn, u.g.Cout, u.raw.device
    u.coef = torch.empty((4, Cout), dtype=torch.float32, device=dev)
    mom = bn.momentum if bn.momentum is not None else momentum_default
    L.check(L.lib().mmnn_rn_bn_coeffs(_p(stats), u.count, _p(bn.weight.detach()), _p(bn.bias.detach()), _p(bn.running_mean),
                                      _p(bn.running_var), _p(bn.num_batches_tracked), bn.eps, mom, 1 if training else 0, Cout,
                                      _p(u.coef), _stream()), "mmnn_rn_bn_coeffs")
    return u


def _conv_bn_forward_pair(blk, x, N, dims, training):
    """conv1 and the 1x1x1 down-sample convolution of a block in one launch when the C ABI has the fused kernel for the
    geometry (layer1.0: 64 -> 8), else None."""
    if blk.downsample is None or tuple(blk.downsample[0].stride) != (1, 1, 1):
        return None
    u1, st1 = _new_unit(blk.conv1[0], blk.conv1[1], x, False, N, dims)
    ud, std = _new_unit(blk.downsample[0], blk.downsample[1], x, False, N, dims)
    rc = L.lib().mmnn_rn_conv_fwd_ds(C.byref(u1.g), _p(x), _p(u1.conv.weight.detach()), _p(ud.conv.weight.detach()), _p(u1.raw),
                                     _p(ud.raw), _p(st1), _p(std), _stream())
    if rc == -9:
        return None
    L.check(rc, "mmnn_rn_conv_fwd_ds")
    return _bn_table(u1, st1, training), _bn_table(ud, std, training)


def _conv_bn_forward(conv, bn, x, x_f32, N, dims, training):
    u, stats = _new_unit(conv, bn, x, x_f32, N, dims)
    L.check(L.lib().mmnn_rn_conv(C.byref(u.g), 0, 1 if x_f32 else 0, _p(x), _p(conv.weight.detach()), _p(u.raw), None, _p(stats),
                                 _stream()), "mmnn_rn_conv (forward)")
    return _bn_table(u, stats, training)


def _bn_act(u, res_mode, res, coef2, relu, drop_p, seed, mask):
    y = torch.empty_like(u.raw)
    L.check(L.lib().mmnn_rn_bn_act(_p(u.raw), _p(u.coef), res_mode, _p(res), _p(coef2), _p(y), y.numel(), y.shape[-1],
                                   1 if relu else 0, float(drop_p), int(seed), _p(mask), _stream()), "mmnn_rn_bn_act")
    return y


class _ResnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, image, *params):
        _require_cuda(image, "Resnet18")
        lib = L.lib()
        training = model.training
        N = image.shape[0]
        if image.dim() != 5 or image.shape[1] != 1:
            # the reference's stem is Conv3d(1, 64, ...) (/root/reference/models/resnet.py:10): same error class as torch's
            raise RuntimeError(f"Resnet18 expects input [B, 1, D, H, W], got {tuple(image.shape)}")
        x = image.detach().contiguous().float()                  # NCDHW with C = 1 is NDHWC already
        dims = tuple(image.shape[2:])
        p = model.dropout.p if training else 0.0
        tape = []
        with torch.cuda.device(image.device):
            u0 = _conv_bn_forward(model.stem[0], model.stem[1], x, True, N, dims, training)
            a = _bn_act(u0, 0, None, None, True, 0.0, 0, None)
            tape.append(("stem", u0, a))
            dims = (u0.g.Do, u0.g.Ho, u0.g.Wo)
            for li, layer in enumerate((model.layer1, model.layer2, model.layer3, model.layer4)):
                for bi, blk in enumerate(layer):
                    last = bi == len(layer) - 1
                    pair = _conv_bn_forward_pair(blk, a, N, dims, training)
                    u1 = pair[0] if pair else _conv_bn_forward(blk.conv1[0], blk.conv1[1], a, False, N, dims, training)
                    a1 = _bn_act(u1, 0, None, None, True, 0.0, 0, None)
                    d1 = (u1.g.Do, u1.g.Ho, u1.g.Wo)
                    u2 = _conv_bn_forward(blk.conv2[0], blk.conv2[1], a1, False, N, d1, training)
                    drop = p if last else 0.0
                    mask = None
                    if last and training and model.injected_masks is not None:
                        mask, drop = model.injected_masks[li].to(device=a.device, dtype=torch.uint8).contiguous(), p
                    seed = model._next_seed() if drop > 0 else 0
                    ud = None
                    if blk.downsample is not None:
                        ud = pair[1] if pair else _conv_bn_forward(blk.downsample[0], blk.downsample[1], a, False, N, dims, training)
                        y = _bn_act(u2, 2, ud.raw, ud.coef, True, drop, seed, mask)
                    else:
                        y = _bn_act(u2, 1, a, None, True, drop, seed, mask)
                    scale = 1.0 / (1.0 - drop) if drop > 0 else 1.0
                    tape.append(("block", u1, a1, u2, ud, y, scale))
                    a, dims = y, d1
            Cl = a.shape[-1]
            V = dims[0] * dims[1] * dims[2]
            K = model.fc.out_features
            pooled = torch.empty((N, Cl), dtype=torch.float32, device=a.device)
            out = torch.empty((N, K), dtype=torch.float32, device=a.device)
            L.check(lib.mmnn_rn_head_fwd(_p(a), N, V, Cl, _p(model.fc.weight.detach()), _p(model.fc.bias.detach()), K,
                                         _p(pooled), _p(out), _stream()), "mmnn_rn_head_fwd")
        if any(ctx.needs_input_grad):
            ctx.tape, ctx.model, ctx.head = tape, model, (pooled, out, a.shape, V)
            ctx.training = training
        return out

    @staticmethod
    def _bn_backward(u, dy, y, scale, ud, want_dz, grads, training):
        """Backward through [ReLU(+Dropout scale)] o (BN(u.raw) [+ BN(ud.raw) | + identity]): returns (draw, draw_ds, dz)."""
        lib = L.lib()
        Cc = y.shape[-1]
        dev = y.device
        two = ud is not None
        sums = torch.zeros((3 * Cc,), dtype=torch.float64, device=dev)
        L.check(lib.mmnn_rn_act_bwd_reduce(_p(dy), _p(y), scale, _p(u.raw), _p(u.coef), _p(ud.raw) if two else None,
                                           _p(ud.coef) if two else None, _p(sums), y.numel(), Cc, _stream()),
                "mmnn_rn_act_bwd_reduce")
        draw = torch.empty_like(y, dtype=torch.bfloat16)
        draw2 = torch.empty_like(draw) if two else None
        dz = torch.empty_like(draw) if want_dz else None
        dg, db = torch.empty((Cc,), device=dev), torch.empty((Cc,), device=dev)
        dg2, db2 = (torch.empty((Cc,), device=dev), torch.empty((Cc,), device=dev)) if two else (None, None)
        L.check(lib.mmnn_rn_bn_bwd_apply(_p(dy), _p(y), scale, _p(u.raw), _p(u.coef), _p(ud.raw) if two else None,
                                         _p(ud.coef) if two else None, _p(sums), 1.0 / u.count, 0 if training else 1,
                                         _p(draw), _p(draw2), _p(dz), _p(dg), _p(db), _p(dg2), _p(db2), y.numel(), Cc,
                                         _stream()), "mmnn_rn_bn_bwd_apply")
        grads[u.bn.weight], grads[u.bn.bias] = dg, db
        if two:
            grads[ud.bn.weight], grads[ud.bn.bias] = dg2, db2
        return draw, draw2, dz

    @staticmethod
    def _wgrad(u, draw, grads):
        dw = torch.zeros_like(u.conv.weight, dtype=torch.float32)
        L.check(L.lib().mmnn_rn_conv_wgrad(C.byref(u.g), 1 if u.x_f32 else 0, _p(u.x), _p(draw), _p(dw), _stream()),
                "mmnn_rn_conv_wgrad")
        grads[u.conv.weight] = dw

    @staticmethod
    def _dgrad(u, draw, add, out=None):
        g = u.g
        if out is None:
            out = torch.empty((g.N, g.Di, g.Hi, g.Wi, g.Cin), dtype=torch.bfloat16, device=draw.device)
        L.check(L.lib().mmnn_rn_conv(C.byref(g), 1, 0, _p(draw), _p(u.conv.weight.detach()), _p(out), _p(add), None, _stream()),
                "mmnn_rn_conv (data gradient)")
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = L.lib()
        model, tape, training = ctx.model, ctx.tape, ctx.training
        pooled, out, ashape, V = ctx.head
        grads = {}
        F = _ResnetFn
        with torch.cuda.device(out.device):
            dout = dout.contiguous().float()
            N, K = out.shape
            Cl = ashape[-1]
            dy = torch.empty(ashape, dtype=torch.bfloat16, device=out.device)
            dW = torch.zeros_like(model.fc.weight, dtype=torch.float32)
            db = torch.zeros_like(model.fc.bias, dtype=torch.float32)
            L.check(lib.mmnn_rn_head_bwd(_p(dout), _p(out), _p(pooled), _p(model.fc.weight.detach()), N, V, Cl, K, _p(dy),
                                         _p(dW), _p(db), _stream()), "mmnn_rn_head_bwd")
            grads[model.fc.weight], grads[model.fc.bias] = dW, db
            for entry in reversed(tape):
                if entry[0] == "block":
                    _, u1, a1, u2, ud, y, scale = entry
                    draw2, drawd, dz = F._bn_backward(u2, dy, y, scale, ud, ud is None, grads, training)
                    F._wgrad(u2, draw2, grads)
                    da1 = F._dgrad(u2, draw2, None)
                    draw1, _, _ = F._bn_backward(u1, da1, a1, 1.0, None, False, grads, training)
                    F._wgrad(u1, draw1, grads)
                    if ud is not None:
                        F._wgrad(ud, drawd, grads)
                        g1 = u1.g
                        dx = torch.empty((g1.N, g1.Di, g1.Hi, g1.Wi, g1.Cin), dtype=torch.bfloat16, device=draw1.device)
                        rc = lib.mmnn_rn_conv_dgrad_ds(C.byref(g1), _p(draw1), _p(u1.conv.weight.detach()), _p(drawd),
                                                       _p(ud.conv.weight.detach()), _p(dx), _stream())
                        if rc == -9:                           # no fused kernel for this pair: two launches, second in place
                            dx = F._dgrad(u1, draw1, None, out=dx)
                            dx = F._dgrad(ud, drawd, dx, out=dx)
                        else:
                            L.check(rc, "mmnn_rn_conv_dgrad_ds")
                    else:
                        dx = F._dgrad(u1, draw1, dz)           # + identity-residual gradient
                    dy = dx
                else:
                    _, u0, a0 = entry
                    draw0, _, _ = F._bn_backward(u0, dy, a0, 1.0, None, False, grads, training)
                    F._wgrad(u0, draw0, grads)
        ctx.tape = None
        params = [q for _, q in model.named_parameters()]
        return (None, None) + tuple(grads.get(q) for q in params)


class Resnet18(nn.Module):
    """/root/reference/models/resnet.py:112-203.  `block` must be BasicBlock (the reference's r3d_18 uses nothing else)."""

    def __init__(self, block, conv_makers, layers, stem, num_classes=400, zero_init_residual=False, dropout_prob=0.2):
        super().__init__()
        if block is not BasicBlock:
            raise NotImplementedError("mmnn_sts_b200 Resnet18 runs BasicBlock stages only (what r3d_18 builds)")
        self.inplanes = 64
        self.stem = stem()
        self.dropout = torch.nn.Dropout(p=dropout_prob)
        self.layer1 = self._make_layer(block, conv_makers[0], 8, layers[0], stride=1)
        self.layer2 = self._make_layer(block, conv_makers[1], 16, layers[1], stride=2)
        self.layer3 = self._make_layer(block, conv_makers[2], 8, layers[2], stride=2)
        self.layer4 = self._make_layer(block, conv_makers[3], 16, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool3d((1, 1, 1))
        self.fc = nn.Linear(16 * block.expansion, num_classes)
        self._initialize_weights()
        self.injected_masks = None          # tests: list of 4 uint8 keep-masks (channels-last stage outputs)
        self._seed_gen = torch.Generator().manual_seed(torch.initial_seed() & 0x7FFFFFFF)

    def _next_seed(self):
        return int(torch.randint(0, 2 ** 62, (1,), generator=self._seed_gen).item())

    def _make_layer(self, block, conv_builder, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            ds_stride = conv_builder.get_downsample_stride(stride)
            downsample = nn.Sequential(
                nn.Conv3d(self.inplanes, planes * block.expansion, kernel_size=1, stride=ds_stride, bias=False),
                nn.BatchNorm3d(planes * block.expansion))
        layers = [block(self.inplanes, planes, conv_builder, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, conv_builder))
        return nn.Sequential(*layers)

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        params = [q for _, q in self.named_parameters()]
        return _ResnetFn.apply(self, x, *params)


def algorithmic_cost(model, batch, spatial):
    """Algorithmic bytes and FLOPs of one training step per kernel class: every logical tensor a kernel
    reads or writes counted once in its storage type (fp32 image, fp16 activations, bf16 gradients); conv FLOPs = 2 M N K.
    Returns {class: {"bytes": .., "flops": ..}} for bench.py's roofline block."""
    cost = {k: {"bytes": 0.0, "flops": 0.0} for k in ("rn_fprop", "rn_dgrad", "rn_wgrad", "rn_eltwise")}

    def conv(c, dims, f32=False, dgrad=True):
        g = _geom(c, batch, dims)
        vin, vout = batch * g.Di * g.Hi * g.Wi * g.Cin, batch * g.Do * g.Ho * g.Wo * g.Cout
        fl = 2.0 * vout * g.Cin * g.kd * g.kh * g.kw
        cost["rn_fprop"]["bytes"] += vin * (4 if f32 else 2) + vout * 2
        cost["rn_fprop"]["flops"] += fl
        cost["rn_wgrad"]["bytes"] += vin * (4 if f32 else 2) + vout * 2
        cost["rn_wgrad"]["flops"] += fl
        if dgrad:
            cost["rn_dgrad"]["bytes"] += vin * 2 + vout * 2
            cost["rn_dgrad"]["flops"] += fl
        return (g.Do, g.Ho, g.Wo), vout

    e = cost["rn_eltwise"]
    dims, v = conv(model.stem[0], tuple(spatial), f32=True, dgrad=False)
    e["bytes"] += v * 2 * (2 + 3 + 4)                       # bn_act: raw, y; reduce: dy, y, raw; apply: dy, y, raw, draw
    for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
        for blk in layer:
            d1, v1 = conv(blk.conv1[0], dims)
            e["bytes"] += v1 * 2 * (2 + 3 + 4)
            _, v2 = conv(blk.conv2[0], d1)
            if blk.downsample is not None:
                conv(blk.downsample[0], dims)
                e["bytes"] += v2 * 2 * (3 + 4 + 6)          # + raw2 in every pass, + draw2
            else:
                e["bytes"] += v2 * 2 * (3 + 3 + 5) + v2 * 2  # + identity residual in the forward pass, + dz out, + dz add in dgrad
            dims = d1
    return cost


def r3d_18(num_classes):
    """/root/reference/models/resnet.py:205-226."""
    return Resnet18(BasicBlock, [Conv3DSimple] * 4, [2, 2, 2, 2], BasicStem, num_classes=num_classes)
