"""Hot-path helpers with the reference's names and argument meaning (/root/reference/utils/utils.py:20-29,238-251)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def criterion(loss_func, preds, labels, device):
    return loss_func(preds, labels).to(device)


def surv_criterion(loss_func, preds, events, durations, device):
    """SUM over classes of loss_func(preds[:, i], events[:, i], durations[:, i]) (/root/reference/utils/utils.py:24-29).
    When loss_func is this package's CoxPH all classes go to the GPU in ONE launch (one segment per class)."""
    from ..losses.losses import CoxPH, _coxph_columns
    if loss_func is CoxPH and preds.is_cuda:
        return _coxph_columns(preds, events, durations).sum().to(device)
    losses = 0
    for i in range(preds.shape[1]):
        losses += loss_func(preds[:, i], events[:, i], durations[:, i]).to(device)
    return losses


class BackpropagatableFeatureExtractor(nn.Module):
    """features(backbone(x)) of a model split into `backbone` and `features` (/root/reference/utils/utils.py:238-251)."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        x = self.model.backbone(x)
        return self.model.features(x)


def remap_bhb_checkpoint(checkpoint):
    """State-dict keys of the BHB-10K yAware-contrastive DenseNet121 checkpoint renamed to this package's (= the
    reference's MONAI-style) layout.  Same mapping as /root/reference/utils/utils.py:373-382: the DataParallel prefix
    `module.` goes away, and the dense layers of that checkpoint (`features.denseblockB.denselayerL.<leaf>`) gain the
    `layers` container level (`features.denseblockB.denselayerL.layers.<leaf>`).  `checkpoint` is the loaded file, a
    dict whose 'model' entry is the state_dict."""
    def rename(key):
        parts = key.replace('module.', '').split('.')
        in_dense_block = len(parts) > 1 and parts[0] == 'features' and parts[1].startswith('dense')
        if in_dense_block:
            parts = parts[:3] + ['layers'] + parts[3:]
        return '.'.join(parts)

    return {rename(key): tensor for key, tensor in checkpoint['model'].items()}


def loadWeights(model, path, device):
    """`loadWeights` of the reference (/root/reference/utils/utils.py:357-390) without the S3 branch (storage is out of
    scope): a plain state_dict loads strictly; the BHB pretrained backbone is remapped and loaded
    with strict=False (it has no classification head).  Works because the parameter containers of
    mmnn_sts_b200.models keep the reference's state_dict keys and shapes."""
    try:
        checkpoint = torch.load(path, map_location=device)
        model.load_state_dict(checkpoint)
    except Exception as e1:
        if str(path).endswith('DenseNet121_BHB-10K_yAwareContrastive.pth'):
            checkpoint = torch.load(path, map_location=device)
            model.load_state_dict(remap_bhb_checkpoint(checkpoint), strict=False)
        else:
            raise e1
    return model


class MultiModalGradCAM(nn.Module):
    """GradCAM of the reference (/root/reference/utils/utils.py:253-344) on the fused trunk.

    The reference hooks the LAST nn.Conv3d of `model.gradcam_layer` (= denseblock4.denselayer16.layers.conv2) for its
    output and the gradient of that output.  Here the trunk is one fused call, so both are read from its workspace: the
    conv's output is the last 32 channels of the final dense block's buffer, its gradient the same channels of that
    block's fp32 gradient accumulator after a backward pass (mmnn_encoder_debug_offsets).  Everything else follows the
    reference line by line -- including that the activations are scaled IN PLACE class after class (:306-307), so the
    map of class c carries the pooled gradients of classes 0..c -- and that only batch size 1 is accepted (:330).
    The model must be in eval mode (the reference calls it from inference_survival) on a CUDA device."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.input_shape = None

    def forward(self, x):
        outputs = self.model(x)
        self.input_shape = x['image'].shape
        att_maps = self.attentionMaps(outputs)
        return outputs, att_maps

    def _last_conv_views(self):
        bb = self.model.gradcam_layer
        ws, xshape = getattr(bb, "_last_ws", None), getattr(bb, "_last_xshape", None)
        if ws is None:
            raise RuntimeError("GradCAM needs a forward pass with autograd enabled (do not wrap it in torch.no_grad())")
        nb = len(bb._cfg[1])
        act, grad = bb.block_buffer_views(ws, xshape, nb - 1)
        d, h, w = bb._last_out_dims
        B, growth = xshape[0], bb._cfg[3]

        def to_ncdhw(t):
            return t[:, t.shape[1] - growth:].float().reshape(B, d, h, w, growth).permute(0, 4, 1, 2, 3).contiguous()
        return act, grad, to_ncdhw

    def attentionMaps(self, outputs):
        """One trilinearly up-sampled, [0, 1]-normalised map per output class (/root/reference/utils/utils.py:293-344).
        Reference behaviours kept on purpose: the activation tensor is re-weighted cumulatively (class c sees the pooled
        gradients of classes 0..c multiplied together) and a batch of more than one volume is rejected."""
        act, grad, to_ncdhw = self._last_conv_views()
        weighted = to_ncdhw(act)                                     # [1, growth, d, h, w], re-weighted in place below
        volume_shape = tuple(self.input_shape[2:])
        maps = []
        for class_index in range(outputs.shape[1]):
            outputs[0, class_index].backward(retain_graph=True)      # fills the block's gradient accumulator
            channel_weight = to_ncdhw(grad).mean(dim=(0, 2, 3, 4))    # GradCAM alpha: gradient pooled over voxels
            weighted.mul_(channel_weight.view(1, -1, 1, 1, 1))
            cam = weighted.mean(dim=1).squeeze()
            cam = cam - cam.min()
            cam = cam / cam.max()
            if cam.ndim != 3:
                raise AssertionError('attention maps are computed one volume at a time: call with batch size 1')
            maps.append(F.interpolate(cam[None, None], volume_shape, mode='trilinear').squeeze())
        return maps
