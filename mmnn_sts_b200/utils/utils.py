"""Hot-path helpers with the reference's names and argument meaning (/root/reference/utils/utils.py:20-29,238-251)."""
import torch
import torch.nn as nn


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
    """Key remapping the reference applies to the BHB-10K yAware-contrastive DenseNet121 checkpoint so that it loads into
    the MONAI-style `backbone` (/root/reference/utils/utils.py:373-382): strip `module.`, and insert `layers` after the
    dense-layer name of every `features.dense*` key.  `checkpoint` is the loaded file (a dict with a 'model' entry)."""
    new_checkpoint = {}
    for key in checkpoint['model'].keys():
        new_key = key.replace('module.', '')
        heirarchy = new_key.split('.')
        if heirarchy[0] == 'features' and heirarchy[1].startswith('dense'):
            heirarchy.insert(3, 'layers')
        new_checkpoint['.'.join(heirarchy)] = checkpoint['model'][key]
    return new_checkpoint


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
