"""Clinical MLP with the reference's constructor, sub-module names and state_dict layout
(/root/reference/models/mlp.py:7-63); the arithmetic runs in the fused mmnn_mlp_heads kernel."""
from collections import OrderedDict

import torch
import torch.nn as nn


class MLP(nn.Module):
    def __init__(self, in_channels=1, out_channels=3, feature_channels=12, dropout_prob=0.2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.feature_channels, self.dropout_prob = feature_channels, dropout_prob
        self.relu = nn.ReLU()
        widths = [in_channels, 32, 16, 8, 8, 8]
        layers = []
        for i in range(5):
            # the reference passes `3` as nn.Linear's bias argument for dense1..4 (mlp.py:25,29,33,37) -> bias=True
            layers.append((f"dense{i}", nn.Linear(widths[i], widths[i + 1])))
            layers.append((f"bn{i}", nn.BatchNorm1d(widths[i + 1])))
            if i == 0:
                layers += [("relu0", self.relu), ("drop0", nn.Dropout1d(dropout_prob))]
            else:
                layers += [(f"drop{i}", nn.Dropout1d(dropout_prob)), (f"relu{i}", self.relu)]
        self.backbone = nn.Sequential(OrderedDict(layers))
        self.features = nn.Sequential(OrderedDict([
            ("dense5", nn.Linear(8, feature_channels)), ("bn5", nn.BatchNorm1d(feature_channels)),
            ("drop5", nn.Dropout1d(dropout_prob)), ("relu5", self.relu)]))
        self.output_head = nn.Sequential(OrderedDict([("dense6", nn.Linear(feature_channels, out_channels))]))
        self.injected_masks = None  # tests: [6, B] per-sample keep-mask / (1-p)

    # parameter / buffer tables in the order the fused kernel expects
    def kernel_params(self):
        mods = [(getattr(self.backbone, f"dense{i}"), getattr(self.backbone, f"bn{i}")) for i in range(5)]
        mods.append((self.features.dense5, self.features.bn5))
        params, buffers = [], []
        for d, b in mods:
            params += [d.weight, d.bias, b.weight, b.bias]
            buffers.append((b.running_mean, b.running_var, b.num_batches_tracked))
        return params, buffers

    def sample_masks(self, batch, device):
        if not self.training:
            return None
        if self.injected_masks is not None:
            return self.injected_masks.to(device=device, dtype=torch.float32).contiguous()
        if self.dropout_prob <= 0:
            return None
        keep = 1.0 - self.dropout_prob
        return torch.bernoulli(torch.full((6, batch), keep, device=device)) / keep

    def forward(self, x):
        """Clinical-only model of the reference (/root/reference/models/mlp.py:57-63): output_head(features(backbone(x))).
        Runs the same fused kernel as the multimodal path: `output_head.dense6` takes the clinical head's slot, the image
        features are zeros and the two other heads are unused zero weights (their outputs are discarded)."""
        from ..ops import MLPHeads
        params, buffers = self.kernel_params()
        C_, F_ = self.out_channels, self.feature_channels

        def zeros(*shape):
            return torch.zeros(*shape, dtype=torch.float32, device=x.device)
        heads = [zeros(C_, 2 * F_), zeros(C_), zeros(C_, F_), zeros(C_), self.output_head.dense6.weight, self.output_head.dense6.bias]
        mask = self.sample_masks(x.shape[0], x.device)
        preds = MLPHeads.apply(x, zeros(x.shape[0], F_), mask, (self.training, True, C_), buffers, *params, *heads)
        return preds[2]
