"""fp32 CPU restatement of the reference network forward pass, written functionally over a state_dict.

TEST INFRASTRUCTURE.  Follows (does not copy) the module structure of the reference:
  * 3-D DenseNet-121 trunk      /root/reference/models/densenet.py:46-148 (_DenseLayer/_DenseBlock/_Transition),
                                 :196-231 (backbone), :234-247 (features head)
  * clinical MLP                 /root/reference/models/mlp.py:19-51
  * fusion + blending heads      /root/reference/models/multimodal.py:51-80
  * BackpropagatableFeatureExtractor = features(backbone(x))   /root/reference/utils/utils.py:238-251
State-dict key names are the reference's (779 entries) so one set of weights drives the reference (under
oracle/shim.py), this restatement and the CUDA build.  Validated against the unchanged reference files by
tests/golden/make_golden.py (max-abs difference recorded in the fixture).

Dropout: masks are INJECTED (`masks` dict) so all three implementations can share them:
  masks['dense'][(block, layer)] -> [B, growth] keep-mask already divided by (1-p)   (Dropout3d = channel dropout)
  masks['image_features']        -> [B, F]      elementwise keep-mask / (1-p)        (nn.Dropout)
  masks['mlp'][i]                -> [B]         per-SAMPLE keep-mask / (1-p)  (quirk Q3: Dropout1d on a 2-D
                                                 [B,F] input is treated as unbatched (C,L): drops whole rows)
"""
import torch
import torch.nn.functional as F

BLOCK_CONFIG = (6, 12, 24, 16)


def _bn(x, sd, prefix, training, momentum=0.1, eps=1e-5):
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and (prefix + ".num_batches_tracked") in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)


def densenet_backbone(sd, x, training, masks=None, prefix="", block_config=BLOCK_CONFIG, collect=None):
    p = prefix + "backbone."
    x = F.conv3d(x, sd[p + "conv0.weight"], None, stride=2, padding=3)
    if collect is not None: collect["conv0"] = x
    x = F.relu(_bn(x, sd, p + "norm0", training))
    x = F.max_pool3d(x, kernel_size=3, stride=2, padding=1)
    if collect is not None: collect["pool0"] = x
    for b, nl in enumerate(block_config):
        for l in range(nl):
            q = f"{p}denseblock{b + 1}.denselayer{l + 1}.layers."
            y = F.relu(_bn(x, sd, q + "norm1", training))
            y = F.conv3d(y, sd[q + "conv1.weight"])
            y = F.relu(_bn(y, sd, q + "norm2", training))
            y = F.conv3d(y, sd[q + "conv2.weight"], padding=1)
            if masks is not None and masks.get("dense") is not None:
                y = y * masks["dense"][(b, l)][:, :, None, None, None]
            x = torch.cat([x, y], 1)
        if collect is not None: collect[f"block{b + 1}"] = x
        if b == len(block_config) - 1:
            x = _bn(x, sd, p + "norm5", training)
        else:
            q = f"{p}transition{b + 1}."
            y = F.relu(_bn(x, sd, q + "norm", training))
            y = F.conv3d(y, sd[q + "conv.weight"])
            x = F.avg_pool3d(y, kernel_size=2, stride=2)
    return x


def densenet_features(sd, x, masks=None, prefix=""):
    x = F.relu(x)
    x = F.adaptive_avg_pool3d(x, 1).flatten(1)
    x = F.linear(x, sd[prefix + "features.feature_layer.weight"], sd[prefix + "features.feature_layer.bias"])
    if masks is not None and masks.get("image_features") is not None:
        x = x * masks["image_features"]
    return x


def mlp_features(sd, x, training, masks=None, prefix=""):
    """dense0->bn0->ReLU->drop ; then 4x dense->bn->drop->ReLU ; features: dense5->bn5->drop->ReLU."""
    names = [("backbone.dense0", "backbone.bn0"), ("backbone.dense1", "backbone.bn1"),
             ("backbone.dense2", "backbone.bn2"), ("backbone.dense3", "backbone.bn3"),
             ("backbone.dense4", "backbone.bn4"), ("features.dense5", "features.bn5")]
    for i, (d, b) in enumerate(names):
        x = F.linear(x, sd[prefix + d + ".weight"], sd[prefix + d + ".bias"])
        x = _bn(x, sd, prefix + b, training)
        m = None if masks is None or masks.get("mlp") is None else masks["mlp"][i][:, None]
        if i == 0:
            x = F.relu(x)
            if m is not None: x = x * m
        else:
            if m is not None: x = x * m
            x = F.relu(x)
    return x


def multimodal_forward(sd, image, clinical, training, blend, masks=None, collect=None):
    img_f = densenet_features(sd, densenet_backbone(sd, image, training, masks, "image_model.model.", collect=collect),
                              masks, "image_model.model.")
    clin_f = mlp_features(sd, clinical, training, masks, "clinical_model.model.")
    if collect is not None:
        collect["image_features"] = img_f; collect["clinical_features"] = clin_f
    feats = torch.cat([img_f, clin_f], 1)
    out = F.linear(feats, sd["output_head.weight"], sd["output_head.bias"])
    if blend:
        ip = F.linear(img_f, sd["image_output_head.weight"], sd["image_output_head.bias"])
        cp = F.linear(clin_f, sd["clinical_output_head.weight"], sd["clinical_output_head.bias"])
        out = torch.stack((out, ip, cp), dim=0)
    return out
