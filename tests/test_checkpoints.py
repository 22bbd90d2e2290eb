"""checkpoint compatibility (loadWeights, BHB key remap) on CPU, TinyDensenet on the GPU."""
import os

import pytest
import torch


def _model(cls="DenseNet121", cin=1):
    from mmnn_sts_b200.models import densenet as D
    return getattr(D, cls)(spatial_dims=3, in_channels=cin, out_channels=2, feature_channels=12, dropout_prob=0.0)


def test_load_weights_roundtrip(tmp_path):
    """best_surv_model.pth written by one model (main.py:577 `torch.save(model.state_dict())`) loads into another."""
    from mmnn_sts_b200.models.multimodal import MultiModalModel
    from mmnn_sts_b200.utils.utils import loadWeights
    a = MultiModalModel(_model(cin=2), ["x"] * 20, 2, 12, blend=True)
    b = MultiModalModel(_model(cin=2), ["x"] * 20, 2, 12, blend=True)
    path = os.path.join(tmp_path, "best_surv_model.pth")
    torch.save(a.state_dict(), path)
    loadWeights(b, path, "cpu")
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k
    with pytest.raises(Exception):
        loadWeights(_model(cin=1), path, "cpu")            # wrong architecture still fails loudly, as in the reference


def test_bhb_checkpoint_remap(tmp_path):
    """A checkpoint laid out like DenseNet121_BHB-10K_yAwareContrastive.pth ({'model': {'module.features.denseblockB.
    denselayerL.<leaf>': ...}}) must land on the backbone through the reference's key remap (utils/utils.py:373-382)."""
    from mmnn_sts_b200.utils.utils import loadWeights, remap_bhb_checkpoint
    src, dst = _model(), _model()
    bhb = {}
    for k, v in src.backbone.state_dict().items():          # our backbone keys -> the BHB file's layout
        parts = k.split(".")
        if parts[0].startswith("denseblock"):
            assert parts[2] == "layers"
            parts.pop(2)
        bhb["module.features." + ".".join(parts)] = v.clone()
    ck = {"model": bhb, "epoch": 49}
    remapped = remap_bhb_checkpoint(ck)
    assert "features.denseblock1.denselayer1.layers.conv1.weight" in remapped
    assert "features.conv0.weight" in remapped and not any(k.startswith("module.") for k in remapped)

    class Wrapper(torch.nn.Module):                          # the reference loads it into a module whose trunk is called `features`
        def __init__(self, bb):
            super().__init__()
            self.features = bb
            self.classifier = torch.nn.Linear(4, 2)            # absent from the checkpoint: strict=False must tolerate it
    w = Wrapper(dst.backbone)
    path = os.path.join(tmp_path, "DenseNet121_BHB-10K_yAwareContrastive.pth")
    torch.save(ck, path)
    loadWeights(w, path, "cpu")
    for (k, va), (_, vb) in zip(src.backbone.state_dict().items(), dst.backbone.state_dict().items()):
        assert torch.equal(va, vb), k


def test_tiny_densenet_state_dict_layout():
    from oracle import synth
    m = _model("TinyDensenet", cin=1)
    spec = synth.state_dict_spec(in_channels=1, block_config=(6, 12, 4))
    keys = [k for k, _, _ in spec if k.startswith("image_model.model.")]
    sd = m.state_dict()
    assert [("image_model.model." + k) for k in sd.keys()] == keys
    assert m.backbone.out_channels == 384


@pytest.mark.gpu
def test_tiny_densenet_matches_oracle():
    """TinyDensenet (block_config (6,12,4), /root/reference/models/densenet.py:333-356): forward and backward of the
    trunk + features head against the fp32 oracle in train mode (batch statistics), 16-bit tolerances of DESIGN.md 5."""
    from oracle import model as om, synth
    torch.manual_seed(0)
    cfg = (6, 12, 4)
    sd_full = synth.make_state_dict(7, in_channels=1, block_config=cfg)
    sd = {k[len("image_model.model."):]: v for k, v in sd_full.items() if k.startswith("image_model.model.")}
    m = _model("TinyDensenet", cin=1)
    m.load_state_dict(sd)
    m = m.cuda().train()
    x = torch.rand(4, 1, 64, 64, 32)
    feats = m.features(m.backbone(x.cuda()))
    feats.square().sum().backward()
    p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
    ref = om.densenet_features(p, om.densenet_backbone(p, x, True, None, "", block_config=cfg), None, "")
    ref.square().sum().backward()
    err = float((feats.detach().cpu() - ref.detach()).abs().max() / ref.detach().abs().max())
    assert err < 2e-2, err
    for name in ("backbone.denseblock3.denselayer4.layers.conv2.weight", "backbone.transition2.conv.weight",
                 "backbone.denseblock1.denselayer1.layers.conv1.weight", "features.feature_layer.weight"):
        g = dict(m.named_parameters())[name].grad.detach().cpu().flatten()
        r = p[name].grad.flatten()
        cos = float(torch.dot(g, r) / (g.norm() * r.norm()))
        # same criteria as tests/test_trunk_gpu.py: ReLU-mask flips bound gradient agreement
        assert cos > 0.97 and 0.9 < float(g.norm() / r.norm()) < 1.1, (name, cos, float(g.norm() / r.norm()))
