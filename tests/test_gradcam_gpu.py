"""MultiModalGradCAM on the fused trunk against golden attention maps produced by the
UNCHANGED reference class (tests/golden/make_gradcam_golden.py).  Also covers eval-mode backward of the trunk (running-
statistics BatchNorm) that GradCAM relies on.  Tolerance: maps are min-max normalised to [0, 1]; with 16-bit activation /
gradient storage they must agree to 0.05 absolute and have the same arg-max voxel."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "gradcam.npz")


def _model(sd, blend=False):
    from mmnn_sts_b200.models.densenet import DenseNet121
    from mmnn_sts_b200.models.multimodal import MultiModalModel
    m = MultiModalModel(DenseNet121(spatial_dims=3, in_channels=1, out_channels=2, feature_channels=12, dropout_prob=0.0),
                        ["x"] * 20, 2, 12, blend=blend)
    m.load_state_dict(sd)
    return m.cuda().eval()


def test_gradcam_matches_reference_golden():
    from oracle import synth
    g = np.load(GOLD)
    sd = synth.make_state_dict(int(g["state_seed"]), in_channels=1)
    image, clinical, _, _ = synth.make_batch(int(g["batch_seed"]), 1, 1, (64, 64, 64))
    m = _model(sd)
    cam = m.add_gradcam(".")
    outputs, maps = cam({"image": image.cuda(), "clinical": clinical.cuda()})
    ref_logits = torch.from_numpy(g["logits"])
    assert float((outputs.detach().cpu() - ref_logits).abs().max()) < 2e-2 * float(ref_logits.abs().max()) + 5e-3
    maps = torch.stack(maps).cpu()
    assert tuple(maps.shape) == (2, 64, 64, 64) and float(maps.min()) >= 0.0 and float(maps.max()) <= 1.0 + 1e-6
    ref = torch.from_numpy(g["maps_sub4"])
    sub = maps[:, ::4, ::4, ::4]
    err = float((sub - ref).abs().max())
    print("gradcam max abs err", err, "mean", float((sub - ref).abs().mean()))
    assert err < 0.05, err
    for c in range(2):
        assert int(sub[c].argmax()) == int(ref[c].argmax())
    np.testing.assert_allclose(maps.mean(dim=(1, 2, 3)).numpy(), g["maps_mean"], atol=0.03)


def test_eval_mode_backward_matches_oracle():
    """Gradients through the trunk in eval mode (BatchNorm = fixed affine map of its running statistics)."""
    from oracle import model as om, synth
    sd = synth.make_state_dict(31, in_channels=1)
    image, clinical, _, _ = synth.make_batch(32, 2, 1, (64, 64, 32))
    m = _model(sd)
    out = m({"image": image.cuda(), "clinical": clinical.cuda()})
    out.square().sum().backward()
    p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
    ref = om.multimodal_forward(p, image, clinical, False, False, None)
    ref.square().sum().backward()
    assert float((out.detach().cpu() - ref.detach()).abs().max()) < 2e-2 * float(ref.detach().abs().max()) + 5e-3
    named = dict(m.named_parameters())
    for name in ("image_model.model.backbone.denseblock4.denselayer16.layers.conv2.weight",
                 "image_model.model.backbone.denseblock2.denselayer3.layers.norm1.weight",
                 "image_model.model.backbone.transition1.conv.weight",
                 "image_model.model.backbone.denseblock1.denselayer2.layers.conv1.weight"):
        gq, r = named[name].grad.detach().cpu().flatten(), p[name].grad.flatten()
        cos = float(torch.dot(gq, r) / (gq.norm() * r.norm()))
        assert cos > 0.97 and 0.9 < float(gq.norm() / r.norm()) < 1.1, (name, cos, float(gq.norm() / r.norm()))
