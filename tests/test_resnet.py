"""CPU tests of the 3-D ResNet path: the oracle restatement (oracle/resnet.py) against the golden
outputs of the UNCHANGED reference module (tests/golden/make_resnet_golden.py), and the drop-in module's layout."""
import os

import numpy as np
import pytest
import torch

from oracle import resnet as orn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NUM_CLASSES, SPATIAL, BATCH = 5, (12, 32, 24), 3


def _inputs():
    sd = orn.make_state_dict(7, NUM_CLASSES)
    image, labels = orn.make_batch(11, BATCH, SPATIAL, NUM_CLASSES)
    return sd, image, labels


def test_oracle_eval_matches_reference_golden():
    sd, image, _ = _inputs()
    gold = np.load(os.path.join(GOLD, "resnet_eval.npz"))
    with torch.no_grad():
        out = orn.resnet_forward(sd, image, training=False)
    np.testing.assert_allclose(out.numpy(), gold["out"], rtol=0, atol=2e-6)


def test_oracle_train_matches_reference_golden():
    sd, image, labels = _inputs()
    gold = np.load(os.path.join(GOLD, "resnet_train.npz"))
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    out = orn.resnet_forward(p, image, training=True, masks=None)
    loss = orn.train_step_loss(out, labels, torch.from_numpy(gold["pos_weight"]))
    loss.backward()
    np.testing.assert_allclose(out.detach().numpy(), gold["out"], rtol=0, atol=2e-6)
    assert abs(loss.item() - float(gold["loss"])) < 1e-5 * abs(float(gold["loss"]))
    for k in gold.files:
        if k.startswith("grad:"):
            g, ref = p[k[5:]].grad.numpy(), gold[k]
            assert np.abs(g - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-7, k
        if k.startswith("stat:"):
            np.testing.assert_allclose(p[k[5:]].numpy(), gold[k], rtol=1e-5, atol=1e-6, err_msg=k)


def test_module_layout_matches_reference_spec():
    from mmnn_sts_b200.models.resnet import r3d_18
    torch.manual_seed(0)
    m = r3d_18(NUM_CLASSES)
    spec = orn.state_dict_spec(NUM_CLASSES)
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sum(q.numel() for q in m.parameters()) == 80757          # 80 706 at the reference's 2 classes
    # the reference's initialisation law (models/resnet.py:189-203)
    assert float(m.fc.bias.abs().max()) == 0.0 and float(m.fc.weight.std()) < 0.02
    assert float(m.stem[1].weight.min()) == 1.0
    m.load_state_dict(orn.make_state_dict(7, NUM_CLASSES))
    shapes = orn.stage_shapes(BATCH, SPATIAL)
    assert shapes[0] == (3, 8, 14, 16, 12) and shapes[3] == (3, 16, 2, 2, 2)


def test_no_cpu_path():
    from mmnn_sts_b200 import _lib
    from mmnn_sts_b200.models.resnet import r3d_18
    m = r3d_18(2)
    with pytest.raises(_lib.MMNNLibraryError):
        m(torch.zeros(2, 1, 4, 16, 16))
