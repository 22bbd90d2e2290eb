"""Input pipeline: Normalize -> ScaleIntensity -> Resize.
CPU: the oracle restatement against the golden vectors produced by the UNCHANGED reference Normalize class.
GPU: the fused kernel (mmnn_preprocess_volumes through mmnn_sts_b200.data.transforms.ValTransformsGPU) against the
oracle and the golden vectors; fp32, tolerance 2e-6 absolute (window sums are accumulated in a different order)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")
MEAN, STD = 286.90859071507913, 581.7816096485366
CASES = ["mri_like", "upsample", "odd", "negative"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    from oracle import preprocess as op
    g = np.load(GOLD)
    raw = g[case + "_raw"]
    assert np.array_equal(op.normalize(raw, MEAN, STD).astype(np.float32), g[case + "_normalized"])
    out = op.val_transforms(raw, MEAN, STD, tuple(g[case + "_size"]))
    assert np.array_equal(out, g[case + "_out"])
    assert out.min() >= 0.0 and out.max() <= 1.0


def test_oracle_edge_cases():
    from oracle import preprocess as op
    const = np.full((1, 4, 4, 4), 7.0, np.float32)
    assert np.array_equal(op.val_transforms(const, MEAN, STD, (2, 2, 2)), np.zeros((1, 2, 2, 2), np.float32))   # all-equal image -> zeros
    x = np.arange(2 * 4 * 4 * 4, dtype=np.float32).reshape(2, 4, 4, 4) + 1
    same = op.val_transforms(x, MEAN, STD, (4, 4, 4))                  # identity resize: pure min-max scaling over BOTH channels
    np.testing.assert_allclose(same, (x - 1) / (x.max() - 1), rtol=0, atol=5e-5)   # fp32 cancellation in x - mean*max(x), as in the reference


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_matches_golden(case):
    from mmnn_sts_b200.data.transforms import ValTransformsGPU
    g = np.load(GOLD)
    raw = torch.from_numpy(g[case + "_raw"]).cuda()
    tf = ValTransformsGPU(MEAN, STD, tuple(int(v) for v in g[case + "_size"]))
    out = tf(raw).cpu().numpy()
    np.testing.assert_allclose(out, g[case + "_out"], rtol=0, atol=2e-6)
    batch = torch.stack([raw, raw * 0.5 + 3.0])                          # per-patient statistics inside a batch
    outb = tf(batch).cpu().numpy()
    np.testing.assert_allclose(outb[0], g[case + "_out"], rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_gpu_matches_oracle_bench_shape():
    """One raw 2-channel 160x192x96 volume pair -> 128x128x64 (the bench workload's input shape), non-integer windows."""
    from mmnn_sts_b200.data.transforms import ValTransformsGPU
    from oracle import preprocess as op
    rng = np.random.RandomState(3)
    raw = (rng.gamma(2.0, 200.0, size=(2, 2, 160, 192, 96)) * (rng.rand(2, 2, 160, 192, 96) > 0.1)).astype(np.float32)
    out = ValTransformsGPU(MEAN, STD, (128, 128, 64))(torch.from_numpy(raw).cuda()).cpu().numpy()
    for b in range(2):
        ref = op.val_transforms(raw[b], MEAN, STD, (128, 128, 64))
        np.testing.assert_allclose(out[b], ref, rtol=0, atol=2e-6)
    const = torch.full((1, 1, 8, 8, 8), 5.0, device="cuda")
    assert float(ValTransformsGPU(MEAN, STD, (4, 4, 4))(const).abs().max()) == 0.0


def test_cpu_input_is_refused():
    from mmnn_sts_b200 import _lib as L
    from mmnn_sts_b200.data.transforms import ValTransformsGPU
    with pytest.raises(L.MMNNLibraryError):
        ValTransformsGPU()(torch.rand(1, 4, 4, 4))
