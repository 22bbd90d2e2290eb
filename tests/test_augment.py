"""Training-time augmentations of the reference's `train_transforms` (/root/reference/main.py:64-85) on the GPU
(mmnn_sts_b200/data/transforms.py TrainTransformsGPU, csrc/augment.cu).  The reference never executes this branch as shipped and
MONAI is not vendored, so there is no golden output: each transform is checked against a plain torch restatement of its published
definition ON THE SAME PARAMETERS (parity unpinned, stated in DESIGN.md section 9)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mmnn_sts_b200.data.transforms import IMAGE_DATA_MEAN, IMAGE_DATA_STDDEV, TrainTransformsGPU, ValTransformsGPU

NEUTRAL = {"rotate": None, "flip_axis": None, "zoom": None, "shift": 0.0, "gamma": 0.0, "smooth": None, "sharpen": None,
           "hist": None, "noise_std": 0.0, "seed": 1}


def _q(**kw):
    q = dict(NEUTRAL)
    q.update(kw)
    return q


def test_draw_follows_the_reference_probabilities_and_ranges():
    t = TrainTransformsGPU(seed=3)
    qs = t.draw(4000)
    frac = lambda f: sum(1 for q in qs if f(q)) / len(qs)
    assert abs(frac(lambda q: q["rotate"] is not None) - 0.5) < 0.03
    assert abs(frac(lambda q: q["flip_axis"] is not None) - 0.5) < 0.03
    assert abs(frac(lambda q: q["zoom"] is not None) - 0.5) < 0.03
    assert abs(frac(lambda q: q["shift"] != 0.0) - 0.3) < 0.03
    assert abs(frac(lambda q: q["gamma"] > 0.0) - 0.3) < 0.03
    assert abs(frac(lambda q: q["smooth"] is not None) - 0.2) < 0.03
    assert abs(frac(lambda q: q["sharpen"] is not None) - 0.2) < 0.03
    assert abs(frac(lambda q: q["hist"] is not None) - 0.3) < 0.03
    assert abs(frac(lambda q: q["noise_std"] > 0.0) - 0.3) < 0.03
    for q in qs:
        assert q["rotate"] is None or -15.0 <= q["rotate"] <= 15.0
        assert q["zoom"] is None or 0.9 <= q["zoom"] <= 1.1
        assert -0.1 <= q["shift"] <= 0.1 and (q["gamma"] == 0.0 or 0.5 <= q["gamma"] <= 4.5) and 0.0 <= q["noise_std"] <= 0.05
        if q["sharpen"] is not None:
            s1, s2, a = q["sharpen"]
            assert all(0.5 <= v <= 1.0 for v in s1) and all(0.5 <= w <= v for v, w in zip(s1, s2)) and 10.0 <= a <= 30.0
        if q["hist"] is not None:
            ref, flt = q["hist"]
            assert ref[0] == flt[0] == 0.0 and ref[-1] == flt[-1] == 1.0 and all(b >= a for a, b in zip(flt, flt[1:]))
    # same seed, same parameters
    a, b = TrainTransformsGPU(seed=11).draw(6), TrainTransformsGPU(seed=11).draw(6)
    assert a == b


def test_affine_composition():
    A = TrainTransformsGPU.affine(_q(flip_axis=1, zoom=1.1))
    assert np.allclose(A, np.diag([1 / 1.1, -1 / 1.1, 1 / 1.1]))
    A = TrainTransformsGPU.affine(_q(rotate=math.pi / 2))
    assert np.allclose(A @ np.array([0.0, 1.0, 0.0]), [0.0, 0.0, 1.0], atol=1e-12)
    assert np.allclose(TrainTransformsGPU.affine(NEUTRAL), np.eye(3))


def _raw(B, C, shape, seed, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, C, *shape, generator=g) * 1500.0 + 20.0).to(device)


def _normalized(raw):      # Normalize -> ScaleIntensity over the whole multi-channel image of one patient
    out = []
    for x in raw:
        M = x.max()
        n = (x - IMAGE_DATA_MEAN * M) / (IMAGE_DATA_STDDEV * M)
        out.append((n - n.min()) / (n.max() - n.min()))
    return torch.stack(out)


def _spatial_ref(raw, params, out_size):
    n = _normalized(raw)
    B, C, X, Y, Z = n.shape
    res = []
    dev = raw.device
    ctr = torch.tensor([(X - 1) / 2, (Y - 1) / 2, (Z - 1) / 2], dtype=torch.float64, device=dev)
    gx, gy, gz = torch.meshgrid(torch.arange(X, device=dev), torch.arange(Y, device=dev), torch.arange(Z, device=dev), indexing="ij")
    g = torch.stack([gx, gy, gz], -1).double() - ctr
    for b, q in enumerate(params):
        A = torch.tensor(TrainTransformsGPU.affine(q), dtype=torch.float64, device=dev)
        src = g @ A.T + ctr                                   # voxel coordinates in the source volume
        lim = torch.tensor([X - 1, Y - 1, Z - 1], dtype=torch.float64, device=dev)
        src = torch.minimum(torch.maximum(src, torch.zeros_like(lim)), lim)
        # grid_sample: last dim (x, y, z) = (W, H, D) = (Z, Y, X) order, align_corners=True maps -1..1 to 0..N-1
        norm = torch.stack([src[..., 2] / max(Z - 1, 1), src[..., 1] / max(Y - 1, 1), src[..., 0] / max(X - 1, 1)], -1) * 2 - 1
        s = F.grid_sample(n[b:b + 1].double(), norm[None], mode="bilinear", padding_mode="border", align_corners=True)
        res.append(F.adaptive_avg_pool3d(s, out_size)[0])
    return torch.stack(res).float()


@pytest.mark.gpu
def test_no_transform_fired_equals_the_validation_chain():
    raw = _raw(2, 2, (24, 20, 16), 1)
    t = TrainTransformsGPU(spatial_size=(8, 10, 8))
    got = t.apply(raw, [dict(NEUTRAL), dict(NEUTRAL)])
    ref = ValTransformsGPU(spatial_size=(8, 10, 8))(raw)
    # not bit-equal: the validation kernel normalises every voxel and then averages (the reference's order, bit-exact to the oracle);
    # the augmentation kernel averages the interpolated RAW values and normalises once -- fp32 cancellation in (x - mean * max)
    # puts the two ~2e-5 apart on a [0, 1] output
    assert torch.allclose(got, ref, atol=1e-4), (got - ref).abs().max()


@pytest.mark.gpu
@pytest.mark.parametrize("q", [
    _q(flip_axis=0), _q(flip_axis=2), _q(zoom=0.9), _q(zoom=1.1), _q(rotate=0.3), _q(rotate=-14.2, flip_axis=1, zoom=1.05),
])
def test_spatial_transforms_match_the_torch_restatement(q):
    raw = _raw(2, 2, (32, 32, 16), 2)
    t = TrainTransformsGPU(spatial_size=(16, 16, 8))
    got = t.apply(raw, [q, dict(NEUTRAL)])
    ref = _spatial_ref(raw, [q, dict(NEUTRAL)], (16, 16, 8))
    assert (got - ref).abs().max() < 1e-4, (got - ref).abs().max()
    if q["flip_axis"] is not None and q["rotate"] is None and q["zoom"] is None:      # a pure flip is exact
        plain = ValTransformsGPU(spatial_size=(16, 16, 8))(raw)
        assert torch.allclose(got[0], plain[0].flip(1 + q["flip_axis"]), atol=1e-4)


def _gauss1d(sigma, device="cuda"):
    tail = int(max(sigma * 4.0, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float32, device=device)
    t = 0.70710678 / abs(sigma)
    return (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)


def _blur(v, sigmas):      # v [C][x][y][z], separable, zero padding
    out = v[None]
    for ax, s in enumerate(sigmas):
        k = _gauss1d(s, v.device)
        shape = [1, 1, 1, 1, 1]
        shape[2 + ax] = k.numel()
        pad = [0, 0, 0]
        pad[ax] = k.numel() // 2
        C = out.shape[1]
        out = F.conv3d(out, k.view(shape).repeat(C, 1, 1, 1, 1), padding=pad, groups=C)
    return out[0]


@pytest.mark.gpu
def test_intensity_transforms_match_the_torch_restatement():
    raw = _raw(5, 2, (16, 16, 16), 3)
    size = (16, 16, 16)
    t = TrainTransformsGPU(spatial_size=size)
    base = ValTransformsGPU(spatial_size=size)(raw)
    ref_pts = np.linspace(0, 1, 10).tolist()
    flt_pts = [0.0, 0.05, 0.3, 0.32, 0.5, 0.52, 0.7, 0.9, 0.95, 1.0]
    params = [_q(shift=0.07, gamma=2.3), _q(smooth=[0.4, 1.2, 0.8]), _q(sharpen=([0.6, 0.9, 0.7], [0.5, 0.6, 0.55], 17.0)),
              _q(hist=(ref_pts, flt_pts)), _q(shift=-0.05, gamma=0.6, smooth=[1.5, 0.25, 1.0], sharpen=([1.0, 0.5, 0.8], [0.7, 0.5, 0.5], 11.0),
                                              hist=(ref_pts, flt_pts))]
    got = t.apply(raw, params)

    def restate(v, q):
        v = v + q["shift"]
        if q["gamma"] > 0:
            mn, rg = v.min(), v.max() - v.min()
            v = ((v - mn) / (rg + 1e-7)) ** q["gamma"] * rg + mn
        if q["smooth"] is not None:
            v = _blur(v, q["smooth"])
        if q["sharpen"] is not None:
            b1 = _blur(v, q["sharpen"][0])
            b2 = _blur(b1, q["sharpen"][1])
            v = b1 + q["sharpen"][2] * (b1 - b2)
        if q["hist"] is not None:
            mn, mx = float(v.min()), float(v.max())
            xp = np.array(q["hist"][0]) * (mx - mn) + mn
            fp = np.array(q["hist"][1]) * (mx - mn) + mn
            v = torch.from_numpy(np.interp(v.cpu().numpy().astype(np.float64), xp, fp)).float().cuda()
        return v

    for b, q in enumerate(params):
        ref = restate(base[b], q)
        scale = float(ref.abs().max())
        assert (got[b] - ref).abs().max() < 3e-5 * max(scale, 1.0), (b, (got[b] - ref).abs().max())


@pytest.mark.gpu
def test_gaussian_noise_statistics_and_reproducibility():
    raw = _raw(2, 1, (32, 32, 32), 4)
    t = TrainTransformsGPU(spatial_size=(32, 32, 32))
    base = ValTransformsGPU(spatial_size=(32, 32, 32))(raw)
    a = t.apply(raw, [_q(noise_std=0.04, seed=5), _q(noise_std=0.02, seed=6)])
    b = t.apply(raw, [_q(noise_std=0.04, seed=5), _q(noise_std=0.02, seed=7)])
    d = a - base
    assert abs(float(d[0].std()) - 0.04) < 0.002 and abs(float(d[0].mean())) < 0.001
    assert abs(float(d[1].std()) - 0.02) < 0.001
    assert torch.equal(a[0], b[0]) and not torch.equal(a[1], b[1])          # the stream is a function of (seed, element)
    # roughly normal: 4th standardized moment ~ 3
    z = d[0].flatten() / d[0].std()
    assert abs(float((z ** 4).mean()) - 3.0) < 0.15


@pytest.mark.gpu
def test_random_call_runs_and_stays_finite():
    raw = _raw(6, 2, (40, 36, 20), 5)
    out = TrainTransformsGPU(seed=0)(raw)
    assert out.shape == (6, 2, 64, 64, 64) and torch.isfinite(out).all()


# ---------------------------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_restatement_properties():
    """oracle/augment.py (numpy restatement of the published MONAI algorithms, parity unpinned) against exact identities."""
    from oracle import augment as OA
    from oracle import preprocess as OP
    rng = np.random.RandomState(0)
    q = _q(rotate=0.7, flip_axis=2, zoom=1.07)
    assert np.allclose(OA.affine_matrix(q["rotate"], q["flip_axis"], q["zoom"]), TrainTransformsGPU.affine(q))
    img = rng.rand(2, 12, 10, 8)
    # identity map: the resampling is the area resize alone
    ident = OA.resample(img, np.eye(3), (6, 5, 4))
    assert np.allclose(ident, OP.resize(img.astype(np.float32), (6, 5, 4)), atol=1e-6)
    # a pure flip commutes with the area resize
    flip = OA.resample(img, OA.affine_matrix(flip_axis=1), (6, 5, 4))
    assert np.allclose(flip, ident[:, :, ::-1, :], atol=1e-12)
    # Gaussian kernel: symmetric, non-negative, mass = erf of the truncation point; a constant stays constant away from the border
    for s in (0.25, 0.8, 1.5):
        k = OA.gaussian_1d(s)
        assert k.size == 2 * int(max(4 * s, 0.5) + 0.5) + 1 and np.allclose(k, k[::-1]) and (k >= 0).all() and 0.99 < k.sum() <= 1.0 + 1e-12
    const = OA.gaussian_smooth(np.ones((1, 16, 16, 16)), [1.0, 0.5, 1.5])
    assert np.allclose(const[0, 7:9, 7:9, 7:9], 1.0, atol=1e-3) and const[0, 0, 0, 0] < 0.5          # zero padding at the corner
    # contrast: the extremes are fixed points, gamma = 1 is the identity up to the 1e-7 guard
    v = rng.rand(1, 6, 6, 6) * 3 - 1
    c = OA.adjust_contrast(v, 2.0)
    assert abs(c.min() - v.min()) < 1e-6 and abs(c.max() - v.max()) < 1e-5 and np.allclose(OA.adjust_contrast(v, 1.0), v, atol=1e-6)
    # histogram shift: identical control points = identity; any increasing floating set keeps the order and the extremes
    ref = np.linspace(0, 1, 10)
    assert np.allclose(OA.histogram_shift(v, ref, ref), v, atol=1e-12)
    flt = np.array([0.0, 0.05, 0.3, 0.32, 0.5, 0.52, 0.7, 0.9, 0.95, 1.0])
    h = OA.histogram_shift(v, ref, flt)
    order = np.argsort(v.ravel())
    assert (np.diff(h.ravel()[order]) >= -1e-12).all() and abs(h.min() - v.min()) < 1e-12 and abs(h.max() - v.max()) < 1e-12
    # sharpen with alpha = 0 is the first blur
    assert np.allclose(OA.gaussian_sharpen(v, [0.6] * 3, [0.5] * 3, 0.0), OA.gaussian_smooth(v, [0.6] * 3))


def test_torch_restatement_of_the_gpu_tests_equals_the_oracle():
    """The GPU tests above check the CUDA path against torch restatements (`_spatial_ref`, `_blur`); this CPU test ties those to
    oracle/augment.py: same formulas, float64 numpy."""
    from oracle import augment as OA
    raw = _raw(2, 2, (16, 12, 8), 9, device="cpu")
    params = [_q(rotate=0.3), _q(rotate=-14.2, flip_axis=1, zoom=1.05)]
    ref = _spatial_ref(raw, params, (8, 6, 4)).numpy()
    n = _normalized(raw).numpy().astype(np.float64)
    for b, q in enumerate(params):
        o = OA.resample(n[b], OA.affine_matrix(q["rotate"], q["flip_axis"], q["zoom"]), (8, 6, 4))
        assert np.abs(o - ref[b]).max() < 1e-6
    v = torch.rand(2, 16, 16, 16, generator=torch.Generator().manual_seed(3))
    assert np.abs(OA.gaussian_smooth(v.numpy(), [0.4, 1.2, 0.8]) - _blur(v, [0.4, 1.2, 0.8]).numpy()).max() < 1e-6
    b1 = _blur(v, [0.6, 0.9, 0.7])
    sharp = (b1 + 17.0 * (b1 - _blur(b1, [0.5, 0.6, 0.55]))).numpy()
    assert np.abs(OA.gaussian_sharpen(v.numpy(), [0.6, 0.9, 0.7], [0.5, 0.6, 0.55], 17.0) - sharp).max() < 2e-5
