"""classification loss path (BCEWithLogitsLoss(pos_weight), GradientBlender classification
branch, F1 counters).  CPU: the oracle against torch.nn.BCEWithLogitsLoss and the reference's unchanged GradientBlender
(when /root/reference is present).  GPU: fused kernel against the oracle, fp32 tolerance 1e-6 relative; counters exact."""
import pytest
import torch


def _case(seed=0, H=3, N=37, C=5):
    g = torch.Generator().manual_seed(seed)
    preds = torch.randn(H, N, C, generator=g) * 3
    targets = torch.randint(0, 2, (N, C), generator=g).float()
    freqs = torch.rand(C, generator=g) * 0.8 + 0.1
    return preds, targets, (1 - freqs) / freqs


def test_oracle_bce_matches_torch():
    from oracle import classification as oc
    preds, targets, pw = _case()
    ref = torch.nn.BCEWithLogitsLoss(pos_weight=pw, reduction="none")(preds, targets.expand_as(preds))
    torch.testing.assert_close(oc.bce_with_logits(preds, targets, pw).float(), ref, rtol=1e-6, atol=1e-6)
    big = torch.tensor([[[-80.0, 80.0, 0.0]]]); y = torch.tensor([[1.0, 0.0, 1.0]])
    assert torch.isfinite(oc.bce_with_logits(big, y)).all()


def test_oracle_blender_matches_reference_class():
    from oracle import classification as oc, shim
    if not shim.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    ref_mod = shim.load_reference().blender
    preds, targets, pw = _case(1)
    vp, vt, _ = _case(2, N=21)
    lf = torch.nn.BCEWithLogitsLoss(pos_weight=pw, reduction="none")
    ref = ref_mod.GradientBlender(lf, survival=False, reduction="sum", device="cpu")
    mine = oc.GradientBlenderClassOracle(lf, "sum")
    assert torch.equal(ref.computeLoss(preds, targets), mine.computeLoss(preds, targets))
    for k in range(3):
        ref.updateWeights(preds * (1 + 0.1 * k), targets, vp * (1 - 0.05 * k), vt)
        mine.updateWeights(preds * (1 + 0.1 * k), targets, vp * (1 - 0.05 * k), vt)
        assert torch.equal(ref.weights, mine.weights), k
    assert torch.equal(ref.computeLoss(preds, targets, reduceToHeads=True), mine.computeLoss(preds, targets, reduceToHeads=True))


def test_f1_helpers_match_reference_formula():
    from mmnn_sts_b200.main import classification_pos_weights, getF1Score
    from oracle import classification as oc
    tps, fps, fns = torch.tensor([3, 0, 5]), torch.tensor([1, 2, 0]), torch.tensor([2, 1, 0])
    assert getF1Score(tps, fps, fns) == oc.f1_score(tps, fps, fns)
    assert getF1Score(tps, fps, fns) == pytest.approx([3 / 4.5, 0.0, 1.0], rel=1e-6)
    f = torch.tensor([0.25, 0.5])
    assert torch.equal(classification_pos_weights(f), torch.tensor([3.0, 1.0]))


@pytest.mark.gpu
def test_gpu_bce_blender_and_counts():
    from mmnn_sts_b200.losses.GradientBlender import GradientBlender
    from mmnn_sts_b200.losses.losses import BCEWithLogitsLoss
    from oracle import classification as oc
    preds, targets, pw = _case(3, N=301, C=7)
    dev = torch.device("cuda", 0)
    p = preds.to(dev).requires_grad_(True)
    lf = BCEWithLogitsLoss(pos_weight=pw.to(dev), reduction="none", count_threshold=0.5).to(dev)
    el = lf(p, targets.to(dev))
    torch.testing.assert_close(el.detach().cpu().double(), oc.bce_with_logits(preds, targets, pw), rtol=2e-6, atol=2e-6)
    tps, fps, fns = oc.f1_counts(preds[0], targets)
    assert torch.equal(lf.last_counts.cpu().long(), torch.stack([tps, fps, fns]))
    # blended loss + gradient against autograd through torch's own loss under the oracle blender
    gb = GradientBlender(BCEWithLogitsLoss(pos_weight=pw.to(dev), reduction="none").to(dev), survival=False, reduction="sum")
    loss = gb.computeLoss(p, targets.to(dev))
    loss.backward()
    pr = preds.clone().requires_grad_(True)
    ob = oc.GradientBlenderClassOracle(torch.nn.BCEWithLogitsLoss(pos_weight=pw, reduction="none"), "sum")
    lr = ob.computeLoss(pr, targets)
    lr.backward()
    torch.testing.assert_close(loss.detach().cpu(), lr.detach(), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(p.grad.cpu(), pr.grad, rtol=1e-5, atol=1e-6)
    vp, vt, _ = _case(4, N=55, C=7)
    for k in range(3):
        gb.updateWeights(p.detach() * (1 + 0.1 * k), targets.to(dev), vp.to(dev) * (1 - 0.05 * k), vt.to(dev))
        ob.updateWeights(preds * (1 + 0.1 * k), targets, vp * (1 - 0.05 * k), vt)
        torch.testing.assert_close(gb.weights.cpu(), ob.weights, rtol=1e-3, atol=1e-4)
    # plain (non-blended) path of main.py:210: criterion(train_loss_function, outputs, labels, device) with reduction='sum'
    s = BCEWithLogitsLoss(pos_weight=pw.to(dev), reduction="sum").to(dev)(p.detach()[0], targets.to(dev))
    torch.testing.assert_close(s.cpu().double(), oc.bce_with_logits(preds[0], targets, pw).sum(), rtol=1e-5, atol=1e-4)


def test_cpu_logits_are_refused():
    from mmnn_sts_b200 import _lib as L
    from mmnn_sts_b200.losses.losses import BCEWithLogitsLoss
    with pytest.raises(L.MMNNLibraryError):
        BCEWithLogitsLoss()(torch.zeros(2, 3), torch.zeros(2, 3))
