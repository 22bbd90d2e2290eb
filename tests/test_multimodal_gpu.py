"""End-to-end GPU parity of MultiModalModel + Cox / GradientBlender loss against the committed golden vectors
(generated from the UNCHANGED reference files, tests/golden/make_golden.py).
Tolerances: BASELINE.json north star -- logits within 2e-2 of the logit range, loss within 1e-3 (configs[0]);
fp16 activation storage / bf16 gradient storage / fp32 accumulation vs the fp32 reference (DESIGN.md "Numerics")."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(meta, sd):
    from mmnn_sts_b200.models.densenet import DenseNet121
    from mmnn_sts_b200.models.multimodal import MultiModalModel
    seed_w, seed_x, batch, cin, sx, sy, sz, blend, training, dropout, tie_free = meta
    m = MultiModalModel(DenseNet121(spatial_dims=3, in_channels=cin, out_channels=2, feature_channels=12,
                                    dropout_prob=0.2 if dropout else 0.0), ["x"] * 20, 2, 12, blend=bool(blend))
    m.load_state_dict(sd)
    return m.cuda()


@pytest.mark.parametrize("name", ["tiny_blend_train", "tiny_blend_train_dropout", "tiny_eval", "cfg1_train", "cfg1_eval", "odd_train"])
def test_against_reference_golden(name):
    from mmnn_sts_b200.losses.GradientBlender import GradientBlender
    from mmnn_sts_b200.losses.losses import CoxPH
    from mmnn_sts_b200.utils.utils import surv_criterion
    from oracle import synth
    g = np.load(os.path.join(GOLD, name + ".npz"))
    meta = [int(v) for v in g["meta"]]
    seed_w, seed_x, batch, cin, sx, sy, sz, blend, training, dropout, tie_free = meta
    sd = synth.make_state_dict(seed_w, in_channels=cin)
    image, clinical, events, durations = synth.make_batch(seed_x, batch, cin, (sx, sy, sz), tie_free=bool(tie_free))
    m = _build(meta, sd).train(bool(training))
    if training:
        masks = synth.make_masks(seed_x + 1000, batch) if dropout else synth.make_masks(0, batch, 0, 0, 0)
        m.image_model.model.backbone.injected_dropmask = torch.stack([masks["dense"][(b, l)] for b, nl in enumerate(synth.BLOCK_CONFIG) for l in range(nl)])
        m.image_model.model.features.injected_mask = masks["image_features"]
        m.clinical_model.model.injected_masks = torch.stack(masks["mlp"])
    with torch.set_grad_enabled(bool(training)):
        out = m({"image": image.cuda(), "clinical": clinical.cuda()})
    ref = torch.tensor(g["logits"])
    scale = float(ref.abs().max())
    err = float((out.detach().cpu() - ref).abs().max()) / scale
    print(f"\n{name}: logits max-abs err / max|logit| = {err:.3e}")
    # north star: rel 2e-2 on logits
    assert err < (2e-2 if training else 1e-2)
    if training:
        if blend:
            loss, _ = GradientBlender(CoxPH, survival=True, surv_criterion=surv_criterion).computeLoss(out, events.cuda(), durations.cuda())
        else:
            loss = surv_criterion(CoxPH, out, events.cuda(), durations.cuda(), "cuda")
        rel = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
        print(f"{name}: loss {loss.item():.5f} vs reference {float(g['loss']):.5f} (rel {rel:.3e})")
        assert rel < 2e-3      # north star: 1e-3 on the loss at configs[0] (asserted below for cfg1)
        if name.startswith('cfg1'):
            assert rel < 1e-3
        loss.backward()
        names = [str(s) for s in g["param_names"]]
        norms = dict(zip(names, g["grad_norms"]))
        ratios = []
        for k, p in m.named_parameters():
            if np.isnan(norms[k]):
                assert p.grad is None, f"{k}: the reference leaves this gradient None"
            else:
                assert p.grad is not None and torch.isfinite(p.grad).all(), k
                if norms[k] > 1e-8:
                    ratios.append(float(p.grad.double().norm()) / norms[k])
        print(f"{name}: gradient-norm ratio median {np.median(ratios):.3f}  [p5 {np.percentile(ratios, 5):.3f}, p95 {np.percentile(ratios, 95):.3f}]")
        assert 0.9 < np.median(ratios) < 1.1
        for k in ("output_head.weight", "clinical_model.model.backbone.dense0.weight", "image_model.model.features.feature_layer.weight"):
            a = p_grad = dict(m.named_parameters())[k].grad.cpu().double().flatten()
            b = torch.tensor(g["grad:" + k]).double().flatten()
            cos = float(a @ b / (a.norm() * b.norm()))
            print(f"   cosine({k}) = {cos:.4f}")
            assert cos > 0.98, k


# Stated gradient tolerance at the BENCHMARKED configuration.  Reference = the fp32 gradients of the
# unchanged reference (tests/golden/cfg2_train.npz: a strided subsample of <= 2048 elements of EVERY parameter gradient).
#   * per tensor:   cosine >= GRAD_COS_MIN  and  |g - g_ref|_2 / |g_ref|_2 <= GRAD_REL_MAX
#   * over tensors: median cosine >= GRAD_COS_MEDIAN, median rel-L2 <= GRAD_REL_MEDIAN
#   * calibration:  the reference ITSELF under torch.autocast(bfloat16) (tests/golden/cfg2_train_amp_bfloat16.json, made by
#                   tests/golden/amp_reference_gradient_error.py) sits at median rel-L2 0.211 / median cosine 0.978 / conv0.weight
#                   rel-L2 0.58 on this case: 16-bit activation storage flips the ReLU masks of pre-activations within rounding
#                   distance of zero in ANY implementation.  This build must stay at least 2x closer to fp32 than that.
#   * tensors whose reference gradient is analytically ZERO (a Linear bias feeding a BatchNorm; the head biases, because the
#     Cox gradient sums to zero over the risk set) hold rounding noise ~1e-8 in the reference: they must be noise here too.
#   * d(gamma) of norm0 is a ~3e7-term sum that cancels to 1e-4 of its terms (reference |g| 8.8e-5 next to |d(beta)| 6.7e-4):
#     ill-conditioned under ANY perturbation of the incoming gradient (2.5x its own norm under bf16 autocast), so it is held to
#     an ABSOLUTE bound relative to its companion d(beta) instead.
GRAD_COS_MIN, GRAD_REL_MAX = 0.96, 0.30
GRAD_COS_MEDIAN, GRAD_REL_MEDIAN = 0.995, 0.10
AMP_BF16_REL_MEDIAN = 0.211
ILL_CONDITIONED = {"image_model.model.backbone.norm0.weight": "image_model.model.backbone.norm0.bias"}


def _gsub_index(numel, gmax=2048):
    stride = max(1, -(-numel // gmax))
    return np.arange(0, numel, stride)


def _per_tensor_grad_errors(model, g):
    """[(name, cosine, rel-L2, |ref|)] of every parameter gradient against the golden subsample."""
    ref = torch.tensor(g["gsub"]).double()
    names = [str(s) for s in g["param_names"]]
    norms = dict(zip(names, g["grad_norms"]))
    out, off = [], 0
    for k, p in model.named_parameters():
        if np.isnan(norms[k]):
            assert p.grad is None, f"{k}: the reference leaves this gradient None"
            continue
        idx = torch.as_tensor(_gsub_index(p.numel()))
        r = ref[off:off + len(idx)]
        off += len(idx)
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        a = p.grad.detach().flatten().cpu().double()[idx]
        cos = float(a @ r / (a.norm() * r.norm() + 1e-300))
        rel = float((a - r).norm() / (r.norm() + 1e-300))
        out.append((k, cos, rel, float(r.norm())))
    assert off == len(ref)
    return out


def _cfg2_model_and_batch(g):
    from oracle import synth
    meta = [int(v) for v in g["meta"]]
    seed_w, seed_x, batch, cin, sx, sy, sz, blend, training, dropout, tie_free = meta
    sd = synth.make_state_dict(seed_w, in_channels=cin)
    image, clinical, events, durations = synth.make_batch(seed_x, batch, cin, (sx, sy, sz), tie_free=bool(tie_free))
    m = _build(meta, sd).train(bool(training))
    if training:
        masks = synth.make_masks(seed_x + 1000, batch) if dropout else synth.make_masks(0, batch, 0, 0, 0)
        m.image_model.model.backbone.injected_dropmask = torch.stack([masks["dense"][(b, l)] for b, nl in enumerate(synth.BLOCK_CONFIG) for l in range(nl)])
        m.image_model.model.features.injected_mask = masks["image_features"]
        m.clinical_model.model.injected_masks = torch.stack(masks["mlp"])
    return m, image, clinical, events, durations


def test_benchmarked_shape_cfg2_parity_and_determinism():
    """BASELINE configs[1] ITSELF (the shape bench.py quotes volumes/s on): B 16, 2x128x128x64, --blend, dropout 0.2 with
    injected masks, against the golden vectors of the UNCHANGED reference (/root/reference/main.py:445-469,
    models/multimodal.py:51-80).  North star: logits <= 2e-2, loss <= 1e-3, C-index of the risks bit-exact; gradients
    within the stated per-tensor tolerance; and the same input gives the SAME bits twice (ordered reductions)."""
    from mmnn_sts_b200 import main as M
    from mmnn_sts_b200.losses.GradientBlender import GradientBlender
    from mmnn_sts_b200.losses.losses import CoxPH
    from mmnn_sts_b200.utils.utils import surv_criterion
    from oracle import cindex
    g = np.load(os.path.join(GOLD, "cfg2_train.npz"))
    m, image, clinical, events, durations = _cfg2_model_and_batch(g)

    def run():
        m.zero_grad(set_to_none=True)
        out = m({"image": image.cuda(), "clinical": clinical.cuda()})
        loss, _ = GradientBlender(CoxPH, survival=True, surv_criterion=surv_criterion).computeLoss(out, events.cuda(), durations.cuda())
        loss.backward()
        torch.cuda.synchronize()
        return out.detach().clone(), loss.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}

    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out, loss, grads = run()
    sd_after = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ref = torch.tensor(g["logits"])
    err = float((out.cpu() - ref).abs().max()) / float(ref.abs().max())
    rel = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
    print(f"\ncfg2_train: logits max-abs err / max|logit| = {err:.3e}; loss {loss.item():.6f} vs reference {float(g['loss']):.6f} (rel {rel:.3e})")
    assert err < 2e-2
    assert rel < 1e-3
    # C-index of the risks: the GPU kernel's integer counts equal the oracle sweep on the SAME scores (bit-exact), and the
    # build's risks rank the patients exactly as the reference's do (same counts as from the golden logits)
    risks = out[0]
    for c in range(2):
        got = tuple(int(v) for v in M.concordance_counts(durations[:, c].cuda(), risks[:, c], events[:, c].cuda())[0].cpu())
        want = cindex.concordance_counts(durations[:, c].numpy(), risks[:, c].cpu().numpy(), events[:, c].numpy())
        gold = cindex.concordance_counts(durations[:, c].numpy(), g["logits"][0][:, c], events[:, c].numpy())
        assert got == want == gold, (c, got, want, gold)
    assert M.getCIndices(risks, events.cuda(), durations.cuda()) == cindex.getCIndices(g["logits"][0], events.numpy(), durations.numpy())
    # gradients, per tensor
    errs = _per_tensor_grad_errors(m, g)
    med_norm = float(np.median([n for _, _, _, n in errs]))
    null = [e for e in errs if e[3] < 1e-5 * med_norm]
    ill = [e for e in errs if e[0] in ILL_CONDITIONED]
    reg = [e for e in errs if e[3] >= 1e-5 * med_norm and e[0] not in ILL_CONDITIONED]
    worst_cos = sorted(reg, key=lambda e: e[1])[:5]
    worst_rel = sorted(reg, key=lambda e: -e[2])[:5]
    cos_med, rel_med = float(np.median([e[1] for e in reg])), float(np.median([e[2] for e in reg]))
    print("   gradients (%d tensors + %d analytically-zero + %d ill-conditioned): cosine median %.5f min %.5f; rel-L2 median %.3e max %.3e" % (
        len(reg), len(null), len(ill), cos_med, worst_cos[0][1], rel_med, worst_rel[0][2]))
    print("   worst cosine:", [(k.replace("image_model.model.backbone.", ""), f"{c:.4f}") for k, c, r, n in worst_cos])
    print("   worst rel-L2:", [(k.replace("image_model.model.backbone.", ""), f"{r:.3e}") for k, c, r, n in worst_rel])
    named = dict(m.named_parameters())
    for k, c, r, n in ill:
        comp = [e for e in errs if e[0] == ILL_CONDITIONED[k]][0]
        print(f"   ill-conditioned {k}: |g - g_ref| = {r * n:.3e} = {r * n / comp[3]:.2f} x |d(beta)_ref|")
    # determinism: same weights, same inputs -> bit-identical logits, loss and every gradient
    m.load_state_dict(sd0)
    out2, loss2, grads2 = run()
    diff = [k for k in grads if not torch.equal(grads[k], grads2[k])]
    print(f"   second run of the same input: logits identical {torch.equal(out, out2)}, loss identical {torch.equal(loss, loss2)}, "
          f"{len(diff)} of {len(grads)} gradient tensors differ")
    bad = [(k, c, r) for k, c, r, n in reg if not (c >= GRAD_COS_MIN and r <= GRAD_REL_MAX)]
    assert not bad, bad[:10]
    assert cos_med >= GRAD_COS_MEDIAN and rel_med <= GRAD_REL_MEDIAN, (cos_med, rel_med)
    assert rel_med <= 0.5 * AMP_BF16_REL_MEDIAN
    assert len(null) == 9, [e[0] for e in null]
    for k, c, r, n in null:
        assert float(named[k].grad.double().norm()) < 1e-4 * med_norm, k
    for k, c, r, n in ill:
        comp = [e for e in errs if e[0] == ILL_CONDITIONED[k]][0]
        assert r * n <= 0.5 * comp[3], (k, r * n, comp[3])
    # running statistics of the step
    new_sd = {k: v for k, v in sd_after.items()}
    for k in ["image_model.model.backbone.norm0", "image_model.model.backbone.denseblock2.denselayer3.layers.norm2",
              "image_model.model.backbone.norm5", "clinical_model.model.backbone.bn0"]:
        for kind, key in (("rm", ".running_mean"), ("rv", ".running_var")):
            a, b = new_sd[k + key].cpu().double(), torch.tensor(g[kind + ":" + k]).double()
            assert float((a - b).norm() / b.norm()) < 5e-3, (k, kind)
    assert torch.equal(out, out2) and torch.equal(loss, loss2)
    assert not diff, f"{len(diff)} gradient tensors differ between two runs of the same input: {diff[:5]}"


def test_benchmarked_shape_cfg2_eval():
    g = np.load(os.path.join(GOLD, "cfg2_eval.npz"))
    m, image, clinical, events, durations = _cfg2_model_and_batch(g)
    with torch.no_grad():
        out = m({"image": image.cuda(), "clinical": clinical.cuda()})
    ref = torch.tensor(g["logits"])
    err = float((out.cpu() - ref).abs().max()) / float(ref.abs().max())
    print(f"\ncfg2_eval: logits max-abs err / max|logit| = {err:.3e}")
    assert err < 1e-2


def test_cindex_of_risks_bit_exact_and_bootstrap():
    """C-index computed from the build's own risk scores: GPU counts == CPU oracle counts on the same scores
    (integers), C-index equal as float64; bootstrap mean/std equal to the oracle's loop."""
    from mmnn_sts_b200 import main as M
    from oracle import cindex, synth
    rng = np.random.RandomState(3)
    n = 400
    preds = torch.tensor(np.round(rng.randn(n, 2), 2), dtype=torch.float32)
    events = torch.tensor(rng.randint(0, 2, (n, 2))); durations = torch.tensor(rng.randint(1, 200, (n, 2)))
    got = M.getCIndices(preds.cuda(), events.cuda(), durations.cuda())
    ref = cindex.getCIndices(preds.numpy(), events.numpy(), durations.numpy())
    assert got == ref
    idx = np.stack([rng.randint(0, n, n) for _ in range(20)])
    c, mean, std, _ = M.bootstrap_cindices(preds.cuda(), events.cuda(), durations.cuda(), torch.tensor(idx).cuda())
    c_ref, mean_ref, std_ref = cindex.bootstrap_cindex(preds.numpy(), events.numpy(), durations.numpy(), idx)
    assert np.array_equal(c, c_ref) and np.array_equal(mean, mean_ref) and np.array_equal(std, std_ref)
    with pytest.raises(ZeroDivisionError):
        M.getCIndices(preds.cuda(), torch.zeros_like(events).cuda(), durations.cuda())


def test_train_survival_trajectory_matches_the_reference_loop():
    """mmnn_sts_b200.main.train_survival against the trajectory of the reference's loop body (/root/reference/main.py:402-414 setup,
    :445-481 accumulate / step, :509-569 validation, :584-588 blending-weight update) run with the UNCHANGED reference classes
    (tests/golden/trajectory.npz): 72 patients in micro-batches of 8, 2 epochs.  Checked: WHEN the optimiser steps (after the 8th
    micro-batch = 64 patients and after the last one), the OneCycleLR lr / momentum after every step, per-epoch train / validation
    loss, head-0 C-indices, blending weights after each update, running statistics bookkeeping, and the parameter UPDATE of tensors
    from every part of the network."""
    from mmnn_sts_b200 import main as M
    from mmnn_sts_b200.models.densenet import DenseNet121
    from mmnn_sts_b200.models.multimodal import MultiModalModel
    from oracle import train_loop
    g = np.load(os.path.join(GOLD, "trajectory.npz"))
    args, train, val, sd = train_loop.trajectory_case()
    m = MultiModalModel(DenseNet121(spatial_dims=3, in_channels=1, out_channels=2, feature_channels=12, dropout_prob=0.0), ["x"] * 20, 2, 12, blend=True)
    m.load_state_dict(sd)
    m.clinical_model.model.dropout_prob = 0.0
    hist = M.train_survival(m, train, val, args, torch.device("cuda"))
    assert [list(x) for x in hist.step_at] == g["step_at"].tolist()                 # [[0, 7], [0, 8], [1, 7], [1, 8]]
    assert hist.lr_trace == g["lr_trace"].tolist() and hist.momentum_trace == g["momentum_trace"].tolist()
    print("\ntrajectory: train loss", hist.train_loss, "vs", g["train_loss"].tolist(), "; val loss", hist.val_loss, "vs", g["val_loss"].tolist())
    np.testing.assert_allclose(hist.train_loss, g["train_loss"], rtol=2e-3)
    np.testing.assert_allclose(hist.val_loss, g["val_loss"], rtol=2e-3)
    print("   C-index train", hist.train_c, "vs", g["train_c"].tolist(), "; val", hist.val_c, "vs", g["val_c"].tolist())
    np.testing.assert_allclose(np.array(hist.train_c), g["train_c"], atol=0.02)
    # 16 validation patients = 37 / 65 admissible pairs per class: ONE pair of near-tied eval-mode risks swapping moves the index by
    # 0.027 / 0.015, so this is a 3-pair tolerance
    np.testing.assert_allclose(np.array(hist.val_c), g["val_c"], atol=0.085)
    w = np.array(hist.blender.history)
    print("   blending weights", w.tolist(), "vs", g["blender_weights"].tolist())
    assert w.shape == g["blender_weights"].shape
    np.testing.assert_allclose(w[0], g["blender_weights"][0], atol=1e-6)            # first update: uniform
    np.testing.assert_allclose(w[1], g["blender_weights"][1], atol=0.05)            # softmax(dG / dO^2): ratios of loss DIFFERENCES
    new_sd = m.state_dict()
    assert int(new_sd["image_model.model.backbone.norm0.num_batches_tracked"]) == int(g["nbt:norm0"]) == 18
    a, b = new_sd["image_model.model.backbone.norm5.running_mean"].cpu().double(), torch.tensor(g["rm:norm5"]).double()
    assert float((a - b).norm() / b.norm()) < 1e-2
    worst = []
    for k in train_loop.TRACKED:
        w0, ref, got = sd[k].double(), torch.tensor(g["final:" + k]).double(), new_sd[k].cpu().double()
        du_ref, du = (ref - w0).flatten(), (got - w0).flatten()
        cos = float(du @ du_ref / (du.norm() * du_ref.norm()))
        rel = float((du - du_ref).norm() / du_ref.norm())
        worst.append((k, cos, rel))
    print("   parameter updates (cosine, rel-L2):", [(".".join(k.split(".")[-3:-1]), f"{c:.4f}", f"{r:.3f}") for k, c, r in worst])
    for k, cos, rel in worst:
        assert cos > 0.95 and rel < 0.35, (k, cos, rel)
    res = M.inference_survival(m, train, torch.device("cuda"), bootstrap=True, num_resamples=10, seed=1)
    assert res.per_resample.shape == (10, 2)


def test_bootstrap_cindex_config5_scale():
    """BASELINE configs[4] scale: 10 000 patients, 1000 resamples (indices np.random.RandomState(42+r).randint).
    The GPU counts of a sample of resamples are compared bit-exactly with the CPU oracle sweep."""
    import time
    from mmnn_sts_b200.ops import concordance_counts
    from oracle import cindex
    rng = np.random.RandomState(0)
    n, R = 10000, 1000
    risks = np.round(rng.randn(n), 3).astype(np.float32)
    t = rng.randint(1, 3651, n); e = rng.randint(0, 2, n)
    idx = np.stack([np.random.RandomState(42 + r).randint(0, n, n) for r in range(R)])
    args = [torch.tensor(v, device="cuda") for v in (t, risks, e)]
    concordance_counts(*args, torch.tensor(idx[:4], device="cuda"))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    counts = concordance_counts(*args, torch.tensor(idx, device="cuda")).cpu().numpy()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"\n10k patients x 1000 resamples on the GPU: {dt * 1e3:.1f} ms")
    for r in (0, 1, 499, 999):
        assert tuple(counts[r]) == cindex.concordance_counts(t[idx[r]], risks[idx[r]], e[idx[r]]), r
    assert (counts[:, 2] > 0).all()


def test_cuda_graph_step_matches_eager():
    """A captured training step (forward, blended Cox loss, backward, SGD) replays to the same parameters as the eager
    step: the C-ABI path is allocation- and sync-free, side-stream fork/join included."""
    from mmnn_sts_b200.graph import GraphedTrainStep
    from mmnn_sts_b200.losses.GradientBlender import GradientBlender
    from mmnn_sts_b200.losses.losses import CoxPH
    from mmnn_sts_b200.models.densenet import DenseNet121
    from mmnn_sts_b200.models.multimodal import MultiModalModel
    from mmnn_sts_b200.utils.utils import surv_criterion
    from oracle import synth
    sd = synth.make_state_dict(42, in_channels=1)
    batches = [[t.cuda() for t in synth.make_batch(50 + i, 4, 1, (64, 64, 32))] for i in range(3)]

    def build():
        m = MultiModalModel(DenseNet121(spatial_dims=3, in_channels=1, out_channels=2, feature_channels=12, dropout_prob=0.0), ["x"] * 20, 2, 12, blend=True)
        m.load_state_dict(sd)
        m = m.cuda().train()
        m.clinical_model.model.dropout_prob = 0.0
        opt = torch.optim.SGD(m.parameters(), 1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)
        gb = GradientBlender(CoxPH, survival=True, surv_criterion=surv_criterion)

        def step(im, cl, ev, du):
            out = m({"image": im, "clinical": cl})
            loss, _ = gb.computeLoss(out, ev, du)
            loss.backward()
            opt.step(); opt.zero_grad(set_to_none=True)
            return loss.detach()
        return m, step

    m1, step1 = build()
    for _ in range(3):
        step1(*batches[0])                 # same warm-up steps as the graphed run performs
    losses1 = [float(step1(*b)) for b in batches]
    m2, step2 = build()
    g = GraphedTrainStep(step2, batches[0], warmup=3)
    losses2 = [float(g(*b)) for b in batches]
    torch.cuda.synchronize()
    for a, b in zip(losses1, losses2):
        assert abs(a - b) < 2e-3 * max(1.0, abs(a)), (losses1, losses2)
    # compare the UPDATES (run-to-run atomics order + ReLU-flip sensitivity make the last digits of a gradient differ)
    for key in ("image_model.model.backbone.conv0.weight", "image_model.model.backbone.denseblock4.denselayer16.layers.conv2.weight", "output_head.weight"):
        w0 = sd[key].cuda().double()
        u1 = (dict(m1.named_parameters())[key].double() - w0).flatten(); u2 = (dict(m2.named_parameters())[key].double() - w0).flatten()
        cos = float((u1 @ u2).detach() / (u1.norm() * u2.norm()).detach())
        # the stem gradient sits behind all 121 ReLU layers: two runs of 6 chained steps already differ by a few percent
        assert cos > (0.9 if "conv0" in key else 0.98) and 0.9 < float((u1.norm() / u2.norm()).detach()) < 1.1, (key, cos)


def test_fp16_volumes_give_bit_identical_results():
    """A loader that ships 16-bit volumes (mmnn_encoder_forward_f16): the stem rounds the image to the activation format first
    thing, so fp16 input == fp32 input rounded, bit for bit, in logits and in the stem's weight gradient."""
    from mmnn_sts_b200.losses.losses import CoxPH
    from mmnn_sts_b200.utils.utils import surv_criterion
    from oracle import synth
    sd = synth.make_state_dict(42, in_channels=2)
    image, clinical, events, durations = synth.make_batch(31, 4, 2, (64, 64, 32))
    outs = []
    for dt in (torch.float32, torch.float16):
        m = _build([42, 31, 4, 2, 64, 64, 32, 1, 1, 0, 1], sd).train()
        m.clinical_model.model.dropout_prob = 0.0
        out = m({"image": image.to(dt).cuda(), "clinical": clinical.cuda()})
        surv_criterion(CoxPH, out[0], events.cuda(), durations.cuda(), "cuda").backward()
        outs.append((out.detach().clone(), m.image_model.model.backbone.conv0.weight.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
