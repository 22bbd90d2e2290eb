"""How far do the gradients of the UNCHANGED reference move when it runs in 16-bit autocast instead of fp32?

    python tests/golden/amp_reference_gradient_error.py [bfloat16|float16] [cfg2|tiny]

Build-container only (imports /root/reference under oracle/shim.py).  Runs the reference MultiModalModel + GradientBlender
(/root/reference/main.py:445-469) on the inputs / weights / dropout masks of a golden case under
torch.autocast('cpu', dtype) and compares every parameter gradient with the fp32 gradients stored in the golden file
(same strided subsample).  The result calibrates the gradient tolerance stated in DESIGN.md section 5: a ReLU network whose
activations are stored in 16 bits flips the masks of pre-activations within rounding distance of zero, which moves the
gradient by ~sqrt(fraction flipped) per layer in ANY implementation -- the reference's own AMP path included.
Writes tests/golden/<case>_amp_<dtype>.json (per-tensor cosine / rel-L2 summary)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import shim, synth  # noqa: E402
import make_golden as mg  # noqa: E402

dtype = getattr(torch, sys.argv[1] if len(sys.argv) > 1 else "bfloat16")
case = (sys.argv[2] if len(sys.argv) > 2 else "cfg2") + "_train"
case = {"cfg2_train": "cfg2_train", "tiny_train": "tiny_blend_train_dropout"}[case]
torch.set_num_threads(8)
ns = shim.load_reference()
g = np.load(os.path.join(mg.OUT, case + ".npz"))
seed_w, seed_x, batch, cin, sx, sy, sz, blend, training, dropout, tie_free = [int(v) for v in g["meta"]]
sd = synth.make_state_dict(seed_w, in_channels=cin)
mm = mg.build_reference(ns, cin, bool(blend), 0.2 if dropout else 0.0, sd)
image, clinical, events, durations = synth.make_batch(seed_x, batch, cin, (sx, sy, sz), tie_free=bool(tie_free))
masks = synth.make_masks(seed_x + 1000, batch) if dropout else synth.make_masks(0, batch, 0, 0, 0)
mm.train(True)
mg.inject_masks(mm, masks)
with torch.autocast("cpu", dtype=dtype):
    out = mm({"image": image, "clinical": clinical})
gb = ns.blender.GradientBlender(ns.losses.CoxPH, survival=True, surv_criterion=ns.utils.surv_criterion)
loss, _ = gb.computeLoss(out.float(), events, durations)
loss.backward()
ref_logits = torch.tensor(g["logits"])
res = {"dtype": str(dtype), "case": case,
       "logits_err": float((out.float() - ref_logits).abs().max() / ref_logits.abs().max()),
       "loss_rel": abs(float(loss) - float(g["loss"])) / abs(float(g["loss"])), "tensors": {}}
if "gsub" in g.files:
    ref = torch.tensor(g["gsub"]).double()
    off = 0
    for k, p in mm.named_parameters():
        if p.grad is None:
            continue
        idx = torch.as_tensor(mg.gsub_index(p.numel()))
        r = ref[off:off + len(idx)]; off += len(idx)
        a = p.grad.flatten().double()[idx]
        res["tensors"][k] = [float(a @ r / (a.norm() * r.norm() + 1e-300)), float((a - r).norm() / (r.norm() + 1e-300)), float(r.norm())]
else:
    for k, p in mm.named_parameters():
        if "grad:" + k in g.files:
            r = torch.tensor(g["grad:" + k]).flatten().double(); a = p.grad.flatten().double()
            res["tensors"][k] = [float(a @ r / (a.norm() * r.norm() + 1e-300)), float((a - r).norm() / (r.norm() + 1e-300)), float(r.norm())]
cos = np.array([v[0] for v in res["tensors"].values()]); rel = np.array([v[1] for v in res["tensors"].values()])
res["summary"] = {"cos_median": float(np.median(cos)), "cos_p5": float(np.percentile(cos, 5)), "rel_median": float(np.median(rel)), "rel_p95": float(np.percentile(rel, 95))}
print(json.dumps({k: v for k, v in res.items() if k != "tensors"}, indent=1))
json.dump(res, open(os.path.join(mg.OUT, f"{case}_amp_{sys.argv[1] if len(sys.argv) > 1 else 'bfloat16'}.json"), "w"))
