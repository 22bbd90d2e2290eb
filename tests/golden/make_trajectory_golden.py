"""tests/golden/trajectory.npz: the reference's survival training loop (body of /root/reference/main.py:385-601, restated in
oracle/train_loop.py because main.py cannot be imported) driven with the UNCHANGED reference classes -- MultiModalModel,
DenseNet121, GradientBlender, CoxPH, surv_criterion -- under oracle/shim.py.  Build container only.

    python tests/golden/make_trajectory_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cindex, shim, train_loop  # noqa: E402

def run(ns, mini):
    args, train, val, sd = train_loop.trajectory_case(mini)
    dn = ns.densenet.DenseNet121(spatial_dims=3, in_channels=1, out_channels=2, feature_channels=12, dropout_prob=0.0)
    mm = ns.multimodal.MultiModalModel(dn, ["x"] * 20, 2, 12, blend=True)
    mm.load_state_dict(sd)
    for mod in list(mm.modules()):                       # the clinical MLP's Dropout1d(p = 0.2) modules -> p = 0 (deterministic run)
        for name, child in list(mod.named_children()):
            if isinstance(child, (torch.nn.Dropout, torch.nn.Dropout1d, torch.nn.Dropout3d)):
                child.p = 0.0
    gb = ns.blender.GradientBlender(ns.losses.CoxPH, survival=True, surv_criterion=ns.utils.surv_criterion)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hist = train_loop.train_survival_loop(mm, train, val, args, gb, ns.utils.surv_criterion, ns.losses.CoxPH, cindex.getCIndices)
    new_sd = mm.state_dict()
    res = {"train_loss": np.array(hist.train_loss), "val_loss": np.array(hist.val_loss), "train_c": np.array(hist.train_c),
           "val_c": np.array(hist.val_c), "lr_trace": np.array(hist.lr_trace), "momentum_trace": np.array(hist.momentum_trace),
           "step_at": np.array(hist.step_at), "selection": np.array(hist.selection), "blender_weights": np.array(hist.blender_weights)}
    for k in train_loop.TRACKED:
        res["final:" + k] = new_sd[k].detach().numpy()
    res["rm:norm5"] = new_sd["image_model.model.backbone.norm5.running_mean"].numpy()
    res["nbt:norm0"] = new_sd["image_model.model.backbone.norm0.num_batches_tracked"].numpy()
    name = "trajectory_mini.npz" if mini else "trajectory.npz"
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name), **res)
    for k in ("train_loss", "val_loss", "train_c", "val_c", "lr_trace", "momentum_trace", "step_at", "blender_weights"):
        print(name, k, res[k].tolist())


if __name__ == "__main__":
    torch.set_num_threads(8)
    ns = shim.load_reference()
    if "--full-only" not in sys.argv:
        run(ns, True)
    if "--mini-only" not in sys.argv:
        run(ns, False)
