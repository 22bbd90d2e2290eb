"""Generates tests/golden/preprocess.npz: small raw volumes and the output of the UNCHANGED reference `Normalize`
(/root/reference/utils/utils.py:346-355, imported under oracle/shim.py stubs) followed by the oracle's restatements of
MONAI ScaleIntensity / Resize.  Run in the build container (needs /root/reference):  python tests/golden/make_preprocess_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import preprocess as op, shim  # noqa: E402

Normalize = shim.load_reference().utils.Normalize      # the reference's class, unchanged

MEAN, STD = 286.90859071507913, 581.7816096485366
rng = np.random.RandomState(11)
out = {}
cases = {"mri_like": (2, 20, 18, 12, (8, 8, 6)), "upsample": (1, 5, 6, 7, (8, 9, 10)), "odd": (2, 13, 11, 9, (4, 5, 3))}
for name, (c, x, y, z, size) in cases.items():
    raw = (rng.gamma(2.0, 200.0, size=(c, x, y, z)) * (rng.rand(c, x, y, z) > 0.2)).astype(np.float32)
    ref_norm = Normalize(MEAN, STD)(raw)
    assert np.array_equal(ref_norm, op.normalize(raw, MEAN, STD)), "oracle normalize differs from the reference class"
    out[name + "_raw"] = raw
    out[name + "_normalized"] = ref_norm.astype(np.float32)
    out[name + "_out"] = op.resize(op.scale_intensity(ref_norm), size).astype(np.float32)
    out[name + "_size"] = np.asarray(size)
out["negative_raw"] = -np.abs(rng.randn(1, 6, 6, 6)).astype(np.float32) - 1.0      # max < 0: Normalize flips the order
out["negative_normalized"] = Normalize(MEAN, STD)(out["negative_raw"]).astype(np.float32)
out["negative_out"] = op.resize(op.scale_intensity(out["negative_normalized"]), (3, 3, 3)).astype(np.float32)
out["negative_size"] = np.asarray((3, 3, 3))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "preprocess.npz"), **out)
print("wrote preprocess.npz:", {k: v.shape for k, v in out.items()})
