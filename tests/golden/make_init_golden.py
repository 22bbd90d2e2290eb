"""SHA-1 of every tensor of the UNCHANGED reference MultiModalModel's state_dict built under torch.manual_seed(42)
(/root/reference/models/densenet.py:258-265 initialisation law, /root/reference/parser/parser.py:106-113,162-168 construction
order).  Build-container only (needs /root/reference);  writes tests/golden/init_seed42.json.

    python tests/golden/make_init_golden.py
"""
import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shim  # noqa: E402


def digest(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


if __name__ == "__main__":
    ns = shim.load_reference()
    out = {}
    for cin in (1, 2):
        torch.manual_seed(42)
        dn = ns.densenet.DenseNet121(spatial_dims=3, in_channels=cin, out_channels=2, feature_channels=12, dropout_prob=0.2)
        mm = ns.multimodal.MultiModalModel(dn, ["x"] * 20, 2, 12, blend=True)
        out[str(cin)] = {k: digest(v) for k, v in mm.state_dict().items()}
        print(cin, len(out[str(cin)]), "tensors")
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "init_seed42.json"), "w"))
