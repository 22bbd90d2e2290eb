"""Import the UNCHANGED reference hot-path files from /root/reference under stubbed third-party modules.

TEST INFRASTRUCTURE (build container only: /root/reference does not exist on the GPU box).
Used by tests/golden/make_golden.py to generate the committed golden vectors and by the optional
"reference present" tests that validate oracle/model.py against the real thing.

What is stubbed: monai, medcam, boto3, botocore, matplotlib, nibabel, SimpleITK,
skmultilearn, torch_lr_finder, lifelines, pycox, s3fs -> MagicMock;  five functional shims map the MONAI
layer factories used at /root/reference/models/densenet.py:71-85,142-148,190-203,223,237-238 to torch.nn
(monai-weekly==1.2.dev2313 semantics), and pycox.models.loss.CoxPHLoss is restated (oracle/cox.py).
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MMNN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "densenet.py"))


class _Factory:
    """monai.networks.layers.factories.LayerFactory look-alike: Conv[Conv.CONV, 3] -> nn.Conv3d."""

    def __init__(self, table):
        self._table = table
        for name in table:
            setattr(self, name.upper(), name.upper())

    def __getitem__(self, key):
        name, dim = key
        return self._table[name.lower()][dim - 1]


def _get_norm_layer(name, spatial_dims=1, channels=1):
    assert name == "batch", name
    return (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)[spatial_dims - 1](channels)


def _get_act_layer(name):
    act, kwargs = name if isinstance(name, tuple) else (name, {})
    assert act == "relu", act
    return nn.ReLU(**kwargs)


def install_stubs():
    for mod in ["monai", "monai.networks", "monai.networks.nets", "monai.networks.layers",
                "monai.networks.layers.factories", "monai.networks.layers.utils", "monai.utils",
                "monai.utils.module", "monai.utils.type_conversion", "monai.transforms", "monai.data",
                "medcam", "boto3", "botocore", "botocore.exceptions", "matplotlib", "matplotlib.pyplot",
                "nibabel", "SimpleITK", "skmultilearn", "skmultilearn.model_selection", "torch_lr_finder",
                "lifelines", "lifelines.utils", "pycox", "pycox.models", "pycox.models.loss", "s3fs",
                "torchvision"]:
        if mod not in sys.modules or isinstance(sys.modules[mod], MagicMock):
            sys.modules[mod] = MagicMock(name=mod)

    fac = types.ModuleType("monai.networks.layers.factories")
    fac.Conv = _Factory({"conv": (nn.Conv1d, nn.Conv2d, nn.Conv3d)})
    fac.Pool = _Factory({"max": (nn.MaxPool1d, nn.MaxPool2d, nn.MaxPool3d),
                         "avg": (nn.AvgPool1d, nn.AvgPool2d, nn.AvgPool3d),
                         "adaptiveavg": (nn.AdaptiveAvgPool1d, nn.AdaptiveAvgPool2d, nn.AdaptiveAvgPool3d)})
    fac.Dropout = _Factory({"dropout": (nn.Dropout, nn.Dropout2d, nn.Dropout3d)})
    sys.modules["monai.networks.layers.factories"] = fac

    ut = types.ModuleType("monai.networks.layers.utils")
    ut.get_norm_layer = _get_norm_layer
    ut.get_act_layer = _get_act_layer
    sys.modules["monai.networks.layers.utils"] = ut

    tc = types.ModuleType("monai.utils.type_conversion")
    tc.convert_to_tensor = torch.as_tensor
    sys.modules["monai.utils.type_conversion"] = tc

    tr = types.ModuleType("monai.transforms")
    tr.Transform = object
    sys.modules["monai.transforms"] = tr

    from . import cox as _cox
    px = types.ModuleType("pycox.models.loss")
    px.CoxPHLoss = _cox.CoxPHLoss
    sys.modules["pycox.models.loss"] = px

    from . import cindex as _ci
    ll = types.ModuleType("lifelines.utils")
    ll.concordance_index = _ci.concordance_index
    sys.modules["lifelines.utils"] = ll


def load_reference():
    """Returns a namespace with the reference's own classes/functions (unchanged source files)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    # the reference's `data/__init__.py` imports SimpleITK-dependent code; only constants are needed
    data_pkg = types.ModuleType("data")
    data_pkg.__path__ = [os.path.join(REFERENCE_ROOT, "data")]
    sys.modules.setdefault("data", data_pkg)
    du = types.ModuleType("data.utils")
    du._stratifiedSplit = None
    sys.modules.setdefault("data.utils", du)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.densenet = importlib.import_module("models.densenet")
    ns.mlp = importlib.import_module("models.mlp")
    ns.utils = importlib.import_module("utils.utils")
    ns.multimodal = importlib.import_module("models.multimodal")
    ns.losses = importlib.import_module("losses.losses")
    ns.blender = importlib.import_module("losses.GradientBlender")
    return ns
