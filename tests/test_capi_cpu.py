"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/mmnn_b200.h declares, the
ctypes struct mirrors have the library's sizes, the Python mirrors keep the reference's constructor signatures and
state_dict layout, and the product refuses to run without CUDA (no CPU fallback).  No compute call is made."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from mmnn_sts_b200 import _lib
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mmnn_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mmnn_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmnn_b200.h but not exported"


def test_struct_sizes_match(lib):
    from mmnn_sts_b200 import _lib as L
    assert lib.mmnn_sizeof_rows_params() == ctypes.sizeof(L.RowsParams)
    assert lib.mmnn_sizeof_wgrad_params() == ctypes.sizeof(L.WgradParams)
    assert lib.mmnn_sizeof_pack_desc() == ctypes.sizeof(L.PackDesc)
    assert lib.mmnn_sizeof_mlp_args() == ctypes.sizeof(L.MlpArgs)
    assert lib.mmnn_sizeof_cox_args() == ctypes.sizeof(L.CoxArgs)
    assert lib.mmnn_sizeof_cindex_args() == ctypes.sizeof(L.CindexArgs)
    assert lib.mmnn_sizeof_aug_spatial() == ctypes.sizeof(L.AugSpatial)
    assert lib.mmnn_sizeof_aug_intensity() == ctypes.sizeof(L.AugIntensity)


def test_encoder_plan_tables(lib):
    cfg = (ctypes.c_int * 4)(6, 12, 24, 16)
    h = lib.mmnn_encoder_create(2, cfg, 4, 64, 32, 4)
    assert h
    h = ctypes.c_void_p(h)
    assert lib.mmnn_encoder_num_params(h) == 364 - 2  # backbone: 120 conv + 2*121 BN tensors
    assert lib.mmnn_encoder_num_buffers(h) == 363
    assert lib.mmnn_encoder_num_layers(h) == 58
    assert lib.mmnn_encoder_out_channels(h) == 1024
    dims = (ctypes.c_int * 3)()
    assert lib.mmnn_encoder_out_dims(h, 16, 128, 128, 64, dims) == 0 and list(dims) == [4, 4, 2]
    assert lib.mmnn_encoder_workspace_bytes(h, 16, 128, 128, 64) > (1 << 30)
    assert lib.mmnn_encoder_workspace_bytes(h, 1, 8, 8, 8) == -1          # too small for five halvings
    lib.mmnn_encoder_destroy(h)
    assert not lib.mmnn_encoder_create(3, cfg, 4, 64, 32, 4)              # unsupported -> NULL, never a fallback


def test_state_dict_layout_matches_reference_spec():
    from mmnn_sts_b200.models.densenet import DenseNet121
    from mmnn_sts_b200.models.multimodal import MultiModalModel
    from oracle import synth
    m = MultiModalModel(DenseNet121(spatial_dims=3, in_channels=2, out_channels=2, feature_channels=12, dropout_prob=0.2), ["x"] * 20, 2, 12, blend=True)
    sd = m.state_dict()
    spec = synth.state_dict_spec(in_channels=2)          # key order / shapes validated against the unchanged reference
    assert len(sd) == 779 and list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sum(p.numel() for p in m.parameters()) == 11278786    # 2-channel model
    m.load_state_dict(synth.make_state_dict(1))


def test_no_cpu_fallback():
    from mmnn_sts_b200 import _lib as L
    from mmnn_sts_b200.losses.losses import CoxPH
    from mmnn_sts_b200.models.densenet import DenseNet121
    if torch.cuda.is_available():
        pytest.skip("CPU-refusal test runs on the GPU-less builder")
    m = DenseNet121(spatial_dims=3, in_channels=1, out_channels=2, feature_channels=12)
    with pytest.raises(L.MMNNLibraryError):
        m.backbone(torch.rand(2, 1, 32, 32, 32))
    with pytest.raises(L.MMNNLibraryError):
        CoxPH(torch.randn(4), torch.ones(4), torch.arange(4))
    from mmnn_sts_b200.optim import SGD
    p = torch.nn.Parameter(torch.ones(4)); p.grad = torch.ones(4)
    with pytest.raises(L.MMNNLibraryError):
        SGD([p], 0.1, momentum=0.9, nesterov=True).step()


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


@pytest.mark.parametrize("cin", [1, 2])
def test_seeded_initialisation_is_bit_identical_to_the_reference(cin):
    """under torch.manual_seed(42) every one of the 779 state_dict tensors of this package's
    MultiModalModel equals, bit for bit, the tensor the UNCHANGED reference constructs (same module construction order, same
    initialisation law: /root/reference/models/densenet.py:258-265).  Digests: tests/golden/make_init_golden.py."""
    import hashlib
    import json
    import torch
    from mmnn_sts_b200.models.densenet import DenseNet121
    from mmnn_sts_b200.models.multimodal import MultiModalModel
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "init_seed42.json")))[str(cin)]
    torch.manual_seed(42)
    m = MultiModalModel(DenseNet121(spatial_dims=3, in_channels=cin, out_channels=2, feature_channels=12, dropout_prob=0.2),
                        ["x"] * 20, 2, 12, blend=True)
    got = {k: hashlib.sha1(v.detach().cpu().contiguous().numpy().tobytes()).hexdigest() for k, v in m.state_dict().items()}
    assert list(got) == list(want)
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, bad[:5]
