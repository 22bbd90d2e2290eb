"""Stage-by-stage comparison of the trunk's intermediate buffers with the oracle (debug aid, run on the GPU box)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L
from mmnn_sts_b200.models.densenet import DenseNet121
from oracle import model as om, synth

PFX = "image_model.model."
cin, spatial, batch = (1, (64, 64, 32), 4) if "--cfg1" in sys.argv else (2, (32, 32, 32), 4)
if "--mid" in sys.argv:
    cin, spatial, batch = 2, (64, 64, 64), 8
training = "--eval" not in sys.argv
cfg = (6, 12, 24, 16)
for a in sys.argv:
    if a.startswith("--cfg="):
        cfg = tuple(int(v) for v in a[6:].split(","))
from oracle import emul
sd = synth.make_state_dict(42, in_channels=cin, block_config=cfg)
image, _, _, _ = synth.make_batch(1, batch, cin, spatial)
coll = {}
p = {k: v.clone() for k, v in sd.items()}
fn = emul.backbone_bf16 if "--emul" in sys.argv else om.densenet_backbone
with torch.no_grad():
    y_ref = fn(p, image, training, None, PFX, collect=coll, block_config=cfg)
m = DenseNet121(spatial_dims=3, in_channels=cin, out_channels=2, feature_channels=12, dropout_prob=0.0, block_config=cfg)
m.load_state_dict({k[len(PFX):]: v for k, v in sd.items() if k.startswith(PFX)})
m = m.cuda().train(training)
bb = m.backbone
with torch.no_grad():
    y = bb(image.cuda())
torch.cuda.synchronize()
ws = list(bb._workspaces.values())[0][0].tensor
offs = (C.c_longlong * 32)(); dims = (C.c_longlong * 32)()
L.lib().mmnn_encoder_debug_offsets(bb._plan, batch, *spatial, offs, dims)
nb = len(cfg)


def view(off, rows, cols):
    return ws[off:off + rows * cols * 2].view(torch.bfloat16).view(rows, cols).float().cpu()


def cmp(name, got, ref):
    err = (got - ref).abs().max() / ref.abs().max()
    l2 = (got - ref).norm() / ref.norm()
    print(f"{name:12s} rel-max err {float(err):.3e}  rel-L2 {float(l2):.3e}  ref absmax {float(ref.abs().max()):.3f}")


M0 = dims[0]
cmp("conv0", view(offs[1], M0, 64), coll["conv0"].permute(0, 2, 3, 4, 1).reshape(M0, 64))
for b in range(nb):
    M, ctot, c0 = dims[4 + 4 * b], dims[5 + 4 * b], dims[6 + 4 * b]
    got = view(offs[3 + b], M, ctot)
    ref = coll[f"block{b + 1}"].permute(0, 2, 3, 4, 1).reshape(M, ctot)
    cmp(f"block{b+1}[:c0]", got[:, :c0], ref[:, :c0])
    for l in range((ctot - c0) // 32):
        e = (got[:, c0 + 32 * l:c0 + 32 * l + 32] - ref[:, c0 + 32 * l:c0 + 32 * l + 32]).abs().max() / ref[:, c0 + 32 * l:c0 + 32 * l + 32].abs().max()
        e2 = (got[:, c0 + 32 * l:c0 + 32 * l + 32] - ref[:, c0 + 32 * l:c0 + 32 * l + 32]).norm() / ref[:, c0 + 32 * l:c0 + 32 * l + 32].norm()
        print(f"   L{l + 1}: {float(e):.1e}/{float(e2):.1e}", end="")
    print()
cmp("norm5", y.cpu(), y_ref)

f_ref = om.densenet_features(p, y_ref, None, PFX)
with torch.no_grad():
    f = m.features(y)
print("features:", float((f.cpu() - f_ref).abs().max() / f_ref.abs().max()), float((f.cpu() - f_ref).norm() / f_ref.norm()))
print(f_ref[0], f[0].cpu())

# ---- statistics arena vs statistics recomputed from the stored buffers
FC = dims[4 + 4 * nb]
fst = ws[offs[3 + 3 * nb]:offs[3 + 3 * nb] + FC * 16].view(torch.float64).view(2, FC).cpu()
for b in range(nb):
    M, ctot, c0, foff = dims[4 + 4 * b], dims[5 + 4 * b], dims[6 + 4 * b], dims[7 + 4 * b]
    got = view(offs[3 + b], M, ctot).double()
    s1, s2 = got.sum(0), (got ** 2).sum(0)
    print(f"block{b+1} stats: sum rel err {float((fst[0, foff:foff+ctot]-s1).norm()/s1.norm()):.2e}  sumsq rel err {float((fst[1, foff:foff+ctot]-s2).norm()/s2.norm()):.2e}")
    # bottleneck stats of layer 1: follows the block channels in the arena
    bott = view(offs[3 + nb + b], M, 128).double()
    o = foff + ctot
    print(f"   bott L1 stats: sum {float((fst[0, o:o+128]-bott.sum(0)).norm()/bott.sum(0).norm()):.2e} sumsq {float((fst[1, o:o+128]-(bott**2).sum(0)).norm()/(bott**2).sum(0).norm()):.2e}")
m0 = view(offs[1], M0, 64).double()
print("stem stats:", float((fst[0, :64] - m0.sum(0)).norm() / m0.sum(0).norm()), float((fst[1, :64] - (m0 ** 2).sum(0)).norm() / (m0 ** 2).sum(0).norm()))

# ---- recompute block-1 layer-1 from the STORED inputs in float64 on the GPU and compare with the stored outputs
import torch.nn.functional as F
bf = lambda t: t.to(torch.bfloat16).double()
M, ctot, c0 = dims[4], dims[5], dims[6]
Bn, D1 = batch, round((M // batch) ** (1 / 3))
x = ws[offs[3]:offs[3] + M * ctot * 2].view(torch.bfloat16).view(M, ctot)[:, :c0].double()
q = PFX + "backbone.denseblock1.denselayer1.layers."
dev = lambda k: sd[k].cuda().double()
def bn(t, pre):
    mean = t.mean(0); var = t.var(0, unbiased=False)
    return (t - mean) / torch.sqrt(var + 1e-5) * dev(pre + ".weight") + dev(pre + ".bias")
a1 = bf(F.relu(bn(x, q + "norm1")))
bott_ref = bf(a1 @ bf(dev(q + "conv1.weight").view(128, c0)).t())
bott = ws[offs[3 + nb]:offs[3 + nb] + M * 128 * 2].view(torch.bfloat16).view(M, 128).double()
print("layer1 bott vs f64 recompute: rel", float((bott - bott_ref).norm() / bott_ref.norm()), "exact", float((bott == bott_ref).double().mean()))
a2 = bf(F.relu(bn(bott, q + "norm2")))
a25 = a2.view(Bn, D1, D1, D1, 128).permute(0, 4, 1, 2, 3)
y_ref2 = bf(F.conv3d(a25, bf(dev(q + "conv2.weight")), padding=1).permute(0, 2, 3, 4, 1).reshape(M, 32))
y_got = ws[offs[3]:offs[3] + M * ctot * 2].view(torch.bfloat16).view(M, ctot)[:, c0:c0 + 32].double()
print("layer1 new slice vs f64 recompute from stored bott: rel", float((y_got - y_ref2).norm() / y_ref2.norm()), "exact", float((y_got == y_ref2).double().mean()))
