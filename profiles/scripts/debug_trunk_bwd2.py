import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from mmnn_sts_b200 import _lib as L
from mmnn_sts_b200.models.densenet import DenseNet121
from oracle import model as om, synth
PFX = "image_model.model."
cfg = (1,)
cin, spatial, batch = 2, (32, 32, 32), 4
sd = synth.make_state_dict(42, in_channels=cin, block_config=cfg)
image, _, _, _ = synth.make_batch(1, batch, cin, spatial)
g = torch.Generator().manual_seed(1)
gw = torch.randn(batch, 12, generator=g)
p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
coll = {}
y_ref = om.densenet_backbone(p, image, True, None, PFX, block_config=cfg, collect=coll)
for v in coll.values():
    v.retain_grad()
y_ref.retain_grad()
f_ref = om.densenet_features(p, y_ref, None, PFX)
(f_ref * gw).sum().backward()
m = DenseNet121(spatial_dims=3, in_channels=cin, out_channels=2, feature_channels=12, dropout_prob=0.0, block_config=cfg)
m.load_state_dict({k[len(PFX):]: v for k, v in sd.items() if k.startswith(PFX)})
m = m.cuda().train()
bb = m.backbone
yb = bb(image.cuda())
yb.retain_grad()
f = m.features(yb)
(f * gw.cuda()).sum().backward()
torch.cuda.synchronize()
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
print("dy (grad wrt norm5 out):", rel(yb.grad.cpu(), y_ref.grad))
ws = list(bb._workspaces.values())[0][0].tensor
offs = (C.c_longlong * 32)(); dims = (C.c_longlong * 32)()
L.lib().mmnn_encoder_debug_offsets(bb._plan, batch, *spatial, offs, dims)
nb = 1
M, ctot, c0 = dims[4], dims[5], dims[6]
dbuf = ws[offs[3 + 2 * nb]:offs[3 + 2 * nb] + M * ctot * 4].view(torch.float32).view(M, ctot).cpu()
ref = coll["block1"].grad.permute(0, 2, 3, 4, 1).reshape(M, ctot)
print("dbuf new-slice:", rel(dbuf[:, c0:], ref[:, c0:]), " dbuf first c0 (accumulated):", rel(dbuf[:, :c0], ref[:, :c0]))
print("pool0 grad:", rel(dbuf[:, :c0], coll["pool0"].grad.permute(0, 2, 3, 4, 1).reshape(M, c0)))
print(dbuf[:2, c0:c0 + 6], ref[:2, c0:c0 + 6])
