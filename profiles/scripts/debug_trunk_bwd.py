import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mmnn_sts_b200.models.densenet import DenseNet121
from oracle import model as om, synth
PFX = "image_model.model."
cin, spatial, batch = (2, (64, 64, 64), 8) if "--mid" in sys.argv else (1, (64, 64, 32), 4)
cfg = (6, 12, 24, 16)
for a in sys.argv:
    if a.startswith("--cfg="):
        cfg = tuple(int(v) for v in a[6:].split(","))
if "--small" in sys.argv:
    cin, spatial, batch = 2, (32, 32, 32), 4
sd = synth.make_state_dict(42, in_channels=cin, block_config=cfg)
image, _, _, _ = synth.make_batch(1, batch, cin, spatial)
g = torch.Generator().manual_seed(1)
gw = torch.randn(batch, 12, generator=g)
p = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
f_ref = om.densenet_features(p, om.densenet_backbone(p, image, True, None, PFX, block_config=cfg), None, PFX)
(f_ref * gw).sum().backward()
m = DenseNet121(spatial_dims=3, in_channels=cin, out_channels=2, feature_channels=12, dropout_prob=0.0, block_config=cfg)
m.load_state_dict({k[len(PFX):]: v for k, v in sd.items() if k.startswith(PFX)})
m = m.cuda().train()
f = m.features(m.backbone(image.cuda()))
(f * gw.cuda()).sum().backward()
torch.cuda.synchronize()
print("features err", float((f.detach().cpu() - f_ref).norm() / f_ref.norm()))
for k, q in m.named_parameters():
    ref = p[PFX + k].grad
    if ref is None:
        continue
    got = q.grad.cpu().double(); ref = ref.double()
    e = float((got - ref).norm() / (ref.norm() + 1e-30))
    flag = "  <<<<" if not (e < 0.2) else ""
    print(f"{e:10.3e}  |ref|={float(ref.norm()):9.3e} |got|={float(got.norm()):9.3e}  {k}{flag}")
