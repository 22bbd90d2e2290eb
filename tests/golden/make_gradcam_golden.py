"""Generates tests/golden/gradcam.npz from the UNCHANGED reference MultiModalGradCAM (/root/reference/utils/utils.py:
253-344) + reference model classes under oracle/shim.py, fp32 on CPU: T1-only model, one 1x64x64x64 volume, eval mode.
Stores the weights' seed, the input seed, the logits, and the attention maps subsampled every 4th voxel per axis."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shim, synth  # noqa: E402

ref = shim.load_reference()
torch.manual_seed(0)
image_model = ref.densenet.DenseNet121(spatial_dims=3, in_channels=1, out_channels=2, feature_channels=12, dropout_prob=0.0)
model = ref.multimodal.MultiModalModel(image_model, ["x"] * 20, 2, 12, blend=False)
sd = synth.make_state_dict(21, in_channels=1)
model.load_state_dict(sd)
model.eval()
image, clinical, _, _ = synth.make_batch(22, 1, 1, (64, 64, 64))
cam = ref.utils.MultiModalGradCAM(model)
outputs, maps = cam({"image": image, "clinical": clinical})
maps = torch.stack([m.detach() for m in maps])
print("logits", outputs.detach().numpy(), "maps", tuple(maps.shape), float(maps.min()), float(maps.max()))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gradcam.npz"),
                    state_seed=21, batch_seed=22, logits=outputs.detach().numpy(), maps_sub4=maps[:, ::4, ::4, ::4].numpy(),
                    maps_mean=maps.mean(dim=(1, 2, 3)).numpy())
