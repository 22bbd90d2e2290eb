"""GPU form of the reference's deterministic image transforms:

    val_transforms = Compose([EnsureChannelFirst(0), Normalize(IMAGE_DATA_MEAN, IMAGE_DATA_STDDEV), ScaleIntensity(),
                              Resize(spatial_size=SPATIAL_SIZE), ToTensor()])          /root/reference/main.py:86-92

`Normalize` is the reference's own class (/root/reference/utils/utils.py:346-355); `ScaleIntensity` and `Resize` are
MONAI 1.2 transforms with their default arguments (min-max scaling to [0, 1] over the whole image; "area"
interpolation = adaptive average pooling).  One fused C-ABI call (`mmnn_preprocess_volumes`, two HBM-bound passes)
processes a whole batch of raw volumes already on the device.  No CPU path: CPU tensors raise.
"""
import torch

from .. import _lib as L

IMAGE_DATA_MEAN = 286.90859071507913      # /root/reference/data/constants.py:91
IMAGE_DATA_STDDEV = 581.7816096485366     # /root/reference/data/constants.py:92
SPATIAL_SIZE = (64, 64, 64)               # /root/reference/main.py:60


class ValTransformsGPU:
    """callable(raw) -> float32 [B, C, *spatial_size] in [0, 1];  raw: CUDA tensor [B, C, X, Y, Z] (or [C, X, Y, Z])."""

    def __init__(self, mean=IMAGE_DATA_MEAN, std=IMAGE_DATA_STDDEV, spatial_size=SPATIAL_SIZE):
        self.mean, self.std, self.spatial_size = float(mean), float(std), tuple(int(v) for v in spatial_size)
        self._scratch = {}

    def __call__(self, raw):
        if not raw.is_cuda:
            raise L.MMNNLibraryError("mmnn_sts_b200 has no CPU path: move the raw volumes to a CUDA device")
        single = raw.dim() == 4
        x = (raw[None] if single else raw).contiguous().float()
        B, C, X, Y, Z = x.shape
        ox, oy, oz = self.spatial_size
        out = torch.empty((B, C, ox, oy, oz), dtype=torch.float32, device=x.device)
        key = (x.device.index, B)
        if key not in self._scratch:
            self._scratch[key] = torch.empty(2 * B, dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            rc = L.lib().mmnn_preprocess_volumes(x.data_ptr(), out.data_ptr(), self._scratch[key].data_ptr(), B, C, X, Y, Z,
                                                 ox, oy, oz, self.mean, self.std, torch.cuda.current_stream().cuda_stream)
        L.check(rc, "mmnn_preprocess_volumes")
        return out[0] if single else out
