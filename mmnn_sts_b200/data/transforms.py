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


class TrainTransformsGPU:
    """The reference's `train_transforms` (/root/reference/main.py:64-85) on the GPU: the deterministic chain of
    `ValTransformsGPU` with MONAI's random transforms in between, each with its default arguments as written there:

        RandRotate(range_x=15, prob=0.5, keep_size=True), RandAxisFlip(prob=0.5), RandZoom(0.9, 1.1, prob=0.5, keep_size=True),
        Resize, RandShiftIntensity(0.1, prob=0.3), RandAdjustContrast(prob=0.3), RandGaussianSmooth(prob=0.2),
        RandGaussianSharpen(prob=0.2), RandHistogramShift(prob=0.3), RandGaussianNoise(prob=0.3, mean=0, std=0.05)

    The parameters of every transform are drawn HERE (numpy RandomState, the distributions of MONAI 1.2's `randomize`
    methods) and handed to two C-ABI calls (`mmnn_augment_resample`: one affine resampling fused with the area resize;
    `mmnn_augment_intensity`).  Parity unpinned (csrc/augment.cu): the reference never runs this branch as shipped, and the
    three spatial resamplings are composed into one tri-linear pass.  `draw()` / `apply()` are separate so that a test (or a
    caller that wants reproducibility across devices) can fix the parameters."""

    def __init__(self, mean=IMAGE_DATA_MEAN, std=IMAGE_DATA_STDDEV, spatial_size=SPATIAL_SIZE, seed=None):
        import numpy as np
        self.mean, self.std, self.spatial_size = float(mean), float(std), tuple(int(v) for v in spatial_size)
        self.R = np.random.RandomState(seed)

    def draw(self, B):
        """Per-sample parameters of one batch (plain Python / numpy values)."""
        import numpy as np
        R = self.R
        out = []
        for _ in range(B):
            q = {"rotate": None, "flip_axis": None, "zoom": None, "shift": 0.0, "gamma": 0.0, "smooth": None, "sharpen": None,
                 "hist": None, "noise_std": 0.0, "seed": int(R.randint(0, 2 ** 31 - 1))}
            if R.rand() < 0.5:
                q["rotate"] = float(R.uniform(-15.0, 15.0))             # radians about the first spatial axis (range_x=15)
            if R.rand() < 0.5:
                q["flip_axis"] = int(R.randint(3))
            if R.rand() < 0.5:
                q["zoom"] = float(R.uniform(0.9, 1.1))
            if R.rand() < 0.3:
                q["shift"] = float(R.uniform(-0.1, 0.1))
            if R.rand() < 0.3:
                q["gamma"] = float(R.uniform(0.5, 4.5))
            if R.rand() < 0.2:
                q["smooth"] = [float(R.uniform(0.25, 1.5)) for _ in range(3)]
            if R.rand() < 0.2:
                s1 = [float(R.uniform(0.5, 1.0)) for _ in range(3)]
                q["sharpen"] = (s1, [float(R.uniform(0.5, v)) for v in s1], float(R.uniform(10.0, 30.0)))
            if R.rand() < 0.3:
                ref = np.linspace(0.0, 1.0, L.AUG_HIST_POINTS)
                flt = ref.copy()
                for i in range(1, L.AUG_HIST_POINTS - 1):
                    flt[i] = R.uniform(flt[i - 1], flt[i + 1])
                q["hist"] = (ref.tolist(), flt.tolist())
            if R.rand() < 0.3:
                q["noise_std"] = float(R.uniform(0.0, 0.05))
            out.append(q)
        return out

    @staticmethod
    def affine(q):
        """3x3 matrix A of `source = A (grid - centre) + centre`: the image passes rotate -> flip -> zoom, so the output voxel g
        reads the source at R F g / zoom."""
        import numpy as np
        A = np.eye(3)
        if q["rotate"] is not None:
            c, s = np.cos(q["rotate"]), np.sin(q["rotate"])
            A = A @ np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
        if q["flip_axis"] is not None:
            F = np.eye(3)
            F[q["flip_axis"], q["flip_axis"]] = -1.0
            A = A @ F
        if q["zoom"] is not None:
            A = A / q["zoom"]
        return A

    def apply(self, raw, params):
        import ctypes as C
        if not raw.is_cuda:
            raise L.MMNNLibraryError("mmnn_sts_b200 has no CPU path: move the raw volumes to a CUDA device")
        single = raw.dim() == 4
        x = (raw[None] if single else raw).contiguous().float()
        B, Cc, X, Y, Z = x.shape
        assert len(params) == B
        ox, oy, oz = self.spatial_size
        dev = x.device
        sp = (L.AugSpatial * B)()
        it = (L.AugIntensity * B)()
        smooth = torch.zeros(B, 3)
        sharpen = torch.zeros(7 * B)
        for b, q in enumerate(params):
            A = self.affine(q)
            for i in range(9):
                sp[b].a[i] = float(A[i // 3, i % 3])
            it[b].shift, it[b].gamma, it[b].noise_std, it[b].seed = q["shift"], q["gamma"], q["noise_std"], q["seed"]
            it[b].hist_on = 0
            if q["hist"] is not None:
                it[b].hist_on = 1
                for i in range(L.AUG_HIST_POINTS):
                    it[b].hist_ref[i], it[b].hist_flt[i] = q["hist"][0][i], q["hist"][1][i]
            if q["smooth"] is not None:
                smooth[b] = torch.tensor(q["smooth"])
            if q["sharpen"] is not None:
                sharpen[3 * b:3 * b + 3] = torch.tensor(q["sharpen"][0])
                sharpen[3 * B + 3 * b:3 * B + 3 * b + 3] = torch.tensor(q["sharpen"][1])
                sharpen[6 * B + b] = q["sharpen"][2]
        sp_d = torch.frombuffer(bytearray(bytes(sp)), dtype=torch.uint8).to(dev)
        it_d = torch.frombuffer(bytearray(bytes(it)), dtype=torch.uint8).to(dev)
        smooth_d, sharpen_d = smooth.to(dev), sharpen.to(dev)
        out = torch.empty((B, Cc, ox, oy, oz), dtype=torch.float32, device=dev)
        tmp = torch.empty((3,) + tuple(out.shape), dtype=torch.float32, device=dev)
        scratch = torch.empty(2 * B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            L.check(L.lib().mmnn_augment_resample(x.data_ptr(), out.data_ptr(), scratch.data_ptr(), sp_d.data_ptr(), B, Cc, X, Y, Z,
                                                  ox, oy, oz, self.mean, self.std, st), "mmnn_augment_resample")
            L.check(L.lib().mmnn_augment_intensity(out.data_ptr(), tmp.data_ptr(), scratch.data_ptr(), it_d.data_ptr(),
                                                   smooth_d.data_ptr(), sharpen_d.data_ptr(), B, Cc, ox, oy, oz,
                                                   int(any(q["smooth"] is not None for q in params)),
                                                   int(any(q["sharpen"] is not None for q in params)), st), "mmnn_augment_intensity")
        return out[0] if single else out

    def __call__(self, raw):
        B = 1 if raw.dim() == 4 else raw.shape[0]
        return self.apply(raw, self.draw(B))
