"""CPU restatement of the RANDOM transforms of the reference's `train_transforms` (/root/reference/main.py:64-85).
TEST INFRASTRUCTURE.  PARITY UNPINNED: every transform below is a class of monai-weekly==1.2.dev2313 (requirements.txt:2; NOT
vendored under /root/reference, absent here), and the reference as shipped never executes this branch (the validation wrapper
overwrites the shared dataset's transforms: main.py:355-356, data/ImageDatasets.py:316-318).  Each function restates the published
algorithm of one transform on EXPLICIT parameters (numpy, float64 where it is cheap), in the order of the reference's Compose:

  affine_matrix       RandRotate(range_x) -> RandAxisFlip -> RandZoom(keep_size) composed: output voxel g (about the volume centre)
                      reads the source at R F g / zoom (the CUDA path resamples ONCE with this map; MONAI resamples three times)
  resample            tri-linear sampling with edge clamping on the keep_size grid, then Resize(mode="area") = adaptive average
  shift_intensity     ShiftIntensity: img + offset
  adjust_contrast     AdjustContrast: ((img - min) / (range + 1e-7)) ** gamma * range + min
  gaussian_1d         monai.networks.layers.gaussian_1d(sigma, truncated=4, approx="erf")
  gaussian_smooth     GaussianSmooth: separable convolution with that kernel, zero padding
  gaussian_sharpen    GaussianSharpen: b1 = G(s1) img;  b1 + alpha * (b1 - G(s2) b1)
  histogram_shift     RandHistogramShift: np.interp(img, reference control points, floating control points) over [min, max]
  gaussian_noise      RandGaussianNoise: img + N(mean, std)   (only the formula; the generator is the caller's)
"""
import numpy as np


def affine_matrix(rotate=None, flip_axis=None, zoom=None):
    A = np.eye(3)
    if rotate is not None:
        c, s = np.cos(rotate), np.sin(rotate)
        A = A @ np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
    if flip_axis is not None:
        F = np.eye(3)
        F[flip_axis, flip_axis] = -1.0
        A = A @ F
    if zoom is not None:
        A = A / zoom
    return A


def _trilinear(vol, x, y, z):
    X, Y, Z = vol.shape
    x = np.clip(x, 0, X - 1); y = np.clip(y, 0, Y - 1); z = np.clip(z, 0, Z - 1)
    x0 = np.floor(x).astype(int); y0 = np.floor(y).astype(int); z0 = np.floor(z).astype(int)
    x1 = np.minimum(x0 + 1, X - 1); y1 = np.minimum(y0 + 1, Y - 1); z1 = np.minimum(z0 + 1, Z - 1)
    fx, fy, fz = x - x0, y - y0, z - z0
    c00 = vol[x0, y0, z0] * (1 - fz) + vol[x0, y0, z1] * fz
    c01 = vol[x0, y1, z0] * (1 - fz) + vol[x0, y1, z1] * fz
    c10 = vol[x1, y0, z0] * (1 - fz) + vol[x1, y0, z1] * fz
    c11 = vol[x1, y1, z0] * (1 - fz) + vol[x1, y1, z1] * fz
    return (c00 * (1 - fy) + c01 * fy) * (1 - fx) + (c10 * (1 - fy) + c11 * fy) * fx


def resample(image, A, out_size):
    """image [C][X][Y][Z] (already normalised / scaled), A 3x3: -> [C][ox][oy][oz]."""
    C, X, Y, Z = image.shape
    ctr = np.array([(X - 1) / 2, (Y - 1) / 2, (Z - 1) / 2])
    g = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij"), -1).astype(np.float64) - ctr
    src = g @ np.asarray(A, dtype=np.float64).T + ctr
    warped = np.stack([_trilinear(image[c].astype(np.float64), src[..., 0], src[..., 1], src[..., 2]) for c in range(C)])
    ox, oy, oz = out_size
    out = np.empty((C, ox, oy, oz))
    for i in range(ox):
        x0, x1 = (i * X) // ox, -((-(i + 1) * X) // ox)
        for j in range(oy):
            y0, y1 = (j * Y) // oy, -((-(j + 1) * Y) // oy)
            for k in range(oz):
                z0, z1 = (k * Z) // oz, -((-(k + 1) * Z) // oz)
                out[:, i, j, k] = warped[:, x0:x1, y0:y1, z0:z1].mean(axis=(1, 2, 3))
    return out


def shift_intensity(img, offset):
    return img + offset


def adjust_contrast(img, gamma):
    mn = img.min()
    rg = img.max() - mn
    return ((img - mn) / (rg + 1e-7)) ** gamma * rg + mn


def gaussian_1d(sigma, truncated=4.0):
    from math import erf
    tail = int(max(sigma * truncated, 0.5) + 0.5)
    x = np.arange(-tail, tail + 1, dtype=np.float64)
    t = 0.70710678 / abs(sigma)
    k = 0.5 * (np.array([erf(v) for v in t * (x + 0.5)]) - np.array([erf(v) for v in t * (x - 0.5)]))
    return np.clip(k, 0, None)


def gaussian_smooth(img, sigmas):
    """img [C][x][y][z]; one sigma per spatial axis; zero padding ("same" convolution)."""
    out = img.astype(np.float64)
    for ax, s in enumerate(sigmas):
        k = gaussian_1d(s)
        out = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1 + ax, out)
    return out


def gaussian_sharpen(img, sigma1, sigma2, alpha):
    b1 = gaussian_smooth(img, sigma1)
    b2 = gaussian_smooth(b1, sigma2)
    return b1 + alpha * (b1 - b2)


def histogram_shift(img, reference_fractions, floating_fractions):
    mn, mx = img.min(), img.max()
    xp = np.asarray(reference_fractions) * (mx - mn) + mn
    fp = np.asarray(floating_fractions) * (mx - mn) + mn
    return np.interp(img, xp, fp)


def gaussian_noise(img, noise):
    return img + noise
