// GPU form of the reference's TRAINING transform chain (/root/reference/main.py:64-85):
//   Normalize -> ScaleIntensity -> RandRotate(range_x=15, p=.5, keep_size) -> RandAxisFlip(p=.5) -> RandZoom(.9..1.1, p=.5, keep_size)
//   -> Resize(spatial_size) -> RandShiftIntensity(.1, p=.3) -> RandAdjustContrast(p=.3) -> RandGaussianSmooth(p=.2)
//   -> RandGaussianSharpen(p=.2) -> RandHistogramShift(p=.3) -> RandGaussianNoise(p=.3, std=.05)
// The random transforms are MONAI 1.2 classes (not vendored in /root/reference) and -- as shipped -- the reference never runs them
// (the validation wrapper overwrites the shared dataset's transforms), so there is no reference output to pin
// against: PARITY UNPINNED.  What is built here is each transform's published definition with EXPLICIT parameters (the host
// draws them, mmnn_sts_b200/data/transforms.py), checked against a plain torch restatement on the same parameters (tests/).
//
// Passes (all HBM-bound, one thread per output voxel):
//   1. min / max of the raw volume (preprocess.cu)
//   2. augment_resample_kernel: the three spatial transforms are ONE affine map of the keep_size grid (about the volume centre,
//      voxel units, composed on the host: source = A * grid + t); every voxel of the adaptive-average window of the final Resize is
//      sampled tri-linearly from the RAW volume with edge clamping, averaged, and the Normalize / ScaleIntensity affine map is applied
//      to the average (it commutes with interpolation and averaging).  One pass instead of MONAI's three resamplings + one resize.
//   3. intensity passes over the small output volume: shift + gamma contrast, separable Gaussian blur (smooth, sharpen),
//      histogram shift (monotone piecewise-linear remap through control points) + Gaussian noise (counter-based hash, Box-Muller).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "prof.h"

namespace mmnn {

__device__ __forceinline__ float aug_dec_f32(uint32_t e) {     // inverse of preprocess.cu's order-preserving encoding
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}
__device__ __forceinline__ uint32_t aug_enc_f32(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct AugSpatial {      // per sample: source = A * (grid - centre) + centre + t   (rows of A then t), identity when nothing fired
  float a[9];
  float t[3];
};

struct AugResampleParams {
  const float* src;      // [B][C][X][Y][Z] raw intensities
  float* dst;            // [B][C][ox][oy][oz]
  const uint32_t* ext;   // per sample: enc(max x), enc(max -x)   (preprocess.cu pass 1)
  const AugSpatial* sp;  // [B] (device)
  int B, C, X, Y, Z, ox, oy, oz;
  float mean, std;
};

__device__ __forceinline__ float aug_trilinear(const float* __restrict__ s, int X, int Y, int Z, float x, float y, float z) {
  x = fminf(fmaxf(x, 0.f), (float)(X - 1)); y = fminf(fmaxf(y, 0.f), (float)(Y - 1)); z = fminf(fmaxf(z, 0.f), (float)(Z - 1));
  const int x0 = (int)x, y0 = (int)y, z0 = (int)z;
  const int x1 = min(x0 + 1, X - 1), y1 = min(y0 + 1, Y - 1), z1 = min(z0 + 1, Z - 1);
  const float fx = x - (float)x0, fy = y - (float)y0, fz = z - (float)z0;
  const long long r00 = ((long long)x0 * Y + y0) * Z, r01 = ((long long)x0 * Y + y1) * Z;
  const long long r10 = ((long long)x1 * Y + y0) * Z, r11 = ((long long)x1 * Y + y1) * Z;
  const float c00 = __ldg(s + r00 + z0) * (1.f - fz) + __ldg(s + r00 + z1) * fz;
  const float c01 = __ldg(s + r01 + z0) * (1.f - fz) + __ldg(s + r01 + z1) * fz;
  const float c10 = __ldg(s + r10 + z0) * (1.f - fz) + __ldg(s + r10 + z1) * fz;
  const float c11 = __ldg(s + r11 + z0) * (1.f - fz) + __ldg(s + r11 + z1) * fz;
  const float c0 = c00 * (1.f - fy) + c01 * fy, c1 = c10 * (1.f - fy) + c11 * fy;
  return c0 * (1.f - fx) + c1 * fx;
}

__global__ void __launch_bounds__(256) augment_resample_kernel(const __grid_constant__ AugResampleParams p) {
  const long long total = (long long)p.B * p.C * p.ox * p.oy * p.oz;
  const float cx = 0.5f * (float)(p.X - 1), cy = 0.5f * (float)(p.Y - 1), cz = 0.5f * (float)(p.Z - 1);
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    long long t = idx;
    const int k = (int)(t % p.oz); t /= p.oz;
    const int j = (int)(t % p.oy); t /= p.oy;
    const int i = (int)(t % p.ox); t /= p.ox;
    const int c = (int)(t % p.C);
    const int b = (int)(t / p.C);
    const AugSpatial sp = p.sp[b];
    const float M = aug_dec_f32(p.ext[2 * b]), m = -aug_dec_f32(p.ext[2 * b + 1]);
    const float sub = __fmul_rn(p.mean, M), den = __fmul_rn(p.std, M);
    const float na = (m - sub) / den, nb = (M - sub) / den;
    const float nmin = fminf(na, nb), range = fmaxf(na, nb) - nmin;
    const int x0 = (int)(((long long)i * p.X) / p.ox), x1 = (int)((((long long)i + 1) * p.X + p.ox - 1) / p.ox);
    const int y0 = (int)(((long long)j * p.Y) / p.oy), y1 = (int)((((long long)j + 1) * p.Y + p.oy - 1) / p.oy);
    const int z0 = (int)(((long long)k * p.Z) / p.oz), z1 = (int)((((long long)k + 1) * p.Z + p.oz - 1) / p.oz);
    const float* s = p.src + ((long long)b * p.C + c) * p.X * p.Y * p.Z;
    float acc = 0.f;
    for (int x = x0; x < x1; ++x)
      for (int y = y0; y < y1; ++y)
        for (int z = z0; z < z1; ++z) {
          const float gx = (float)x - cx, gy = (float)y - cy, gz = (float)z - cz;
          const float sx = sp.a[0] * gx + sp.a[1] * gy + sp.a[2] * gz + cx + sp.t[0];
          const float sy = sp.a[3] * gx + sp.a[4] * gy + sp.a[5] * gz + cy + sp.t[1];
          const float sz = sp.a[6] * gx + sp.a[7] * gy + sp.a[8] * gz + cz + sp.t[2];
          acc += aug_trilinear(s, p.X, p.Y, p.Z, sx, sy, sz);
        }
    const float avg = acc / (float)((x1 - x0) * (y1 - y0) * (z1 - z0));
    const float n = (avg - sub) / den;
    p.dst[idx] = (range == 0.f) ? n * 0.f : (n - nmin) / range;
  }
}

// per-sample min / max of a small volume (one block per sample): ext2[b] = {enc(max), enc(max -x)}
__global__ void __launch_bounds__(256) aug_minmax_kernel(const float* __restrict__ v, long long per_image, uint32_t* __restrict__ ext2) {
  const int b = blockIdx.x;
  const float* s = v + (long long)b * per_image;
  float mx = -INFINITY, mn = INFINITY;
  for (long long i = threadIdx.x; i < per_image; i += 256) { const float x = s[i]; mx = fmaxf(mx, x); mn = fminf(mn, x); }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  __shared__ float smx[8], smn[8];
  if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, smx[w]); mn = fminf(mn, smn[w]); }
    ext2[2 * b] = aug_enc_f32(mx);
    ext2[2 * b + 1] = aug_enc_f32(-mn);
  }
}

constexpr int AUG_HIST_POINTS = 10;      // RandHistogramShift(num_control_points=10), MONAI default

struct AugIntensity {     // per sample; a transform that did not fire carries its neutral value
  float shift;            // RandShiftIntensity: img + shift                                   (0)
  float gamma;            // RandAdjustContrast: ((img - min) / (range + 1e-7)) ** gamma * range + min   (<= 0: off)
  float noise_std;        // RandGaussianNoise: img + N(0, noise_std)                           (0)
  int hist_on;            // RandHistogramShift: interp(img, reference control points, floating control points)
  float hist_ref[AUG_HIST_POINTS];    // fractions of the intensity range, increasing, [0] = 0, [last] = 1
  float hist_flt[AUG_HIST_POINTS];
  unsigned long long seed;
};

// phase 0: shift + gamma (extremes of the volume BEFORE this pass in ext2);  phase 1: histogram shift + noise (extremes again)
__global__ void __launch_bounds__(256) augment_intensity_kernel(float* __restrict__ v, long long per_image, int B,
                                                                 const AugIntensity* __restrict__ prm, const uint32_t* __restrict__ ext2,
                                                                 int phase) {
  const long long total = per_image * B;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int b = (int)(idx / per_image);
    const AugIntensity q = prm[b];
    const float mx = aug_dec_f32(ext2[2 * b]), mn = -aug_dec_f32(ext2[2 * b + 1]);
    float x = v[idx];
    if (phase == 0) {
      x += q.shift;
      if (q.gamma > 0.f) {
        const float lo = mn + q.shift, rg = mx - mn;
        x = powf((x - lo) / (rg + 1e-7f), q.gamma) * rg + lo;
      }
    } else {
      if (q.hist_on) {
        const float rg = mx - mn;
        const float u = rg > 0.f ? (x - mn) / rg : 0.f;      // position in the range; control points are fractions of it
        int seg = 0;
#pragma unroll
        for (int i = 1; i < AUG_HIST_POINTS - 1; ++i) seg += (u >= q.hist_ref[i]) ? 1 : 0;
        const float r0 = q.hist_ref[seg], r1 = q.hist_ref[seg + 1], f0 = q.hist_flt[seg], f1 = q.hist_flt[seg + 1];
        const float w = r1 > r0 ? (u - r0) / (r1 - r0) : 0.f;
        x = mn + rg * (f0 + (f1 - f0) * fminf(fmaxf(w, 0.f), 1.f));
      }
      if (q.noise_std > 0.f) {
        // counter-based generator: two 32-bit hashes of (seed, element index) -> Box-Muller
        unsigned long long h = q.seed ^ ((unsigned long long)idx * 0x9E3779B97F4A7C15ull);
        h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 27; h *= 0x94D049BB133111EBull; h ^= h >> 31;
        const float u1 = ((float)(uint32_t)(h & 0xffffffu) + 1.f) * (1.f / 16777217.f);
        const float u2 = (float)(uint32_t)((h >> 32) & 0xffffffu) * (1.f / 16777216.f);
        x += q.noise_std * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
      }
    }
    v[idx] = x;
  }
}

// Separable Gaussian blur along one axis of [B*C][ox][oy][oz] (MONAI GaussianFilter, truncated = 4 sigma, "erf" kernel:
// w[d] = 0.5 (erf((d + .5) / (sigma sqrt 2)) - erf((d - .5) / (sigma sqrt 2))), zero padding).  sigma <= 0 for a sample: copy.
__global__ void __launch_bounds__(256) aug_blur_axis_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C,
                                                             int ox, int oy, int oz, int axis, const float* __restrict__ sigma /*[B][3]*/) {
  const long long per_image = (long long)C * ox * oy * oz;
  const long long total = per_image * B;
  const int n_ax = axis == 0 ? ox : (axis == 1 ? oy : oz);
  const long long stride = axis == 0 ? (long long)oy * oz : (axis == 1 ? oz : 1);
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int b = (int)(idx / per_image);
    const float sg = sigma[3 * b + axis];
    if (!(sg > 0.f)) { dst[idx] = src[idx]; continue; }
    const int pos = (int)((idx / stride) % n_ax);
    const int tail = max((int)(4.f * sg + 0.5f), 1);
    const float inv = 1.f / (sg * 1.41421356237f);
    float acc = 0.f;
    for (int d = -tail; d <= tail; ++d) {
      const int q = pos + d;
      if (q < 0 || q >= n_ax) continue;
      const float w = 0.5f * (erff(((float)d + 0.5f) * inv) - erff(((float)d - 0.5f) * inv));
      acc += w * src[idx + (long long)d * stride];
    }
    dst[idx] = acc;
  }
}

// RandGaussianSharpen: out = blur1 + alpha * (blur1 - blur2), blur2 = gaussian(sigma2)(blur1);  alpha == 0 for a sample: keep `cur`
__global__ void __launch_bounds__(256) aug_sharpen_combine_kernel(float* __restrict__ cur, const float* __restrict__ blur1,
                                                                   const float* __restrict__ blur2, long long per_image, int B,
                                                                   const float* __restrict__ alpha /*[B]*/) {
  const long long total = per_image * B;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const float a = alpha[idx / per_image];
    if (a != 0.f) cur[idx] = blur1[idx] + a * (blur1[idx] - blur2[idx]);
  }
}

static inline unsigned aug_grid(long long total) {
  long long blocks = (total + 255) / 256;
  return (unsigned)(blocks > 148 * 16 ? 148 * 16 : (blocks < 1 ? 1 : blocks));
}

}  // namespace mmnn

extern "C" {

int mmnn_sizeof_aug_spatial() { return (int)sizeof(mmnn::AugSpatial); }
int mmnn_sizeof_aug_intensity() { return (int)sizeof(mmnn::AugIntensity); }

// Spatial part: raw [B][C][X][Y][Z] -> [B][C][ox][oy][oz] in [0, 1].  scratch: 2*B uint32 (the min / max pass of preprocess.cu runs
// first); sp: [B] AugSpatial (device).
int mmnn_preprocess_minmax(const float* src, void* scratch, int B, long long per_image, void* stream);

int mmnn_augment_resample(const float* src, float* dst, void* scratch, const void* sp, int B, int C, int X, int Y, int Z, int ox, int oy,
                          int oz, float mean, float std, void* stream) {
  using namespace mmnn;
  if (B <= 0 || C <= 0 || X <= 0 || Y <= 0 || Z <= 0 || ox <= 0 || oy <= 0 || oz <= 0) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  const int rc = mmnn_preprocess_minmax(src, scratch, B, (long long)C * X * Y * Z, stream);
  if (rc != 0) return rc;
  AugResampleParams p = {src, dst, (const uint32_t*)scratch, (const AugSpatial*)sp, B, C, X, Y, Z, ox, oy, oz, mean, std};
  ProfScope ps(PC_PREPROC, st, 1);
  augment_resample_kernel<<<aug_grid((long long)B * C * ox * oy * oz), 256, 0, st>>>(p);
  return (int)cudaGetLastError();
}

// Intensity part, in place on v [B][C][ox][oy][oz].  prm: [B] AugIntensity; smooth_sigma: [B][3] (<= 0: off); sharpen: sigma1 [B][3],
// sigma2 [B][3], alpha [B] (alpha == 0: off) in one array of 7*B floats; tmp: 3 volumes of the same size as v; scratch: 2*B uint32.
// Order of the reference chain: shift, contrast, smooth, sharpen, histogram shift, noise.
int mmnn_augment_intensity(float* v, float* tmp, void* scratch, const void* prm, const float* smooth_sigma, const float* sharpen, int B,
                           int C, int ox, int oy, int oz, int any_smooth, int any_sharpen, void* stream) {
  using namespace mmnn;
  if (B <= 0 || C <= 0 || ox <= 0 || oy <= 0 || oz <= 0) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  const long long per_image = (long long)C * ox * oy * oz;
  const unsigned grid = aug_grid(per_image * B);
  uint32_t* ext2 = (uint32_t*)scratch;
  float* t0 = tmp;
  float* t1 = tmp + per_image * B;
  float* t2 = tmp + 2 * per_image * B;
  aug_minmax_kernel<<<B, 256, 0, st>>>(v, per_image, ext2);
  augment_intensity_kernel<<<grid, 256, 0, st>>>(v, per_image, B, (const AugIntensity*)prm, ext2, 0);
  if (any_smooth) {
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(v, t0, B, C, ox, oy, oz, 0, smooth_sigma);
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(t0, t1, B, C, ox, oy, oz, 1, smooth_sigma);
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(t1, v, B, C, ox, oy, oz, 2, smooth_sigma);
  }
  if (any_sharpen) {
    // three arrays laid out one after the other: sigma1 [B][3], sigma2 [B][3], alpha [B]
    const float* s1 = sharpen;
    const float* s2 = sharpen + 3 * B;
    const float* alpha = sharpen + 6 * B;
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(v, t0, B, C, ox, oy, oz, 0, s1);
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(t0, t1, B, C, ox, oy, oz, 1, s1);
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(t1, t0, B, C, ox, oy, oz, 2, s1);      // t0 = blur1
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(t0, t1, B, C, ox, oy, oz, 0, s2);
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(t1, t2, B, C, ox, oy, oz, 1, s2);
    aug_blur_axis_kernel<<<grid, 256, 0, st>>>(t2, t1, B, C, ox, oy, oz, 2, s2);      // t1 = blur2
    aug_sharpen_combine_kernel<<<grid, 256, 0, st>>>(v, t0, t1, per_image, B, alpha);
  }
  aug_minmax_kernel<<<B, 256, 0, st>>>(v, per_image, ext2);
  augment_intensity_kernel<<<grid, 256, 0, st>>>(v, per_image, B, (const AugIntensity*)prm, ext2, 1);
  return (int)cudaGetLastError();
}
}
