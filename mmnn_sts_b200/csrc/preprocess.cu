// GPU input pipeline of the reference's validation / inference transform chain:
//   Normalize(IMAGE_DATA_MEAN, IMAGE_DATA_STDDEV)   /root/reference/utils/utils.py:346-355
//       n = (x - mean*max(x)) / (std*max(x)),  max over the whole multi-channel image of ONE patient
//   ScaleIntensity()                                  monai 1.2 (not vendored): minv 0, maxv 1, whole image
//       s = (n - min(n)) / (max(n) - min(n));  all-equal image -> n * 0
//   Resize(spatial_size)                              monai 1.2 default mode "area" = adaptive average pooling
// (/root/reference/main.py:86-92 val_transforms; the deterministic head and tail of train_transforms :64-83).
// Two HBM-bound passes over the raw volume: (1) per-patient min / max (order-preserving uint encoding, atomicMax),
// (2) one thread per output voxel: the affine maps applied element-wise in the reference's operation order (fp32),
// summed over the adaptive window, divided by the window size.  Algorithmic bytes: 2 x raw volume + output.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "prof.h"

namespace mmnn {

__device__ __forceinline__ uint32_t enc_f32(float f) {   // monotone float -> uint (0 is below every real value)
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(uint32_t e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// ext[b][0] = enc(max x), ext[b][1] = enc(max -x): both start at 0 (one memset)
__global__ void __launch_bounds__(256) preproc_minmax_kernel(const float* __restrict__ src, long long per_image, uint32_t* __restrict__ ext) {
  const int b = blockIdx.y;
  const float* s = src + (long long)b * per_image;
  float mx = -INFINITY, mn = INFINITY;
  const long long n4 = ((reinterpret_cast<uintptr_t>(s) & 15u) == 0) ? per_image >> 2 : 0;
  const float4* s4 = reinterpret_cast<const float4*>(s);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = __ldg(s4 + i);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    mn = fminf(fminf(mn, fminf(v.x, v.y)), fminf(v.z, v.w));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < per_image; i += (long long)gridDim.x * 256) {
    const float v = __ldg(s + i);
    mx = fmaxf(mx, v); mn = fminf(mn, v);
  }
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
    if (mx >= mn) {   // the block saw at least one element
      atomicMax(ext + 2 * b, enc_f32(mx));
      atomicMax(ext + 2 * b + 1, enc_f32(-mn));
    }
  }
}

struct PreprocParams {
  const float* src;   // [B][C][X][Y][Z] raw intensities
  float* dst;         // [B][C][ox][oy][oz]
  const uint32_t* ext;
  int B, C, X, Y, Z, ox, oy, oz;
  float mean, std;
};

__global__ void __launch_bounds__(256) preproc_resize_kernel(const __grid_constant__ PreprocParams p) {
  const long long total = (long long)p.B * p.C * p.ox * p.oy * p.oz;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    long long t = idx;
    const int k = (int)(t % p.oz); t /= p.oz;
    const int j = (int)(t % p.oy); t /= p.oy;
    const int i = (int)(t % p.ox); t /= p.ox;
    const int c = (int)(t % p.C);
    const int b = (int)(t / p.C);
    const float M = dec_f32(p.ext[2 * b]), m = -dec_f32(p.ext[2 * b + 1]);
    // Normalize: n = (x - mean*M) / (std*M);  ScaleIntensity over n: the extremes of n are the images of the extremes of x
    // __fmul_rn: the products must be ROUNDED to fp32 like numpy's (nvcc would otherwise contract x - mean*M into one
    // FMA; with mean*M ~ 5e5 against intensities ~ 1e3 that single rounding shifts every output by ~3e-5)
    const float sub = __fmul_rn(p.mean, M), den = __fmul_rn(p.std, M);
    const float na = (m - sub) / den, nb = (M - sub) / den;
    const float nmin = fminf(na, nb), nmax = fmaxf(na, nb);
    const float range = nmax - nmin;
    // adaptive windows (torch adaptive_avg_pool3d): [floor(i*X/ox), ceil((i+1)*X/ox))
    const int x0 = (int)(((long long)i * p.X) / p.ox), x1 = (int)((((long long)i + 1) * p.X + p.ox - 1) / p.ox);
    const int y0 = (int)(((long long)j * p.Y) / p.oy), y1 = (int)((((long long)j + 1) * p.Y + p.oy - 1) / p.oy);
    const int z0 = (int)(((long long)k * p.Z) / p.oz), z1 = (int)((((long long)k + 1) * p.Z + p.oz - 1) / p.oz);
    const float* s = p.src + ((long long)b * p.C + c) * p.X * p.Y * p.Z;
    float acc = 0.f;
    for (int x = x0; x < x1; ++x)
      for (int y = y0; y < y1; ++y) {
        const float* row = s + ((long long)x * p.Y + y) * p.Z;
        for (int z = z0; z < z1; ++z) {
          const float n = (__ldg(row + z) - sub) / den;
          acc += (range == 0.f) ? n * 0.f : (n - nmin) / range;
        }
      }
    p.dst[idx] = acc / (float)((x1 - x0) * (y1 - y0) * (z1 - z0));
  }
}

}  // namespace mmnn

extern "C" {
// pass 1 alone (per-patient extremes into scratch[2*B]); shared with the training chain (augment.cu)
int mmnn_preprocess_minmax(const float* src, void* scratch, int B, long long per_image, void* stream) {
  using namespace mmnn;
  if (B <= 0 || per_image <= 0) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(uint32_t) * 2 * B, st);
  if (e != cudaSuccess) return (int)e;
  int gx = (int)((per_image / 4 + 255) / 256);
  gx = gx < 1 ? 1 : (gx > 148 * 8 / (B < 8 ? B : 8) ? 148 * 8 / (B < 8 ? B : 8) : gx);
  preproc_minmax_kernel<<<dim3(gx, B), 256, 0, st>>>(src, per_image, (uint32_t*)scratch);
  return (int)cudaGetLastError();
}
// src fp32 [B][C][X][Y][Z] (device), dst fp32 [B][C][ox][oy][oz], scratch: 2*B uint32 (device).  Stream-ordered.
int mmnn_preprocess_volumes(const float* src, float* dst, void* scratch, int B, int C, int X, int Y, int Z, int ox, int oy,
                            int oz, float mean, float std, void* stream) {
  using namespace mmnn;
  if (B <= 0 || C <= 0 || X <= 0 || Y <= 0 || Z <= 0 || ox <= 0 || oy <= 0 || oz <= 0) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(uint32_t) * 2 * B, st);
  if (e != cudaSuccess) return (int)e;
  const long long per_image = (long long)C * X * Y * Z;
  ProfScope ps(PC_PREPROC, st, 2);
  int gx = (int)((per_image / 4 + 255) / 256);
  gx = gx < 1 ? 1 : (gx > 148 * 8 / (B < 8 ? B : 8) ? 148 * 8 / (B < 8 ? B : 8) : gx);
  preproc_minmax_kernel<<<dim3(gx, B), 256, 0, st>>>(src, per_image, (uint32_t*)scratch);
  PreprocParams p = {src, dst, (const uint32_t*)scratch, B, C, X, Y, Z, ox, oy, oz, mean, std};
  const long long total = (long long)B * C * ox * oy * oz;
  long long blocks = (total + 255) / 256;
  preproc_resize_kernel<<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, st>>>(p);
  return (int)cudaGetLastError();
}
}
