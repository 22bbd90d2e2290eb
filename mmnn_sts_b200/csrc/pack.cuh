// Weight packing: fp32 master weights (reference layouts) -> bf16 k-block images in the engine's chunk-plane layout
//   dst[n tile][k-block = (tap, cb)][chunk][n (NT)][8]      so that one 1-D bulk copy brings a whole B k-block to smem.
#pragma once
#include "common.cuh"

namespace mmnn {

enum { PACK_GENERIC = 0, PACK_STEM = 1, PACK_STEM_SW32 = 2 };

struct PackDesc {
  const float* src;
  bf16* dst;
  int N;      // valid output rows of the packed operand
  int NT;     // tile width (rows per tile, zero padded)
  int Cin;    // channels per tap
  int kbw;    // k-block width
  int ntaps;
  int mode;
  int cin_real;  // stem: real input channels (1 or 2); the packed form always has 2
  int f16;       // 1: emit IEEE fp16 (forward images when activations are fp16), 0: bf16
  long long sn, sc, st;  // source strides (elements) for n, channel, tap
};

static __global__ void pack_weights_kernel(const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  if (d.mode == PACK_STEM_SW32) {
    // stem.cuh B operand: dst[tap = (dz*4+dy)*4+dx][n (64)][32 B] = the 16 space-to-depth channels (pz,py,px,ci) of one tap,
    // K-major rows of 32 bytes written with the 32-byte swizzle (16-byte half h of row n at h ^ bit2(n): the image is
    // copied to a 256-byte aligned shared-memory address, so address bit 7 of a row is bit 2 of n)
    const long long cells = 64LL * 64 * 2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < cells; idx += (long long)gridDim.x * blockDim.x) {
      const int hp = (int)(idx & 1), n = (int)((idx >> 1) & 63), tap = (int)(idx >> 7);
      const int half = hp ^ ((n >> 2) & 1);
      const int dz = tap >> 4, dy = (tap >> 2) & 3, dx = tap & 3;
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int kz = 2 * dz + half, ky = 2 * dy + ((e >> 2) & 1), kx = 2 * dx + ((e >> 1) & 1), ci = e & 1;
        float val = 0.f;
        if (kz < 7 && ky < 7 && kx < 7 && ci < d.cin_real && n < d.N)
          val = d.src[((((long long)n * d.cin_real + ci) * 7 + kz) * 7 + ky) * 7 + kx];
        w[e] = val;
      }
      uint4 o;
      if (d.f16) {
        o.x = pack2<true>(w[0], w[1]); o.y = pack2<true>(w[2], w[3]); o.z = pack2<true>(w[4], w[5]); o.w = pack2<true>(w[6], w[7]);
      } else {
        o.x = pack_bf16(w[0], w[1]); o.y = pack_bf16(w[2], w[3]); o.z = pack_bf16(w[4], w[5]); o.w = pack_bf16(w[6], w[7]);
      }
      reinterpret_cast<uint4*>(d.dst)[idx] = o;
    }
    return;
  }
  const int planes = d.kbw / 8;
  const int kb_per_tap = (d.Cin + d.kbw - 1) / d.kbw;
  const int KB = d.ntaps * kb_per_tap;
  const int ntile = (d.N + d.NT - 1) / d.NT;
  const long long total = (long long)ntile * KB * planes * d.NT;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int n = (int)(t % d.NT); t /= d.NT;
    const int chunk = (int)(t % planes); t /= planes;
    const int kb = (int)(t % KB);
    const int tile = (int)(t / KB);
    const int tap = kb / kb_per_tap, cb = kb - tap * kb_per_tap;
    const int ng = tile * d.NT + n;
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cb * d.kbw + chunk * 8 + e;
      float val = 0.f;
      if (ng < d.N && c < d.Cin) {
        if (d.mode == PACK_GENERIC) {
          val = d.src[(long long)ng * d.sn + (long long)c * d.sc + (long long)tap * d.st];
        } else {
          const int dz = tap >> 2, dy = tap & 3, dx = c >> 4;
          const int kz = 2 * dz + ((c >> 3) & 1), ky = 2 * dy + ((c >> 2) & 1), kx = 2 * dx + ((c >> 1) & 1), ci = c & 1;
          if (kz < 7 && ky < 7 && kx < 7 && ci < d.cin_real)
            val = d.src[((((long long)ng * d.cin_real + ci) * 7 + kz) * 7 + ky) * 7 + kx];
        }
      }
      w[e] = val;
    }
    uint4 o;
    if (d.f16) {
      o.x = pack2<true>(w[0], w[1]); o.y = pack2<true>(w[2], w[3]); o.z = pack2<true>(w[4], w[5]); o.w = pack2<true>(w[6], w[7]);
    } else {
      o.x = pack_bf16(w[0], w[1]); o.y = pack_bf16(w[2], w[3]); o.z = pack_bf16(w[4], w[5]); o.w = pack_bf16(w[6], w[7]);
    }
    reinterpret_cast<uint4*>(d.dst)[idx] = o;
  }
}

static inline size_t packed_elems(int N, int NT, int Cin, int kbw, int ntaps) {
  const int kb_per_tap = (Cin + kbw - 1) / kbw;
  const int ntile = (N + NT - 1) / NT;
  return (size_t)ntile * ntaps * kb_per_tap * (kbw / 8) * NT * 8;
}

}  // namespace mmnn
