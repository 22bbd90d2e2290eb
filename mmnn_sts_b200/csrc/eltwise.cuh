// Memory-bound kernels of the DenseNet-3D trunk (everything that is not a GEMM): layout conversion, BN+ReLU+max-pool,
// BN+ReLU+avg-pool (transition), BatchNorm backward application, slice extraction, norm5, and their backwards.
// Design rules: one thread owns a 16-byte cell (8 bf16 channels of one voxel); consecutive threads own consecutive
// cells so every warp access is a run of contiguous 128-bit vectors; per-channel statistics are reduced in
// registers -> shared memory -> one fp64 atomic per channel per block; grids are sized in multiples of the SM count.
#pragma once
#include "common.cuh"
#include "engine.cuh"

namespace mmnn {

constexpr int EW_THREADS = 256;

// ACT = forward activation tensor (fp16 when kActF16), GRD = gradient tensor (bf16)
constexpr bool ACT = kActF16;
constexpr bool GRD = false;
template <bool F16>
MMNN_DEVINL void unpack8(const uint4& v, float* f) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) unpack2<F16>(w[i], f[2 * i], f[2 * i + 1]);
}
template <bool F16>
MMNN_DEVINL uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack2<F16>(f[0], f[1]); o.y = pack2<F16>(f[2], f[3]); o.z = pack2<F16>(f[4], f[5]); o.w = pack2<F16>(f[6], f[7]);
  return o;
}

// Block-level reduction of per-thread 8-channel partial sums (two quantities) followed by fp64 atomics.
// Thread layout: tid = rowslot * cpr + chunk  (cpr = chunks per row = C/8, cpr divides EW_THREADS).
MMNN_DEVINL void block_channel_reduce(float* red /*[EW_THREADS][16]*/, const float* s1, const float* s2, int cpr,
                                      double* dst1, double* dst2) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int e = 0; e < 8; ++e) { red[tid * 16 + e] = s1[e]; red[tid * 16 + 8 + e] = s2[e]; }
  __syncthreads();
  // cpr*16 outputs, each the sum of floor(EW_THREADS/cpr) entries (threads beyond rows*cpr carry zeros and are skipped)
  for (int o = tid; o < cpr * 16; o += EW_THREADS) {
    const int chunk = o >> 4, e = o & 15;
    float acc = 0.f;
    for (int r = 0; r < EW_THREADS / cpr; ++r) acc += red[(r * cpr + chunk) * 16 + e];
    if (e < 8) atomicAdd(dst1 + chunk * 8 + e, (double)acc);
    else atomicAdd(dst2 + chunk * 8 + (e - 8), (double)acc);
  }
}

// ------------------------------------------------------------------------------------------------- E1: input pack
// image NCDHW fp32 [B][cin][X][Y][Z]  ->  padded space-to-depth bf16 [B][Sz][Sy][Sx][(pz,py,px,c2)]   (pad 3, stride 2)
// TIN: float (the reference's collate output, /root/reference/utils/utils.py:98-99) or __half (a loader that ships 16-bit volumes:
// half the host->device bytes; the values are rounded to the activation format here either way, so the results are identical)
template <typename TIN>
// dst_w (optional): the same image in bf16 (each value = the bf16 rounding of the stored activation-format value), the TMA operand
// of the stem weight gradient, whose MMA runs in bf16 next to the bf16 gradient tiles
static __global__ void s2d_pack_kernel(const TIN* __restrict__ img, bf16* __restrict__ dst, bf16* __restrict__ dst_w, int B, int cin,
                                       int X, int Y, int Z, int Sz, int Sy, int Sx) {
  const long long total = (long long)B * Sz * Sy * Sx * 2;  // 16-byte cells (pz = cell & 1)
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int pz = (int)(t & 1); t >>= 1;
    const int qx = (int)(t % Sx); t /= Sx;
    const int qy = (int)(t % Sy); t /= Sy;
    const int qz = (int)(t % Sz);
    const int b = (int)(t / Sz);
    float f[8];
    const int iz = 2 * qz + pz - 3;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int py = e >> 2, px = (e >> 1) & 1, c = e & 1;
      const int iy = 2 * qy + py - 3, ix = 2 * qx + px - 3;
      float v = 0.f;
      if (c < cin && iz >= 0 && iz < X && iy >= 0 && iy < Y && ix >= 0 && ix < Z)
        v = (float)img[((((long long)b * cin + c) * X + iz) * Y + iy) * Z + ix];
      f[e] = v;
    }
    uint4 o = pack8<ACT>(f);
    reinterpret_cast<uint4*>(dst)[idx] = o;
    if (dst_w != nullptr) {
      convert8<kActF16, false>(o);
      reinterpret_cast<uint4*>(dst_w)[idx] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------- E2: BN0+ReLU+maxpool
struct PoolParams {
  int B, D0, H0, W0, D1, H1, W1;
  const bf16* src;  // [B*D0*H0*W0][64]
  BnSrc bn;
  bf16* dst;        // block-1 buffer, channels [0,64)
  long long dst_pitch;
  uint8_t* argmax;  // [M1][64] window code 0..26 (first maximum in (dz,dy,dx) scan order, like torch)
  double* st_sum;
  double* st_sq;
};

static __global__ void __launch_bounds__(EW_THREADS) bnrelu_maxpool_kernel(const __grid_constant__ PoolParams p) {
  pdl_wait(); pdl_trigger();   // launched with launch_pdl (launch.h)
  __shared__ float red[EW_THREADS * 16];
  __shared__ float coef[128];
  for (int c = threadIdx.x; c < 64; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(p.bn, c, mean, rstd);
    const float s = p.bn.gamma[c] * rstd;
    coef[c] = s; coef[64 + c] = p.bn.beta[c] - mean * s;
  }
  __syncthreads();
  const int chunk = threadIdx.x & 7;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = coef[chunk * 8 + e]; sh[e] = coef[64 + chunk * 8 + e]; }
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long M1 = (long long)p.B * p.D1 * p.H1 * p.W1;
  for (long long m = (long long)blockIdx.x * (EW_THREADS / 8) + (threadIdx.x >> 3); m < M1; m += (long long)gridDim.x * (EW_THREADS / 8)) {
    unsigned t = (unsigned)m;                       // M1 < 2^31 (checked by the host): 32-bit index arithmetic
    const int x = (int)(t % (unsigned)p.W1); t /= (unsigned)p.W1;
    const int y = (int)(t % (unsigned)p.H1); t /= (unsigned)p.H1;
    const int z = (int)(t % (unsigned)p.D1);
    const int b = (int)(t / (unsigned)p.D1);
    float best[8];
    int code[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; code[e] = 0; }
    // one z-plane of the window at a time: its 9 loads are issued together (predicated), then compared in the
    // (dz, dy, dx) scan order so that the first maximum wins like torch
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
      const int iz = 2 * z + dz - 1;
      const bool zok = (unsigned)iz < (unsigned)p.D0;
      uint4 v[9];
      bool ok[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int iy = 2 * y + k / 3 - 1, ix = 2 * x + k % 3 - 1;
        ok[k] = zok && (unsigned)iy < (unsigned)p.H0 && (unsigned)ix < (unsigned)p.W0;
        v[k] = make_uint4(0, 0, 0, 0);
        if (ok[k]) v[k] = ldg16(p.src + ((((long long)b * p.D0 + iz) * p.H0 + iy) * p.W0 + ix) * 64 + chunk * 8);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (ok[k]) {
          float f[8];
          unpack8<ACT>(v[k], f);
          const int cd = dz * 9 + k;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float a = fmaxf(fmaf(f[e], sc[e], sh[e]), 0.f);
            if (a > best[e]) { best[e] = a; code[e] = cd; }
          }
        }
      }
    }
    float r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { r[e] = round16<ACT>(best[e]); s1[e] += r[e]; s2[e] += r[e] * r[e]; }
    *reinterpret_cast<uint4*>(p.dst + m * p.dst_pitch + chunk * 8) = pack8<ACT>(r);
    uint2 cdv;
    cdv.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
    cdv.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
    *reinterpret_cast<uint2*>(p.argmax + m * 64 + chunk * 8) = cdv;
  }
  block_channel_reduce(red, s1, s2, 8, p.st_sum, p.st_sq);
}

// backward of pool0+relu0: gather form. dR[m0][c] = relu'(.) * sum over windows whose argmax is m0 of dPool[o][c];
// accumulates the BN0 backward statistics (sum dR, sum dR*xhat).
struct PoolBwdParams {
  int B, D0, H0, W0, D1, H1, W1;
  const bf16* x;        // stem conv output [M0][64]
  BnSrc bn;
  const float* dpool;   // fp32 gradient buffer of block 1, channels [0,64)
  long long dpool_pitch;
  const uint8_t* argmax;
  bf16* dr;             // [M0][64]
  double* g_sum;
  double* g_dot;
};

static __global__ void __launch_bounds__(EW_THREADS) maxpool_bnrelu_bwd_kernel(const __grid_constant__ PoolBwdParams p) {
  pdl_wait(); pdl_trigger();   // launched with launch_pdl (launch.h)
  __shared__ float red[EW_THREADS * 16];
  __shared__ float coef[256];
  for (int c = threadIdx.x; c < 64; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(p.bn, c, mean, rstd);
    const float s = p.bn.gamma[c] * rstd;
    coef[c] = s; coef[64 + c] = p.bn.beta[c] - mean * s; coef[128 + c] = mean; coef[192 + c] = rstd;
  }
  __syncthreads();
  const int chunk = threadIdx.x & 7;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long M0 = (long long)p.B * p.D0 * p.H0 * p.W0;
  for (long long m = (long long)blockIdx.x * (EW_THREADS / 8) + (threadIdx.x >> 3); m < M0; m += (long long)gridDim.x * (EW_THREADS / 8)) {
    unsigned t = (unsigned)m;                       // M0 < 2^31 (checked by the host)
    const int ix = (int)(t % (unsigned)p.W0); t /= (unsigned)p.W0;
    const int iy = (int)(t % (unsigned)p.H0); t /= (unsigned)p.H0;
    const int iz = (int)(t % (unsigned)p.D0);
    const int b = (int)(t / (unsigned)p.D0);
    const uint4 xraw = ldg16(p.x + m * 64 + chunk * 8);   // issued before the window loop
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // windows o with 2o-1 <= i <= 2o+1
    for (int oz = iz >> 1; oz <= (iz + 1) >> 1; ++oz) {
      if (oz >= p.D1) continue;
      const int dz = iz - 2 * oz + 1;
      for (int oy = iy >> 1; oy <= (iy + 1) >> 1; ++oy) {
        if (oy >= p.H1) continue;
        const int dy = iy - 2 * oy + 1;
        for (int ox = ix >> 1; ox <= (ix + 1) >> 1; ++ox) {
          if (ox >= p.W1) continue;
          const int dx = ix - 2 * ox + 1;
          const int cd = (dz * 3 + dy) * 3 + dx;
          const long long o = (((long long)b * p.D1 + oz) * p.H1 + oy) * p.W1 + ox;
          const uint2 cdv = *reinterpret_cast<const uint2*>(p.argmax + o * 64 + chunk * 8);
          const float4 ga = *reinterpret_cast<const float4*>(p.dpool + o * p.dpool_pitch + chunk * 8);
          const float4 gb = *reinterpret_cast<const float4*>(p.dpool + o * p.dpool_pitch + chunk * 8 + 4);
          const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = (e < 4 ? (cdv.x >> (8 * e)) : (cdv.y >> (8 * (e - 4)))) & 0xff;
            if (c == cd) g[e] += gg[e];
          }
        }
      }
    }
    float f[8], r[8];
    unpack8<ACT>(xraw, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = chunk * 8 + e;
      const bool act = fmaf(f[e], coef[c], coef[64 + c]) > 0.f;
      r[e] = act ? round_bf16(g[e]) : 0.f;
      s1[e] += r[e];
      s2[e] += r[e] * (f[e] - coef[128 + c]) * coef[192 + c];
    }
    *reinterpret_cast<uint4*>(p.dr + m * 64 + chunk * 8) = pack8<GRD>(r);
  }
  block_channel_reduce(red, s1, s2, 8, p.g_sum, p.g_dot);
}

// Tiled scatter form of the same pass (the default).  The gather kernel above visits up to 8 windows per INPUT voxel (argmax codes
// + fp32 gradients re-read 8x, a compare per window and channel): 0.46 ms at configs[1] for 620 MB of algorithmic traffic.  Here a
// CTA owns an input tile of MPB_TZ x MPB_TY x MPB_TX voxels x 64 channels as an fp32 accumulator in shared memory; every pooling
// window that overlaps the tile routes its gradient to its arg-max voxel (ONE shared-memory add per window and channel, dropped when
// the voxel belongs to a neighbouring tile); then the tile is streamed out: mask, round, store dR, BN-backward statistics.
// Windows of the same parity class (oz & 1, oy & 1, ox & 1) have disjoint 3x3x3 footprints (stride 2), so the eight classes are
// applied one after the other with plain read-modify-writes: no atomics, a fixed summation order, bit-reproducible.
constexpr int MPB_TZ = 4, MPB_TY = 8, MPB_TX = 8, MPB_VOX = MPB_TZ * MPB_TY * MPB_TX;
constexpr int MPB_WZ = MPB_TZ / 2 + 1, MPB_WY = MPB_TY / 2 + 1, MPB_WX = MPB_TX / 2 + 1, MPB_WIN = MPB_WZ * MPB_WY * MPB_WX;   // 75 windows
constexpr int MPB_SMEM = MPB_VOX * 64 * 4 + MPB_WIN * 64 * 4 + MPB_WIN * 64;   // accumulator 64 KB + staged gradients 19 KB + codes 5 KB: two CTAs per SM

static __global__ void __launch_bounds__(EW_THREADS) maxpool_bnrelu_bwd_tiled_kernel(const __grid_constant__ PoolBwdParams p) {
  pdl_wait(); pdl_trigger();
  extern __shared__ __align__(16) float acc[];          // [MPB_VOX][64]
  float* wgr = acc + MPB_VOX * 64;                       // [MPB_WIN][64] gradients of the windows overlapping the tile
  uint8_t* wcd = reinterpret_cast<uint8_t*>(wgr + MPB_WIN * 64);   // [MPB_WIN][64] arg-max codes (255: window outside the volume)
  __shared__ float red[EW_THREADS * 16];
  __shared__ float coef[256];
  for (int c = threadIdx.x; c < 64; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(p.bn, c, mean, rstd);
    const float s = p.bn.gamma[c] * rstd;
    coef[c] = s; coef[64 + c] = p.bn.beta[c] - mean * s; coef[128 + c] = mean; coef[192 + c] = rstd;
  }
  const int tid = threadIdx.x;
  const int chunk = tid & 7;
  const int tz = (p.D0 + MPB_TZ - 1) / MPB_TZ, ty = (p.H0 + MPB_TY - 1) / MPB_TY, tx = (p.W0 + MPB_TX - 1) / MPB_TX;
  const long long ntiles = (long long)p.B * tz * ty * tx;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    long long r = t;
    const int x0 = (int)(r % tx) * MPB_TX; r /= tx;
    const int y0 = (int)(r % ty) * MPB_TY; r /= ty;
    const int z0 = (int)(r % tz) * MPB_TZ;
    const int b = (int)(r / tz);
    __syncthreads();                                     // previous tile fully streamed out (also orders the coef table on the first pass)
    // ---- stage every overlapping window (o in [i0/2, (i0+T)/2] per axis) in ONE round of loads; zero the accumulator meanwhile
    for (int i = tid; i < MPB_WIN * 8; i += EW_THREADS) {
      int w = i >> 3;
      const int lx = w % MPB_WX, ly = (w / MPB_WX) % MPB_WY, lz = w / (MPB_WX * MPB_WY);
      const int oz = z0 / 2 + lz, oy = y0 / 2 + ly, ox = x0 / 2 + lx;
      uint2 cdv = make_uint2(0xffffffffu, 0xffffffffu);
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
      if (oz < p.D1 && oy < p.H1 && ox < p.W1) {
        const long long o = (((long long)b * p.D1 + oz) * p.H1 + oy) * p.W1 + ox;
        cdv = *reinterpret_cast<const uint2*>(p.argmax + o * 64 + chunk * 8);
        ga = *reinterpret_cast<const float4*>(p.dpool + o * p.dpool_pitch + chunk * 8);
        gb = *reinterpret_cast<const float4*>(p.dpool + o * p.dpool_pitch + chunk * 8 + 4);
      }
      *reinterpret_cast<uint2*>(wcd + w * 64 + chunk * 8) = cdv;
      *reinterpret_cast<float4*>(wgr + w * 64 + chunk * 8) = ga;
      *reinterpret_cast<float4*>(wgr + w * 64 + chunk * 8 + 4) = gb;
    }
    for (int i = tid; i < MPB_VOX * 16; i += EW_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // the tile's forward activations: issued now, consumed after the scatter
    uint4 xr[MPB_VOX * 8 / EW_THREADS];
#pragma unroll
    for (int k = 0; k < MPB_VOX * 8 / EW_THREADS; ++k) {
      const int v = (tid + k * EW_THREADS) >> 3;
      const int ix = x0 + v % MPB_TX, iy = y0 + (v / MPB_TX) % MPB_TY, iz = z0 + v / (MPB_TX * MPB_TY);
      xr[k] = make_uint4(0, 0, 0, 0);
      if (iz < p.D0 && iy < p.H0 && ix < p.W0)
        xr[k] = ldg16(p.x + ((((long long)b * p.D0 + iz) * p.H0 + iy) * p.W0 + ix) * 64 + chunk * 8);
    }
    __syncthreads();
    // ---- scatter, one parity class at a time (disjoint footprints inside a class: plain read-modify-write)
#pragma unroll 1
    for (int cls = 0; cls < 8; ++cls) {
      const int pz = cls >> 2, py = (cls >> 1) & 1, px = cls & 1;
      const int nz = (MPB_WZ - pz + 1) / 2, ny = (MPB_WY - py + 1) / 2, nx = (MPB_WX - px + 1) / 2;
      const int items = nz * ny * nx * 64;               // one thread per (window, channel)
      for (int i = tid; i < items; i += EW_THREADS) {
        const int c = i & 63;
        int w = i >> 6;
        const int lx = (w % nx) * 2 + px; w /= nx;
        const int ly = (w % ny) * 2 + py;
        const int lz = (w / ny) * 2 + pz;
        const int wi = (lz * MPB_WY + ly) * MPB_WX + lx;
        const int cd = wcd[wi * 64 + c];
        if (cd == 255) continue;
        const int dz = cd / 9, dy = (cd / 3) % 3, dx = cd % 3;
        const int iz = 2 * lz - 1 + dz, iy = 2 * ly - 1 + dy, ix = 2 * lx - 1 + dx;   // tile-local voxel of the arg-max (z0, y0, x0 are even)
        if ((unsigned)iz < (unsigned)MPB_TZ && (unsigned)iy < (unsigned)MPB_TY && (unsigned)ix < (unsigned)MPB_TX)
          acc[((iz * MPB_TY + iy) * MPB_TX + ix) * 64 + c] += wgr[wi * 64 + c];
      }
      __syncthreads();
    }
    // ---- stream the tile out
#pragma unroll
    for (int k = 0; k < MPB_VOX * 8 / EW_THREADS; ++k) {
      const int v = (tid + k * EW_THREADS) >> 3;
      const int ix = x0 + v % MPB_TX, iy = y0 + (v / MPB_TX) % MPB_TY, iz = z0 + v / (MPB_TX * MPB_TY);
      if (iz >= p.D0 || iy >= p.H0 || ix >= p.W0) continue;
      const long long m = (((long long)b * p.D0 + iz) * p.H0 + iy) * p.W0 + ix;
      float f[8], rr[8];
      unpack8<ACT>(xr[k], f);
      const float4 ga = *reinterpret_cast<const float4*>(acc + v * 64 + chunk * 8);
      const float4 gb = *reinterpret_cast<const float4*>(acc + v * 64 + chunk * 8 + 4);
      const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = chunk * 8 + e;
        const bool act = fmaf(f[e], coef[c], coef[64 + c]) > 0.f;
        rr[e] = act ? round_bf16(g[e]) : 0.f;
        s1[e] += rr[e];
        s2[e] += rr[e] * (f[e] - coef[128 + c]) * coef[192 + c];
      }
      *reinterpret_cast<uint4*>(p.dr + m * 64 + chunk * 8) = pack8<GRD>(rr);
    }
  }
  __syncthreads();
  block_channel_reduce(red, s1, s2, 8, p.g_sum, p.g_dot);
}

// ------------------------------------------------------------------------------------------------- E3: transition pooling
struct AvgPoolParams {
  int B, D, H, W;   // input dims; output dims = floor(/2)
  int C;
  const bf16* x;    // [M][pitch]
  long long x_pitch;
  BnSrc bn;
  bf16* pooled;     // [M/8][C]   forward output: mean over 2x2x2 of relu(bn(x))
  // backward
  const bf16* dpooled;  // [Mout][C]
  float* dx;            // fp32 [M][dx_pitch]   (overwritten)
  long long dx_pitch;
  double* g_sum;
  double* g_dot;
  const double* g_sum_in;  // pass 2
  const double* g_dot_in;
  float inv_count;
  int pre_rstd;            // pass 2: write gamma * (v - c1 - xhat c2) (the consumer's finalising pass applies rstd) instead of gamma * rstd * (...)
};

static __global__ void __launch_bounds__(EW_THREADS) bnrelu_avgpool_kernel(const __grid_constant__ AvgPoolParams p) {
  pdl_wait(); pdl_trigger();   // launched with launch_pdl (launch.h)
  extern __shared__ float coef[];  // [2][C]
  for (int c = threadIdx.x; c < p.C; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(p.bn, c, mean, rstd);
    const float s = p.bn.gamma[c] * rstd;
    coef[c] = s; coef[p.C + c] = p.bn.beta[c] - mean * s;
  }
  __syncthreads();
  const int cpr = p.C / 8;
  const int Do = p.D / 2, Ho = p.H / 2, Wo = p.W / 2;
  const long long total = (long long)p.B * Do * Ho * Wo * cpr;
  for (long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x; idx < total; idx += (long long)gridDim.x * EW_THREADS) {
    const int chunk = (int)(idx % cpr);
    long long t = idx / cpr;
    const long long mo = t;
    const int x = (int)(t % Wo); t /= Wo;
    const int y = (int)(t % Ho); t /= Ho;
    const int z = (int)(t % Do);
    const int b = (int)(t / Do);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long mi = (((long long)b * p.D + 2 * z + (k >> 2)) * p.H + 2 * y + ((k >> 1) & 1)) * p.W + 2 * x + (k & 1);
      float f[8];
      unpack8<ACT>(ldg16(p.x + mi * p.x_pitch + chunk * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += fmaxf(fmaf(f[e], coef[chunk * 8 + e], coef[p.C + chunk * 8 + e]), 0.f);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= 0.125f;
    *reinterpret_cast<uint4*>(p.pooled + mo * p.C + chunk * 8) = pack8<ACT>(acc);
  }
}

// PASS == 1: accumulate sum v, sum v*xhat with v = relu'(bn(x)) * dpooled[parent]/8.   PASS == 2: dx = k (v - c1 - xhat c2)
// PASS == 3: both in ONE pass with the deferred BatchNorm backward: dx = k v is written while the statistics are accumulated, and the
//            (c1, c2) terms are applied later by grad_finalize_kernel, which lists this BatchNorm among the channel's contributors
//            (k = gamma in batch mode; in eval mode k = gamma * rstd and c1 = c2 = 0, so the pass is complete by itself).
template <int PASS>
static __global__ void __launch_bounds__(EW_THREADS) avgpool_bnrelu_bwd_kernel(const __grid_constant__ AvgPoolParams p) {
  pdl_wait(); pdl_trigger();   // launched with launch_pdl (launch.h)
  extern __shared__ float coef[];  // [7][C]: s, t, mean, rstd, (pass2) c1, c2, output factor
  __shared__ float red[PASS != 2 ? EW_THREADS * 16 : 1];
  for (int c = threadIdx.x; c < p.C; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(p.bn, c, mean, rstd);
    const float s = p.bn.gamma[c] * rstd;
    coef[c] = s; coef[p.C + c] = p.bn.beta[c] - mean * s; coef[2 * p.C + c] = mean; coef[3 * p.C + c] = rstd;
    if (PASS == 2) {
      coef[4 * p.C + c] = (float)(p.g_sum_in[c] * (double)p.inv_count);
      coef[5 * p.C + c] = (float)(p.g_dot_in[c] * (double)p.inv_count);
    }
    if (PASS >= 2) coef[6 * p.C + c] = p.pre_rstd ? p.bn.gamma[c] : s;
  }
  __syncthreads();
  const int cpr = p.C / 8;
  const int rows_per_block = EW_THREADS / cpr;
  const int chunk = threadIdx.x % cpr;
  const int Do = p.D / 2, Ho = p.H / 2, Wo = p.W / 2;
  const long long M = (long long)p.B * p.D * p.H * p.W;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long m = (long long)blockIdx.x * rows_per_block + threadIdx.x / cpr; m < M && threadIdx.x < rows_per_block * cpr;
       m += (long long)gridDim.x * rows_per_block) {
    unsigned t = (unsigned)m;                       // M < 2^31 (checked by the host)
    const int x = (int)(t % (unsigned)p.W); t /= (unsigned)p.W;
    const int y = (int)(t % (unsigned)p.H); t /= (unsigned)p.H;
    const int z = (int)(t % (unsigned)p.D);
    const int b = (int)(t / (unsigned)p.D);
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if ((z >> 1) < Do && (y >> 1) < Ho && (x >> 1) < Wo) {
      const long long mo = (((long long)b * Do + (z >> 1)) * Ho + (y >> 1)) * Wo + (x >> 1);
      unpack8<GRD>(ldg16(p.dpooled + mo * p.C + chunk * 8), g);
    }
    float f[8];
    unpack8<ACT>(ldg16(p.x + m * p.x_pitch + chunk * 8), f);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = chunk * 8 + e;
      const bool act = fmaf(f[e], coef[c], coef[p.C + c]) > 0.f;
      const float v = act ? g[e] * 0.125f : 0.f;
      const float xh = (f[e] - coef[2 * p.C + c]) * coef[3 * p.C + c];
      if (PASS != 2) { s1[e] += v; s2[e] += v * xh; }
      if (PASS == 2) o[e] = coef[6 * p.C + c] * (v - coef[4 * p.C + c] - xh * coef[5 * p.C + c]);
      if (PASS == 3) o[e] = coef[6 * p.C + c] * v;
    }
    if (PASS >= 2) {
      float4* d = reinterpret_cast<float4*>(p.dx + m * p.dx_pitch + chunk * 8);
      d[0] = make_float4(o[0], o[1], o[2], o[3]);
      d[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  if (PASS != 2) block_channel_reduce(red, s1, s2, cpr, p.g_sum, p.g_dot);
}

// ------------------------------------------------------------------------------------------------- E4: BN backward apply
// out = gamma*rstd * ( v - mean(v) - xhat * mean(v*xhat) )       v: masked upstream gradient, x: forward BN input
enum { BA_OUT_BF16 = 0, BA_OUT_F32_ADD = 1, BA_OUT_F32_STORE = 2 };
struct BnApplyParams {
  long long M;
  int C;
  const bf16* v;      // bf16 gradient (or nullptr when v32 is used)
  const float* v32;   // fp32 gradient
  long long v_pitch;
  const bf16* x;
  long long x_pitch;
  BnSrc bn;
  const double* g_sum;
  const double* g_dot;
  float inv_count;
  void* out;
  long long out_pitch;
  // optional fused slice extraction (BA_OUT_F32_ADD only): the 32 channels [slice_c0, slice_c0+32) of the updated fp32
  // accumulator are final after this pass; emit them as the next layer's bf16 gradient operand (x dropout keep-scale)
  bf16* slice_out;          // [M][32] or null
  int slice_c0;
  const float* slice_scale; // [samples][32] or null
  int vps;
  int pre_rstd;             // leading factor gamma instead of gamma * rstd (deferred BatchNorm backward: the finalising pass applies rstd)
};

template <int OUT>
static __global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_kernel(const __grid_constant__ BnApplyParams p) {
  pdl_wait(); pdl_trigger();   // launched with launch_pdl (launch.h)
  extern __shared__ float coef[];  // [3][C]: a (on v), b (on x), d (const):  out = a*v + b*x + d
  for (int c = threadIdx.x; c < p.C; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(p.bn, c, mean, rstd);
    const float k = p.pre_rstd ? p.bn.gamma[c] : p.bn.gamma[c] * rstd;
    const float c1 = (float)(p.g_sum[c] * (double)p.inv_count);
    const float c2 = (float)(p.g_dot[c] * (double)p.inv_count);
    coef[c] = k;
    coef[p.C + c] = -k * rstd * c2;
    coef[2 * p.C + c] = -k * c1 + k * rstd * c2 * mean;
  }
  __syncthreads();
  const int cpr = p.C / 8;
  const long long total = p.M * cpr;
  // (row, chunk) of a cell are carried incrementally: one 64-bit division per thread instead of one per 16-byte cell
  // (the division made this pass issue-bound: 4.0 TB/s at 28 % issue utilisation with long-scoreboard stalls, ncu r01f)
  const long long stride = (long long)gridDim.x * EW_THREADS;
  const long long sm_ = stride / cpr;
  const int sc_ = (int)(stride - sm_ * cpr);
  long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x;
  long long m = idx / cpr;
  int chunk = (int)(idx - m * cpr);
  auto cell = [&](long long m, int chunk, const uint4& vraw, const float4& va, const float4& vb, const uint4& xraw) {
    float v[8], f[8], o[8];
    if (p.v != nullptr) {
      unpack8<GRD>(vraw, v);
    } else {
      v[0] = va.x; v[1] = va.y; v[2] = va.z; v[3] = va.w; v[4] = vb.x; v[5] = vb.y; v[6] = vb.z; v[7] = vb.w;
    }
    unpack8<ACT>(xraw, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = chunk * 8 + e;
      o[e] = fmaf(coef[c], v[e], fmaf(coef[p.C + c], f[e], coef[2 * p.C + c]));
    }
    if (OUT == BA_OUT_BF16) {
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + m * p.out_pitch + chunk * 8) = pack8<GRD>(o);
    } else {
      float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.out_pitch + chunk * 8);
      if (OUT == BA_OUT_F32_ADD) {
        const float4 a = d[0], b = d[1];
        float nv[8] = {a.x + o[0], a.y + o[1], a.z + o[2], a.w + o[3], b.x + o[4], b.y + o[5], b.z + o[6], b.w + o[7]};
        d[0] = make_float4(nv[0], nv[1], nv[2], nv[3]);
        d[1] = make_float4(nv[4], nv[5], nv[6], nv[7]);
        const int sc0 = chunk * 8 - p.slice_c0;
        if (p.slice_out != nullptr && sc0 >= 0 && sc0 < 32) {
          if (p.slice_scale != nullptr) {
            const float* cs = p.slice_scale + (m / p.vps) * 32 + sc0;
#pragma unroll
            for (int e = 0; e < 8; ++e) nv[e] *= __ldg(cs + e);
          }
          *reinterpret_cast<uint4*>(p.slice_out + m * 32 + sc0) = pack8<GRD>(nv);
        }
      } else {
        d[0] = make_float4(o[0], o[1], o[2], o[3]);
        d[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  };
  auto load = [&](long long m, int chunk, uint4& vraw, float4& va, float4& vb, uint4& xraw) {
    if (p.v != nullptr) {
      vraw = ldg16(p.v + m * p.v_pitch + chunk * 8);
    } else {
      va = *reinterpret_cast<const float4*>(p.v32 + m * p.v_pitch + chunk * 8);
      vb = *reinterpret_cast<const float4*>(p.v32 + m * p.v_pitch + chunk * 8 + 4);
    }
    xraw = ldg16(p.x + m * p.x_pitch + chunk * 8);
  };
  // two cells per iteration: both cells' loads are issued before either is consumed
  while (idx < total) {
    long long m1 = m + sm_;
    int chunk1 = chunk + sc_;
    if (chunk1 >= cpr) { chunk1 -= cpr; ++m1; }
    const bool two = idx + stride < total;
    uint4 vr0 = make_uint4(0, 0, 0, 0), xr0, vr1 = make_uint4(0, 0, 0, 0), xr1 = make_uint4(0, 0, 0, 0);
    float4 a0 = make_float4(0, 0, 0, 0), b0 = a0, a1 = a0, b1 = a0;
    load(m, chunk, vr0, a0, b0, xr0);
    if (two) load(m1, chunk1, vr1, a1, b1, xr1);
    cell(m, chunk, vr0, a0, b0, xr0);
    if (two) cell(m1, chunk1, vr1, a1, b1, xr1);
    idx += 2 * stride;
    m = m1 + sm_;
    chunk = chunk1 + sc_;
    if (chunk >= cpr) { chunk -= cpr; ++m; }
  }
}

// ------------------------------------------------------------------------------------------------- E4b: deferred BN backward
// Every norm1 of a dense block normalises the SAME tensor (the block buffer) with the SAME batch statistics; only gamma / beta
// differ per layer (/root/reference/models/densenet.py:76).  So the gradient of buffer channel c,
//     dX_c = T_c + sum over the layers l reading c of  gamma_lc rstd_c (dA1_lc - mean(dA1_lc) - xhat_c mean(dA1_lc xhat_c)),
// (T_c = the term of the block's consumer: transition / norm5, a BatchNorm over the same statistics too) factors as
//     dX_c = rstd_c (G_c - C1_c - xhat_c C2_c),   G_c = T_c / rstd_c + sum_l gamma_lc dA1_lc,
//     C1_c = sum_l gamma_lc mean(dA1_lc),  C2_c = sum_l gamma_lc mean(dA1_lc xhat_c).
// G is accumulated by the epilogue of every layer's 1x1x1 data-gradient GEMM (EP_MASK_STATS_ACC, engine.cuh) while the tile
// is in registers; this pass finalises a channel range once its last contributor has run: instead of one read-modify-write of
// ALL cin channels per layer (O(L^2) traffic, 12 B per element) each channel is finalised ONCE (O(L)).
// Eval mode (running statistics differ per layer): G already holds the complete gradient, the pass is the identity + cast.
struct FinalizeParams {
  long long M;
  int c_lo, nch;            // channel range of the block buffer (nch multiple of 8)
  float* G;                 // [M][g_pitch] fp32 accumulator; the final gradient is written back in place
  long long g_pitch;
  const bf16* x;            // block buffer (forward activations), same channel indexing
  long long x_pitch;
  BnSrc bn;                 // statistics of the block buffer's channels (any norm of the block: only mean / rstd are used)
  int nlayers;              // contributing layers (0: the range only holds T)
  const float* gamma[25];   // per contributor (the reading layers' norm1, + the transition's BatchNorm): gamma, and the g_sum / g_dot
  const double* g_sum[25];  // of its backward statistics (all indexed from channel 0 of the block buffer)
  const double* g_dot[25];
  float inv_count;
  int batch;                // 0: eval mode
  bf16* out;                // optional bf16 copy [M][out_pitch] (the next GEMM's gradient operand), x out_scale[sample][nch]
  long long out_pitch;
  const float* out_scale;
  int vps;
  int write_back;           // store the final fp32 gradient back into G (needed when someone reads G afterwards: the stem's max-pool
                            // backward, GradCAM); a slice that only feeds the next GEMM through `out` skips the 4 B/element store
};

static __global__ void __launch_bounds__(EW_THREADS) grad_finalize_kernel(const __grid_constant__ FinalizeParams p) {
  pdl_wait(); pdl_trigger();
  extern __shared__ float coef[];   // [3][nch]: a (on G), b (on x), d;  then double part[2][T][nch] (reduction scratch)
  // C1 / C2: nlayers x nch independent loads.  T threads share a channel (layers j, j+T, ...), four loads in flight per thread,
  // partial sums combined in a fixed order -- the serial per-channel loop cost ~0.5 us of L2 latency per contributing layer in
  // EVERY block of the grid (24 layers: ~10 us per launch).
  int T = EW_THREADS / p.nch;
  T = T < 1 ? 1 : (T > 8 ? 8 : T);
  double* part = reinterpret_cast<double*>(coef + 3 * p.nch + ((3 * p.nch) & 1));
  if (p.batch) {
    for (int i = threadIdx.x; i < p.nch * T; i += EW_THREADS) {
      const int c = i % p.nch, j = i / p.nch;
      double a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
      int l = j;
      for (; l + 3 * T < p.nlayers; l += 4 * T) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double ga = (double)p.gamma[l + u * T][p.c_lo + c];
          a1[u] += ga * p.g_sum[l + u * T][p.c_lo + c];
          a2[u] += ga * p.g_dot[l + u * T][p.c_lo + c];
        }
      }
      for (; l < p.nlayers; l += T) {
        const double ga = (double)p.gamma[l][p.c_lo + c];
        a1[0] += ga * p.g_sum[l][p.c_lo + c];
        a2[0] += ga * p.g_dot[l][p.c_lo + c];
      }
      part[(0 * T + j) * p.nch + c] = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      part[(1 * T + j) * p.nch + c] = (a2[0] + a2[1]) + (a2[2] + a2[3]);
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < p.nch; c += EW_THREADS) {
    float a = 1.f, b = 0.f, d = 0.f;
    if (p.batch) {
      float mean, rstd;
      bn_mean_rstd(p.bn, p.c_lo + c, mean, rstd);
      double C1 = 0.0, C2 = 0.0;
      for (int j = 0; j < T; ++j) { C1 += part[(0 * T + j) * p.nch + c]; C2 += part[(1 * T + j) * p.nch + c]; }
      const float c1 = (float)(C1 * (double)p.inv_count), c2 = (float)(C2 * (double)p.inv_count);
      a = rstd; b = -rstd * rstd * c2; d = rstd * (-c1 + mean * rstd * c2);
    }
    coef[c] = a; coef[p.nch + c] = b; coef[2 * p.nch + c] = d;
  }
  __syncthreads();
  const int cpr = p.nch / 8;
  const long long total = p.M * cpr;
  for (long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x; idx < total; idx += (long long)gridDim.x * EW_THREADS) {
    const int chunk = (int)(idx % cpr);
    const long long m = idx / cpr;
    float4* gp = reinterpret_cast<float4*>(p.G + m * p.g_pitch + p.c_lo + chunk * 8);
    const float4 ga = gp[0], gb = gp[1];
    float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    if (p.batch) {
      float f[8];
      unpack8<ACT>(ldg16(p.x + m * p.x_pitch + p.c_lo + chunk * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = chunk * 8 + e;
        g[e] = fmaf(coef[c], g[e], fmaf(coef[p.nch + c], f[e], coef[2 * p.nch + c]));
      }
      if (p.write_back) {
        gp[0] = make_float4(g[0], g[1], g[2], g[3]);
        gp[1] = make_float4(g[4], g[5], g[6], g[7]);
      }
    }
    if (p.out != nullptr) {
      if (p.out_scale != nullptr) {
        const float* cs = p.out_scale + (m / p.vps) * p.nch + chunk * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] *= __ldg(cs + e);
      }
      *reinterpret_cast<uint4*>(p.out + m * p.out_pitch + chunk * 8) = pack8<GRD>(g);
    }
  }
}

// ------------------------------------------------------------------------------------------------- E5: slice extraction
// dst bf16 [M][C] = src fp32 [M][pitch](channels [0,C) from the given pointer) * colscale[sample][C]
static __global__ void __launch_bounds__(EW_THREADS) extract_slice_kernel(const float* __restrict__ src, long long src_pitch,
                                                                   bf16* __restrict__ dst, long long M, int C,
                                                                   const float* __restrict__ colscale, int vps) {
  pdl_wait(); pdl_trigger();   // launched with launch_pdl (launch.h)
  const int cpr = C / 8;
  const long long total = M * cpr;
  for (long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x; idx < total; idx += (long long)gridDim.x * EW_THREADS) {
    const int chunk = (int)(idx % cpr);
    const long long m = idx / cpr;
    const float4 a = *reinterpret_cast<const float4*>(src + m * src_pitch + chunk * 8);
    const float4 b = *reinterpret_cast<const float4*>(src + m * src_pitch + chunk * 8 + 4);
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (colscale != nullptr) {
      const float* cs = colscale + (m / vps) * C + chunk * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] *= __ldg(cs + e);
    }
    *reinterpret_cast<uint4*>(dst + m * C + chunk * 8) = pack8<GRD>(f);
  }
}

// ------------------------------------------------------------------------------------------------- E6: norm5
// forward: y fp32 [M][C] = bn(x);  backward pass 1: statistics of (dy, dy*xhat) from the fp32 upstream gradient
static __global__ void __launch_bounds__(EW_THREADS) bn_apply_f32_kernel(const bf16* __restrict__ x, long long x_pitch, BnSrc bn,
                                                                  float* __restrict__ y, long long M, int C) {
  extern __shared__ float coef[];
  for (int c = threadIdx.x; c < C; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(bn, c, mean, rstd);
    const float s = bn.gamma[c] * rstd;
    coef[c] = s; coef[C + c] = bn.beta[c] - mean * s;
  }
  __syncthreads();
  const int cpr = C / 8;
  const long long total = M * cpr;
  for (long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x; idx < total; idx += (long long)gridDim.x * EW_THREADS) {
    const int chunk = (int)(idx % cpr);
    const long long m = idx / cpr;
    float f[8];
    unpack8<ACT>(ldg16(x + m * x_pitch + chunk * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], coef[chunk * 8 + e], coef[C + chunk * 8 + e]);
    float4* d = reinterpret_cast<float4*>(y + m * C + chunk * 8);
    d[0] = make_float4(f[0], f[1], f[2], f[3]);
    d[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

static __global__ void __launch_bounds__(EW_THREADS) bn_bwd_stats_f32_kernel(const float* __restrict__ dy, const bf16* __restrict__ x,
                                                                      long long x_pitch, BnSrc bn, long long M, int C,
                                                                      double* g_sum, double* g_dot) {
  extern __shared__ float coef[];  // mean, rstd
  __shared__ float red[EW_THREADS * 16];
  for (int c = threadIdx.x; c < C; c += EW_THREADS) {
    float mean, rstd;
    bn_mean_rstd(bn, c, mean, rstd);
    coef[c] = mean; coef[C + c] = rstd;
  }
  __syncthreads();
  const int cpr = C / 8;
  const int rows_per_block = EW_THREADS / cpr;
  const int chunk = threadIdx.x % cpr;
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long m = (long long)blockIdx.x * rows_per_block + threadIdx.x / cpr; m < M && threadIdx.x < rows_per_block * cpr;
       m += (long long)gridDim.x * rows_per_block) {
    const float4 a = *reinterpret_cast<const float4*>(dy + m * C + chunk * 8);
    const float4 b = *reinterpret_cast<const float4*>(dy + m * C + chunk * 8 + 4);
    const float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float f[8];
    unpack8<ACT>(ldg16(x + m * x_pitch + chunk * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s1[e] += g[e];
      s2[e] += g[e] * (f[e] - coef[chunk * 8 + e]) * coef[C + chunk * 8 + e];
    }
  }
  block_channel_reduce(red, s1, s2, cpr, g_sum, g_dot);
}

// ------------------------------------------------------------------------------------------------- E8-E10: table-driven tails
struct BnTableEntry {
  const double* sum;     // forward batch statistics of this BN's input channels
  const double* sumsq;
  const double* g_sum;   // backward statistics
  const double* g_dot;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  long long grad_gamma_off;   // element offsets from the gradient base pointer passed to the kernel: with the
  long long grad_beta_off;    // gradients laid out in one flat buffer the table is identical every step (graph-safe)
  int C;
  float count;
};

// running_mean/var momentum update (momentum 0.1, unbiased variance) for every BN of the trunk in one launch
static __global__ void bn_running_update_kernel(const BnTableEntry* __restrict__ tab, float momentum) {
  const BnTableEntry e = tab[blockIdx.x];
  for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
    const double mean = e.sum[c] / (double)e.count;
    double var = e.sumsq[c] / (double)e.count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const double unbiased = e.count > 1.f ? var * (double)e.count / ((double)e.count - 1.0) : var;
    e.running_mean[c] = (1.f - momentum) * e.running_mean[c] + momentum * (float)mean;
    e.running_var[c] = (1.f - momentum) * e.running_var[c] + momentum * (float)unbiased;
  }
  if (threadIdx.x == 0 && e.num_batches_tracked != nullptr) *e.num_batches_tracked += 1;
}

// dgamma = sum dy*xhat, dbeta = sum dy
static __global__ void bn_param_grad_kernel(const BnTableEntry* __restrict__ tab, float* __restrict__ gbase) {
  const BnTableEntry e = tab[blockIdx.x];
  for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
    gbase[e.grad_gamma_off + c] = (float)e.g_dot[c];
    gbase[e.grad_beta_off + c] = (float)e.g_sum[c];
  }
}

// Ordered reduction of the weight-gradient slots (engine.cuh, WgradParams::slot_stride): dst[i] = sum over s = 0 .. S-1, IN THAT
// ORDER, of src[s * numel + i] -- the same bits every run, whatever order the CTAs of the voxel split finished in.  The 3x3x3
// gradients are accumulated lane-contiguously as [tap][co][ci] and written out here as the reference's [co][ci][tap]
// (taps = 27, coci = co * ci); taps = 0: same layout on both sides.  HBM-bound: S * numel floats read once, numel written.
struct WReduceEntry {
  const float* src;
  long long dst_off;   // element offset from the gradient base pointer
  int numel;
  int S;
  int taps;
  int coci;
};
static __global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WReduceEntry* __restrict__ tab, float* __restrict__ gbase) {
  const WReduceEntry e = tab[blockIdx.y];
  const int nvec = e.numel >> 2;                      // numel is a multiple of 4 for every convolution of the trunk
  for (int iv = blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += gridDim.x * blockDim.x) {
    const float4* sp = reinterpret_cast<const float4*>(e.src) + iv;
    float4 acc = __ldcs(sp);
    for (int s = 1; s < e.S; ++s) {
      const float4 v = __ldcs(sp + (size_t)s * nvec);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int i = iv << 2;
    if (e.taps == 0) {
      *reinterpret_cast<float4*>(gbase + e.dst_off + i) = acc;
    } else {
      const int tap = i / e.coci, rest = i - tap * e.coci;   // 4 consecutive (co, ci) cells of one tap
      float* d = gbase + e.dst_off + (long long)rest * e.taps + tap;
      d[0] = acc.x; d[e.taps] = acc.y; d[2 * e.taps] = acc.z; d[3 * e.taps] = acc.w;
    }
  }
}

}  // namespace mmnn
