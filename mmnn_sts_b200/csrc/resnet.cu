// The 3-D ResNet encoder of BASELINE configs[3]: /root/reference/models/resnet.py -- BasicStem
// (Conv3d(1 -> 64, k (1,7,7), s (1,2,2), p (1,3,3)) + BN + ReLU, :5-13), BasicBlock (two Conv3DSimple 3x3x3 + BN, residual
// add, ReLU, :61-95) with the 1x1x1 strided down-sample branch (:172-179), planes 8 / 16 / 8 / 16 (:134-137), element-wise
// Dropout after every stage (:156-163), AdaptiveAvgPool3d(1) + Linear + sigmoid (:165-170) -- forward and backward.
//
// Every layer behind the stem has 8 or 16 channels (80 706 parameters in total): there is no dense contraction worth a
// tcgen05 tile (a 128 x 8 MMA would be bound by re-reading A), the activations are what costs: the
// stem output alone is 8 x 258 x 128 x 32 x 64 x 2 B = 1.08 GB at configs[3].  So these are direct convolutions on
// channels-last (NDHWC) 16-bit tensors (activations fp16, gradients bf16): one thread per output voxel x CT output channels with the weight slice of the
// block resident in shared memory (fp32, broadcast float4 reads), fp32 accumulate, fused per-channel batch statistics
// of the STORED (rounded) output for the BatchNorm that follows; the data gradient is the same kernel run as the
// transposed gather; the weight gradient keeps 8 accumulators per (tap, c_in, c_out-group) work item in registers
// across a run of voxels and finishes with fp32 atomics.  BatchNorm / residual / ReLU / Dropout are fused element-wise
// passes (128-bit accesses, HBM-bound).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "prof.h"

namespace mmnn {
namespace rn {

struct RnConvGeom {
  int N, Di, Hi, Wi, Cin, Do, Ho, Wo, Cout;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
};

constexpr int THREADS = 256;

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ void unpack8(const uint4 q, float* f) {
  f[0] = bf_lo(q.x); f[1] = bf_hi(q.x); f[2] = bf_lo(q.y); f[3] = bf_hi(q.y);
  f[4] = bf_lo(q.z); f[5] = bf_hi(q.z); f[6] = bf_lo(q.w); f[7] = bf_hi(q.w);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
__device__ __forceinline__ float round_bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// forward activations are IEEE fp16 (3 more mantissa bits than bf16; every tensor here is O(1..100)), gradients bf16
__device__ __forceinline__ void unpack8h(const uint4 q, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8h(const float* f) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return q;
}
__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }
template <bool BF>
__device__ __forceinline__ void unpack8t(const uint4 q, float* f) { if (BF) unpack8(q, f); else unpack8h(q, f); }
template <bool BF>
__device__ __forceinline__ uint4 pack8t(const float* f) { return BF ? pack8(f) : pack8h(f); }

// ------------------------------------------------------------------------------------------------ direct convolution
// DGRAD == false: dst[n,od,oh,ow,co] = sum_{tap,ci} src[n, o*s - p + tap, ci] * w[co][ci][tap]      (nn.Conv3d forward)
// DGRAD == true : dst[n,id,ih,iw,ci] = sum_{tap,co} src[n, (i + p - tap)/s, co] * w[co][ci][tap]    (its data gradient)
// KC = channels of `src` (the reduction), CT = channels of `dst` one thread owns; blockIdx.y = channel group of dst.
template <typename TIN, int KC, int CT, bool DGRAD>
__global__ void __launch_bounds__(THREADS) rn_conv_kernel(const RnConvGeom g, const TIN* __restrict__ src,
                                                          const float* __restrict__ w, uint16_t* __restrict__ dst,
                                                          const uint16_t* add, double* __restrict__ stats) {
  extern __shared__ float ws[];                       // [taps][KC][CT] then [2][CT] block statistics
  const int taps = g.kd * g.kh * g.kw;
  const int cg = blockIdx.y;
  const int tid = threadIdx.x;
  float* sstat = ws + taps * KC * CT;
  for (int i = tid; i < taps * KC * CT; i += THREADS) {
    const int c = i % CT, k = (i / CT) % KC, t = i / (CT * KC);
    const int co = DGRAD ? k : cg * CT + c;
    const int ci = DGRAD ? cg * CT + c : k;
    ws[i] = w[((long long)co * g.Cin + ci) * taps + t];
  }
  if (tid < 2 * CT) sstat[tid] = 0.f;
  __syncthreads();

  const int OD = DGRAD ? g.Di : g.Do, OH = DGRAD ? g.Hi : g.Ho, OW = DGRAD ? g.Wi : g.Wo;   // dst extent
  const int ID = DGRAD ? g.Do : g.Di, IH = DGRAD ? g.Ho : g.Hi, IW = DGRAD ? g.Wo : g.Wi;   // src extent
  const int OC = DGRAD ? g.Cin : g.Cout;
  const long long total = (long long)g.N * OD * OH * OW;
  float ssum[CT], ssq[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }

  for (long long v = (long long)blockIdx.x * THREADS + tid; v < total; v += (long long)gridDim.x * THREADS) {
    const int ow = (int)(v % OW);
    long long t = v / OW;
    const int oh = (int)(t % OH);
    t /= OH;
    const int od = (int)(t % OD);
    const int n = (int)(t / OD);
    float acc[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[c] = 0.f;
    for (int a = 0; a < g.kd; ++a) {
      int zd;
      if (!DGRAD) { zd = od * g.sd - g.pd + a; if ((unsigned)zd >= (unsigned)ID) continue; }
      else { const int q = od + g.pd - a; if (q < 0) continue; zd = q / g.sd; if (zd * g.sd != q || zd >= ID) continue; }
      for (int b = 0; b < g.kh; ++b) {
        int zh;
        if (!DGRAD) { zh = oh * g.sh - g.ph + b; if ((unsigned)zh >= (unsigned)IH) continue; }
        else { const int q = oh + g.ph - b; if (q < 0) continue; zh = q / g.sh; if (zh * g.sh != q || zh >= IH) continue; }
        for (int c = 0; c < g.kw; ++c) {
          int zw;
          if (!DGRAD) { zw = ow * g.sw - g.pw + c; if ((unsigned)zw >= (unsigned)IW) continue; }
          else { const int q = ow + g.pw - c; if (q < 0) continue; zw = q / g.sw; if (zw * g.sw != q || zw >= IW) continue; }
          const TIN* px = src + ((((long long)n * ID + zd) * IH + zh) * IW + zw) * KC;
          const float* pw = ws + ((a * g.kh + b) * g.kw + c) * (KC * CT);
          if constexpr (std::is_same<TIN, float>::value) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
              const float xv = __ldg(px + k);
#pragma unroll
              for (int j = 0; j < CT; j += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(pw + k * CT + j);
                acc[j] = fmaf(xv, wv.x, acc[j]); acc[j + 1] = fmaf(xv, wv.y, acc[j + 1]);
                acc[j + 2] = fmaf(xv, wv.z, acc[j + 2]); acc[j + 3] = fmaf(xv, wv.w, acc[j + 3]);
              }
            }
          } else {
#pragma unroll
            for (int k0 = 0; k0 < KC; k0 += 8) {
              float xf[8];
              unpack8t<DGRAD>(__ldg(reinterpret_cast<const uint4*>(px + k0)), xf);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int j = 0; j < CT; j += 4) {
                  const float4 wv = *reinterpret_cast<const float4*>(pw + (k0 + k) * CT + j);
                  acc[j] = fmaf(xf[k], wv.x, acc[j]); acc[j + 1] = fmaf(xf[k], wv.y, acc[j + 1]);
                  acc[j + 2] = fmaf(xf[k], wv.z, acc[j + 2]); acc[j + 3] = fmaf(xf[k], wv.w, acc[j + 3]);
                }
              }
            }
          }
        }
      }
    }
    const long long o = v * OC + cg * CT;
    if (add != nullptr) {
#pragma unroll
      for (int j = 0; j < CT; j += 8) {
        float af[8];
        unpack8t<DGRAD>(*reinterpret_cast<const uint4*>(add + o + j), af);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[j + k] += af[k];
      }
    }
#pragma unroll
    for (int j = 0; j < CT; j += 8) *reinterpret_cast<uint4*>(dst + o + j) = pack8t<DGRAD>(acc + j);
    if (stats != nullptr) {
#pragma unroll
      for (int c = 0; c < CT; ++c) { const float r = DGRAD ? round_bf(acc[c]) : round_h(acc[c]); ssum[c] += r; ssq[c] = fmaf(r, r, ssq[c]); }
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      float s = ssum[c], q = ssq[c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
      }
      if ((tid & 31) == 0) { atomicAdd(&sstat[c], s); atomicAdd(&sstat[CT + c], q); }
    }
    __syncthreads();
    if (tid < CT) atomicAdd(&stats[cg * CT + tid], (double)sstat[tid]);
    else if (tid < 2 * CT) atomicAdd(&stats[OC + cg * CT + (tid - CT)], (double)sstat[tid]);
  }
}

template <typename TIN, int KC, int CT, bool DGRAD>
static int launch_conv(const RnConvGeom& g, const void* src, const float* w, void* dst, const void* add, double* stats,
                       cudaStream_t st) {
  const int taps = g.kd * g.kh * g.kw;
  const size_t smem = ((size_t)taps * KC * CT + 2 * CT) * sizeof(float);
  auto kern = rn_conv_kernel<TIN, KC, CT, DGRAD>;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  const long long total = DGRAD ? (long long)g.N * g.Di * g.Hi * g.Wi : (long long)g.N * g.Do * g.Ho * g.Wo;
  const int OC = DGRAD ? g.Cin : g.Cout;
  long long bx = (total + THREADS - 1) / THREADS;
  const long long cap = 148LL * 8;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  kern<<<dim3((unsigned)bx, (unsigned)(OC / CT)), THREADS, smem, st>>>(g, (const TIN*)src, w, (uint16_t*)dst,
                                                                      (const uint16_t*)add, stats);
  return (int)cudaGetLastError();
}

template <bool DGRAD>
static int dispatch_conv(const RnConvGeom& g, int src_f32, const void* src, const float* w, void* dst, const void* add,
                         double* stats, cudaStream_t st) {
  const int KC = DGRAD ? g.Cout : g.Cin, OC = DGRAD ? g.Cin : g.Cout;
  if (OC % 8 != 0) return -2;
  const bool c16 = (OC % 16 == 0);
  if (src_f32) {
    if (KC != 1 || DGRAD) return -3;
    if (c16) return launch_conv<float, 1, 16, false>(g, src, w, dst, add, stats, st);
    return launch_conv<float, 1, 8, false>(g, src, w, dst, add, stats, st);
  }
#define RN_CASE(kc)                                                                                          \
  if (KC == kc) {                                                                                            \
    if (c16) return launch_conv<uint16_t, kc, 16, DGRAD>(g, src, w, dst, add, stats, st);               \
    return launch_conv<uint16_t, kc, 8, DGRAD>(g, src, w, dst, add, stats, st);                         \
  }
  RN_CASE(8)
  RN_CASE(16)
  RN_CASE(64)
#undef RN_CASE
  return -4;
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dw[co][ci][tap] += sum_voxels dy[n,o,co] * x[n, o*s - p + tap, ci].  Tile = TW consecutive output voxels of one output row
// (n, od, oh): the kd x kh input rows the tile touches ((TW-1) sw + kw voxels each, zero-filled outside the volume) and the
// tile's dy are staged in shared memory, so the inner loop has no bounds checks and only fixed-latency operands.  Work
// item = (tap, ci, group of 8 co); a thread owns up to MAXI items (8 fp32 accumulators each) across all tiles of its block
// (grid-stride), and finishes with fp32 atomics.
constexpr int WG_TW = 32;

template <typename TX>
__device__ __forceinline__ float wg_ld(const TX* p) {
  if constexpr (std::is_same<TX, float>::value) return *p;
  else return __half2float(*p);
}

template <typename TX, int MAXI>
__global__ void __launch_bounds__(THREADS, (MAXI > 2 ? 2 : 3)) rn_wgrad_kernel(const RnConvGeom g, const TX* __restrict__ x,
                                                                              const uint16_t* __restrict__ dy,
                                                                              float* __restrict__ dw, int tiles_w,
                                                                              long long ntiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int taps = g.kd * g.kh * g.kw;
  const int items = taps * g.Cin * (g.Cout / 8);
  const int tid = threadIdx.x;
  const int XW = (WG_TW - 1) * g.sw + g.kw;               // staged voxels per input row
  const int rowlen = XW * g.Cin;                          // elements per staged input row
  float* dy_s = reinterpret_cast<float*>(smem_raw);       // [WG_TW][Cout] fp32
  TX* x_s = reinterpret_cast<TX*>(smem_raw + (size_t)WG_TW * g.Cout * sizeof(float));   // [kd*kh][XW][Cin]
  int xo[MAXI], icog[MAXI];
  float acc[MAXI][8];
#pragma unroll
  for (int j = 0; j < MAXI; ++j) {
    const int it = tid + j * THREADS;
    const int ci = it % g.Cin, rest = it / g.Cin;
    const int tap = rest % taps;
    icog[j] = (it < items) ? rest / taps : -1;
    const int a = tap / (g.kh * g.kw), b = (tap / g.kw) % g.kh, c = tap % g.kw;
    xo[j] = ((a * g.kh + b) * XW + c) * g.Cin + ci;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
  }
  const bool one = (g.Cout == 8);
  const int step = g.sw * g.Cin;
  constexpr int VEC = 16 / (int)sizeof(TX);               // elements per 16-byte copy
  const bool vec_ok = (g.Cin % VEC) == 0;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    long long t = tile / tiles_w;
    const int oh = (int)(t % g.Ho);
    t /= g.Ho;
    const int od = (int)(t % g.Do);
    const int n = (int)(t / g.Do);
    const int ow0 = tw * WG_TW;
    __syncthreads();                                      // previous tile fully consumed
    // ---- stage dy (bf16 -> fp32; voxels past the row end are zero)
    for (int i = tid; i < WG_TW * g.Cout / 8; i += THREADS) {
      const int v = i / (g.Cout / 8), c8 = i % (g.Cout / 8);
      float d[8];
      if (ow0 + v < g.Wo) {
        const long long o = ((((long long)n * g.Do + od) * g.Ho + oh) * g.Wo + ow0 + v) * g.Cout + c8 * 8;
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy + o)), d);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(dy_s + v * g.Cout + c8 * 8);
      dst[0] = make_float4(d[0], d[1], d[2], d[3]);
      dst[1] = make_float4(d[4], d[5], d[6], d[7]);
    }
    // ---- stage the kd x kh input rows
    const int iw0 = ow0 * g.sw - g.pw;
    const int nrows = g.kd * g.kh;
    if (vec_ok) {
      const int per_row = rowlen / VEC;
      for (int i = tid; i < nrows * per_row; i += THREADS) {
        const int r = i / per_row, e = (i % per_row) * VEC;
        const int zd = od * g.sd - g.pd + r / g.kh, zh = oh * g.sh - g.ph + r % g.kh;
        const int zw = iw0 + e / g.Cin;
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)zd < (unsigned)g.Di && (unsigned)zh < (unsigned)g.Hi && (unsigned)zw < (unsigned)g.Wi)
          q = __ldg(reinterpret_cast<const uint4*>(x + ((((long long)n * g.Di + zd) * g.Hi + zh) * g.Wi + zw) * g.Cin + e % g.Cin));
        *reinterpret_cast<uint4*>(x_s + (long long)r * rowlen + e) = q;
      }
    } else {
      for (int i = tid; i < nrows * rowlen; i += THREADS) {
        const int r = i / rowlen, e = i % rowlen;
        const int zd = od * g.sd - g.pd + r / g.kh, zh = oh * g.sh - g.ph + r % g.kh;
        const int zw = iw0 + e / g.Cin;
        TX q = TX(0);
        if ((unsigned)zd < (unsigned)g.Di && (unsigned)zh < (unsigned)g.Hi && (unsigned)zw < (unsigned)g.Wi)
          q = x[((((long long)n * g.Di + zd) * g.Hi + zh) * g.Wi + zw) * g.Cin + e % g.Cin];
        x_s[(long long)r * rowlen + e] = q;
      }
    }
    __syncthreads();
    // ---- accumulate
#pragma unroll 4
    for (int v = 0; v < WG_TW; ++v) {
      float d[8];
      if (one) {
        const float4 d0 = *reinterpret_cast<const float4*>(dy_s + v * 8), d1 = *reinterpret_cast<const float4*>(dy_s + v * 8 + 4);
        d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
      }
#pragma unroll
      for (int j = 0; j < MAXI; ++j) {
        if (icog[j] < 0) continue;
        const float xv = wg_ld(x_s + xo[j] + v * step);
        if (!one) {
          const float* pd = dy_s + v * g.Cout + icog[j] * 8;
          const float4 d0 = *reinterpret_cast<const float4*>(pd), d1 = *reinterpret_cast<const float4*>(pd + 4);
          d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[j][k] = fmaf(xv, d[k], acc[j][k]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXI; ++j) {
    if (icog[j] < 0) continue;
    const int it = tid + j * THREADS;
    const int ci = it % g.Cin, tap = (it / g.Cin) % taps;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      atomicAdd(&dw[((long long)(icog[j] * 8 + k) * g.Cin + ci) * taps + tap], acc[j][k]);
  }
}

template <typename TX, int MAXI>
static int launch_wgrad(const RnConvGeom& g, const void* x, const void* dy, float* dw, cudaStream_t st) {
  const int tiles_w = (g.Wo + WG_TW - 1) / WG_TW;
  const long long ntiles = (long long)g.N * g.Do * g.Ho * tiles_w;
  const int XW = (WG_TW - 1) * g.sw + g.kw;
  const size_t smem = (size_t)WG_TW * g.Cout * sizeof(float) + (size_t)g.kd * g.kh * XW * g.Cin * sizeof(TX);
  auto kern = rn_wgrad_kernel<TX, MAXI>;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  long long blocks = ntiles;
  const long long cap = 148LL * (MAXI > 2 ? 2 : 3);
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, THREADS, smem, st>>>(g, (const TX*)x, (const uint16_t*)dy, dw, tiles_w, ntiles);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ BatchNorm coefficients
// coef [4][C]: scale = gamma * rstd, shift = beta - mean * scale, mean, rstd.  Training: batch statistics (biased variance
// for the normalisation, unbiased for the running update, momentum 0.1 -- torch defaults).
__global__ void rn_bn_coeffs_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float* rmean, float* rvar, long long* nbt, float eps,
                                    float momentum, int training, int C, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  double mean, var;
  if (training) {
    mean = stats[c] / count;
    var = stats[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    rmean[c] = (float)((1.0 - (double)momentum) * (double)rmean[c] + (double)momentum * mean);
    rvar[c] = (float)((1.0 - (double)momentum) * (double)rvar[c] + (double)momentum * unb);
  } else {
    mean = (double)rmean[c];
    var = (double)rvar[c];
  }
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  coef[c] = sc;
  coef[C + c] = beta[c] - (float)mean * sc;
  coef[2 * C + c] = (float)mean;
  coef[3 * C + c] = rstd;
}

// ------------------------------------------------------------------------------------------------ element-wise passes
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long e) {
  unsigned long long z = seed + e * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// y = [relu]( raw * scale + shift  [+ res | + raw2 * scale2 + shift2] )  [ * keep / (1 - p) ]
// RES 0: none, 1: identity tensor, 2: second BatchNorm'ed raw tensor (the down-sample branch).
template <int RES>
__global__ void __launch_bounds__(THREADS) rn_bn_act_kernel(const uint4* __restrict__ raw, const float* __restrict__ coef,
                                                            const uint4* __restrict__ res, const float* __restrict__ coef2,
                                                            uint4* __restrict__ y, long long n8, int C, int relu,
                                                            float drop_p, unsigned long long seed,
                                                            const unsigned char* __restrict__ mask) {
  const int c0 = (threadIdx.x * 8) % C;
  float sc[8], sh[8], sc2[8], sh2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = coef[c0 + k]; sh[k] = coef[C + c0 + k];
    if (RES == 2) { sc2[k] = coef2[c0 + k]; sh2[k] = coef2[C + c0 + k]; }
  }
  const bool drop = drop_p > 0.f || mask != nullptr;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n8; i += (long long)gridDim.x * THREADS) {
    float v[8], r[8];
    unpack8h(__ldg(raw + i), v);
    if (RES != 0) unpack8h(__ldg(res + i), r);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float o = fmaf(v[k], sc[k], sh[k]);
      if (RES == 1) o += r[k];
      if (RES == 2) o += fmaf(r[k], sc2[k], sh2[k]);
      if (relu) o = fmaxf(o, 0.f);
      if (drop) {
        const unsigned long long e = (unsigned long long)i * 8ull + k;
        const bool keep = mask != nullptr ? (mask[e] != 0) : (hash_uniform(seed, e) >= drop_p);
        o = keep ? o * keep_scale : 0.f;
      }
      v[k] = o;
    }
    y[i] = pack8h(v);
  }
}

// dz = dy * [y > 0] * post_scale.  sums [3][C] (fp64): sum dz, sum dz * xhat(raw), sum dz * xhat(raw2) (TWO only).
template <bool TWO>
__global__ void __launch_bounds__(THREADS) rn_act_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y,
                                                                    float post_scale, const uint4* __restrict__ raw,
                                                                    const float* __restrict__ coef,
                                                                    const uint4* __restrict__ raw2,
                                                                    const float* __restrict__ coef2, double* __restrict__ sums,
                                                                    long long n8, int C) {
  __shared__ float sm[3 * 64];
  for (int i = threadIdx.x; i < 3 * C; i += THREADS) sm[i] = 0.f;
  __syncthreads();
  const int c0 = (threadIdx.x * 8) % C;
  float mean[8], rstd[8], mean2[8], rstd2[8], s0[8], s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = coef[2 * C + c0 + k]; rstd[k] = coef[3 * C + c0 + k];
    if (TWO) { mean2[k] = coef2[2 * C + c0 + k]; rstd2[k] = coef2[3 * C + c0 + k]; }
    s0[k] = 0.f; s1[k] = 0.f; s2[k] = 0.f;
  }
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n8; i += (long long)gridDim.x * THREADS) {
    float g[8], yy[8], x[8], x2[8];
    unpack8(__ldg(dy + i), g);
    unpack8h(__ldg(y + i), yy);
    unpack8h(__ldg(raw + i), x);
    if (TWO) unpack8h(__ldg(raw2 + i), x2);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float dz = yy[k] > 0.f ? g[k] * post_scale : 0.f;
      s0[k] += dz;
      s1[k] = fmaf(dz, (x[k] - mean[k]) * rstd[k], s1[k]);
      if (TWO) s2[k] = fmaf(dz, (x2[k] - mean2[k]) * rstd2[k], s2[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(&sm[c0 + k], s0[k]);
    atomicAdd(&sm[C + c0 + k], s1[k]);
    if (TWO) atomicAdd(&sm[2 * C + c0 + k], s2[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (TWO ? 3 : 2) * C; i += THREADS) atomicAdd(&sums[i], (double)sm[i]);
}

// draw = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))  (BatchNorm backward, training mode); eval_mode: draw =
// gamma * rstd * dz.  Optional second branch (down-sample BatchNorm) and optional dz output (identity residual).
template <bool TWO>
__global__ void __launch_bounds__(THREADS) rn_bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y,
                                                                  float post_scale, const uint4* __restrict__ raw,
                                                                  const float* __restrict__ coef, const uint4* __restrict__ raw2,
                                                                  const float* __restrict__ coef2,
                                                                  const double* __restrict__ sums, double inv_count,
                                                                  int eval_mode, uint4* __restrict__ draw,
                                                                  uint4* __restrict__ draw2, uint4* __restrict__ dzout,
                                                                  float* dgamma, float* dbeta, float* dgamma2, float* dbeta2,
                                                                  long long n8, int C) {
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += THREADS) {
      dbeta[c] = (float)sums[c];
      dgamma[c] = (float)sums[C + c];
      if (TWO) { dbeta2[c] = (float)sums[c]; dgamma2[c] = (float)sums[2 * C + c]; }
    }
  }
  const int c0 = (threadIdx.x * 8) % C;
  float mean[8], rstd[8], sc[8], m1[8], m2[8], mean2[8], rstd2[8], sc2[8], m22[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = coef[c0 + k]; mean[k] = coef[2 * C + c0 + k]; rstd[k] = coef[3 * C + c0 + k];
    m1[k] = eval_mode ? 0.f : (float)(sums[c0 + k] * inv_count);
    m2[k] = eval_mode ? 0.f : (float)(sums[C + c0 + k] * inv_count);
    if (TWO) {
      sc2[k] = coef2[c0 + k]; mean2[k] = coef2[2 * C + c0 + k]; rstd2[k] = coef2[3 * C + c0 + k];
      m22[k] = eval_mode ? 0.f : (float)(sums[2 * C + c0 + k] * inv_count);
    }
  }
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n8; i += (long long)gridDim.x * THREADS) {
    float g[8], yy[8], x[8], x2[8], o[8], o2[8];
    unpack8(__ldg(dy + i), g);
    unpack8h(__ldg(y + i), yy);
    unpack8h(__ldg(raw + i), x);
    if (TWO) unpack8h(__ldg(raw2 + i), x2);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float dz = yy[k] > 0.f ? g[k] * post_scale : 0.f;
      g[k] = dz;
      o[k] = sc[k] * (dz - m1[k] - (x[k] - mean[k]) * rstd[k] * m2[k]);
      if (TWO) o2[k] = sc2[k] * (dz - m1[k] - (x2[k] - mean2[k]) * rstd2[k] * m22[k]);
    }
    draw[i] = pack8(o);
    if (TWO) draw2[i] = pack8(o2);
    if (dzout != nullptr) dzout[i] = pack8(g);
  }
}

// ------------------------------------------------------------------------------------------------ head
// AdaptiveAvgPool3d(1) -> flatten -> Linear(C -> K) -> sigmoid (/root/reference/models/resnet.py:165-170); one block per sample.
__global__ void __launch_bounds__(128) rn_head_fwd_kernel(const __half* __restrict__ y, int V, int C,
                                                          const float* __restrict__ W, const float* __restrict__ bias, int K,
                                                          float* __restrict__ pooled, float* __restrict__ out) {
  __shared__ float sp[64];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < C; i += 128) sp[i] = 0.f;
  __syncthreads();
  const __half* py = y + (long long)b * V * C;
  // thread owns channel (tid % C) when 128 % C == 0 (C = 8 / 16 / 64)
  const int c = threadIdx.x % C;
  float s = 0.f;
  for (long long i = threadIdx.x; i < (long long)V * C; i += 128) s += __half2float(py[i]);
  atomicAdd(&sp[c], s);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 128) { sp[i] *= 1.f / (float)V; pooled[b * C + i] = sp[i]; }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += 128) {
    float z = bias[k];
    for (int i = 0; i < C; ++i) z = fmaf(W[k * C + i], sp[i], z);
    out[b * K + k] = 1.f / (1.f + expf(-z));
  }
}

// dlogit = dout * out * (1 - out); dW += dlogit^T pooled; db += sum dlogit; dy[b][v][c] = (dlogit W)[b][c] / V
__global__ void __launch_bounds__(128) rn_head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                          const float* __restrict__ pooled, const float* __restrict__ W, int V,
                                                          int C, int K, __nv_bfloat16* __restrict__ dy, float* dW, float* db) {
  __shared__ float sdl[64];
  __shared__ float sdp[64];
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += 128) {
    const float o = out[b * K + k];
    const float dl = dout[b * K + k] * o * (1.f - o);
    sdl[k] = dl;
    atomicAdd(&db[k], dl);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += 128) atomicAdd(&dW[i], sdl[i / C] * pooled[b * C + (i % C)]);
  for (int c = threadIdx.x; c < C; c += 128) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(sdl[k], W[k * C + c], s);
    sdp[c] = s / (float)V;
  }
  __syncthreads();
  __nv_bfloat16* pdy = dy + (long long)b * V * C;
  const __nv_bfloat16 val = __float2bfloat16_rn(sdp[threadIdx.x % C]);
  for (long long i = threadIdx.x; i < (long long)V * C; i += 128) pdy[i] = val;
}

static inline unsigned elt_grid(long long n8) {
  long long b = (n8 + THREADS - 1) / THREADS;
  const long long cap = 148LL * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace rn
}  // namespace mmnn

using namespace mmnn;
using namespace mmnn::rn;

extern "C" {

int mmnn_sizeof_rn_conv_geom(void) { return (int)sizeof(RnConvGeom); }

int mmnn_rn_conv(const RnConvGeom* g, int dgrad, int src_is_f32, const void* src, const float* w, void* dst, const void* add,
                 double* stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(dgrad ? PC_RN_DGRAD : PC_RN_FPROP, st);
  if (dgrad) return dispatch_conv<true>(*g, src_is_f32, src, w, dst, add, stats, st);
  return dispatch_conv<false>(*g, src_is_f32, src, w, dst, add, stats, st);
}

int mmnn_rn_conv_wgrad(const RnConvGeom* g, int x_is_f32, const void* x, const void* dy, float* dw, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(PC_RN_WGRAD, st);
  if (g->Cout % 8 != 0) return -2;
  const int items = g->kd * g->kh * g->kw * g->Cin * (g->Cout / 8);
  const int per = (items + THREADS - 1) / THREADS;
  if (x_is_f32) {
    if (per <= 2) return launch_wgrad<float, 2>(*g, x, dy, dw, st);
    if (per <= 4) return launch_wgrad<float, 4>(*g, x, dy, dw, st);
    return -3;
  }
  if (per <= 1) return launch_wgrad<__half, 1>(*g, x, dy, dw, st);
  if (per <= 2) return launch_wgrad<__half, 2>(*g, x, dy, dw, st);
  if (per <= 4) return launch_wgrad<__half, 4>(*g, x, dy, dw, st);
  if (per <= 7) return launch_wgrad<__half, 7>(*g, x, dy, dw, st);
  return -4;
}

int mmnn_rn_bn_coeffs(const double* stats, double count, const float* gamma, const float* beta, float* rmean, float* rvar,
                      long long* nbt, float eps, float momentum, int training, int C, float* coef, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(PC_RN_ELTWISE, st);
  rn_bn_coeffs_kernel<<<1, 64, 0, st>>>(stats, count, gamma, beta, rmean, rvar, nbt, eps, momentum, training, C, coef);
  return C > 64 ? -2 : (int)cudaGetLastError();
}

int mmnn_rn_bn_act(const void* raw, const float* coef, int res_mode, const void* res, const float* coef2, void* y,
                   long long elems, int C, int relu, float drop_p, unsigned long long seed, const unsigned char* mask,
                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (elems % 8 != 0 || (C != 8 && C != 16 && C != 64)) return -2;
  ProfScope ps(PC_RN_ELTWISE, st);
  const long long n8 = elems / 8;
  const unsigned grid = elt_grid(n8);
  if (res_mode == 0)
    rn_bn_act_kernel<0><<<grid, THREADS, 0, st>>>((const uint4*)raw, coef, nullptr, nullptr, (uint4*)y, n8, C, relu, drop_p, seed, mask);
  else if (res_mode == 1)
    rn_bn_act_kernel<1><<<grid, THREADS, 0, st>>>((const uint4*)raw, coef, (const uint4*)res, nullptr, (uint4*)y, n8, C, relu, drop_p, seed, mask);
  else
    rn_bn_act_kernel<2><<<grid, THREADS, 0, st>>>((const uint4*)raw, coef, (const uint4*)res, coef2, (uint4*)y, n8, C, relu, drop_p, seed, mask);
  return (int)cudaGetLastError();
}

int mmnn_rn_act_bwd_reduce(const void* dy, const void* y, float post_scale, const void* raw, const float* coef,
                           const void* raw2, const float* coef2, double* sums, long long elems, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (elems % 8 != 0 || (C != 8 && C != 16 && C != 64)) return -2;
  ProfScope ps(PC_RN_ELTWISE, st);
  const long long n8 = elems / 8;
  const unsigned grid = elt_grid(n8);
  if (raw2 != nullptr)
    rn_act_bwd_reduce_kernel<true><<<grid, THREADS, 0, st>>>((const uint4*)dy, (const uint4*)y, post_scale, (const uint4*)raw, coef, (const uint4*)raw2, coef2, sums, n8, C);
  else
    rn_act_bwd_reduce_kernel<false><<<grid, THREADS, 0, st>>>((const uint4*)dy, (const uint4*)y, post_scale, (const uint4*)raw, coef, nullptr, nullptr, sums, n8, C);
  return (int)cudaGetLastError();
}

int mmnn_rn_bn_bwd_apply(const void* dy, const void* y, float post_scale, const void* raw, const float* coef, const void* raw2,
                         const float* coef2, const double* sums, double inv_count, int eval_mode, void* draw, void* draw2,
                         void* dz, float* dgamma, float* dbeta, float* dgamma2, float* dbeta2, long long elems, int C,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (elems % 8 != 0 || (C != 8 && C != 16 && C != 64)) return -2;
  ProfScope ps(PC_RN_ELTWISE, st);
  const long long n8 = elems / 8;
  const unsigned grid = elt_grid(n8);
  if (raw2 != nullptr)
    rn_bn_bwd_apply_kernel<true><<<grid, THREADS, 0, st>>>((const uint4*)dy, (const uint4*)y, post_scale, (const uint4*)raw, coef, (const uint4*)raw2, coef2, sums, inv_count, eval_mode, (uint4*)draw, (uint4*)draw2, (uint4*)dz, dgamma, dbeta, dgamma2, dbeta2, n8, C);
  else
    rn_bn_bwd_apply_kernel<false><<<grid, THREADS, 0, st>>>((const uint4*)dy, (const uint4*)y, post_scale, (const uint4*)raw, coef, nullptr, nullptr, sums, inv_count, eval_mode, (uint4*)draw, nullptr, (uint4*)dz, dgamma, dbeta, nullptr, nullptr, n8, C);
  return (int)cudaGetLastError();
}

int mmnn_rn_head_fwd(const void* y, int B, int V, int C, const float* W, const float* bias, int K, float* pooled, float* out,
                     void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C > 64 || 128 % C != 0 || K > 64) return -2;
  ProfScope ps(PC_RN_HEAD, st);
  rn_head_fwd_kernel<<<B, 128, 0, st>>>((const __half*)y, V, C, W, bias, K, pooled, out);
  return (int)cudaGetLastError();
}

int mmnn_rn_head_bwd(const float* dout, const float* out, const float* pooled, const float* W, int B, int V, int C, int K,
                     void* dy, float* dW, float* db, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C > 64 || 128 % C != 0 || K > 64) return -2;
  ProfScope ps(PC_RN_HEAD, st);
  rn_head_bwd_kernel<<<B, 128, 0, st>>>(dout, out, pooled, W, V, C, K, (__nv_bfloat16*)dy, dW, db);
  return (int)cudaGetLastError();
}

}  // extern "C"
