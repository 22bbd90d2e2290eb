// tcgen05 tile engine: implicit-GEMM kernels shared by every convolution of the DenseNet-3D trunk.
//
// Operand staging.  Every operand tile in shared memory is a "chunk-plane" matrix
//        tile[chunk][row] of 16-byte cells   (cell = 8 consecutive channels of one voxel / one weight row)
// with rows contiguous at 16 B inside a plane.  8 consecutive rows of a plane are therefore one contiguous 128-byte
// core matrix, which is exactly the SWIZZLE_NONE canonical layout of the tcgen05 shared-memory descriptor:
//   * as a K-major operand (fprop / dgrad: M or N = row, K = channel)  LBO = plane stride, SBO = 128 B
//   * as an MN-major operand (wgrad: M or N = channel, K = row/voxel)  SBO = plane stride, LBO = 128 B
// so one producer routine feeds forward, data-gradient and weight-gradient GEMMs.
//
// Roles (288 threads): warps 0-7 = producers (gather + BN/ReLU transform of the activation operand, 16-byte
// vector loads, conflict-free 16-byte shared stores); warps 0-3 additionally run the epilogue after the K loop
// (TMEM -> registers -> global, fused per-channel statistics);  warp 8 = TMEM allocator + single-thread tcgen05.mma issuer.
// Weights arrive with one cp.async.bulk (TMA unit, 1-D) per k-block from a pre-packed image.
// Pipelines: full[stage] (128 producer arrivals + 1 expect_tx arrival), empty[stage] (tcgen05.commit), accum (commit).
#pragma once
#include "common.cuh"

namespace mmnn {

constexpr int TILE_ROWS = 128;
constexpr int PLANE_BYTES = TILE_ROWS * 16 + 16;  // +16 B pad: chunk planes land on distinct 16-byte bank groups
constexpr int PRODUCER_WARPS = 8;   // a lone warp per scheduler cannot hide its own ALU/LDS latency: 8 warps share a k-block
constexpr int NUM_PRODUCER_THREADS = PRODUCER_WARPS * 32;
constexpr int EPILOGUE_THREADS = 128;  // warps 0-3: one TMEM lane quarter each
constexpr int MMA_WARP = PRODUCER_WARPS;
constexpr int ENGINE_THREADS = NUM_PRODUCER_THREADS + 32;
constexpr int MAX_PASSES = 32 / PRODUCER_WARPS;

// Where a BatchNorm's per-channel scale/shift comes from (batch statistics accumulated by a producer kernel's
// epilogue, or running statistics in eval mode).
struct BnSrc {
  const double* sum;    // [C] sum x      (batch mode)
  const double* sumsq;  // [C] sum x^2
  const float* gamma;
  const float* beta;
  const float* rmean;  // eval mode
  const float* rvar;
  float inv_count;
  float eps;
  int use_batch;
};

MMNN_DEVINL void bn_mean_rstd(const BnSrc& b, int c, float& mean, float& rstd) {
  if (b.use_batch) {
    const double m = b.sum[c] * (double)b.inv_count;
    double var = b.sumsq[c] * (double)b.inv_count - m * m;
    var = var < 0.0 ? 0.0 : var;
    mean = (float)m;
    // fp64 only where cancellation matters (E[x^2] - E[x]^2); the reciprocal square root in fp32 + one Newton step
    const float vf = (float)var + b.eps;
    float r0 = rsqrtf(vf);
    rstd = r0 * (1.5f - 0.5f * vf * r0 * r0);
  } else {
    mean = b.rmean[c];
    rstd = rsqrtf(b.rvar[c] + b.eps);
  }
}

enum { A_LINEAR_CONV = 0, A_STEM = 1 };
// T_BNBWD (data-gradient GEMMs of small grids): the A operand is the BatchNorm-backward of the raw masked gradient,
//   a_c * v + b_c * x + d_c   (= gamma rstd (v - mean(v) - xhat mean(v xhat)),  x = the BatchNorm's forward input, from t_src),
// applied by the producers while the tile passes through registers -- the separate in-place pass (bn_bwd_apply, one more launch
// in the serial chain of every late dense layer) disappears; the CTAs of column tile 0 also store the transformed tile to t_out
// (the weight-gradient GEMM on the side stream reads it through TMA).
enum { T_NONE = 0, T_BNRELU = 1, T_BNBWD = 2 };
// EP_MASK_STATS_ACC (1x1x1 data gradient of a dense layer): like EP_MASK_STATS, but the masked gradient is not stored as a bf16
// tensor for a separate BatchNorm-backward pass -- it is ADDED, scaled per column by coefG, into the fp32 gradient accumulator
// of the block buffer (`out` is then a float*, one vector RED per 4 columns; every element has exactly one writer per launch,
// so the result does not depend on any ordering).  See encoder.cu ("deferred BatchNorm backward") for the algebra.
enum { EP_STORE = 0, EP_STORE_STATS = 1, EP_MASK_STATS = 2, EP_MASK_STATS_ACC = 3 };

struct RowsParams {
  int M;       // rows (output voxels)
  int NT;      // MMA N (tile width, multiple of 32, <= 256)
  int Ncols;   // total valid output columns over all N tiles
  int Cin;     // channels per tap of the A operand
  int kbw;     // k-block width in channels (64 or 32)
  int ntaps;   // 1 (1x1x1), 27 (3x3x3) or 16 (stem: (dz,dy) pairs of the space-to-depth 4x4x4 form)
  int tap_sign;
  int Dz, Dy, Dx;  // row space dims (voxels per sample = Dz*Dy*Dx)
  int Sz, Sy, Sx;  // stem only: padded space-to-depth source dims
  const bf16* a_src;
  long long a_pitch;  // elements between consecutive voxels of the source
  BnSrc bnA;
  const bf16* b_packed;  // [n tile][k-block][chunk][NT][8]
  bf16* out;
  long long out_pitch;
  const float* colscale;  // optional [batch][Ncols] multiplier (channel dropout keep-mask / (1-p))
  double* st_sum;         // EP_STORE_STATS: sum / sumsq of the stored values;  EP_MASK_STATS: sum dy / sum dy*xhat
  double* st_sq;
  const bf16* e_src;  // EP_MASK_STATS: forward input of the BN whose ReLU gates the gradient
  long long e_pitch;
  BnSrc bnE;
  int stages;
  int acc_rstd;       // EP_MASK_STATS_ACC: per-column scale of the accumulated gradient: 0 -> gamma (batch statistics: rstd is
                      // applied once, by the finalising pass), 1 -> gamma * rstd (running statistics differ per layer)
  // Early start (1x1x1 forward GEMM of a dense layer, small grids): the first early_ch input channels -- and their BatchNorm
  // statistics -- were final BEFORE the preceding kernel in the stream (the previous layer's 3x3x3 convolution, which only
  // appends 32 channels) started.  Launched with the programmatic-dependent-launch attribute, the kernel works through the
  // k-blocks below early_ch while that convolution is still running and calls griddepcontrol.wait only before the k-blocks
  // that contain its output.  0: wait first (plain stream order).
  int early_ch;
  // T_BNBWD
  const bf16* t_src;       // forward input of the BatchNorm (activation format), same row / channel indexing as a_src
  long long t_pitch;
  const double* t_gsum;    // backward statistics of that BatchNorm: sum v, sum v * xhat
  const double* t_gdot;
  float t_inv_count;       // 1 / rows in batch mode, 0 in eval mode (then the transform is gamma * rstd * v)
  bf16* t_out;             // transformed gradient [M][t_out_pitch] (bf16), written by the CTAs with blockIdx.y == 0
  long long t_out_pitch;
  // Small-grid 1x1x1 forward GEMM (late dense blocks): tma_a != 0 -> the RAW activation k-block (128 rows x 64 channels) arrives by
  // TMA (2-D map of [M][a_pitch], box 64 x 128, SWIZZLE_128B = the canonical K-major operand) up to stages - 2 k-blocks ahead, and
  // the producers apply BN+ReLU in place in shared memory: no global load in their instruction stream (the register prefetch kept at
  // most 2 k-blocks in flight; a depth of 4 spills under the 2-CTA register cap).
  int tma_a;
};

// 8 elements (16 B): load format IN, BN scale/shift + ReLU in fp32, store format OUT
template <bool IN_F16, bool OUT_F16>
MMNN_DEVINL void apply_bnrelu8(uint4& v, const float* sc, const float* sh) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    unpack2<IN_F16>(w[i], lo, hi);
    const float a = fmaxf(fmaf(lo, sc[2 * i], sh[2 * i]), 0.f);
    const float b = fmaxf(fmaf(hi, sc[2 * i + 1], sh[2 * i + 1]), 0.f);
    w[i] = pack2<OUT_F16>(a, b);
  }
}
// fp16 in / fp16 out fast path: y = relu(x*s + t) with s = s_hi + s_lo, t = t_hi + t_lo split into fp16 pairs, 3 packed
// HFMA2 per channel pair (12 per 16-byte cell instead of 28 scalar instructions); every step rounds to fp16, the result
// carries ~2 fp16 roundings instead of 1.
struct H2Coef { __half2 s_hi, s_lo, t_hi, t_lo; };   // one channel pair
MMNN_DEVINL void apply_bnrelu8_h2(uint4& v, const H2Coef (&c)[4]) {
  __half2* w = reinterpret_cast<__half2*>(&v);
  const __half2 one = __floats2half2_rn(1.f, 1.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 r0 = __hfma2(w[i], c[i].s_lo, c[i].t_lo);
    const __half2 r1 = __hfma2(w[i], c[i].s_hi, c[i].t_hi);
    w[i] = __hfma2_relu(r0, one, r1);
  }
}
// fills the per-pair coefficient table for `C` channels from fp32 scale/shift arrays
MMNN_DEVINL void fill_h2coef(H2Coef* dst, const float* scale, const float* shift, int C, int tid, int nthreads) {
  for (int j = tid; j < C / 2; j += nthreads) {
    const float s0 = scale[2 * j], s1 = scale[2 * j + 1], t0 = shift[2 * j], t1 = shift[2 * j + 1];
    H2Coef c;
    c.s_hi = __floats2half2_rn(s0, s1);
    c.t_hi = __floats2half2_rn(t0, t1);
    const float2 sh = __half22float2(c.s_hi), th = __half22float2(c.t_hi);
    c.s_lo = __floats2half2_rn(s0 - sh.x, s1 - sh.y);
    c.t_lo = __floats2half2_rn(t0 - th.x, t1 - th.y);
    dst[j] = c;
  }
}

template <bool IN_F16, bool OUT_F16>
MMNN_DEVINL void convert8(uint4& v) {
  if (IN_F16 == OUT_F16) return;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    unpack2<IN_F16>(w[i], lo, hi);
    w[i] = pack2<OUT_F16>(lo, hi);
  }
}

struct EngineSmem {
  uint32_t full, empty, accum, tmem_ptr, rowinfo, coefA, coefE, red, stage0;
  uint32_t stage_bytes, a_bytes;
};

// dynamic smem carve-up shared by host (size) and device (offsets)
__host__ __device__ inline uint32_t rows_stage_bytes(int NT, int kbw, bool tma) {
  const uint32_t stage = (uint32_t)(kbw / 8) * PLANE_BYTES + (uint32_t)(kbw / 8) * NT * 16;
  return tma ? (stage + 1023u) & ~1023u : stage;      // TMA mode: [A 16 KB swizzled, 1 KB aligned][B]
}
__host__ __device__ inline uint32_t rows_smem_layout(int Cin, int NT, int kbw, int stages, uint32_t* offs /*[6]*/, bool tma = false) {
  uint32_t o = 0;
  offs[0] = o; o += 192;                 // barriers + tmem ptr (+128: raw[6] of the TMA mode)
  offs[1] = o; o += TILE_ROWS * 16;      // rowinfo
  offs[2] = o; o += 3u * Cin * 4;        // coefA: scale, shift (T_BNRELU) / a, b, d (T_BNBWD)
  offs[3] = o; o += 5u * NT * 4;         // coefE: scale, shift, mean, rstd, accumulate scale
  offs[4] = o; o += 8u * NT * 4;         // red[2][4][NT]
  o = (o + 127u) & ~127u;
  offs[5] = o;
  const uint32_t stage = rows_stage_bytes(NT, kbw, tma);
  const uint32_t ring = stages * stage + (tma ? 1024u : 0u);
  // the operand ring is reused by the epilogue to stage the rounded output tile (16 chunk planes) + 512 B of ones for the
  // tensor-core column statistics (NT == 128 launches)
  const uint32_t tcstats = 16u * PLANE_BYTES + 512u;
  return o + (ring > tcstats ? ring : tcstats);
}

// GRAD == false: forward GEMM  -- A = activations (act format), weights packed in act format, output act format.
// GRAD == true : data-gradient -- A = gradients (bf16), weights packed bf16, output bf16; e_src (forward activation
//                                 gating the ReLU / feeding x-hat) is in act format.
// PF = register prefetch distance in k-blocks: 1 for grids of many tiles (registers -> occupancy), 2 for grids that do not
// fill the GPU (tiny late-block layers are bound by the serial chain of K iterations, not by occupancy).
template <int AMODE, int TRANS, int EPI, bool GRAD, int PF>
__global__ void __launch_bounds__(ENGINE_THREADS, 2) conv_rows_kernel(const __grid_constant__ RowsParams p,
                                                                      const __grid_constant__ CUtensorMap tma) {
  constexpr bool OP_F16 = !GRAD && kActF16;   // MMA operand + output format of this launch
  constexpr bool E_F16 = kActF16;
  constexpr bool MASK = EPI == EP_MASK_STATS || EPI == EP_MASK_STATS_ACC;
  constexpr bool ACC = EPI == EP_MASK_STATS_ACC;
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t offs[6];
  // raw activation k-blocks by TMA, BN+ReLU in place (small-grid forward launches; RowsParams::tma_a)
  const bool tma_a = PF == 2 && AMODE == A_LINEAR_CONV && TRANS == T_BNRELU && OP_F16 && !GRAD && p.ntaps == 1 && p.kbw == 64 && p.tma_a != 0;
  rows_smem_layout(p.Cin, p.NT, p.kbw, p.stages & 0xff, offs, tma_a);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + offs[0];
  const uint32_t bar_empty = bar_full + 8 * 6;
  const uint32_t bar_accum = bar_full + 8 * 12;
  const uint32_t bar_staged = bar_full + 8 * 14, bar_stats = bar_full + 8 * 15;
  const uint32_t bar_raw = bar_full + 128;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + offs[0] + 8 * 13);
  int4* rowinfo = reinterpret_cast<int4*>(smem + offs[1]);
  float* coefA = reinterpret_cast<float*>(smem + offs[2]);
  float* coefE = reinterpret_cast<float*>(smem + offs[3]);
  float* red = reinterpret_cast<float*>(smem + offs[4]);
  const int planes = p.kbw / 8;
  const uint32_t a_bytes = planes * PLANE_BYTES;
  const uint32_t b_bytes = planes * p.NT * 16;
  const uint32_t stage_bytes = rows_stage_bytes(p.NT, p.kbw, tma_a);
  const uint32_t stage0 = tma_a ? ((sbase + offs[5] + 1023u) & ~1023u) : sbase + offs[5];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int S = p.stages & 0xff;
  const int kb_per_tap = (p.Cin + p.kbw - 1) / p.kbw;
  const int KB = p.ntaps * kb_per_tap;
  const int vps = p.Dz * p.Dy * p.Dx;
  // Column statistics on the tensor core (NT == 128): the epilogue stages the ROUNDED output tile in the (by then free)
  // operand ring as an MN-major operand G[row][col]; the MMA warp computes  sum_rows G = G^T x ones  (N = 16, TMEM columns
  // NT..NT+15) and, for the forward statistics, sum_rows G^2 = diag(G^T x G) (into the drained accumulator columns) --
  // exact products, fp32 accumulation -- instead of one / two 32x32 warp transposes per column chunk, which are 40 % of
  // the epilogue's instructions.  Parity-tested, but in a same-box A/B at configs[1] it is NEUTRAL (step 14.92 vs 14.90 ms:
  // forward -0.6 %, data gradient +2.4 %): in a one-tile-per-CTA kernel the extra stage -> MMA -> commit -> TMEM-load round
  // trip at the end of the CTA costs what the transposes did.  OFF unless MMNN_TC_STATS=1; the place for it is a
  // persistent kernel where that round trip overlaps the next tile.
  const bool tcs = (EPI != EP_STORE) && !ACC && p.NT == 128 && (p.stages & 0x100) != 0;   // bit 8 of `stages`: opt-in switch (MMNN_TC_STATS=1)
  const uint32_t sG = stage0, sOnes = stage0 + 16u * PLANE_BYTES;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.NT + (tcs ? 32 : 0)) tmem_cols <<= 1;

  // ---------------- prologue
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(bar_full + 8 * s, NUM_PRODUCER_THREADS + 1);
        mbar_init(bar_empty + 8 * s, 1);
        mbar_init(bar_raw + 8 * s, 1);
      }
      if (tma_a) tma_prefetch_desc(&tma);
      mbar_init(bar_accum, 1);
      mbar_init(bar_staged, NUM_PRODUCER_THREADS);
      mbar_init(bar_stats, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
  }
  if (PF == 2 && warp == 0 && lane == 0) {
    // Small grids (late dense blocks) are latency chains: every k-block waits for its weight image (one cp.async.bulk from
    // wherever it lives -- packed at the start of the forward pass, usually evicted to HBM by now).  The weights do not depend on
    // the preceding kernel, so the whole image of this CTA's column tile is pulled into L2 right away, in pieces of <= 64 KB.
    const uint8_t* w = reinterpret_cast<const uint8_t*>(p.b_packed + (size_t)tile_n * KB * (size_t)(planes * p.NT * 8));
    const uint32_t total = (uint32_t)KB * b_bytes;
    for (uint32_t o = 0; o < total; o += 65536u) bulk_prefetch_l2(w + o, total - o < 65536u ? total - o : 65536u);
  }
  // early start: only the leading `early_ch` channels may be touched before the preceding kernel has completed
  const bool early = PF == 2 && TRANS == T_BNRELU && OP_F16 && EPI == EP_STORE_STATS && p.ntaps == 1 && p.early_ch >= p.kbw;
  const int coef_ch = early ? (p.early_ch / p.kbw) * p.kbw : p.Cin;     // channels whose coefficients are computed in the prologue
  if (!early) { pdl_wait(); pdl_trigger(); }   // nothing above touches global memory written by the preceding kernel
  H2Coef* coefH = reinterpret_cast<H2Coef*>(coefA);   // fp16 operands: packed per-pair table in place of the fp32 one (same size)
  if (TRANS == T_BNRELU) {
    if (OP_F16) {
      for (int j = tid; j < coef_ch / 2; j += ENGINE_THREADS) {
        float m0, r0, m1, r1;
        bn_mean_rstd(p.bnA, 2 * j, m0, r0);
        bn_mean_rstd(p.bnA, 2 * j + 1, m1, r1);
        const float s0 = p.bnA.gamma[2 * j] * r0, s1 = p.bnA.gamma[2 * j + 1] * r1;
        const float t0 = p.bnA.beta[2 * j] - m0 * s0, t1 = p.bnA.beta[2 * j + 1] - m1 * s1;
        H2Coef c;
        c.s_hi = __floats2half2_rn(s0, s1);
        c.t_hi = __floats2half2_rn(t0, t1);
        const float2 sh = __half22float2(c.s_hi), th = __half22float2(c.t_hi);
        c.s_lo = __floats2half2_rn(s0 - sh.x, s1 - sh.y);
        c.t_lo = __floats2half2_rn(t0 - th.x, t1 - th.y);
        coefH[j] = c;
      }
    } else {
      for (int c = tid; c < p.Cin; c += ENGINE_THREADS) {
        float mean, rstd;
        bn_mean_rstd(p.bnA, c, mean, rstd);
        const float s = p.bnA.gamma[c] * rstd;
        coefA[c] = s;
        coefA[p.Cin + c] = p.bnA.beta[c] - mean * s;
      }
    }
  }
  if (TRANS == T_BNBWD) {
    for (int c = tid; c < p.Cin; c += ENGINE_THREADS) {
      float mean, rstd;
      bn_mean_rstd(p.bnA, c, mean, rstd);
      const float k = p.bnA.gamma[c] * rstd;
      const float c1 = (float)(p.t_gsum[c] * (double)p.t_inv_count);
      const float c2 = (float)(p.t_gdot[c] * (double)p.t_inv_count);
      coefA[c] = k;
      coefA[p.Cin + c] = -k * rstd * c2;
      coefA[2 * p.Cin + c] = -k * c1 + k * rstd * c2 * mean;
    }
  }
  if (MASK) {
    for (int c = tid; c < p.NT; c += ENGINE_THREADS) {
      const int col = tile_n * p.NT + c;
      float mean = 0.f, rstd = 0.f, s = 0.f, t = -1.f, ga = 0.f;
      if (col < p.Ncols) {
        bn_mean_rstd(p.bnE, col, mean, rstd);
        ga = p.bnE.gamma[col];
        s = ga * rstd;
        t = p.bnE.beta[col] - mean * s;
      }
      coefE[c] = s; coefE[p.NT + c] = t; coefE[2 * p.NT + c] = mean; coefE[3 * p.NT + c] = rstd;
      if (ACC) coefE[4 * p.NT + c] = p.acc_rstd ? s : ga;
    }
  }
  if (tid < TILE_ROWS) {
    const long long m = (long long)tile_m * TILE_ROWS + tid;
    int4 ri;
    if (m < p.M) {
      const int n = (int)(m / vps);
      int rem = (int)(m - (long long)n * vps);
      const int z = rem / (p.Dy * p.Dx);
      rem -= z * p.Dy * p.Dx;
      const int y = rem / p.Dx;
      const int x = rem - y * p.Dx;
      if (AMODE == A_STEM) {
        ri.x = ((n * p.Sz + z) * p.Sy + y) * p.Sx + x;
      } else {
        ri.x = (int)m;
      }
      ri.y = z; ri.z = y; ri.w = x;
    } else {
      ri.x = 0; ri.y = -100000; ri.z = 0; ri.w = 0;
    }
    rowinfo[tid] = ri;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < PRODUCER_WARPS) {
    // ================= producers
    // A k-block is 128 rows x cpl 16-byte cells.  One warp-wide 128-bit load covers rpp = 32/cpl rows x (cpl*16)
    // contiguous bytes ("group"); warp w owns groups w, w+8, ...  Register double buffering: the loads of k-block
    // kb+1 are issued BEFORE k-block kb is transformed and stored, so loads stay in flight across the slot wait.
    struct KbGeom { int cb, cpl, dz, dy, dx; long long delta; };
    auto geom = [&](int tap, int cb) {
      KbGeom gq;
      gq.cb = cb;
      int cpl = (p.Cin - cb * p.kbw) / 8;
      gq.cpl = cpl > planes ? planes : cpl;
      gq.dz = gq.dy = gq.dx = 0; gq.delta = 0;
      if (AMODE == A_STEM) {
        gq.delta = (long long)((tap >> 2) * p.Sy + (tap & 3)) * p.Sx;
      } else if (p.ntaps == 27) {
        const int t9 = tap / 9, t3 = (tap - t9 * 9) / 3;
        gq.dz = (t9 - 1) * p.tap_sign; gq.dy = (t3 - 1) * p.tap_sign; gq.dx = (tap - t9 * 9 - t3 * 3 - 1) * p.tap_sign;
        gq.delta = (long long)(gq.dz * p.Dy + gq.dy) * p.Dx + gq.dx;
      }
      return gq;
    };
    auto load_kb = [&](const KbGeom& gq, uint4 (&regs)[MAX_PASSES], uint32_t& okmask) {
      const int cshift = (gq.cpl == 8) ? 3 : 2;
      const int chunk = lane & (gq.cpl == 8 ? 7 : 3);
      const int rsub = lane >> cshift;
      const int rpp = 32 >> cshift;
      const int npass = (TILE_ROWS / rpp) / PRODUCER_WARPS;
      const int ch0 = gq.cb * p.kbw + chunk * 8;
      okmask = 0;
#pragma unroll
      for (int ps = 0; ps < MAX_PASSES; ++ps) {
        regs[ps] = make_uint4(0, 0, 0, 0);
        if (ps < npass) {
          const int r = (warp + ps * PRODUCER_WARPS) * rpp + rsub;
          const int4 ri = rowinfo[r];
          bool ok = ri.y > -1000;
          if (AMODE == A_LINEAR_CONV) {
            const int zz = ri.y + gq.dz, yy = ri.z + gq.dy, xx = ri.w + gq.dx;
            ok = ok && zz >= 0 && zz < p.Dz && yy >= 0 && yy < p.Dy && xx >= 0 && xx < p.Dx;
          }
          if (ok) {
            regs[ps] = ldg16(p.a_src + ((long long)ri.x + gq.delta) * p.a_pitch + ch0);
            okmask |= 1u << ps;
          }
        }
      }
    };
    // PF+1 rotating register sets: the loads of k-blocks kb+1 .. kb+PF are in flight while kb is transformed and stored
    uint4 R[PF + 1][MAX_PASSES];
    uint32_t OK[PF + 1];
    KbGeom G[PF + 1];
    int tap_n = 0, cb_n = 0;   // (tap, cb) of the NEXT k-block to load
    auto next_geom = [&]() {
      const KbGeom gq = geom(tap_n, cb_n);
      if (++cb_n == kb_per_tap) { cb_n = 0; ++tap_n; }
      return gq;
    };
    int s = 0;
    uint32_t ph = 0;
    auto process = [&](int kb, const KbGeom& gq, uint4 (&regs)[MAX_PASSES], uint32_t okm) {
      mbar_wait(bar_empty + 8 * s, ph ^ 1u, 1);
      const uint32_t sA = stage0 + s * stage_bytes;
      if (tid == 0) {
        mbar_arrive_expect_tx(bar_full + 8 * s, b_bytes);
        bulk_g2s(sA + a_bytes, p.b_packed + ((size_t)tile_n * KB + kb) * (size_t)(planes * p.NT * 8), b_bytes, bar_full + 8 * s);
      }
      const int cshift = (gq.cpl == 8) ? 3 : 2;
      const int chunk = lane & (gq.cpl == 8 ? 7 : 3);
      const int rsub = lane >> cshift;
      const int rpp = 32 >> cshift;
      const int npass = (TILE_ROWS / rpp) / PRODUCER_WARPS;
      const int ch0 = gq.cb * p.kbw + chunk * 8;
      float sc[8], sh[8];
      H2Coef hc[4];
      if (TRANS == T_BNRELU) {
        if (OP_F16) {
#pragma unroll
          for (int i = 0; i < 4; ++i) hc[i] = coefH[ch0 / 2 + i];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { sc[e] = coefA[ch0 + e]; sh[e] = coefA[p.Cin + ch0 + e]; }
        }
      }
      uint4 xr[MAX_PASSES];
      if (TRANS == T_BNBWD) {
        // the BatchNorm's forward input of the same cells: issued together, one L2 round trip per k-block (two k-blocks per launch)
#pragma unroll
        for (int ps = 0; ps < MAX_PASSES; ++ps) {
          xr[ps] = make_uint4(0, 0, 0, 0);
          if (ps < npass && ((okm >> ps) & 1u)) {
            const int r = (warp + ps * PRODUCER_WARPS) * rpp + rsub;
            xr[ps] = ldg16(p.t_src + (long long)rowinfo[r].x * p.t_pitch + ch0);
          }
        }
      }
#pragma unroll
      for (int ps = 0; ps < MAX_PASSES; ++ps) {
        if (ps < npass) {
          const int r = (warp + ps * PRODUCER_WARPS) * rpp + rsub;
          uint4 v = regs[ps];
          if (TRANS == T_BNRELU && ((okm >> ps) & 1u)) {
            if (OP_F16) apply_bnrelu8_h2(v, hc);     // 12 packed HFMA2 per cell instead of 28 scalar instructions
            else apply_bnrelu8<OP_F16, OP_F16>(v, sc, sh);
          }
          if (TRANS == T_BNBWD && ((okm >> ps) & 1u)) {
            uint32_t* vw = reinterpret_cast<uint32_t*>(&v);
            const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xr[ps]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float v0, v1, x0, x1;
              unpack2<false>(vw[i], v0, v1);
              unpack2<E_F16>(xw[i], x0, x1);
              const int c = ch0 + 2 * i;
              const float o0 = fmaf(coefA[c], v0, fmaf(coefA[p.Cin + c], x0, coefA[2 * p.Cin + c]));
              const float o1 = fmaf(coefA[c + 1], v1, fmaf(coefA[p.Cin + c + 1], x1, coefA[2 * p.Cin + c + 1]));
              vw[i] = pack2<false>(o0, o1);
            }
            if (tile_n == 0 && p.t_out != nullptr)
              *reinterpret_cast<uint4*>(p.t_out + (long long)rowinfo[r].x * p.t_out_pitch + ch0) = v;
          }
          sts16(sA + chunk * PLANE_BYTES + r * 16, v);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_full + 8 * s);
      if (++s == S) { s = 0; ph ^= 1u; }
    };
    // the k loop runs over [kb0, kb1): once for the whole K, or -- early start -- first over the k-blocks of channels that were
    // final before the preceding kernel started, then (after griddepcontrol.wait) over the rest
    // TMA mode: thread 0 feeds the ring LA = S - 2 k-blocks ahead (raw activation box -> raw[s], weight image -> full[s]); waiting
    // for the stage of k-block j - S cannot deadlock: that k-block was handed to the MMA warp two iterations ago.  Every thread
    // then transforms its 4 cells of the swizzled tile in place: logical 16-byte chunk c = tid & 7 of rows (tid >> 3) + 32 j, at
    // physical chunk c ^ (row & 7); rows beyond M stay the zeros the TMA unit delivered.
    const int LA = S > 2 ? S - 2 : 1;
    auto tma_issue = [&](int j) {
      const int sj = j % S;
      mbar_wait(bar_empty + 8 * sj, ((uint32_t)(j / S) & 1u) ^ 1u, 1);
      const uint32_t sAj = stage0 + sj * stage_bytes;
      mbar_arrive_expect_tx(bar_raw + 8 * sj, (uint32_t)TILE_ROWS * 128u);
      tma_load_2d(sAj, &tma, j * 64, tile_m * TILE_ROWS, bar_raw + 8 * sj);
      mbar_arrive_expect_tx(bar_full + 8 * sj, b_bytes);
      bulk_g2s(sAj + (uint32_t)TILE_ROWS * 128u, p.b_packed + ((size_t)tile_n * KB + j) * (size_t)(planes * p.NT * 8), b_bytes, bar_full + 8 * sj);
    };
    auto run_range_tma = [&](int kb0, int kb1) {
      if (tid == 0)
        for (int j = kb0; j < kb1 && j < kb0 + LA; ++j) tma_issue(j);
      const int c = tid & 7, rb = tid >> 3;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (tid == 0 && kb + LA < kb1) tma_issue(kb + LA);
        const int s2 = kb % S;
        mbar_wait(bar_raw + 8 * s2, (uint32_t)(kb / S) & 1u, 3);
        const uint32_t sA = stage0 + s2 * stage_bytes;
        int cpl = (p.Cin - kb * 64) / 8;
        cpl = cpl > 8 ? 8 : cpl;
        if (c < cpl) {
          H2Coef hc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) hc[i] = coefH[(kb * 64 + c * 8) / 2 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = rb + 32 * j;
            if ((long long)tile_m * TILE_ROWS + r < p.M) {
              const uint32_t addr = sA + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
              uint4 v = lds16(addr);
              apply_bnrelu8_h2(v, hc);
              sts16(addr, v);
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_full + 8 * s2);
      }
    };
    auto run_range = [&](int kb0, int kb1) {
      if (tma_a) { run_range_tma(kb0, kb1); return; }
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        OK[u] = 0;
        if (kb0 + u < kb1) { G[u] = next_geom(); load_kb(G[u], R[u], OK[u]); }
      }
      for (int kb = kb0; kb < kb1; kb += PF + 1) {
#pragma unroll
        for (int u = 0; u <= PF; ++u) {
          const int k = kb + u;
          if (k < kb1) {
            if (k + PF < kb1) {
              G[(u + PF) % (PF + 1)] = next_geom();
              load_kb(G[(u + PF) % (PF + 1)], R[(u + PF) % (PF + 1)], OK[(u + PF) % (PF + 1)]);
            }
            process(k, G[u], R[u], OK[u]);
          }
        }
      }
    };
    if (early) {
      const int kb_early = coef_ch / p.kbw;
      run_range(0, kb_early);
      pdl_wait();                                   // the preceding kernel (the previous layer's 3x3x3 conv) is complete and visible
      pdl_trigger();
      for (int j = coef_ch / 2 + tid; j < p.Cin / 2; j += NUM_PRODUCER_THREADS) {
        float m0, r0, m1, r1;
        bn_mean_rstd(p.bnA, 2 * j, m0, r0);
        bn_mean_rstd(p.bnA, 2 * j + 1, m1, r1);
        const float s0 = p.bnA.gamma[2 * j] * r0, s1 = p.bnA.gamma[2 * j + 1] * r1;
        const float t0 = p.bnA.beta[2 * j] - m0 * s0, t1 = p.bnA.beta[2 * j + 1] - m1 * s1;
        H2Coef c;
        c.s_hi = __floats2half2_rn(s0, s1);
        c.t_hi = __floats2half2_rn(t0, t1);
        const float2 sh = __half22float2(c.s_hi), th = __half22float2(c.t_hi);
        c.s_lo = __floats2half2_rn(s0 - sh.x, s1 - sh.y);
        c.t_lo = __floats2half2_rn(t0 - th.x, t1 - th.y);
        coefH[j] = c;
      }
      named_bar_sync(1, NUM_PRODUCER_THREADS);
      run_range(kb_early, KB);
    } else {
      run_range(0, KB);
    }
  }
  if (warp < PRODUCER_WARPS) {
    // ================= epilogue on all 8 producer warps: warps w and w+4 share TMEM lane quarter w%4 and split the
    // 32-column chunks between them (the epilogue, not the MMA, bounds the thin-K GEMMs: ncu, profiles/)
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const long long m = (long long)tile_m * TILE_ROWS + r;
    const bool row_ok = m < p.M;
    const int nb = row_ok ? (int)(m / vps) : 0;
    // the forward activations that gate the gradient do not depend on the MMA: fetch this row's (up to 128 columns)
    // before waiting for the accumulator so their latency hides behind the tail of the K loop
    uint4 xpre[2][4];
    if (MASK) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int cc = (warp >> 2) + 2 * k;
        const int col0 = tile_n * p.NT + cc * 32;
        const bool on = row_ok && cc * 32 < p.NT && col0 < p.Ncols;
#pragma unroll
        for (int i = 0; i < 4; ++i) xpre[k][i] = on ? ldg16(p.e_src + m * p.e_pitch + col0 + i * 8) : make_uint4(0, 0, 0, 0);
      }
    }
    mbar_wait(bar_accum, 0, 3);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cc = (warp >> 2) + 2 * k;
      if (cc >= p.NT / 32) break;
      const int col0 = tile_n * p.NT + cc * 32;
      if (col0 >= p.Ncols) {
        if (EPI != EP_STORE) {
          red[(0 * 4 + qd) * p.NT + cc * 32 + lane] = 0.f;
          red[(1 * 4 + qd) * p.NT + cc * 32 + lane] = 0.f;
        }
        continue;
      }
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(cc * 32), v);
      float q[32];
      if (MASK) {
        uint4 xv[4];
        if (k < 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = xpre[k][i];
        } else {       // column tiles wider than 128 (single-tile data gradients of 160..256 channels): chunks 4..7 are fetched here
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = row_ok ? ldg16(p.e_src + m * p.e_pitch + col0 + i * 8) : make_uint4(0, 0, 0, 0);
        }
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float xlo, xhi;
          unpack2<E_F16>(xw[j >> 1], xlo, xhi);
          const float x = (j & 1) ? xhi : xlo;
          const int c = cc * 32 + j;
          const bool act = fmaf(x, coefE[c], coefE[p.NT + c]) > 0.f;
          const float g = (row_ok && act) ? (ACC ? v[j] : round16<OP_F16>(v[j])) : 0.f;   // ACC: the fp32 value itself is accumulated
          v[j] = g;
          q[j] = g * (x - coefE[2 * p.NT + c]) * coefE[3 * p.NT + c];
        }
      } else {
        if (p.colscale != nullptr) {
          const float* cs = p.colscale + (size_t)nb * p.Ncols + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= __ldg(cs + j);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float g = row_ok ? round16<OP_F16>(v[j]) : 0.f;
          v[j] = g;
          q[j] = g * g;
        }
      }
      if (ACC) {
        if (row_ok) {
          float4* gp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.out_pitch + col0);
          const float* cg = coefE + 4 * p.NT + cc * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            atomicAdd(gp + i, make_float4(v[4 * i] * cg[4 * i], v[4 * i + 1] * cg[4 * i + 1], v[4 * i + 2] * cg[4 * i + 2], v[4 * i + 3] * cg[4 * i + 3]));
        }
      } else {
        uint4* op = reinterpret_cast<uint4*>(p.out + m * p.out_pitch + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack2<OP_F16>(v[8 * i + 0], v[8 * i + 1]); o.y = pack2<OP_F16>(v[8 * i + 2], v[8 * i + 3]);
          o.z = pack2<OP_F16>(v[8 * i + 4], v[8 * i + 5]); o.w = pack2<OP_F16>(v[8 * i + 6], v[8 * i + 7]);
          if (row_ok) op[i] = o;
          if (tcs) sts16(sG + (uint32_t)(cc * 4 + i) * PLANE_BYTES + (uint32_t)r * 16u, o);   // rows beyond M stage zeros
        }
      }
      if (EPI != EP_STORE) {
        if (!tcs) {
          const float s1 = warp_transpose_sum32(v, lane);
          red[(0 * 4 + qd) * p.NT + cc * 32 + lane] = s1;
        }
        if (!tcs || MASK) {
          const float s2 = warp_transpose_sum32(q, lane);
          red[(1 * 4 + qd) * p.NT + cc * 32 + lane] = s2;
        }
      }
    }
    if (tcs) {
      if (warp == 0) {   // 512 B of 1.0 in the operand format
        const uint32_t one = OP_F16 ? 0x3c003c00u : 0x3f803f80u;
        sts16(sOnes + lane * 16, make_uint4(one, one, one, one));
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_staged);
      mbar_wait(bar_stats, 0, 4);
      tc_fence_after();
      if (warp < 4) {    // one warp per TMEM lane quarter: lane = output column qd*32 + lane
        float d[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)p.NT, d);
        const int c = qd * 32 + lane;
        red[(0 * 4 + 0) * p.NT + c] = d[0];
        red[(0 * 4 + 1) * p.NT + c] = 0.f; red[(0 * 4 + 2) * p.NT + c] = 0.f; red[(0 * 4 + 3) * p.NT + c] = 0.f;
        if (EPI == EP_STORE_STATS) {
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(qd * 32), d);   // the 32x32 block holding the diagonal
          float x = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) x = (lane == j) ? d[j] : x;
          red[(1 * 4 + 0) * p.NT + c] = x;
          red[(1 * 4 + 1) * p.NT + c] = 0.f; red[(1 * 4 + 2) * p.NT + c] = 0.f; red[(1 * 4 + 3) * p.NT + c] = 0.f;
        }
      }
    }
    if (EPI != EP_STORE) {
      named_bar_sync(1, NUM_PRODUCER_THREADS);
      for (int c = tid; c < p.NT; c += NUM_PRODUCER_THREADS) {
        const int col = tile_n * p.NT + c;
        if (col < p.Ncols) {
          const float a = red[(0 * 4 + 0) * p.NT + c] + red[(0 * 4 + 1) * p.NT + c] + red[(0 * 4 + 2) * p.NT + c] + red[(0 * 4 + 3) * p.NT + c];
          const float b = red[(1 * 4 + 0) * p.NT + c] + red[(1 * 4 + 1) * p.NT + c] + red[(1 * 4 + 2) * p.NT + c] + red[(1 * 4 + 3) * p.NT + c];
          atomicAdd(p.st_sum + col, (double)a);
          atomicAdd(p.st_sq + col, (double)b);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer (one elected lane)
    const uint32_t idesc = make_idesc(TILE_ROWS, p.NT, 0, 0, OP_F16);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (uint32_t)(kb / S) & 1u;
      mbar_wait(bar_full + 8 * s, ph, 2);
      tc_fence_after();
      if (elect_one()) {
        const int tap = kb / kb_per_tap;
        const int cb = kb - tap * kb_per_tap;
        int cpl = (p.Cin - cb * p.kbw) / 8;
        cpl = cpl > planes ? planes : cpl;
        const uint32_t sA = stage0 + s * stage_bytes;
        // A: SWIZZLE_NONE chunk planes, or (TMA mode) 128-byte rows with the 128-byte swizzle: 8-row atoms 1 KB apart, a K step of
        // 16 channels = 32 B inside the row
        const uint64_t ad0 = tma_a ? make_smem_desc_sw(sA, 16, 1024, 2u) : make_smem_desc(sA, PLANE_BYTES, 128);
        const uint64_t bd0 = make_smem_desc(sA + (tma_a ? (uint32_t)TILE_ROWS * 128u : a_bytes), p.NT * 16, 128);
        const uint32_t a_step = tma_a ? 32u : 2 * PLANE_BYTES, b_step = 2 * p.NT * 16;
        tc_mma_bf16(tmem_base, ad0, bd0, idesc, kb > 0 ? 1u : 0u);
#pragma unroll
        for (int k16 = 1; k16 < 4; ++k16)
          if (k16 < cpl / 2) tc_mma_bf16(tmem_base, desc_advance(ad0, k16 * a_step), desc_advance(bd0, k16 * b_step), idesc, 1u);
        tc_commit(bar_empty + 8 * s);
        if (kb == KB - 1) tc_commit(bar_accum);
      }
      __syncwarp();
    }
    if (tcs) {
      mbar_wait(bar_staged, 0, 5);     // the rounded tile is staged (generic-proxy stores fenced by their writers)
      tc_fence_after();
      if (elect_one()) {
        const uint64_t gd = make_smem_desc(sG, 128, PLANE_BYTES);          // MN-major: 8-column chunks PLANE_BYTES apart, 8-row groups 128 B apart
        const uint64_t od = make_smem_desc(sOnes, 128, 128);
        const uint32_t id_sum = make_idesc(TILE_ROWS, 16, 1, 1, OP_F16);
#pragma unroll
        for (int k16 = 0; k16 < TILE_ROWS / 16; ++k16)
          tc_mma_bf16(tmem_base + p.NT, desc_advance(gd, k16 * 256), od, id_sum, k16 > 0 ? 1u : 0u);
        if (EPI == EP_STORE_STATS) {
          const uint32_t id_sq = make_idesc(TILE_ROWS, 128, 1, 1, OP_F16);
#pragma unroll
          for (int k16 = 0; k16 < TILE_ROWS / 16; ++k16)
            tc_mma_bf16(tmem_base, desc_advance(gd, k16 * 256), desc_advance(gd, k16 * 256), id_sq, k16 > 0 ? 1u : 0u);
        }
        tc_commit(bar_stats);
      }
      __syncwarp();
    }
  }
  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mmnn

// =====================================================================================================================
// Weight-gradient kernel:  D_j[a][b] = sum over voxels v of  Aop[v][a] * Bop_j[v][b]        (K = voxels, split across CTAs)
// Both operands are chunk-plane tiles of 128 voxels read as MN-major (channel-major) matrices.  A CTA walks a
// contiguous range of voxel tiles, accumulating in TMEM, then adds its partial result into the fp32 gradient with
// lane-contiguous atomics.
namespace mmnn {

enum { WA_LINEAR = 0, WA_STEM_PAIR = 1 };
enum { WE_STRIDED = 0, WE_STEM = 1 };

struct WgradParams {
  int M;           // voxels
  int CB;          // MMA N (B channels per tile), multiple of 32, <= 128
  int NB;          // B tiles (taps) per stage: 1 or 9
  int na_total;    // valid A channels over all z tiles (A tile = 128 channels at z*128)
  int nb_total;    // valid B channels over all y tiles (NB == 1) -- B tile = CB channels at y*CB
  int Dz, Dy, Dx;
  int Sz, Sy, Sx;  // stem
  const bf16* a_src;
  long long a_pitch;
  BnSrc bnA;       // used when A transform is T_BNRELU (channel index = z*128 + a)
  const bf16* b_src;
  long long b_pitch;
  BnSrc bnB;       // used when B transform is T_BNRELU (channel index = y*CB + b)
  float* dw;
  long long so_a, so_b, so_j;  // element strides of the gradient for (a, b, tap j); tap index = y*NB + j when NB > 1
  int cin_real;    // stem
  int stages;
  int NP;          // stem: k-block PAIRS handled per CTA (A tile = NP x 128 "channels", NP accumulators); else 1
  // Ordered reduction of the voxel split (deterministic weight gradients): slot_stride > 0 -> CTA x of the split WRITES its
  // partial result (plain stores) to dw + x * slot_stride instead of adding it to dw with floating-point atomics, and a tail
  // kernel sums the slots in index order (encoder.cu, wgrad_reduce_kernel).  0 -> atomics into dw (run-dependent last bits).
  long long slot_stride;
  // TMA path of the B (gradient) operand: tma_b != 0 -> the kernel's CUtensorMap argument describes b_src as the 5-D tensor
  // [N][Dz][Dy][Dx][C] and a 128-voxel tile of 32 channels is the box (32 channels, bx, by, bz, bn) -- see wgrad_tma_box() --
  // written with the 64-byte swizzle: 128 rows of 64 B = the canonical SWIZZLE_64B MN-major tcgen05 operand (K = voxel rows:
  // 8-row atoms 512 B apart = SBO; the next 32 channels / the next tap: one 8 KB group further = LBO).  A shifted tile (3x3x3
  // taps) is the same box at shifted coordinates; out-of-volume voxels are zero-filled by the TMA unit.
  // tma_b == 2 (3x3x3, tile = bx x by voxels of ONE z slice, bx a multiple of 8): the three dy taps of a (dz, dx) pair are row
  // windows of ONE box of by + 2 rows (y0 - 1 .. y0 + by), bx rows = a multiple of the 512-byte swizzle period apart: 3 boxes per
  // stage instead of 9, so the ring holds 3 stages instead of 2 (this kernel is bound by the fill latency of its ring).
  int tma_b;
  int bx, by, bz, bn;
  // Stem (WA_STEM_PAIR): a_bf16 != 0 -> a_src already holds bf16 (1: TMA allowed, 2: register path only -- tests); (the trunk keeps a bf16 copy of the space-to-depth image for
  // this kernel; otherwise the activation format, converted in the producers).  tma_a != 0 (needs a_bf16 and tma_b) -> the A
  // operand arrives by TMA as well: the kernel's second CUtensorMap views the image as [N][Sz][Sy][Sx-3][64], dimension x
  // strided by ONE 16-channel cell (32 B) under a 64-element row -- overlapping rows, so the 128 "channels" of a k-block pair
  // are four boxes (32 elements, bx, by, bz, bn) at (32 g mod 64, x0, y0 + dy + g / 2, z0 + dz, n0) in the same SWIZZLE_64B
  // layout as the gradient tiles.  No producer warp touches the data: one thread issues 4 NP + 2 boxes per voxel tile.
  // tma_a == 2 (tile = bx x by voxels of ONE z slice, bx a multiple of 8): the four k-blocks (dz, dy = 0..3) of a CTA read rows
  // y0 + dy .. of the same slice, so ONE box of by + 3 rows per 32-element half (instead of four boxes of by rows per half) feeds
  // all of them: k-block dy's operand starts bx rows (bx x 64 B, a multiple of the 512-byte swizzle period) further into the box.
  // The MMA's M = 128 is then (dy = 0..3) x 32 elements of one half -- LBO = bx x 64 B -- and accumulator j holds half j.
  // WA_LINEAR with a BN+ReLU A operand (1x1x1 / 3x3x3 weight gradients): tma_a == 1 (needs tma_b) -> the RAW activation tile
  // arrives by TMA as well (4 boxes of 32 channels, SWIZZLE_64B, straight into the operand stage; issued by warp 9 as soon as the
  // stage is free, i.e. up to `stages - 1` tiles ahead), and the 256 producer threads apply BN+ReLU -> bf16 IN PLACE
  // (LDS.128 -> 8 FMAs -> STS.128 to the same cell).  No global load sits in the producers' instruction stream any more: with
  // register staging the tile period was one exposed HBM round trip (ring depth 1, 2 or 3 gave the same 105-118 us per block-1
  // launch, profiles/scripts/microbench_wgrad.py).
  int a_bf16;
  int tma_a;
};

constexpr int TMA_GROUP_BYTES = TILE_ROWS * 64;   // one TMA box: 128 voxel rows x 32 bf16 channels

// 3x3x3, tma_b == 2: bytes of the three (by + 2)-row boxes (one per dx) of a stage
__host__ __device__ inline uint32_t wgrad_tap_halo_bytes(int bx, int by) { return 3u * (uint32_t)bx * (uint32_t)(by + 2) * 64u; }
// stem, tma_a == 2: bytes of the two (by + 3)-row boxes of a stage
__host__ __device__ inline uint32_t wgrad_stem_halo_bytes(int bx, int by) { return 2u * (uint32_t)bx * (uint32_t)(by + 3) * 64u; }

// Box decomposition of a tile of 128 consecutive voxels of an [N][Dz][Dy][Dx] volume (x fastest).  Returns false when a tile is
// not a box (dims that do not chain-divide 128): the kernel then keeps the register path.
__host__ __device__ inline bool wgrad_tma_box(int Dz, int Dy, int Dx, int& bx, int& by, int& bz, int& bn) {
  int rem = TILE_ROWS;
  if (Dx >= rem) { if (Dx % rem) return false; bx = rem; rem = 1; } else { if (rem % Dx) return false; bx = Dx; rem /= Dx; }
  if (Dy >= rem) { if (Dy % rem) return false; by = rem; rem = 1; } else { if (rem % Dy) return false; by = Dy; rem /= Dy; }
  if (Dz >= rem) { if (Dz % rem) return false; bz = rem; rem = 1; } else { if (rem % Dz) return false; bz = Dz; rem /= Dz; }
  bn = rem;
  return true;
}

__host__ __device__ inline uint32_t wgrad_smem_layout(int CB, int NB, int stages, int NP, uint32_t* offs /*[4]*/, bool tma_b = false,
                                                      uint32_t a_halo_bytes = 0, uint32_t b_halo_bytes = 0) {
  uint32_t o = 0;
  offs[0] = o; o += 192;                          // barriers (full[6], empty[6], accum, tmem ptr, raw[6] at +128) + debug slot
  offs[1] = o; o += stages * TILE_ROWS * 16;      // rowinfo per stage
  offs[2] = o; o += 2u * 128 * 4 + 2u * CB * 4;   // coefA (scale, shift) [128], coefB [CB]
  o = (o + 127u) & ~127u;
  offs[3] = o;
  uint32_t stage = 16u * NP * PLANE_BYTES + (uint32_t)NB * (CB / 8) * PLANE_BYTES;
  // TMA mode: [B groups (swizzled, the stage base is 1024-byte aligned at run time)][A planes], stage size a multiple of 1 KB
  if (tma_b) stage = ((b_halo_bytes ? b_halo_bytes : (uint32_t)NB * (CB / 32) * TMA_GROUP_BYTES) + (a_halo_bytes ? a_halo_bytes : 16u * NP * PLANE_BYTES) + 1023u) & ~1023u;
  const uint32_t ring = stages * stage + (tma_b ? 1024u : 0u);
  const uint32_t epi = 32u * 132u * 4u;           // the epilogue's transpose staging reuses the ring
  return o + (ring > epi ? ring : epi);
}
__host__ __device__ inline uint32_t wgrad_stage_bytes(int CB, int NB, int NP, bool tma_b, uint32_t a_halo_bytes = 0, uint32_t b_halo_bytes = 0) {
  if (tma_b) return ((b_halo_bytes ? b_halo_bytes : (uint32_t)NB * (CB / 32) * TMA_GROUP_BYTES) + (a_halo_bytes ? a_halo_bytes : 16u * NP * PLANE_BYTES) + 1023u) & ~1023u;
  return 16u * NP * PLANE_BYTES + (uint32_t)NB * (CB / 8) * PLANE_BYTES;
}

// Fill `planes` chunk planes of one operand tile. lane -> (chunk within a group of G, row sub-index).
// All loads of a group are issued before the first use (8 or 4 independent 128-bit loads in flight per thread).
// IN_F16: storage format of the source; the tile is always written as bf16 (weight-gradient GEMMs run in bf16).
template <int TRANS, bool SHIFTED, bool IN_F16>
MMNN_DEVINL void produce_planes(uint32_t sdst, int planes, const bf16* src, long long pitch, const int4* rowinfo, int warp,
                                int lane, int dz, int dy, int dx, long long delta, int Dz, int Dy, int Dx,
                                const float* scale, const float* shift) {
  const int G = planes >= 8 ? 8 : 4;
  const int gshift = planes >= 8 ? 3 : 2;
  const int rsub = lane >> gshift;
  const int rpp = 32 >> gshift;
  const int npass = (TILE_ROWS / rpp) / PRODUCER_WARPS;
  for (int grp = 0; grp < (planes + G - 1) / G; ++grp) {
    const int chunk = grp * G + (lane & (G - 1));
    if (chunk >= planes) continue;  // planes is a multiple of 4 but not necessarily of G
    uint4 regs[MAX_PASSES];
    uint32_t okmask = 0;
#pragma unroll
    for (int ps = 0; ps < MAX_PASSES; ++ps) {
      regs[ps] = make_uint4(0, 0, 0, 0);
      if (ps < npass) {
        const int r = (warp + ps * PRODUCER_WARPS) * rpp + rsub;
        const int4 ri = rowinfo[r];
        bool ok = ri.y > -1000;
        if (SHIFTED) {
          const int zz = ri.y + dz, yy = ri.z + dy, xx = ri.w + dx;
          ok = ok && zz >= 0 && zz < Dz && yy >= 0 && yy < Dy && xx >= 0 && xx < Dx;
        }
        if (ok) {
          regs[ps] = ldg16(src + ((long long)ri.x + delta) * pitch + chunk * 8);
          okmask |= 1u << ps;
        }
      }
    }
    float sc[8], sh[8];
    if (TRANS == T_BNRELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { sc[e] = scale[chunk * 8 + e]; sh[e] = shift[chunk * 8 + e]; }
    }
#pragma unroll
    for (int ps = 0; ps < MAX_PASSES; ++ps) {
      if (ps < npass) {
        const int r = (warp + ps * PRODUCER_WARPS) * rpp + rsub;
        uint4 v = regs[ps];
        if ((okmask >> ps) & 1u) {
          if (TRANS == T_BNRELU) apply_bnrelu8<IN_F16, false>(v, sc, sh);
          else convert8<IN_F16, false>(v);
        }
        sts16(sdst + chunk * PLANE_BYTES + r * 16, v);
      }
    }
  }
}

// Two-phase variant of produce_planes for tiles of up to 16 planes: load_planes issues every 128-bit load of the tile
// into registers (up to 8 per thread), store_planes transforms and stores them.  The weight-gradient kernel issues the
// loads of ALL operands of a stage before the first store: one L2 round trip per voxel tile instead of one per group.
template <bool SHIFTED>
MMNN_DEVINL uint32_t load_planes(uint4 (&regs)[2 * MAX_PASSES], int planes, const bf16* src, long long pitch, const int4* rowinfo,
                                 int warp, int lane, int dz, int dy, int dx, long long delta, int Dz, int Dy, int Dx,
                                 long long linear_base = -1) {
  const int G = planes >= 8 ? 8 : 4;
  const int gshift = planes >= 8 ? 3 : 2;
  const int rsub = lane >> gshift;
  const int rpp = 32 >> gshift;
  const int npass = (TILE_ROWS / rpp) / PRODUCER_WARPS;
  uint32_t okmask = 0;
#pragma unroll
  for (int grp = 0; grp < 2; ++grp) {
    const int chunk = grp * G + (lane & (G - 1));
#pragma unroll
    for (int ps = 0; ps < MAX_PASSES; ++ps) {
      regs[grp * MAX_PASSES + ps] = make_uint4(0, 0, 0, 0);
      if (ps < npass && chunk < planes) {
        const int r = (warp + ps * PRODUCER_WARPS) * rpp + rsub;
        const int4 ri = rowinfo[r];
        bool ok = ri.y > -1000;
        if (SHIFTED) {
          const int zz = ri.y + dz, yy = ri.z + dy, xx = ri.w + dx;
          ok = ok && zz >= 0 && zz < Dz && yy >= 0 && yy < Dy && xx >= 0 && xx < Dx;
        }
        if (ok) {
          const long long vox = linear_base >= 0 ? linear_base + r : (long long)ri.x;
          regs[grp * MAX_PASSES + ps] = ldg16(src + (vox + delta) * pitch + chunk * 8);
          okmask |= 1u << (grp * MAX_PASSES + ps);
        }
      }
    }
  }
  return okmask;
}

// load_planes for a tile of 128 CONSECUTIVE voxels starting at m0 (no row table: row r is voxel m0 + r, valid while < M)
MMNN_DEVINL uint32_t load_planes_linear(uint4 (&regs)[2 * MAX_PASSES], int planes, const bf16* src, long long pitch, long long m0,
                                        long long M, int warp, int lane) {
  const int G = planes >= 8 ? 8 : 4;
  const int gshift = planes >= 8 ? 3 : 2;
  const int rsub = lane >> gshift;
  const int rpp = 32 >> gshift;
  const int npass = (TILE_ROWS / rpp) / PRODUCER_WARPS;
  uint32_t okmask = 0;
#pragma unroll
  for (int grp = 0; grp < 2; ++grp) {
    const int chunk = grp * G + (lane & (G - 1));
#pragma unroll
    for (int ps = 0; ps < MAX_PASSES; ++ps) {
      regs[grp * MAX_PASSES + ps] = make_uint4(0, 0, 0, 0);
      if (ps < npass && chunk < planes) {
        const long long m = m0 + (warp + ps * PRODUCER_WARPS) * rpp + rsub;
        if (m < M) {
          regs[grp * MAX_PASSES + ps] = ldg16(src + m * pitch + chunk * 8);
          okmask |= 1u << (grp * MAX_PASSES + ps);
        }
      }
    }
  }
  return okmask;
}

// OUT_F16: format the tile is written in (the weight-gradient MMA takes an fp16 A operand next to the bf16 gradient operand, so
// forward activations need no conversion).  IN_F16 && OUT_F16 && BN+ReLU: `scale` points to an H2Coef table (one entry per
// channel pair) and the transform is the 12-HFMA2 fast path of the forward producers.
template <int TRANS, bool IN_F16, bool OUT_F16 = false>
MMNN_DEVINL void store_planes(const uint4 (&regs)[2 * MAX_PASSES], uint32_t okmask, uint32_t sdst, int planes, int warp, int lane,
                              const float* scale, const float* shift) {
  constexpr bool H2 = TRANS == T_BNRELU && IN_F16 && OUT_F16;
  const int G = planes >= 8 ? 8 : 4;
  const int gshift = planes >= 8 ? 3 : 2;
  const int rsub = lane >> gshift;
  const int rpp = 32 >> gshift;
  const int npass = (TILE_ROWS / rpp) / PRODUCER_WARPS;
#pragma unroll
  for (int grp = 0; grp < 2; ++grp) {
    const int chunk = grp * G + (lane & (G - 1));
    if (chunk >= planes) continue;
    float sc[8], sh[8];
    H2Coef hc[4];
    if (H2) {
      const H2Coef* tab = reinterpret_cast<const H2Coef*>(scale);
#pragma unroll
      for (int i = 0; i < 4; ++i) hc[i] = tab[chunk * 4 + i];
    } else if (TRANS == T_BNRELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { sc[e] = scale[chunk * 8 + e]; sh[e] = shift[chunk * 8 + e]; }
    }
#pragma unroll
    for (int ps = 0; ps < MAX_PASSES; ++ps) {
      if (ps < npass) {
        const int r = (warp + ps * PRODUCER_WARPS) * rpp + rsub;
        uint4 v = regs[grp * MAX_PASSES + ps];
        if ((okmask >> (grp * MAX_PASSES + ps)) & 1u) {
          if (H2) apply_bnrelu8_h2(v, hc);
          else if (TRANS == T_BNRELU) apply_bnrelu8<IN_F16, OUT_F16>(v, sc, sh);
          else convert8<IN_F16, OUT_F16>(v);
        }
        sts16(sdst + chunk * PLANE_BYTES + r * 16, v);
      }
    }
  }
}

#ifndef MMNN_WGRAD_A_F16
#define MMNN_WGRAD_A_F16 0
#endif
constexpr int WGRAD_THREADS = ENGINE_THREADS + 32;   // + warp 9: the TMA issuer of the raw-A mode (idle otherwise)
constexpr int WGRAD_TMA_WARP = MMA_WARP + 1;
#ifdef MMNN_WGRAD_TIMING
// debug build (-DMMNN_WGRAD_TIMING): cycle counters summed over CTAs -- [0] MMA warp waiting for a full stage, [1] MMA issue,
// [2] producer warp 0 waiting for an empty stage, [3] its A stores, [4] fill latency free->full (warp 2), [5] whole kernel, [6] CTAs, [7] tiles
static __device__ unsigned long long g_wgrad_dbg[12];   // [8] fence + elect, [9] the MMAs, [10] commits, [11] syncwarp + loop
#define WG_T(var) const long long var = clock64()
#define WG_ADD(i, v) do { if (lane == 0) atomicAdd(&g_wgrad_dbg[i], (unsigned long long)(v)); } while (0)
#else
#define WG_T(var)
#define WG_ADD(i, v)
#endif
template <int AMODE, int ATRANS, int BTRANS, int EMODE>
__global__ void __launch_bounds__(WGRAD_THREADS) conv_wgrad_kernel(const __grid_constant__ WgradParams p,
                                                                    const __grid_constant__ CUtensorMap tmb,
                                                                    const __grid_constant__ CUtensorMap tma) {
  // NEGATIVE RESULT (round 2, B200): the instruction descriptor of tcgen05 kind::f16 has separate A / B format fields, but an
  // fp16 A operand (forward activations, no conversion, packed-half BN+ReLU) next to the bf16 B operand (gradients) raises
  // "illegal instruction" on sm_100a -- both operands must have the same format.  -DMMNN_WGRAD_A_F16=1 builds that variant
  // (kept to document the experiment); the default converts the activation tile to bf16 in the producers.
  constexpr bool A_F16 = kActF16 && (MMNN_WGRAD_A_F16 != 0);
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t offs[4];
  const int NP = p.NP < 1 ? 1 : p.NP;
  const bool tma_b = p.tma_b != 0;
  const bool tma_a = AMODE == WA_STEM_PAIR && tma_b && p.tma_a != 0;
  const bool halo = tma_a && p.tma_a == 2;
  const bool araw = AMODE == WA_LINEAR && ATRANS == T_BNRELU && tma_b && p.tma_a == 1;      // raw A by TMA, transformed in place
  const uint32_t a_grp = halo ? (uint32_t)p.bx * (uint32_t)(p.by + 3) * 64u : (uint32_t)TMA_GROUP_BYTES;   // one 32-element half / group
  const uint32_t halo_bytes = halo ? 2u * a_grp : 0u;
  const bool bhalo = tma_b && p.tma_b == 2 && p.NB == 9 && p.CB == 32 && AMODE == WA_LINEAR;
  const uint32_t b_box = bhalo ? (uint32_t)p.bx * (uint32_t)(p.by + 2) * 64u : 0u;      // one (by + 2)-row box of 32 channels
  wgrad_smem_layout(p.CB, p.NB, p.stages, NP, offs, tma_b, halo_bytes, 3u * b_box);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + offs[0];
  const uint32_t bar_empty = bar_full + 8 * 6;
  const uint32_t bar_accum = bar_full + 8 * 12;
  const uint32_t bar_raw = bar_full + 128;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + offs[0] + 8 * 13);
  int4* rowinfo_all = reinterpret_cast<int4*>(smem + offs[1]);
  float* coefA = reinterpret_cast<float*>(smem + offs[2]);
  float* coefB = coefA + 256;
  const int bplanes = p.CB / 8;
  const uint32_t a_bytes = 16u * NP * PLANE_BYTES;
  const uint32_t bgroups = (uint32_t)p.CB / 32u;                        // TMA mode: 32-channel groups per B tile
  const uint32_t bt_bytes = tma_b ? bgroups * (uint32_t)TMA_GROUP_BYTES : (uint32_t)bplanes * PLANE_BYTES;
  const uint32_t stage_bytes = wgrad_stage_bytes(p.CB, p.NB, NP, tma_b, halo_bytes, 3u * b_box);
  const uint32_t stage0 = tma_b ? ((sbase + offs[3] + 1023u) & ~1023u) : sbase + offs[3];
  // operand offsets inside a stage: register path [A][B]; TMA path [B][A] (the swizzled B groups need the 1 KB alignment)
  const uint32_t a_off = bhalo ? 3u * b_box : tma_b ? p.NB * bt_bytes : 0u, b_off = tma_b ? 0u : a_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  WG_T(wg_t_start);
  const int S = p.stages;
  const int ntiles_total = (p.M + TILE_ROWS - 1) / TILE_ROWS;
  const int split = gridDim.x;
  const int t_begin = (int)((long long)ntiles_total * blockIdx.x / split);
  const int t_end = (int)((long long)ntiles_total * (blockIdx.x + 1) / split);
  const int nt = t_end - t_begin;
  if (nt <= 0) return;
  const int ytile = blockIdx.y, ztile = blockIdx.z;
  const int vps = p.Dz * p.Dy * p.Dx;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (p.NB > NP ? p.NB : NP) * p.CB) tmem_cols <<= 1;

  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        // producers + the expect_tx arrival of the TMA issuer; all-TMA stem: the issuer alone
        mbar_init(bar_full + 8 * s, tma_a ? 1 : NUM_PRODUCER_THREADS + (tma_b ? 1 : 0));
        mbar_init(bar_empty + 8 * s, 1);
        mbar_init(bar_raw + 8 * s, 1);
      }
      mbar_init(bar_accum, 1);
      fence_mbar_init();
      if (tma_b) tma_prefetch_desc(&tmb);
      if (tma_a || araw) tma_prefetch_desc(&tma);
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
  }
  pdl_wait();   // nothing above touches global memory
  pdl_trigger();   // AFTER the wait: a dependent that starts early may rely on everything before THIS kernel being complete
  if (ATRANS == T_BNRELU) {
    if (A_F16) {
      H2Coef* coefH = reinterpret_cast<H2Coef*>(coefA);   // 64 channel pairs x 16 B: same 1 KB as the fp32 scale / shift table
      for (int j = tid; j < 64; j += ENGINE_THREADS) {
        float s[2] = {0.f, 0.f}, t[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = ztile * 128 + 2 * j + h;
          if (ch < p.na_total) {
            float mean, rstd;
            bn_mean_rstd(p.bnA, ch, mean, rstd);
            s[h] = p.bnA.gamma[ch] * rstd;
            t[h] = p.bnA.beta[ch] - mean * s[h];
          }
        }
        H2Coef c;
        c.s_hi = __floats2half2_rn(s[0], s[1]);
        c.t_hi = __floats2half2_rn(t[0], t[1]);
        const float2 sh = __half22float2(c.s_hi), th = __half22float2(c.t_hi);
        c.s_lo = __floats2half2_rn(s[0] - sh.x, s[1] - sh.y);
        c.t_lo = __floats2half2_rn(t[0] - th.x, t[1] - th.y);
        coefH[j] = c;
      }
    } else
    for (int c = tid; c < 128; c += ENGINE_THREADS) {
      const int ch = ztile * 128 + c;
      float s = 0.f, t = 0.f;
      if (ch < p.na_total) {
        float mean, rstd;
        bn_mean_rstd(p.bnA, ch, mean, rstd);
        s = p.bnA.gamma[ch] * rstd;
        t = p.bnA.beta[ch] - mean * s;
      }
      coefA[c] = s; coefA[128 + c] = t;
    }
  }
  if (BTRANS == T_BNRELU) {
    for (int c = tid; c < p.CB; c += ENGINE_THREADS) {
      const int ch = ytile * p.CB + c;
      float s = 0.f, t = 0.f;
      if (ch < p.nb_total) {
        float mean, rstd;
        bn_mean_rstd(p.bnB, ch, mean, rstd);
        s = p.bnB.gamma[ch] * rstd;
        t = p.bnB.beta[ch] - mean * s;
      }
      coefB[c] = s; coefB[p.CB + c] = t;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < PRODUCER_WARPS) {
    // valid plane counts (multiples of 4) of the A / B tiles handled by this CTA
    int aplanes = 16;
    if (AMODE == WA_LINEAR) { int rem = (p.na_total - ztile * 128) / 8; aplanes = rem < 16 ? rem : 16; }
    int bplanes_valid = bplanes;
    if (p.NB == 1) { int rem = (p.nb_total - ytile * p.CB) / 8; bplanes_valid = rem < bplanes ? rem : bplanes; }
    // ---- software-pipelined path (1x1x1 and 3x3x3 weight gradients, raw B operand): the B tiles need no transform, so
    // they go global -> shared with cp.async (zero-fill for rows / taps outside the volume) and never touch registers;
    // the loads of tile it+1 (A into registers, B asynchronously) are issued BEFORE tile it is transformed, stored
    // and handed to the MMA warp.  Before, every tile exposed one full L2 / HBM round trip (ncu r01f: 10 % of all
    // samples on the first use of the A loads, 288 threads at 31 % issue utilisation).
    // MEASURED SLOWER than the register-staged path below in a same-box A/B at configs[1] (3x3x3 weight gradients 2.56 vs
    // 2.37 ms, step 15.17 vs 14.94 ms; 2.63 ms when the copies bypass L1 -- the 9 shifted tiles share their rows), so it
    // is OFF unless MMNN_WGRAD_PIPED=1 (kept as a parity-tested experiment).
    const bool piped = !tma_b && AMODE == WA_LINEAR && BTRANS == T_NONE && S >= 2 && p.NB == 9 && bplanes == 4 && p.NP == -1;   // NP == -1: opt-in experiment switch (MMNN_WGRAD_PIPED=1)
    if (araw) {
      // ---- raw A by TMA (issued by warp 9 below), BN+ReLU -> bf16 in place.  Thread -> logical 16-byte chunk c = tid & 3 of every
      // 32-channel group and rows r0, r0 + 64: its 32 channels never change, so their scale / shift live in registers; a warp touches
      // 8 rows x 64 B = one 512-byte swizzle atom per access (conflict-free), the chunk's physical position is c ^ ((row >> 1) & 3).
      const int c = tid & 3, r0 = tid >> 2;
      float sc[4][8], sh[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc[g][e] = coefA[g * 32 + c * 8 + e]; sh[g][e] = coefA[128 + g * 32 + c * 8 + e]; }
      const uint32_t o0 = (uint32_t)r0 * 64u + (uint32_t)((c ^ ((r0 >> 1) & 3)) * 16);
      const uint32_t o1 = (uint32_t)(r0 + 64) * 64u + (uint32_t)((c ^ (((r0 + 64) >> 1) & 3)) * 16);
      for (int it = 0; it < nt; ++it) {
        const int s = it % S;
        WG_T(w0);
        mbar_wait(bar_raw + 8 * s, (uint32_t)(it / S) & 1u, 15);
        WG_T(w1);
        const uint32_t sAs = stage0 + s * stage_bytes + a_off;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v0 = lds16(sAs + (uint32_t)g * (uint32_t)TMA_GROUP_BYTES + o0);
          uint4 v1 = lds16(sAs + (uint32_t)g * (uint32_t)TMA_GROUP_BYTES + o1);
          apply_bnrelu8<kActF16, false>(v0, sc[g], sh[g]);
          apply_bnrelu8<kActF16, false>(v1, sc[g], sh[g]);
          sts16(sAs + (uint32_t)g * (uint32_t)TMA_GROUP_BYTES + o0, v0);
          sts16(sAs + (uint32_t)g * (uint32_t)TMA_GROUP_BYTES + o1, v1);
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_full + 8 * s);
        WG_T(w2);
        if (warp == 1) { WG_ADD(2, w1 - w0); WG_ADD(3, w2 - w1); }
      }
    } else if (tma_a) {
      // ---- all-TMA stem: both operands are boxes; one thread feeds the ring, the other producer threads go straight to the epilogue
      if (tid == 4 * 32) {     // a warp without an epilogue role (the epilogue runs on warps 0-3)
        for (int it = 0; it < nt; ++it) {
          const int s = it % S;
          mbar_wait(bar_empty + 8 * s, ((uint32_t)(it / S) & 1u) ^ 1u, 11);
          const long long m0 = (long long)(t_begin + it) * TILE_ROWS;
          const int n0 = (int)(m0 / vps);
          int rem = (int)(m0 - (long long)n0 * vps);
          const int z0 = rem / (p.Dy * p.Dx);
          rem -= z0 * p.Dy * p.Dx;
          const int y0 = rem / p.Dx, x0 = rem - (rem / p.Dx) * p.Dx;
          const uint32_t sBs = stage0 + s * stage_bytes + b_off, sAs = stage0 + s * stage_bytes + a_off;
          // (the A region of a stage is sized in padded chunk planes, the boxes fill 4 NP x 8 KB of it)
          mbar_arrive_expect_tx(bar_full + 8 * s, bt_bytes + (halo ? halo_bytes : 4u * NP * (uint32_t)TMA_GROUP_BYTES));
          for (uint32_t c = 0; c < bgroups; ++c)
            tma_load_5d(sBs + c * (uint32_t)TMA_GROUP_BYTES, &tmb, (int)c * 32, x0, y0, z0, n0, bar_full + 8 * s);
          if (halo) {      // this CTA's k-blocks are (dz = ztile, dy = 0..3): rows y0 .. y0 + by + 2 of slice z0 + ztile, two halves
            tma_load_5d(sAs, &tma, 0, x0, y0, z0 + ztile, n0, bar_full + 8 * s);
            tma_load_5d(sAs + a_grp, &tma, 32, x0, y0, z0 + ztile, n0, bar_full + 8 * s);
          } else
          for (int g = 0; g < 4 * NP; ++g) {     // group g: k-block kb = ztile*2*NP + g/2, elements 32 (g & 1) .. + 32 of its 64
            const int kb = ztile * 2 * NP + (g >> 1);
            tma_load_5d(sAs + (uint32_t)g * (uint32_t)TMA_GROUP_BYTES, &tma, (g & 1) * 32, x0, y0 + (kb & 3), z0 + (kb >> 2), n0,
                        bar_full + 8 * s);
          }
        }
      }
    } else if (tma_b && AMODE == WA_LINEAR) {
      // ---- TMA mode (default when the tile is a box of the volume): the gradient tiles arrive by cp.async.bulk.tensor, so
      // the 256 producer threads only build the A tile -- and do it software-pipelined: the 128-bit loads of tile it+1 are in
      // flight while tile it is transformed and stored (8 loads per thread, no row table, no per-tile barrier).
      const bf16* asrc = p.a_src + ztile * 128;
      auto issue_b = [&](int it, int s) {       // one thread: expect_tx + the boxes of this tile
        const long long m0 = (long long)(t_begin + it) * TILE_ROWS;
        const int n0 = (int)(m0 / vps);
        int rem = (int)(m0 - (long long)n0 * vps);
        const int z0 = rem / (p.Dy * p.Dx);
        rem -= z0 * p.Dy * p.Dx;
        const int y0 = rem / p.Dx, x0 = rem - (rem / p.Dx) * p.Dx;
        const uint32_t sBs = stage0 + s * stage_bytes + b_off;
        if (p.cin_real == 2 || p.cin_real == 3) { mbar_arrive_expect_tx(bar_full + 8 * s, 0u); return; }
        if (bhalo) {     // box c = taps (dz of this CTA, dx = 1 - c), rows y0 - 1 .. y0 + by
          mbar_arrive_expect_tx(bar_full + 8 * s, 3u * b_box);
          const int dz = -(ytile - 1);
          for (int c = 0; c < 3; ++c) tma_load_5d(sBs + (uint32_t)c * b_box, &tmb, 0, x0 + 1 - c, y0 - 1, z0 + dz, n0, bar_full + 8 * s);
          return;
        }
        mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)p.NB * bt_bytes);
        for (int j = 0; j < p.NB; ++j) {
          int dz = 0, dy = 0, dx = 0;
          if (p.NB > 1) { const int tap = ytile * p.NB + j; dz = -(tap / 9 - 1); dy = -((tap / 3) % 3 - 1); dx = -(tap % 3 - 1); }
          const int cbase = p.NB == 1 ? ytile * p.CB : 0;
          for (uint32_t c = 0; c < bgroups; ++c)     // channels beyond the tensor arrive as zeros (their columns are discarded)
            tma_load_5d(sBs + ((uint32_t)j * bgroups + c) * (uint32_t)TMA_GROUP_BYTES, &tmb, cbase + (int)c * 32, x0 + dx, y0 + dy, z0 + dz, n0,
                        bar_full + 8 * s);
        }
      };
      uint4 RA[2 * MAX_PASSES], RB[2 * MAX_PASSES];
      uint32_t okA = load_planes_linear(RA, aplanes, asrc, p.a_pitch, (long long)t_begin * TILE_ROWS, p.M, warp, lane), okB = 0;
      const int dbg = p.cin_real;     // ablation switches of profiles/scripts/microbench_wgrad.py (0 in the product): 1 = no A stores, 2 = no B boxes
      auto finish = [&](int it, const uint4 (&regs)[2 * MAX_PASSES], uint32_t ok) {
        const int s = it % S;
        WG_T(w0);
        mbar_wait(bar_empty + 8 * s, ((uint32_t)(it / S) & 1u) ^ 1u, 11);
        WG_T(w1);
        if (tid == 0) issue_b(it, s);
        if (dbg != 1 && dbg != 3) store_planes<ATRANS, kActF16, A_F16>(regs, ok, stage0 + s * stage_bytes + a_off, aplanes, warp, lane, coefA, coefA + 128);
        fence_proxy_async_smem();
        mbar_arrive(bar_full + 8 * s);
        WG_T(w2);
        if (warp == 1) { WG_ADD(2, w1 - w0); WG_ADD(3, w2 - w1); }
#ifdef MMNN_WGRAD_TIMING
        if (warp == 2) {     // fill latency: stage free -> stage full (this warp then lags the others a little)
          mbar_wait(bar_full + 8 * s, (uint32_t)(it / S) & 1u, 14);
          WG_ADD(4, clock64() - w1);
        }
#endif
      };
      for (int it = 0; it < nt; it += 2) {
        if (it + 1 < nt) okB = load_planes_linear(RB, aplanes, asrc, p.a_pitch, (long long)(t_begin + it + 1) * TILE_ROWS, p.M, warp, lane);
        finish(it, RA, okA);
        if (it + 1 < nt) {
          if (it + 2 < nt) okA = load_planes_linear(RA, aplanes, asrc, p.a_pitch, (long long)(t_begin + it + 2) * TILE_ROWS, p.M, warp, lane);
          finish(it + 1, RB, okB);
        }
      }
    } else if (piped) {
      auto issue = [&](int it, uint4 (&aregs)[2 * MAX_PASSES], uint32_t& aok) {
        const int s = it % S;
        mbar_wait(bar_empty + 8 * s, ((uint32_t)(it / S) & 1u) ^ 1u, 11);
        int4* rowinfo = rowinfo_all + s * TILE_ROWS;
        if (tid < TILE_ROWS) {
          const long long m = (long long)(t_begin + it) * TILE_ROWS + tid;
          int4 ri;
          if (m < p.M) {
            const int n = (int)(m / vps);
            int rem = (int)(m - (long long)n * vps);
            const int z = rem / (p.Dy * p.Dx);
            rem -= z * p.Dy * p.Dx;
            const int y = rem / p.Dx;
            ri.x = (int)m; ri.y = z; ri.z = y; ri.w = rem - y * p.Dx;
          } else {
            ri.x = 0; ri.y = -100000; ri.z = 0; ri.w = 0;
          }
          rowinfo[tid] = ri;
        }
        named_bar_sync(1, NUM_PRODUCER_THREADS);
        const uint32_t sB = stage0 + s * stage_bytes + b_off;
        aok = load_planes<false>(aregs, aplanes, p.a_src + ztile * 128, p.a_pitch, rowinfo, warp, lane, 0, 0, 0, 0, p.Dz, p.Dy, p.Dx);
        if (p.NB == 1) {
          const int rsub = lane >> 3;
#pragma unroll
          for (int grp = 0; grp < 2; ++grp) {
            const int chunk = grp * 8 + (lane & 7);
            if (chunk < bplanes_valid) {
#pragma unroll
              for (int ps = 0; ps < MAX_PASSES; ++ps) {
                const int r = (warp + ps * PRODUCER_WARPS) * 4 + rsub;
                const int4 ri = rowinfo[r];
                const bool ok = ri.y > -1000;
                cp_async16(sB + chunk * PLANE_BYTES + r * 16, p.b_src + ytile * p.CB + (long long)ri.x * p.b_pitch + chunk * 8, ok ? 16u : 0u);
              }
            }
          }
        } else {
          const int chunk = lane & 3, rsub = lane >> 2;
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            const int tap = ytile * 9 + j;
            const int dz = -(tap / 9 - 1), dy = -((tap / 3) % 3 - 1), dx = -(tap % 3 - 1);
            const long long delta = (long long)(dz * p.Dy + dy) * p.Dx + dx;
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
              const int r = (warp + ps * PRODUCER_WARPS) * 8 + rsub;
              const int4 ri = rowinfo[r];
              const int zz = ri.y + dz, yy = ri.z + dy, xx = ri.w + dx;
              const bool ok = ri.y > -1000 && zz >= 0 && zz < p.Dz && yy >= 0 && yy < p.Dy && xx >= 0 && xx < p.Dx;
              cp_async16_ca(sB + j * bt_bytes + chunk * PLANE_BYTES + r * 16,
                            p.b_src + (ok ? ((long long)ri.x + delta) * p.b_pitch : 0) + chunk * 8, ok ? 16u : 0u);
            }
          }
        }
        cp_async_commit();
      };
      auto finish = [&](int it, const uint4 (&aregs)[2 * MAX_PASSES], uint32_t aok, bool newer_pending) {
        const int s = it % S;
        store_planes<ATRANS, kActF16, A_F16>(aregs, aok, stage0 + s * stage_bytes, aplanes, warp, lane, coefA, coefA + 128);
        if (newer_pending) cp_async_wait<1>(); else cp_async_wait<0>();
        fence_proxy_async_smem();
        mbar_arrive(bar_full + 8 * s);
      };
      uint4 RA[2 * MAX_PASSES], RB[2 * MAX_PASSES];
      uint32_t okA = 0, okB = 0;
      issue(0, RA, okA);
      for (int it = 0; it < nt; it += 2) {
        if (it + 1 < nt) issue(it + 1, RB, okB);
        finish(it, RA, okA, it + 1 < nt);
        if (it + 1 < nt) {
          if (it + 2 < nt) issue(it + 2, RA, okA);
          finish(it + 1, RB, okB, it + 2 < nt);
        }
      }
    } else
    for (int it = 0; it < nt; ++it) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(bar_empty + 8 * s, ph ^ 1u, 11);
      int4* rowinfo = rowinfo_all + s * TILE_ROWS;
      if (tid < TILE_ROWS) {
        const long long m = (long long)(t_begin + it) * TILE_ROWS + tid;
        int4 ri;
        if (m < p.M) {
          const int n = (int)(m / vps);
          int rem = (int)(m - (long long)n * vps);
          const int z = rem / (p.Dy * p.Dx);
          rem -= z * p.Dy * p.Dx;
          const int y = rem / p.Dx;
          const int x = rem - y * p.Dx;
          ri.x = (int)m; ri.y = z; ri.z = y; ri.w = x;
          if (AMODE == WA_STEM_PAIR) ri.x = ((n * p.Sz + z) * p.Sy + y) * p.Sx + x;  // A-side index; B uses m (see below)
        } else {
          ri.x = 0; ri.y = -100000; ri.z = 0; ri.w = 0;
        }
        rowinfo[tid] = ri;
      }
      named_bar_sync(1, NUM_PRODUCER_THREADS);
      const uint32_t sA = stage0 + s * stage_bytes + a_off;
      const uint32_t sB = stage0 + s * stage_bytes + b_off;
      uint4 aregs[2 * MAX_PASSES];
      uint32_t aok = 0;
      if (AMODE == WA_LINEAR) {
        // loads of the A tile are issued now and consumed after the B loads have been issued as well
        aok = load_planes<false>(aregs, aplanes, p.a_src + ztile * 128, p.a_pitch, rowinfo, warp, lane, 0, 0, 0, 0, p.Dz, p.Dy, p.Dx);
      }
      uint4 sregs[4][MAX_PASSES];   // stem: up to 2 pairs = 4 k-blocks of 8 planes, all loads in flight before the stores
      uint32_t sok = 0;
      if (AMODE == WA_STEM_PAIR) {
        const int chunk = lane & 7, rsub = lane >> 3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g < 2 * NP) {
            const int kb = ztile * 2 * NP + g;
            const long long delta = (long long)((kb >> 2) * p.Sy + (kb & 3)) * p.Sx;
#pragma unroll
            for (int ps = 0; ps < MAX_PASSES; ++ps) {
              const int r = (warp + ps * PRODUCER_WARPS) * 4 + rsub;
              const int4 ri = rowinfo[r];
              const bool ok = ri.y > -1000;
              sregs[g][ps] = ok ? ldg16(p.a_src + ((long long)ri.x + delta) * p.a_pitch + chunk * 8) : make_uint4(0, 0, 0, 0);
              sok |= (uint32_t)ok << (g * MAX_PASSES + ps);
            }
          }
        }
      }
      if (tma_b) {
        // B operand by TMA: one elected thread issues NB x (valid planes) boxes of 128 voxels x 8 channels; rows outside the
        // volume (taps) or beyond M arrive as zeros.  The 256 producer threads only build the A tile.
        if (tid == 0) {
          const long long m0 = (long long)(t_begin + it) * TILE_ROWS;
          const int n0 = (int)(m0 / vps);
          int rem = (int)(m0 - (long long)n0 * vps);
          const int z0 = rem / (p.Dy * p.Dx);
          rem -= z0 * p.Dy * p.Dx;
          const int y0 = rem / p.Dx, x0 = rem - (rem / p.Dx) * p.Dx;
          mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)p.NB * bt_bytes);
          for (int j = 0; j < p.NB; ++j) {
            int dz = 0, dy = 0, dx = 0;
            if (p.NB > 1) { const int tap = ytile * p.NB + j; dz = -(tap / 9 - 1); dy = -((tap / 3) % 3 - 1); dx = -(tap % 3 - 1); }
            const int cbase = p.NB == 1 ? ytile * p.CB : 0;
            for (uint32_t c = 0; c < bgroups; ++c)     // channels beyond the tensor arrive as zeros (their columns are discarded)
              tma_load_5d(sB + ((uint32_t)j * bgroups + c) * (uint32_t)TMA_GROUP_BYTES, &tmb, cbase + (int)c * 32, x0 + dx, y0 + dy, z0 + dz, n0,
                          bar_full + 8 * s);
          }
        }
        if (AMODE == WA_LINEAR) store_planes<ATRANS, kActF16, A_F16>(aregs, aok, sA, aplanes, warp, lane, coefA, coefA + 128);
        if (AMODE == WA_STEM_PAIR) {
          const int chunk = lane & 7, rsub = lane >> 3;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < 2 * NP) {
#pragma unroll
              for (int ps = 0; ps < MAX_PASSES; ++ps) {
                const int r = (warp + ps * PRODUCER_WARPS) * 4 + rsub;
                uint4 v = sregs[g][ps];
                if (((sok >> (g * MAX_PASSES + ps)) & 1u) && !p.a_bf16) convert8<kActF16, A_F16>(v);
                sts16(sA + (g * 8 + chunk) * PLANE_BYTES + r * 16, v);
              }
            }
          }
        }
      } else if (p.NB == 1) {
        uint4 bregs[2 * MAX_PASSES];
        // stem: the rowinfo index is the space-to-depth row; the B rows are plain output voxels (linear index)
        const uint32_t bok = load_planes<false>(bregs, bplanes_valid, p.b_src + ytile * p.CB, p.b_pitch, rowinfo, warp, lane, 0, 0, 0,
                                                0, p.Dz, p.Dy, p.Dx,
                                                AMODE == WA_STEM_PAIR ? (long long)(t_begin + it) * TILE_ROWS : -1);
        if (AMODE == WA_LINEAR) store_planes<ATRANS, kActF16, A_F16>(aregs, aok, sA, aplanes, warp, lane, coefA, coefA + 128);
        if (AMODE == WA_STEM_PAIR) {
          const int chunk = lane & 7, rsub = lane >> 3;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < 2 * NP) {
#pragma unroll
              for (int ps = 0; ps < MAX_PASSES; ++ps) {
                const int r = (warp + ps * PRODUCER_WARPS) * 4 + rsub;
                uint4 v = sregs[g][ps];
                if (((sok >> (g * MAX_PASSES + ps)) & 1u) && !p.a_bf16) convert8<kActF16, A_F16>(v);
                sts16(sA + (g * 8 + chunk) * PLANE_BYTES + r * 16, v);
              }
            }
          }
        }
        store_planes<T_NONE, false>(bregs, bok, sB, bplanes_valid, warp, lane, nullptr, nullptr);
      } else if (p.NB == 9 && bplanes == 4) {
        // 9 shifted raw gradient tiles of 4 planes: B_j[v] = g[v - tap offset].  All 18 loads of this thread are
        // issued before the first store (one L2 round trip instead of nine).
        const int chunk = lane & 3, rsub = lane >> 2;
        uint4 regs[9][2];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int tap = ytile * 9 + j;
          const int dz = -(tap / 9 - 1), dy = -((tap / 3) % 3 - 1), dx = -(tap % 3 - 1);
          const long long delta = (long long)(dz * p.Dy + dy) * p.Dx + dx;
#pragma unroll
          for (int ps = 0; ps < 2; ++ps) {
            const int r = (warp + ps * PRODUCER_WARPS) * 8 + rsub;
            const int4 ri = rowinfo[r];
            const int zz = ri.y + dz, yy = ri.z + dy, xx = ri.w + dx;
            const bool ok = ri.y > -1000 && zz >= 0 && zz < p.Dz && yy >= 0 && yy < p.Dy && xx >= 0 && xx < p.Dx;
            regs[j][ps] = ok ? ldg16(p.b_src + ((long long)ri.x + delta) * p.b_pitch + chunk * 8) : make_uint4(0, 0, 0, 0);
          }
        }
        if (AMODE == WA_LINEAR) store_planes<ATRANS, kActF16, A_F16>(aregs, aok, sA, aplanes, warp, lane, coefA, coefA + 128);
#pragma unroll
        for (int j = 0; j < 9; ++j)
#pragma unroll
          for (int ps = 0; ps < 2; ++ps)
            sts16(sB + j * bt_bytes + chunk * PLANE_BYTES + ((warp + ps * PRODUCER_WARPS) * 8 + rsub) * 16, regs[j][ps]);
      } else {
        if (AMODE == WA_LINEAR) store_planes<ATRANS, kActF16, A_F16>(aregs, aok, sA, aplanes, warp, lane, coefA, coefA + 128);
        for (int j = 0; j < p.NB; ++j) {
          const int tap = ytile * p.NB + j;
          const int dz = -(tap / 9 - 1), dy = -((tap / 3) % 3 - 1), dx = -(tap % 3 - 1);
          const long long delta = (long long)(dz * p.Dy + dy) * p.Dx + dx;
          produce_planes<BTRANS, true, false>(sB + j * bt_bytes, bplanes, p.b_src, p.b_pitch, rowinfo, warp, lane, dz, dy, dx, delta,
                                       p.Dz, p.Dy, p.Dx, coefB, coefB + p.CB);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_full + 8 * s);
    }
  }
  if (warp < 4) {
    // ---- epilogue: TMEM lane = A channel
    mbar_wait(bar_accum, 0, 13);
    tc_fence_after();
    WG_T(wg_e0);
#ifdef MMNN_WGRAD_TIMING
    if (warp == 0) WG_ADD(11, wg_e0 - *reinterpret_cast<volatile long long*>(smem + offs[0] + 184));   // last commit -> accumulator complete
#endif
    const int a = warp * 32 + lane;
    const int nacc = (EMODE == WE_STEM) ? NP : p.NB;
    const bool slotted = p.slot_stride > 0;
    float* const dwb = p.dw + (long long)blockIdx.x * p.slot_stride;   // this CTA's slot (== dw when the split is reduced by atomics)
    for (int j = 0; j < nacc; ++j) {
      for (int cc = 0; cc < p.CB / 32; ++cc) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * p.CB + cc * 32), v);
        if (EMODE == WE_STRIDED) {
          // Transpose the 128(a) x 32(b) chunk through shared memory (the operand stages are free once the accumulator
          // is final) so that every warp instruction adds one 512-byte contiguous run of the gradient with 128-bit
          // vector atomics: 4x fewer atomic instructions than one RED per element.
          float* stg = reinterpret_cast<float*>(smem + offs[3]);      // [32][132]
          const int b0 = (p.NB == 1 ? ytile * p.CB : 0) + cc * 32;
          const int tap = p.NB == 1 ? 0 : ytile * p.NB + j;
          named_bar_sync(2, EPILOGUE_THREADS);                          // previous chunk fully drained
#pragma unroll
          for (int i = 0; i < 32; ++i) stg[i * 132 + a] = v[i];
          named_bar_sync(2, EPILOGUE_THREADS);
          const int ag4 = ztile * 128 + lane * 4;
          if (ag4 < p.na_total) {
            for (int i = warp; i < 32; i += 4) {
              if (b0 + i < p.nb_total) {
                const float4 val = *reinterpret_cast<const float4*>(stg + i * 132 + lane * 4);
                float* dst = dwb + (long long)ag4 + (long long)tap * p.so_j + (long long)(b0 + i) * p.so_b;
                if (slotted) *reinterpret_cast<float4*>(dst) = val;
                else atomicAdd(reinterpret_cast<float4*>(dst), val);
              }
            }
          }
        } else {
          // stem: a = (g, c64) with kb = ztile*2 + g, c64 = (dx, pz, py, px, ci); b = output channel
          // (halo mode: accumulator j = element half j of the four k-blocks dy = a / 32)
          const int kb = halo ? ztile * 4 + (a >> 5) : (ztile * NP + j) * 2 + (a >> 6), c = halo ? j * 32 + (a & 31) : (a & 63);
          const int kz = 2 * (kb >> 2) + ((c >> 3) & 1), ky = 2 * (kb & 3) + ((c >> 2) & 1), kx = 2 * (c >> 4) + ((c >> 1) & 1);
          const int ci = c & 1;
          if (kz < 7 && ky < 7 && kx < 7 && ci < p.cin_real) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int co = cc * 32 + i;
              float* dst = dwb + ((((long long)co * p.cin_real + ci) * 7 + kz) * 7 + ky) * 7 + kx;
              if (slotted) *dst = v[i];
              else atomicAdd(dst, v[i]);
            }
          }
        }
      }
    }
#ifdef MMNN_WGRAD_TIMING
    if (warp == 0) { WG_ADD(5, clock64() - wg_t_start); WG_ADD(6, 1); WG_ADD(7, nt); }
#endif
  } else if (warp == MMA_WARP) {
    const uint32_t idesc = make_idesc_ab(128, p.CB, 1, 1, A_F16, false);
    for (int it = 0; it < nt; ++it) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      WG_T(m0);
      mbar_wait(bar_full + 8 * s, ph, 12);
      WG_T(m1);
      WG_ADD(0, m1 - m0);
      tc_fence_after();
#ifdef MMNN_WGRAD_TIMING
      long long q0 = 0, q1 = 0, q2 = 0;
#endif
      if (elect_one()) {
#ifdef MMNN_WGRAD_TIMING
        q0 = clock64();
#endif
        const uint32_t sA = stage0 + s * stage_bytes + a_off;
        const uint32_t sB = stage0 + s * stage_bytes + b_off;
        // A: chunk planes, or (all-TMA stem) the same swizzled 32-channel groups as B
        const uint64_t ad0 = (tma_a || araw) ? make_smem_desc_sw(sA, halo ? (uint32_t)p.bx * 64u : (uint32_t)TMA_GROUP_BYTES, 512, 4u)
                                             : make_smem_desc(sA, 128, PLANE_BYTES);
        const uint32_t ak16 = (tma_a || araw) ? 1024u : 256u;
        // B: register path = SWIZZLE_NONE chunk planes (8-row K groups 128 B apart, 8-channel chunks one plane apart);
        //    TMA path = SWIZZLE_64B rows of 64 B (8-row atoms 512 B apart, 32-channel groups 8 KB apart)
        uint64_t bd0 = tma_b ? make_smem_desc_sw(sB, TMA_GROUP_BYTES, 512, 4u) : make_smem_desc(sB, 128, PLANE_BYTES);
        const uint32_t bk16 = tma_b ? 1024u : 256u;      // byte advance of the B start address per K step of 16 voxel rows
        const uint32_t acc0 = it > 0 ? 1u : 0u;
        if (AMODE == WA_STEM_PAIR) {
          // NP pairs share the B tile; pair pi reads its own 16 A planes and owns accumulator pi
          for (int pi = 0; pi < NP; ++pi) {
            const uint64_t ad = desc_advance(ad0, halo ? pi * a_grp : tma_a ? pi * 4 * TMA_GROUP_BYTES : pi * 16 * PLANE_BYTES);
            const uint32_t td = tmem_base + pi * p.CB;
            tc_mma_bf16(td, ad, bd0, idesc, acc0);
#pragma unroll
            for (int k16 = 1; k16 < TILE_ROWS / 16; ++k16)
              tc_mma_bf16(td, desc_advance(ad, k16 * ak16), desc_advance(bd0, k16 * bk16), idesc, 1u);
          }
        } else if (AMODE == WA_LINEAR && p.cin_real == 4) {
          // ablation: no MMA at all (fill pipeline alone)
        } else if (p.NB == 9 && p.CB == 32 && tma_b) {
          // swizzled operand: N must cover whole 32-channel groups -> three N = 96 MMAs (3 taps each) per K step
          const uint32_t idesc3 = make_idesc_ab(128, 96, 1, 1, A_F16, false);
          // halo boxes: MMA h = the taps dy = 1 - h of the three dx boxes (one box apart = LBO), rows (2 - h) * bx .. + 128
          if (bhalo) bd0 = make_smem_desc_sw(sB, b_box, 512, 4u);
#pragma unroll
          for (int h = 0; h < 3; ++h) {
            const uint32_t td = tmem_base + h * 96;
            const uint64_t bd = desc_advance(bd0, bhalo ? (uint32_t)(2 - h) * (uint32_t)p.bx * 64u : (uint32_t)h * 3u * TMA_GROUP_BYTES);
            tc_mma_bf16(td, ad0, bd, idesc3, acc0);
#pragma unroll
            for (int k16 = 1; k16 < TILE_ROWS / 16; ++k16)
              tc_mma_bf16(td, desc_advance(ad0, k16 * ak16), desc_advance(bd, k16 * 1024), idesc3, 1u);
          }
        } else if (p.NB == 9 && p.CB == 32) {
          // the 9 shifted gradient tiles are 36 consecutive chunk planes = ONE MN-major operand of N = 288 columns
          // (column = tap*32 + co, the accumulator layout the epilogue expects): two N = 144 MMAs per K step instead of
          // nine N = 32 ones, so the A tile is read from shared memory 2x instead of 9x per step
          const uint32_t idesc2 = make_idesc_ab(128, 144, 1, 1, A_F16, false);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t td = tmem_base + h * 144;
            const uint64_t bd = desc_advance(bd0, h * 18 * PLANE_BYTES);
            tc_mma_bf16(td, ad0, bd, idesc2, acc0);
#pragma unroll
            for (int k16 = 1; k16 < TILE_ROWS / 16; ++k16)
              tc_mma_bf16(td, desc_advance(ad0, k16 * 256), desc_advance(bd, k16 * 256), idesc2, 1u);
          }
        } else
        for (int j = 0; j < p.NB; ++j) {
          const uint32_t td = tmem_base + j * p.CB;
          tc_mma_bf16(td, ad0, bd0, idesc, acc0);
#pragma unroll
          for (int k16 = 1; k16 < TILE_ROWS / 16; ++k16)
            tc_mma_bf16(td, desc_advance(ad0, k16 * ak16), desc_advance(bd0, k16 * bk16), idesc, 1u);
          bd0 = desc_advance(bd0, bt_bytes);
        }
#ifdef MMNN_WGRAD_TIMING
        q1 = clock64();
#endif
        tc_commit(bar_empty + 8 * s);
        if (it == nt - 1) tc_commit(bar_accum);
#ifdef MMNN_WGRAD_TIMING
        q2 = clock64();
        if (it == nt - 1) *reinterpret_cast<volatile long long*>(smem + offs[0] + 184) = q2;     // spare bytes of the barrier block
        atomicAdd(&g_wgrad_dbg[8], (unsigned long long)(q0 - m1));
        atomicAdd(&g_wgrad_dbg[9], (unsigned long long)(q1 - q0));
        atomicAdd(&g_wgrad_dbg[10], (unsigned long long)(q2 - q1));
#endif
      }
      __syncwarp();
      WG_ADD(1, clock64() - m1);
    }
  }
  if (warp == WGRAD_TMA_WARP && araw && lane == 0) {
    // ---- TMA issuer of the raw-A mode: refills a stage the moment the MMAs that read it are done -- up to S - 1 tiles ahead of
    // the transform -- with the 4 activation boxes (-> raw[s]) and the gradient boxes (-> full[s], next to the 256 producer arrivals)
    for (int it = 0; it < nt; ++it) {
      const int s = it % S;
      mbar_wait(bar_empty + 8 * s, ((uint32_t)(it / S) & 1u) ^ 1u, 16);
      const long long m0 = (long long)(t_begin + it) * TILE_ROWS;
      const int n0 = (int)(m0 / vps);
      int rem = (int)(m0 - (long long)n0 * vps);
      const int z0 = rem / (p.Dy * p.Dx);
      rem -= z0 * p.Dy * p.Dx;
      const int y0 = rem / p.Dx, x0 = rem - (rem / p.Dx) * p.Dx;
      const uint32_t sBs = stage0 + s * stage_bytes + b_off, sAs = stage0 + s * stage_bytes + a_off;
      mbar_arrive_expect_tx(bar_raw + 8 * s, 4u * (uint32_t)TMA_GROUP_BYTES);
      for (int g = 0; g < 4; ++g)     // channels beyond the tensor arrive as zeros, beyond na_total their accumulator rows are discarded
        tma_load_5d(sAs + (uint32_t)g * (uint32_t)TMA_GROUP_BYTES, &tma, ztile * 128 + g * 32, x0, y0, z0, n0, bar_raw + 8 * s);
      if (bhalo) {
        mbar_arrive_expect_tx(bar_full + 8 * s, 3u * b_box);
        const int dz = -(ytile - 1);
        for (int c = 0; c < 3; ++c) tma_load_5d(sBs + (uint32_t)c * b_box, &tmb, 0, x0 + 1 - c, y0 - 1, z0 + dz, n0, bar_full + 8 * s);
      } else {
        mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)p.NB * bt_bytes);
        for (int j = 0; j < p.NB; ++j) {
          int dz = 0, dy = 0, dx = 0;
          if (p.NB > 1) { const int tap = ytile * p.NB + j; dz = -(tap / 9 - 1); dy = -((tap / 3) % 3 - 1); dx = -(tap % 3 - 1); }
          const int cbase = p.NB == 1 ? ytile * p.CB : 0;
          for (uint32_t c = 0; c < bgroups; ++c)
            tma_load_5d(sBs + ((uint32_t)j * bgroups + c) * (uint32_t)TMA_GROUP_BYTES, &tmb, cbase + (int)c * 32, x0 + dx, y0 + dy, z0 + dz, n0,
                        bar_full + 8 * s);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mmnn
