// Brick-mode 3x3x3 convolution (forward and data-gradient) for the large dense blocks.
//
// The gather engine (engine.cuh) rebuilds the A operand once per tap: 27 loads + 27 BN/ReLU transforms per input
// element -- it is instruction-issue bound (ncu: 71 % issue slots busy, 4.6 % tensor pipe; profiles/r01_*).
// Here a persistent CTA stages the input HALO BRICK of one output tile (1 x 16 x 8 voxels -> 3 x 18 x 10 voxel slots)
// in shared memory ONCE (one load + one transform per element, 4.2x the tile instead of 27x), 32 channels (4 chunk
// planes) per brick buffer, in the chunk-plane layout
//      brick[chunk][slot]  (16-byte cells, slot = (z'*18 + y')*10 + x')
// and every tap is just a different START ADDRESS of the same SWIZZLE_NONE K-major descriptor:
//      start = plane0 + ((dz*18 + dy)*10 + dx)*16,  SBO (next 8-row group = next y) = 10*16 B,  LBO = plane stride.
// Warp roles (448 threads = 14 warps), fprop: 0-7 producers, 8 MMA issuer + TMEM owner, 9 weight loader (cp.async.bulk
// ring), 10-13 epilogue (TMEM double-buffered, so the epilogue of tile t overlaps the MMAs of tile t+1).
// dgrad (32 input channels, 128 output columns): the producer is light and the mask + statistics epilogue is the
// critical path (ncu r01f: MMA warp waits on acc_empty, 4 epilogue warps 87 % busy at 16 % issue rate), so the roles
// are 0-3 producers, 4 MMA, 5 loader, 6-13 epilogue: two warps per TMEM lane quarter, each taking half the columns.
//
// Tile pairs (template TP = 2, experiment switch MMNN_BRICK_TP=2): a CTA keeps the bricks of TWO tiles in shared memory
// and issues the MMAs of both against each weight stage, halving the weight re-fetch from L2 (221 KB per tile; ncu
// r01c: 925 MB L2->SM per layer at 5.5 TB/s, half of it weights).  Measured: no gain (dgrad) / slower (forward), so
// it is off by default -- what bounds the N = 32 forward is the 4 KB A-operand read per MMA from shared memory
// (6 912 + 1 728 of the ~11 k shared-memory wavefronts per tile; the shared pipe saturates near 70 %).
#pragma once
#include "engine.cuh"

namespace mmnn {

constexpr int BR_TY = 16, BR_TX = 8, BR_HY = BR_TY + 2, BR_HX = BR_TX + 2;
constexpr int BR_SLOTS = 3 * BR_HY * BR_HX;            // 540
#ifdef MMNN_BRICK_TEST_ALIGNED   // timing experiment only (wrong results): every A core matrix 128-byte aligned
constexpr int BR_PLANE = 8704;
#else
constexpr int BR_PLANE = BR_SLOTS * 16 + 16;           // 8656 B: odd multiple of 16 -> conflict-free chunk planes
#endif
constexpr int BR_THREADS = 448;          // forward: 14 warps
__host__ __device__ constexpr int brick_threads(bool grad) { return 448; }   // (16 epilogue + 2 producer warps for dgrad, 640 threads, was tried: 105 -> 149 us)
// Weight ring: forward (2 KB per tap and 32-channel buffer) 4 stages of 9 taps (one dz plane), data gradient (8 KB per
// tap) 6 stages of 3 taps (one dx row).  The MMA warp pays ~280 cycles of wait / fence / commit per ring stage
// (measured: 36 stages per tile instead of 18 cost +5 k cycles per tile), so forward stages carry as many taps as fit.
// With tile pairs (TP = 2, below) the data gradient keeps 4 brick buffers, so its ring shrinks to 3 stages.
// With tile pairs the data gradient uses 9 stages of ONE tap (8 KB): same 72 KB, three times the look-ahead.
__host__ __device__ constexpr int brick_btaps(bool grad, int tp) { return grad ? (tp == 2 ? 1 : 3) : 9; }
__host__ __device__ constexpr int brick_bstages(bool grad, int tp) { return grad ? (tp == 2 ? 9 : 6) : 4; }

struct BrickParams {
  int B, Dz, Dy, Dx;
  int CH;        // A channels per tap: 128 (fprop) or 32 (dgrad)
  int NT;        // output columns: 32 (fprop) or 128 (dgrad)
  int tap_sign;
  const bf16* a_src;
  long long a_pitch;
  BnSrc bnA;
  const bf16* b_packed;   // [tap*NH + h][planes][NT][8]
  bf16* out;
  long long out_pitch;
  const float* colscale;
  double* st_sum;
  double* st_sq;
  const bf16* e_src;
  long long e_pitch;
  BnSrc bnE;
};

constexpr int BR_PH = 4;         // chunk planes (of 8 channels) per brick buffer
__host__ __device__ inline int brick_nbuf(int CH, int tp) { return (CH >= 64 || tp == 2) ? 4 : 2; }
__host__ __device__ inline uint32_t brick_smem_layout(int CH, int NT, int tp, uint32_t* offs /*[6]*/) {
  const int PH = BR_PH;
  uint32_t o = 0;
  offs[0] = o; o += 256;                       // barriers + tmem ptr
  offs[1] = o; o += 2u * CH * 4 + 8u * CH;    // coefA (fp32 scale, shift) + packed half2 hi/lo table (16 B per channel pair)
  offs[2] = o; o += 4u * NT * 4;               // coefE
  offs[3] = o; o += 8u * NT * 4;               // red
  o = (o + 127u) & ~127u;
  offs[4] = o; o += (uint32_t)brick_nbuf(CH, tp) * PH * BR_PLANE;   // brick buffer ring
  o = (o + 127u) & ~127u;
  offs[5] = o; o += (uint32_t)(brick_bstages(CH < 64, tp) * brick_btaps(CH < 64, tp)) * PH * NT * 16;
  return o;
}

template <int TRANS, int EPI, bool GRAD, int TP>
__global__ void __launch_bounds__(brick_threads(GRAD), 1) conv3_brick_kernel(const __grid_constant__ BrickParams p) {
  constexpr bool OP_F16 = !GRAD && kActF16;
  constexpr bool E_F16 = kActF16;
  constexpr int NTHREADS = brick_threads(GRAD);
  constexpr int NPW = GRAD ? 4 : 8;                    // producer warps
  constexpr int NPT = NPW * 32;                        // producer threads
  constexpr int NEW = GRAD ? 8 : 4;                    // epilogue warps (dgrad: two per TMEM lane quarter, two 32-column chunks each)
  constexpr int NET = NEW * 32;
  constexpr int BR_MMA_WARP = NPW, BR_LOAD_WARP = NPW + 1, BR_EPI_WARP0 = NPW + 2;
  static_assert(BR_EPI_WARP0 + NEW == NTHREADS / 32, "warp roles must fill the CTA");
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t offs[6];
  brick_smem_layout(p.CH, p.NT, TP, offs);
  const uint32_t sbase = smem_u32(smem);
  constexpr int BR_BTAPS = brick_btaps(GRAD, TP), BR_BSTAGES = brick_bstages(GRAD, TP);
  constexpr int NBUF = (!GRAD || TP == 2) ? 4 : 2;   // brick buffers (GRAD launches have CH = 32, forward ones CH = 128: checked by the host)
  constexpr int LAG = NBUF == 4 ? 2 : 1;             // producer look-ahead: buffers whose copies are in flight while an older one is finished
  // barrier map (8 B each): brick_full[4] | brick_empty[4] | b_full[BSTAGES] | b_empty[BSTAGES] | acc_full[2] | acc_empty[2]
  constexpr int BF = 0, BE = 4, WF = 8, WE = 8 + BR_BSTAGES, AF = 8 + 2 * BR_BSTAGES, AE = AF + 2;
  const uint32_t bars = sbase + offs[0];
  auto BAR = [&](int i) { return bars + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + offs[0] + 8 * (AE + 2));
  float* coefA = reinterpret_cast<float*>(smem + offs[1]);
  float* coefE = reinterpret_cast<float*>(smem + offs[2]);
  float* red = reinterpret_cast<float*>(smem + offs[3]);
  constexpr int PH = BR_PH;                       // planes per brick buffer
  const int NH = p.CH / (8 * PH);                // brick buffers per tile: 4 (fprop) or 1 (dgrad)
  const uint32_t brick0 = sbase + offs[4];
  constexpr uint32_t brick_bytes = (uint32_t)PH * BR_PLANE;
  const uint32_t bst0 = sbase + offs[5];
  const uint32_t b_bytes = (uint32_t)PH * p.NT * 16;              // one tap
  const uint32_t bs_bytes = BR_BTAPS * b_bytes;                  // one ring stage

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_y = (p.Dy + BR_TY - 1) / BR_TY, tiles_x = (p.Dx + BR_TX - 1) / BR_TX;
  const int ntiles = p.B * p.Dz * tiles_y * tiles_x;
  // the unit of work is a GROUP of TP consecutive tiles (the last group may hold an invalid tile: zero-filled, not stored)
  const int ngroups = (ntiles + TP - 1) / TP;
  const int my_groups = ((int)blockIdx.x < ngroups) ? (ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int rot = (int)(blockIdx.x % (27 / BR_BTAPS));   // per-CTA rotation of the tap-group order
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * TP * p.NT) tmem_cols <<= 1;   // TP accumulators per group, double-buffered

  if (warp == BR_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < NBUF; ++i) { mbar_init(BAR(BF + i), NPT); mbar_init(BAR(BE + i), 1); }
      for (int i = 0; i < BR_BSTAGES; ++i) { mbar_init(BAR(WF + i), 1); mbar_init(BAR(WE + i), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(AF + i), 1); mbar_init(BAR(AE + i), NET); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
  }
  pdl_wait();   // nothing above touches global memory
  pdl_trigger();   // AFTER the wait: a dependent that starts early may rely on everything before THIS kernel being complete
  if (TRANS == T_BNRELU) {
    for (int c = tid; c < p.CH; c += NTHREADS) {
      float mean, rstd;
      bn_mean_rstd(p.bnA, c, mean, rstd);
      const float s = p.bnA.gamma[c] * rstd;
      coefA[c] = s;
      coefA[p.CH + c] = p.bnA.beta[c] - mean * s;
    }
  }
  if (EPI == EP_MASK_STATS) {
    for (int c = tid; c < p.NT; c += NTHREADS) {
      float mean, rstd;
      bn_mean_rstd(p.bnE, c, mean, rstd);
      const float s = p.bnE.gamma[c] * rstd;
      coefE[c] = s; coefE[p.NT + c] = p.bnE.beta[c] - mean * s; coefE[2 * p.NT + c] = mean; coefE[3 * p.NT + c] = rstd;
    }
  }
  H2Coef* coefH = reinterpret_cast<H2Coef*>(coefA + 2 * p.CH);
  if (TRANS == T_BNRELU && OP_F16) {
    __syncthreads();
    fill_h2coef(coefH, coefA, coefA + p.CH, p.CH, tid, NTHREADS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile index -> coordinates; returns false for the padding tile of the last group
  auto tile_coords = [&](int t, int& n, int& z, int& y0, int& x0) {
    const bool valid = t < ntiles;
    if (!valid) t = 0;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y; t /= tiles_y;
    z = t % p.Dz; n = t / p.Dz;
    y0 = ty * BR_TY; x0 = tx * BR_TX;
    return valid;
  };
  // Brick buffers are filled and consumed in the order (group, channel quarter h, tile-of-group t):
  //   seq = ((group_iteration * NH) + h) * TP + t,   ring slot = seq % NBUF.

  if (warp < NPW) {
    // ================= producers: one load + one transform per brick cell.  Software-pipelined over the buffer ring:
    // the asynchronous copies of buffer seq (and seq-1) are in flight while buffer seq-LAG is transformed in place and
    // handed to the MMA warp, so the L2 / HBM round trip is not exposed.
    constexpr int cells = BR_SLOTS * PH;
    // A thread owns the same brick cells for every buffer: cell c = tid + u*NPT -> chunk = c % 4 (constant per thread),
    // slot = c / 4.  The slot's halo coordinates are decoded ONCE (packed z|y|x) instead of per tile.
    constexpr int MAXU = (cells + NPT - 1) / NPT;   // 9 (forward) / 17 (dgrad)
    const int chunk = tid & 3;
    int pk[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int c = tid + u * NPT;
      const int slot = c >> 2;
      const int xx = slot % BR_HX, r2 = slot / BR_HX;
      pk[u] = (c < cells) ? (((r2 / BR_HY) << 16) | ((r2 % BR_HY) << 8) | xx) : -1;
    }
    const int total = my_groups * NH * TP;
    uint32_t m_prev1 = 0, m_prev2 = 0;     // validity masks of the buffers issued 1 and 2 iterations ago
    int it = 0, h = 0, t = 0;               // (group, quarter, tile-of-group) of the buffer being ISSUED
    int hq1 = 0, hq2 = 0;                   // quarter index of the buffers issued 1 and 2 iterations ago
    for (int seq = 0; seq < total + LAG; ++seq) {
      uint32_t okmask = 0;
      const int h_cur = h;
      if (seq < total) {
        int n, z, y0, x0;
        const bool tvalid = tile_coords(((int)blockIdx.x + it * (int)gridDim.x) * TP + t, n, z, y0, x0);
        const int q = seq % NBUF;
        mbar_wait(BAR(BE + q), ((uint32_t)(seq / NBUF) & 1u) ^ 1u, 21);
        const uint32_t dst = brick0 + q * brick_bytes + chunk * BR_PLANE;
        const bf16* src = p.a_src + h * 32 + chunk * 8;
        const long long nbase = (long long)n * p.Dz;
#pragma unroll
        for (int u = 0; u < MAXU; ++u) {
          if (pk[u] >= 0) {
            const int sz = z + (pk[u] >> 16) - 1, sy = y0 + ((pk[u] >> 8) & 0xff) - 1, sx = x0 + (pk[u] & 0xff) - 1;
            const bool ok = tvalid && (unsigned)sz < (unsigned)p.Dz && (unsigned)sy < (unsigned)p.Dy && (unsigned)sx < (unsigned)p.Dx;
            const long long m = ok ? ((nbase + sz) * p.Dy + sy) * p.Dx + sx : 0;
            const int slot = (tid >> 2) + u * (NPT / 4);
            cp_async16(dst + slot * 16, src + m * p.a_pitch, ok ? 16u : 0u);   // zero-fill outside the volume
            okmask |= (uint32_t)ok << u;
          }
        }
        if (++t == TP) { t = 0; if (++h == NH) { h = 0; ++it; } }
      }
      cp_async_commit();                    // (an empty group past the end keeps the group count uniform)
      if (seq >= LAG) {
        const int j = seq - LAG;            // buffer to finish: its copies are complete once at most LAG groups are pending
        cp_async_wait<LAG>();
        const int qj = j % NBUF;
        if (TRANS == T_BNRELU) {
          const uint32_t okj = (LAG == 1) ? m_prev1 : m_prev2;
          const uint32_t dstj = brick0 + qj * brick_bytes + chunk * BR_PLANE;
          const int ch0 = ((LAG == 1) ? hq1 : hq2) * 32 + chunk * 8;
          if (OP_F16) {
            H2Coef hc[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) hc[i] = coefH[ch0 / 2 + i];
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
              if ((okj >> u) & 1u) {
                const int slot = (tid >> 2) + u * (NPT / 4);
                uint4 v = lds16(dstj + slot * 16);
                apply_bnrelu8_h2(v, hc);
                sts16(dstj + slot * 16, v);
              }
            }
          } else {
            float sc[8], sh[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { sc[e] = coefA[ch0 + e]; sh[e] = coefA[p.CH + ch0 + e]; }
#pragma unroll
            for (int u = 0; u < MAXU; ++u) {
              if ((okj >> u) & 1u) {
                const int slot = (tid >> 2) + u * (NPT / 4);
                uint4 v = lds16(dstj + slot * 16);
                apply_bnrelu8<OP_F16, OP_F16>(v, sc, sh);
                sts16(dstj + slot * 16, v);
              }
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(BAR(BF + qj));
      }
      m_prev2 = m_prev1; m_prev1 = okmask;
      hq2 = hq1; hq1 = h_cur;
    }
  } else if (warp == BR_LOAD_WARP) {
    // ================= weight loader: ring of BR_BSTAGES stages of BR_BTAPS tap images each (one pass per group)
    if (lane == 0) {
      int j = 0;
      for (int it = 0; it < my_groups; ++it)
        for (int h = 0; h < NH; ++h)
          for (int tg = 0; tg < 27 / BR_BTAPS; ++tg, ++j) {
            const int s = j % BR_BSTAGES;
            const uint32_t par = (uint32_t)(j / BR_BSTAGES) & 1u;
            mbar_wait(BAR(WE + s), par ^ 1u, 22);
            mbar_arrive_expect_tx(BAR(WF + s), bs_bytes);
            // every CTA walks the tap groups in a different rotation: otherwise all 148 SMs request the same weights
            // at the same moment and the few L2 slices holding those lines serialise them
            const int tgr = (tg + rot) % (27 / BR_BTAPS);
            for (int u = 0; u < BR_BTAPS; ++u)
              bulk_g2s(bst0 + s * bs_bytes + u * b_bytes,
                       p.b_packed + (size_t)((tgr * BR_BTAPS + u) * NH + h) * (size_t)(PH * p.NT * 8), b_bytes, BAR(WF + s));
          }
    }
  } else if (warp == BR_MMA_WARP) {
    // ================= MMA issuer
    const uint32_t idesc = make_idesc(TILE_ROWS, p.NT, 0, 0, OP_F16);
    int j = 0;
    for (int it = 0; it < my_groups; ++it) {
      const int abuf = it & 1;
      const uint32_t apar = (uint32_t)(it >> 1) & 1u;
      mbar_wait(BAR(AE + abuf), apar ^ 1u, 23);   // epilogue has drained this accumulator set
      tc_fence_after();
      for (int h = 0; h < NH; ++h) {
        const int seq0 = (it * NH + h) * TP;      // buffers seq0 .. seq0+TP-1 hold quarter h of the group's tiles
#pragma unroll
        for (int t = 0; t < TP; ++t) mbar_wait(BAR(BF + (seq0 + t) % NBUF), (uint32_t)((seq0 + t) / NBUF) & 1u, 24);
        tc_fence_after();
        for (int tg = 0; tg < 27 / BR_BTAPS; ++tg, ++j) {
          const int s = j % BR_BSTAGES;
          mbar_wait(BAR(WF + s), (uint32_t)(j / BR_BSTAGES) & 1u, 25);
          tc_fence_after();
          if (elect_one()) {
            // taps tgr*BR_BTAPS + v.  Window start slot = (d+1) per axis, d = (t-1)*tap_sign.
            const int tgr = (tg + rot) % (27 / BR_BTAPS);
#pragma unroll
            for (int t = 0; t < TP; ++t) {
              const uint64_t ad_base = make_smem_desc(brick0 + ((seq0 + t) % NBUF) * brick_bytes, BR_PLANE, BR_HX * 16);
              const uint32_t td = tmem_base + (abuf * TP + t) * p.NT;
              uint64_t bd = make_smem_desc(bst0 + s * bs_bytes, p.NT * 16, 128);
#pragma unroll
              for (int v = 0; v < BR_BTAPS; ++v) {
                const int t9 = BR_BTAPS == 9 ? tgr : (BR_BTAPS == 3 ? tgr / 3 : tgr / 9);
                const int t3 = BR_BTAPS == 9 ? v / 3 : (BR_BTAPS == 3 ? tgr - (tgr / 3) * 3 : (tgr / 3) % 3);
                const int t1 = BR_BTAPS == 9 ? v % 3 : (BR_BTAPS == 3 ? v : tgr % 3);
                const int oz = (t9 - 1) * p.tap_sign + 1, oy = (t3 - 1) * p.tap_sign + 1, ox = (t1 - 1) * p.tap_sign + 1;
                const uint64_t ad = desc_advance(ad_base, (uint32_t)((oz * BR_HY + oy) * BR_HX + ox) * 16u);
                tc_mma_bf16(td, ad, bd, idesc, (h > 0 || tg > 0 || v > 0) ? 1u : 0u);
                tc_mma_bf16(td, desc_advance(ad, 2 * BR_PLANE), desc_advance(bd, 2 * p.NT * 16), idesc, 1u);   // K 16..31
                bd = desc_advance(bd, b_bytes);
              }
            }
            tc_commit(BAR(WE + s));
          }
          __syncwarp();
        }
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < TP; ++t) tc_commit(BAR(BE + (seq0 + t) % NBUF));
          if (h == NH - 1) tc_commit(BAR(AF + abuf));
        }
        __syncwarp();
      }
    }
  } else if (warp >= BR_EPI_WARP0) {
    // ================= epilogue (TMEM lane quarter = warp % 4; with 8 warps, warp e and e+4 share a quarter and split
    // the 32-column chunks: CCW chunks per warp starting at cc0)
    constexpr int CCW = 16 / NEW;      // 32-column chunks per warp: 4 (forward, only one in use), 2 (dgrad)
    const int qd = warp & 3;
    const int etid = (warp - BR_EPI_WARP0) * 32 + lane;
    const int cc0 = ((warp - BR_EPI_WARP0) >> 2) * CCW;
    const int r = qd * 32 + lane;
    const int ry = r >> 3, rx = r & 7;
    float acc1[CCW], acc2[CCW];   // per-lane column partials
#pragma unroll
    for (int k = 0; k < CCW; ++k) { acc1[k] = 0.f; acc2[k] = 0.f; }
    for (int it = 0; it < my_groups; ++it) {
      const int abuf = it & 1;
#pragma unroll
      for (int t = 0; t < TP; ++t) {
        int n, z, y0, x0;
        const bool tvalid = tile_coords(((int)blockIdx.x + it * (int)gridDim.x) * TP + t, n, z, y0, x0);
        const bool row_ok = tvalid && (y0 + ry < p.Dy) && (x0 + rx < p.Dx);
        const long long m = (((long long)n * p.Dz + z) * p.Dy + (y0 + ry)) * p.Dx + (x0 + rx);
        uint4 xpre[CCW][4];   // gating activations of this row: fetched before the accumulator is ready
        if (EPI == EP_MASK_STATS) {
#pragma unroll
          for (int k = 0; k < CCW; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              xpre[k][i] = (row_ok && (cc0 + k) * 32 < p.NT) ? ldg16(p.e_src + m * p.e_pitch + (cc0 + k) * 32 + i * 8) : make_uint4(0, 0, 0, 0);
        }
        if (EPI == EP_MASK_STATS && TP == 1 && it + 1 < my_groups) {
          // the gating activations come from HBM (written a whole forward pass ago): pull the NEXT tile's rows into L2
          // now, so that its loads at the top of the next iteration are L2 hits instead of an exposed HBM round trip
          int n2, z2, y2, x2;
          tile_coords((int)blockIdx.x + (it + 1) * (int)gridDim.x, n2, z2, y2, x2);
          if ((y2 + ry < p.Dy) && (x2 + rx < p.Dx)) {
            const long long m2 = (((long long)n2 * p.Dz + z2) * p.Dy + (y2 + ry)) * p.Dx + (x2 + rx);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.e_src + m2 * p.e_pitch + cc0 * 32));
          }
        }
        if (t == 0) {
          mbar_wait(BAR(AF + abuf), (uint32_t)(it >> 1) & 1u, 26);
          tc_fence_after();
        }
#pragma unroll
        for (int k = 0; k < CCW; ++k) {
          const int cc = cc0 + k;
          if (cc * 32 >= p.NT) break;
          float v[32], qv[32];
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((abuf * TP + t) * p.NT + cc * 32), v);
          if (EPI == EP_MASK_STATS) {
            uint4 xv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = xpre[k][i];
            const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              float xlo, xhi;
              unpack2<E_F16>(xw[jj >> 1], xlo, xhi);
              const float x = (jj & 1) ? xhi : xlo;
              const int c = cc * 32 + jj;
              const bool act = fmaf(x, coefE[c], coefE[p.NT + c]) > 0.f;
              const float g = (row_ok && act) ? round16<OP_F16>(v[jj]) : 0.f;
              v[jj] = g;
              qv[jj] = g * (x - coefE[2 * p.NT + c]) * coefE[3 * p.NT + c];
            }
          } else {
            if (p.colscale != nullptr) {
              const float* cs = p.colscale + (size_t)n * p.NT + cc * 32;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] *= __ldg(cs + jj);
            }
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const float g = row_ok ? round16<OP_F16>(v[jj]) : 0.f;
              v[jj] = g;
              qv[jj] = g * g;
            }
          }
          if (row_ok) {
            uint4* op = reinterpret_cast<uint4*>(p.out + m * p.out_pitch + cc * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 o;
              o.x = pack2<OP_F16>(v[8 * i + 0], v[8 * i + 1]); o.y = pack2<OP_F16>(v[8 * i + 2], v[8 * i + 3]);
              o.z = pack2<OP_F16>(v[8 * i + 4], v[8 * i + 5]); o.w = pack2<OP_F16>(v[8 * i + 6], v[8 * i + 7]);
              op[i] = o;
            }
          }
          if (EPI != EP_STORE) {
            acc1[k] += warp_transpose_sum32(v, lane);
            acc2[k] += warp_transpose_sum32(qv, lane);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(BAR(AE + abuf));
    }
    if (EPI != EP_STORE) {
#pragma unroll
      for (int k = 0; k < CCW; ++k) {
        const int cc = cc0 + k;
        if (cc * 32 < p.NT) {
          red[(0 * 4 + qd) * p.NT + cc * 32 + lane] = acc1[k];
          red[(1 * 4 + qd) * p.NT + cc * 32 + lane] = acc2[k];
        }
      }
      named_bar_sync(1, NET);
      for (int c = etid; c < p.NT; c += NET) {
        const float a = red[(0 * 4 + 0) * p.NT + c] + red[(0 * 4 + 1) * p.NT + c] + red[(0 * 4 + 2) * p.NT + c] + red[(0 * 4 + 3) * p.NT + c];
        const float b = red[(1 * 4 + 0) * p.NT + c] + red[(1 * 4 + 1) * p.NT + c] + red[(1 * 4 + 2) * p.NT + c] + red[(1 * 4 + 3) * p.NT + c];
        atomicAdd(p.st_sum + c, (double)a);
        atomicAdd(p.st_sq + c, (double)b);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BR_MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mmnn
