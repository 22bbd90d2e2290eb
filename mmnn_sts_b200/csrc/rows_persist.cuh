// Persistent, warp-specialised form of the 1x1x1 GEMM kernels (forward and data gradient of conv1 / transitions in the
// large dense blocks).  conv_rows_kernel (engine.cuh) runs ONE 128-row tile per CTA: prologue (barrier init, TMEM
// allocation, BN coefficient tables), K loop and epilogue are serialised inside a CTA, and a block-1 layer needs 2048
// such CTAs -- measured 83 us for a layer whose HBM traffic is worth ~25 us.  Here a CTA per SM loops over tiles with
// dedicated roles, as the brick kernel does: warps 0-7 producers (LDG -> BN+ReLU as packed HFMA2 -> STS, loads of the
// next k-block issued before the current one is transformed; they simply continue into the next tile), warp 8 MMA
// issuer (TMEM accumulator double-buffered), warps 9-16 epilogue (two warps per TMEM lane quarter, each taking half of
// the column chunks; statistics partials stay in registers across the CTA's tiles and are reduced once).  The
// epilogue of tile t overlaps the loads and MMAs of tile t+1.
// All tiles of a CTA share one N tile (so the per-column partials and coefficient tables are per CTA).
#pragma once
#include "engine.cuh"

namespace mmnn {

constexpr int RP_NPW = 8, RP_NPT = RP_NPW * 32, RP_MMA_WARP = RP_NPW, RP_EPI_WARP0 = RP_NPW + 1, RP_NEW = 8, RP_NET = RP_NEW * 32;
constexpr int RP_THREADS = (RP_EPI_WARP0 + RP_NEW) * 32;   // 544

// offs[5]: two staging buffers of the rounded output tile (16 chunk planes each) + 512 B of ones for the tensor-core
// column statistics (NT == 128 launches with statistics; size 0 otherwise)
__host__ __device__ inline uint32_t rowsp_smem_layout(int Cin, int NT, int kbw, int stages, bool tcs, uint32_t* offs /*[6]*/, bool tma = false) {
  uint32_t o = 0;
  offs[0] = o; o += 256;                 // barriers: full[6] empty[6] acc_full[2] acc_empty[2] staged[2] gfree[2] stats_final + tmem ptr, raw[6] at +192
  offs[1] = o; o += 2u * Cin * 4;        // coefA: fp32 scale / shift or packed half2 table (same size)
  offs[2] = o; o += 4u * NT * 4;         // coefE: scale, shift, mean, rstd
  offs[3] = o; o += 8u * NT * 4;         // red[2][4][NT]
  o = (o + 127u) & ~127u;
  offs[4] = o;
  const uint32_t stage = rows_stage_bytes(NT, kbw, tma);      // TMA mode (RowsParams::tma_a): 1 KB multiples behind a 1 KB aligned base
  o += stages * stage + (tma ? 1024u : 0u);
  o = (o + 127u) & ~127u;
  offs[5] = o;
  if (tcs) o += 2u * 16u * PLANE_BYTES + 512u;
  return o;
}

template <int TRANS, int EPI, bool GRAD>
__global__ void __launch_bounds__(RP_THREADS, 1) conv1_persist_kernel(const __grid_constant__ RowsParams p,
                                                                      const __grid_constant__ CUtensorMap tma) {
  constexpr bool OP_F16 = !GRAD && kActF16;   // MMA operand + output format of this launch
  constexpr bool E_F16 = kActF16;
  extern __shared__ __align__(128) uint8_t smem[];
  // Column statistics on the tensor core (see conv_rows_kernel): here the staged tile of tile t is consumed by the MMA
  // warp AFTER it has issued the main MMAs of tile t+1, and the statistics accumulate in TMEM across all tiles of the
  // CTA -- no transposes and no per-tile read-out for sum / sum of squares (forward) and sum (data gradient).
  const bool tcs = (EPI != EP_STORE) && p.NT == 128 && (p.stages & 0x100) != 0;
  const int S = p.stages & 0xff;
  uint32_t offs[6];
  // raw activation k-blocks by TMA (SWIZZLE_128B K-major rows), BN+ReLU applied in place: see conv_rows_kernel / RowsParams::tma_a
  const bool tma_a = TRANS == T_BNRELU && OP_F16 && !GRAD && p.kbw == 64 && p.tma_a != 0;
  rowsp_smem_layout(p.Cin, p.NT, p.kbw, S, tcs, offs, tma_a);
  const uint32_t sbase = smem_u32(smem);
  constexpr int FULL = 0, EMPTY = 6, AF = 12, AE = 14, STAGED = 16, GFREE = 18, SFINAL = 20, RAW = 24;
  const uint32_t sG0 = sbase + offs[5], sOnes = sG0 + 2u * 16u * PLANE_BYTES;
  constexpr uint32_t D1_COL = 256, D2_COL = 288;   // statistics accumulators in TMEM (main accumulators: 0..255)
  auto BAR = [&](int i) { return sbase + offs[0] + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + offs[0] + 8 * 22);
  float* coefA = reinterpret_cast<float*>(smem + offs[1]);
  H2Coef* coefH = reinterpret_cast<H2Coef*>(coefA);
  float* coefE = reinterpret_cast<float*>(smem + offs[2]);
  float* red = reinterpret_cast<float*>(smem + offs[3]);
  const int planes = p.kbw / 8;
  const uint32_t a_bytes = planes * PLANE_BYTES;
  const uint32_t b_bytes = planes * p.NT * 16;
  const uint32_t stage_bytes = rows_stage_bytes(p.NT, p.kbw, tma_a);
  const uint32_t stage0 = tma_a ? ((sbase + offs[4] + 1023u) & ~1023u) : sbase + offs[4];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = (p.Cin + p.kbw - 1) / p.kbw;
  const int vps = p.Dz * p.Dy * p.Dx;
  const int tiles_m = (p.M + TILE_ROWS - 1) / TILE_ROWS;
  const int ntn = (p.Ncols + p.NT - 1) / p.NT;
  // CTA c works on N tile (c % ntn) and the M tiles (c / ntn), (c / ntn) + gridDim/ntn, ...   (gridDim % ntn == 0)
  const int tile_n = (int)blockIdx.x % ntn;
  const int m_first = (int)blockIdx.x / ntn, m_step = (int)gridDim.x / ntn;
  const int my_tiles = m_first < tiles_m ? (tiles_m - 1 - m_first) / m_step + 1 : 0;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (tcs ? 512 : 2 * p.NT)) tmem_cols <<= 1;

  if (warp == RP_MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(BAR(FULL + s), RP_NPT + 1); mbar_init(BAR(EMPTY + s), 1); mbar_init(BAR(RAW + s), 1); }
      if (tma_a) tma_prefetch_desc(&tma);
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(AF + i), 1); mbar_init(BAR(AE + i), RP_NET); }
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(STAGED + i), RP_NET); mbar_init(BAR(GFREE + i), 1); }
      mbar_init(BAR(SFINAL), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
  }
  pdl_wait();   // nothing above touches global memory
  pdl_trigger();   // AFTER the wait: a dependent that starts early may rely on everything before THIS kernel being complete
  if (TRANS == T_BNRELU) {
    if (OP_F16) {
      for (int j = tid; j < p.Cin / 2; j += RP_THREADS) {
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
      for (int c = tid; c < p.Cin; c += RP_THREADS) {
        float mean, rstd;
        bn_mean_rstd(p.bnA, c, mean, rstd);
        const float s = p.bnA.gamma[c] * rstd;
        coefA[c] = s;
        coefA[p.Cin + c] = p.bnA.beta[c] - mean * s;
      }
    }
  }
  if (EPI == EP_MASK_STATS) {
    for (int c = tid; c < p.NT; c += RP_THREADS) {
      const int col = tile_n * p.NT + c;
      float mean = 0.f, rstd = 0.f, s = 0.f, t = -1.f;
      if (col < p.Ncols) {
        bn_mean_rstd(p.bnE, col, mean, rstd);
        s = p.bnE.gamma[col] * rstd;
        t = p.bnE.beta[col] - mean * s;
      }
      coefE[c] = s; coefE[p.NT + c] = t; coefE[2 * p.NT + c] = mean; coefE[3 * p.NT + c] = rstd;
    }
  }
  if (tcs && warp == 0) {   // 512 B of 1.0 in the operand format
    const uint32_t one = OP_F16 ? 0x3c003c00u : 0x3f803f80u;
    sts16(sOnes + lane * 16, make_uint4(one, one, one, one));
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < RP_NPW) {
    // ================= producers (see conv_rows_kernel): one warp-wide 128-bit load covers rpp rows x cpl chunks
    auto geom_cpl = [&](int cb) { int cpl = (p.Cin - cb * p.kbw) / 8; return cpl > planes ? planes : cpl; };
    auto load_kb = [&](int tile_m, int cb, uint4 (&regs)[MAX_PASSES], uint32_t& okmask) {
      const int cpl = geom_cpl(cb);
      const int cshift = (cpl == 8) ? 3 : 2;
      const int chunk = lane & (cpl == 8 ? 7 : 3);
      const int rsub = lane >> cshift;
      const int rpp = 32 >> cshift;
      const int npass = (TILE_ROWS / rpp) / RP_NPW;
      const int ch0 = cb * p.kbw + chunk * 8;
      okmask = 0;
#pragma unroll
      for (int ps = 0; ps < MAX_PASSES; ++ps) {
        regs[ps] = make_uint4(0, 0, 0, 0);
        if (ps < npass) {
          const long long m = (long long)tile_m * TILE_ROWS + (warp + ps * RP_NPW) * rpp + rsub;
          if (m < p.M) {
            regs[ps] = ldg16(p.a_src + m * p.a_pitch + ch0);
            okmask |= 1u << ps;
          }
        }
      }
    };
    const long long total = (long long)my_tiles * KB;       // k-blocks of this CTA, in order
    uint4 R0[MAX_PASSES], R1[MAX_PASSES];
    uint32_t ok0 = 0, ok1 = 0;
    int it_n = 0, cb_n = 0;                                   // (tile iteration, k-block) of the NEXT load
    auto next_load = [&](uint4 (&regs)[MAX_PASSES], uint32_t& okm) {
      load_kb(m_first + it_n * m_step, cb_n, regs, okm);
      if (++cb_n == KB) { cb_n = 0; ++it_n; }
    };
    auto process = [&](long long g, int cb, const uint4 (&regs)[MAX_PASSES], uint32_t okm) {
      const int s = (int)(g % S);
      mbar_wait(BAR(EMPTY + s), ((uint32_t)(g / S) & 1u) ^ 1u, 61);
      const uint32_t sA = stage0 + s * stage_bytes;
      if (tid == 0) {
        mbar_arrive_expect_tx(BAR(FULL + s), b_bytes);
        bulk_g2s(sA + a_bytes, p.b_packed + ((size_t)tile_n * KB + cb) * (size_t)(planes * p.NT * 8), b_bytes, BAR(FULL + s));
      }
      const int cpl = geom_cpl(cb);
      const int cshift = (cpl == 8) ? 3 : 2;
      const int chunk = lane & (cpl == 8 ? 7 : 3);
      const int rsub = lane >> cshift;
      const int rpp = 32 >> cshift;
      const int npass = (TILE_ROWS / rpp) / RP_NPW;
      const int ch0 = cb * p.kbw + chunk * 8;
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
#pragma unroll
      for (int ps = 0; ps < MAX_PASSES; ++ps) {
        if (ps < npass) {
          const int r = (warp + ps * RP_NPW) * rpp + rsub;
          uint4 v = regs[ps];
          if (TRANS == T_BNRELU && ((okm >> ps) & 1u)) {
            if (OP_F16) apply_bnrelu8_h2(v, hc);
            else apply_bnrelu8<OP_F16, OP_F16>(v, sc, sh);
          }
          sts16(sA + chunk * PLANE_BYTES + r * 16, v);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(BAR(FULL + s));
    };
    if (tma_a) {
      // thread 0 feeds the ring LA = S - 2 k-blocks ahead, across tile boundaries; all threads transform their 4 cells in place
      const int LA = S > 2 ? S - 2 : 1;
      auto tma_issue = [&](long long j) {
        const int sj = (int)(j % S);
        mbar_wait(BAR(EMPTY + sj), ((uint32_t)(j / S) & 1u) ^ 1u, 61);
        const int itj = (int)(j / KB), cbj = (int)(j - (long long)itj * KB);
        const uint32_t sAj = stage0 + sj * stage_bytes;
        mbar_arrive_expect_tx(BAR(RAW + sj), (uint32_t)TILE_ROWS * 128u);
        tma_load_2d(sAj, &tma, cbj * 64, (m_first + itj * m_step) * TILE_ROWS, BAR(RAW + sj));
        mbar_arrive_expect_tx(BAR(FULL + sj), b_bytes);
        bulk_g2s(sAj + (uint32_t)TILE_ROWS * 128u, p.b_packed + ((size_t)tile_n * KB + cbj) * (size_t)(planes * p.NT * 8), b_bytes, BAR(FULL + sj));
      };
      if (tid == 0)
        for (long long j = 0; j < total && j < LA; ++j) tma_issue(j);
      const int c = tid & 7, rb = tid >> 3;
      int cbq = 0, itq = 0;
      for (long long g = 0; g < total; ++g) {
        if (tid == 0 && g + LA < total) tma_issue(g + LA);
        const int s2 = (int)(g % S);
        mbar_wait(BAR(RAW + s2), (uint32_t)(g / S) & 1u, 68);
        const uint32_t sA = stage0 + s2 * stage_bytes;
        const int cpl = geom_cpl(cbq);
        if (c < cpl) {
          H2Coef hc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) hc[i] = coefH[(cbq * 64 + c * 8) / 2 + i];
          const long long row0 = (long long)(m_first + itq * m_step) * TILE_ROWS;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = rb + 32 * j;
            if (row0 + r < p.M) {
              const uint32_t addr = sA + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
              uint4 v = lds16(addr);
              apply_bnrelu8_h2(v, hc);
              sts16(addr, v);
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(BAR(FULL + s2));
        if (++cbq == KB) { cbq = 0; ++itq; }
      }
    } else {
    if (total > 0) next_load(R0, ok0);
    int cb = 0;
    for (long long g = 0; g < total; g += 2) {
      if (g + 1 < total) next_load(R1, ok1);
      process(g, cb, R0, ok0);
      if (++cb == KB) cb = 0;
      if (g + 1 < total) {
        if (g + 2 < total) next_load(R0, ok0);
        process(g + 1, cb, R1, ok1);
        if (++cb == KB) cb = 0;
      }
    }
    }
  } else if (warp == RP_MMA_WARP) {
    // ================= MMA issuer
    const uint32_t idesc = make_idesc(TILE_ROWS, p.NT, 0, 0, OP_F16);
    auto issue_stats = [&](int t) {   // statistics MMAs over the staged tile of tile iteration t
      const int b = t & 1;
      mbar_wait(BAR(STAGED + b), (uint32_t)(t >> 1) & 1u, 65);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t gd = make_smem_desc(sG0 + (uint32_t)b * 16u * PLANE_BYTES, 128, PLANE_BYTES);   // MN-major
        const uint64_t od = make_smem_desc(sOnes, 128, 128);
        const uint32_t id_sum = make_idesc(TILE_ROWS, 16, 1, 1, OP_F16);
#pragma unroll
        for (int k16 = 0; k16 < TILE_ROWS / 16; ++k16)
          tc_mma_bf16(tmem_base + D1_COL, desc_advance(gd, k16 * 256), od, id_sum, (t > 0 || k16 > 0) ? 1u : 0u);
        if (EPI == EP_STORE_STATS) {
          const uint32_t id_sq = make_idesc(TILE_ROWS, 128, 1, 1, OP_F16);
#pragma unroll
          for (int k16 = 0; k16 < TILE_ROWS / 16; ++k16)
            tc_mma_bf16(tmem_base + D2_COL, desc_advance(gd, k16 * 256), desc_advance(gd, k16 * 256), id_sq, (t > 0 || k16 > 0) ? 1u : 0u);
        }
        tc_commit(BAR(GFREE + b));
      }
      __syncwarp();
    };
    long long g = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int abuf = it & 1;
      mbar_wait(BAR(AE + abuf), ((uint32_t)(it >> 1) & 1u) ^ 1u, 62);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t td = tmem_base + abuf * p.NT;
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int s = (int)(g % S);
        mbar_wait(BAR(FULL + s), (uint32_t)(g / S) & 1u, 63);
        tc_fence_after();
        if (elect_one()) {
          int cpl = (p.Cin - kb * p.kbw) / 8;
          cpl = cpl > planes ? planes : cpl;
          const uint32_t sA = stage0 + s * stage_bytes;
          const uint64_t ad0 = tma_a ? make_smem_desc_sw(sA, 16, 1024, 2u) : make_smem_desc(sA, PLANE_BYTES, 128);
          const uint64_t bd0 = make_smem_desc(sA + (tma_a ? (uint32_t)TILE_ROWS * 128u : a_bytes), p.NT * 16, 128);
          const uint32_t a_step = tma_a ? 32u : 2 * PLANE_BYTES, b_step = 2 * p.NT * 16;
          tc_mma_bf16(td, ad0, bd0, idesc, kb > 0 ? 1u : 0u);
#pragma unroll
          for (int k16 = 1; k16 < 4; ++k16)
            if (k16 < cpl / 2) tc_mma_bf16(td, desc_advance(ad0, k16 * a_step), desc_advance(bd0, k16 * b_step), idesc, 1u);
          tc_commit(BAR(EMPTY + s));
          if (kb == KB - 1) tc_commit(BAR(AF + abuf));
        }
        __syncwarp();
      }
      if (tcs && it > 0) issue_stats(it - 1);   // its epilogue ran while this tile's MMAs were being fed
    }
    if (tcs && my_tiles > 0) {
      issue_stats(my_tiles - 1);
      if (elect_one()) tc_commit(BAR(SFINAL));
      __syncwarp();
    }
  } else {
    // ================= epilogue: warps e and e+4 share TMEM lane quarter (warp & 3); warp handles chunks cc0, cc0+1 (+4, +5 for NT = 256)
    const int e = warp - RP_EPI_WARP0;
    const int qd = warp & 3;
    const int half = e >> 2;               // 0 or 1
    const int etid = e * 32 + lane;
    const int r = qd * 32 + lane;
    const int nchunks = p.NT / 32;
    float acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};   // per-lane column partials of chunks half, half+2, half+4, half+6
    for (int it = 0; it < my_tiles; ++it) {
      const int tile_m = m_first + it * m_step;
      const int abuf = it & 1;
      const long long m = (long long)tile_m * TILE_ROWS + r;
      const bool row_ok = m < p.M;
      const int nb = row_ok ? (int)(m / vps) : 0;
      uint4 xpre[2][4];   // gating activations (first two chunks of this warp), fetched before the accumulator is ready
      if (EPI == EP_MASK_STATS) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int cc = half + 2 * k;
          const int col0 = tile_n * p.NT + cc * 32;
          const bool on = row_ok && cc < nchunks && col0 < p.Ncols;
#pragma unroll
          for (int i = 0; i < 4; ++i) xpre[k][i] = on ? ldg16(p.e_src + m * p.e_pitch + col0 + i * 8) : make_uint4(0, 0, 0, 0);
        }
        if (it + 1 < my_tiles) {   // pull the next tile's gating rows into L2
          const long long m2 = (long long)(tile_m + m_step) * TILE_ROWS + r;
          const int col0 = tile_n * p.NT + half * 32;
          if (m2 < p.M && col0 < p.Ncols) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.e_src + m2 * p.e_pitch + col0));
        }
      }
      mbar_wait(BAR(AF + abuf), (uint32_t)(it >> 1) & 1u, 64);
      tc_fence_after();
      const uint32_t sG = sG0 + (uint32_t)(it & 1) * 16u * PLANE_BYTES;
      if (tcs) mbar_wait(BAR(GFREE + (it & 1)), ((uint32_t)(it >> 1) & 1u) ^ 1u, 66);   // statistics MMAs of tile it-2 have read this buffer
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cc = half + 2 * k;
        if (cc >= nchunks) break;
        const int col0 = tile_n * p.NT + cc * 32;
        if (col0 >= p.Ncols) continue;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(abuf * p.NT + cc * 32), v);
        float q[32];
        if (EPI == EP_MASK_STATS) {
          uint4 xv[4];
          if (k < 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = xpre[k < 2 ? k : 1][i];
          } else {
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
            const float gq = (row_ok && act) ? round16<OP_F16>(v[j]) : 0.f;
            v[j] = gq;
            q[j] = gq * (x - coefE[2 * p.NT + c]) * coefE[3 * p.NT + c];
          }
        } else {
          if (p.colscale != nullptr) {
            const float* cs = p.colscale + (size_t)nb * p.Ncols + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= __ldg(cs + j);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float gq = row_ok ? round16<OP_F16>(v[j]) : 0.f;
            v[j] = gq;
            q[j] = gq * gq;
          }
        }
        {
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
          if (!tcs) acc1[k] += warp_transpose_sum32(v, lane);
          if (!tcs || EPI == EP_MASK_STATS) acc2[k] += warp_transpose_sum32(q, lane);
        }
      }
      tc_fence_before();
      mbar_arrive(BAR(AE + abuf));
      if (tcs) {
        fence_proxy_async_smem();
        mbar_arrive(BAR(STAGED + (it & 1)));
      }
    }
    if (tcs && my_tiles > 0) {
      // read the TMEM statistics accumulators once per CTA: lane quarter qd, lane = output column qd*32 + lane
      mbar_wait(BAR(SFINAL), 0, 67);
      tc_fence_after();
      if (e < 4) {
        float d[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + D1_COL, d);
        const int c = qd * 32 + lane;
        const int col = tile_n * p.NT + c;
        if (col < p.Ncols) atomicAdd(p.st_sum + col, (double)d[0]);
        if (EPI == EP_STORE_STATS) {
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + D2_COL + (uint32_t)(qd * 32), d);   // block holding the diagonal
          float x = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) x = (lane == j) ? d[j] : x;
          if (col < p.Ncols) atomicAdd(p.st_sq + col, (double)x);
        }
      }
    }
    const bool reg1 = EPI != EP_STORE && !tcs, reg2 = EPI != EP_STORE && (!tcs || EPI == EP_MASK_STATS);   // statistics still in registers
    if (reg1 || reg2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cc = half + 2 * k;
        if (cc < nchunks) {
          red[(0 * 4 + qd) * p.NT + cc * 32 + lane] = acc1[k];
          red[(1 * 4 + qd) * p.NT + cc * 32 + lane] = acc2[k];
        }
      }
      named_bar_sync(1, RP_NET);
      for (int c = etid; c < p.NT; c += RP_NET) {
        const int col = tile_n * p.NT + c;
        if (col < p.Ncols) {
          const float a = red[(0 * 4 + 0) * p.NT + c] + red[(0 * 4 + 1) * p.NT + c] + red[(0 * 4 + 2) * p.NT + c] + red[(0 * 4 + 3) * p.NT + c];
          const float b = red[(1 * 4 + 0) * p.NT + c] + red[(1 * 4 + 1) * p.NT + c] + red[(1 * 4 + 2) * p.NT + c] + red[(1 * 4 + 3) * p.NT + c];
          if (reg1) atomicAdd(p.st_sum + col, (double)a);
          if (reg2) atomicAdd(p.st_sq + col, (double)b);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RP_MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace mmnn
