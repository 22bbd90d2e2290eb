// Brick-mode stem convolution: Conv3d(C_in -> 64, k 7, s 2, p 3) of /root/reference/models/densenet.py:199 on the padded
// space-to-depth image xs2d [B][D0+3][H0+3][W0+3][16] (the 7^3 / s2 conv is a 4^3 / s1 conv over 16 s2d channels).
//
// The gather engine (conv_rows_kernel<A_STEM>) rebuilds a 128 x 64 A tile per (dz, dy) tap pair with LDG -> STS and is
// instruction-issue bound on its producers (ncu r01f: 52 % issue slots, 968 us at configs[1]).  Here, as in
// brick.cuh, a persistent CTA stages the HALO BRICK of one 1 x 16 x 8 output tile (4 x 19 x 11 voxel slots x 16
// channels = 26 KB) ONCE with cp.async -- no transform, the image is already padded -- and each of the 64 taps
// (dz, dy, dx) is one tcgen05.mma (M 128, N 64, K 16) whose A descriptor merely starts at a different slot:
//      start = plane0 + ((dz*19 + dy)*11 + dx)*16,   SBO (next 8 rows = next y) = 11*16 B,   LBO = plane stride.
// All 64 weight taps (128 KB packed image, PACK_STEM layout) stay resident in shared memory for the CTA's lifetime.
// Warp roles (416 threads): 0-3 producers (3-deep brick ring), 4 MMA issuer + TMEM owner + weight load, 5-12 epilogue
// (two warps per TMEM lane quarter, one 32-column chunk each; TMEM double-buffered).  The per-channel statistics of
// the stored values (for norm0) are accumulated in REGISTERS across the CTA's tiles and reduced across lanes once.
#pragma once
#include "engine.cuh"

namespace mmnn {

constexpr int SB_TY = 16, SB_TX = 8, SB_HZ = 4, SB_HY = SB_TY + 3, SB_HX = SB_TX + 3;
constexpr int SB_SLOTS = SB_HZ * SB_HY * SB_HX;          // 836
#ifndef MMNN_STEM_SW32
#define MMNN_STEM_SW32 1
#endif
// Brick layout.  SW32 (default): one 32-byte row per slot (the 16 channels of a voxel = exactly the K = 16 of one MMA),
// written with the 32-byte swizzle (16-byte half h of the slot at byte address a goes to h ^ bit7(a)); the A descriptor
// uses SWIZZLE_32B, 8-row groups = 8 consecutive x slots (256 B atoms), SBO = 11 slots.  Fallback (0): two chunk planes,
// SWIZZLE_NONE core matrices -- the operand fetch of that layout runs at ~64 B/clk.
constexpr bool SB_SW32 = MMNN_STEM_SW32 != 0;
constexpr int SB_PLANE = SB_SLOTS * 16 + 16;             // 13392 B: odd multiple of 16 (plane layout)
constexpr int SB_BRICK = SB_SW32 ? ((SB_SLOTS * 32 + 255) / 256) * 256 : 2 * SB_PLANE;   // 26880 / 26784 B
constexpr int SB_STAGES = 3;
constexpr int SB_NPW = 4, SB_NPT = SB_NPW * 32, SB_MMA_WARP = SB_NPW, SB_EPI_WARP0 = SB_NPW + 1, SB_NEW = 8, SB_NET = SB_NEW * 32;
constexpr int SB_THREADS = (SB_EPI_WARP0 + SB_NEW) * 32;  // 416
constexpr int SB_WBYTES = 64 * 2048;                     // 64 taps x [2 planes][64 n][8] 16-bit
constexpr uint32_t SB_OFF_RED = 256, SB_OFF_W = 256 + 2 * 4 * 64 * 4, SB_OFF_BRICK = SB_OFF_W + SB_WBYTES;   // 133376 = 521 * 256
constexpr uint32_t SB_SMEM = SB_OFF_BRICK + SB_STAGES * SB_BRICK + 1024;   // + slack to align the dynamic base to 1 KB

struct StemBrickParams {
  int B, D0, H0, W0;       // output dims
  int Sz, Sy, Sx;          // xs2d dims (D0+3, H0+3, W0+3)
  const bf16* xs2d;
  const bf16* w_packed;    // SW32: PACK_STEM_SW32 image [tap64][64 n][32 B swizzled]; else PACK_STEM [tap16][chunk8][64 n][8]
  bf16* out;               // [B*D0*H0*W0][out_pitch]
  long long out_pitch;
  double* st_sum;          // [64] or nullptr
  double* st_sq;
};

#ifdef MMNN_STEM_TIMING
__device__ long long g_stem_dbg[8];
#endif
static __global__ void __launch_bounds__(SB_THREADS, 1) stem_brick_kernel(const __grid_constant__ StemBrickParams p) {
  constexpr bool F16 = kActF16;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // swizzle patterns are functions of the shared-memory byte address: work from a 1 KB aligned base
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  // barriers (8 B each): brick_full[3] 0..2 | brick_empty[3] 3..5 | w_full 6 | acc_full[2] 7,8 | acc_empty[2] 9,10
  auto BAR = [&](int i) { return sbase + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 128);
  float* red = reinterpret_cast<float*>(smem + SB_OFF_RED);
  const uint32_t wsm = sbase + SB_OFF_W, brick0 = sbase + SB_OFF_BRICK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_y = (p.H0 + SB_TY - 1) / SB_TY, tiles_x = (p.W0 + SB_TX - 1) / SB_TX;
  const int ntiles = p.B * p.D0 * tiles_y * tiles_x;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // (rotating over NACC > 1 accumulators per tile, summed by the epilogue, was measured: no gain -- back-to-back MMAs
  // into one accumulator already run at the operand-fetch rate, profiles/scripts/microbench_mma.py)
  constexpr int NACC = 1;
  constexpr uint32_t TMEM_COLS = 2 * NACC * 64;

  if (warp == SB_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < SB_STAGES; ++i) { mbar_init(BAR(i), SB_NPT); mbar_init(BAR(3 + i), 1); }
      mbar_init(BAR(6), 1);
      for (int i = 0; i < 2; ++i) { mbar_init(BAR(7 + i), 1); mbar_init(BAR(9 + i), SB_NET); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
  }
  pdl_wait();   // nothing above touches global memory
  pdl_trigger();   // AFTER the wait: a dependent that starts early may rely on everything before THIS kernel being complete
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == SB_MMA_WARP && lane == 0) {
    mbar_arrive_expect_tx(BAR(6), SB_WBYTES);
    for (int i = 0; i < 16; ++i) bulk_g2s(wsm + i * 8192, p.w_packed + (size_t)i * 4096, 8192, BAR(6));
  }

  auto tile_coords = [&](int t, int& n, int& z, int& y0, int& x0) {
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y; t /= tiles_y;
    z = t % p.D0; n = t / p.D0;
    y0 = ty * SB_TY; x0 = tx * SB_TX;
  };

  if (warp < SB_NPW) {
    // ================= producers: cell c = tid + u*128 -> plane = c & 1, slot = c >> 1 (decoded once)
    constexpr int CELLS = SB_SLOTS * 2;
    constexpr int MAXU = (CELLS + SB_NPT - 1) / SB_NPT;   // 14
    const int plane = tid & 1;
    int pk[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int c = tid + u * SB_NPT;
      const int slot = c >> 1;
      const int sx = slot % SB_HX, r2 = slot / SB_HX;
      pk[u] = (c < CELLS) ? (((r2 / SB_HY) << 16) | ((r2 % SB_HY) << 8) | sx) : -1;
    }
    const bf16* src0 = p.xs2d + plane * 8;
    for (int it = 0; it < my_tiles; ++it) {
      int n, z, y0, x0;
      tile_coords((int)blockIdx.x + it * (int)gridDim.x, n, z, y0, x0);
      const int s = it % SB_STAGES;
      mbar_wait(BAR(3 + s), ((uint32_t)(it / SB_STAGES) & 1u) ^ 1u, 41);
      const uint32_t dst = brick0 + s * SB_BRICK + (SB_SW32 ? 0 : plane * SB_PLANE);
      const long long nbase = (long long)n * p.Sz + z;
#pragma unroll
      for (int u = 0; u < MAXU; ++u) {
        if (pk[u] >= 0) {
          const int sz = pk[u] >> 16, sy = y0 + ((pk[u] >> 8) & 0xff), sx = x0 + (pk[u] & 0xff);
          const bool ok = sy < p.Sy && sx < p.Sx;
          const long long cell = ok ? ((nbase + sz) * p.Sy + sy) * p.Sx + sx : 0;
          const int slot = (tid >> 1) + u * (SB_NPT / 2);
          const uint32_t off = SB_SW32 ? (uint32_t)slot * 32u + (uint32_t)((plane ^ ((slot >> 2) & 1)) << 4) : (uint32_t)slot * 16u;
          cp_async16(dst + off, src0 + cell * 16, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      if (it > 0) {   // the previous tile's copies have landed once at most this tile's group is pending
        cp_async_wait<1>();
        fence_proxy_async_smem();
        mbar_arrive(BAR((it - 1) % SB_STAGES));
      }
    }
    if (my_tiles > 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(BAR((my_tiles - 1) % SB_STAGES));
    }
  } else if (warp == SB_MMA_WARP) {
    // ================= MMA issuer: 64 taps per tile, weights resident
    const uint32_t idesc = make_idesc(TILE_ROWS, 64, 0, 0, F16);
    const uint64_t bd_base = SB_SW32 ? make_smem_desc_sw(wsm, 16, 256, 6) : make_smem_desc(wsm, 1024, 128);
    if (my_tiles > 0) mbar_wait(BAR(6), 0u, 42);
#ifdef MMNN_STEM_TIMING
    long long w_acc = 0, w_brick = 0, t_all0 = clock64();
#endif
    for (int it = 0; it < my_tiles; ++it) {
      const int abuf = it & 1, s = it % SB_STAGES;
#ifdef MMNN_STEM_TIMING
      long long c0 = clock64();
#endif
      mbar_wait(BAR(9 + abuf), ((uint32_t)(it >> 1) & 1u) ^ 1u, 43);   // epilogue has drained this accumulator
#ifdef MMNN_STEM_TIMING
      long long c1 = clock64();
#endif
      mbar_wait(BAR(s), (uint32_t)(it / SB_STAGES) & 1u, 44);
#ifdef MMNN_STEM_TIMING
      long long c2 = clock64(); w_acc += c1 - c0; w_brick += c2 - c1;
#endif
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad_base = SB_SW32 ? make_smem_desc_sw(brick0 + s * SB_BRICK, 16, SB_HX * 32, 6)
                                         : make_smem_desc(brick0 + s * SB_BRICK, SB_PLANE, SB_HX * 16);
        const uint32_t td = tmem_base + abuf * NACC * 64;
#pragma unroll
        for (int dz = 0; dz < 4; ++dz)
#pragma unroll
          for (int dy = 0; dy < 4; ++dy)
#pragma unroll
            for (int dx = 0; dx < 4; ++dx)
              tc_mma_bf16(td + (dx % NACC) * 64, desc_advance(ad_base, (uint32_t)((dz * SB_HY + dy) * SB_HX + dx) * (SB_SW32 ? 32u : 16u)),
                          desc_advance(bd_base, SB_SW32 ? (uint32_t)((dz * 4 + dy) * 4 + dx) * 2048u : (uint32_t)((dz * 4 + dy) * 8 + dx * 2) * 1024u),
                          idesc, ((dz | dy) != 0 || dx >= NACC) ? 1u : 0u);
        tc_commit(BAR(3 + s));
        tc_commit(BAR(7 + abuf));
      }
      __syncwarp();
    }
#ifdef MMNN_STEM_TIMING
    if (blockIdx.x == 0 && lane == 0) { g_stem_dbg[0] = w_acc; g_stem_dbg[1] = w_brick; g_stem_dbg[2] = clock64() - t_all0; g_stem_dbg[3] = my_tiles; }
#endif
  } else {
    // ================= epilogue: warp e and e+4 share TMEM lane quarter (warp & 3); chunk cc = e >> 2
    const int e = warp - SB_EPI_WARP0;
    const int qd = warp & 3, cc = e >> 2;
    const int etid = e * 32 + lane;
    const int r = qd * 32 + lane;
    const int ry = r >> 3, rx = r & 7;
    float a1[32], a2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
    for (int it = 0; it < my_tiles; ++it) {
      int n, z, y0, x0;
      tile_coords((int)blockIdx.x + it * (int)gridDim.x, n, z, y0, x0);
      const int abuf = it & 1;
      const bool row_ok = (y0 + ry < p.H0) && (x0 + rx < p.W0);
      const long long m = (((long long)n * p.D0 + z) * p.H0 + (y0 + ry)) * p.W0 + (x0 + rx);
      mbar_wait(BAR(7 + abuf), (uint32_t)(it >> 1) & 1u, 45);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(abuf * NACC * 64 + cc * 32), v);
#pragma unroll
      for (int a = 1; a < NACC; ++a) {
        float w[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((abuf * NACC + a) * 64 + cc * 32), w);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += w[j];
      }
      tc_fence_before();
      mbar_arrive(BAR(9 + abuf));            // values are in registers: the MMA warp may overwrite this accumulator
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float g = row_ok ? round16<F16>(v[j]) : 0.f;
        v[j] = g;
        a1[j] += g;
        a2[j] = fmaf(g, g, a2[j]);
      }
      if (row_ok) {
        uint4* op = reinterpret_cast<uint4*>(p.out + m * p.out_pitch + cc * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack2<F16>(v[8 * i + 0], v[8 * i + 1]); o.y = pack2<F16>(v[8 * i + 2], v[8 * i + 3]);
          o.z = pack2<F16>(v[8 * i + 4], v[8 * i + 5]); o.w = pack2<F16>(v[8 * i + 6], v[8 * i + 7]);
          op[i] = o;
        }
      }
    }
    if (p.st_sum != nullptr) {
      const float s1 = warp_transpose_sum32(a1, lane), s2 = warp_transpose_sum32(a2, lane);
      red[(0 * 4 + qd) * 64 + cc * 32 + lane] = s1;
      red[(1 * 4 + qd) * 64 + cc * 32 + lane] = s2;
      named_bar_sync(1, SB_NET);
      if (etid < 64) {
        const int c = etid;
        const float a = red[(0 * 4 + 0) * 64 + c] + red[(0 * 4 + 1) * 64 + c] + red[(0 * 4 + 2) * 64 + c] + red[(0 * 4 + 3) * 64 + c];
        const float b = red[(1 * 4 + 0) * 64 + c] + red[(1 * 4 + 1) * 64 + c] + red[(1 * 4 + 2) * 64 + c] + red[(1 * 4 + 3) * 64 + c];
        atomicAdd(p.st_sum + c, (double)a);
        atomicAdd(p.st_sq + c, (double)b);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SB_MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace mmnn
