// C-ABI entry points for the tcgen05 tile engine (fine-grained: used by the per-kernel parity tests and by the
// encoder orchestrator in encoder.cu).  Plain pointers and sizes only; see include/mmnn_b200.h.
#include "brick.cuh"
#include "engine.cuh"
#include "launch.h"
#include "pack.cuh"
#include "prof.h"
#include "rows_persist.cuh"
#include "stem.cuh"

#include <map>
#include <mutex>
#include <tuple>

using namespace mmnn;

namespace mmnn {
// 5-D TMA tensor map of a channels-last bf16 tensor [N][Dz][Dy][Dx][pitch] with the box (32 channels, bx, by, bz, bn), 64-byte swizzle:
// the descriptor the weight-gradient kernel's TMA issuer hands to cp.async.bulk.tensor (engine.cuh).  The encoder comes from the
// driver through cudaGetDriverEntryPoint (no link against libcuda); maps are cached per (pointer, geometry).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
bool make_tmap_ndhwc(CUtensorMap* out, const void* ptr, long long pitch, int N, int Dz, int Dy, int Dx, int bx, int by, int bz, int bn) {
  typedef std::tuple<const void*, long long, int, int, int, int, int, int, int, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  const Key key(ptr, pitch, N, Dz, Dy, Dx, bx, by, bz, bn);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr || ((uintptr_t)ptr & 15u) != 0 || (pitch * 2) % 16 != 0) return false;
  const cuuint64_t gdim[5] = {(cuuint64_t)pitch, (cuuint64_t)Dx, (cuuint64_t)Dy, (cuuint64_t)Dz, (cuuint64_t)N};
  const cuuint64_t gstr[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)Dx * pitch * 2, (cuuint64_t)Dy * Dx * pitch * 2,
                              (cuuint64_t)Dz * Dy * Dx * pitch * 2};
  const cuuint32_t box[5] = {32u, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bn};
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

// The bf16 space-to-depth image [N][Sz][Sy][Sx][16] seen as [N][Sz][Sy][Sx-3][64]: a "row" is 4 consecutive 16-channel cells and
// rows advance by one cell (32 B), so consecutive rows overlap -- the operand of the stem weight gradient (engine.cuh, tma_a).
bool make_tmap_s2d_rows(CUtensorMap* out, const void* ptr, int N, int Sz, int Sy, int Sx, int bx, int by, int bz, int bn) {
  typedef std::tuple<const void*, int, int, int, int, int, int, int, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  const Key key(ptr, N, Sz, Sy, Sx, bx, by, bz, bn);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr || ((uintptr_t)ptr & 15u) != 0 || Sx < 4) return false;
  const cuuint64_t gdim[5] = {64u, (cuuint64_t)(Sx - 3), (cuuint64_t)Sy, (cuuint64_t)Sz, (cuuint64_t)N};
  const cuuint64_t gstr[4] = {32u, (cuuint64_t)Sx * 32u, (cuuint64_t)Sy * Sx * 32u, (cuuint64_t)Sz * Sy * Sx * 32u};
  const cuuint32_t box[5] = {32u, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, (cuuint32_t)bn};
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 256) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

// [M][pitch] 2-byte elements as a 2-D tensor, box = 64 channels x 128 rows, SWIZZLE_128B: the canonical K-major tcgen05 operand of a
// 1x1x1 GEMM k-block (engine.cuh, RowsParams::tma_a)
bool make_tmap_rows_kmajor(CUtensorMap* out, const void* ptr, long long pitch, long long M) {
  typedef std::tuple<const void*, long long, long long> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  const Key key(ptr, pitch, M);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr || ((uintptr_t)ptr & 15u) != 0 || (pitch * 2) % 16 != 0 || pitch < 64) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)M};
  const cuuint64_t gstr[1] = {(cuuint64_t)pitch * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)TILE_ROWS};
  const cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 1024) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}
}  // namespace mmnn

#define MMNN_CHECK_LAUNCH()                      \
  do {                                           \
    cudaError_t e_ = cudaGetLastError();         \
    if (e_ != cudaSuccess) return (int)e_;       \
  } while (0)

namespace mmnn {

int choose_stages(const RowsParams& p, uint32_t budget) {
  uint32_t offs[6];
  const int kb_per_tap = (p.Cin + p.kbw - 1) / p.kbw;
  const int KB = p.ntaps * kb_per_tap;
  int best = 1;
  for (int s = 1; s <= 6 && s <= (KB > 1 ? KB : 1); ++s)
    if (rows_smem_layout(p.Cin, p.NT, p.kbw, s, offs) <= budget) best = s;
  return best;
}

template <int AMODE, int TRANS, int EPI, bool GRAD, int PF>
int launch_rows_pf(RowsParams p, cudaStream_t stream) {
  uint32_t offs[6];
  // grids that fill the GPU keep two CTAs per SM (100 KB each); grids of <= 148 tiles (late dense blocks) have an SM to themselves
  // and are latency chains over their k-blocks, so they take the deepest operand ring that fits (ncu, round 2: with Cin = 992
  // the 100 KB budget left TWO stages and every k-block exposed its weight fetch: 4 200 cycles per k-block)
  static const int small_kb = [] { const char* e = getenv("MMNN_ROWS_SMALL_SMEM_KB"); return e ? atoi(e) : 200; }();
  // small-grid forward launches: raw activation k-blocks by TMA, BN+ReLU in place (MMNN_ROWS_TMA=0: register staging)
  static const bool rows_tma_on = [] { const char* e = getenv("MMNN_ROWS_TMA"); return !(e != nullptr && e[0] == '0'); }();
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  p.tma_a = 0;
  if (PF == 2 && rows_tma_on && AMODE == A_LINEAR_CONV && TRANS == T_BNRELU && !GRAD && kActF16 && p.ntaps == 1 && p.kbw == 64 &&
      make_tmap_rows_kmajor(&tmap, p.a_src, p.a_pitch, p.M))
    p.tma_a = 1;
  if (p.stages <= 0) {
    if (p.tma_a) {
      const int KB = (p.Cin + p.kbw - 1) / p.kbw;
      p.stages = 1;
      for (int s = 1; s <= 6 && s <= (KB > 1 ? KB : 1); ++s)
        if (rows_smem_layout(p.Cin, p.NT, p.kbw, s, offs, true) <= (uint32_t)small_kb * 1024) p.stages = s;
    } else {
      p.stages = choose_stages(p, (PF == 2 ? small_kb : (GRAD ? 112 : 100)) * 1024);
    }
  }
  const uint32_t smem = rows_smem_layout(p.Cin, p.NT, p.kbw, p.stages, offs, p.tma_a != 0);
  static const bool tc_stats_on = [] { const char* e = getenv("MMNN_TC_STATS"); return e != nullptr && e[0] == '1'; }();
  if (tc_stats_on) p.stages |= 0x100;    // experiment switch: column statistics on the tensor core (engine.cuh)
  auto kern = conv_rows_kernel<AMODE, TRANS, EPI, GRAD, PF>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((p.M + TILE_ROWS - 1) / TILE_ROWS, (p.Ncols + p.NT - 1) / p.NT);
  static const bool early_on = [] { const char* e = getenv("MMNN_EARLY_START"); return !(e != nullptr && e[0] == '0'); }();
  if (PF < 2 || !early_on) p.early_ch = 0;
  if (p.early_ch > 0) launch_pdl_forced(kern, grid, dim3(ENGINE_THREADS), smem, stream, p, tmap);   // starts while the previous layer's 3x3x3 conv runs
  else launch_pdl(kern, grid, dim3(ENGINE_THREADS), smem, stream, p, tmap);
  MMNN_CHECK_LAUNCH();
  return 0;
}

// Persistent warp-specialised 1x1x1 kernel (rows_persist.cuh): one CTA per SM, all tiles of a CTA share one N tile
template <int TRANS, int EPI, bool GRAD>
int launch_rows_persist(RowsParams p, cudaStream_t stream) {
  uint32_t offs[6];
  static const bool tc_stats_off = [] { const char* e = getenv("MMNN_TC_STATS"); return e != nullptr && e[0] == '0'; }();
  const bool tcs = EPI != EP_STORE && p.NT == 128 && !tc_stats_off;   // persistent kernel: tensor-core statistics by default
  static const bool rows_tma_on = [] { const char* e = getenv("MMNN_ROWS_TMA"); return !(e != nullptr && e[0] == '0'); }();
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  p.tma_a = 0;
  if (rows_tma_on && TRANS == T_BNRELU && !GRAD && kActF16 && p.kbw == 64 && make_tmap_rows_kmajor(&tmap, p.a_src, p.a_pitch, p.M)) p.tma_a = 1;
  int stages = 1;
  for (int s = 1; s <= 6; ++s)
    if (rowsp_smem_layout(p.Cin, p.NT, p.kbw, s, tcs, offs, p.tma_a != 0) <= 220 * 1024) stages = s;
  const uint32_t smem = rowsp_smem_layout(p.Cin, p.NT, p.kbw, stages, tcs, offs, p.tma_a != 0);
  p.stages = stages | (tcs ? 0x100 : 0);
  auto kern = conv1_persist_kernel<TRANS, EPI, GRAD>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int tiles_m = (p.M + TILE_ROWS - 1) / TILE_ROWS, ntn = (p.Ncols + p.NT - 1) / p.NT;
  int per_n = 148 / ntn;
  if (per_n > tiles_m) per_n = tiles_m;
  if (per_n < 1) return -2;
  launch_pdl(kern, dim3(per_n * ntn), dim3(RP_THREADS), smem, stream, p, tmap);
  MMNN_CHECK_LAUNCH();
  return 0;
}

template <int AMODE, int TRANS, int EPI, bool GRAD>
int launch_rows_t(const RowsParams& p, cudaStream_t stream) {
  const long long tiles = (long long)((p.M + TILE_ROWS - 1) / TILE_ROWS) * ((p.Ncols + p.NT - 1) / p.NT);
  // Big 1x1x1 layers (at least two tiles per SM): the persistent kernel (rows_persist.cuh) with tensor-core column statistics
  // accumulated in TMEM.  Same-box A/B at configs[1] (round 1): forward 1.68 vs 1.76 ms -> ON for the forward; data
  // gradient 2.04 vs 1.86 ms (its ReLU-mask + second-statistic epilogue has 8 warps per SM there instead of 16) -> OFF.
  // MMNN_ROWS_PERSIST=0 / 1 forces the one-tile-per-CTA / the persistent kernel for both (parity tests run both).
  static const int persist_env = [] { const char* e = getenv("MMNN_ROWS_PERSIST"); return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0); }();
  const bool persist = persist_env == 1 || (persist_env == -1 && !GRAD && EPI == EP_STORE_STATS);
  if (AMODE == A_LINEAR_CONV && persist && EPI != EP_MASK_STATS_ACC && p.ntaps == 1 && p.NT <= 256 && (p.Ncols + p.NT - 1) / p.NT <= 148 && tiles >= 2 * 148)
    return launch_rows_persist<TRANS, EPI, GRAD>(p, stream);
  // (a deeper register prefetch, PF = 4, spills under the 2-CTA register cap: 1343 vs 1423 volumes/s)
  if (tiles <= 148) return launch_rows_pf<AMODE, TRANS, EPI, GRAD, 2>(p, stream);   // latency-bound small grids
  return launch_rows_pf<AMODE, TRANS, EPI, GRAD, 1>(p, stream);
}

// grad == 0: forward GEMM (activation-format operands and output); grad == 1: data-gradient GEMM (bf16 operands/output)
int launch_rows(const RowsParams& p, int amode, int trans, int epi, int grad, cudaStream_t stream) {
  if (p.NT % 32 != 0 || p.NT > 256 || p.Cin % 32 != 0 || (p.kbw != 32 && p.kbw != 64)) return -2;
#define CASE(A, T, E, G) if (amode == A && trans == T && epi == E && grad == (G ? 1 : 0)) return launch_rows_t<A, T, E, G>(p, stream);
  CASE(A_LINEAR_CONV, T_NONE, EP_STORE, false)
  CASE(A_LINEAR_CONV, T_NONE, EP_STORE_STATS, false)
  CASE(A_LINEAR_CONV, T_BNRELU, EP_STORE_STATS, false)
  CASE(A_LINEAR_CONV, T_BNRELU, EP_STORE, false)
  CASE(A_STEM, T_NONE, EP_STORE_STATS, false)
  CASE(A_LINEAR_CONV, T_NONE, EP_STORE, true)
  CASE(A_LINEAR_CONV, T_NONE, EP_MASK_STATS, true)
  CASE(A_LINEAR_CONV, T_NONE, EP_MASK_STATS_ACC, true)
  CASE(A_LINEAR_CONV, T_BNBWD, EP_MASK_STATS_ACC, true)
#undef CASE
  return -3;
}

template <int TRANS, int EPI, bool GRAD, int TP>
int launch_brick_tp(const BrickParams& p, int ntiles, cudaStream_t stream) {
  uint32_t offs[6];
  const uint32_t smem = brick_smem_layout(p.CH, p.NT, TP, offs);
  auto kern = conv3_brick_kernel<TRANS, EPI, GRAD, TP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int ngroups = (ntiles + TP - 1) / TP;
  const int grid = ngroups < 148 ? ngroups : 148;   // persistent: one CTA per SM
  launch_pdl(kern, dim3(grid), dim3(brick_threads(GRAD)), smem, stream, p);
  MMNN_CHECK_LAUNCH();
  return 0;
}

// Tile pairs (two tiles per weight pass, brick.cuh) halve the weight re-fetch from L2.  Measured on B200 at block-1 size
// (round 1): data gradient 110 us either way, forward 151 -> 222 us (the 4-buffer ring then looks only one channel
// quarter ahead) -- the weight traffic is not what bounds these kernels, so pairs are OFF unless MMNN_BRICK_TP=2 asks
// for them (kept for experiments; covered by the parity tests under that setting).
template <int TRANS, int EPI, bool GRAD>
int launch_brick_t(const BrickParams& p, cudaStream_t stream) {
  const int ntiles = p.B * p.Dz * ((p.Dy + BR_TY - 1) / BR_TY) * ((p.Dx + BR_TX - 1) / BR_TX);
  static const int forced = [] { const char* e = getenv("MMNN_BRICK_TP"); return e ? atoi(e) : 0; }();
  const bool pair = forced == 2;
  return pair ? launch_brick_tp<TRANS, EPI, GRAD, 2>(p, ntiles, stream) : launch_brick_tp<TRANS, EPI, GRAD, 1>(p, ntiles, stream);
}

// 3x3x3 convolution in brick mode: grad == 0 forward (BN+ReLU prologue, store + statistics), grad == 1 data gradient
// (raw gradient operand, ReLU mask + BN-backward statistics epilogue)
int launch_brick(const BrickParams& p, int grad, cudaStream_t stream) {
  if (grad == 2) {   // ablation (microbenchmarks only): forward shape without the BN/ReLU transform
    if (p.CH != 128 || p.NT != 32) return -2;
    return launch_brick_t<T_NONE, EP_STORE_STATS, false>(p, stream);
  }
  if (grad == 0) {
    if (p.CH != 128 || p.NT != 32) return -2;
    return launch_brick_t<T_BNRELU, EP_STORE_STATS, false>(p, stream);
  }
  if (p.CH != 32 || p.NT != 128) return -2;
  return launch_brick_t<T_NONE, EP_MASK_STATS, true>(p, stream);
}

// Stem convolution in brick mode (stem.cuh): persistent, one CTA per SM
int launch_stem_brick(const StemBrickParams& p, cudaStream_t stream) {
  if (p.Sz != p.D0 + 3 || p.Sy != p.H0 + 3 || p.Sx != p.W0 + 3) return -2;
  cudaError_t e = cudaFuncSetAttribute(stem_brick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM);
  if (e != cudaSuccess) return (int)e;
  const long long ntiles = (long long)p.B * p.D0 * ((p.H0 + SB_TY - 1) / SB_TY) * ((p.W0 + SB_TX - 1) / SB_TX);
  if (ntiles <= 0 || ntiles > 0x7fffffffLL) return -2;
  const int grid = ntiles < 148 ? (int)ntiles : 148;
  launch_pdl(stem_brick_kernel, dim3(grid), dim3(SB_THREADS), (size_t)SB_SMEM, stream, p);
  MMNN_CHECK_LAUNCH();
  return 0;
}

// Output-tile grid (gy, gz) of a weight-gradient launch and the voxel split (grid x).  Atomic reduction: as many CTAs as fill the
// GPU.  Slotted (deterministic) reduction: every CTA of the split writes a full partial copy of its output tile, so the split
// is additionally capped at one CTA per MMNN_WGRAD_MIN_TILES voxel tiles (default 2; measured on B200 at configs[1]: 1, 2, 3 match the atomic reduction at 1072-1076 volumes/s, 4 -> 1058, 8 -> 1021, 16 -> 935): the small late-block layers (32 / 4
// voxel tiles) then write 16 / 2 partial copies instead of 32 / 4.
static int stem_np_default() {   // k-block pairs per CTA of the stem weight gradient (MMNN_STEM_NP=1: 8 z tiles, deeper ring)
  static const int v = [] { const char* e = getenv("MMNN_STEM_NP"); return (e != nullptr && e[0] == '1') ? 1 : 2; }();
  return v;
}
void wgrad_grid(const WgradParams& p, int kind, int& gy, int& gz) {
  gz = (p.na_total + 127) / 128;
  gy = (p.nb_total + p.CB - 1) / p.CB;
  if (kind == 1) { gy = 3; gz = 1; }
  if (kind == 3) { const int np = (p.NP == 1 || p.NP == 2) ? p.NP : stem_np_default(); gy = 1; gz = 8 / np; }
}
int wgrad_split(const WgradParams& p, int kind, bool slotted) {
  int gy, gz;
  wgrad_grid(p, kind, gy, gz);
  const int ntiles = (p.M + TILE_ROWS - 1) / TILE_ROWS;
  int split = 148 / (gy * gz);
  if (slotted) {
    static const int min_tiles = [] { const char* e = getenv("MMNN_WGRAD_MIN_TILES"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    static const int small_tiles = [] { const char* e = getenv("MMNN_WGRAD_SMALL_TILES"); return e ? atoi(e) : 0; }();
    const int mt = (small_tiles > 0 && ntiles <= 64) ? small_tiles : min_tiles;   // experiment: a different cap for the late blocks only
    const int cap = (ntiles + mt - 1) / mt;
    if (split > cap) split = cap;
  }
  if (split < 1) split = 1;
  if (split > ntiles) split = ntiles;
  return split;
}

template <int AMODE, int ATRANS, int BTRANS, int EMODE>
int launch_wgrad_t(WgradParams p, int split, int gy, int gz, cudaStream_t stream) {
  static const bool piped_on = [] { const char* e = getenv("MMNN_WGRAD_PIPED"); return e != nullptr && e[0] == '1'; }();
  if (piped_on && p.NB == 9) p.NP = -1;   // experiment switch: cp.async-pipelined producer for the 3x3x3 weight gradient (engine.cuh)
  // gradient (B) operand through the TMA unit when a 128-voxel tile is a box of the volume and the operand needs no transform
  static const bool tma_on = [] { const char* e = getenv("MMNN_WGRAD_TMA"); return !(e != nullptr && e[0] == '0'); }();
  CUtensorMap tmb;
  memset(&tmb, 0, sizeof(tmb));
  p.tma_b = 0;
  if (tma_on && BTRANS == T_NONE && p.NP != -1 && wgrad_tma_box(p.Dz, p.Dy, p.Dx, p.bx, p.by, p.bz, p.bn)) {
    const long long vps = (long long)p.Dz * p.Dy * p.Dx;
    const int N = (int)((p.M + vps - 1) / vps);
    // 3x3x3: one (by + 2)-row box per (dz, dx) when the tile lies in one z slice (engine.cuh, tma_b == 2); MMNN_WGRAD_HALO=0: nine boxes
    static const bool bhalo_on = [] { const char* e = getenv("MMNN_WGRAD_HALO"); return !(e != nullptr && e[0] == '0'); }();
    const bool bhalo = bhalo_on && AMODE == WA_LINEAR && p.NB == 9 && p.CB == 32 && p.b_pitch == 32 && p.bz == 1 && p.bn == 1 && p.bx % 8 == 0;
    if (make_tmap_ndhwc(&tmb, p.b_src, p.b_pitch, N, p.Dz, p.Dy, p.Dx, p.bx, bhalo ? p.by + 2 : p.by, p.bz, p.bn)) p.tma_b = bhalo ? 2 : 1;
  }
  const uint32_t b_halo_bytes = p.tma_b == 2 ? wgrad_tap_halo_bytes(p.bx, p.by) : 0u;
  // stem: the space-to-depth operand by TMA too when the caller holds it in bf16 (MMNN_STEM_WGRAD_TMA=0: register path)
  static const bool tma_a_on = [] { const char* e = getenv("MMNN_STEM_WGRAD_TMA"); return !(e != nullptr && e[0] == '0'); }();
  CUtensorMap tma;
  memset(&tma, 0, sizeof(tma));
  p.tma_a = 0;
  if (AMODE == WA_STEM_PAIR && tma_a_on && p.tma_b && p.a_bf16 == 1 && p.a_pitch == 16 && p.Sz >= p.Dz + 3 && p.Sy >= p.Dy + 3 && p.Sx >= p.Dx + 3) {
    const long long vps = (long long)p.Dz * p.Dy * p.Dx;
    const int N = (int)((p.M + vps - 1) / vps);
    // one (by + 3)-row box per element half when the tile lies in one z slice (engine.cuh, tma_a == 2); MMNN_STEM_WGRAD_HALO=0: four boxes
    static const bool halo_on = [] { const char* e = getenv("MMNN_STEM_WGRAD_HALO"); return !(e != nullptr && e[0] == '0'); }();
    const bool halo = halo_on && p.NP == 2 && p.bz == 1 && p.bn == 1 && p.bx % 8 == 0 && p.by + 3 <= 256;
    if (make_tmap_s2d_rows(&tma, p.a_src, N, p.Sz, p.Sy, p.Sx, p.bx, halo ? p.by + 3 : p.by, p.bz, p.bn)) p.tma_a = halo ? 2 : 1;
  }
  // 1x1x1 / 3x3x3: the raw activation tile by TMA too, BN+ReLU applied in place (engine.cuh, araw); MMNN_WGRAD_TMA_A=0: register staging
  static const bool araw_on = [] { const char* e = getenv("MMNN_WGRAD_TMA_A"); return !(e != nullptr && e[0] == '0'); }();
  if (AMODE == WA_LINEAR && ATRANS == T_BNRELU && araw_on && p.tma_b && p.NP != -1) {
    const long long vps = (long long)p.Dz * p.Dy * p.Dx;
    const int N = (int)((p.M + vps - 1) / vps);
    if (make_tmap_ndhwc(&tma, p.a_src, p.a_pitch, N, p.Dz, p.Dy, p.Dx, p.bx, p.by, p.bz, p.bn)) p.tma_a = 1;
  }
  const uint32_t halo_bytes = p.tma_a == 2 ? wgrad_stem_halo_bytes(p.bx, p.by) : 0u;
  uint32_t offs[4];
  if (p.stages <= 0) {
    p.stages = 1;
    for (int s = 1; s <= 4; ++s)
      if (wgrad_smem_layout(p.CB, p.NB, s, p.NP < 1 ? 1 : p.NP, offs, p.tma_b != 0, halo_bytes, b_halo_bytes) <= 225 * 1024) p.stages = s;
  }
  const uint32_t smem = wgrad_smem_layout(p.CB, p.NB, p.stages, p.NP < 1 ? 1 : p.NP, offs, p.tma_b != 0, halo_bytes, b_halo_bytes);
  auto kern = conv_wgrad_kernel<AMODE, ATRANS, BTRANS, EMODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  launch_pdl(kern, dim3(split, gy, gz), dim3(WGRAD_THREADS), smem, stream, p, tmb, tma);
  MMNN_CHECK_LAUNCH();
  return 0;
}

// kind: 0 = conv1x1 (A = BN+ReLU(activation tile), B = raw gradient tile), 1 = conv3x3x3 (A = BN+ReLU(bottleneck),
// B = 9 shifted raw gradient tiles per CTA), 2 = raw x raw (transition), 3 = stem (A = space-to-depth rows, B = raw)
// split <= 0: chosen by wgrad_split().  With p.slot_stride > 0 the caller provides `split` slots of slot_stride floats at p.dw.
int launch_wgrad(const WgradParams& p, int kind, int split, cudaStream_t stream) {
  if (p.CB % 32 != 0 || p.CB > 128) return -2;
  if (p.slot_stride < 0 || (p.slot_stride & 3) != 0) return -2;
  int gy, gz;
  wgrad_grid(p, kind, gy, gz);
  const int ntiles = (p.M + TILE_ROWS - 1) / TILE_ROWS;
  if (split <= 0) split = wgrad_split(p, kind, p.slot_stride > 0);
  if (split > ntiles) split = ntiles;
  switch (kind) {
    case 0: return launch_wgrad_t<WA_LINEAR, T_BNRELU, T_NONE, WE_STRIDED>(p, split, gy, gz, stream);
    case 1: return launch_wgrad_t<WA_LINEAR, T_BNRELU, T_NONE, WE_STRIDED>(p, split, gy, gz, stream);
    case 2: return launch_wgrad_t<WA_LINEAR, T_NONE, T_NONE, WE_STRIDED>(p, split, gy, gz, stream);
    case 3: {
      WgradParams q = p;
      if (q.NP != 1 && q.NP != 2) q.NP = stem_np_default();   // two k-block pairs per CTA share one gradient tile
      return launch_wgrad_t<WA_STEM_PAIR, T_NONE, T_NONE, WE_STEM>(q, split, gy, gz, stream);
    }
  }
  return -3;
}

}  // namespace mmnn

namespace mmnn {
ProfState& prof_state() {
  static ProfState s;
  return s;
}
}  // namespace mmnn

extern "C" {

void mmnn_profile_enable(int on) { prof_state().on = on != 0; prof_state().timeline = on == 2; }
// Timeline of the records taken with mmnn_profile_enable(2): cls[i], start / end in ms relative to the first record's start.
// Returns the number of records (clears them).  Events sit between kernels of the real streams: a launch's interval includes the
// time it waited for its stream predecessors / cross-stream events after its start event was reached.
int mmnn_profile_timeline(int* cls, float* t0, float* t1, int cap) {
  ProfState& s = prof_state();
  int n = 0;
  if (!s.recs.empty()) {
    cudaEventSynchronize(s.recs.back().b);
    cudaDeviceSynchronize();
    for (auto& r : s.recs) {
      if (n < cap) {
        float a = 0.f, b = 0.f;
        cudaError_t ea = cudaEventElapsedTime(&a, s.recs[0].a, r.a);
        cudaError_t eb = cudaEventElapsedTime(&b, s.recs[0].a, r.b);
        if (ea != cudaSuccess) a = -(float)ea;
        if (eb != cudaSuccess) b = -(float)eb;
        (void)cudaGetLastError();
        cls[n] = r.cls; t0[n] = a; t1[n] = b; ++n;
      }
    }
    for (auto& r : s.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  }
  s.recs.clear();
  return n;
}
long long mmnn_launch_count() { return prof_state().launches; }
// Synchronises, sums the event-timed durations per kernel class, clears the records. ms / counts: PC_COUNT entries.
int mmnn_profile_collect(float* ms, int* counts) {
  ProfState& s = prof_state();
  for (int i = 0; i < PC_COUNT; ++i) { ms[i] = 0.f; counts[i] = 0; }
  for (auto& r : s.recs) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.cls] += t; counts[r.cls] += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  s.recs.clear();
  return PC_COUNT;
}

int mmnn_conv3_brick(const BrickParams* p, int grad, void* stream) { return launch_brick(*p, grad, (cudaStream_t)stream); }
int mmnn_sizeof_brick_params() { return (int)sizeof(BrickParams); }
int mmnn_stem_brick(const StemBrickParams* p, void* stream) { return launch_stem_brick(*p, (cudaStream_t)stream); }
int mmnn_sizeof_stem_brick_params() { return (int)sizeof(StemBrickParams); }

int mmnn_conv_wgrad(const WgradParams* p, int kind, int split, void* stream) {
  return launch_wgrad(*p, kind, split, (cudaStream_t)stream);
}
int mmnn_sizeof_wgrad_params() { return (int)sizeof(WgradParams); }

int mmnn_conv_rows(const RowsParams* p, int amode, int trans, int epi, int grad, void* stream) {
  return launch_rows(*p, amode, trans, epi, grad, (cudaStream_t)stream);
}
int mmnn_act_is_fp16() { return kActF16 ? 1 : 0; }

// descs: HOST array of n PackDesc; dev_descs: device scratch of n*sizeof(PackDesc) bytes.
int mmnn_pack_weights(const PackDesc* descs, int n, void* dev_descs, void* stream) {
  cudaError_t e = cudaMemcpyAsync(dev_descs, descs, sizeof(PackDesc) * n, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  pack_weights_kernel<<<dim3(32, n), 256, 0, (cudaStream_t)stream>>>((const PackDesc*)dev_descs);
  MMNN_CHECK_LAUNCH();
  return 0;
}

int mmnn_sizeof_rows_params() { return (int)sizeof(RowsParams); }
int mmnn_sizeof_pack_desc() { return (int)sizeof(PackDesc); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05.mma issue-rate microbenchmark (instrumentation; profiles/scripts/microbench_mma.py): one CTA, one elected thread issues
// `iters` back-to-back MMAs  D[128 x N] += A[128 x 16] * B[N x 16]  from fixed shared-memory operands and commits; the
// elapsed SM clock cycles between the first issue and the arrival of the commit are written to out[0].
// layout 0: SWIZZLE_NONE core matrices (LBO = plane stride, SBO = 128 B), 6: SWIZZLE_32B rows of 32 B (SBO = 256 B).
namespace mmnn {
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int layout, int iters, int a_step_bytes, long long* out, int nacc = 2) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase, tptr = sbase + 32, sA = sbase + 1024, sB = sA + 64 * 1024;   // B: 32 KB (N = 128 in MN-major groups)
  for (int i = threadIdx.x; i < (64 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;  // fp16 1.0
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(tptr, 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(smem + 32);
  if (warp == 0) {
    // layout 100 / 101: the weight-gradient kernel's MN-major operands (K = voxel rows).  100: A = SWIZZLE_NONE chunk planes,
    // B = SWIZZLE_64B 32-channel groups (3x3x3 / 1x1x1 weight gradient); 101: both SWIZZLE_64B groups (all-TMA stem)
    const bool mn = layout >= 100;
    const uint32_t idesc = mn ? make_idesc_ab(128, N, 1, 1, false, false) : make_idesc(128, N, 0, 0, true);
    const uint64_t ad0 = layout == 100 ? make_smem_desc(sA, 128, PLANE_BYTES)
                         : layout == 101 ? make_smem_desc_sw(sA, 8192, 512, 4u)
                         : layout == 0 ? make_smem_desc(sA, 2064, 128) : make_smem_desc_sw(sA, 16, 256, (uint32_t)layout);
    const uint64_t bd0 = mn ? make_smem_desc_sw(sB, 8192, 512, 4u)
                         : layout == 0 ? make_smem_desc(sB, (uint32_t)N * 16, 128) : make_smem_desc_sw(sB, 16, 256, (uint32_t)layout);
    long long t0 = 0;
    if (elect_one()) {
      const int amask = nacc - 1;
      const uint32_t acols = 512u / (uint32_t)nacc;
      t0 = clock64();
      if (layout == 102) {
        // the 3x3x3 weight gradient's issue pattern: per tile 3 accumulators (columns 0 / 96 / 192) x 8 K steps, halo-box B operand
        const uint64_t a0 = make_smem_desc(sA, 128, PLANE_BYTES);
        const uint64_t b0 = make_smem_desc_sw(sB, 10240, 512, 4u);
        for (int i = 0; i < iters / 24; ++i) {
#pragma unroll
          for (int h = 0; h < 3; ++h) {
            const uint64_t bd = desc_advance(b0, (uint32_t)(2 - h) * 1024u);
#pragma unroll
            for (int k16 = 0; k16 < 8; ++k16)
              tc_mma_bf16(tmem_base + h * 96, desc_advance(a0, k16 * 256), desc_advance(bd, k16 * 1024), idesc, (i > 0 || k16 > 0) ? 1u : 0u);
          }
          if (nacc == 1) tc_commit(bar + 8);     // a commit per tile like the kernel's (to a second, never awaited barrier)
        }
      } else
      for (int i = 0; i < iters; ++i)
        tc_mma_bf16(tmem_base + (uint32_t)(i & amask) * acols, desc_advance(ad0, (uint32_t)((i & 7) * a_step_bytes)), bd0, idesc, i >= nacc ? 1u : 0u);
      tc_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0, 90);
    const long long t1 = clock64();
    if (t0 != 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}
}  // namespace mmnn

extern "C" int mmnn_mma_rate(int N, int layout, int iters, int a_step_bytes, long long* out_cycles, void* stream) {
  const int smem = (1 + 64 + 32 + 1) * 1024;
  cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, layout, iters, a_step_bytes, out_cycles);
  return (int)cudaGetLastError();
}

// nacc: accumulators the back-to-back MMAs rotate over (1 = every MMA accumulates into the same TMEM columns)
extern "C" int mmnn_mma_rate_acc(int N, int layout, int iters, int a_step_bytes, int nacc, long long* out_cycles) {
  const int smem = (1 + 64 + 32 + 1) * 1024;
  if (nacc < 1 || nacc > 4 || nacc == 3 || N * nacc > 512) return -2;
  cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  mma_rate_kernel<<<1, 128, smem, 0>>>(N, layout, iters, a_step_bytes, out_cycles, nacc);
  return (int)cudaGetLastError();
}
#ifdef MMNN_WGRAD_TIMING
extern "C" int mmnn_wgrad_dbg(unsigned long long* out, int reset) {
  if (reset) { unsigned long long z[12] = {0}; return (int)cudaMemcpyToSymbol(mmnn::g_wgrad_dbg, z, sizeof(z)); }
  return (int)cudaMemcpyFromSymbol(out, mmnn::g_wgrad_dbg, 12 * sizeof(unsigned long long));
}
#endif
#ifdef MMNN_STEM_TIMING
extern "C" int mmnn_stem_dbg(long long* out) { return (int)cudaMemcpyFromSymbol(out, mmnn::g_stem_dbg, 8 * sizeof(long long)); }
#endif
