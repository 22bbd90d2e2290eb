// Host-side orchestration of the 3-D DenseNet trunk (reference: /root/reference/models/densenet.py:196-231 backbone,
// :46-148 dense layer / block / transition).  One C-ABI call enqueues the whole forward (or backward) on a stream:
// no allocation, no synchronisation, no host round trip inside (CUDA-graph capturable).
//
// Data layout in HBM (all activations NDHWC, bf16):
//   xs2d      padded space-to-depth image      [B][D0+3][H0+3][W0+3][16]
//   stem_out  raw conv0 output                 [M0][64]
//   buf[b]    ONE buffer per dense block       [M_b][Ctot_b]   layer l writes channels [c0+32l, c0+32l+32) in place
//                                               (the reference's torch.cat copies disappear)
//   bott[b,l] raw bottleneck (conv1 output)    [M_b][128]      kept for backward
//   pooled[b] avgpool(relu(bn(buf[b])))        [M_b/8][Ctot_b] transition: pooling commutes with the 1x1x1 conv
//   statistics: fp64 per-channel sum / sum-of-squares, produced by the epilogue of whichever kernel writes a channel,
//   consumed by the prologue of every kernel that normalises it.
#include <string.h>

#include <algorithm>
#include <vector>

#include "brick.cuh"
#include "eltwise.cuh"
#include "engine.cuh"
#include "pack.cuh"
#include "launch.h"
#include "prof.h"
#include "stem.cuh"

namespace mmnn {
int launch_stem_brick(const StemBrickParams& p, cudaStream_t stream);
int launch_rows(const RowsParams& p, int amode, int trans, int epi, int grad, cudaStream_t stream);
int launch_wgrad(const WgradParams& p, int kind, int split, cudaStream_t stream);
int wgrad_split(const WgradParams& p, int kind, bool slotted);
int launch_brick(const BrickParams& p, int grad, cudaStream_t stream);
}  // namespace mmnn

using namespace mmnn;

namespace {

constexpr int GROWTH = 32, BOTT = 128, INIT_F = 64, NUM_SMS = 148;
// Dense blocks of at most this many voxel rows are serial chains of small launches: there the BatchNorm backward of norm2 is
// applied by the producers of the 1x1x1 data-gradient GEMM (T_BNBWD, engine.cuh) instead of a launch of its own.
constexpr long long FUSE_BN2_MAX_M = 8192;

struct BnInfo {
  int C;
  int fwd_off;    // channel offset in the forward statistics arena
  int bwd_off;    // channel offset in the backward statistics arena
  int param_idx;  // gamma; beta = +1
  int buf_idx;    // running_mean; running_var = +1; num_batches_tracked = +2
};

struct LayerInfo {
  int cin;
  int nt_d;     // column-tile width of the 1x1x1 data gradient: cin itself for 128 < cin <= 256 (ONE tile: the gradient operand is
                // read once and no CTA works on a mostly empty second tile), else 128
  BnInfo n1, n2;
  int conv1_idx, conv2_idx;
  size_t pk_c1f, pk_c1d, pk_c2f, pk_c2d;  // packed weight offsets (elements)
  int index;                              // global dense-layer index (dropout mask / conv2 scratch slot)
};

struct BlockInfo {
  int c0, ctot, fwd_off;
  std::vector<LayerInfo> layers;
  bool has_trans;
  BnInfo tn;
  int tconv_idx;
  size_t pk_tf, pk_td;
};

struct Plan {
  int cin_real;
  std::vector<BlockInfo> blocks;
  BnInfo n0, n5;
  int conv0_idx;
  size_t pk_stem;
  int num_params, num_buffers, num_bn, num_layers;
  int fwd_channels, bwd_channels;
  size_t packed_elems_total;
  std::vector<long long> param_numel;
  std::vector<BnInfo*> bn_order;
  // device tables are re-uploaded only when their content (pointer sets) or destination changes: keeps the steady
  // state free of pageable host->device copies (which synchronise the stream and cannot be graph-captured)
  std::vector<uint8_t> cache[4];
  const void* cache_dst[4] = {nullptr, nullptr, nullptr, nullptr};
  // pinned staging of the table uploads: a ring of STAGE_RING buffers per slot, each guarded by an event recorded right
  // after its copy was enqueued -- two workspaces used alternately (their tables differ) never overwrite the source of a
  // copy that has not executed yet
  static constexpr int STAGE_RING = 4;
  void* pinned[4][STAGE_RING] = {};
  size_t pinned_bytes[4][STAGE_RING] = {};
  cudaEvent_t pinned_ev[4][STAGE_RING] = {};
  bool pinned_used[4][STAGE_RING] = {};
  int pinned_next[4] = {0, 0, 0, 0};
  // Weight-gradient GEMMs are off the critical path of backward (nothing downstream reads them): they run on a second
  // stream, forked / joined with events, so the small late-block launches overlap the data-gradient chain.
  cudaStream_t side = nullptr;
  // Weight gradients: the voxel split of every weight-gradient GEMM is reduced through slots summed in index order
  // (deterministic, the default) or, with MMNN_DETERMINISTIC=0, by floating-point atomics (run-dependent last bits).
  bool det = true;
  // Gradient groups, in the order backward finalises them: group k = dense block nb-1-k with the transition that
  // follows it (the last block's group also holds norm5), group nb = the stem (conv0, norm0).  Each group is a
  // contiguous range of the parameter-order gradient buffer; grad_ev[k] is recorded when its last writer has been
  // enqueued, so a data-parallel reducer can all-reduce it while the earlier blocks are still in backward.
  std::vector<cudaEvent_t> grad_ev;
  std::vector<cudaEvent_t> events;
  size_t ev_next = 0;
  cudaEvent_t next_event() {
    if (ev_next == events.size()) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      events.push_back(e);
    }
    return events[ev_next++];
  }
};

int upload_table(Plan* pl, int slot, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  std::vector<uint8_t>& c = pl->cache[slot];
  if (pl->cache_dst[slot] == dst && c.size() == bytes && memcmp(c.data(), src, bytes) == 0) return 0;
  // the copy is sourced from a pinned staging buffer owned by the plan: asynchronous, and still valid if the copy was
  // captured into a CUDA graph and is replayed later
  const int r = pl->pinned_next[slot];
  pl->pinned_next[slot] = (r + 1) % Plan::STAGE_RING;
  cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap_status);
  const bool capturing = cap_status != cudaStreamCaptureStatusNone;
  if (pl->pinned_used[slot][r] && !capturing) cudaEventSynchronize(pl->pinned_ev[slot][r]);   // the copy that last read this buffer has run
  if (pl->pinned[slot][r] == nullptr || pl->pinned_bytes[slot][r] < bytes) {
    if (pl->pinned[slot][r] != nullptr) cudaFreeHost(pl->pinned[slot][r]);
    const size_t cap = bytes < (1u << 16) ? (1u << 16) : bytes;
    if (cudaMallocHost(&pl->pinned[slot][r], cap) != cudaSuccess) return -20;
    pl->pinned_bytes[slot][r] = cap;
  }
  memcpy(pl->pinned[slot][r], src, bytes);
  cudaError_t e = cudaMemcpyAsync(dst, pl->pinned[slot][r], bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  if (!capturing) {
    if (pl->pinned_ev[slot][r] == nullptr) cudaEventCreateWithFlags(&pl->pinned_ev[slot][r], cudaEventDisableTiming);
    cudaEventRecord(pl->pinned_ev[slot][r], st);
    pl->pinned_used[slot][r] = true;
  }
  c.assign((const uint8_t*)src, (const uint8_t*)src + bytes);
  pl->cache_dst[slot] = dst;
  return 0;
}

struct Geo {
  int B, X, Y, Z;
  int D0, H0, W0, Sz, Sy, Sx;
  long long M0;
  int D[8], H[8], W[8];
  long long M[8];
  size_t xs2d, xs2dw, stem_out, argmax, fstats, bstats, packed, tables, dA2, dA1, dB2, gslice, gout, dpooled, wslots, dr, total;
  size_t wslots_bytes;
  std::vector<size_t> wslot_off;   // per parameter index: element offset of the weight-gradient slots (convolutions only)
  std::vector<int> wslot_S;        // per parameter index: number of slots == voxel split of that weight-gradient launch
  size_t maxM_;
  size_t buf[8], bott[8], pooled[8], dbuf[8];
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
bool nt_cin_enabled() {
  static const bool on = [] { const char* e = getenv("MMNN_DGRAD_NT_CIN"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}

// voxel split of a weight-gradient launch (same rule as launch_wgrad applies: capi_conv.cu)
int split_for(int kind, long long M, int na_total, int nb_total, int CB, bool slotted) {
  WgradParams w = {};
  w.M = (int)M; w.na_total = na_total; w.nb_total = nb_total; w.CB = CB; w.NP = 2;
  return wgrad_split(w, kind, slotted);
}

bool make_geo(const Plan& pl, int B, int X, int Y, int Z, Geo& g) {
  g.B = B; g.X = X; g.Y = Y; g.Z = Z;
  g.D0 = (X - 1) / 2 + 1; g.H0 = (Y - 1) / 2 + 1; g.W0 = (Z - 1) / 2 + 1;
  g.Sz = g.D0 + 3; g.Sy = g.H0 + 3; g.Sx = g.W0 + 3;
  g.M0 = (long long)B * g.D0 * g.H0 * g.W0;
  if (g.M0 <= 0 || g.M0 * 128 > 0x7fffffffLL) return false;   // kernels index voxel rows (x chunk counts) in 32 bits
  int d = (g.D0 - 1) / 2 + 1, h = (g.H0 - 1) / 2 + 1, w = (g.W0 - 1) / 2 + 1;
  const int nb = (int)pl.blocks.size();
  for (int b = 0; b < nb; ++b) {
    if (d < 1 || h < 1 || w < 1) return false;
    g.D[b] = d; g.H[b] = h; g.W[b] = w;
    g.M[b] = (long long)B * d * h * w;
    d /= 2; h /= 2; w /= 2;
  }
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  g.xs2d = take((size_t)B * g.Sz * g.Sy * g.Sx * 16 * 2 + 256);
  // bf16 copy for the stem weight gradient's TMA operand (the image itself when the activation format already is bf16)
  g.xs2dw = kActF16 ? take((size_t)B * g.Sz * g.Sy * g.Sx * 16 * 2 + 256) : g.xs2d;
  g.stem_out = take((size_t)g.M0 * 64 * 2);
  g.argmax = take((size_t)g.M[0] * 64);
  size_t maxM = 0, maxMC = 0, maxMoutC = 0, maxMoutC2 = 0;
  for (int b = 0; b < nb; ++b) {
    const BlockInfo& bi = pl.blocks[b];
    g.buf[b] = take((size_t)g.M[b] * bi.ctot * 2);
    g.bott[b] = take((size_t)g.M[b] * BOTT * 2 * bi.layers.size());
    g.dbuf[b] = take((size_t)g.M[b] * bi.ctot * 4);
    if (bi.has_trans) {
      g.pooled[b] = take((size_t)g.M[b + 1] * bi.ctot * 2);
      maxMoutC = std::max(maxMoutC, (size_t)g.M[b + 1] * bi.ctot);
      maxMoutC2 = std::max(maxMoutC2, (size_t)g.M[b + 1] * (bi.ctot / 2));
    } else {
      g.pooled[b] = 0;
    }
    maxM = std::max(maxM, (size_t)g.M[b]);
    maxMC = std::max(maxMC, (size_t)g.M[b] * (bi.ctot - GROWTH));
  }
  g.maxM_ = maxM;
  g.fstats = take((size_t)pl.fwd_channels * 2 * sizeof(double));
  g.bstats = take((size_t)pl.bwd_channels * 2 * sizeof(double));
  g.packed = take(pl.packed_elems_total * 2);
  g.tables = take(1 << 20);
  g.dA2 = take(2 * maxM * BOTT * 2);          // x2: layer parity (the side-stream wgrad of layer l reads it while l+1 runs)
  g.dA1 = take(256);                            // (the bf16 dA1 tensor of round 1 is gone: deferred BatchNorm backward)
  g.dB2 = take(2 * (size_t)FUSE_BN2_MAX_M * BOTT * 2);   // BN2-backward output of the small blocks (x2: layer parity), see backward
  g.gslice = take(2 * maxM * GROWTH * 2);
  g.gout = take(maxMoutC2 * 2 + 256);
  g.dpooled = take(maxMoutC * 2 + 256);
  {
    // weight-gradient slots: S partial copies per convolution weight (S = 1 without the ordered reduction, where only the
    // 3x3x3 gradients go through this scratch -- in [tap][co][ci] order -- and are zeroed + accumulated with atomics)
    g.wslot_off.assign(pl.num_params, 0);
    g.wslot_S.assign(pl.num_params, 0);
    size_t e = 0;
    auto add = [&](int pidx, int kind, long long M, int na, int nb, int CB) {
      const int S = pl.det ? split_for(kind, M, na, nb, CB, true) : 1;
      g.wslot_off[pidx] = e; g.wslot_S[pidx] = S;
      e += (size_t)S * (size_t)pl.param_numel[pidx];
    };
    for (int b = 0; b < nb; ++b) {
      const BlockInfo& bi = pl.blocks[b];
      for (auto& li : bi.layers) {
        if (pl.det) add(li.conv1_idx, 0, g.M[b], li.cin, BOTT, 128);
        add(li.conv2_idx, 1, g.M[b], BOTT, GROWTH, GROWTH);
      }
      if (bi.has_trans && pl.det) add(bi.tconv_idx, 2, g.M[b + 1], bi.ctot, bi.ctot / 2, 128);
    }
    if (pl.det) add(pl.conv0_idx, 3, g.M0, 128, 64, 64);
    g.wslots_bytes = e * sizeof(float);
    g.wslots = take(g.wslots_bytes + 256);
  }
  g.dr = take((size_t)g.M0 * 64 * 2);
  g.total = o;
  return true;
}

int ew_grid(long long work_items) {
  long long blocks = (work_items + EW_THREADS - 1) / EW_THREADS;
  long long cap = (long long)NUM_SMS * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

BnSrc make_bn(const BnInfo& bn, const void* const* params, void* const* buffers, const double* fstats, int fwd_channels,
              long long count, bool batch, int ch_off = 0) {
  BnSrc s;
  s.sum = fstats + bn.fwd_off + ch_off;
  s.sumsq = fstats + fwd_channels + bn.fwd_off + ch_off;
  s.gamma = (const float*)params[bn.param_idx] + ch_off;
  s.beta = (const float*)params[bn.param_idx + 1] + ch_off;
  s.rmean = (const float*)buffers[bn.buf_idx] + ch_off;
  s.rvar = (const float*)buffers[bn.buf_idx + 1] + ch_off;
  s.inv_count = 1.0f / (float)count;
  s.eps = 1e-5f;
  s.use_batch = batch ? 1 : 0;
  return s;
}

#define RET_IF(x)          \
  do {                     \
    int rc_ = (x);         \
    if (rc_ != 0) return rc_; \
  } while (0)
#define CUDA_RET(x)                          \
  do {                                       \
    cudaError_t e_ = (x);                    \
    if (e_ != cudaSuccess) return (int)e_;   \
  } while (0)
#define LAUNCH_RET()                          \
  do {                                        \
    cudaError_t e_ = cudaGetLastError();      \
    if (e_ != cudaSuccess) return (int)e_;    \
  } while (0)

}  // namespace

extern "C" {

void* mmnn_encoder_create(int in_channels, const int* block_config, int nblocks, int init_features, int growth_rate,
                          int bn_size) {
  if (in_channels < 1 || in_channels > 2 || init_features != INIT_F || growth_rate != GROWTH ||
      bn_size * growth_rate != BOTT || nblocks < 1 || nblocks > 6)
    return nullptr;
  Plan* pl = new Plan();
  pl->cin_real = in_channels;
  { const char* e = getenv("MMNN_DETERMINISTIC"); pl->det = !(e != nullptr && e[0] == '0'); }
  int pidx = 0, bidx = 0, foff = 0, boff = 0, lidx = 0;
  size_t pk = 0;
  auto add_param = [&](long long numel) { pl->param_numel.push_back(numel); return pidx++; };
  auto add_bn = [&](BnInfo& bn, int C, int fwd_off) {
    bn.C = C; bn.fwd_off = fwd_off; bn.bwd_off = boff; boff += C;
    bn.param_idx = add_param(C); add_param(C);
    bn.buf_idx = bidx; bidx += 3;
  };
  pl->conv0_idx = add_param((long long)INIT_F * in_channels * 343);
  pl->pk_stem = pk; pk += packed_elems(64, 64, 64, 64, 16);
  add_bn(pl->n0, INIT_F, foff); foff += INIT_F;
  int c = INIT_F;
  for (int b = 0; b < nblocks; ++b) {
    BlockInfo bi;
    bi.c0 = c; bi.ctot = c + block_config[b] * GROWTH; bi.fwd_off = foff; foff += bi.ctot;
    if (bi.ctot > 1024) { delete pl; return nullptr; }
    for (int l = 0; l < block_config[b]; ++l) {
      LayerInfo li;
      li.cin = c + l * GROWTH; li.index = lidx++;
      li.nt_d = (li.cin > 128 && li.cin <= 256 && nt_cin_enabled()) ? li.cin : 128;
      add_bn(li.n1, li.cin, bi.fwd_off);
      li.conv1_idx = add_param((long long)BOTT * li.cin);
      add_bn(li.n2, BOTT, foff); foff += BOTT;
      li.conv2_idx = add_param((long long)GROWTH * BOTT * 27);
      li.pk_c1f = pk; pk += packed_elems(BOTT, 128, li.cin, 64, 1);
      li.pk_c1d = pk; pk += packed_elems(li.cin, li.nt_d, BOTT, 64, 1);
      li.pk_c2f = pk; pk += packed_elems(GROWTH, 32, BOTT, 64, 27);
      li.pk_c2d = pk; pk += packed_elems(BOTT, 128, GROWTH, 32, 27);
      bi.layers.push_back(li);
    }
    c = bi.ctot;
    bi.has_trans = b != nblocks - 1;
    if (bi.has_trans) {
      if (c % 64 != 0) { delete pl; return nullptr; }
      add_bn(bi.tn, c, bi.fwd_off);
      bi.tconv_idx = add_param((long long)(c / 2) * c);
      bi.pk_tf = pk; pk += packed_elems(c / 2, 128, c, 64, 1);
      bi.pk_td = pk; pk += packed_elems(c, 128, c / 2, 64, 1);
      c /= 2;
    }
    pl->blocks.push_back(bi);
  }
  add_bn(pl->n5, c, pl->blocks.back().fwd_off);
  pl->num_params = pidx; pl->num_buffers = bidx; pl->num_bn = bidx / 3; pl->num_layers = lidx;
  pl->fwd_channels = foff; pl->bwd_channels = boff; pl->packed_elems_total = pk;
  // BN order == buffer order
  pl->bn_order.push_back(&pl->n0);
  for (auto& bi : pl->blocks) {
    for (auto& li : bi.layers) { pl->bn_order.push_back(&li.n1); pl->bn_order.push_back(&li.n2); }
    if (bi.has_trans) pl->bn_order.push_back(&bi.tn);
  }
  pl->bn_order.push_back(&pl->n5);
  return pl;
}

void mmnn_encoder_destroy(void* h) { delete (Plan*)h; }
int mmnn_encoder_num_grad_groups(void* h) { return (int)((Plan*)h)->blocks.size() + 1; }
// element range [lo, hi) of gradient group k in a buffer holding all gradients back to back in parameter order
int mmnn_encoder_grad_group_range(void* h, int k, long long* lo, long long* hi) {
  Plan* pl = (Plan*)h;
  const int nb = (int)pl->blocks.size();
  if (k < 0 || k > nb) return -2;
  int p0, p1;
  if (k == nb) { p0 = 0; p1 = pl->blocks[0].layers[0].n1.param_idx; }
  else {
    const BlockInfo& bi = pl->blocks[nb - 1 - k];
    p0 = bi.layers[0].n1.param_idx;
    p1 = bi.has_trans ? bi.tconv_idx + 1 : pl->n5.param_idx + 2;
  }
  long long off = 0, a = 0, b = 0;
  for (int i = 0; i < pl->num_params; ++i) {
    if (i == p0) a = off;
    off += pl->param_numel[i];
    if (i == p1 - 1) b = off;
  }
  *lo = a; *hi = b;
  return 0;
}
// makes `stream` wait until gradient group k of the most recent mmnn_encoder_backward is final
int mmnn_encoder_wait_grad_group(void* h, int k, void* stream) {
  Plan* pl = (Plan*)h;
  if (k < 0 || k >= (int)pl->grad_ev.size()) return -2;
  return (int)cudaStreamWaitEvent((cudaStream_t)stream, pl->grad_ev[k], 0);
}
int mmnn_encoder_num_params(void* h) { return ((Plan*)h)->num_params; }
int mmnn_encoder_num_buffers(void* h) { return ((Plan*)h)->num_buffers; }
int mmnn_encoder_num_layers(void* h) { return ((Plan*)h)->num_layers; }
long long mmnn_encoder_param_numel(void* h, int i) { return ((Plan*)h)->param_numel[i]; }
int mmnn_encoder_out_channels(void* h) { return ((Plan*)h)->n5.C; }

long long mmnn_encoder_workspace_bytes(void* h, int B, int X, int Y, int Z) {
  Geo g;
  if (!make_geo(*(Plan*)h, B, X, Y, Z, g)) return -1;
  return (long long)g.total;
}

// out_dhw: spatial dims of the trunk output (the last block's)
int mmnn_encoder_out_dims(void* h, int B, int X, int Y, int Z, int* out_dhw) {
  Geo g;
  Plan* pl = (Plan*)h;
  if (!make_geo(*pl, B, X, Y, Z, g)) return -1;
  const int nb = (int)pl->blocks.size();
  out_dhw[0] = g.D[nb - 1]; out_dhw[1] = g.H[nb - 1]; out_dhw[2] = g.W[nb - 1];
  return 0;
}

// Debug / test hook: byte offsets of the main activation buffers inside the workspace.
// offs[0]=xs2d, [1]=stem_out, [2]=argmax, [3..3+nb)=buf[b], [3+nb..3+2nb)=bott[b], [3+2nb..3+3nb)=dbuf[b], then fstats, bstats
int mmnn_encoder_debug_offsets(void* h, int B, int X, int Y, int Z, long long* offs, long long* dims) {
  Geo g;
  Plan* pl = (Plan*)h;
  if (!make_geo(*pl, B, X, Y, Z, g)) return -1;
  const int nb = (int)pl->blocks.size();
  int k = 0;
  offs[k++] = g.xs2d; offs[k++] = g.stem_out; offs[k++] = g.argmax;
  for (int b = 0; b < nb; ++b) offs[k++] = g.buf[b];
  for (int b = 0; b < nb; ++b) offs[k++] = g.bott[b];
  for (int b = 0; b < nb; ++b) offs[k++] = g.dbuf[b];
  offs[k++] = g.fstats; offs[k++] = g.bstats;
  int d = 0;
  dims[d++] = g.M0; dims[d++] = g.D0; dims[d++] = g.H0; dims[d++] = g.W0;
  for (int b = 0; b < nb; ++b) { dims[d++] = g.M[b]; dims[d++] = pl->blocks[b].ctot; dims[d++] = pl->blocks[b].c0; dims[d++] = pl->blocks[b].fwd_off; }
  dims[d++] = pl->fwd_channels;
  return 0;
}

// image   : fp32 NCDHW [B][cin][X][Y][Z]
// params  : device pointers in backbone.named_parameters() order;  buffers: in backbone.named_buffers() order
// dropmask: optional fp32 [num_layers][B][32] channel keep-mask already divided by (1-p) (Dropout3d), or null
// out     : fp32 [M_last][C_last] = norm5 output, NDHWC
static int encoder_forward_impl(void* h, int B, int X, int Y, int Z, const void* image, int image_f16, const void* const* params,
                                void* const* buffers, const float* dropmask, void* workspace, float* out, int training,
                                void* stream_) {
  Plan* pl = (Plan*)h;
  cudaStream_t st = (cudaStream_t)stream_;
  Geo g;
  if (!make_geo(*pl, B, X, Y, Z, g)) return -10;
  uint8_t* ws = (uint8_t*)workspace;
  double* fstats = (double*)(ws + g.fstats);
  const int FC = pl->fwd_channels;
  const bool batch = training != 0;
  bf16* packed = (bf16*)(ws + g.packed);
  const int nb = (int)pl->blocks.size();

  CUDA_RET(cudaMemsetAsync(fstats, 0, (size_t)FC * 2 * sizeof(double), st));

  // ---- 1. pack every conv weight (fprop + dgrad images) with one launch
  {
    std::vector<PackDesc> descs;
    auto add = [&](const void* src, size_t off, int N, int NT, int Cin, int kbw, int ntaps, int mode, long long sn,
                   long long sc, long long stt, bool fwd) {
      PackDesc d;
      d.src = (const float*)src; d.dst = packed + off; d.N = N; d.NT = NT; d.Cin = Cin; d.kbw = kbw; d.ntaps = ntaps;
      d.mode = mode; d.cin_real = pl->cin_real; d.f16 = (fwd && kActF16) ? 1 : 0; d.sn = sn; d.sc = sc; d.st = stt;
      descs.push_back(d);
    };
    add(params[pl->conv0_idx], pl->pk_stem, 64, 64, 64, 64, 16, SB_SW32 ? PACK_STEM_SW32 : PACK_STEM, 0, 0, 0, true);
    for (auto& bi : pl->blocks) {
      for (auto& li : bi.layers) {
        add(params[li.conv1_idx], li.pk_c1f, BOTT, 128, li.cin, 64, 1, PACK_GENERIC, li.cin, 1, 0, true);
        add(params[li.conv1_idx], li.pk_c1d, li.cin, li.nt_d, BOTT, 64, 1, PACK_GENERIC, 1, li.cin, 0, false);
        add(params[li.conv2_idx], li.pk_c2f, GROWTH, 32, BOTT, 64, 27, PACK_GENERIC, BOTT * 27, 27, 1, true);
        add(params[li.conv2_idx], li.pk_c2d, BOTT, 128, GROWTH, 32, 27, PACK_GENERIC, 27, BOTT * 27, 1, false);
      }
      if (bi.has_trans) {
        add(params[bi.tconv_idx], bi.pk_tf, bi.ctot / 2, 128, bi.ctot, 64, 1, PACK_GENERIC, bi.ctot, 1, 0, true);
        add(params[bi.tconv_idx], bi.pk_td, bi.ctot, 128, bi.ctot / 2, 64, 1, PACK_GENERIC, 1, bi.ctot, 0, false);
      }
    }
    if (descs.size() * sizeof(PackDesc) > (1 << 19)) return -11;
    RET_IF(upload_table(pl, 0, ws + g.tables, descs.data(), descs.size() * sizeof(PackDesc), st));
    ProfScope ps_(PC_PACK, st);
    pack_weights_kernel<<<dim3(16, (unsigned)descs.size()), 256, 0, st>>>((const PackDesc*)(ws + g.tables));
    LAUNCH_RET();
  }

  // ---- 2. stem
  bf16* xs2d = (bf16*)(ws + g.xs2d);
  bf16* xs2dw = (kActF16 && training) ? (bf16*)(ws + g.xs2dw) : nullptr;   // eval-mode forward: no weight gradient follows
  {
    const long long cells = (long long)B * g.Sz * g.Sy * g.Sx * 2;
    { ProfScope ps_(PC_S2D, st);
      if (image_f16) s2d_pack_kernel<__half><<<ew_grid(cells), EW_THREADS, 0, st>>>((const __half*)image, xs2d, xs2dw, B, pl->cin_real, X, Y, Z, g.Sz, g.Sy, g.Sx);
      else s2d_pack_kernel<float><<<ew_grid(cells), EW_THREADS, 0, st>>>((const float*)image, xs2d, xs2dw, B, pl->cin_real, X, Y, Z, g.Sz, g.Sy, g.Sx); }
    LAUNCH_RET();
    StemBrickParams p = {};
    p.B = B; p.D0 = g.D0; p.H0 = g.H0; p.W0 = g.W0; p.Sz = g.Sz; p.Sy = g.Sy; p.Sx = g.Sx;
    p.xs2d = xs2d; p.w_packed = packed + pl->pk_stem;
    p.out = (bf16*)(ws + g.stem_out); p.out_pitch = 64;
    p.st_sum = fstats + pl->n0.fwd_off; p.st_sq = fstats + FC + pl->n0.fwd_off;
    { ProfScope ps_(PC_STEM_FPROP, st); RET_IF(launch_stem_brick(p, st)); }
    PoolParams q = {};
    q.B = B; q.D0 = g.D0; q.H0 = g.H0; q.W0 = g.W0; q.D1 = g.D[0]; q.H1 = g.H[0]; q.W1 = g.W[0];
    q.src = (const bf16*)(ws + g.stem_out);
    q.bn = make_bn(pl->n0, params, buffers, fstats, FC, g.M0, batch);
    q.dst = (bf16*)(ws + g.buf[0]); q.dst_pitch = pl->blocks[0].ctot;
    q.argmax = ws + g.argmax;
    q.st_sum = fstats + pl->blocks[0].fwd_off; q.st_sq = fstats + FC + pl->blocks[0].fwd_off;
    int blocks = (int)std::min<long long>((g.M[0] + 31) / 32, NUM_SMS * 8);
    { ProfScope ps_(PC_MAXPOOL, st); launch_pdl(bnrelu_maxpool_kernel, dim3(blocks), dim3(EW_THREADS), 0, st, q); }
    LAUNCH_RET();
  }

  // ---- 3. dense blocks + transitions
  for (int b = 0; b < nb; ++b) {
    const BlockInfo& bi = pl->blocks[b];
    bf16* buf = (bf16*)(ws + g.buf[b]);
    const long long M = g.M[b];
    for (size_t l = 0; l < bi.layers.size(); ++l) {
      const LayerInfo& li = bi.layers[l];
      bf16* bott = (bf16*)(ws + g.bott[b]) + (size_t)l * M * BOTT;
      RowsParams p = {};
      p.M = (int)M; p.NT = 128; p.Ncols = BOTT; p.Cin = li.cin; p.kbw = 64; p.ntaps = 1; p.tap_sign = 1;
      p.Dz = g.D[b]; p.Dy = g.H[b]; p.Dx = g.W[b];
      p.a_src = buf; p.a_pitch = bi.ctot;
      p.bnA = make_bn(li.n1, params, buffers, fstats, FC, M, batch);
      p.b_packed = packed + li.pk_c1f;
      p.out = bott; p.out_pitch = BOTT;
      p.st_sum = fstats + li.n2.fwd_off; p.st_sq = fstats + FC + li.n2.fwd_off;
      // every channel but the 32 the previous layer's 3x3x3 conv (the kernel right before this one in the stream) is writing was
      // final before that conv started: small-grid launches work through them while it runs (engine.cuh, early start).
      // Not while bench.py's per-class timing is on (classes must not overlap).
      p.early_ch = (l > 0 && (!prof_state().on || prof_state().timeline)) ? li.cin - GROWTH : 0;
      { ProfScope ps_(PC_CONV1_FPROP, st); RET_IF(launch_rows(p, A_LINEAR_CONV, T_BNRELU, EP_STORE_STATS, 0, st)); }
      const float* cs = (dropmask != nullptr && training) ? dropmask + (size_t)li.index * B * GROWTH : nullptr;
      if (true) {  // brick mode for every spatial size: partial tiles only cost idle MMA rows, tiny layers are latency-bound anyway
        // large blocks: halo brick staged once per tile, taps are descriptor offsets (brick.cuh)
        BrickParams q = {};
        q.B = B; q.Dz = g.D[b]; q.Dy = g.H[b]; q.Dx = g.W[b]; q.CH = BOTT; q.NT = GROWTH; q.tap_sign = 1;
        q.a_src = bott; q.a_pitch = BOTT;
        q.bnA = make_bn(li.n2, params, buffers, fstats, FC, M, batch);
        q.b_packed = packed + li.pk_c2f;
        q.out = buf + li.cin; q.out_pitch = bi.ctot;
        q.colscale = cs;
        q.st_sum = fstats + bi.fwd_off + li.cin; q.st_sq = fstats + FC + bi.fwd_off + li.cin;
        ProfScope ps_(PC_CONV2_FPROP, st);
        RET_IF(launch_brick(q, 0, st));
      } else {
        RowsParams q = {};
        q.M = (int)M; q.NT = 32; q.Ncols = GROWTH; q.Cin = BOTT; q.kbw = 64; q.ntaps = 27; q.tap_sign = 1;
        q.Dz = g.D[b]; q.Dy = g.H[b]; q.Dx = g.W[b];
        q.a_src = bott; q.a_pitch = BOTT;
        q.bnA = make_bn(li.n2, params, buffers, fstats, FC, M, batch);
        q.b_packed = packed + li.pk_c2f;
        q.out = buf + li.cin; q.out_pitch = bi.ctot;
        q.colscale = cs;
        q.st_sum = fstats + bi.fwd_off + li.cin; q.st_sq = fstats + FC + bi.fwd_off + li.cin;
        ProfScope ps_(PC_CONV2_FPROP, st);
        RET_IF(launch_rows(q, A_LINEAR_CONV, T_BNRELU, EP_STORE_STATS, 0, st));
      }
    }
    if (bi.has_trans) {
      const BlockInfo& nx = pl->blocks[b + 1];
      AvgPoolParams a = {};
      a.B = B; a.D = g.D[b]; a.H = g.H[b]; a.W = g.W[b]; a.C = bi.ctot;
      a.x = buf; a.x_pitch = bi.ctot;
      a.bn = make_bn(bi.tn, params, buffers, fstats, FC, M, batch);
      a.pooled = (bf16*)(ws + g.pooled[b]);
      { ProfScope ps_(PC_TRANS_POOL, st);
        launch_pdl(bnrelu_avgpool_kernel, dim3(ew_grid(g.M[b + 1] * (bi.ctot / 8))), dim3(EW_THREADS), 2 * bi.ctot * sizeof(float), st, a); }
      LAUNCH_RET();
      RowsParams p = {};
      p.M = (int)g.M[b + 1]; p.NT = 128; p.Ncols = bi.ctot / 2; p.Cin = bi.ctot; p.kbw = 64; p.ntaps = 1; p.tap_sign = 1;
      p.Dz = g.D[b + 1]; p.Dy = g.H[b + 1]; p.Dx = g.W[b + 1];
      p.a_src = a.pooled; p.a_pitch = bi.ctot;
      p.b_packed = packed + bi.pk_tf;
      p.out = (bf16*)(ws + g.buf[b + 1]); p.out_pitch = nx.ctot;
      p.st_sum = fstats + nx.fwd_off; p.st_sq = fstats + FC + nx.fwd_off;
      { ProfScope ps_(PC_TRANS_FPROP, st); RET_IF(launch_rows(p, A_LINEAR_CONV, T_NONE, EP_STORE_STATS, 0, st)); }
    }
  }

  // ---- 4. norm5 -> fp32 output
  {
    const BlockInfo& bi = pl->blocks[nb - 1];
    const long long M = g.M[nb - 1];
    BnSrc bn = make_bn(pl->n5, params, buffers, fstats, FC, M, batch);
    ProfScope ps_(PC_NORM5, st);
    bn_apply_f32_kernel<<<ew_grid(M * (bi.ctot / 8)), EW_THREADS, 2 * bi.ctot * sizeof(float), st>>>(
        (const bf16*)(ws + g.buf[nb - 1]), bi.ctot, bn, out, M, bi.ctot);
    LAUNCH_RET();
  }

  // ---- 5. running statistics (training only): one launch for all BNs
  if (batch) {
    std::vector<BnTableEntry> tab;
    const long long counts0 = g.M0;
    auto count_of = [&](const BnInfo* bn) -> long long {
      if (bn == &pl->n0) return counts0;
      for (int b = 0; b < nb; ++b) {
        const BlockInfo& bi = pl->blocks[b];
        for (auto& li : bi.layers)
          if (bn == &li.n1 || bn == &li.n2) return g.M[b];
        if (bi.has_trans && bn == &bi.tn) return g.M[b];
      }
      return g.M[nb - 1];
    };
    for (BnInfo* bn : pl->bn_order) {
      BnTableEntry e = {};
      e.sum = fstats + bn->fwd_off; e.sumsq = fstats + FC + bn->fwd_off;
      e.running_mean = (float*)buffers[bn->buf_idx]; e.running_var = (float*)buffers[bn->buf_idx + 1];
      e.num_batches_tracked = (long long*)buffers[bn->buf_idx + 2];
      e.C = bn->C; e.count = (float)count_of(bn);
      tab.push_back(e);
    }
    uint8_t* dtab = ws + g.tables + (1 << 19);
    RET_IF(upload_table(pl, 1, dtab, tab.data(), tab.size() * sizeof(BnTableEntry), st));
    ProfScope ps_(PC_BN_RUNNING, st);
    bn_running_update_kernel<<<(unsigned)tab.size(), 128, 0, st>>>((const BnTableEntry*)dtab, 0.1f);
    LAUNCH_RET();
  }
  return 0;
}

int mmnn_encoder_forward(void* h, int B, int X, int Y, int Z, const float* image, const void* const* params,
                         void* const* buffers, const float* dropmask, void* workspace, float* out, int training,
                         void* stream_) {
  return encoder_forward_impl(h, B, X, Y, Z, image, 0, params, buffers, dropmask, workspace, out, training, stream_);
}
// same with an IEEE fp16 image [B][cin][X][Y][Z] (a loader that ships 16-bit volumes: half the host->device traffic)
int mmnn_encoder_forward_f16(void* h, int B, int X, int Y, int Z, const void* image_f16, const void* const* params,
                             void* const* buffers, const float* dropmask, void* workspace, float* out, int training,
                             void* stream_) {
  return encoder_forward_impl(h, B, X, Y, Z, image_f16, 1, params, buffers, dropmask, workspace, out, training, stream_);
}

// grad_out: fp32 [M_last][C_last] gradient w.r.t. the norm5 output (NDHWC)
// grads   : device pointers (param order) of ZERO-INITIALISED fp32 tensors shaped like the parameters
int mmnn_encoder_backward(void* h, int B, int X, int Y, int Z, const void* const* params, void* const* buffers,
                          void* const* grads, const float* dropmask, void* workspace, const float* grad_out,
                          int training, void* stream_) {
  Plan* pl = (Plan*)h;
  cudaStream_t st = (cudaStream_t)stream_;
  Geo g;
  if (!make_geo(*pl, B, X, Y, Z, g)) return -10;
  uint8_t* ws = (uint8_t*)workspace;
  const double* fstats = (const double*)(ws + g.fstats);
  double* bstats = (double*)(ws + g.bstats);
  const int FC = pl->fwd_channels, BC = pl->bwd_channels;
  // Eval-mode backward (GradCAM, frozen-BN fine-tuning): BatchNorm is the fixed affine map of its running statistics, so the
  // masks use those and the batch-statistic terms of the BN gradient vanish (inv_count = 0 -> c1 = c2 = 0); the statistics
  // the epilogues still accumulate are then exactly d(gamma) = sum dy*xhat and d(beta) = sum dy.
  const bool batch = training != 0;   // mode of the forward pass that filled this workspace (passed per call)
  bf16* packed = (bf16*)(ws + g.packed);
  const int nb = (int)pl->blocks.size();
  auto gsum = [&](const BnInfo& bn) { return bstats + bn.bwd_off; };
  auto gdot = [&](const BnInfo& bn) { return bstats + BC + bn.bwd_off; };

  CUDA_RET(cudaMemsetAsync(bstats, 0, (size_t)BC * 2 * sizeof(double), st));
  if (!pl->det) CUDA_RET(cudaMemsetAsync(ws + g.wslots, 0, g.wslots_bytes, st));   // atomics accumulate into the 3x3x3 scratch
  float* const wslots = (float*)(ws + g.wslots);
  const bool det = pl->det;
  if (pl->side == nullptr) CUDA_RET(cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking));
  // while bench.py's per-kernel profiling is on, everything stays on one stream so that class times are not overlapped
  cudaStream_t sd = (prof_state().on && !prof_state().timeline) ? st : pl->side;
  pl->ev_next = 0;
  {  // fork: the side stream starts after everything already queued on the caller's stream (zeroed gradients etc.)
    cudaEvent_t e = pl->next_event();
    CUDA_RET(cudaEventRecord(e, st));
    CUDA_RET(cudaStreamWaitEvent(sd, e, 0));
  }
  cudaEvent_t side_done[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [parity][0: conv2 wgrad read gslice, 1: conv1 wgrad read dA2]
  int parity = 0;
  cudaEvent_t trans_done = nullptr;

  auto bn_apply = [&](int out_mode, long long M, int C, const bf16* v, const float* v32, long long v_pitch, const bf16* x,
                      long long x_pitch, const BnSrc& bn, const double* gs, const double* gd, void* outp,
                      long long out_pitch, bf16* slice_out = nullptr, int slice_c0 = 0, const float* slice_scale = nullptr,
                      int vps_ = 1, int pre_rstd = 0) -> int {
    BnApplyParams a = {};
    a.slice_out = slice_out; a.slice_c0 = slice_c0; a.slice_scale = slice_scale; a.vps = vps_; a.pre_rstd = pre_rstd;
    a.M = M; a.C = C; a.v = v; a.v32 = v32; a.v_pitch = v_pitch; a.x = x; a.x_pitch = x_pitch; a.bn = bn;
    a.g_sum = gs; a.g_dot = gd; a.inv_count = batch ? 1.0f / (float)M : 0.f; a.out = outp; a.out_pitch = out_pitch;
    const int grid = ew_grid(M * (C / 8));
    const size_t sm = 3 * C * sizeof(float);
    ProfScope ps_(PC_BN_APPLY, st);
    if (out_mode == BA_OUT_BF16) launch_pdl(bn_bwd_apply_kernel<BA_OUT_BF16>, dim3(grid), dim3(EW_THREADS), sm, st, a);
    else if (out_mode == BA_OUT_F32_ADD) launch_pdl(bn_bwd_apply_kernel<BA_OUT_F32_ADD>, dim3(grid), dim3(EW_THREADS), sm, st, a);
    else launch_pdl(bn_bwd_apply_kernel<BA_OUT_F32_STORE>, dim3(grid), dim3(EW_THREADS), sm, st, a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
  };

  // Deferred BatchNorm backward of the dense blocks (eltwise.cuh E4b): dbuf[b] is the fp32 accumulator G.  It is initialised by
  // the block's consumer (norm5 / the transition's BatchNorm backward, written WITHOUT the rstd factor in batch mode), every
  // layer's 1x1x1 data-gradient epilogue adds gamma_l * dA1_l, and finalize() turns a channel range into the true gradient
  // rstd (G - C1 - xhat C2) once its lowest reader has run: layers first_layer .. L-1 of block b contribute.
  auto finalize = [&](int b, int c_lo, int nch, int first_layer, bf16* outp, long long out_pitch, const float* out_scale) -> int {
    const BlockInfo& bi = pl->blocks[b];
    const int L = (int)bi.layers.size();
    FinalizeParams f = {};
    f.M = g.M[b]; f.c_lo = c_lo; f.nch = nch;
    f.G = (float*)(ws + g.dbuf[b]); f.g_pitch = bi.ctot;
    f.x = (const bf16*)(ws + g.buf[b]); f.x_pitch = bi.ctot;
    f.bn = make_bn(bi.layers[L - 1].n1, params, buffers, fstats, FC, g.M[b], batch);   // any norm of the block: same statistics
    f.nlayers = 0;
    if (batch && bi.has_trans) {   // the transition's own BatchNorm reads every channel of the block: its (c1, c2) terms are deferred too
      f.gamma[0] = (const float*)params[bi.tn.param_idx];
      f.g_sum[0] = gsum(bi.tn); f.g_dot[0] = gdot(bi.tn);
      f.nlayers = 1;
    }
    for (int l = first_layer; l < L && batch; ++l) {
      const LayerInfo& li = bi.layers[l];
      if (f.nlayers >= 25) return -12;
      f.gamma[f.nlayers] = (const float*)params[li.n1.param_idx];
      f.g_sum[f.nlayers] = gsum(li.n1); f.g_dot[f.nlayers] = gdot(li.n1);
      ++f.nlayers;
    }
    f.inv_count = batch ? 1.0f / (float)g.M[b] : 0.f;
    f.batch = batch ? 1 : 0;
    f.out = outp; f.out_pitch = out_pitch; f.out_scale = out_scale; f.vps = g.D[b] * g.H[b] * g.W[b];
    // the fp32 gradient itself is read again only by the stem's max-pool backward (block 0 input) and by GradCAM (last block)
    f.write_back = (outp == nullptr || b == (int)pl->blocks.size() - 1) ? 1 : 0;
    if (!batch && outp == nullptr) return 0;      // eval mode: the accumulator already holds the gradient
    ProfScope ps_(PC_EXTRACT, st);
    int T = EW_THREADS / nch;
    T = T < 1 ? 1 : (T > 8 ? 8 : T);
    const size_t fsm = (3 * nch + 1) * sizeof(float) + (size_t)2 * T * nch * sizeof(double);
    launch_pdl(grad_finalize_kernel, dim3(ew_grid(f.M * (nch / 8))), dim3(EW_THREADS), fsm, st, f);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
  };

  // ---- tails (BN parameter gradients from the statistics arena, conv2 gradient scratch -> reference layout): the
  // tables are uploaded now, the launches happen per gradient group as soon as the group's last writer is enqueued
  float* gbase = (float*)grads[0];
  std::vector<int> bn_lo(nb + 2, 0), tt_lo(nb + 2, 0);      // table ranges per group (group k = block nb-1-k; nb = stem)
  uint8_t* dtab = ws + g.tables + (1 << 19) + (1 << 17);
  uint8_t* dtt = dtab + (1 << 18);
  {
    // table order: [block nb-1 (+n5)] [block nb-2 + its transition] ... [block 0 + its transition] [stem]
    std::vector<BnTableEntry> tab;
    std::vector<WReduceEntry> tt;
    auto add_bn_entry = [&](const BnInfo& bn) {
      BnTableEntry e = {};
      e.g_sum = gsum(bn); e.g_dot = gdot(bn);
      e.grad_gamma_off = (float*)grads[bn.param_idx] - gbase; e.grad_beta_off = (float*)grads[bn.param_idx + 1] - gbase;
      e.C = bn.C;
      tab.push_back(e);
    };
    auto add_w_entry = [&](int pidx, int taps, int coci) {
      WReduceEntry e;
      e.src = wslots + g.wslot_off[pidx];
      e.dst_off = (float*)grads[pidx] - gbase;
      e.numel = (int)pl->param_numel[pidx]; e.S = g.wslot_S[pidx]; e.taps = taps; e.coci = coci;
      tt.push_back(e);
    };
    for (int k = 0; k < nb; ++k) {
      const BlockInfo& bi = pl->blocks[nb - 1 - k];
      bn_lo[k] = (int)tab.size(); tt_lo[k] = (int)tt.size();
      for (auto& li : bi.layers) {
        add_bn_entry(li.n1); add_bn_entry(li.n2);
        add_w_entry(li.conv2_idx, 27, GROWTH * BOTT);
        if (det) add_w_entry(li.conv1_idx, 0, 0);
      }
      if (bi.has_trans) add_bn_entry(bi.tn); else add_bn_entry(pl->n5);
      if (bi.has_trans && det) add_w_entry(bi.tconv_idx, 0, 0);
    }
    bn_lo[nb] = (int)tab.size(); tt_lo[nb] = (int)tt.size();
    add_bn_entry(pl->n0);
    if (det) add_w_entry(pl->conv0_idx, 0, 0);
    bn_lo[nb + 1] = (int)tab.size(); tt_lo[nb + 1] = (int)tt.size();
    RET_IF(upload_table(pl, 2, dtab, tab.data(), tab.size() * sizeof(BnTableEntry), st));
    RET_IF(upload_table(pl, 3, dtt, tt.data(), tt.size() * sizeof(WReduceEntry), st));
  }
  while ((int)pl->grad_ev.size() < nb + 1) {
    cudaEvent_t e;
    CUDA_RET(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    pl->grad_ev.push_back(e);
  }
  // launches the tails of group k on stream `ts` (which must already be ordered after the group's writers) and
  // records the group's gradient-ready event there
  auto group_tail = [&](int k, cudaStream_t ts) -> int {
    const int nbn = bn_lo[k + 1] - bn_lo[k], ntt = tt_lo[k + 1] - tt_lo[k];
    ProfScope ps_(PC_TAILS, ts, (nbn > 0) + (ntt > 0));
    if (nbn > 0) bn_param_grad_kernel<<<(unsigned)nbn, 128, 0, ts>>>((const BnTableEntry*)dtab + bn_lo[k], gbase);
    if (ntt > 0) wgrad_reduce_kernel<<<dim3(32, (unsigned)ntt), 256, 0, ts>>>((const WReduceEntry*)dtt + tt_lo[k], gbase);
    LAUNCH_RET();
    CUDA_RET(cudaEventRecord(pl->grad_ev[k], ts));
    return 0;
  };

  // ---- norm5
  {
    const BlockInfo& bi = pl->blocks[nb - 1];
    const long long M = g.M[nb - 1];
    const int C = bi.ctot;
    BnSrc bn = make_bn(pl->n5, params, buffers, fstats, FC, M, batch);
    const bf16* x = (const bf16*)(ws + g.buf[nb - 1]);
    const int rows_per_block = EW_THREADS / (C / 8);
    int blocks = (int)std::min<long long>((M + rows_per_block - 1) / rows_per_block, NUM_SMS * 4);
    { ProfScope ps_(PC_NORM5_BWD, st);
      bn_bwd_stats_f32_kernel<<<blocks, EW_THREADS, 2 * C * sizeof(float), st>>>(grad_out, x, C, bn, M, C, gsum(pl->n5), gdot(pl->n5)); }
    LAUNCH_RET();
    RET_IF(bn_apply(BA_OUT_F32_STORE, M, C, nullptr, grad_out, C, x, C, bn, gsum(pl->n5), gdot(pl->n5), ws + g.dbuf[nb - 1], C,
                    nullptr, 0, nullptr, 1, batch ? 1 : 0));
  }

  for (int b = nb - 1; b >= 0; --b) {
    const BlockInfo& bi = pl->blocks[b];
    const long long M = g.M[b];
    const int vps = g.D[b] * g.H[b] * g.W[b];
    bf16* buf = (bf16*)(ws + g.buf[b]);
    float* dbuf = (float*)(ws + g.dbuf[b]);
    bf16* dA1 = (bf16*)(ws + g.dA1);
    for (int l = (int)bi.layers.size() - 1; l >= 0; --l, parity ^= 1) {
      const LayerInfo& li = bi.layers[l];
      bf16* dA2 = (bf16*)(ws + g.dA2) + (size_t)parity * g.maxM_ * BOTT;
      bf16* gslice = (bf16*)(ws + g.gslice) + (size_t)parity * g.maxM_ * GROWTH;
      // the buffers of this parity were last read by the side-stream wgrads of two layers ago
      if (side_done[parity][0]) CUDA_RET(cudaStreamWaitEvent(st, side_done[parity][0], 0));
      if (side_done[parity][1]) CUDA_RET(cudaStreamWaitEvent(st, side_done[parity][1], 0));
      bf16* bott = (bf16*)(ws + g.bott[b]) + (size_t)l * M * BOTT;
      const BnSrc bn1 = make_bn(li.n1, params, buffers, fstats, FC, M, batch);
      const BnSrc bn2 = make_bn(li.n2, params, buffers, fstats, FC, M, batch);
      // gradient of the layer's 32 new channels (all later consumers have already accumulated into dbuf): for the top
      // layer of a block it is extracted here; for every other layer the previous iteration's BN-backward pass already
      // emitted it while it had the final values in registers (fused slice extraction).
      if (l == (int)bi.layers.size() - 1)
        RET_IF(finalize(b, li.cin, GROWTH, (int)bi.layers.size(), gslice, GROWTH, dropmask ? dropmask + (size_t)li.index * B * GROWTH : nullptr));
      // conv2 wgrad -> scratch [tap][co][ci]
      {
        WgradParams w = {};
        w.M = (int)M; w.CB = GROWTH; w.NB = 9; w.na_total = BOTT; w.nb_total = GROWTH;
        w.Dz = g.D[b]; w.Dy = g.H[b]; w.Dx = g.W[b];
        w.a_src = bott; w.a_pitch = BOTT; w.bnA = bn2;
        w.b_src = gslice; w.b_pitch = GROWTH;
        w.dw = wslots + g.wslot_off[li.conv2_idx];
        w.slot_stride = det ? pl->param_numel[li.conv2_idx] : 0;
        w.so_a = 1; w.so_b = BOTT; w.so_j = GROWTH * BOTT;
        cudaEvent_t ready = pl->next_event();
        CUDA_RET(cudaEventRecord(ready, st));          // gslice written
        CUDA_RET(cudaStreamWaitEvent(sd, ready, 0));
        {
          ProfScope ps_(PC_CONV2_WGRAD, sd);
          RET_IF(launch_wgrad(w, 1, det ? g.wslot_S[li.conv2_idx] : 0, sd));
        }
        side_done[parity][0] = pl->next_event();
        CUDA_RET(cudaEventRecord(side_done[parity][0], sd));
      }
      // conv2 dgrad (+ ReLU mask of norm2/relu2, + BN2 backward statistics)
      if (true) {  // brick mode for every spatial size: partial tiles only cost idle MMA rows, tiny layers are latency-bound anyway
        BrickParams q = {};
        q.B = B; q.Dz = g.D[b]; q.Dy = g.H[b]; q.Dx = g.W[b]; q.CH = GROWTH; q.NT = BOTT; q.tap_sign = -1;
        q.a_src = gslice; q.a_pitch = GROWTH;
        q.b_packed = packed + li.pk_c2d;
        q.out = dA2; q.out_pitch = BOTT;
        q.st_sum = gsum(li.n2); q.st_sq = gdot(li.n2);
        q.e_src = bott; q.e_pitch = BOTT; q.bnE = bn2;
        ProfScope ps_(PC_CONV2_DGRAD, st);
        RET_IF(launch_brick(q, 1, st));
      } else {
        RowsParams p = {};
        p.M = (int)M; p.NT = 128; p.Ncols = BOTT; p.Cin = GROWTH; p.kbw = 32; p.ntaps = 27; p.tap_sign = -1;
        p.Dz = g.D[b]; p.Dy = g.H[b]; p.Dx = g.W[b];
        p.a_src = gslice; p.a_pitch = GROWTH;
        p.b_packed = packed + li.pk_c2d;
        p.out = dA2; p.out_pitch = BOTT;
        p.st_sum = gsum(li.n2); p.st_sq = gdot(li.n2);
        p.e_src = bott; p.e_pitch = BOTT; p.bnE = bn2;
        ProfScope ps_(PC_CONV2_DGRAD, st);
        RET_IF(launch_rows(p, A_LINEAR_CONV, T_NONE, EP_MASK_STATS, 1, st));
      }
      // NEGATIVE RESULT (round 2, same-box A/B at configs[1]): applying norm2's backward in the producers of the late-block 1x1x1
      // data-gradient GEMM (T_BNBWD) removes 40 bn_bwd_apply launches (class 0.96 -> 0.63 ms) but lengthens the GEMMs' producer chain
      // (2.19 -> 2.38 ms) and delays the side-stream weight gradient: 1220 vs 1233 volumes/s.  OFF unless MMNN_FUSE_BN2=1.
      static const bool fuse_env = [] { const char* e = getenv("MMNN_FUSE_BN2"); return e != nullptr && e[0] == '1'; }();
      const bool fuse_bn2 = fuse_env && M <= FUSE_BN2_MAX_M;
      bf16* dB2 = (bf16*)(ws + g.dB2) + (size_t)parity * FUSE_BN2_MAX_M * BOTT;
      const bf16* dbott = fuse_bn2 ? dB2 : dA2;      // BN2-backward output = gradient of the bottleneck tensor
      if (!fuse_bn2) RET_IF(bn_apply(BA_OUT_BF16, M, BOTT, dA2, nullptr, BOTT, bott, BOTT, bn2, gsum(li.n2), gdot(li.n2), dA2, BOTT));
      auto conv1_wgrad = [&]() -> int {              // D[ci][co] -> dW1[co][ci], on the side stream
        WgradParams w = {};
        w.M = (int)M; w.CB = 128; w.NB = 1; w.na_total = li.cin; w.nb_total = BOTT;
        w.Dz = g.D[b]; w.Dy = g.H[b]; w.Dx = g.W[b];
        w.a_src = buf; w.a_pitch = bi.ctot; w.bnA = bn1;
        w.b_src = dbott; w.b_pitch = BOTT;
        w.dw = det ? wslots + g.wslot_off[li.conv1_idx] : (float*)grads[li.conv1_idx];
        w.slot_stride = det ? pl->param_numel[li.conv1_idx] : 0;
        w.so_a = 1; w.so_b = li.cin; w.so_j = 0;
        cudaEvent_t ready = pl->next_event();
        CUDA_RET(cudaEventRecord(ready, st));          // dBott is final
        CUDA_RET(cudaStreamWaitEvent(sd, ready, 0));
        {
          ProfScope ps_(PC_CONV1_WGRAD, sd);
          RET_IF(launch_wgrad(w, 0, det ? g.wslot_S[li.conv1_idx] : 0, sd));
        }
        side_done[parity][1] = pl->next_event();
        CUDA_RET(cudaEventRecord(side_done[parity][1], sd));
        return 0;
      };
      if (!fuse_bn2) RET_IF(conv1_wgrad());
      // conv1 dgrad (+ ReLU mask of norm1/relu1, + BN1 backward statistics): the masked gradient, scaled by gamma, is ADDED to the
      // block's fp32 accumulator by the epilogue (deferred BatchNorm backward: no bf16 dA1 tensor, no per-layer pass over cin channels)
      {
        RowsParams p = {};
        p.M = (int)M; p.NT = li.nt_d; p.Ncols = li.cin; p.Cin = BOTT; p.kbw = 64; p.ntaps = 1; p.tap_sign = 1;
        p.Dz = g.D[b]; p.Dy = g.H[b]; p.Dx = g.W[b];
        p.a_src = dA2; p.a_pitch = BOTT;
        p.b_packed = packed + li.pk_c1d;
        p.out = (bf16*)dbuf; p.out_pitch = bi.ctot; p.acc_rstd = batch ? 0 : 1;
        p.st_sum = gsum(li.n1); p.st_sq = gdot(li.n1);
        p.e_src = buf; p.e_pitch = bi.ctot; p.bnE = bn1;
        if (fuse_bn2) {     // the producers apply BN2's backward to the raw masked gradient and materialise it for the weight gradient
          p.bnA = bn2; p.t_src = bott; p.t_pitch = BOTT; p.t_gsum = gsum(li.n2); p.t_gdot = gdot(li.n2);
          p.t_inv_count = batch ? 1.0f / (float)M : 0.f; p.t_out = dB2; p.t_out_pitch = BOTT;
        }
        ProfScope ps_(PC_CONV1_DGRAD, st);
        RET_IF(launch_rows(p, A_LINEAR_CONV, fuse_bn2 ? T_BNBWD : T_NONE, EP_MASK_STATS_ACC, 1, st));
      }
      if (fuse_bn2) RET_IF(conv1_wgrad());
      if (l > 0) {
        // channels [cin-32, cin) = the new channels of layer l-1 have seen their last reader: finalise them and emit that
        // layer's gradient slice into the other parity's buffer (after the side-stream wgrad that last read it has finished)
        const LayerInfo& lp = bi.layers[l - 1];
        if (side_done[parity ^ 1][0]) CUDA_RET(cudaStreamWaitEvent(st, side_done[parity ^ 1][0], 0));
        bf16* gnext = (bf16*)(ws + g.gslice) + (size_t)(parity ^ 1) * g.maxM_ * GROWTH;
        RET_IF(finalize(b, lp.cin, GROWTH, l, gnext, GROWTH, dropmask ? dropmask + (size_t)lp.index * B * GROWTH : nullptr));
      }
    }
    {
      // every gradient of block b (and of the transition after it, done in the previous iteration) has its last writer
      // enqueued: statistics on this stream, weight gradients on the side stream.  Finish the group on the side stream.
      cudaEvent_t m = pl->next_event();
      CUDA_RET(cudaEventRecord(m, st));
      CUDA_RET(cudaStreamWaitEvent(sd, m, 0));
      RET_IF(group_tail(nb - 1 - b, sd));
    }
    if (b > 0) {
      // transition b-1: buf[b][:, :c0] = avgpool(conv(relu(bn(buf[b-1]))))  ==  conv(pooled[b-1])
      const BlockInfo& pv = pl->blocks[b - 1];
      const long long Mp = g.M[b - 1];
      bf16* gout = (bf16*)(ws + g.gout);
      bf16* dpooled = (bf16*)(ws + g.dpooled);
      const bf16* pooled = (const bf16*)(ws + g.pooled[b - 1]);
      if (trans_done) CUDA_RET(cudaStreamWaitEvent(st, trans_done, 0));   // previous transition's wgrad still reads gout
      RET_IF(finalize(b, 0, bi.c0, 0, gout, bi.c0, nullptr));     // the block's input channels: every layer of the block read them
      {
        WgradParams w = {};
        w.M = (int)M; w.CB = 128; w.NB = 1; w.na_total = pv.ctot; w.nb_total = bi.c0;
        w.Dz = g.D[b]; w.Dy = g.H[b]; w.Dx = g.W[b];
        w.a_src = pooled; w.a_pitch = pv.ctot;
        w.b_src = gout; w.b_pitch = bi.c0;
        w.dw = det ? wslots + g.wslot_off[pv.tconv_idx] : (float*)grads[pv.tconv_idx];
        w.slot_stride = det ? pl->param_numel[pv.tconv_idx] : 0;
        w.so_a = 1; w.so_b = pv.ctot; w.so_j = 0;
        cudaEvent_t ready = pl->next_event();
        CUDA_RET(cudaEventRecord(ready, st));
        CUDA_RET(cudaStreamWaitEvent(sd, ready, 0));
        {
          ProfScope ps_(PC_TRANS_WGRAD, sd);
          RET_IF(launch_wgrad(w, 2, det ? g.wslot_S[pv.tconv_idx] : 0, sd));
        }
        trans_done = pl->next_event();
        CUDA_RET(cudaEventRecord(trans_done, sd));
      }
      {
        RowsParams p = {};
        p.M = (int)M; p.NT = 128; p.Ncols = pv.ctot; p.Cin = bi.c0; p.kbw = 64; p.ntaps = 1; p.tap_sign = 1;
        p.Dz = g.D[b]; p.Dy = g.H[b]; p.Dx = g.W[b];
        p.a_src = gout; p.a_pitch = bi.c0;
        p.b_packed = packed + pv.pk_td;
        p.out = dpooled; p.out_pitch = pv.ctot;
        ProfScope ps_(PC_TRANS_DGRAD, st);
        RET_IF(launch_rows(p, A_LINEAR_CONV, T_NONE, EP_STORE, 1, st));
      }
      AvgPoolParams a = {};
      a.B = B; a.D = g.D[b - 1]; a.H = g.H[b - 1]; a.W = g.W[b - 1]; a.C = pv.ctot;
      a.x = (const bf16*)(ws + g.buf[b - 1]); a.x_pitch = pv.ctot;
      a.bn = make_bn(pv.tn, params, buffers, fstats, FC, Mp, batch);
      a.dpooled = dpooled;
      a.dx = (float*)(ws + g.dbuf[b - 1]); a.dx_pitch = pv.ctot;
      a.g_sum = gsum(pv.tn); a.g_dot = gdot(pv.tn); a.g_sum_in = gsum(pv.tn); a.g_dot_in = gdot(pv.tn);
      a.inv_count = batch ? 1.0f / (float)Mp : 0.f;
      a.pre_rstd = batch ? 1 : 0;                                  // dbuf[b-1] is that block's accumulator G (finalised per channel range)
      const int rows_per_block = EW_THREADS / (pv.ctot / 8);
      int blocks = (int)std::min<long long>((Mp + rows_per_block - 1) / rows_per_block, NUM_SMS * 8);
      { ProfScope ps_(PC_AVGPOOL_BWD, st, 1);     // ONE pass: statistics + gamma * v into the block's accumulator (c1 / c2 deferred to finalize)
        launch_pdl(avgpool_bnrelu_bwd_kernel<3>, dim3(blocks), dim3(EW_THREADS), 7 * pv.ctot * sizeof(float), st, a); }
      LAUNCH_RET();
    } else {
      // pool0 + relu0 + norm0 + conv0
      RET_IF(finalize(0, 0, bi.c0, 0, nullptr, 0, nullptr));      // block 1's input channels (the pooled stem output), in place
      PoolBwdParams q = {};
      q.B = B; q.D0 = g.D0; q.H0 = g.H0; q.W0 = g.W0; q.D1 = g.D[0]; q.H1 = g.H[0]; q.W1 = g.W[0];
      q.x = (const bf16*)(ws + g.stem_out);
      q.bn = make_bn(pl->n0, params, buffers, fstats, FC, g.M0, batch);
      q.dpool = dbuf; q.dpool_pitch = bi.ctot;
      q.argmax = ws + g.argmax;
      q.dr = (bf16*)(ws + g.dr);
      q.g_sum = gsum(pl->n0); q.g_dot = gdot(pl->n0);
      static const bool mpb_gather = [] { const char* e = getenv("MMNN_MAXPOOL_BWD_GATHER"); return e != nullptr && e[0] == '1'; }();
      if (mpb_gather) {   // round-1 gather form (kept for A/B and as the reference of the parity test)
        int blocks = (int)std::min<long long>((g.M0 + 31) / 32, NUM_SMS * 8);
        ProfScope ps_(PC_MAXPOOL_BWD, st); launch_pdl(maxpool_bnrelu_bwd_kernel, dim3(blocks), dim3(EW_THREADS), 0, st, q);
      } else {
        const long long tiles = (long long)B * ((g.D0 + MPB_TZ - 1) / MPB_TZ) * ((g.H0 + MPB_TY - 1) / MPB_TY) * ((g.W0 + MPB_TX - 1) / MPB_TX);
        const int blocks = (int)std::min<long long>(tiles, NUM_SMS * 2);
        CUDA_RET(cudaFuncSetAttribute(maxpool_bnrelu_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MPB_SMEM));
        ProfScope ps_(PC_MAXPOOL_BWD, st); launch_pdl(maxpool_bnrelu_bwd_tiled_kernel, dim3(blocks), dim3(EW_THREADS), MPB_SMEM, st, q);
      }
      LAUNCH_RET();
      RET_IF(bn_apply(BA_OUT_BF16, g.M0, 64, q.dr, nullptr, 64, q.x, 64, q.bn, gsum(pl->n0), gdot(pl->n0), q.dr, 64));
      WgradParams w = {};
      w.M = (int)g.M0; w.CB = 64; w.NB = 1; w.na_total = 128; w.nb_total = 64;
      w.Dz = g.D0; w.Dy = g.H0; w.Dx = g.W0; w.Sz = g.Sz; w.Sy = g.Sy; w.Sx = g.Sx;
      // after a training-mode forward the workspace holds the bf16 copy (TMA operand); after an eval-mode forward (GradCAM) only
      // the activation-format image, read by the register path
      w.a_bf16 = (!kActF16 || batch) ? 1 : 0;
      w.a_src = (const bf16*)(ws + (w.a_bf16 ? g.xs2dw : g.xs2d)); w.a_pitch = 16;
      w.b_src = q.dr; w.b_pitch = 64;
      w.dw = det ? wslots + g.wslot_off[pl->conv0_idx] : (float*)grads[pl->conv0_idx];
      w.slot_stride = det ? pl->param_numel[pl->conv0_idx] : 0;
      w.cin_real = pl->cin_real;
      ProfScope ps_(PC_STEM_WGRAD, st);
      RET_IF(launch_wgrad(w, 3, det ? g.wslot_S[pl->conv0_idx] : 0, st));
    }
  }

  {  // join: everything queued on the side stream becomes a dependency of the caller's stream
    cudaEvent_t e = pl->next_event();
    CUDA_RET(cudaEventRecord(e, sd));
    CUDA_RET(cudaStreamWaitEvent(st, e, 0));
  }
  // ---- tail of the stem group (conv0, norm0), then its event
  RET_IF(group_tail(nb, st));
  return 0;
}

}  // extern "C"
