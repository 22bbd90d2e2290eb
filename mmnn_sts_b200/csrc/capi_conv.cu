// C-ABI entry points for the tcgen05 tile engine (fine-grained: used by the per-kernel parity tests and by the
// encoder orchestrator in encoder.cu).  Plain pointers and sizes only; see include/mmnn_b200.h.
#include "engine.cuh"
#include "pack.cuh"

using namespace mmnn;

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

template <int AMODE, int TRANS, int EPI>
int launch_rows_t(RowsParams p, cudaStream_t stream) {
  uint32_t offs[6];
  if (p.stages <= 0) p.stages = choose_stages(p, 100 * 1024);
  const uint32_t smem = rows_smem_layout(p.Cin, p.NT, p.kbw, p.stages, offs);
  auto kern = conv_rows_kernel<AMODE, TRANS, EPI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((p.M + TILE_ROWS - 1) / TILE_ROWS, (p.Ncols + p.NT - 1) / p.NT);
  kern<<<grid, ENGINE_THREADS, smem, stream>>>(p);
  MMNN_CHECK_LAUNCH();
  return 0;
}

int launch_rows(const RowsParams& p, int amode, int trans, int epi, cudaStream_t stream) {
  if (p.NT % 32 != 0 || p.NT > 256 || p.Cin % 32 != 0 || (p.kbw != 32 && p.kbw != 64)) return -2;
#define CASE(A, T, E) if (amode == A && trans == T && epi == E) return launch_rows_t<A, T, E>(p, stream);
  CASE(A_LINEAR_CONV, T_NONE, EP_STORE)
  CASE(A_LINEAR_CONV, T_NONE, EP_STORE_STATS)
  CASE(A_LINEAR_CONV, T_NONE, EP_MASK_STATS)
  CASE(A_LINEAR_CONV, T_BNRELU, EP_STORE_STATS)
  CASE(A_LINEAR_CONV, T_BNRELU, EP_STORE)
  CASE(A_STEM, T_NONE, EP_STORE_STATS)
  CASE(A_STEM, T_NONE, EP_STORE)
#undef CASE
  return -3;
}

}  // namespace mmnn

extern "C" {

int mmnn_conv_rows(const RowsParams* p, int amode, int trans, int epi, void* stream) {
  return launch_rows(*p, amode, trans, epi, (cudaStream_t)stream);
}

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
