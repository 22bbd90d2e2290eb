// Host-side launch helper: cudaLaunchKernelEx with the programmatic-stream-serialization attribute (PDL), see common.cuh
// pdl_wait / pdl_trigger.  ON by default since round 2 (MMNN_PDL=0 disables it): with every kernel triggering right AFTER its wait the
// next kernel's launch latency and pre-wait prologue (barrier init, TMEM allocation, weight prefetch) overlap the running kernel's
// body -- same-box A/B on B200 at configs[1]: 1241 vs 1220 volumes/s device-resident, 1232 vs 1219 end to end, parity and
// bit-reproducibility tests green under both settings.  (Round 1 triggered at kernel START and measured no gain: 15.33 vs 15.35 ms.)
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace mmnn {

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MMNN_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// Launch WITH the attribute regardless of MMNN_PDL: for kernels that are written to do useful work before griddepcontrol.wait
// (the early-start 1x1x1 GEMMs of the late dense blocks, engine.cuh).  The preceding kernel in the stream must itself have been
// launched WITHOUT the attribute (so that everything before it is complete when the dependent starts).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_forced(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace mmnn
