// Host-side launch helper: cudaLaunchKernelEx, optionally with the programmatic-stream-serialization attribute (PDL),
// see common.cuh pdl_trigger / pdl_wait.  OFF by default: measured on B200 at configs[1] (round 1) the step is bound by
// kernel time, not by launch gaps -- 15.33 ms with PDL vs 15.35 ms without, and the end-to-end arm (H2D prefetch on a
// copy stream) got slower -- so plain stream-ordered launches stay the default; MMNN_PDL=1 enables it for experiments.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace mmnn {

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MMNN_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
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
