// Multi-tensor SGD (momentum, Nesterov, weight decay) in ONE launch: the optimiser step of the reference's training
// loop (torch.optim.SGD(lr, momentum=0.9, nesterov=True, weight_decay=1e-4), /root/reference/main.py:410-414,479-481)
// over all ~400 parameter tensors.  HBM-bound: 20 B per parameter (read p, g, m; write p, m); torch's foreach path
// takes 3 passes x 11 launches for the same update.
#include <cuda_runtime.h>
#include <stdint.h>

#include "prof.h"

namespace mmnn {

struct SgdTensor { float* p; const float* g; float* m; int n; int first_chunk; };
constexpr int SGD_MAX_TENSORS = 512;
struct SgdTable { SgdTensor t[SGD_MAX_TENSORS]; };   // 16 KB, passed BY VALUE as the kernel parameter: the pointers are
                                                     // captured at launch (gradient tensors move between steps), no
                                                     // device-side table to keep alive, CUDA-graph safe
constexpr int SGD_CHUNK = 8192;      // elements per block
constexpr int SGD_THREADS = 256;

// g' = g + wd p;  m = mu m + g';  step = nesterov ? g' + mu m : m;  p -= lr step   (dampening 0, torch semantics)
__device__ __forceinline__ void sgd_one(float& p, float g, float& m, float lr, float mu, float wd, int nesterov) {
  g = fmaf(wd, p, g);
  m = fmaf(mu, m, g);
  const float s = nesterov ? fmaf(mu, m, g) : m;
  p = fmaf(-lr, s, p);
}

// hyper != nullptr: lr / momentum / weight decay are read from DEVICE memory ({lr, mu, wd}) at run time, so a step captured
// in a CUDA graph follows a schedule (OneCycleLR cycles lr AND momentum every optimiser step, /root/reference/main.py:414)
// that the host updates in place between replays; otherwise the launch-time scalars are used.
__global__ void __launch_bounds__(SGD_THREADS) sgd_step_kernel(const __grid_constant__ SgdTable tab, int ntensors,
                                                               float lr, float mu, float wd, int nesterov,
                                                               const float* __restrict__ hyper) {
  if (hyper != nullptr) { lr = hyper[0]; mu = hyper[1]; wd = hyper[2]; }
  // the tensor owning this block's chunk: last entry with first_chunk <= blockIdx.x
  int lo_t = 0, hi_t = ntensors - 1;
  while (lo_t < hi_t) {
    const int mid = (lo_t + hi_t + 1) >> 1;
    if (tab.t[mid].first_chunk <= (int)blockIdx.x) lo_t = mid; else hi_t = mid - 1;
  }
  const SgdTensor t = tab.t[lo_t];
  const long long lo = (long long)((int)blockIdx.x - t.first_chunk) * SGD_CHUNK;
  const long long hi = (lo + SGD_CHUNK < t.n) ? lo + SGD_CHUNK : t.n;
  const bool vec = ((((uintptr_t)t.p) | ((uintptr_t)t.g) | ((uintptr_t)t.m)) & 15u) == 0;
  if (vec) {
    const long long nv = (hi - lo) >> 2;
    float4* p4 = reinterpret_cast<float4*>(t.p + lo);
    const float4* g4 = reinterpret_cast<const float4*>(t.g + lo);
    float4* m4 = reinterpret_cast<float4*>(t.m + lo);
    for (long long i = threadIdx.x; i < nv; i += SGD_THREADS) {
      float4 p = p4[i], m = m4[i];
      const float4 g = g4[i];
      sgd_one(p.x, g.x, m.x, lr, mu, wd, nesterov); sgd_one(p.y, g.y, m.y, lr, mu, wd, nesterov);
      sgd_one(p.z, g.z, m.z, lr, mu, wd, nesterov); sgd_one(p.w, g.w, m.w, lr, mu, wd, nesterov);
      p4[i] = p; m4[i] = m;
    }
    for (long long i = lo + (nv << 2) + threadIdx.x; i < hi; i += SGD_THREADS) {
      float p = t.p[i], m = t.m[i];
      sgd_one(p, t.g[i], m, lr, mu, wd, nesterov);
      t.p[i] = p; t.m[i] = m;
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += SGD_THREADS) {
      float p = t.p[i], m = t.m[i];
      sgd_one(p, t.g[i], m, lr, mu, wd, nesterov);
      t.p[i] = p; t.m[i] = m;
    }
  }
}

}  // namespace mmnn

extern "C" {
int mmnn_sgd_chunk_elems() { return mmnn::SGD_CHUNK; }
int mmnn_sgd_max_tensors() { return mmnn::SGD_MAX_TENSORS; }
// p / g / m: HOST arrays of device pointers (parameter, gradient, momentum buffer), n: HOST array of element counts
static int sgd_step_impl(void* const* p, const void* const* g, void* const* m, const long long* n, int ntensors, float lr,
                         float momentum, float weight_decay, int nesterov, const float* hyper, void* stream) {
  using namespace mmnn;
  for (int base = 0; base < ntensors; base += SGD_MAX_TENSORS) {
    const int cnt = ntensors - base < SGD_MAX_TENSORS ? ntensors - base : SGD_MAX_TENSORS;
    static thread_local SgdTable tab;
    int chunks = 0;
    for (int i = 0; i < cnt; ++i) {
      if (n[base + i] <= 0 || n[base + i] > 0x7fffffffLL) return -2;
      tab.t[i].p = (float*)p[base + i]; tab.t[i].g = (const float*)g[base + i]; tab.t[i].m = (float*)m[base + i];
      tab.t[i].n = (int)n[base + i]; tab.t[i].first_chunk = chunks;
      chunks += (int)((n[base + i] + SGD_CHUNK - 1) / SGD_CHUNK);
    }
    ProfScope ps(PC_SGD, (cudaStream_t)stream);
    sgd_step_kernel<<<chunks, SGD_THREADS, 0, (cudaStream_t)stream>>>(tab, cnt, lr, momentum, weight_decay, nesterov, hyper);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}
int mmnn_sgd_step(void* const* p, const void* const* g, void* const* m, const long long* n, int ntensors, float lr,
                  float momentum, float weight_decay, int nesterov, void* stream) {
  return sgd_step_impl(p, g, m, n, ntensors, lr, momentum, weight_decay, nesterov, nullptr, stream);
}
// hyper: DEVICE float[3] = {lr, momentum, weight_decay}, read by the kernel when it RUNS (CUDA-graph replays included)
int mmnn_sgd_step_dev(void* const* p, const void* const* g, void* const* m, const long long* n, int ntensors,
                      const float* hyper, int nesterov, void* stream) {
  if (hyper == nullptr) return -2;
  return sgd_step_impl(p, g, m, n, ntensors, 0.f, 0.f, 0.f, nesterov, hyper, stream);
}
}
