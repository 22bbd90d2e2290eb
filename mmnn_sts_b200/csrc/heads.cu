// Small latency-bound kernels after the trunk: image feature head (ReLU -> global average pool -> Linear -> Dropout),
// clinical MLP + fusion heads, Cox partial-likelihood loss, bootstrap concordance index.
// References: /root/reference/models/densenet.py:234-247, /root/reference/models/mlp.py:19-51,
// /root/reference/models/multimodal.py:51-80, /root/reference/losses/losses.py:6-9 (pycox CoxPHLoss),
// /root/reference/main.py:106-123 (lifelines concordance_index).
#include "common.cuh"

using namespace mmnn;

#define LAUNCH_RET()                          \
  do {                                        \
    cudaError_t e_ = cudaGetLastError();      \
    if (e_ != cudaSuccess) return (int)e_;    \
  } while (0)

namespace {

// ---------------------------------------------------------------------------------------------- feature head
// y: fp32 [B][V][C] (norm5 output, NDHWC);  pooled[b][c] = mean_v relu(y);  out[b][f] = (W[f]·pooled[b] + bias[f]) * mask[b][f]
__global__ void __launch_bounds__(256) gap_linear_fwd_kernel(const float* __restrict__ y, int V, int C, const float* __restrict__ W,
                                                             const float* __restrict__ bias, const float* __restrict__ mask,
                                                             int F, float* __restrict__ pooled, float* __restrict__ out) {
  extern __shared__ float sp[];  // [C]
  const int b = blockIdx.x;
  const float* yb = y + (size_t)b * V * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int v = 0; v < V; ++v) acc += fmaxf(yb[(size_t)v * C + c], 0.f);
    acc /= (float)V;
    sp[c] = acc;
    pooled[(size_t)b * C + c] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int f = warp; f < F; f += blockDim.x >> 5) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(W[(size_t)f * C + c], sp[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float r = acc + bias[f];
      if (mask != nullptr) r *= mask[(size_t)b * F + f];
      out[(size_t)b * F + f] = r;
    }
  }
}

// dy[b][v][c] = relu'(y) * (1/V) * sum_f g[b][f] W[f][c],   g = dout * mask
__global__ void __launch_bounds__(256) gap_linear_bwd_dy_kernel(const float* __restrict__ y, int V, int C, const float* __restrict__ W,
                                                                const float* __restrict__ dout, const float* __restrict__ mask,
                                                                int F, float* __restrict__ dy) {
  extern __shared__ float sg[];  // [F]
  const int b = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += blockDim.x)
    sg[f] = dout[(size_t)b * F + f] * (mask != nullptr ? mask[(size_t)b * F + f] : 1.f);
  __syncthreads();
  const float* yb = y + (size_t)b * V * C;
  float* dyb = dy + (size_t)b * V * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc = fmaf(sg[f], W[(size_t)f * C + c], acc);
    acc /= (float)V;
    for (int v = 0; v < V; ++v) dyb[(size_t)v * C + c] = yb[(size_t)v * C + c] > 0.f ? acc : 0.f;
  }
}

// dW[f][c] = sum_b g[b][f] pooled[b][c];  db[f] = sum_b g[b][f]
__global__ void __launch_bounds__(256) gap_linear_bwd_w_kernel(const float* __restrict__ pooled, const float* __restrict__ dout,
                                                               const float* __restrict__ mask, int B, int C, int F,
                                                               float* __restrict__ dW, float* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < F * C) {
    const int f = idx / C, c = idx - f * C;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const float g = dout[(size_t)b * F + f] * (mask != nullptr ? mask[(size_t)b * F + f] : 1.f);
      acc = fmaf(g, pooled[(size_t)b * C + c], acc);
    }
    dW[idx] = acc;
  }
  if (idx < F) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dout[(size_t)b * F + idx] * (mask != nullptr ? mask[(size_t)b * F + idx] : 1.f);
    db[idx] = acc;
  }
}

}  // namespace

extern "C" {

int mmnn_gap_linear_fwd(const float* y, int B, int V, int C, const float* W, const float* bias, const float* mask, int F,
                        float* pooled, float* out, void* stream) {
  gap_linear_fwd_kernel<<<B, 256, C * sizeof(float), (cudaStream_t)stream>>>(y, V, C, W, bias, mask, F, pooled, out);
  LAUNCH_RET();
  return 0;
}

int mmnn_gap_linear_bwd(const float* y, const float* pooled, int B, int V, int C, const float* W, const float* dout,
                        const float* mask, int F, float* dy, float* dW, float* db, void* stream) {
  gap_linear_bwd_dy_kernel<<<B, 256, F * sizeof(float), (cudaStream_t)stream>>>(y, V, C, W, dout, mask, F, dy);
  LAUNCH_RET();
  gap_linear_bwd_w_kernel<<<(F * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pooled, dout, mask, B, C, F, dW, db);
  LAUNCH_RET();
  return 0;
}

}  // extern "C"
