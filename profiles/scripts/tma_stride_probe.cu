// Probe: which global strides does a cp.async.bulk.tensor (tile mode, SWIZZLE_64B, bf16) load accept on sm_100a?
// usage: tma_stride_probe <rank> <s0 bytes> <s1 bytes> [swizzle 0|2(64B)]     dims {64, 16, 12, (2, 2)}, box {32, 16, 8, 1, 1}
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride_probe tma_stride_probe.cu   (no -lcuda: driver entry point)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int rank, int c1, int c2, unsigned short* out, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned sb = (unsigned)__cvta_generic_to_shared(&bar);
  const unsigned dst = ((unsigned)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(8192u) : "memory");
    if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                   "l"(&tm), "r"(sb), "r"(0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                   "l"(&tm), "r"(sb), "r"(0), "r"(c1), "r"(c2), "r"(0), "r"(0) : "memory");
    long long t0 = clock64();
    unsigned ok = 0;
    while (!ok && clock64() - t0 < 200000000LL) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(sb), "r"(0u) : "memory");
    }
    *status = (int)ok;
    if (ok) {
      const unsigned short* s = reinterpret_cast<const unsigned short*>(smem + (dst - (unsigned)__cvta_generic_to_shared(smem)));
      for (int i = 0; i < 4096; ++i) out[i] = s[i];
    }
  }
}

int main(int argc, char** argv) {
  const int rank = argc > 1 ? atoi(argv[1]) : 3;
  const unsigned long long s0 = argc > 2 ? atoll(argv[2]) : 128, s1 = argc > 3 ? atoll(argv[3]) : 2048;
  const int swz = argc > 4 ? atoi(argv[4]) : 2;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) { printf("no entry point\n"); return 1; }
  const size_t elems = 1 << 22;
  std::vector<unsigned short> h(elems);
  for (size_t i = 0; i < elems; ++i) h[i] = (unsigned short)(i & 0xffff);
  unsigned short *d, *dout;
  int* dst;
  cudaMalloc(&d, elems * 2); cudaMalloc(&dout, 8192); cudaMalloc(&dst, 4);
  cudaMemcpy(d, h.data(), elems * 2, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0xff, 4);
  cuuint64_t gdim[5] = {64, 16, 12, 2, 2};
  cuuint64_t gstr[4] = {s0, s1, s1 * 16, s1 * 32};
  if (argc > 5 && atoi(argv[5]) == 1) {   // the stem weight gradient's test case: image [3][11][11][19][16]
    gdim[2] = 11; gdim[3] = 11; gdim[4] = 3;
    gstr[0] = 32; gstr[1] = 608; gstr[2] = 6688; gstr[3] = 73568;
  }
  const cuuint32_t box[5] = {32, 16, 8, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap tm;
  const CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rank %d s0 %llu s1 %llu swizzle %d: encode -> %d", rank, s0, s1, swz, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe<<<1, 32, 16384>>>(tm, rank, 1, 2, dout, dst);
  const cudaError_t e = cudaDeviceSynchronize();
  int st = -1;
  cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost);
  std::vector<unsigned short> o(4096);
  cudaMemcpy(o.data(), dout, 8192, cudaMemcpyDeviceToHost);
  // expected first element of the box: coordinates (0, 1, 2): byte offset s0*1 + s1*2
  printf("  sync %s  completed %d  first elems %u %u (expect %llu)  row1 %u\n", cudaGetErrorString(e), st, o[0], o[1], ((s0 + 2 * s1) / 2) & 0xffff, o[32]);
  return 0;
}
