// Shared device helpers for the sm_100a kernels: PTX wrappers (mbarrier, bulk copy, tcgen05/TMEM), bf16 packing.
// Everything here is written for compute_100a only (tcgen05.* does not exist elsewhere).
#pragma once
#include <cuda.h>        // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mmnn {

typedef __nv_bfloat16 bf16;   // storage type of every 16-bit tensor pointer; the interpretation (bf16 or fp16) is per tensor

// Forward tensors (activations + forward weight images) are stored as IEEE fp16, gradient tensors as bf16:
// same tensor-core rate (tcgen05 kind::f16 takes either), but fp16's 3 extra mantissa bits cut the forward rounding
// noise 8x (DESIGN.md "Numerics"); activations are BatchNorm-bounded so fp16's range is safe (stores saturate),
// while gradients keep bf16's fp32-like range.  -DMMNN_ACT_FP16=0 switches activations back to bf16.
#ifndef MMNN_ACT_FP16
#define MMNN_ACT_FP16 1
#endif
constexpr bool kActF16 = MMNN_ACT_FP16 != 0;

#define MMNN_DEVINL __device__ __forceinline__

MMNN_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
MMNN_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
MMNN_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MMNN_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
MMNN_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
MMNN_DEVINL uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
MMNN_DEVINL uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error), never as a hung GPU.
MMNN_DEVINL void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > 2000000000ull) {
      printf("mmnn: mbarrier wait timeout tag=%d block=(%d,%d,%d) thread=%d parity=%u\n", tag, blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- async proxy / bulk copy (TMA unit, 1-D form)
MMNN_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MMNN_DEVINL void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Asynchronous L2 prefetch of a contiguous global range (TMA unit): warms L2 for bulk copies issued later in the kernel.
MMNN_DEVINL void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// TMA tensor copy global -> shared (cp.async.bulk.tensor, SASS UTMALDG): one box of a 5-D tensor map; coordinates may lie
// outside the tensor -- those elements arrive as ZEROS, which is exactly the zero padding of a convolution operand that needs
// no transform (gradient tensors).  Completion is signalled on `bar` with the full box byte count.
MMNN_DEVINL void tma_load_5d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst_smem),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
MMNN_DEVINL void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
               "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
MMNN_DEVINL void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// One lane of a fully converged warp, chosen by the hardware (elect.sync).  Guarding the tcgen05.mma / commit block
// with THIS predicate (instead of `lane == 0`) lets ptxas emit the uniform-datapath instructions straight-line; with
// an ordinary divergent predicate every UTCHMMA is wrapped in an ELECT / PLOP3 / BRA.U.ANY loop (~130 cycles each,
// measured with ncu source counters: profiles/r01_brick_fprop_sass_mma_loop.txt).
MMNN_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel launched through launch_pdl() (launch.h) calls pdl_wait() before its first access to global memory that the
// preceding kernel may write (returns when that kernel has completed and its writes are visible) and pdl_trigger() right AFTER
// it: the next kernel in the stream may then be scheduled, so its launch latency and pre-wait prologue (barrier init, TMEM
// allocation, weight prefetch -- or, for the early-start 1x1x1 GEMMs, most of the K loop) overlap this kernel's body.
// Trigger-after-wait makes the overlap transitive-safe: when a dependent starts, every kernel BEFORE its predecessor is complete.
// Without the launch attribute (the default, see launch.h) both are no-ops.
MMNN_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
MMNN_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
MMNN_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
MMNN_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
MMNN_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MMNN_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
MMNN_DEVINL void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
MMNN_DEVINL void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t), columns [col, col+32)
MMNN_DEVINL void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" canonical layout): the unit is a core matrix of
// 8 rows x 16 bytes stored contiguously (128 B).
//   K-major operand : LBO = byte stride between the two 16-byte K chunks of one K=16 step,
//                     SBO = byte stride between consecutive 8-row groups along M/N.
//   MN-major operand: SBO = byte stride between consecutive 8-element MN chunks,
//                     LBO = byte stride between consecutive 8-deep K groups.
// bits [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
MMNN_DEVINL uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
// Same descriptor with a swizzle mode in bits [61,64): 0 none, 1 128B (32B atom base), 2 128B, 4 64B, 6 32B.
MMNN_DEVINL uint64_t make_smem_desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return make_smem_desc(addr, lbo, sbo) | ((uint64_t)(layout & 7u) << 61);
}
// Advancing a descriptor's start address by `bytes` (multiple of 16) is an add on the low word: the 14-bit field never
// carries because shared-memory addresses stay below 256 KB.
MMNN_DEVINL uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// Instruction descriptor for kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), majors (bits 15,16), N>>3 (17..22), M>>4 (24..28)
// operand format field: 0 = fp16, 1 = bf16 (both operands the same)
__host__ __device__ inline uint32_t make_idesc_ab(int M, int N, int a_mn_major, int b_mn_major, bool a_f16, bool b_f16) {
  // the A and B formats are separate fields: an fp16 operand (forward activations) can meet a bf16 one (gradients) in one MMA
  const uint32_t fa = a_f16 ? 0u : 1u, fb = b_f16 ? 0u : 1u;
  return (1u << 4) | (fa << 7) | (fb << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, bool f16) {
  return make_idesc_ab(M, N, a_mn_major, b_mn_major, f16, f16);
}

// ---------------------------------------------------------------- bf16 helpers
MMNN_DEVINL float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
MMNN_DEVINL float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
MMNN_DEVINL uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
MMNN_DEVINL float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// 16-bit pair <-> fp32 for either storage format
template <bool F16>
MMNN_DEVINL void unpack2(uint32_t w, float& lo, float& hi) {
  if (F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x; hi = f.y;
  } else {
    lo = bf16_lo(w); hi = bf16_hi(w);
  }
}
template <bool F16>
MMNN_DEVINL uint32_t pack2(float lo, float hi) {
  if (F16) {
    lo = fminf(fmaxf(lo, -65504.f), 65504.f);
    hi = fminf(fmaxf(hi, -65504.f), 65504.f);
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  return pack_bf16(lo, hi);
}
template <bool F16>
MMNN_DEVINL float round16(float x) {
  if (F16) return __half2float(__float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f)));
  return round_bf16(x);
}

MMNN_DEVINL uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
MMNN_DEVINL void sts16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Ampere-style asynchronous 16-byte copy global -> shared (LDGSTS); src_bytes == 0 writes 16 zero bytes.
// Used by every producer: all loads of a stage are in flight at once without occupying registers.
MMNN_DEVINL void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// .ca variant: the line is also kept in L1 -- for operands whose rows are re-read within the tile (the 9 shifted
// gradient tiles of the 3x3x3 weight gradient share most of their rows)
MMNN_DEVINL void cp_async16_ca(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
MMNN_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
MMNN_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
MMNN_DEVINL void cp_async_wait_dyn(int n) {   // at most n groups still pending
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    default: cp_async_wait<5>(); break;
  }
}
MMNN_DEVINL uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
MMNN_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// Sum over the 32 lanes of a warp of a per-lane vector v[0..31]; on return lane j holds the total of element j in v[0].
MMNN_DEVINL float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

}  // namespace mmnn
