// Shared device helpers of the tcgen05 kernels (tc_gemm.cu, tc_conv3.cu): mbarrier / bulk-copy / tcgen05 PTX wrappers,
// UMMA shared-memory descriptors, the fp16 hi/lo split and the coalesced row store.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace dawn {
namespace tc {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Relaxed arrive for the A producers: the default .release form compiles to MEMBAR.ALL.CTA, which also waits for the
// producers' outstanding register-prefetch loads (two panels ahead) and serialised the whole prefetch (measured ~1000
// cycles per panel).  Ordering of the operand writes is provided by the preceding fence.proxy.async; the consumer side
// (mbarrier try_wait, acquire) is unchanged.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");   // compiler barrier only
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Elected-lane forms for an issuer WARP that runs its control flow warp-uniformly (all 32 lanes wait on the barriers and compute the
// operands; one lane -- always the same one, so the commits track its MMAs -- executes the instruction).  With provably uniform operands
// the descriptors stay in uniform registers and an MMA costs one UTCHMMA; issued from inside an `if (lane == 0)` region every operand
// goes through a per-lane R2UR loop (~75 cycles of issue per MMA, more than the tensor pipe needs for a 128 x 64 x 16 MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t r;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t"
      "}" : "=r"(r));
  return r != 0;
}
__device__ __forceinline__ void tc_mma_f16_elected(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elected(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled operand panel: rows of 128 B, 8-row atoms of 1024 B (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// descriptor for an operand whose first row sits `shift` rows (of 128 B) into a 1024-byte swizzle atom
// byte offset of 16-byte chunk c (0..7) of row r inside a swizzled panel
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }

// (x0, x1) -> packed fp16 hi pair and fp16 lo pair.  hi is rounded to 11 significant bits in fp32 with two integer
// ops (so its fp16 conversion is exact and no f16->f32 unpack is needed: the conversion pipe was the measured
// producer bottleneck); lo = x - hi is exact in fp32 and rounded once to fp16.  Below fp16's normal range the
// conversions go subnormal: absolute error <= 2^-25, irrelevant next to O(1) outputs.
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float h0 = __uint_as_float((__float_as_uint(x0) + 0x1000u) & 0xFFFFE000u);
  const float h1 = __uint_as_float((__float_as_uint(x1) + 0x1000u) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(h0, h1);
  const __half2 l = __floats2half2_rn(x0 - h0, x1 - h1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Store this warp's 32 rows x 64 columns (row-per-lane registers) to global memory with full-sector transactions:
// 16 columns at a time go through a per-warp shared-memory buffer so that one store instruction writes 8 rows x 64
// contiguous bytes instead of 32 rows x 16 bytes at a multi-KB stride (measured: the strided form capped the qkv
// projection's output stream at ~1.6 TB/s).
__device__ __forceinline__ void store_rows_coalesced(float* wbuf, const float (&acc)[64], float* out, size_t opix, int ldo,
                                                     int n0, bool rv, int lane) {
  // destination rows of this lane in the read-back phase: r = lane/4 + 8k (fetched once from the owning lanes)
  float* dst[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = (lane >> 2) + 8 * k;
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)opix, r);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)((unsigned long long)opix >> 32), r);
    const int ok = __shfl_sync(0xffffffffu, rv ? 1 : 0, r);
    const size_t px = ((size_t)hi << 32) | lo;
    dst[k] = ok ? (out + px * ldo + n0 + (lane & 3) * 4) : nullptr;
  }
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(wbuf + lane * 20 + j * 4) =
          make_float4(acc[pass * 16 + j * 4], acc[pass * 16 + j * 4 + 1], acc[pass * 16 + j * 4 + 2], acc[pass * 16 + j * 4 + 3]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = (lane >> 2) + 8 * k;
      const float4 v = *reinterpret_cast<const float4*>(wbuf + r * 20 + (lane & 3) * 4);
      if (dst[k]) *reinterpret_cast<float4*>(dst[k] + pass * 16) = v;
    }
  }
}

}  // namespace tc
}  // namespace dawn
