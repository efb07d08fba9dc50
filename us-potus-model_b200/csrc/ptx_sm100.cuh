// Thin inline-PTX wrappers for the sm_100a features the sampler kernel uses:
// tcgen05 (MMA / TMEM alloc / ld / st / commit / fences), mbarrier, bulk async copy (TMA 1-D).
// Descriptor bit layouts follow the PTX ISA "tcgen05 shared memory descriptor" /
// "instruction descriptor" tables (cross-checked against the CuTe headers shipped in this image:
// cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- bulk async copy (TMA, 1-D)
// global -> shared, completion signalled on an mbarrier (complete_tx).  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM allocation
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layout.
//   K-major : element (row r, k) lives at  (r%8)*16 + (r/8)*SBO + (k/8)*LBO + (k%8)*2   [2-byte elements]
//   MN-major: element (mn, k)   lives at  (mn%8)*2 + (mn/8)*SBO + (k%8)*16 + (k/8)*LBO
// bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor for kind::f16, fp16 A/B, fp32 accumulate.
//   [4,6) c_format=1(F32) | [7,10) a_format=0(F16) | [10,13) b_format=0 | bit15 a_major | bit16 b_major
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.  SASS: UTCHMMA.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- TMEM <-> registers (32 lanes x 32-bit, N columns)
// Thread i of warp w touches TMEM lane 32*(w%4)+i, columns [col, col+N).
#define PTX_R16(v, o) \
  "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), \
      "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), \
      "=r"(v[o + 14]), "=r"(v[o + 15])
#define PTX_W16(v, o) \
  "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]), \
      "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), \
      "r"(v[o + 15])

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : PTX_R16(v, 0)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      PTX_W16(v, 0)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : PTX_R16(v, 0), PTX_R16(v, 16)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      PTX_W16(v, 0), PTX_W16(v, 16)
      : "memory");
}


// float-typed variants (no integer<->float register moves around the asm)
#define PTX_RF16(v, o) \
  "=f"(v[o + 0]), "=f"(v[o + 1]), "=f"(v[o + 2]), "=f"(v[o + 3]), "=f"(v[o + 4]), "=f"(v[o + 5]), "=f"(v[o + 6]), \
      "=f"(v[o + 7]), "=f"(v[o + 8]), "=f"(v[o + 9]), "=f"(v[o + 10]), "=f"(v[o + 11]), "=f"(v[o + 12]), "=f"(v[o + 13]), \
      "=f"(v[o + 14]), "=f"(v[o + 15])
#define PTX_WF16(v, o) \
  "f"(v[o + 0]), "f"(v[o + 1]), "f"(v[o + 2]), "f"(v[o + 3]), "f"(v[o + 4]), "f"(v[o + 5]), "f"(v[o + 6]), "f"(v[o + 7]), \
      "f"(v[o + 8]), "f"(v[o + 9]), "f"(v[o + 10]), "f"(v[o + 11]), "f"(v[o + 12]), "f"(v[o + 13]), "f"(v[o + 14]), \
      "f"(v[o + 15])
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : PTX_RF16(v, 0)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      PTX_WF16(v, 0)
      : "memory");
}


// bulk L2 prefetch of a global range (bytes: multiple of 16)
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// 8-column variants (thread-private half vectors; keep register pressure low at 1024 threads / 64 regs)
__device__ __forceinline__ void tmem_ld8f(uint32_t taddr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8f(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// ---------------------------------------------------------------- fp32 -> (hi, lo) fp16 split
// x*scale ~= hi + lo * 2^-11, 22 significant bits.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
}

}  // namespace ptx
