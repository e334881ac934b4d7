// Thin inline-PTX wrappers around the Blackwell (sm_100a) 5th-generation tensor-core instructions:
// tensor-memory allocation, register <-> tensor-memory moves, the single-thread tcgen05.mma issue,
// commit to an mbarrier, and the fences between the generic / async / tensor proxies.
//
// Conventions used by the kernels of this library:
//   * cta_group::1 only (one CTA drives its own SM's tensor core);
//   * kind::tf32 with FP32 accumulation, M = 128: accumulator row m lives in tensor-memory lane m,
//     accumulator column n in tensor-memory column base + n;
//   * A operand in tensor memory (row m in lane m, K index in consecutive 32-bit columns), written
//     straight from registers with tcgen05.st -- no shared-memory round trip;
//   * B operand in shared memory, N-major ("MN-major"), no swizzle: 8 (k) x 16 B (4 n) core matrices,
//     the descriptor's stride byte offset steps along N.
// A tensor-memory address is (lane << 16) | column. A warp can only touch the 32 lanes of its own
// quadrant: lanes 32 * (warp_id % 4) .. + 31.
#pragma once
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// one lane of a converged warp (the compiler can keep the operands of the elected lane's tcgen05 instructions in
// uniform registers, which a plain `lane == 0` branch prevents)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- tensor memory allocation (one warp, all 32 lanes) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

// ---- fences ----
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core's operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ----
#ifndef LLSM_MBAR_SUSPEND_NS
#define LLSM_MBAR_SUSPEND_NS 20000u   // try_wait time hint: a waiting warp sleeps instead of spinning through issue slots
#endif
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "LLSM_MBAR_WAIT:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"   // suspends up to the hint, wakes on completion
    "@p bra LLSM_MBAR_DONE;\n\t"
    "bra LLSM_MBAR_WAIT;\n\t"
    "LLSM_MBAR_DONE:\n\t}"
    :: "r"(smem_u32(bar)), "r"(parity), "r"(LLSM_MBAR_SUSPEND_NS) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
// named barrier among a subset of the CTA's warps (nthreads a multiple of 32)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// ---- tcgen05.mma, A from tensor memory, B from shared memory, one issuing thread ----
// accumulate == 0 overwrites D, otherwise D += A B.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
    :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// both operands from shared memory
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
    :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all tcgen05.mma issued so far by this thread: arrive on the mbarrier when they have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}

// instruction descriptor: kind::tf32, FP32 accumulate, A K-major (tensor memory), B N-major or K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool b_n_major) {
  return (1u << 4)                       // D format F32
       | (2u << 7) | (2u << 10)          // A, B format TF32
       | (0u << 15)                      // A K-major
       | ((b_n_major ? 1u : 0u) << 16)   // B major
       | ((uint32_t)(N >> 3) << 17)
       | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, no swizzle. Offsets in bytes (multiples of 16).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu)
       | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16)
       | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32)
       | (1ull << 46);                   // descriptor version of sm_100
}

// ---- registers <-> tensor memory, shape 32x32b: lane l of the warp <-> tensor-memory lane base + l,
//      register j <-> column base + j ----
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
    :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
    :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
    : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc
