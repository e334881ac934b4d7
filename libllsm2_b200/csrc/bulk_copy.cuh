// Bulk asynchronous global -> shared copies (the TMA engine's 1-D form, cp.async.bulk; SASS: UBLKCP) completed on an
// mbarrier, as the analysis kernels use them to stage a frame's sample slices one frame ahead of the arithmetic.
// Contract: source, destination and size are multiples of 16 bytes; ONE thread arms the barrier with the byte total
// (bulk_expect) and issues the copies of a stage; every consumer thread waits on the stage's phase parity (bulk_wait),
// after which the bytes are visible to it without a further block barrier.
// The g++ -DLLSM_EMU build (tests/emu, CPU thread emulation) performs the copy synchronously in the issuing thread
// and releases waiters through an atomic phase counter.
#pragma once
#include "common.cuh"

#ifdef LLSM_EMU
#include <sched.h>
typedef uint64_t bulk_bar_t;       // low half: completed phases; high half: bytes still expected (producer-private)
__device__ __forceinline__ void bulk_bar_init(bulk_bar_t* bar) { *bar = 0; }
__device__ __forceinline__ void bulk_expect(bulk_bar_t* bar, uint32_t bytes) {
  uint32_t* w = (uint32_t*)bar;
  if(bytes == 0) { __atomic_fetch_add(&w[0], 1u, __ATOMIC_RELEASE); return; }
  w[1] = bytes;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, bulk_bar_t* bar) {
  uint32_t* w = (uint32_t*)bar;
  memcpy(dst, src, bytes);
  w[1] -= bytes;
  if(w[1] == 0) __atomic_fetch_add(&w[0], 1u, __ATOMIC_RELEASE);
}
__device__ __forceinline__ void bulk_wait(bulk_bar_t* bar, uint32_t parity) {
  uint32_t* w = (uint32_t*)bar;
  while((__atomic_load_n(&w[0], __ATOMIC_ACQUIRE) & 1u) == parity) sched_yield();
}
#else
typedef uint64_t bulk_bar_t;
__device__ __forceinline__ uint32_t bulk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_bar_init(bulk_bar_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bulk_smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// the single producer's arrival, announcing the bytes of the stage (0 bytes: the phase completes at once)
__device__ __forceinline__ void bulk_expect(bulk_bar_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bulk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, bulk_bar_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
    :: "r"(bulk_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bulk_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait(bulk_bar_t* bar, uint32_t parity) {
  const uint32_t a = bulk_smem_u32(bar);
  asm volatile("{\n\t.reg .pred p;\n\tBULK_WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra BULK_DONE_%=;\n\tbra BULK_WAIT_%=;\n\tBULK_DONE_%=:\n\t}" :: "r"(a), "r"(parity) : "memory");
}
#endif
