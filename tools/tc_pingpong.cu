// Round-trip latency of mbarrier hand-over between two warps (try_wait vs test_wait polling), and of
// tcgen05.commit -> mbarrier, in cycles.
#include <cstdio>
#include <cuda_runtime.h>
#include "../libllsm2_b200/csrc/tcgen05.cuh"
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while(! ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
  }
}
template <int MODE>
__global__ void pingpong(long long* out, int iters) {
  __shared__ uint64_t a, b;
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if(warp == 0) tc::tmem_alloc(&tb, 32);
  if(threadIdx.x == 0) { tc::mbar_init(&a, 1); tc::mbar_init(&b, 1); tc::fence_mbar_init(); }
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  long long t0 = clock64();
  if(warp == 0) {
    for(int i = 0; i < iters; i ++) {
      if(lane == 0) tc::mbar_arrive(&a);
      __syncwarp();
      if(MODE == 1) mbar_spin(&b, i & 1); else tc::mbar_wait(&b, i & 1);
    }
  } else if(warp == 1) {
    for(int i = 0; i < iters; i ++) {
      if(MODE == 1) mbar_spin(&a, i & 1); else tc::mbar_wait(&a, i & 1);
      if(MODE == 2) { if(tc::elect_one()) tc::mma_commit(&b); __syncwarp(); }
      else { if(lane == 0) tc::mbar_arrive(&b); __syncwarp(); }
    }
  }
  long long t1 = clock64();
  if(threadIdx.x == 0) out[0] = (t1 - t0) / iters;
  __syncthreads();
  if(warp == 0) tc::tmem_dealloc(tb, 32);
}
int main() {
  long long* d; cudaMalloc(&d, 8); long long h;
  pingpong<0><<<1, 64>>>(d, 1000); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("try_wait ping-pong round trip: %lld cycles\n", h);
  pingpong<1><<<1, 64>>>(d, 1000); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("test_wait spin ping-pong round trip: %lld cycles\n", h);
  pingpong<2><<<1, 64>>>(d, 1000); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("arrive -> try_wait -> tcgen05.commit -> try_wait round trip: %lld cycles\n", h);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
