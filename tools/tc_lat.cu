// Latency of the synchronisation primitives around tcgen05 (single warp, back to back), in cycles.
#include <cstdio>
#include <cuda_runtime.h>
#include "../libllsm2_b200/csrc/tcgen05.cuh"
__global__ void lat(long long* out, int nwarps_active) {
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tb;
  __shared__ float buf[1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if(warp == 0) tc::tmem_alloc(&tb, 64);
  if(threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar + 1, 1); tc::fence_mbar_init(); }
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t t = tb + ((uint32_t)(32 * (warp & 3)) << 16);
  const int R = 64;
  long long c[12];
  uint32_t r[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  c[0] = clock64();
  for(int i = 0; i < R; i ++) { asm volatile("" ::: "memory"); c[1] = clock64(); }
  c[1] = clock64();
  for(int i = 0; i < R; i ++) tc::mbar_wait(bar, 1);               // passes immediately (phase 0 pending, parity 1 "done")
  c[2] = clock64();
  for(int i = 0; i < R; i ++) tc::fence_after_sync();
  c[3] = clock64();
  for(int i = 0; i < R; i ++) tc::fence_before_sync();
  c[4] = clock64();
  for(int i = 0; i < R; i ++) { tc::tmem_st8(t, r); tc::tmem_st_wait(); }
  c[5] = clock64();
  for(int i = 0; i < R; i ++) { buf[threadIdx.x & 1023] = (float)i; tc::fence_smem_to_async(); }
  c[6] = clock64();
  for(int i = 0; i < R; i ++) { tc::tmem_ld8(t, r); tc::tmem_ld_wait(); }
  c[7] = clock64();
  for(int i = 0; i < R; i ++) { __syncwarp(); if(lane == 0) tc::mbar_arrive(bar + 1); __syncwarp(); }
  c[8] = clock64();
  for(int i = 0; i < R; i ++) { tc::tmem_st8(t, r); tc::tmem_st8(t + 8, r); tc::tmem_st8(t + 16, r); tc::tmem_st8(t + 24, r); tc::tmem_st_wait(); }
  c[9] = clock64();
  if(threadIdx.x == 0 && blockIdx.x == 0) for(int i = 0; i < 9; i ++) out[i] = (c[i + 1] - c[i]) / R;
  if(r[0] == 12345) out[11] = r[1];
  __syncthreads();
  if(warp == 0) tc::tmem_dealloc(tb, 64);
}
int main() {
  long long* d; cudaMalloc(&d, 128);
  for(int nthr : {32, 512}) {
    lat<<<1, nthr>>>(d, 0); cudaDeviceSynchronize();
    long long h[12]; cudaMemcpy(h, d, 96, cudaMemcpyDeviceToHost);
    printf("threads %d: clock %lld | mbar try_wait(pass) %lld | fence::after %lld | fence::before %lld | st8+wait %lld | sts+fence.proxy.async %lld | ld8+wait %lld | arrive %lld | 4xst8+wait %lld\n",
      nthr, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
  }
  return 0;
}
