// tcgen05.mma issue-rate probe: cycles per kind::tf32 MMA (M = 128, K = 8, A in tensor memory, B K-major in
// shared memory) as a function of N and of the number of independent accumulators.
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -o build/tc_rate tools/tc_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../libllsm2_b200/csrc/tcgen05.cuh"

template <int N, int NACC, bool SS, int NISSUE = 1>
__global__ void __launch_bounds__(128) rate(long long* out, int iters) {
  __shared__ __align__(1024) float sB[256 * 8];     // N x 8 K-major: (n/8) * 256 B + (k/4) * 128 B
  __shared__ __align__(1024) float sA[128 * 8];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5;
  for(int i = threadIdx.x; i < 256 * 8; i += blockDim.x) sB[i] = 0.f;
  for(int i = threadIdx.x; i < 128 * 8; i += blockDim.x) sA[i] = 0.f;
  if(warp == 0) tc::tmem_alloc(&tbase_s, 512);
  if(threadIdx.x == 0) { tc::mbar_init(&bar, NISSUE); tc::fence_mbar_init(); }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tbase_s;
  {
    uint32_t r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    tc::tmem_st8(tbase + ((uint32_t)(32 * warp) << 16), r);
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if(uwarp < NISSUE) {
    tc::fence_after_sync();
    const uint32_t idesc = tc::idesc_tf32(128, N, false);
    const uint64_t bd = tc::smem_desc(tc::smem_u32(sB), 128, 256);
    const uint64_t ad = tc::smem_desc(tc::smem_u32(sA), 128, 256);
    long long t0 = clock64();
    for(int it = 0; it < iters; it ++) {
#pragma unroll
      for(int a = 0; a < NACC; a ++) if(tc::elect_one()) {
        if(SS) tc::mma_tf32_ss(tbase + 16 + (a + uwarp * NACC) * N, ad, bd, idesc, 1u);
        else   tc::mma_tf32_ts(tbase + 16 + (a + uwarp * NACC) * N, tbase, bd, idesc, 1u);
      }
    }
    long long t1 = clock64();
    if(tc::elect_one()) tc::mma_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if(blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if(warp == 0) tc::tmem_dealloc(tbase, 512);
}

template <int N, int NACC, bool SS, int NISSUE = 1>
static void run(long long* d) {
  const int iters = 2048 / NACC;
  rate<N, NACC, SS, NISSUE><<<148, 128>>>(d, iters);
  rate<N, NACC, SS, NISSUE><<<148, 128>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("issuers=%d ", NISSUE); printf("%s N=%3d acc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", SS ? "SS" : "TS", N, NACC,
         (double)h[0] / (iters * NACC * NISSUE), (double)h[1] / (iters * NACC * NISSUE), cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<16, 1, false>(d); run<32, 1, false>(d); run<32, 2, false>(d); run<32, 4, false>(d);
  run<64, 1, false>(d); run<64, 2, false>(d); run<128, 1, false>(d); run<128, 2, false>(d); run<256, 1, false>(d);
  run<32, 1, false, 2>(d); run<32, 1, false, 4>(d); run<32, 2, false, 4>(d); run<16, 1, false, 4>(d); run<64, 1, false, 2>(d); run<64, 1, false, 4>(d);
  run<32, 2, true>(d); run<64, 2, true>(d); run<128, 2, true>(d); run<256, 1, true>(d);
  return 0;
}
