// Hardware probe for the tcgen05 conventions csrc/tcgen05.cuh relies on (run on a B200):
//   D[128 x 32] = A[128 x 32] B[32 x 32], A written to tensor memory with tcgen05.st, B in shared memory
//   N-major without swizzle, four K = 8 tcgen05.mma.kind::tf32, result read back with tcgen05.ld.
// Reports, per variant of the descriptor fields, the maximum deviation from a CPU product with
// operands truncated to TF32 and with operands rounded to nearest TF32.
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -o tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../libllsm2_b200/csrc/tcgen05.cuh"

#define KG 1024   // bytes between k-groups of 8 (one slab per MMA)
#define NG 128    // bytes between 4-column core matrices along N

__global__ void __launch_bounds__(256) probe(const float* A, const float* B, float* D, int variant, int passes, float* dbg) {
  __shared__ __align__(1024) float sB[32 * 32];
  __shared__ __align__(1024) float sA[128 * 32];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, qd = warp & 3, h = warp >> 2;
  if(warp == 0) tc::tmem_alloc(&tbase_s, 256);
  if(threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tbase_s;
  // A: row m = 32 qd + lane, columns 16 h .. 16 h + 15
  {
    uint32_t r[16];
    const int m = 32 * qd + lane;
    for(int j = 0; j < 16; j ++) r[j] = __float_as_uint(A[m * 32 + 16 * h + j]);
    tc::tmem_st16(tbase + ((uint32_t)(32 * qd) << 16) + 16 * h, r);
    tc::tmem_st_wait();
    if(dbg) {
      uint32_t rb[8];
      tc::tmem_ld8(tbase + ((uint32_t)(32 * qd) << 16) + 16 * h, rb);
      tc::tmem_ld_wait();
      for(int j = 0; j < 8; j ++) dbg[16 + m * 32 + 16 * h + j] = __uint_as_float(rb[j]);
      if(threadIdx.x == 0) dbg[0] = (float)tbase;
    }
  }
  // B[k][n]: variants 0/1 N-major core matrices, variants >= 2 K-major core matrices
  for(int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
    int k = e >> 5, n = e & 31;
    int off = variant < 2 ? (k >> 3) * KG + (n >> 2) * NG + (k & 7) * 16 + (n & 3) * 4
                          : (n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;
    sB[off >> 2] = B[k * 32 + n];
  }
  // A[m][k] K-major core matrices in shared memory (variants 4, 5)
  for(int e = threadIdx.x; e < 128 * 32; e += blockDim.x) {
    int m = e >> 5, k = e & 31;
    int off = (m >> 3) * 1024 + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4;
    sA[off >> 2] = A[m * 32 + k];
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  if(threadIdx.x == 0) {
    tc::fence_after_sync();
    const uint32_t idesc = tc::idesc_tf32(128, 32, variant < 2);
    for(int rep = 0; rep < passes; rep ++)
      for(int ks = 0; ks < 4; ks ++) {
        uint64_t bd;
        if(variant < 2) {
          uint32_t sa = tc::smem_u32(sB) + ks * KG;
          bd = variant == 0 ? tc::smem_desc(sa, KG, NG) : tc::smem_desc(sa, NG, KG);
        } else {
          uint32_t sa = tc::smem_u32(sB) + ks * 256;
          bd = (variant & 1) == 0 ? tc::smem_desc(sa, 128, 1024) : tc::smem_desc(sa, 1024, 128);
        }
        if(variant < 4) tc::mma_tf32_ts(tbase + 128, tbase + 8 * ks, bd, idesc, (rep | ks) ? 1u : 0u);
        else {
          uint32_t sa = tc::smem_u32(sA) + ks * 256;
          uint64_t ad = (variant & 1) == 0 ? tc::smem_desc(sa, 128, 1024) : tc::smem_desc(sa, 1024, 128);
          tc::mma_tf32_ss(tbase + 128, ad, bd, idesc, (rep | ks) ? 1u : 0u);
        }
      }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  if(h == 0) {
    for(int c = 0; c < 32; c += 8) {
      uint32_t r[8];
      tc::tmem_ld8(tbase + ((uint32_t)(32 * qd) << 16) + 128 + c, r);
      tc::tmem_ld_wait();
      for(int j = 0; j < 8; j ++) D[(32 * qd + lane) * 32 + c + j] = __uint_as_float(r[j]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if(warp == 0) tc::tmem_dealloc(tbase, 256);
}

static float tf_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float tf_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

static const char* names[6] = {"TS, B N-major, LBO=K SBO=N", "TS, B N-major, LBO=N SBO=K", "TS, B K-major, LBO=K SBO=N",
  "TS, B K-major, LBO=N SBO=K", "SS, K-major, LBO=K SBO=MN", "SS, K-major, LBO=MN SBO=K"};
int main() {
  std::vector<float> A(128 * 32), B(32 * 32), D(128 * 32);
  srand(1);
  for(auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
  for(auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD, *dDbg; cudaMalloc(&dDbg, (16 + 128 * 32) * 4); cudaMemset(dDbg, 0, (16 + 128 * 32) * 4);
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for(int variant = 0; variant < 6; variant ++) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 256>>>(dA, dB, dD, variant, 1, dDbg);
    cudaError_t e = cudaDeviceSynchronize();
    if(e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    { std::vector<float> dbg(16 + 128 * 32); cudaMemcpy(dbg.data(), dDbg, dbg.size() * 4, cudaMemcpyDeviceToHost);
      double ea = 0; for(int m = 0; m < 128; m ++) for(int k = 0; k < 32; k ++) if((k & 15) < 8) ea = fmax(ea, fabs(dbg[16 + m * 32 + k] - A[m * 32 + k]));
      printf("tbase = %g, st/ld round trip max err %.3e\n", dbg[0], ea); }
    double et = 0, er = 0, ef = 0;
    for(int m = 0; m < 128; m ++) for(int n = 0; n < 32; n ++) {
      double st = 0, sr = 0, sf = 0;
      for(int k = 0; k < 32; k ++) {
        st += (double)tf_trunc(A[m * 32 + k]) * tf_trunc(B[k * 32 + n]);
        sr += (double)tf_rn(A[m * 32 + k]) * tf_rn(B[k * 32 + n]);
        sf += (double)A[m * 32 + k] * B[k * 32 + n];
      }
      et = fmax(et, fabs(st - D[m * 32 + n])); er = fmax(er, fabs(sr - D[m * 32 + n]));
      ef = fmax(ef, fabs(sf - D[m * 32 + n]));
    }
    printf("variant %d (%s): max |D - trunc| = %.3e  |D - rn| = %.3e  |D - fp32| = %.3e   D[0][0..3] = %g %g %g %g\n",
           variant, names[variant], et, er, ef,
           D[0], D[1], D[2], D[3]);
  }
  // timing: many accumulating MMAs back to back (tensor-pipe cost of M=128, N=32, K=8)
  for(int passes : {64, 256}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<148, 256>>>(dA, dB, dD, 0, passes, nullptr);
    cudaEventRecord(e0);
    probe<<<148, 256>>>(dA, dB, dD, 0, passes, nullptr);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("passes %d: %.3f us per launch -> %.1f ns per MMA (one CTA per SM)\n", passes, ms * 1e3, ms * 1e6 / (passes * 4));
  }
  return 0;
}
