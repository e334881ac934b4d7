// Stand-alone timing / tracing harness for hm_bank_tc_kernel (no parity check: tests/ do that).
//   nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a [-DBTC_TRACE] -o build/btc_trace tools/btc_trace.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#ifdef BTC_TRACE
__device__ long long btc_trace_buf[3][512];
__device__ int btc_trace_n[3];
#define BTC_STAMP_DECL int btc_si = 0;
#define BTC_STAMP(who, tag) do { if(blockIdx.x == 2 && blockIdx.y == 7 && (threadIdx.x & 31) == 0 && btc_si < 510) { \
  btc_trace_buf[(who)][btc_si ++] = (clock64() << 8) | (tag); btc_trace_n[(who)] = btc_si; } } while(0)
#endif
#include "../libllsm2_b200/csrc/kernels_synth.cuh"

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024, F = 400, K = 128, N = 442;
  const float fs = 44100.f, hop = 220.5f;
  std::vector<float> f0((size_t)B * F), ampl((size_t)B * F * K), phse((size_t)B * F * K), win(N), frac(F, 0.f);
  std::vector<int> nhar((size_t)B * F, K), base(F);
  srand(3);
  for(auto& v : f0) v = 90.f + 80.f * rand() / RAND_MAX;
  for(size_t i = 0; i < (size_t)B * F; i += 7) if((i / 40) % 5 == 0) f0[i] = 0.f;
  for(auto& v : ampl) v = 0.01f * rand() / RAND_MAX;
  for(auto& v : phse) v = 6.28f * rand() / RAND_MAX - 3.14f;
  for(int i = 0; i < N; i ++) win[i] = 0.5f - 0.5f * cosf(2.f * 3.14159265f * i / (N - 1));
  for(int f = 0; f < F; f ++) base[f] = (int)lroundf(f * hop);
  const int ny = (int)lroundf((F + 1) * hop), stride = (ny + 3) & ~3;
  BankParams P; memset(&P, 0, sizeof(P));
  float *d_f0, *d_a, *d_p, *d_w, *d_fr, *d_y; int *d_nh, *d_b;
  cudaMalloc(&d_f0, f0.size() * 4); cudaMalloc(&d_a, ampl.size() * 4); cudaMalloc(&d_p, phse.size() * 4);
  cudaMalloc(&d_w, N * 4); cudaMalloc(&d_fr, F * 4); cudaMalloc(&d_y, (size_t)B * stride * 4);
  cudaMalloc(&d_nh, nhar.size() * 4); cudaMalloc(&d_b, F * 4);
  cudaMemcpy(d_f0, f0.data(), f0.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_a, ampl.data(), ampl.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_p, phse.data(), phse.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_w, win.data(), N * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_fr, frac.data(), F * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_nh, nhar.data(), nhar.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, base.data(), F * 4, cudaMemcpyHostToDevice);
  P.nfrm = F; P.maxnhar = K; P.f0 = d_f0; P.nhar = d_nh; P.ampl = d_a; P.phse = d_p; P.hm_base = d_b; P.hm_frac = d_fr; P.win = d_w;
  P.n_hm = N; P.ny = ny; P.nsamp = stride; P.stride = stride; P.fs = fs; P.has_options = 1; P.use_iczt = 1; P.iczt_a = 0.275f; P.iczt_b = 2.26f;
  P.hop = hop; P.y_sin = d_y;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int tc_on = 1; tc_on >= 0; tc_on --) {
    float best = 1e9f;
    for(int it = 0; it < 6; it ++) {
#ifdef BTC_TRACE
      if(tc_on) { int z[3] = {0, 0, 0}; cudaMemcpyToSymbol(btc_trace_n, z, sizeof(z)); }
#endif
      cudaEventRecord(e0);
      int rc = tc_on ? launch_hm_bank_tc(P, B, F, 0) : (setenv("LLSM_BANK_TC", "0", 1), launch_hm_bank(P, B, F, 0));
      cudaEventRecord(e1); cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if(it > 0) best = std::min(best, ms);
      if(e != cudaSuccess || rc != 0) { printf("error rc %d %s\n", rc, cudaGetErrorString(e)); return 1; }
    }
    std::vector<float> y((size_t)stride); cudaMemcpy(y.data(), d_y + (size_t)7 * stride, stride * 4, cudaMemcpyDeviceToHost);
    double cs = 0; for(float v : y) cs += (double)v * v;
    printf("%s: %.3f ms for %d frames (%.1f M frames/s), row-7 energy %.9g\n", tc_on ? "tensor-core bank" : "CUDA-core bank", best, B * F, B * F / best / 1e3, cs);
  }
#ifdef BTC_TRACE
  {
    static long long h[3][512]; int n[3];
    cudaMemcpyFromSymbol(h, btc_trace_buf, sizeof(h)); cudaMemcpyFromSymbol(n, btc_trace_n, sizeof(n));
    for(int w = 0; w < 3; w ++) {
      printf("trace %d (%d stamps of the last launch...):", w, n[w]);
      long long t0 = h[w][0] >> 8;
      for(int i = 0; i < n[w] && i < 140; i ++) printf(" %lld:%lld", h[w][i] & 255, (h[w][i] >> 8) - t0);
      printf("\n");
    }
  }
#endif
  return 0;
}
