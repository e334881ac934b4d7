// Speed-of-light probe for the tensor-core harmonic bank (kernels_bank_tc.cuh): ONLY the operand generation of
// hm_bank_tc_kernel -- the same per-thread instruction streams of its A warps (harmonic recurrences, TF32 hi / lo
// split, 32 words stored per half chunk) and of its B warps (seeds, coefficient products, split, four 16-byte stores),
// plus the coefficient staging -- with no tcgen05.mma, no tensor-memory traffic, no mbarrier hand-over, no read-back and
// no overlap-add. Stores go to shared memory (the A warps' tcgen05.st become st.shared of the same width). What this
// kernel takes is the floor of the present factorisation: everything above it in hm_bank_tc_kernel is pipeline hand-over.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/bank_floor tools/bank_floor.cu
//   run  : build/bank_floor [ctas_per_sm]          (prints ms per 409 600 frames of 128 harmonics, BASELINE configs[1])
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../libllsm2_b200/csrc/common.cuh"
#include "../libllsm2_b200/csrc/kernels_synth.cuh"
#include "../libllsm2_b200/csrc/kernels_bank_tc.cuh"

struct FloorParams { const float* f0; const float* ampl; const float* phse; int nfrm, nhar; float fs; float* sink; };

__global__ void __launch_bounds__(512) bank_floor_kernel(FloorParams P) {
  extern __shared__ __align__(16) char smem[];
  float* Cr = (float*)smem;                               // [2][4][BTC_CST]
  float* Ci = Cr + 2 * 4 * BTC_CST;
  float* bt = Ci + 2 * 4 * BTC_CST;                       // [BTC_NBUF][4096] B ring
  uint32_t* at = (uint32_t*)(bt + BTC_NBUF * 4096);       // [8 warps][32 lanes][32 words] stand-in for the A ring
  __shared__ BtcFrame finfo[BTC_NSLOT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * (BTC_NSLOT - 2);
  const size_t row = (size_t)b * P.nfrm;
  for(int s = tid; s < BTC_NSLOT; s += 512) {
    const int f = t0 - 1 + s;
    BtcFrame fi; fi.nufix = 0; fi.corr = 0; fi.nh = 0;
    if(f >= 0 && f < P.nfrm && P.f0[row + f] > 0) {
      const float f0n = P.f0[row + f] / P.fs;
      fi.nufix = __double2ull_rn((double)f0n * 18446744073709551616.0); fi.corr = 1e-3f; fi.nh = P.nhar;
    }
    finfo[s] = fi;
  }
  __syncthreads();
  float acc = 0.f;
  const int nck = (P.nhar + BTC_KC - 1) / BTC_KC;
  if(warp < 8) {                                          // ---- A warps
    const int qd = warp & 3, hh = warp >> 2, pp = lane;
    const int gC = tid >> 6, kC = tid & 63;
    uint4* my4 = (uint4*)at + tid;               // word j of thread t at uint4 index j * 256 + t: conflict-free 16-byte stores
    for(int grp = 0; grp < BTC_NSLOT / 4; grp ++) {
      const int s0 = 4 * grp;
      {   // coefficient staging of the group (a_k cos / sin of the corrected phase)
        const BtcFrame fi = finfo[s0 + gC];
        float* cr = Cr + ((grp & 1) * 4 + gC) * BTC_CST; float* cim = Ci + ((grp & 1) * 4 + gC) * BTC_CST;
        for(int k = kC; k < nck * BTC_KC; k += 64) {
          float a = 0.f, phv = 0.f;
          const int f = t0 - 1 + s0 + gC;
          if(k < fi.nh && f >= 0 && f < P.nfrm) { const size_t o = (row + f) * (size_t)P.nhar + k; a = P.ampl[o]; phv = P.phse[o]; }
          const float ph2 = (float)((double)phv - (double)fi.corr * ((double)k + 1.0));
          float sn, cs; __sincosf(ph2, &sn, &cs);
          cr[k] = a * cs; cim[k] = a * sn;
        }
      }
      float2 Wr, Wi, z2, z18;
      {
        const unsigned long long nfA = finfo[s0 + qd].nufix;
        const unsigned tA = 8u * (unsigned)pp, kka = 16u * (unsigned)hh + 1u;
        const float2 z1 = btc_phasor(nfA, tA);
        z2 = cmul(z1, z1); z18 = btc_phasor(nfA, 18u * tA);
        const float2 wa = btc_phasor(nfA, kka * tA), wb = cmul(wa, z1);
        Wr = make_float2(wa.x, wb.x); Wi = make_float2(wa.y, wb.y);
      }
      const float2 z2r = make_float2(z2.x, z2.x), z2i = make_float2(z2.y, z2.y), nz2i = make_float2(-z2.y, -z2.y);
      const float2 z18r = make_float2(z18.x, z18.x), z18i = make_float2(z18.y, z18.y), nz18i = make_float2(-z18.y, -z18.y);
      for(int c = 0; c < nck; c ++) {
#pragma unroll
        for(int hf = 0; hf < 2; hf ++) {
          uint32_t arh[8], arl[8], aih[8], ail[8];
#pragma unroll
          for(int i = 0; i < 4; i ++) {
            btc_split2(Wr, arh[2 * i], arh[2 * i + 1], arl[2 * i], arl[2 * i + 1]);
            btc_split2(Wi, aih[2 * i], aih[2 * i + 1], ail[2 * i], ail[2 * i + 1]);
            if(hf == 1 && i == 3) btc_rot2(Wr, Wi, z18r, z18i, nz18i);
            else btc_rot2(Wr, Wi, z2r, z2i, nz2i);
          }
          my4[0 * 256] = make_uint4(arh[0], arh[1], arh[2], arh[3]); my4[1 * 256] = make_uint4(arh[4], arh[5], arh[6], arh[7]);
          my4[2 * 256] = make_uint4(arl[0], arl[1], arl[2], arl[3]); my4[3 * 256] = make_uint4(arl[4], arl[5], arl[6], arl[7]);
          my4[4 * 256] = make_uint4(aih[0], aih[1], aih[2], aih[3]); my4[5 * 256] = make_uint4(aih[4], aih[5], aih[6], aih[7]);
          my4[6 * 256] = make_uint4(ail[0], ail[1], ail[2], ail[3]); my4[7 * 256] = make_uint4(ail[4], ail[5], ail[6], ail[7]);
        }
        __syncwarp();
      }
      acc += __uint_as_float(at[(grp * 7 + lane) & 31]);
    }
  } else {                                                // ---- B warps
    const int team = (warp - 8) >> 2, tT = tid - 256 - 128 * team;
    const int qB = tT & 7, jB = (tT >> 3) & 7, g0 = tT >> 6;
    int ci = 0, bf = 0;
    for(int grp = 0; grp < BTC_NSLOT / 4; grp ++) {
      const int s0 = 4 * grp, cb = grp & 1;
      const int c_first = (team ^ ci) & 1;
      float2 w[2], rho[2], rho64[2], r2r[2], r2i[2];
#pragma unroll
      for(int m = 0; m < 2; m ++) {
        const unsigned long long nfB = finfo[s0 + g0 + 2 * m].nufix;
        rho[m] = btc_phasor(nfB, (unsigned)qB); rho64[m] = btc_phasor(nfB, 64u * (unsigned)qB);
        const float2 rho2 = cmul(rho[m], rho[m]);
        r2r[m] = make_float2(rho2.x, rho2.x); r2i[m] = make_float2(rho2.y, rho2.y);
        w[m] = btc_phasor(nfB, (unsigned)((32 * c_first + 4 * jB + 1) * qB));
      }
      for(int c = 0; c < nck; c ++, ci ++) {
        if(((ci ^ team) & 1) == 0) {
#pragma unroll
          for(int m = 0; m < 2; m ++) {
            const int gB = g0 + 2 * m;
            const float2 e1 = cmul(w[m], rho[m]);
            float2 Er = make_float2(w[m].x, e1.x), Ei = make_float2(w[m].y, e1.y);
            const float4 cr = *(const float4*)(Cr + (cb * 4 + gB) * BTC_CST + c * BTC_KC + 4 * jB);
            const float4 ci4 = *(const float4*)(Ci + (cb * 4 + gB) * BTC_CST + c * BTC_KC + 4 * jB);
            const float2 cr01 = make_float2(cr.x, cr.y), cr23 = make_float2(cr.z, cr.w);
            const float2 ci01 = make_float2(ci4.x, ci4.y), ci23 = make_float2(ci4.z, ci4.w);
            const float2 br01 = ffma2(Ei, make_float2(-ci4.x, -ci4.y), fmul2(Er, cr01)), bi01 = ffma2(Ei, cr01, fmul2(Er, ci01));
            btc_rot2(Er, Ei, r2r[m], r2i[m], make_float2(-r2i[m].x, -r2i[m].y));
            const float2 br23 = ffma2(Ei, make_float2(-ci4.z, -ci4.w), fmul2(Er, cr23)), bi23 = ffma2(Ei, cr23, fmul2(Er, ci23));
            uint4 rh, rl, ih, il;
            btc_split2(br01, rh.x, rh.y, rl.x, rl.y); btc_split2(br23, rh.z, rh.w, rl.z, rl.w);
            btc_split2(bi01, ih.x, ih.y, il.x, il.y); btc_split2(bi23, ih.z, ih.w, il.z, il.w);
            float* d = bt + bf * 4096 + jB * 32 + gB * 256 + qB * 4;
            *(uint4*)(d) = rh; *(uint4*)(d + 1024) = rl; *(uint4*)(d + 2048) = ih; *(uint4*)(d + 3072) = il;
            w[m] = cmul(w[m], rho64[m]);
          }
          __syncwarp();
        }
        if(++ bf == BTC_NBUF) bf = 0;
      }
      acc += bt[(grp * 13 + tT) & 4095];
    }
  }
  if(acc == 123456.789f) P.sink[0] = acc;                 // keeps the stores alive
}

int main(int argc, char** argv) {
  const int B = 1024, F = 400, K = 128;
  const int per_sm = argc > 1 ? atoi(argv[1]) : 1;
  std::vector<float> f0((size_t)B * F), am((size_t)B * F * K), ph((size_t)B * F * K);
  srand(1);
  for(auto& v : f0) v = (rand() % 5 == 0) ? 0.f : 90.f + (rand() % 8000) / 100.f;
  for(auto& v : am) v = (rand() % 1000) / 1e4f;
  for(auto& v : ph) v = (rand() % 6283) / 1e3f - 3.14f;
  float *df0, *dam, *dph, *sink;
  cudaMalloc(&df0, f0.size() * 4); cudaMalloc(&dam, am.size() * 4); cudaMalloc(&dph, ph.size() * 4); cudaMalloc(&sink, 4);
  cudaMemcpy(df0, f0.data(), f0.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dam, am.data(), am.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dph, ph.data(), ph.size() * 4, cudaMemcpyHostToDevice);
  FloorParams P = {df0, dam, dph, F, K, 44100.f, sink};
  size_t smem = (size_t)(2 * 2 * 4 * BTC_CST + BTC_NBUF * 4096) * 4 + (size_t)256 * 32 * 4 + 64;
  // occupancy is set through the shared-memory footprint: 1 CTA per SM needs > 113 KB
  if(per_sm <= 1) smem = smem > 120 * 1024 ? smem : 120 * 1024;
  cudaFuncSetAttribute(bank_floor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((F + BTC_NSLOT - 3) / (BTC_NSLOT - 2), B);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int i = 0; i < 3; i ++) bank_floor_kernel<<<grid, 512, smem>>>(P);
  cudaEventRecord(e0);
  const int reps = 20;
  for(int i = 0; i < reps; i ++) bank_floor_kernel<<<grid, 512, smem>>>(P);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("bank operand generation only: %.3f ms per %d frames x %d harmonics (%d CTA(s) per SM by shared memory, %zu B; %s)\n",
    ms / reps, B * F, K, per_sm, smem, cudaGetErrorString(err));
  return 0;
}
