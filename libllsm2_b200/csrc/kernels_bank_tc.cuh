// Harmonic bank + Hann + overlap-add on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as hm_bank_ola_kernel (kernels_synth.cuh): llsm_synthesize_harmonics_l0
// (layer0.c:117-146) with the per-frame generator behind it (llsmutils.c:45-58, dsputils.c:328-351);
// every output sample is written exactly once, frames are summed in ascending order.
//
// Factorisation. A frame is y[n] = sum_k a_k cos(k w n + phi_k), n in [-H, H). Write n = 8 p + q
// (q = 0..7) and use cos(-x) = cos(x):
//     y[ 8 p + q] = U[p][q] - V[p][q],   y[-8 p + q] = U[p][q] + V[p][q],   p = 0..31
//     U = Ar Br,  V = Ai Bi,   A[p][k] = e^{i k w 8 p}  (32 x K),   B[k][q] = a_k e^{i (k w q + phi_k)}  (K x 8)
// so a frame costs (32 + 8) K generated phasors and two small real GEMMs instead of 2 H K sinusoid
// samples. Four frames of one utterance share an M = 128 tcgen05.mma: frame g occupies tensor-memory
// lanes 32 g .. 32 g + 31 (rows of A) and columns 8 g .. 8 g + 7 of the N = 32 operand B; only the four
// diagonal 32 x 8 blocks of the product are read back (the tensor pipe has the slack: the kernel is bound
// by generating the operands). Accuracy: each operand is split x = hi + lo with hi = x truncated to TF32
// (what the hardware reads) and lo = x - hi, and three products hi hi + lo hi + hi lo are accumulated in
// FP32 -- 2^-21 relative per term, as good as the FP32 recurrences of the CUDA-core kernel.
//
// A is written from registers straight into tensor memory (tcgen05.st), B goes through 16 KB of shared
// memory (K-major core matrices, no swizzle), the accumulators live in tensor memory. Four warps issue
// the MMAs concurrently (one issuing thread sustains one small MMA per ~45 cycles, the pipe takes one per
// 16: measured with tools/tc_rate.cu), each into its own accumulator. Two CTAs per SM alternate between
// operand generation and MMA.
#pragma once
#ifndef LLSM_EMU
#include "tcgen05.cuh"

#define BTC_THREADS 256
#define BTC_KC 32           // harmonics per MMA chunk (columns of each A tile)
#define BTC_SLOT 512        // floats per parked frame: sample n lives at index n + 256
#define BTC_NSLOT 32        // frame slots per CTA (8 groups of 4), 30 owned output tiles
#define BTC_CST 128         // harmonics whose coefficients are staged together
#define BTC_MAXH 248        // largest half window: n = 8 p + q, p < 32

struct BtcFrame { unsigned long long nufix; float corr; int nh; };   // nufix = nu 2^64 (turns per sample and harmonic)

// e^{i 2 pi t nu}: the product is reduced exactly in 64-bit fixed point (the turn count wraps), the sine and
// cosine come from the special-function unit (absolute error 2^-21.4: below the TF32 split's 2^-21 per term)
__device__ __forceinline__ float2 btc_phasor(unsigned long long nufix, unsigned t) {
  const unsigned long long p = nufix * (unsigned long long)t;
  const float a = (float)(int)(unsigned)(p >> 32) * 1.4629180792671596e-9f;     // 2 pi / 2^32
  return make_float2(__cosf(a), __sinf(a));
}

__device__ __forceinline__ void btc_split2(float2 x, uint32_t& h0, uint32_t& h1, uint32_t& l0, uint32_t& l1) {
  h0 = __float_as_uint(x.x) & 0xffffe000u; h1 = __float_as_uint(x.y) & 0xffffe000u;
  float2 lo = ffma2(make_float2(__uint_as_float(h0), __uint_as_float(h1)), make_float2(-1.f, -1.f), x);
  l0 = __float_as_uint(lo.x); l1 = __float_as_uint(lo.y);
}

__global__ void __launch_bounds__(BTC_THREADS, 2) hm_bank_tc_kernel(BankParams P) {
  extern __shared__ __align__(1024) char smem[];
  float* fb = (float*)smem;                                  // [32][512]
  float* swin = fb + BTC_NSLOT * BTC_SLOT;                   // [512]  window at index n + 256
  float* bt = swin + BTC_SLOT;                               // 4 tiles x 1024 floats: Brh, Brl, Bih, Bil
  float* Cr = bt + 4 * 1024;                                 // [4][128]
  float* Ci = Cr + 4 * BTC_CST;                              // [4][128]
  BtcFrame* finfo = (BtcFrame*)(Ci + 4 * BTC_CST);           // [32]
  int* sb = (int*)(finfo + BTC_NSLOT);                       // [32] frame position
  int* sv = sb + BTC_NSLOT;                                  // [32] slot holds a voiced frame
  uint64_t* bar = (uint64_t*)(sv + BTC_NSLOT);
  uint32_t* tbase_s = (uint32_t*)(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qd = warp & 3, hh = warp >> 2;
  const int b = blockIdx.y, seg = blockIdx.x;
  const int F = BTC_NSLOT - 2;
  const int N = P.n_hm, H = N >> 1;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int ny_b = P.ny_utt ? P.ny_utt[b] : P.ny;
  const int t0 = seg * F;
  const int start = t0 == 0 ? 0 : (t0 < nf ? P.hm_base[t0] : P.nsamp);
  const int end = (t0 + F < nf) ? P.hm_base[t0 + F] : P.nsamp;
  if(start >= end) return;                                   // uniform per CTA
  const size_t row = (size_t)b * P.nfrm;

  // ---- set-up: tensor memory, barrier, window, per-frame scalars ----
  if(warp == 0) tc::tmem_alloc(tbase_s, 256);
  if(tid == 0) { tc::mbar_init(bar, 4); tc::fence_mbar_init(); }
  for(int i = tid; i < BTC_SLOT; i += BTC_THREADS) {
    int j = i - 256 + H;
    swin[i] = (j >= 0 && j < N) ? P.win[j] : 0.f;
  }
  if(tid < BTC_NSLOT) {
    const int s = tid, f = t0 - 1 + s;
    const bool inrange = f >= 0 && f < nf;
    float f0 = 0; int nh = 0;
    if(inrange) { f0 = P.f0[row + f]; nh = P.nhar[row + f]; }
    if(nh > 2048) nh = 2048;                                 // layer0.c:119,130
    bool voiced = inrange && f0 > 0 && nh > 0;               // layer0.c:125
    if(voiced && P.frame_mask) voiced = P.frame_mask[row + f] != 0;
    if(P.frame_hi > 0 && (f < P.frame_lo || f >= P.frame_hi)) voiced = false;
    sb[s] = inrange ? P.hm_base[f] : (f < 0 ? -(1 << 28) : (1 << 28));
    sv[s] = voiced ? 1 : 0;
    BtcFrame fi; fi.nufix = 0; fi.corr = 0; fi.nh = 0;
    if(voiced) {
      // per-frame scalars in the reference's precision (layer0.c:127-134, llsmutils.c:45-58, dsputils.c:338-348)
      const float f0n = f0 / P.fs;
      const float omega0 = (float)(2.0 * LLSM_PI * (double)f0n);
      bool iczt = false;
      if(P.has_options && P.use_iczt)
        iczt = log((double)N) * (double)P.iczt_a < log((double)nh) - (double)P.iczt_b;
      if(iczt && nh > N - 1) nh = N - 1;
      fi.nufix = __double2ull_rn((iczt ? (double)omega0 / (2.0 * LLSM_PI) : (double)f0n) * 18446744073709551616.0);
      const float frac = P.hm_frac[f];
      fi.corr = (float)((double)(frac * 2.0f) * LLSM_PI / (double)P.fs * (double)f0);
      fi.nh = nh;
    }
    finfo[s] = fi;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tbase_s;
  const uint32_t tlane = tbase + ((uint32_t)(32 * qd) << 16);

  // roles
  const int pp = lane;                                        // A: row (coarse time index) of frame qd
  const int gB = tid >> 6, jB = (tid >> 3) & 7, qB = tid & 7; // B: frame, k-quad, fine time index
  uint32_t parity = 0;
  bool pending = false;
  float pf_a[2], pf_p[2];                                     // prefetched amplitudes / phases of group pf_grp
  int pf_grp = 0;
#pragma unroll
  for(int it = 0; it < 2; it ++) {
    const int i = tid + it * BTC_THREADS, g = i >> 7, kl = i & 127;
    pf_a[it] = 0.f; pf_p[it] = 0.f;
    if(kl < finfo[g].nh) { const size_t o = (row + t0 - 1 + g) * (size_t)P.maxnhar + kl; pf_a[it] = P.ampl[o]; pf_p[it] = P.phse[o]; }
  }
  // MMA issue state of warps 0..3 (lane 0): accumulator t = 2 (V ? 1 : 0) + (slab parity)
  const uint32_t mm_idesc = tc::idesc_tf32(128, 32, false);
  const uint32_t mm_d = tbase + 128 + 32 * (warp & 3);
  const uint32_t mm_ah = tbase + 64 * ((warp >> 1) & 1) + 8 * (warp & 1), mm_al = mm_ah + 32;
  uint64_t mm_bh[2], mm_bl[2];
#pragma unroll
  for(int m = 0; m < 2; m ++) {
    const uint32_t bh = tc::smem_u32(bt) + 8192 * ((warp >> 1) & 1) + 256 * ((warp & 1) + 2 * m);
    mm_bh[m] = tc::smem_desc(bh, 128, 1024); mm_bl[m] = tc::smem_desc(bh + 4096, 128, 1024);
  }

  for(int grp = 0; grp < BTC_NSLOT / 4; grp ++) {
    const int s0 = 4 * grp;
    int nhmax = 0;
#pragma unroll
    for(int g = 0; g < 4; g ++) nhmax = max(nhmax, finfo[s0 + g].nh);
    if(nhmax == 0) continue;                                  // uniform: nothing voiced in the group
    const int nck = (nhmax + BTC_KC - 1) / BTC_KC;

    // ---- phasor seeds ----
    float2 Wr, Wi, z2, z16;
    {
      const unsigned long long nfA = finfo[s0 + qd].nufix;
      const unsigned tA = 8u * (unsigned)pp, kka = 16u * (unsigned)hh + 1u;
      const float2 z1 = btc_phasor(nfA, tA);
      z2 = cmul(z1, z1); z16 = btc_phasor(nfA, 16u * tA);
      const float2 wa = btc_phasor(nfA, kka * tA), wb = cmul(wa, z1);
      Wr = make_float2(wa.x, wb.x); Wi = make_float2(wa.y, wb.y);
    }
    const float2 z2r = make_float2(z2.x, z2.x), z2i = make_float2(z2.y, z2.y), nz2i = make_float2(-z2.y, -z2.y);
    const unsigned long long nfB = finfo[s0 + gB].nufix;
    const float2 rho = btc_phasor(nfB, (unsigned)qB), rho2 = cmul(rho, rho), rho32 = btc_phasor(nfB, 32u * (unsigned)qB);
    float2 w = btc_phasor(nfB, (unsigned)((4 * jB + 1) * qB));

    for(int c = 0; c < nck; c ++) {
      if((c & 3) == 0) {
        // stage a_k cos(phi'_k), a_k sin(phi'_k), phi'_k = phse[k] - corr (k + 1)  (layer0.c:132)
#pragma unroll
        for(int it = 0; it < 2; it ++) {
          const int i = tid + it * BTC_THREADS, g = i >> 7, kl = i & 127, k = c * BTC_KC + kl;
          const BtcFrame fi = finfo[s0 + g];
          float a = pf_a[it], phv = pf_p[it];
          if(c > 0 || pf_grp != grp) {
            a = 0.f; phv = 0.f;
            if(k < fi.nh) { const size_t o = (row + t0 - 1 + s0 + g) * (size_t)P.maxnhar + k; a = P.ampl[o]; phv = P.phse[o]; }
          }
          const float ph = (float)((double)phv - (double)fi.corr * ((double)k + 1.0));
          float sn, cs; __sincosf(ph, &sn, &cs);
          Cr[i] = a * cs; Ci[i] = a * sn;
        }
        if(c == 0 && grp + 1 < BTC_NSLOT / 4) {
          pf_grp = grp + 1;
          // the next group's first coefficients: in flight while this group computes
#pragma unroll
          for(int it = 0; it < 2; it ++) {
            const int i = tid + it * BTC_THREADS, g = i >> 7, kl = i & 127;
            pf_a[it] = 0.f; pf_p[it] = 0.f;
            if(kl < finfo[s0 + 4 + g].nh) {
              const size_t o = (row + t0 - 1 + s0 + 4 + g) * (size_t)P.maxnhar + kl; pf_a[it] = P.ampl[o]; pf_p[it] = P.phse[o];
            }
          }
        }
        __syncthreads();
      }
      if(pending) { tc::mbar_wait(bar, parity); parity ^= 1; pending = false; tc::fence_after_sync(); }

      // ---- A: rows of e^{i k theta}, 16 harmonics per thread, into tensor memory ----
      {
        uint32_t arh[16], arl[16], aih[16], ail[16];
#pragma unroll
        for(int i = 0; i < 8; i ++) {
          btc_split2(Wr, arh[2 * i], arh[2 * i + 1], arl[2 * i], arl[2 * i + 1]);
          btc_split2(Wi, aih[2 * i], aih[2 * i + 1], ail[2 * i], ail[2 * i + 1]);
          float2 t1 = fmul2(Wr, z2r), t2 = fmul2(Wi, z2r);
          float2 nWr = ffma2(Wi, nz2i, t1), nWi = ffma2(Wr, z2i, t2);
          Wr = nWr; Wi = nWi;
        }
        {   // skip the other half-chunk's 16 harmonics
          const float2 z16r = make_float2(z16.x, z16.x), z16i = make_float2(z16.y, z16.y), nz16i = make_float2(-z16.y, -z16.y);
          float2 t1 = fmul2(Wr, z16r), t2 = fmul2(Wi, z16r);
          float2 nWr = ffma2(Wi, nz16i, t1), nWi = ffma2(Wr, z16i, t2);
          Wr = nWr; Wi = nWi;
        }
        const uint32_t ta = tlane + 16 * hh;
        tc::tmem_st16(ta, arh); tc::tmem_st16(ta + 32, arl); tc::tmem_st16(ta + 64, aih); tc::tmem_st16(ta + 96, ail);
      }
      // ---- B: four harmonics of one (frame, q) column, K-major core matrices ----
      {
        const float2 e1 = cmul(w, rho);
        const float2 Er = make_float2(w.x, e1.x), Ei = make_float2(w.y, e1.y);
        const float2 r2r = make_float2(rho2.x, rho2.x), r2i = make_float2(rho2.y, rho2.y), nr2i = make_float2(-rho2.y, -rho2.y);
        const float2 Er2 = ffma2(Ei, nr2i, fmul2(Er, r2r)), Ei2 = ffma2(Er, r2i, fmul2(Ei, r2r));
        const float4 cr = *(const float4*)(Cr + gB * BTC_CST + (c & 3) * BTC_KC + 4 * jB);
        const float4 ci = *(const float4*)(Ci + gB * BTC_CST + (c & 3) * BTC_KC + 4 * jB);
        const float2 cr01 = make_float2(cr.x, cr.y), cr23 = make_float2(cr.z, cr.w);
        const float2 ci01 = make_float2(ci.x, ci.y), ci23 = make_float2(ci.z, ci.w);
        const float2 nci01 = make_float2(-ci.x, -ci.y), nci23 = make_float2(-ci.z, -ci.w);
        const float2 br01 = ffma2(Ei, nci01, fmul2(Er, cr01)), bi01 = ffma2(Ei, cr01, fmul2(Er, ci01));
        const float2 br23 = ffma2(Ei2, nci23, fmul2(Er2, cr23)), bi23 = ffma2(Ei2, cr23, fmul2(Er2, ci23));
        uint4 rh, rl, ih, il;
        btc_split2(br01, rh.x, rh.y, rl.x, rl.y); btc_split2(br23, rh.z, rh.w, rl.z, rl.w);
        btc_split2(bi01, ih.x, ih.y, il.x, il.y); btc_split2(bi23, ih.z, ih.w, il.z, il.w);
        float* d = bt + jB * 32 + gB * 256 + qB * 4;            // floats: k-quad 128 B, frame 1024 B, q 16 B
        *(uint4*)(d) = rh; *(uint4*)(d + 1024) = rl; *(uint4*)(d + 2048) = ih; *(uint4*)(d + 3072) = il;
        w = cmul(w, rho32);
      }
      tc::tmem_st_wait();
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
      // ---- MMA: four issuing warps, one accumulator each ----
      if(warp < 4 && lane == 0) {
        tc::fence_after_sync();
#pragma unroll
        for(int m = 0; m < 2; m ++) {
          tc::mma_tf32_ts(mm_d, mm_ah + 16 * m, mm_bh[m], mm_idesc, (c | m) ? 1u : 0u);
          tc::mma_tf32_ts(mm_d, mm_al + 16 * m, mm_bh[m], mm_idesc, 1u);
          tc::mma_tf32_ts(mm_d, mm_ah + 16 * m, mm_bl[m], mm_idesc, 1u);
        }
        tc::mma_commit(bar);
      }
      __syncwarp();
      pending = true;
    }

    // ---- epilogue: diagonal block of the four accumulators, window, park the frame ----
    tc::mbar_wait(bar, parity); parity ^= 1; pending = false;
    tc::fence_after_sync();
    {
      uint32_t u0[8], u1[8], v0[8], v1[8];
      const uint32_t td = tlane + 128 + 8 * qd;
      tc::tmem_ld8(td, u0); tc::tmem_ld8(td + 32, u1); tc::tmem_ld8(td + 64, v0); tc::tmem_ld8(td + 96, v1);
      tc::tmem_ld_wait();
      const int s = s0 + qd;
      if(sv[s] && !(hh == 1 && pp == 0)) {
        const int nb = hh == 0 ? 8 * pp : -8 * pp;
        float* dst = fb + (size_t)s * BTC_SLOT + 256 + nb;
        const float* wv = swin + 256 + nb;
        float o[8];
#pragma unroll
        for(int q = 0; q < 8; q ++) {
          const float U = __uint_as_float(u0[q]) + __uint_as_float(u1[q]);
          const float V = __uint_as_float(v0[q]) + __uint_as_float(v1[q]);
          o[q] = (hh == 0 ? U - V : U + V) * wv[q];
        }
        *(float4*)(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *(float4*)(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    tc::fence_before_sync();
  }
  __syncthreads();

  // ---- overlap-add, one warp per owned tile [base_f, base_f+1): the frames f - 1 .. f + 2 can reach it
  //      (H = round(hop), positions round(f hop)); ascending frame order as in layer0.c:135-140
  float* yrow = P.y_sin + (size_t)b * P.stride;
  for(int ti = warp; ti < F; ti += BTC_THREADS / 32) {
    const int f = t0 + ti, s = ti + 1;
    if(f >= nf) break;
    const int lo = f == 0 ? 0 : sb[s];
    const int hi = f + 1 < nf ? sb[s + 1] : P.nsamp;
    int off[4]; const float* src[4];
#pragma unroll
    for(int c = 0; c < 4; c ++) {
      const int sc = min(s - 1 + c, BTC_NSLOT - 1);
      const bool ok = (s - 1 + c < BTC_NSLOT) && sv[sc];
      off[c] = ok ? sb[sc] - H : (1 << 29);                   // j = idx - off; invalid slots fail j < N
      src[c] = fb + (size_t)sc * BTC_SLOT + 256 - H;
    }
    for(int idx = lo + lane; idx < hi; idx += 32) {
      float acc = 0.f;
      if(idx < ny_b) {
#pragma unroll
        for(int c = 0; c < 4; c ++) {
          const int j = idx - off[c];
          if((unsigned)j < (unsigned)N) acc += src[c][j];
        }
      }
      yrow[idx] = acc;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if(warp == 0) tc::tmem_dealloc(tbase, 256);
}

static inline size_t bank_tc_smem_bytes() {
  return (size_t)(BTC_NSLOT * BTC_SLOT + BTC_SLOT + 4 * 1024 + 8 * BTC_CST) * 4 + BTC_NSLOT * (sizeof(BtcFrame) + 8) + 16;
}

// tensor-core path when the window fits the 32 x 8 sample grid; returns -1 otherwise
static inline int launch_hm_bank_tc(const BankParams& P, int nutt, int nfrm_max, cudaStream_t st) {
  if((P.n_hm >> 1) > BTC_MAXH || (P.n_hm & 1)) return -1;
  const int F = BTC_NSLOT - 2;
  const int nseg = (std::max(nfrm_max, 1) + F - 1) / F;
  const size_t smem = bank_tc_smem_bytes();
  static bool attr = false;
  if(! attr) { cudaFuncSetAttribute(hm_bank_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  hm_bank_tc_kernel<<<dim3(nseg, nutt), dim3(BTC_THREADS), smem, st>>>(P);
  return 0;
}
#endif  // LLSM_EMU
