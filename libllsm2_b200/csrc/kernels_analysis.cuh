// Layer-0 analysis kernels for B200 (sm_100a).
//
//   refine_f0_kernel        llsm_refine_f0 (dsputils.c:72-94) + the IF detector it drives
//   harmonic_dft_kernel     llsm_harmonic_analysis, HMCZT method (dsputils.c:175-228,145-169):
//                           Blackman window, DFT at the harmonics of f0, centre rotation,
//                           amplitude 2/sum(w), phase. Also used on the squared sub-band signals
//                           (layer0.c:443) with maxnhar_e harmonics.
//   residual_kernel         x_res = x - x_sin (layer0.c:500-501)
//   (iir_filtfilt_kernel in kernels_iir.cuh: llsm_subband_energy, dsputils.c:230-235,51-70)
//   frame_dc_kernel         llsm_compute_dc (dsputils.c:117-124) with the window rule of
//                           layer0.c:430
//   noise_spec_kernel       per-frame spectra of llsm_analyze_noise_psd (layer0.c:325-360):
//                           Hann STFT -> cepstral envelope (spec2env) and Blackman PSD of x_res
//   noise_kalman_kernel     per-bin moving variance, Kalman filter + RTS smoother along time
//                           (layer0.c:361-385)
//   noise_psd_out_kernel    interpolation to npsd bins, dB conversion (layer0.c:388-408)
#pragma once
#include "common.cuh"
#include "kernels_synth.cuh"   // ChanFiltDev
#include "bulk_copy.cuh"

// reference index helpers, float/double steps as written in the C source ---------------------
__device__ __forceinline__ int ana_winsize(float fs, float f0, float rel) {
  float t = fs / f0;                 // dsputils.c:190  round(fs / f0[i] * rel_winsize / 2) * 2
  t = __fmul_rn(t, rel);
  t = t / 2.0f;
  return (int)round((double)t) * 2;
}
__device__ __forceinline__ int ana_nhar(float fs, float f0, int maxnhar) {
  float t = fs / f0;                 // dsputils.c:171-173  floor(fs / f0 / 2)
  t = t / 2.0f;
  int n = (int)floor((double)t);
  return n < maxnhar ? n : maxnhar;
}

// ------------------------------------------------------------------------------------------
// F0 refinement
// ------------------------------------------------------------------------------------------
struct RefineParams {
  int nfrm; const int* nfrm_utt;
  const float* x; int nx, xstride;
  const int* center;        // [nfrm] round(i * thop * fs)
  float fs;
  float* f0;                // [B][nfrm] in/out
};

// One WARP per frame (RF_WARPS frames per CTA) estimates the instantaneous frequency around harmonics 1..3 with the
// windowed complex-demodulation detector of the oracle's ciglet shim: Hann window of nh = 2 round(2 / fres) + 1 taps
// (fres = f0 / fs), f = fc - Im(yd / y) / (2 pi). The three detectors share their window (same fres), so a tap costs one
// window rotation and three demodulation rotations, all in double from one seed per lane (m = lane - half, + 32, ...);
// the taps are formed and accumulated in float (the detector stores them as FP_TYPE), ~40 terms per lane, and the
// lane sums are combined in double. (First version: a warp per harmonic with double taps and accumulators -- the
// kernel was bound by the double <-> float conversions, XU pipe 80 % busy. Tried and dropped: the window's cosine / sine
// from a per-plan table instead of the double rotation -- 1.76 ms against 1.43 ms at C2, the table reads add a
// dependent global load to every tap.)
#define RF_WARPS 4

__global__ void __launch_bounds__(32 * RF_WARPS) refine_f0_kernel(RefineParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * RF_WARPS + warp, b = blockIdx.y;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const float f0 = P.f0[(size_t)b * P.nfrm + i];
  if(f0 == 0) return;
  const float fres = f0 / P.fs;
  int half = (int)round(2.0 / (double)fres);
  if(half < 2) half = 2;
  const int nh = 2 * half + 1;
  const double L = 2.0 * half + 2.0;
  const float* x = P.x + (size_t)b * P.xstride;
  const int center = P.center[i];
  float fc[3];
#pragma unroll
  for(int j = 0; j < 3; j ++) fc[j] = fres * (float)(j + 1);          // f0[i] / fs * j (dsputils.c:79)
  float yr[3] = {0.f, 0.f, 0.f}, yi[3] = {0.f, 0.f, 0.f}, dr[3] = {0.f, 0.f, 0.f}, di[3] = {0.f, 0.f, 0.f};
  {
    // The taps m and -m share everything but the sample: w(-m) = w(m), w'(-m) = -w'(m), e^{-i th (-m)} = conj. With
    // e = x(m) + x(-m), o = x(m) - x(-m):  y += e w cos - i o w sin,  yd += o w' cos - i e w' sin  -- one rotation set and
    // one set of double -> float conversions per PAIR of taps (the kernel is bound by exactly those: XU pipe 64 %,
    // FP64 39 % with a tap per iteration). Lane l takes m = l + 1, l + 33, ...; the centre tap (w = 1, w' = 0) is lane 0's.
    const int m0 = lane + 1;
    double sw, cw, sws, cws, sp[3], cp[3], sps[3], cps[3];
    sincospi(2.0 * (double)m0 / L, &sw, &cw);
    sincospi(2.0 * 32.0 / L, &sws, &cws);
#pragma unroll
    for(int j = 0; j < 3; j ++) {
      double u = (double)fc[j] * (double)m0; u -= rint(u);
      sincospi(2.0 * u, &sp[j], &cp[j]);
      double us = (double)fc[j] * 32.0; us -= rint(us);
      sincospi(2.0 * us, &sps[j], &cps[j]);
    }
    const float wdk = (float)(-0.5 * (2.0 * LLSM_PI / L));
    if(lane == 0 && center >= 0 && center < P.nx) {
      const float x0 = x[center];
#pragma unroll
      for(int j = 0; j < 3; j ++) yr[j] = x0;
    }
    int ip = center + m0, im = center - m0;
    float xpn = (m0 <= half && ip >= 0 && ip < P.nx) ? x[ip] : 0.f;
    float xmn = (m0 <= half && im >= 0 && im < P.nx) ? x[im] : 0.f;
    for(int m = m0; m <= half; m += 32) {
      const float xp = xpn, xm = xmn;
      ip += 32; im -= 32;
      xpn = (m + 32 <= half && ip >= 0 && ip < P.nx) ? x[ip] : 0.f;   // the next pair's samples, one step ahead of their use
      xmn = (m + 32 <= half && im >= 0 && im < P.nx) ? x[im] : 0.f;
      const float w = 0.5f + 0.5f * (float)cw, wd = wdk * (float)sw;
      const float e = xp + xm, o = xp - xm;
      const float ew = e * w, ow = o * w, ed = e * wd, od = o * wd;
#pragma unroll
      for(int j = 0; j < 3; j ++) {
        const float c = (float)cp[j], s = (float)sp[j];
        yr[j] = fmaf(ew, c, yr[j]); yi[j] = fmaf(-ow, s, yi[j]);
        dr[j] = fmaf(od, c, dr[j]); di[j] = fmaf(-ed, s, di[j]);
        const double t2 = cp[j] * cps[j] - sp[j] * sps[j]; sp[j] = sp[j] * cps[j] + cp[j] * sps[j]; cp[j] = t2;
      }
      const double t1 = cw * cws - sw * sws; sw = sw * cws + cw * sws; cw = t1;
    }
  }
  // lane sums -> warp sums in double; lane j then finishes harmonic j + 1
  double myr = 0, myi = 0, mdr = 0, mdi = 0;
#pragma unroll
  for(int j = 0; j < 3; j ++) {
    double a = (double)yr[j], bq = (double)yi[j], c = (double)dr[j], d = (double)di[j];
    for(int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); bq += __shfl_xor_sync(0xffffffffu, bq, o);
      c += __shfl_xor_sync(0xffffffffu, c, o); d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if(lane == j) { myr = a; myi = bq; mdr = c; mdi = d; }
  }
  float fj = 0.f; int ok = 0;
  if(lane < 3) {
    const float fcl = fres * (float)(lane + 1);
    const double den = myr * myr + myi * myi;
    float est = fcl;
    if(! (den < 1e-30)) est = (float)((double)fcl - ((mdi * myr - mdr * myi) / den) / (2.0 * LLSM_PI));
    fj = est / (float)(lane + 1);                         // dsputils.c:81
    const float diff = fj - fres;
    ok = fabs((double)diff) < (double)f0 * 0.1 / (double)P.fs;       // dsputils.c:82
  }
  float favg = 0; int n = 0;
#pragma unroll
  for(int q = 0; q < 3; q ++) {                           // same order as the reference's loop over harmonics
    const float fq = __shfl_sync(0xffffffffu, fj, q); const int oq = __shfl_sync(0xffffffffu, ok, q);
    if(oq) { favg += fq; n ++; }
  }
  if(lane == 0 && n > 0) {
    favg = favg / (float)n;
    P.f0[(size_t)b * P.nfrm + i] = favg * P.fs;           // dsputils.c:89-92
  }
}

// ------------------------------------------------------------------------------------------
// Harmonic estimator (HMCZT)
// ------------------------------------------------------------------------------------------
struct HarmDftParams {
  int nfrm; const int* nfrm_utt;
  const float* sig;         // [B][nsig][xstride]
  int nsig, nx, xstride;
  const float* f0;          // [B][nfrm]
  const int* center;        // [nfrm]
  float fs, rel_winsize;
  int maxnhar;              // harmonics to estimate (row length of the outputs)
  int out_stride;           // floats between consecutive frames in ampl/phse ([.., nsig, maxnhar])
  int* nhar_out;            // [B][nfrm][nsig]
  float* ampl; float* phse; // [B][nfrm][nsig][maxnhar]
  int max_half;             // capacity of the staged half-frame (pairs)
  float* edc; float thop;   // optional: short-time mean of every signal, [B][nfrm][nsig]
  int hop_max;              // largest distance between consecutive frame centres (staged kernels)
  int mma_cap_half;         // capacity (half window) of the tensor-core kernel's staging
  // window table (AnaPlan::bwin): w(h +- n) at bwin[bw_off[h] + n], sums bw_sum[h], for half windows h <= bw_cap
  const float* bwin; const int* bw_off; const float* bw_sum; int bw_cap;
  int stage_cap;            // half-window capacity of the staged kernels' shared-memory slices
  int seg_frames, seg_len;  // staged kernels: frames per CTA, samples per staged row (multiple of 4)
  int nutt;                 // utterances of the batch
  int* long_list;           // staged kernels: frames left to the general kernels ([0] = count, [1 ..] = b * nfrm + i);
                            // general kernels: when non-NULL, serve exactly these frames
};


// Amplitude and phase of one harmonic from its DFT sum (re, im), as llsm_harmonic_czt finishes it (dsputils.c:158-166):
// the sum was rotated to the window centre with the exact shift (k + 1) omega0 half; the reference rotates by the
// FLOAT-rounded ishift = (float)(shift 2 pi f0 / fs (k + 1)) instead, so the tiny difference eps (|eps| ~ 1e-4 rad: one
// float rounding of a ~1e3 rad angle) is applied here -- by its Taylor series, exact to 1e-17 for such arguments,
// instead of double-precision sin / cos calls. Amplitude = |.| 2 / sum(w).
__device__ __forceinline__ void harmonic_finish(float re, float im, int k, int half, float f0, float fs, float omega0,
  float winsum, float* ampl, float* phse) {
  const float ishift = (float)((double)half * 2.0 * LLSM_PI * (double)f0 / (double)fs * ((double)k + 1.0));
  const double eps = (double)ishift - (double)(k + 1) * (double)omega0 * (double)half;
  const double e2 = eps * eps;
  const float se = (float)(eps * (1.0 - e2 * (1.0 / 6.0))), ce = (float)(1.0 - 0.5 * e2);
  const float dre = re * ce - im * se, dim = re * se + im * ce;
  *ampl = sqrtf(dre * dre + dim * dim) * (2.0f / winsum);
  *phse = atan2f(dim, dre);
}

#define HD_THREADS 128
#define HD_RESEED 64
#define HD_KW 8           // harmonics per signal in the warp-per-signal variant

// G = threads that share one signal: HD_THREADS (one signal per CTA: the main pass) or 32 (one warp per
// signal, HD_THREADS / 32 signals of the same frame per CTA: the sub-band envelope pass, where the window,
// its sum and the frame geometry are shared by the channels). With P.edc != NULL the group also writes
// the short-time mean of its signal (llsm_compute_dc, dsputils.c:117-124; window rule layer0.c:430,446).
// KW = harmonics the warp-per-signal variant carries in registers (its loops are unrolled to KW: with the usual four
// envelope harmonics the eight-wide instance spent half its instructions on predicated-off harmonics).
template <int G, int KW = HD_KW>
__device__ __forceinline__ void harmonic_dft_frame(const HarmDftParams& P, const int i, const int by, char* smem) {
  constexpr int NG = HD_THREADS / G;                  // signals per CTA
  float* wv = (float*)smem;                           // [max_half + 2] window, w(half +- n)
  double* red = (double*)(wv + ((P.max_half + 2 + 1) & ~1));   // [HD_THREADS] reduction scratch
  float* part = (float*)(red + HD_THREADS);           // [2 * HD_THREADS] slice partials
  float2* spall = (float2*)(part + 2 * HD_THREADS);   // [NG][max_half + 2] (x+ + x-, x+ - x-)

  const int ngrp = (P.nsig + NG - 1) / NG;            // CTAs per (utterance, frame)
  const int b = by / ngrp;
  const int tid = threadIdx.x;
  const int g = tid / G, gt = tid % G;                // group, thread in group
  const int c = (by % ngrp) * NG + g;                 // signal of this group
  const bool live = c < P.nsig;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const size_t fidx = ((size_t)b * P.nfrm + i) * P.nsig + (live ? c : 0);
  if(i >= nf) return;
  const float f0 = P.f0[(size_t)b * P.nfrm + i];
  const int center = P.center[i];
  const float* x = P.sig + ((size_t)b * P.nsig + (live ? c : 0)) * P.xstride;
  auto gsync = [&]() { if(G == 32) __syncwarp(); else __syncthreads(); };

  // ---- short-time mean (sub-band pass): round((f0 == 0 ? thop * 2 : 2 / f0) * fs) samples around the centre
  if(P.edc != nullptr && live) {
    double wlen = f0 == 0 ? (double)(P.thop * 2.0f) : 2.0 / (double)f0;
    const int nw = (int)round(wlen * (double)P.fs);
    double acc = 0;
    for(int j = gt; j < nw; j += 4 * G) {               // four reads in flight, summed in the same order
      float v[4];
#pragma unroll
      for(int u = 0; u < 4; u ++) {
        const int idx = center + j + u * G - nw / 2;
        v[u] = (j + u * G < nw && idx >= 0 && idx < P.nx) ? x[idx] : 0.f;
      }
#pragma unroll
      for(int u = 0; u < 4; u ++) acc += (double)v[u];
    }
    if(G == 32) {
      for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    } else {
      red[tid] = acc;
      __syncthreads();
      for(int o = G >> 1; o > 0; o >>= 1) { if(tid < o) red[tid] += red[tid + o]; __syncthreads(); }
      acc = red[0];
      __syncthreads();
    }
    if(gt == 0) P.edc[fidx] = nw > 0 ? (float)(acc / nw) : 0.f;
  }

  if(! (f0 > 0)) {                                    // unvoiced: no harmonic model (layer0.c:106)
    if(live) {
      if(gt == 0) P.nhar_out[fidx] = 0;
      for(int k = gt; k < P.maxnhar; k += G) { P.ampl[fidx * P.maxnhar + k] = 0; P.phse[fidx * P.maxnhar + k] = 0; }
    }
    return;
  }
  const int ws = ana_winsize(P.fs, f0, P.rel_winsize);
  const int nh = ana_nhar(P.fs, f0, P.maxnhar);
  const int half = ws >> 1;                           // shift = nx / 2 (dsputils.c:151)
  if(half > P.max_half) {                             // window longer than the staging buffer
    if(live && gt == 0) P.nhar_out[fidx] = -1;
    return;
  }

  // ---- Blackman window once per CTA. ws is even, so the periodic window is symmetric about m = half:
  //   w(half +- n) = 0.42 + 0.5 cos(2 pi n / ws) + 0.08 cos(4 pi n / ws).
  // Each thread walks n = tid, tid + T, ... advancing cos / sin (2 pi n / ws) by a fixed rotation in double.
  double wsum = 0;
  {
    double cs, sn, cstep, sstep;
    sincospi(2.0 * (double)tid / (double)ws, &sn, &cs);
    sincospi(2.0 * (double)HD_THREADS / (double)ws, &sstep, &cstep);
    for(int n = tid; n <= half; n += HD_THREADS) {
      const float w = (float)(0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0));
      wv[n] = w;
      if(n < half) wsum += w;                         // m = half + n
      if(n >= 1) wsum += w;                           // m = half - n
      const double c2 = cs * cstep - sn * sstep;
      sn = sn * cstep + cs * sstep; cs = c2;
    }
  }
  red[tid] = wsum;
  __syncthreads();
  for(int o = HD_THREADS >> 1; o > 0; o >>= 1) {
    if(tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  const float winsum = (float)red[0];                 // FP_TYPE winsum = sumfp(w, nx)
  if(! live) return;                                  // (no block barrier below for G == 32)

  // ---- the windowed frame as symmetric / antisymmetric halves: staged in shared memory when the CTA
  //      serves one signal; read straight from global memory (L1/L2-resident, a handful of harmonics) by
  //      the warp-per-signal variant, whose shared memory only holds the window
  auto pair_at = [&](int n) -> float2 {
    const float w = wv[n];
    float xp = 0, xm = 0;
    if(n < half) { int idx = center + n; if(idx >= 0 && idx < P.nx) xp = w * x[idx]; }
    if(n >= 1) { int idx = center - n; if(idx >= 0 && idx < P.nx) xm = w * x[idx]; }
    return make_float2(xp + xm, xp - xm);
  };
  float2* sp = spall;
  if(G != 32) {
    for(int n = gt; n <= half; n += G) sp[n] = pair_at(n);
    gsync();
  }

  // ---- frequencies as the reference rounds them (dsputils.c:156-158)
  const float omega0 = (float)(2.0 * LLSM_PI * (double)f0 / (double)P.fs);   // czt step (FP_TYPE arg)
  const double nu = (double)omega0 / (2.0 * LLSM_PI);                        // turns per sample

  const int npair = half + 1;
  auto finish = [&](int k, float re, float im) {
    harmonic_finish(re, im, k, half, f0, P.fs, omega0, winsum, &P.ampl[fidx * P.maxnhar + k], &P.phse[fidx * P.maxnhar + k]);
  };

  if(G == 32) {
    // ---- a warp per signal, at most HD_KW harmonics: every lane walks its own samples (coalesced reads)
    //      carrying one phasor per harmonic, advanced 32 samples at a time and re-seeded every 8 steps
    float2 w[KW], z32[KW]; float re[KW], im[KW];
#pragma unroll
    for(int k = 0; k < KW; k ++) {
      re[k] = 0.f; im[k] = 0.f;
      z32[k] = k < nh ? unit_phasor_turns((double)(k + 1) * nu * 32.0) : make_float2(1.f, 0.f);
    }
    int step = 0;
    float2 sv_next = gt < npair ? pair_at(gt) : make_float2(0.f, 0.f);
    for(int n = gt; n < npair; n += 32, step ++) {
      const float2 sv = sv_next;                          // read one step ahead of its use
      if(n + 32 < npair) sv_next = pair_at(n + 32);
      if((step & 7) == 0) {
#pragma unroll
        for(int k = 0; k < KW; k ++) if(k < nh) w[k] = unit_phasor_turns((double)(k + 1) * nu * (double)n);
      }
#pragma unroll
      for(int k = 0; k < KW; k ++) if(k < nh) {
        re[k] = fmaf(sv.x, w[k].x, re[k]);
        im[k] = fmaf(-sv.y, w[k].y, im[k]);
        w[k] = cmul(w[k], z32[k]);
      }
    }
    float myre = 0.f, myim = 0.f;
#pragma unroll
    for(int k = 0; k < KW; k ++) {
      float r = re[k], q = im[k];
      for(int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if(gt == k) { myre = r; myim = q; }
    }
    if(gt < nh) finish(gt, myre, myim);
  } else {
    // work split: nhc harmonics across threads, nsl sample slices across the remaining factor
    int nhc = 1; while(nhc < nh && nhc < G) nhc <<= 1;
    const int nslp = G / nhc;                           // parallel slices
    const int kk = gt % nhc, sl0 = gt / nhc;
    const int nslice = (npair + HD_RESEED - 1) / HD_RESEED;
    float* gpart = part + 2 * g * G;

    for(int k0 = 0; k0 < nh; k0 += nhc) {
      const int k = k0 + kk;                            // harmonic index (0-based), frequency (k+1) f0
      float re = 0.f, im = 0.f;
      if(k < nh) {
        const double th = (double)(k + 1) * nu;
        const float2 z = unit_phasor_turns(th);
        for(int sl = sl0; sl < nslice; sl += nslp) {
          const int n0 = sl * HD_RESEED, n1 = min(n0 + HD_RESEED, npair);
          float2 w = unit_phasor_turns(th * (double)n0);
          for(int n = n0; n < n1; n ++) {
            const float2 sv = G != 32 ? sp[n] : pair_at(n);
            re = fmaf(sv.x, w.x, re);                   // sum (x+ + x-) cos
            im = fmaf(-sv.y, w.y, im);                  // -sum (x+ - x-) sin
            w = cmul(w, z);
          }
        }
      }
      if(nslp > 1) {                                    // deterministic cross-slice reduction
        gsync();
        gpart[2 * gt] = re; gpart[2 * gt + 1] = im;
        gsync();
        if(sl0 == 0) {
          for(int q = 1; q < nslp; q ++) { re += gpart[2 * (q * nhc + kk)]; im += gpart[2 * (q * nhc + kk) + 1]; }
        }
      }
      if(sl0 == 0 && k < nh) {
        finish(k, re, im);
      }
    }
  }
  for(int k = nh + gt; k < P.maxnhar; k += G) { P.ampl[fidx * P.maxnhar + k] = 0; P.phse[fidx * P.maxnhar + k] = 0; }
  if(gt == 0) P.nhar_out[fidx] = nh;
}

// grid (nfrm, nutt * groups): one CTA per frame and signal group; or, with P.long_list (one signal per utterance), a
// fixed grid walking the listed frames
template <int G, int KW = HD_KW>
__global__ void __launch_bounds__(HD_THREADS) harmonic_dft_kernel(HarmDftParams P) {
  LLSM_DYN_SMEM(smem);
  if(P.long_list == nullptr) { harmonic_dft_frame<G, KW>(P, blockIdx.x, blockIdx.y, smem); return; }
  const int n = P.long_list[0];
  for(int e = blockIdx.x; e < n; e += gridDim.x) {
    const int f = P.long_list[1 + e];
    harmonic_dft_frame<G, KW>(P, f % P.nfrm, f / P.nfrm, smem);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Tensor-core variant of the main pass (one signal, up to maxnhar harmonics per frame).
// With the pair index n = 16 p + q the sums over the window factor into a small GEMM per frame:
//   Re X_k =  sum_q [ cos(k w q) Ecc(k, q) - sin(k w q) Ecs(k, q) ],
//   Im X_k = -sum_q [ cos(k w q) Ocs(k, q) + sin(k w q) Occ(k, q) ],
//   E{c,s}(k, q) = sum_p e[16 p + q] {cos, sin}(16 k w p)   (same with o),
// i.e. D[2K x 32] = A[2K x P] * B[P x 32]: A holds the stride-16 phasors (generated in the fragment registers by
// rotation -- no operand staging, which is what a tcgen05 form of this contraction would have to pay: both operands
// depend on the frame), B the windowed signal halves e | o (staged once in shared memory). Warp-level FP16 tensor-core
// instructions (m16n8k16) with FP32 accumulation; every operand is split x = hi + lo / 2048 (f16_split) and the product
// taken as hi hi + (hi lo + lo hi) / 2048 in two accumulator sets: 22 bits, the accuracy of the 3xTF32 form this kernel
// used before at half its tensor-pipe cycles (that form ran into the pipe: math-pipe throttle was its first stall reason,
// profiles/r2k). One warp owns tiles of 8 harmonics (16 rows: 8 cosine + 8 sine); the q-phasors and the 16-term q-sum
// are the FP32 epilogue.
// ------------------------------------------------------------------------------------------
#define HM_THREADS 128
#define HM_QS 18                                      // uint4 row stride of a staged p-pair row (16 used): conflict-free LDS.128

// A CTA owns P.seg_frames consecutive frames of one utterance (P.seg_frames == 1: one frame, samples read from global
// memory); for longer segments thread 0 stages the waveform slice those frames touch with one bulk asynchronous copy
// (cp.async.bulk, as envelope_seg_kernel does): neighbouring windows overlap by three quarters, and the frames' split /
// pack passes then read shared memory instead of waiting on L2.
__global__ void __launch_bounds__(HM_THREADS) harmonic_mma_kernel(HarmDftParams P) {
  LLSM_DYN_SMEM(smem);
  const int cap = P.mma_cap_half;
  const int prow = (((cap + 1 + 15) / 16) + 15) & ~15;  // staged p-rows (multiple of 16)
  bulk_bar_t* bar = (bulk_bar_t*)smem;
  uint4* sB = (uint4*)(smem + 16);                    // [prow / 2][HM_QS]: {e_hi, e_lo, o_hi, o_lo} of rows (2 p2, 2 p2 + 1), column q
  float2* sums = (float2*)(sB + (size_t)(prow / 2) * HM_QS);   // [maxnhar] DFT sums of the harmonics
  float* slice = (float*)(sums + P.maxnhar);          // [P.seg_len] staged waveform (segments only)

  const int b = blockIdx.y, i_lo = blockIdx.x * P.seg_frames;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int tid = threadIdx.x;
  if(i_lo >= nf) return;
  const int i_hi = min(i_lo + P.seg_frames, nf);
  const float* xg = P.sig + (size_t)b * P.xstride;
  const bool staged = P.seg_frames > 1;
  int lo = 0;
  if(staged) {
    int hi = P.center[i_hi - 1] + cap + 1;
    lo = max(P.center[i_lo] - cap, 0) & ~3; hi = (min(hi, P.nx) + 3) & ~3;   // (rows are padded to a multiple of four samples)
    const int len = min(max(hi - lo, 0), P.seg_len);
    if(tid == 0) bulk_bar_init(bar);
    __syncthreads();
    if(tid == 0) {
      bulk_expect(bar, (uint32_t)(len * 4));
      if(len > 0) bulk_g2s(slice, xg + lo, (uint32_t)(len * 4), bar);
    }
    bulk_wait(bar, 0);
  }
  const float* x = staged ? slice - lo : xg;          // x[ix], ix = sample index in the utterance

  for(int i = i_lo; i < i_hi; i ++) {
  const size_t fidx = (size_t)b * P.nfrm + i;
  const float f0 = P.f0[fidx];
  if(! (f0 > 0)) {                                    // unvoiced: no harmonic model (layer0.c:106)
    if(tid == 0) P.nhar_out[fidx] = 0;
    for(int k = tid; k < P.maxnhar; k += HM_THREADS) { P.ampl[fidx * P.maxnhar + k] = 0; P.phse[fidx * P.maxnhar + k] = 0; }
    continue;
  }
  const int ws = ana_winsize(P.fs, f0, P.rel_winsize);
  const int nh = ana_nhar(P.fs, f0, P.maxnhar);
  const int half = ws >> 1;
  if(half > cap || half < 1) {                        // left to the general kernel (harmonic_dft_kernel in list mode)
    if(tid == 0) { const int at = atomicAdd(P.long_list, 1); P.long_list[1 + at] = (int)fidx; }
    continue;
  }
  const int center = P.center[i];
  // ---- Blackman window and its sum from the plan's table (w(half +- n) at wv[n])
  const float* wv = P.bwin + P.bw_off[half];
  const float winsum = P.bw_sum[half];

  // ---- stage e | o (symmetric / antisymmetric halves of the windowed frame), split, as FP16 pairs over p
  const int npair = half + 1;
  const int np16 = (((npair + 15) / 16) + 15) & ~15;  // p-rows in use (multiple of 16, <= prow)
  __syncthreads();                                    // the previous frame's operand and sums are no longer read
  for(int idx = tid; idx < (np16 / 2) * 16; idx += HM_THREADS) {
    const int p2 = idx >> 4, q = idx & 15;
    float e[2], o[2];
#pragma unroll
    for(int u = 0; u < 2; u ++) {
      const int n = 32 * p2 + 16 * u + q;
      float w = 0.f, xp = 0.f, xm = 0.f;
      if(n < npair) {
        w = wv[n];
        if(n < half) { const int ix = center + n; if(ix >= 0 && ix < P.nx) xp = x[ix]; }
        if(n >= 1) { const int ix = center - n; if(ix >= 0 && ix < P.nx) xm = x[ix]; }
      }
      const float a = w * xp, d = w * xm;
      e[u] = a + d; o[u] = a - d;
    }
    float eh[2], el[2], oh[2], ol[2];
#pragma unroll
    for(int u = 0; u < 2; u ++) { f16_split(e[u], eh[u], el[u]); f16_split(o[u], oh[u], ol[u]); }
    sB[p2 * HM_QS + q] = make_uint4(pack_f16x2(eh[0], eh[1]), pack_f16x2(el[0], el[1]), pack_f16x2(oh[0], oh[1]), pack_f16x2(ol[0], ol[1]));
  }
  __syncthreads();

  const float omega0 = (float)(2.0 * LLSM_PI * (double)f0 / (double)P.fs);   // czt step (FP_TYPE arg)
  const double nu = (double)omega0 / (2.0 * LLSM_PI);                        // turns per sample
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nsteps = np16 >> 4;

  for(int mt = warp; mt * 8 < nh; mt += HM_THREADS / 32) {
    const int kg = mt * 8 + g;                        // this lane's harmonic (rows g: cosine, g + 8: sine)
    const double th = (double)(kg + 1) * nu;
    // ---- phasors of this harmonic, b = e^{i 2 pi th}: two seeds with the angle reduced in double (b, b^16), the rest
    //      by short float product chains (at most five roundings, ~3e-7): b^{2t}, b^{2t+1}, b^{2t+8}, b^{2t+9} for the
    //      epilogue; b^{32t} for the first fragment rows, b^16 and b^128 to reach their neighbours, b^256 to advance.
    const float2 b1 = unit_phasor_turns(th), b16 = unit_phasor_turns(th * 16.0);
    const float2 b2 = cmul(b1, b1), b4 = cmul(b2, b2), b6 = cmul(b4, b2), b8 = cmul(b4, b4);
    const float2 one = make_float2(1.f, 0.f);
    const float2 q00 = t == 0 ? one : (t == 1 ? b2 : (t == 2 ? b4 : b6));
    const float2 q01 = cmul(q00, b1), q10 = cmul(q00, b8), q11 = cmul(q01, b8);
    const float2 b32 = cmul(b16, b16), b64 = cmul(b32, b32), b96 = cmul(b64, b32), b128 = cmul(b64, b64), b256 = cmul(b128, b128);
    float2 u0 = t == 0 ? one : (t == 1 ? b32 : (t == 2 ? b64 : b96));        // p-row 16 s + 2 t
    float d1[4][4], d2[4][4];
#pragma unroll
    for(int j = 0; j < 4; j ++)
#pragma unroll
      for(int c = 0; c < 4; c ++) { d1[j][c] = 0.f; d2[j][c] = 0.f; }
    for(int s = 0; s < nsteps; s ++) {
      if(s > 0 && (s & 3) == 0) u0 = unit_phasor_turns(th * 16.0 * (double)(16 * s + 2 * t));   // re-seed every 64 p-rows
      const float2 u1 = cmul(u0, b16), u2 = cmul(u0, b128), u3 = cmul(u2, b16);   // rows + 1, + 8, + 9
      float ch[4], cl[4], sh[4], sl[4];
      f16_split(u0.x, ch[0], cl[0]); f16_split(u1.x, ch[1], cl[1]); f16_split(u2.x, ch[2], cl[2]); f16_split(u3.x, ch[3], cl[3]);
      f16_split(u0.y, sh[0], sl[0]); f16_split(u1.y, sh[1], sl[1]); f16_split(u2.y, sh[2], sl[2]); f16_split(u3.y, sh[3], sl[3]);
      const uint32_t ah[4] = {pack_f16x2(ch[0], ch[1]), pack_f16x2(sh[0], sh[1]), pack_f16x2(ch[2], ch[3]), pack_f16x2(sh[2], sh[3])};
      const uint32_t al[4] = {pack_f16x2(cl[0], cl[1]), pack_f16x2(sl[0], sl[1]), pack_f16x2(cl[2], cl[3]), pack_f16x2(sl[2], sl[3])};
      const int r0 = (8 * s + t) * HM_QS + g, r1 = (8 * s + t + 4) * HM_QS + g;
#pragma unroll
      for(int j = 0; j < 2; j ++) {
        const uint4 v0 = sB[r0 + 8 * j], v1 = sB[r1 + 8 * j];
        const uint32_t beh[2] = {v0.x, v1.x}, bel[2] = {v0.y, v1.y}, boh[2] = {v0.z, v1.z}, bol[2] = {v0.w, v1.w};
        mma_f16_16x8x16(d1[j], ah, beh); mma_f16_16x8x16(d2[j], ah, bel); mma_f16_16x8x16(d2[j], al, beh);
        mma_f16_16x8x16(d1[j + 2], ah, boh); mma_f16_16x8x16(d2[j + 2], ah, bol); mma_f16_16x8x16(d2[j + 2], al, boh);
      }
      u0 = cmul(u0, b256);
    }
    // ---- epilogue: q-phasors and the sum over the 16 columns (4 per lane, then across the 4 lanes of a row)
    float re = 0.f, im = 0.f;
    const float2 wqv[2][2] = {{q00, q01}, {q10, q11}};
    const float rs = 1.0f / F16_LO_SCALE;
#pragma unroll
    for(int j = 0; j < 2; j ++)
#pragma unroll
      for(int e = 0; e < 2; e ++) {
        const float2 wq = wqv[j][e];                  // e^{i th (8 j + 2 t + e)}
        const float ecc = fmaf(d2[j][e], rs, d1[j][e]), ecs = fmaf(d2[j][2 + e], rs, d1[j][2 + e]);
        const float occ = fmaf(d2[j + 2][e], rs, d1[j + 2][e]), ocs = fmaf(d2[j + 2][2 + e], rs, d1[j + 2][2 + e]);
        re = fmaf(wq.x, ecc, fmaf(-wq.y, ecs, re));
        im = fmaf(-wq.x, ocs, fmaf(-wq.y, occ, im));
      }
    re += __shfl_xor_sync(0xffffffffu, re, 1); im += __shfl_xor_sync(0xffffffffu, im, 1);
    re += __shfl_xor_sync(0xffffffffu, re, 2); im += __shfl_xor_sync(0xffffffffu, im, 2);
    if(t == 0 && kg < nh) sums[kg] = make_float2(re, im);   // parked: the finish runs with every lane busy below
  }
  __syncthreads();
  for(int k = tid; k < nh; k += HM_THREADS) {
    const float2 v = sums[k];
    harmonic_finish(v.x, v.y, k, half, f0, P.fs, omega0, winsum, &P.ampl[fidx * P.maxnhar + k], &P.phse[fidx * P.maxnhar + k]);
  }
  for(int k = nh + tid; k < P.maxnhar; k += HM_THREADS) { P.ampl[fidx * P.maxnhar + k] = 0; P.phse[fidx * P.maxnhar + k] = 0; }
  if(tid == 0) P.nhar_out[fidx] = nh;
  }
}

static inline size_t harm_mma_smem(int cap, int maxnhar, int seg_len) {
  const int prow = (((cap + 1 + 15) / 16) + 15) & ~15;
  return 16 + (size_t)(prow / 2) * HM_QS * 16 + (size_t)maxnhar * 8 + (size_t)seg_len * 4 + 16;
}

static inline double mma_min_f0() {
  static double v = -1;
  if(v < 0) { const char* e = getenv("LLSM_MMA_MIN_F0"); v = e ? atof(e) : 80.0; if(! (v >= 20.0)) v = 20.0; }
  return v;
}

static inline int env_variant() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_ENV_VARIANT"); v = e ? atoi(e) : 1; }
  return v;
}

// CTAs of a persistent grid: per_sm resident CTAs on every SM of the current device
static inline int ana_resident_ctas(int per_sm) {
#ifdef LLSM_EMU
  return 3 * per_sm;
#else
  int dev = 0, sms = 148;
  if(cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * per_sm;
#endif
}

static inline int mma_seg_frames() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_MMA_SEG"); v = e ? atoi(e) : 8; if(v < 1) v = 1; }
  return v;
}
static inline int env_seg_frames() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_ENV_SEG"); v = e ? atoi(e) : 16; if(v < 1) v = 1; }
  return v;
}

static inline int dft_variant() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_DFT_VARIANT"); v = e ? atoi(e) : 1; }
  return v;
}

// ------------------------------------------------------------------------------------------
// Envelope pass (layer0.c:443-446): the harmonic analysis of the nchannel squared sub-band signals of one frame, at
// most 8 harmonics each, plus their short-time means. The channels share f0, the window and every phasor, so ONE CTA
// serves the frame: thread t walks the sample pairs n = t, t + 128, ... carrying one phasor per harmonic (advanced 128
// samples per step, re-seeded every 8 steps) and accumulates all channels x harmonics in registers -- a phasor update is
// paid once per sample instead of once per sample and channel; the per-thread sums are reduced through shared memory in
// a fixed order. NC = channel capacity, KW = harmonic capacity of the instance.
// ------------------------------------------------------------------------------------------
#define ED_THREADS 128
#define ED_STRIDE (ED_THREADS + 4)          // row stride of the reduction scratch: conflict-free for both access patterns

template <int NC, int KW>
__device__ __forceinline__ void envelope_dft_frame(const HarmDftParams& P, const int i, const int b, char* smem) {
  double* red = (double*)smem;                                       // [ED_THREADS]
  float* part = (float*)(red + ED_THREADS);                          // [2 * NC * KW][ED_STRIDE]
  float2* zst = (float2*)(part + 2 * NC * KW * ED_STRIDE);           // [KW] 128-sample rotation per harmonic
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const int nsig = P.nsig;
  const size_t f0idx = (size_t)b * P.nfrm + i;
  const float f0 = P.f0[f0idx];
  const int center = P.center[i];
  const float* x0 = P.sig + (size_t)b * nsig * P.xstride;

  // ---- short-time means (llsm_compute_dc, dsputils.c:117-124): one warp per channel
  if(P.edc != nullptr) {
    double wlen = f0 == 0 ? (double)(P.thop * 2.0f) : 2.0 / (double)f0;
    const int nw = (int)round(wlen * (double)P.fs);
    for(int c = warp; c < nsig; c += ED_THREADS / 32) {
      const float* x = x0 + (size_t)c * P.xstride;
      double acc = 0;
      for(int j = lane; j < nw; j += 4 * 32) {                       // four reads in flight, summed in the same order
        float v[4];
#pragma unroll
        for(int u = 0; u < 4; u ++) {
          const int idx = center + j + u * 32 - nw / 2;
          v[u] = (j + u * 32 < nw && idx >= 0 && idx < P.nx) ? x[idx] : 0.f;
        }
#pragma unroll
        for(int u = 0; u < 4; u ++) acc += (double)v[u];
      }
      for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if(lane == 0) P.edc[f0idx * nsig + c] = nw > 0 ? (float)(acc / nw) : 0.f;
    }
  }
  if(! (f0 > 0)) {                                                   // unvoiced: no harmonic model (layer0.c:106)
    for(int e = tid; e < nsig * P.maxnhar; e += ED_THREADS) { P.ampl[f0idx * nsig * P.maxnhar + e] = 0; P.phse[f0idx * nsig * P.maxnhar + e] = 0; }
    if(tid < nsig) P.nhar_out[f0idx * nsig + tid] = 0;
    return;
  }
  const int ws = ana_winsize(P.fs, f0, P.rel_winsize);
  const int nh = ana_nhar(P.fs, f0, P.maxnhar);
  const int half = ws >> 1;
  if(half > P.max_half) { if(tid < nsig) P.nhar_out[f0idx * nsig + tid] = -1; return; }   // same limit as the main pass

  const float omega0 = (float)(2.0 * LLSM_PI * (double)f0 / (double)P.fs);   // czt step (FP_TYPE arg)
  const double nu = (double)omega0 / (2.0 * LLSM_PI);                        // turns per sample
  if(tid < KW) zst[tid] = unit_phasor_turns((double)(tid + 1) * nu * (double)ED_THREADS);
  __syncthreads();

  // ---- one pass over the sample pairs n = tid, tid + 128, ... of the window (ws is even, so the periodic Blackman
  //      window is symmetric about m = half: w(half +- n) = 0.42 + 0.5 cos(2 pi n / ws) + 0.08 cos(4 pi n / ws), advanced
  //      by a fixed rotation in double and used in the same iteration): symmetric / antisymmetric pairs
  //      (x+ + x-, x+ - x-) against cos / sin of every harmonic; the window sum rides along
  const int npair = half + 1;
  float re[NC][KW], im[NC][KW];
  float2 w[KW], z[KW];
#pragma unroll
  for(int k = 0; k < KW; k ++) {
    z[k] = zst[k];
#pragma unroll
    for(int c = 0; c < NC; c ++) { re[c][k] = 0.f; im[c][k] = 0.f; }
  }
  double wsum = 0, cs, sn, cstep, sstep;
  sincospi(2.0 * (double)tid / (double)ws, &sn, &cs);
  sincospi(2.0 * (double)ED_THREADS / (double)ws, &sstep, &cstep);
  auto load_pairs = [&](int n, float (&xp)[NC], float (&xm)[NC]) {
    const int ip = center + n, im_ = center - n;
    const bool okp = n < half && ip >= 0 && ip < P.nx, okm = n >= 1 && n < npair && im_ >= 0 && im_ < P.nx;
#pragma unroll
    for(int c = 0; c < NC; c ++) {
      const float* x = x0 + (size_t)c * P.xstride;
      xp[c] = (c < nsig && okp) ? x[ip] : 0.f;
      xm[c] = (c < nsig && okm) ? x[im_] : 0.f;
    }
  };
  float xpn[NC], xmn[NC];
  load_pairs(tid, xpn, xmn);
  int step = 0;
  for(int n = tid; n < npair; n += ED_THREADS, step ++) {
    float xp[NC], xm[NC];
#pragma unroll
    for(int c = 0; c < NC; c ++) { xp[c] = xpn[c]; xm[c] = xmn[c]; }
    if(n + ED_THREADS < npair) load_pairs(n + ED_THREADS, xpn, xmn);       // one step ahead of its use
    if((step & 7) == 0) {
#pragma unroll
      for(int k = 0; k < KW; k ++) if(k < nh) w[k] = unit_phasor_turns((double)(k + 1) * nu * (double)n);
    }
    const float wn = (float)(0.42 + 0.5 * cs + 0.08 * (2.0 * cs * cs - 1.0));
    if(n < half) wsum += wn;                                          // m = half + n
    if(n >= 1) wsum += wn;                                            // m = half - n
    { const double c2 = cs * cstep - sn * sstep; sn = sn * cstep + cs * sstep; cs = c2; }
    float e[NC], o[NC];
#pragma unroll
    for(int c = 0; c < NC; c ++) { const float a = wn * xp[c], d = wn * xm[c]; e[c] = a + d; o[c] = a - d; }
#pragma unroll
    for(int k = 0; k < KW; k ++) if(k < nh) {
#pragma unroll
      for(int c = 0; c < NC; c ++) {
        re[c][k] = fmaf(e[c], w[k].x, re[c][k]);
        im[c][k] = fmaf(-o[c], w[k].y, im[c][k]);
      }
      w[k] = cmul(w[k], z[k]);
    }
  }
  // ---- fixed-order reductions over the 128 threads: the window sum (double), then value v = (c, k, re | im) in row v
  red[tid] = wsum;
#pragma unroll
  for(int c = 0; c < NC; c ++)
#pragma unroll
    for(int k = 0; k < KW; k ++) {
      part[(2 * (c * KW + k)) * ED_STRIDE + tid] = re[c][k];
      part[(2 * (c * KW + k) + 1) * ED_STRIDE + tid] = im[c][k];
    }
  __syncthreads();
  for(int o = ED_THREADS >> 1; o > 0; o >>= 1) { if(tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  const float winsum = (float)red[0];                                 // FP_TYPE winsum = sumfp(w, nx)
  // thread t sums every fourth element of row t / 4 (rows beyond 32 in further rounds), two shuffles finish the row
  for(int r0 = 0; r0 < 2 * NC * KW; r0 += ED_THREADS / 4) {
    const int r = r0 + (tid >> 2), q = tid & 3;
    float acc = 0.f;
    if(r < 2 * NC * KW) {
      const float* row = part + r * ED_STRIDE + q;
#pragma unroll 8
      for(int u = 0; u < ED_THREADS / 4; u ++) acc += row[4 * u];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    __syncthreads();
    if(r < 2 * NC * KW && q == 0) part[r * ED_STRIDE] = acc;          // row total in column 0
  }
  __syncthreads();
  for(int e2 = tid; e2 < nsig * P.maxnhar; e2 += ED_THREADS) {
    const int c = e2 / P.maxnhar, k = e2 - c * P.maxnhar;
    const size_t at = (f0idx * nsig + c) * P.maxnhar + k;
    if(k < nh && k < KW && c < NC) {
      const float sre = part[(2 * (c * KW + k)) * ED_STRIDE], sim = part[(2 * (c * KW + k) + 1) * ED_STRIDE];
      harmonic_finish(sre, sim, k, half, f0, P.fs, omega0, winsum, &P.ampl[at], &P.phse[at]);
    } else { P.ampl[at] = 0; P.phse[at] = 0; }
  }
  if(tid < nsig) P.nhar_out[f0idx * nsig + tid] = nh;
}

// grid (nfrm, nutt): one CTA per frame; or, with P.long_list, a fixed grid walking the listed frames
template <int NC, int KW>
__global__ void __launch_bounds__(ED_THREADS) envelope_dft_kernel(HarmDftParams P) {
  LLSM_DYN_SMEM(smem);
  if(P.long_list == nullptr) { envelope_dft_frame<NC, KW>(P, blockIdx.x, blockIdx.y, smem); return; }
  const int n = P.long_list[0];
  for(int e = blockIdx.x; e < n; e += gridDim.x) {
    const int f = P.long_list[1 + e];
    envelope_dft_frame<NC, KW>(P, f % P.nfrm, f / P.nfrm, smem);
    __syncthreads();
  }
}

template <int NC, int KW>
static inline size_t env_dft_smem(int max_half) {
  (void)max_half;
  return (size_t)ED_THREADS * 8 + (size_t)(2 * NC * KW * ED_STRIDE) * 4 + KW * 8 + 32;
}

// ------------------------------------------------------------------------------------------
// Envelope pass from a STAGED SEGMENT: a CTA owns seg_frames consecutive frames of one utterance; thread 0 brings the
// slice of every sub-band signal those frames touch -- [centre(first) - cap, centre(last) + cap], clipped, 16-byte
// aligned -- into shared memory with one bulk asynchronous copy per channel (cp.async.bulk, the TMA engine's 1-D form,
// completed on an mbarrier), and each WARP then analyses whole frames out of it: the short-time means, then the pairs
// n = lane, lane + 32, ... of the window against one phasor per harmonic (advanced 32 samples per step, re-seeded every
// 16 steps), all channels x harmonics accumulated in registers, reduced by a transposing butterfly inside the warp (31
// shuffles per 32 values; lane q ends up with value q) and finished by the lanes that hold them. Neighbouring frames
// overlap by three quarters of their windows, so a sample crosses L2 -> SM about 1.3 times instead of 4 times, no thread
// ever waits on a global load of sample data (the one-CTA-per-frame kernel above spent 44 % of its stall samples there
// and lived ~10 us per frame: profiles/r2c, r2f), and there is no block barrier after the copy has landed.
// The Blackman window comes from the plan's table (AnaPlan::bwin, L2-resident). Frames whose windows exceed the
// capacity (f0 below ~80 Hz) are appended to P.long_list and served by envelope_dft_kernel in list mode.
// ------------------------------------------------------------------------------------------
#define EVS_WARPS 8
#define EVS_THREADS (EVS_WARPS * 32)

// sum over the 32 lanes of each of 32 per-lane values; lane q returns the total of value q
__device__ __forceinline__ float warp_transpose_sum32(float (&a)[32], int lane) {
#pragma unroll
  for(int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for(int j = 0; j < o; j ++) {
      const float send = up ? a[j] : a[j + o];
      const float keep = up ? a[j + o] : a[j];
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return a[0];
}

static inline size_t env_seg_smem(int nsig, int seg_len) { return 16 + (size_t)nsig * seg_len * 4 + 16; }

template <int NC, int KW>
__global__ void __launch_bounds__(EVS_THREADS, 2) envelope_seg_kernel(HarmDftParams P) {
  LLSM_DYN_SMEM(smem);
  constexpr int NV = 2 * NC * KW;                                    // values reduced per frame
  static_assert(NV % 32 == 0, "whole groups of 32 values");
  bulk_bar_t* bar = (bulk_bar_t*)smem;
  float* sl = (float*)(smem + 16);                                   // [nsig][SL]
  const int SL = P.seg_len, cap = P.stage_cap, nsig = P.nsig;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, i_lo = blockIdx.x * P.seg_frames;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i_lo >= nf) return;
  const int i_hi = min(i_lo + P.seg_frames, nf);
  int lo = P.center[i_lo] - cap, hi = P.center[i_hi - 1] + cap + 1;
  lo = max(lo, 0) & ~3; hi = (min(hi, P.nx) + 3) & ~3;               // (rows are padded to a multiple of four samples)
  const int len = min(max(hi - lo, 0), SL);
  if(tid == 0) bulk_bar_init(bar);
  __syncthreads();
  if(tid == 0) {
    bulk_expect(bar, (uint32_t)(nsig * len * 4));
    if(len > 0)
      for(int c = 0; c < nsig; c ++)
        bulk_g2s(sl + (size_t)c * SL, P.sig + ((size_t)b * nsig + c) * P.xstride + lo, (uint32_t)(len * 4), bar);
  }
  // this warp's first frame is fetched while the copy is in flight
  int i = i_lo + warp;
  float f0_next = i < i_hi ? P.f0[(size_t)b * P.nfrm + i] : 0.f;
  int cen_next = i < i_hi ? P.center[i] : 0;
  bulk_wait(bar, 0);

  for(; i < i_hi; i += EVS_WARPS) {
    const float f0 = f0_next; const int center = cen_next;
    if(i + EVS_WARPS < i_hi) { f0_next = P.f0[(size_t)b * P.nfrm + i + EVS_WARPS]; cen_next = P.center[i + EVS_WARPS]; }
    const size_t fidx = (size_t)b * P.nfrm + i;
    const double wlen = f0 == 0 ? (double)(P.thop * 2.0f) : 2.0 / (double)f0;
    const int nw = P.edc != nullptr ? (int)round(wlen * (double)P.fs) : 0;
    int half = 0, nh = 0;
    bool listed = nw / 2 + 1 > cap;
    if(f0 > 0) {
      const int ws = ana_winsize(P.fs, f0, P.rel_winsize);
      half = ws >> 1; nh = ana_nhar(P.fs, f0, P.maxnhar);
      listed = listed || half > cap || half > P.bw_cap || half > P.max_half || half < 1;
    }
    if(listed) {                                                     // left to the general kernel
      if(lane == 0) { const int at = atomicAdd(P.long_list, 1); P.long_list[1 + at] = (int)fidx; }
      continue;
    }
    // ---- short-time means (llsm_compute_dc, dsputils.c:117-124)
    if(P.edc != nullptr) {
      for(int c = 0; c < nsig; c ++) {
        const float* x = sl + (size_t)c * SL - lo;
        double acc = 0;
        for(int j = lane; j < nw; j += 4 * 32) {
          float v[4];
#pragma unroll
          for(int u = 0; u < 4; u ++) {
            const int idx = center + j + u * 32 - nw / 2;
            v[u] = (j + u * 32 < nw && idx >= 0 && idx < P.nx) ? x[idx] : 0.f;
          }
#pragma unroll
          for(int u = 0; u < 4; u ++) acc += (double)v[u];
        }
        for(int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if(lane == 0) P.edc[fidx * nsig + c] = nw > 0 ? (float)(acc / nw) : 0.f;
      }
    }
    if(! (f0 > 0)) {                                                 // unvoiced: no harmonic model (layer0.c:106)
      for(int q = lane; q < nsig * P.maxnhar; q += 32) { P.ampl[fidx * nsig * P.maxnhar + q] = 0; P.phse[fidx * nsig * P.maxnhar + q] = 0; }
      if(lane < nsig) P.nhar_out[fidx * nsig + lane] = 0;
      continue;
    }
    const float omega0 = (float)(2.0 * LLSM_PI * (double)f0 / (double)P.fs);   // czt step (FP_TYPE arg)
    const double nu = (double)omega0 / (2.0 * LLSM_PI);
    const float* wv = P.bwin + P.bw_off[half];
    const float winsum = P.bw_sum[half];
    const int npair = half + 1;
    float v[NV];
#pragma unroll
    for(int q = 0; q < NV; q ++) v[q] = 0.f;                         // v[2 (c KW + k)] = re, [.. + 1] = im
    float2 w[KW], z[KW];
#pragma unroll
    for(int k = 0; k < KW; k ++) z[k] = k < nh ? unit_phasor_turns((double)(k + 1) * nu * 32.0) : make_float2(1.f, 0.f);
    const float* xb = sl - lo + center;
    int step = 0;
    float wn_next = lane < npair ? __ldg(wv + lane) : 0.f;
    for(int n = lane; n < npair; n += 32, step ++) {
      if((step & 15) == 0) {
#pragma unroll
        for(int k = 0; k < KW; k ++) if(k < nh) w[k] = unit_phasor_turns((double)(k + 1) * nu * (double)n);
      }
      const float wn = wn_next;
      if(n + 32 < npair) wn_next = __ldg(wv + n + 32);
      const int ip = center + n, im_ = center - n;
      const bool okp = n < half && ip >= 0 && ip < P.nx, okm = n >= 1 && im_ >= 0 && im_ < P.nx;
      float e_[NC], o_[NC];
#pragma unroll
      for(int c = 0; c < NC; c ++) {
        const float xp = (c < nsig && okp) ? xb[(size_t)c * SL + n] : 0.f;
        const float xm = (c < nsig && okm) ? xb[(size_t)c * SL - n] : 0.f;
        const float a = wn * xp, d = wn * xm;
        e_[c] = a + d; o_[c] = a - d;
      }
#pragma unroll
      for(int k = 0; k < KW; k ++) if(k < nh) {
#pragma unroll
        for(int c = 0; c < NC; c ++) {
          v[2 * (c * KW + k)] = fmaf(e_[c], w[k].x, v[2 * (c * KW + k)]);
          v[2 * (c * KW + k) + 1] = fmaf(-o_[c], w[k].y, v[2 * (c * KW + k) + 1]);
        }
        w[k] = cmul(w[k], z[k]);
      }
    }
    // ---- lane 2 j (+ 32 g) receives re, its neighbour im, of the pair j = c KW + k
#pragma unroll
    for(int g = 0; g < NV / 32; g ++) {
      float a[32];
#pragma unroll
      for(int q = 0; q < 32; q ++) a[q] = v[32 * g + q];
      const float tot = warp_transpose_sum32(a, lane);
      const float sim = __shfl_down_sync(0xffffffffu, tot, 1);
      const int j = 16 * g + (lane >> 1), c = j / KW, k = j - c * KW;
      if((lane & 1) == 0 && c < nsig && k < P.maxnhar) {
        const size_t at = (fidx * nsig + c) * P.maxnhar + k;
        if(k < nh) harmonic_finish(tot, sim, k, half, f0, P.fs, omega0, winsum, &P.ampl[at], &P.phse[at]);
        else { P.ampl[at] = 0; P.phse[at] = 0; }
      }
    }
    if(lane < nsig) P.nhar_out[fidx * nsig + lane] = nh;
  }
}

static inline size_t harm_dft_smem(int max_half, int ng) {
  return (size_t)((max_half + 3) & ~1) * 4 + HD_THREADS * 8 + 2 * HD_THREADS * 4 + (ng > 1 ? 0 : (size_t)(max_half + 2) * 8) + 16;
}

static inline int launch_harmonic_dft(const HarmDftParams& Pin, int nutt, cudaStream_t st) {
  HarmDftParams P = Pin;
  int* const long_list = P.long_list;
  if(P.nsig == 1 && P.edc == nullptr && dft_variant() == 1 && P.maxnhar > HD_KW && P.bwin != nullptr && long_list != nullptr) {
    // tensor-core kernel for windows up to 4 periods of 80 Hz (its staging buffers are sized by that capacity and decide
    // how many CTAs share an SM); the rare longer windows are listed and served by the direct kernel in list mode
    int cap = (int)ceil((double)P.fs / mma_min_f0() * (double)P.rel_winsize / 4.0 * 2.0) + 4;
    if(cap > P.max_half) cap = P.max_half;
    if(cap > P.bw_cap) cap = P.bw_cap;
    P.mma_cap_half = cap;
    // segments of frames with the waveform slice staged by bulk copy when the rows are 16-byte aligned
    P.seg_frames = 1; P.seg_len = 0;
    if(mma_seg_frames() > 1 && (P.xstride & 3) == 0 && ((uintptr_t)P.sig & 15) == 0 && P.hop_max > 0) {
      P.seg_frames = mma_seg_frames();
      P.seg_len = ((P.seg_frames - 1) * P.hop_max + 2 * cap + 1 + 8 + 3) & ~3;
    }
    size_t smem = harm_mma_smem(cap, P.maxnhar, P.seg_len);
    size_t smem_d = harm_dft_smem(P.max_half, 1);
    if(smem <= 200 * 1024 && smem_d <= 200 * 1024) {
      if(dev_memset(long_list, 0, 4, st) != 0) return -1;
#ifndef LLSM_EMU
      cudaFuncSetAttribute(harmonic_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
      LLSM_LAUNCH(harmonic_mma_kernel, dim3((P.nfrm + P.seg_frames - 1) / P.seg_frames, nutt), dim3(HM_THREADS), smem, st, P);
      auto kfn = harmonic_dft_kernel<HD_THREADS>;
#ifndef LLSM_EMU
      cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
#endif
      const long long total = (long long)nutt * P.nfrm;
      dim3 lgrid((unsigned)std::max(1LL, std::min(total, (long long)ana_resident_ctas(2))));
      LLSM_LAUNCH(kfn, lgrid, dim3(HD_THREADS), smem_d, st, P);
      return 0;
    }
  }
  if(P.nsig > 1 && P.maxnhar <= 8 && P.nsig <= 8) {               // envelope pass: all channels of a frame in one CTA
#ifndef LLSM_EMU
#define LLSM_ENV_ATTR(kfn, smem) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))
#else
#define LLSM_ENV_ATTR(kfn, smem) (void)0
#endif
#define LLSM_ENV_LAUNCH(NC, KW, grid) { auto kfn = envelope_dft_kernel<NC, KW>; size_t smem = env_dft_smem<NC, KW>(P.max_half); \
    if(smem > 200 * 1024) return -1; \
    LLSM_ENV_ATTR(kfn, smem); \
    LLSM_LAUNCH(kfn, grid, dim3(ED_THREADS), smem, st, P); }
#define LLSM_ENVS_LAUNCH(NC, KW, grid, smem) { auto kfn = envelope_seg_kernel<NC, KW>; \
    LLSM_ENV_ATTR(kfn, smem); \
    LLSM_LAUNCH(kfn, grid, dim3(EVS_THREADS), smem, st, P); }
    // segment kernel (bulk asynchronous copy of the sub-band slices, a warp per frame) + the general kernel on the frames
    // it lists; the copies need 16-byte aligned rows
    const bool staged = P.bwin != nullptr && long_list != nullptr && env_variant() == 1 && P.nsig <= 4 &&
      (P.xstride & 3) == 0 && ((uintptr_t)P.sig & 15) == 0 && P.hop_max > 0;
    if(staged) {
      int cap = P.bw_cap;
      if(cap > P.max_half) cap = P.max_half;
      P.stage_cap = cap; P.nutt = nutt;
      P.seg_frames = env_seg_frames();
      P.seg_len = ((P.seg_frames - 1) * P.hop_max + 2 * cap + 1 + 8 + 3) & ~3;
      const size_t smem = env_seg_smem(P.nsig, P.seg_len);
      if(smem <= 110 * 1024) {
        if(dev_memset(long_list, 0, 4, st) != 0) return -1;
        dim3 grid((P.nfrm + P.seg_frames - 1) / P.seg_frames, nutt);
        if(P.maxnhar <= 4) LLSM_ENVS_LAUNCH(4, 4, grid, smem) else LLSM_ENVS_LAUNCH(4, 8, grid, smem)
        const long long total = (long long)nutt * P.nfrm;
        dim3 lgrid((unsigned)std::max(1LL, std::min(total, (long long)ana_resident_ctas(2))));
        if(P.maxnhar <= 4) LLSM_ENV_LAUNCH(4, 4, lgrid) else LLSM_ENV_LAUNCH(4, 8, lgrid)
        return 0;
      }
    }
    P.long_list = nullptr;
    dim3 grid(P.nfrm, nutt);
    if(P.nsig <= 4 && P.maxnhar <= 4) LLSM_ENV_LAUNCH(4, 4, grid)
    else if(P.nsig <= 4) LLSM_ENV_LAUNCH(4, 8, grid)
    else if(P.maxnhar <= 4) LLSM_ENV_LAUNCH(8, 4, grid)
    else LLSM_ENV_LAUNCH(8, 8, grid)
#undef LLSM_ENV_LAUNCH
#undef LLSM_ENVS_LAUNCH
    return 0;
  }
  P.long_list = nullptr;
  dim3 grid(P.nfrm, nutt * P.nsig), block(HD_THREADS);
  size_t smem = harm_dft_smem(P.max_half, 1);
  if(smem > 200 * 1024) return -1;
  auto kfn = harmonic_dft_kernel<HD_THREADS>;
#ifndef LLSM_EMU
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  LLSM_LAUNCH(kfn, grid, block, smem, st, P);
  return 0;
}

// ------------------------------------------------------------------------------------------
// residual, sub-band energies, short-time mean
// ------------------------------------------------------------------------------------------
__global__ void residual_kernel(const float* x, const float* x_sin, float* x_res, int nx, int xstride,
  int sstride, int rstride) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if(n < nx) x_res[(size_t)b * rstride + n] = x[(size_t)b * xstride + n] - x_sin[(size_t)b * sstride + n];
}

struct DcParams {
  int nfrm, nchannel; const int* nfrm_utt;
  const float* ce; int cstride, nx;
  const float* f0; const int* center;
  float fs, thop;
  float* edc;               // [B][nfrm][nchannel]
};

// short-time mean over round((f0 == 0 ? thop * 2 : 2.0 / f0) * fs) samples (layer0.c:430,446)
__global__ void __launch_bounds__(128) frame_dc_kernel(DcParams P) {
  __shared__ double red[128];
  const int i = blockIdx.x, b = blockIdx.y / P.nchannel, c = blockIdx.y % P.nchannel;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const float f0 = P.f0[(size_t)b * P.nfrm + i];
  double wlen = f0 == 0 ? (double)(P.thop * 2.0f) : 2.0 / (double)f0;
  const int nw = (int)round(wlen * (double)P.fs);
  const float* x = P.ce + ((size_t)b * P.nchannel + c) * P.cstride;
  const int center = P.center[i];
  double acc = 0;
  for(int j = threadIdx.x; j < nw; j += blockDim.x) {
    int idx = center + j - nw / 2;
    if(idx >= 0 && idx < P.nx) acc += (double)x[idx];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for(int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if(threadIdx.x == 0)
    P.edc[((size_t)b * P.nfrm + i) * P.nchannel + c] = nw > 0 ? (float)(red[0] / nw) : 0.f;
}

// ------------------------------------------------------------------------------------------
// Noise PSD analysis
// ------------------------------------------------------------------------------------------
struct NoiseSpecParams {
  int nfrm; const int* nfrm_utt;
  const float* x; int xstride;        // original waveform (spectral envelope)
  const float* x_res; int rstride;    // residual (noise PSD)
  int nx;
  const float* f0; const int* center;
  float fs;
  int nwin;                           // round(thop * 4 * fs)                layer0.c:320
  int nfft, lg_nfft, nspec;           // PSD transform size                  layer0.c:321-322
  int nfft_s, lg_nfft_s;              // envelope transform size             layer0.c:325
  const float* win_psd;               // blackman(nwin)
  float win_power;                    // float-accumulated sum of squares    dsputils.c:254-258
  float std_norm;                     // 0.5 * sum(hanning(1024))            dsputils.c:100-105
  const float2* tw_s; const float2* tw_p;
  float* env;                         // [B][nfrm][nspec] log-power envelope (x 2)
  float* lpsd;                        // [B][nfrm][nspec] log PSD of the residual
};

#define NS_THREADS 256

// One CTA per PAIR of consecutive frames: every transform is a complex FFT carrying frame i in the real
// part and frame i + 1 in the imaginary part. The windowed frames are separated by Hermitian symmetry;
// the log spectra and the liftered cepstra are real and even, so their transforms are real and the two
// frames stay separated in re / im without any post-processing.
// Transforms: block_fft8 (radix-8 register butterflies, padded buffers). The last transform of the envelope is only
// read at every (nfs / nfft)-th bin (layer0.c:341-342): when that step is even, bin 2 j of the nfs-point transform
// of the liftered cepstrum D equals bin j of the (nfs / 2)-point transform of D[n] + D[n + nfs / 2] -- half the size.
__global__ void __launch_bounds__(NS_THREADS) noise_spec_kernel(NoiseSpecParams P) {
  LLSM_DYN_SMEM(smem);
  const int nfs = P.nfft_s;
  const int nmax = nfs > P.nfft ? nfs : P.nfft;
  float2* bufa = (float2*)smem;
  float2* bufb = bufa + fpad(nmax) + 1;
  const int i0 = 2 * blockIdx.x, b = blockIdx.y;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i0 >= nf) return;
  const bool two = i0 + 1 < nf;
  const int tid = threadIdx.x, nth = blockDim.x;
  const float* x = P.x + (size_t)b * P.xstride;
  float f0v[2]; int cen[2], wsv[2];
#pragma unroll
  for(int h = 0; h < 2; h ++) {
    const int i = (h == 0 || two) ? i0 + h : i0;
    f0v[h] = P.f0[(size_t)b * P.nfrm + i];
    cen[h] = P.center[i];
    int ws = P.nwin;
    if(f0v[h] != 0) { float t = P.fs / f0v[h]; t = __fmul_rn(t, 3.0f); ws = (int)t; }   // layer0.c:331
    wsv[h] = ws;
  }
  const size_t orow = ((size_t)b * P.nfrm + i0) * P.nspec;

  // ---- (i) spectral envelope of x: Hann STFT (window 3 periods or nwin), cepstral smoothing
  for(int kb = tid; kb < nfs; kb += nth) {
    float acc[2] = {0.f, 0.f};
#pragma unroll
    for(int h = 0; h < 2; h ++) {
      if(h == 1 && ! two) break;
      const int ws = wsv[h];
      const float rws = 2.0f / (float)ws;
      int j0 = (kb + ws / 2) & (nfs - 1);           // nfs is a power of two
      for(int j = j0; j < ws; j += nfs) {           // time aliasing when the window exceeds nfft
        int idx = cen[h] + j - ws / 2;
        if(idx >= 0 && idx < P.nx) {
          float w = 0.5f - 0.5f * cospif((float)j * rws);
          acc[h] += x[idx] * w;
        }
      }
    }
    bufa[fpad(kb)] = make_float2(acc[0], acc[1]);
  }
  __syncthreads();
  float2* X = block_fft8<false>(bufa, bufb, P.lg_nfft_s, P.tw_s, P.lg_nfft_s);
  float2* Y = (X == bufa) ? bufb : bufa;
  {
    float nrm[2];
#pragma unroll
    for(int h = 0; h < 2; h ++) { float t = 1024.0f / P.std_norm; nrm[h] = t / (float)wsv[h]; }   // dsputils.c:111
    for(int k = tid; k <= nfs / 2; k += nth) {
      const float2 zk = X[fpad(k)], zn = X[fpad((nfs - k) & (nfs - 1))];
      // A = (Zk + conj Zn) / 2, B = (Zk - conj Zn) / (2 i)
      const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
      const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
      float ma = sqrtf(ar * ar + ai * ai) * nrm[0], mb = sqrtf(br * br + bi * bi) * nrm[1];
      const float2 lg = make_float2(logf(ma > 1e-10f ? ma : 1e-10f), logf(mb > 1e-10f ? mb : 1e-10f));
      Y[fpad(k)] = lg;
      if(k > 0 && k < nfs / 2) Y[fpad(nfs - k)] = lg;
    }
  }
  __syncthreads();
  float2* Cq = block_fft8<true>(Y, X, P.lg_nfft_s, P.tw_s, P.lg_nfft_s);     // cepstra * nfft (real, even)
  float2* D = (Cq == bufa) ? bufb : bufa;
  const int estep_lg = P.lg_nfft_s - P.lg_nfft;                              // envelope bin step nfs / nfft = 1 << estep_lg
  const bool fold = estep_lg >= 1;
  {
    float f0s[2];
#pragma unroll
    for(int h = 0; h < 2; h ++) f0s[h] = (f0v[h] == 0 ? 200.0f : f0v[h]) / P.fs;   // layer0.c:338
    const float inv = 1.0f / (float)nfs;
    for(int q = tid; q <= nfs / 2; q += nth) {
      const float2 c = Cq[fpad(q)];
      float cv[2] = {c.x, c.y};
#pragma unroll
      for(int h = 0; h < 2; h ++) {
        double xq = (double)f0s[h] * q;
        float sinc = 1.0f;
        double xr_ = xq - 2.0 * rint(xq * 0.5);                       // reduce to [-1, 1]
        float s1 = sinpif((float)xr_);
        if(q > 0) sinc = s1 / (float)(LLSM_PI * xq);
        const float c2 = fmaf(-2.0f * s1, s1, 1.0f);                  // cos(2 pi xq) = 1 - 2 sin^2(pi xq)
        cv[h] = cv[h] * inv * sinc * (1.18f - 0.18f * c2);
      }
      const float2 d = make_float2(cv[0], cv[1]);
      D[fpad(q)] = d;
      if(! fold && q > 0 && q < nfs / 2) D[fpad(nfs - q)] = d;
    }
  }
  __syncthreads();
  float2* Ev;
  if(fold) {
    // D is even: D[n + nfs / 2] = D[nfs / 2 - n]; fold into the first half (written to the other buffer)
    const int nh2 = nfs >> 1;
    for(int n = tid; n < nh2; n += nth) {
      const float2 u = D[fpad(n)], v = D[fpad(nh2 - n)];
      Cq[fpad(n)] = make_float2(u.x + v.x, u.y + v.y);
    }
    __syncthreads();
    Ev = block_fft8<false>(Cq, D, P.lg_nfft_s - 1, P.tw_s, P.lg_nfft_s);
  } else {
    Ev = block_fft8<false>(D, Cq, P.lg_nfft_s, P.tw_s, P.lg_nfft_s);
  }
  {
    const int emask = (fold ? (nfs >> 1) : nfs) - 1;
    for(int j = tid; j < P.nspec; j += nth) {
      int idx = (j * nfs) >> P.lg_nfft;                                // j * nfs / nfft, layer0.c:341-342
      if(fold) idx = (idx >> 1) & emask;                               // (even by construction)
      const float2 e = Ev[fpad(idx)];
      P.env[orow + j] = e.x * 2.0f;
      if(two) P.env[orow + P.nspec + j] = e.y * 2.0f;
    }
  }
  __syncthreads();

  // ---- (ii) PSD of the residual frames (Blackman, zero-padded at the end; dsputils.c:246-265)
  const float* xr = P.x_res + (size_t)b * P.rstride;
  for(int j = tid; j < P.nfft; j += nth) {
    float v0 = 0.f, v1 = 0.f;
    if(j < P.nwin) {
      const float w = P.win_psd[j];
      int idx = cen[0] + j - P.nwin / 2;
      if(idx >= 0 && idx < P.nx) v0 = w * xr[idx];
      idx = cen[1] + j - P.nwin / 2;
      if(two && idx >= 0 && idx < P.nx) v1 = w * xr[idx];
    }
    bufa[fpad(j)] = make_float2(v0, v1);
  }
  __syncthreads();
  float2* Z = block_fft8<false>(bufa, bufb, P.lg_nfft, P.tw_p, P.lg_nfft);
  for(int j = tid; j < P.nspec; j += nth) {
    const float2 zk = Z[fpad(j)], zn = Z[fpad((P.nfft - j) & (P.nfft - 1))];
    const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
    const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
    float pa = __fadd_rn(__fmul_rn(ar, ar), __fmul_rn(ai, ai)) / P.win_power;
    float pb = __fadd_rn(__fmul_rn(br, br), __fmul_rn(bi, bi)) / P.win_power;
    P.lpsd[orow + j] = logf(pa > 1e-10f ? pa : 1e-10f);                // layer0.c:358
    if(two) P.lpsd[orow + P.nspec + j] = logf(pb > 1e-10f ? pb : 1e-10f);
  }
}

// ------------------------------------------------------------------------------------------
// The same spectra with every transform held in REGISTERS: one warp per pair of consecutive frames, six 1024-point
// complex FFTs (warp_fft.cuh: 32 lanes x 32 points, no block barrier) instead of two 2048-point and two 1024-point
// block FFTs through shared memory. Sizes fixed at nfft_s = 2048, nfft = 1024 (44.1 / 48 kHz at a 5 ms hop); other
// configurations use noise_spec_kernel above.
//   step 0 / 2  frame A / B: the Hann-windowed (zero-phase, time-aliased) frame as z[n] = x[2n] + i x[2n + 1];
//               Z = FFT1024(z); the 2048-point real spectrum X[k] = (Z[k] + Z*[N-k]) / 2 - i W^k (Z[k] - Z*[N-k]) / 2
//               (W = e^{-2 pi i / 2048}; the partner bin N - k sits in lane 32 - l, register 31 - r), L[k] = log |X[k]|, and
//               at once the packed spectrum of the inverse real transform, Z'[k] = (L[k] + L[N-k]) + i W^{-k} (L[k] - L[N-k]).
//   step 1 / 3  IFFT1024(Z') = 2048 (c[2n] + i c[2n + 1]): the cepstrum; lifter (sinc and raised cosine by a rotation
//               along each lane's arithmetic progression of quefrencies); fold D[n] + D[n + 1024] (layer0.c:341-342 reads
//               only the even bins of the last transform) -- registers r and r + 16 of the same lane.
//   step 4      both folded, real and even sequences in one complex FFT (A real, B imaginary): their transforms are
//               real, so re / im are the two envelopes. Frame A's sequence waits in its own (not yet written) output
//               rows while the warp's shared memory serves frame B's transforms.
//   step 5      Blackman PSD of the residual pair (dsputils.c:246-265), separated by Hermitian symmetry.
// The six transforms share ONE inlined copy of the FFT (rolled loop over the steps): unrolled, the kernel would not fit
// the instruction cache.
// ------------------------------------------------------------------------------------------
#define NSW_WARPS 8
#define NSW_THREADS (NSW_WARPS * 32)
#define NSW_WBYTES (WFFT_SCRATCH_BYTES)
static inline size_t noise_spec_warp_smem() { return (size_t)2 * 1024 * 8 + (size_t)NSW_WARPS * NSW_WBYTES + 16; }

// natural logarithm through the special-function unit (lg2.approx: ~2 ulp at the magnitudes met here, i.e. ~2e-6
// nepers on log spectra around -10; logf's range reduction and polynomial were 12 % of the kernel's instructions)
__device__ __forceinline__ float nsw_log(float v) {
#ifdef LLSM_EMU
  return logf(v);
#else
  return __logf(v);
#endif
}

// log |X[k]| of the 2048-point real spectrum from the packed transform: Z = Z[k], Zp = Z[N - k], W = W2048^k
__device__ __forceinline__ float nsw_logmag(float2 Z, float2 Zp, float2 W, float nrm2) {
  const float ar = 0.5f * (Z.x + Zp.x), ai = 0.5f * (Z.y - Zp.y);
  const float br = 0.5f * (Z.x - Zp.x), bi = 0.5f * (Z.y + Zp.y);
  const float tr = W.x * br - W.y * bi, ti = W.x * bi + W.y * br;     // W^k (Z - Z*p) / 2
  const float xr = ar + ti, xi = ai - tr;
  const float m = (xr * xr + xi * xi) * nrm2;                           // log(|X| nrm) = log(|X|^2 nrm^2) / 2: no square root
  return 0.5f * nsw_log(m > 1e-20f ? m : 1e-20f);
}
// packed spectrum of the inverse real transform from L = L[k], Lp = L[N - k] (both real), W = W2048^k
__device__ __forceinline__ float2 nsw_pack(float L, float Lp, float2 W) {
  const float e = L + Lp, h = L - Lp;
  return make_float2(e + h * W.y, h * W.x);                           // e + i h conj(W)
}

__device__ __forceinline__ float nsw_rcp(float v) {
#ifdef LLSM_EMU
  return 1.0f / v;
#else
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r;
#endif
}

__global__ void __launch_bounds__(NSW_THREADS, 2) noise_spec_warp_kernel(NoiseSpecParams P) {
  LLSM_DYN_SMEM(smem);
  constexpr int NF = 1024, NSPEC = 513;
  float2* tw2 = (float2*)smem;                                   // [1024] W1024^{lane k2} (warp_fft1024)
  float2* w2k = tw2 + 1024;                                      // [1024] W2048^k
  char* wbase = (char*)(w2k + 1024);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float2* scratch = (float2*)(wbase + (size_t)warp * NSW_WBYTES);   // [1024] (+ padding: the FFT's transposes)
  float* buf = (float*)scratch;                                  // [2048] float view
  wfft_build_tw2(tw2, P.tw_p);
  for(int e = tid; e < 1024; e += blockDim.x) w2k[e] = P.tw_s[e];

  const int b = blockIdx.y;
  const int i0 = 2 * (blockIdx.x * NSW_WARPS + warp);
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const bool live = i0 < nf;                                     // (idle warps keep the block barriers company)
  const bool two = i0 + 1 < nf;
  const float* xs = P.x + (size_t)b * P.xstride;
  const float* xr = P.x_res + (size_t)b * P.rstride;
  const size_t orow = ((size_t)b * P.nfrm + (live ? i0 : 0)) * NSPEC;
  float2 x[32];
#pragma unroll
  for(int r = 0; r < 32; r ++) x[r] = make_float2(0.f, 0.f);

#pragma unroll 1
  for(int step = 0; step < 6; step ++) {
    // the warps of a CTA walk the steps together: the instruction cache then holds one step's code, not all six
    __syncthreads();
    const int h = (step >> 1) & 1;
    if(! live || (step < 4 && h == 1 && ! two)) continue;
    float f0 = 0.f; int cen = 0, ws = P.nwin;
    if(step < 4) {
      f0 = P.f0[(size_t)b * P.nfrm + i0 + h];
      cen = P.center[i0 + h];
      if(f0 != 0) { float t = P.fs / f0; t = __fmul_rn(t, 3.0f); ws = (int)t; }      // layer0.c:331
    }
    // ---------------- before the transform
    if(step == 0 || step == 2) {
      // Hann-windowed frame, zero phase (window centre at index 0), time-aliased beyond 2048 samples
      const float rws = 2.0f / (float)ws;
      if(ws <= 2 * NF) {
        for(int q = lane; q < 512; q += 32) ((float4*)buf)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const int first = cen - ws / 2;
        if(first >= 0 && first + ws <= P.nx) {                   // interior frame: no bounds tests
          // (the window cosine per tap, not by rotation along the lane: 69 float rotations put 4e-6 of leakage under the
          //  spectrum of the longest windows -- invisible at the peaks, 4e-3 nepers in the valleys between harmonics, and the
          //  smoother's process variance reads exactly those: 0.07 dB between the two noise-spectra kernels on a 7.5 s
          //  utterance at 60 Hz, for 0.15 ms)
          // (Tried and dropped: the taps from a per-plan table holding the oracle's own float values, 4.2 MB in L2 -- 5.11 ->
          //  5.06 ms only, although this line is 13 % of the kernel's instructions: the cosine hides under the tap's global
          //  read; and the exact window moved a -93 dB notch of the 7.5 s test utterance by 0.05 dB)
          for(int j = lane; j < ws; j += 32)
            buf[(j - ws / 2) & (2 * NF - 1)] = xs[first + j] * (0.5f - 0.5f * cospif((float)j * rws));
        } else {
          for(int j = lane; j < ws; j += 32) {
            const int idx = first + j;
            float v = 0.f;
            if(idx >= 0 && idx < P.nx) v = xs[idx] * (0.5f - 0.5f * cospif((float)j * rws));
            buf[(j - ws / 2) & (2 * NF - 1)] = v;
          }
        }
      } else {
        for(int kb = lane; kb < 2 * NF; kb += 32) {
          float acc = 0.f;
          for(int j = (kb + ws / 2) & (2 * NF - 1); j < ws; j += 2 * NF) {
            const int idx = cen + j - ws / 2;
            if(idx >= 0 && idx < P.nx) acc += xs[idx] * (0.5f - 0.5f * cospif((float)j * rws));
          }
          buf[kb] = acc;
        }
      }
      __syncwarp();
#pragma unroll
      for(int r = 0; r < 32; r ++) x[r] = scratch[lane + 32 * r];
      __syncwarp();
    } else if(step == 4) {
      // folded, liftered cepstra: A from its output rows (the scratch when the pair is a single frame), B from the scratch
      const float* dA = two ? P.env + orow + (orow & 1) : buf;              // (8-byte aligned start inside the two rows)
#pragma unroll
      for(int r = 0; r < 32; r ++) x[r] = make_float2(dA[lane + 32 * r], two ? buf[lane + 32 * r] : 0.f);
      __syncwarp();
    } else if(step == 5) {
      // Blackman-windowed residual frames, zero-padded at the end
      const int c0 = P.center[i0], c1 = P.center[i0 + (two ? 1 : 0)];
      for(int j = lane; j < NF; j += 32) {
        float v0 = 0.f, v1 = 0.f;
        if(j < P.nwin) {
          const float w = P.win_psd[j];
          int idx = c0 + j - P.nwin / 2;
          if(idx >= 0 && idx < P.nx) v0 = w * xr[idx];
          idx = c1 + j - P.nwin / 2;
          if(two && idx >= 0 && idx < P.nx) v1 = w * xr[idx];
        }
        scratch[j] = make_float2(v0, v1);
      }
      __syncwarp();
#pragma unroll
      for(int r = 0; r < 32; r ++) x[r] = scratch[lane + 32 * r];
      __syncwarp();
    }
    // ---------------- the transform (steps 1 and 3: inverse)
    warp_fft1024(x, scratch, tw2, lane, (step == 1 || step == 3) ? 1 : 0);
    // ---------------- after the transform: the result goes back to shared memory in natural order and is worked on by
    //                  ROLLED loops (an unrolled register version of these passes was 10 k instructions in all: the
    //                  kernel stalled on instruction fetch, profiles/r2k)
    if(step != 4) {
#pragma unroll
      for(int r = 0; r < 32; r ++) scratch[lane + 32 * r] = x[r];
      __syncwarp();
    }
    if(step == 0 || step == 2) {
      float nrm = 1024.0f / P.std_norm; nrm = nrm / (float)ws;                    // dsputils.c:111
      const float nrm2 = nrm * nrm;
      // bins k and N - k are worked on together, in place: L[k], L[N-k] from Z[k], Z[N-k], then Z'[k], Z'[N-k]
      for(int k = lane; k <= 512; k += 32) {
        const float2 za = scratch[k], zb = scratch[(NF - k) & (NF - 1)];
        const float2 wa = w2k[k], wb = make_float2(-wa.x, wa.y);                   // W^{N-k} = -conj(W^k)
        if(k == 0) {
          const float m0 = (za.x + za.y) * (za.x + za.y) * nrm2, m1 = (za.x - za.y) * (za.x - za.y) * nrm2;   // X[0], X[1024]
          const float L0 = 0.5f * nsw_log(m0 > 1e-20f ? m0 : 1e-20f), L1 = 0.5f * nsw_log(m1 > 1e-20f ? m1 : 1e-20f);
          scratch[0] = make_float2(L0 + L1, L0 - L1);
        } else {
          const float La = nsw_logmag(za, zb, wa, nrm2), Lb = nsw_logmag(zb, za, wb, nrm2);
          scratch[k] = nsw_pack(La, Lb, wa);
          if(k < 512) scratch[NF - k] = nsw_pack(Lb, La, wb);
        }
      }
      __syncwarp();
#pragma unroll
      for(int r = 0; r < 32; r ++) x[r] = scratch[lane + 32 * r];
      __syncwarp();
    } else if(step == 1 || step == 3) {
      // scratch[n] = 2048 (c[2n], c[2n + 1]). Lifter (layer0.c:338-340 via spec2env): quefrency q and its mirror 2048 - q
      // carry the same weight; fold D[q] + D[q + 1024] (only the even bins of the next transform are read): elements
      // m and m + 512. Four progressions per lane: q = 2 lane + e + 64 r and q' = 1024 - 2 lane - e - 64 r, e = 0, 1;
      // sin / cos (pi f0s q) advance by a fixed rotation from seeds reduced in double.
      const float f0s = (f0 == 0 ? 200.0f : f0) / P.fs;                           // layer0.c:338
      const float inv = 1.0f / 2048.0f;
      float2 rot, ph[2][2];
      {
        double u = (double)f0s * 64.0 * 0.5; u -= rint(u);
        float sn, cs; sincospif(2.0f * (float)u, &sn, &cs); rot = make_float2(cs, sn);
      }
#pragma unroll
      for(int m = 0; m < 2; m ++)
#pragma unroll
        for(int e = 0; e < 2; e ++) {
          const int q = m == 0 ? 2 * lane + e : 1024 - 2 * lane - e;
          ph[m][e] = unit_phasor_turns((double)f0s * (double)q * 0.5);
        }
      const float rpf = inv / (float)(LLSM_PI * (double)f0s);
      float2* dst = (step == 1 && two) ? (float2*)(P.env + orow + (orow & 1)) : scratch;
#pragma unroll 2
      for(int r = 0; r < 16; r ++) {
        const int mm = lane + 32 * r;
        const float2 ca = scratch[mm], cb = scratch[mm + 512];
        float wgt[2][2];
#pragma unroll
        for(int m = 0; m < 2; m ++)
#pragma unroll
          for(int e = 0; e < 2; e ++) {
            const int q = m == 0 ? 2 * mm + e : 1024 - 2 * mm - e;
            const float s1 = ph[m][e].y;                                           // sin(pi f0s q)
            const float sinc = q > 0 ? s1 * rpf * nsw_rcp((float)q) : inv;
            const float c2 = fmaf(-2.0f * s1, s1, 1.0f);                           // cos(2 pi f0s q)
            wgt[m][e] = sinc * (1.18f - 0.18f * c2);
            ph[m][e] = m == 0 ? cmul(ph[m][e], rot) : cmul(ph[m][e], make_float2(rot.x, -rot.y));
          }
        dst[mm] = make_float2(ca.x * wgt[0][0] + cb.x * wgt[1][0], ca.y * wgt[0][1] + cb.y * wgt[1][1]);
      }
#ifndef LLSM_EMU
      __threadfence_block();
#endif
      __syncwarp();
    } else if(step == 4) {
#pragma unroll
      for(int r = 0; r <= 16; r ++) {
        const int j = lane + 32 * r;
        if(j < NSPEC) {
          P.env[orow + j] = x[r].x * 2.0f;
          if(two) P.env[orow + NSPEC + j] = x[r].y * 2.0f;
        }
      }
    } else {
      for(int j = lane; j < NSPEC; j += 32) {
        const float2 zk = scratch[j], zn = scratch[(NF - j) & (NF - 1)];
        const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
        const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
        const float pa = __fadd_rn(__fmul_rn(ar, ar), __fmul_rn(ai, ai)) / P.win_power;
        const float pb = __fadd_rn(__fmul_rn(br, br), __fmul_rn(bi, bi)) / P.win_power;
        P.lpsd[orow + j] = nsw_log(pa > 1e-10f ? pa : 1e-10f);                     // layer0.c:358
        if(two) P.lpsd[orow + NSPEC + j] = nsw_log(pb > 1e-10f ? pb : 1e-10f);
      }
    }
  }
}

#define KCH 8
struct KalmanParams {
  int nfrm, nspec; const int* nfrm_utt;
  const float* env; // in: envelope
  float* pvar;     // scratch: posterior variances
  float* lpsd;     // in: raw log PSD; out: smoothed + Euler gamma
  float* res;      // out: residual (scratch for Q on the way)
  float* filt;     // scratch: filtered means
};

// One thread per (utterance, bin): process variance from a 3-frame moving variance of the
// envelope, observation variance pi^2/6, random-walk Kalman filter + RTS smoother along time
// (layer0.c:361-385; filter conventions of the oracle's ciglet shim: x0 = z0, P0 = R0).
// (Tried and dropped: the output stage fused into the backward pass -- 127 bins + a halo bin per CTA, the smoother's
//  results of eight frames parked in shared memory, the process variance recomputed instead of stored: 7 array passes
//  instead of 13, but 3.01 ms against 2.06 + 0.79 ms for the two kernels at C2: the double-precision exp / log10 of the
//  output stage do not overlap with the latency-bound time recursion inside one CTA.)
// (Also dropped: the process variance recomputed from the envelope in the backward pass instead of stored and read back --
//  one array pass less, bit-identical, but 2.12 against 2.06 ms: the two float divisions per frame cost more than the pass.)
__global__ void __launch_bounds__(128) noise_kalman_kernel(KalmanParams P) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if(j >= P.nspec) return;
  const int n = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(n <= 0) return;
  const size_t base = (size_t)b * P.nfrm * P.nspec + j;
  const size_t st = P.nspec;
  const double R = (double)(float)(LLSM_PI * LLSM_PI / 6.0);      // LOGCHI2VAR stored as FP_TYPE
  float e_prev = P.env[base], e_cur = e_prev;
  double xk = 0, Pk = 0;
  // The time loop is latency-bound (one dependent chain per thread): loads of the next KCH frames are
  // issued together before the chain advances.
  for(int i0 = 0; i0 < n; i0 += KCH) {
    float ev[KCH], zv[KCH];
#pragma unroll
    for(int u = 0; u < KCH; u ++) {
      int ie = i0 + u + 1; if(ie > n - 1) ie = n - 1;
      ev[u] = P.env[base + (size_t)ie * st];
      int iz = i0 + u; if(iz > n - 1) iz = n - 1;
      zv[u] = P.lpsd[base + (size_t)iz * st];
    }
#pragma unroll
    for(int u = 0; u < KCH; u ++) {
      const int i = i0 + u;
      if(i < n) {
        const float e_next = ev[u];
        // Q[i] = max(1e-8, m2 / 3 - m1 * m1 / 9) over frames clamp(i-1..i+1), float arithmetic
        float m1 = 0.f, m2 = 0.f;
        m1 = __fadd_rn(m1, e_prev); m2 = __fadd_rn(m2, __fmul_rn(e_prev, e_prev));
        m1 = __fadd_rn(m1, e_cur);  m2 = __fadd_rn(m2, __fmul_rn(e_cur, e_cur));
        m1 = __fadd_rn(m1, e_next); m2 = __fadd_rn(m2, __fmul_rn(e_next, e_next));
        float qv = __fadd_rn(m2 / 3.0f, -(__fmul_rn(m1, m1) / 9.0f));
        float Q = qv > 1e-8f ? qv : 1e-8f;
        const double z = (double)zv[u];
        if(i == 0) { xk = z; Pk = R; }
        else {
          double Pp = Pk + (double)Q;
          double K = Pp / (Pp + R);
          xk += K * (z - xk);
          Pk = (1.0 - K) * Pp;
        }
        P.filt[base + i * st] = (float)xk;
        P.pvar[base + i * st] = (float)Pk;       // posterior variance (FP_TYPE array P)
        P.res[base + i * st] = Q;
        e_prev = e_cur; e_cur = e_next;
      }
    }
  }
  // RTS smoother, residual, bias removal
  double sn = (double)P.filt[base + (size_t)(n - 1) * st];
  float qnext = 0.f;
  for(int t0 = n - 1; t0 >= 0; t0 -= KCH) {
    float yv[KCH], qv_[KCH], pv[KCH], rv[KCH];
#pragma unroll
    for(int u = 0; u < KCH; u ++) {
      int t = t0 - u; if(t < 0) t = 0;
      yv[u] = P.filt[base + (size_t)t * st]; qv_[u] = P.res[base + (size_t)t * st];
      pv[u] = P.pvar[base + (size_t)t * st]; rv[u] = P.lpsd[base + (size_t)t * st];
    }
#pragma unroll
    for(int u = 0; u < KCH; u ++) {
      const int t = t0 - u;
      if(t >= 0) {
        const float yt = yv[u], qt = qv_[u];
        if(t < n - 1) {
          double Pt = (double)pv[u];
          double Pp = Pt + (double)qnext;
          double Cg = Pt / Pp;
          sn = (double)yt + Cg * (sn - (double)yt);
        }
        float s = (float)sn;                        // stored as FP_TYPE; the recursion stays in double
        P.res[base + t * st] = rv[u] - s;           // layer0.c:381
        P.lpsd[base + t * st] = (float)((double)s + 0.57721566);   // layer0.c:382 EULERGAMMA
        qnext = qt;
      }
    }
  }
}

struct PsdOutParams {
  int nfrm, nspec, npsd; const int* nfrm_utt;
  const float* lpsd; const float* res;
  const int* ip_k; const float* ip_r;  // [npsd] interp1u plan (exclusive-end grid)
  float fs;
  float* psd; float* psdres;           // [B][nfrm][npsd]
};

// (Arithmetic in float: the reference evaluates these in FP_TYPE; the double-precision exp / log10 this kernel used first
//  made it FP64- and conversion-bound -- 44 % / 49 % of those pipes, profiles/r2t -- for differences of 1e-6 dB.)
__global__ void __launch_bounds__(128) noise_psd_out_kernel(PsdOutParams P) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t irow = ((size_t)b * P.nfrm + i) * P.nspec, orow = ((size_t)b * P.nfrm + i) * P.npsd;
  const float sc = 44100.0f / P.fs;
  for(int j = threadIdx.x; j < P.npsd; j += blockDim.x) {
    const int k = P.ip_k[j]; const float r = P.ip_r[j];
    float a = P.lpsd[irow + k], rr = P.res[irow + k];
    if(r != 0.f) {
      a = fmaf(P.lpsd[irow + k + 1] - a, r, a);
      rr = fmaf(P.res[irow + k + 1] - rr, r, rr);
    }
    const float lin = expf(a) * sc;                                               // layer0.c:402
    P.psd[orow + j] = 10.0f * log10f(lin + 1e-12f);                               // layer0.c:403
    P.psdres[orow + j] = rr * (float)(10.0 / 2.3025851);                          // LOG2IN
  }
}

// ------------------------------------------------------------------------------------------
// Harmonic estimator, peak-picking method (LLSM_AOPTION_HMPP): llsm_compute_spectrogram
// (dsputils.c:96-115, Blackman, zero-phase, one FFT size per utterance from its lowest voiced F0:
// llsm_get_fftsize dsputils.c:318-326), log magnitude (dsputils.c:203-204), then per harmonic the
// arg-max within +-0.3 f0, parabolic refinement on the log spectrum and a linear interpolation of
// the WRAPPED phase between the two neighbouring bins (llsm_harmonic_peakpicking dsputils.c:126-143).
// ------------------------------------------------------------------------------------------
struct MinF0Params { int nfrm; const int* nfrm_utt; const float* f0; float fs, rel_winsize; int* nfft_utt; };

__global__ void __launch_bounds__(128) utt_fftsize_kernel(MinF0Params P) {
  __shared__ float red[128];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  float m = 1000.f;                                   // dsputils.c:319
  for(int i = tid; i < nf; i += blockDim.x) { float f = P.f0[(size_t)b * P.nfrm + i]; if(f > 0 && f < m) m = f; }
  red[tid] = m;
  __syncthreads();
  for(int o = blockDim.x >> 1; o > 0; o >>= 1) { if(tid < o) red[tid] = fminf(red[tid], red[tid + o]); __syncthreads(); }
  if(tid == 0) {
    int mw = ana_winsize(P.fs, red[0], P.rel_winsize);  // round(fs / minf0 * rel_winsize / 2) * 2
    P.nfft_utt[b] = pow2_ceil(log2((double)mw));
  }
}

struct HarmPpParams {
  int nfrm; const int* nfrm_utt;
  const float* sig; int nsig, nx, xstride;
  const float* f0; const int* center; const int* nfft_utt;
  float fs, rel_winsize, std_norm;                    // std_norm = 0.5 * sum(blackman(1024))
  int maxnhar;
  int* nhar_out; float* ampl; float* phse;
  const float2* tw; int ntw; int max_nfft;
  int min_nfft;                                       // this launch serves utterances with min_nfft < nfft <= max_nfft
  const float* bwin; const int* bw_off; int bw_cap;   // window table (AnaPlan::bwin), optional
};

#define HP_THREADS 256

__global__ void __launch_bounds__(HP_THREADS) harmonic_pp_kernel(HarmPpParams P) {
  LLSM_DYN_SMEM(smem);
  float2* bufa = (float2*)smem;
  float2* bufb = bufa + P.max_nfft;
  const int i = blockIdx.x;
  const int b = blockIdx.y / P.nsig, c = blockIdx.y % P.nsig;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int tid = threadIdx.x, nth = blockDim.x;
  const size_t fidx = ((size_t)b * P.nfrm + i) * P.nsig + c;
  if(i >= nf) return;
  const float f0 = P.f0[(size_t)b * P.nfrm + i];
  if(! (f0 > 0)) {
    if(tid == 0) P.nhar_out[fidx] = 0;
    for(int k = tid; k < P.maxnhar; k += nth) { P.ampl[fidx * P.maxnhar + k] = 0; P.phse[fidx * P.maxnhar + k] = 0; }
    return;
  }
  const int nfft = P.nfft_utt[b];
  if(nfft <= P.min_nfft) return;                      // served by the launch of a smaller tier
  if(nfft > P.max_nfft) { if(tid == 0 && P.max_nfft >= 8192) P.nhar_out[fidx] = -1; return; }
  int lg = 0; while((1 << lg) < nfft) lg ++;
  const int ws = ana_winsize(P.fs, f0, P.rel_winsize);
  const int nh = ana_nhar(P.fs, f0, P.maxnhar);
  const float* x = P.sig + ((size_t)b * P.nsig + c) * P.xstride;
  const int center = P.center[i];
  // zero-phase Blackman frame (window centre at buffer index 0)
  for(int k = tid; k < nfft; k += nth) bufa[k] = make_float2(0.f, 0.f);
  __syncthreads();
  const int hw = ws >> 1;
  const float* wtab = (P.bwin != nullptr && hw >= 1 && hw <= P.bw_cap) ? P.bwin + P.bw_off[hw] : nullptr;   // w(hw +- n) at [n]
  for(int j = tid; j < ws; j += nth) {
    int idx = center + j - ws / 2;
    float v = 0.f;
    if(idx >= 0 && idx < P.nx) {
      float w;
      if(wtab) w = wtab[j >= hw ? j - hw : hw - j];
      else {
        double s1, c1, s2, c2;
        sincospi(2.0 * (double)j / (double)ws, &s1, &c1); sincospi(4.0 * (double)j / (double)ws, &s2, &c2);
        w = (float)(0.42 - 0.5 * c1 + 0.08 * c2);
      }
      v = x[idx] * w;
    }
    int k = ((j - ws / 2) % nfft + nfft) % nfft;      // ws <= nfft here: no aliasing, plain placement
    bufa[k] = make_float2(v, 0.f);
  }
  __syncthreads();
  float2* X = block_fft<false>(bufa, bufb, lg, P.tw, P.ntw);
  float2* Y = (X == bufa) ? bufb : bufa;
  // log magnitude (x: log(|X| normaliser + 1e-8)) and phase (y) per bin
  float normalizer = 1024.0f / P.std_norm; normalizer = normalizer / (float)ws;
  // only the bins the peak picker can read: up to the upper search bound of the last harmonic, plus the parabola's and
  // the phase interpolation's neighbours (the envelope pass asks for 5 harmonics: 1 / 8 of the spectrum)
  int kneed = nfft / 2;
  {
    float up_f = __fmul_rn(f0, (float)nh + 0.3f); up_f = up_f / P.fs; up_f = __fmul_rn(up_f, (float)nfft);
    const int u = (int)round((double)up_f) + 3;
    if(u < kneed) kneed = u;
    if(kneed < 3) kneed = 3;
  }
  for(int k = tid; k <= kneed; k += nth) {
    float2 v = X[k];
    float mag = (float)sqrt((double)v.x * v.x + (double)v.y * v.y) * normalizer;
    Y[k] = make_float2((float)log((double)mag + 1e-8), (float)atan2((double)v.y, (double)v.x));
  }
  __syncthreads();
  for(int k = tid; k < nh; k += nth) {
    const int hi = k + 1;
    float lo_f = __fmul_rn(f0, (float)hi - 0.3f); lo_f = lo_f / P.fs; lo_f = __fmul_rn(lo_f, (float)nfft);
    float up_f = __fmul_rn(f0, (float)hi + 0.3f); up_f = up_f / P.fs; up_f = __fmul_rn(up_f, (float)nfft);
    int l = (int)round((double)lo_f), u = (int)round((double)up_f);
    if(l < 1) l = 1;
    if(u > nfft / 2 - 1) u = nfft / 2 - 1;
    int pk = l;
    for(int q = l + 1; q <= u; q ++) if(Y[q].x > Y[pk].x) pk = q;
    double a = Y[pk - 1].x, bq = Y[pk].x, cq = Y[pk + 1].x;
    double a1 = (a + cq) * 0.5 - bq, a2 = (cq - a) * 0.5;
    double xo = a1 != 0 ? -a2 / (2.0 * a1) : 0;
    if(! (fabs(xo) < 1.0)) xo = 0;
    float pf = (float)(pk + xo);
    float pa = (float)(a1 * xo * xo + a2 * xo + bq);
    P.ampl[fidx * P.maxnhar + k] = (float)exp((double)pa);
    int ib = (int)pf;
    float pa0 = Y[ib].y, pa1 = Y[ib + 1].y;
    P.phse[fidx * P.maxnhar + k] = (float)((double)pa0 + ((double)pa1 - (double)pa0) * fmod((double)pf, 1.0));
  }
  for(int k = nh + tid; k < P.maxnhar; k += nth) { P.ampl[fidx * P.maxnhar + k] = 0; P.phse[fidx * P.maxnhar + k] = 0; }
  if(tid == 0) P.nhar_out[fidx] = nh;
}

// ------------------------------------------------------------------------------------------
// Unwindowed harmonic frames: llsm_synthesize_harmonic_frame (dsputils.c:328-336, gensins) and
// llsm_synthesize_harmonic_frame_iczt (:338-351), the per-frame routines dsputils.h exports and the reference's
// tests call directly (test/test-harmonic.c:40-43). y[f][j] = sum_k a_k cos(2 pi nu (k + 1)(j - nx / 2) + phi_k),
// j < nx. iczt != 0 reproduces that branch's two quirks: the fundamental is the FLOAT-rounded omega0 = 2 pi f0 and
// harmonics with index >= nx - 1 are dropped (the transform only has nx bins). One CTA per frame, a thread per
// sample; phases are reduced in double (they reach 1e3 rad). A convenience entry, not a hot path (the synthesis
// loops use the harmonic-bank kernels).
// ------------------------------------------------------------------------------------------
struct HarmFrameParams {
  int nfrm, maxnhar, nx, iczt;
  const int* nhar; const float* f0n;          // [nfrm] harmonics in use, fundamental in cycles per sample
  const float* ampl; const float* phse;       // [nfrm][maxnhar]
  float* y;                                   // [nfrm][nx]
};

__global__ void __launch_bounds__(256) harmonic_frame_kernel(HarmFrameParams P) {
  const int f = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j >= P.nx) return;
  int nh = P.nhar[f];
  if(nh > P.maxnhar) nh = P.maxnhar;
  const float f0 = P.f0n[f];
  double nu = (double)f0;
  if(P.iczt) {
    const float omega0 = (float)(2.0 * LLSM_PI * (double)f0);
    nu = (double)omega0 / (2.0 * LLSM_PI);
    if(nh > P.nx - 1) nh = P.nx - 1;
  }
  const double t = (double)(j - P.nx / 2);
  const float* a = P.ampl + (size_t)f * P.maxnhar; const float* ph = P.phse + (size_t)f * P.maxnhar;
  double acc = 0;
  for(int k = 0; k < nh; k ++) {
    double u = nu * (double)(k + 1) * t;
    u -= rint(u);
    acc += (double)a[k] * cos(2.0 * LLSM_PI * u + (double)ph[k]);
  }
  P.y[(size_t)f * P.nx + j] = (float)acc;
}
