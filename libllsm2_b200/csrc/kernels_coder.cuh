// Frame coder on the device: llsm_coder_encode (coder.c:85-163) and
// llsm_coder_decode_layer0 / _layer1 (coder.c:165-292) for a batch of frames, one CTA per frame.
//
//   encode: F0, noise PSD (dB), Rd, VTMAGN (dB)  ->  [voicing, f0, rd, order_spec mel-cepstral numbers, order_bap
//           band aperiodicities]: total power spectrum (noise PSD + LF source x vocal tract x lip radiation) on the
//           full linear axis, band aperiodicity from the harmonic / total ratio, log, mel axis, DCT, truncation
//   decode: the inverse -- mel-cepstrum -> log spectrum -> linear axis, aperiodicity -> harmonic magnitude and noise
//           power, then either the layer-1 members (VTMAGN after removing lip and source, LF source phases) or the
//           layer-0 harmonics (amplitudes by interpolation, minimum-phase vocal tract + lip + source phase).
//
// The two cosine transforms of the reference (Ooura ddct, sizes nspec - 1 and order_spec) are only ever needed at
// order_spec coefficients, so they are evaluated directly: order_spec x (nspec - 1) terms, cosines by a rotation
// recurrence in double. Everything else follows the reference's evaluation types (double where C promotes).
// The axes (coder.c:61-71) are built once on the host (driver_coder.h) with the reference's float expressions.
#pragma once
#include "common.cuh"
#include "lf_model.cuh"
#include "kernels_layer1.cuh"

#define CODER_THREADS 256

struct CoderParams {
  int nfrm; const int* nfrm_utt;
  int npsd, nspec, order_spec, order_bap, maxnhar;
  float fnyq, lip_radius;
  const float* psdaxis;     // [npsd]   linspace(0, fnyq, npsd)
  const float* faxis;       // [nspec]  fnyq * 2 * i / nfullspec
  const float* melaxis;     // [nspec]
  const float* apaxis;      // [order_bap + 1]
  // encode
  const float* f0; const float* psd; const float* rd; const float* vtmagn;
  float* enc;               // [B][F][order_spec + order_bap + 3]
  // decode
  const float* enc_in;
  int use_layer1;
  float* o_f0; float* o_rd; float* o_psd; int* o_nhar;
  float* o_ampl; float* o_phse;             // layer 0
  float* o_vtmagn; float* o_vsphse;         // layer 1
  const float2* tw; int ntw; int max_nfft;  // minimum-phase FFT (layer-0 decode)
};

// interp1 of the oracle's ciglet shim: clamped ends, binary search, linear in double
__device__ __forceinline__ float coder_interp1(const float* __restrict__ xi, const float* yi, int ni, float v) {
  if(! (v > xi[0])) return yi[0];
  if(v >= xi[ni - 1]) return yi[ni - 1];
  int lo = 0, hi = ni - 1;
  while(hi - lo > 1) {
    const int mid = (lo + hi) / 2;
    if(xi[mid] <= v) lo = mid; else hi = mid;
  }
  const double r = ((double)v - (double)xi[lo]) / ((double)xi[hi] - (double)xi[lo]);
  return (float)((double)yi[lo] + ((double)yi[hi] - (double)yi[lo]) * r);
}

// out[k] = sum_{j < n} a[j] cos(pi (j + 1/2) k / n), k < nk  (ddct(n, -1, a) at its first nk outputs). The sum over j
// is cut into four slices per k; partial sums are added in slice order. part: [4][nk] doubles of shared memory.
__device__ void coder_dct(const float* a, int n, int nk, float* out, double* part) {
  const int tid = threadIdx.x, nth = blockDim.x;
  for(int e = tid; e < 4 * nk; e += nth) {
    const int k = e % nk, sl = e / nk;
    const int j0 = (int)((long long)n * sl / 4), j1 = (int)((long long)n * (sl + 1) / 4);
    const double step = (double)k / (double)n;                    // turns / 2 per unit j
    double c = cospi(((double)j0 + 0.5) * step), s = sinpi(((double)j0 + 0.5) * step);
    const double dc = cospi(step), ds = sinpi(step);
    double acc = 0;
    for(int j = j0; j < j1; j ++) {
      acc += (double)a[j] * c;
      const double nc = c * dc - s * ds; s = s * dc + c * ds; c = nc;
      if(((j - j0) & 63) == 63) { c = cospi(((double)j + 1.5) * step); s = sinpi(((double)j + 1.5) * step); }
    }
    part[sl * nk + k] = acc;
  }
  __syncthreads();
  for(int k = tid; k < nk; k += nth) out[k] = (float)(((part[k] + part[nk + k]) + part[2 * nk + k]) + part[3 * nk + k]);
  __syncthreads();
}
// out[k] = sum_{j < nj} a[j] cos(pi j (k + 1/2) / n), k < nk  (ddct(n, 1, a) with a[j] = 0 for j >= nj)
__device__ void coder_idct(const float* a, int nj, int n, int nk, float* out) {
  for(int k = threadIdx.x; k < nk; k += blockDim.x) {
    const double step = ((double)k + 0.5) / (double)n;
    const double dc = cospi(step), ds = sinpi(step);
    double c = 1.0, s = 0.0, acc = 0;
    for(int j = 0; j < nj; j ++) {
      acc += (double)a[j] * c;
      const double nc = c * dc - s * ds; s = s * dc + c * ds; c = nc;
      if((j & 63) == 63) { c = cospi(((double)j + 1.0) * step); s = sinpi(((double)j + 1.0) * step); }
    }
    out[k] = (float)acc;
  }
  __syncthreads();
}

// llsm_lipfilter (dsputils.c:396-412) on magnitudes: element i at frequency f0 (1 + i)
__device__ __forceinline__ float coder_lip_abs(float radius, float f0, int i) {
  const float omega = (float)((double)f0 * (1.0 + i) * 2.0 * LLSM_PI);
  return lip_abs(lip_response(radius, omega));
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CODER_THREADS) coder_encode_kernel(CoderParams P) {
  LLSM_DYN_SMEM(smem);
  const int ns = P.nspec, os = P.order_spec, ob = P.order_bap;
  double* part = (double*)smem;                       // [4][os]
  float* spsd = (float*)(part + 4 * os);              // [ns] total power -> log intensity
  float* senv = spsd + ns;                            // [ns] harmonic power / mel-axis spectrum
  float* coef = senv + ns;                            // [os]
  LfSolved* lfs = (LfSolved*)(coef + ((os + 1) & ~1));
  float* lf0 = (float*)(lfs + 1);
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t r = (size_t)b * P.nfrm + i;
  const int dim = os + ob + 3;
  float* enc = P.enc + r * dim;
  const float f0 = P.f0[r];
  const bool voiced = f0 > 0;
  const float* psd = P.psd + r * P.npsd;

  // noise PSD: scaled axis -> full axis, intensity (dB) -> power (coder.c:96-100)
  for(int j = tid; j < ns; j += nth)
    spsd[j] = (float)exp((double)coder_interp1(P.psdaxis, psd, P.npsd, P.faxis[j]) * 2.3025851 / 10.0);
  if(tid == 0) {
    enc[0] = voiced ? 1.f : 0.f; enc[1] = f0;
    if(voiced) {
      *lfs = lf_solve(lf_from_rd(P.rd[r], (float)(1.0 / (double)f0), 1.0f));
      double m, p; lf_spectrum(*lfs, (double)f0, &m, &p);
      *lf0 = (float)m;
      enc[2] = P.rd[r];
    } else enc[2] = 0.f;
  }
  __syncthreads();
  if(voiced) {
    const float* vt = P.vtmagn + r * (size_t)ns;
    const float lfmagnf0 = *lf0;
    // spectral synthesis (coder.c:107-115): vocal tract x source / f, then lip radiation at (j + 1) fnyq / ns
    for(int j = tid + 1; j < ns; j += nth) {
      double m, p; lf_spectrum(*lfs, (double)P.faxis[j], &m, &p);
      const float lfm = (float)m;
      senv[j] = (float)(exp((double)vt[j] * 2.3025851 / 20.0) * (double)lfm / (double)lfmagnf0 * (double)f0 / (double)P.faxis[j]);
    }
    __syncthreads();
    if(tid == 0) senv[0] = senv[1];
    __syncthreads();
    const float fstep = P.fnyq / (float)ns;
    for(int j = tid; j < ns; j += nth) {
      float v = senv[j] * coder_lip_abs(P.lip_radius, fstep, j);
      if(j >= 1) { float t = v * 44100; t = t / 4; t = t / f0; v = v * t; }    // magnitude -> PSD (coder.c:118-119)
      senv[j] = v;
      spsd[j] = spsd[j] + v;                                                 // total power
    }
    __syncthreads();
    // band aperiodicity: sequential float sums as in the reference (coder.c:124-131)
    if(tid < ob) {
      const int n0 = tid * (ns - 1) / ob, n1 = (tid + 1) * (ns - 1) / ob;
      float apsum = 0;
      for(int k = n0; k < n1; k ++) apsum = apsum + (1 - senv[k] / spsd[k]);
      enc[3 + os + tid] = apsum / (n1 - n0);
    }
  } else if(tid < ob) enc[3 + os + tid] = 1.0f;
  __syncthreads();
  // power -> log intensity, linear axis -> mel axis (coder.c:141-146)
  for(int j = tid; j < ns; j += nth) spsd[j] = (float)(log((double)spsd[j]) * 0.5);
  __syncthreads();
  for(int j = tid; j < ns; j += nth) senv[j] = coder_interp1(P.faxis, spsd, ns, P.melaxis[j]);
  __syncthreads();
  // DCT of size ns - 1, of which the order_spec-point inverse only reads order_spec numbers (coder.c:149-152)
  coder_dct(senv, ns - 1, os, coef, part);
  if(tid == 0) coef[0] = (float)((double)coef[0] * 0.5);
  __syncthreads();
  coder_idct(coef, os, os, os, spsd);
  for(int j = tid; j < os; j += nth) enc[3 + j] = (float)((double)spsd[j] * (2.0 / (ns - 1)));
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CODER_THREADS) coder_decode_kernel(CoderParams P) {
  LLSM_DYN_SMEM(smem);
  const int ns = P.nspec, os = P.order_spec, ob = P.order_bap;
  float2* bufa = (float2*)smem;                       // minimum-phase FFT buffers (layer-0 decode)
  float2* bufb = bufa + P.max_nfft;
  double* part = (double*)(bufb + P.max_nfft);        // [4][os]
  float* mel = (float*)(part + 4 * os);               // [ns]
  float* fsp = mel + ns;                              // [ns] full_psd -> full_spec
  float* fap = fsp + ns;                              // [ns] full_ap -> full_noise
  float* coef = fap + ns;                             // [os + 1]
  float* bap = coef + os + 1;                         // [ob + 1]
  float* ha = bap + ob + 1;                           // [maxnhar + 2]
  float* am = ha + P.maxnhar + 2;                     // [maxnhar]
  float* vtp = am + P.maxnhar;                        // [maxnhar]
  float* vsp = vtp + P.maxnhar;                       // [maxnhar]
  float* lfm = vsp + P.maxnhar;                       // [maxnhar] (layer 0) / [ns] is not needed at once
  LfSolved* lfs = (LfSolved*)(((uintptr_t)(lfm + P.maxnhar) + 7) & ~(uintptr_t)7);
  float* lf0 = (float*)(lfs + 1);
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t r = (size_t)b * P.nfrm + i;
  const float* src = P.enc_in + r * (os + ob + 3);
  const bool voicing = src[0] > 0.5;
  const float f0 = (float)fmax(20.0, (double)src[1]);
  const float rd = (float)fmin(3.0, fmax(0.02, (double)src[2]));
  const int nhar_ref = voicing ? (int)(P.fnyq / f0) : 0;       // coder.c:172
  const int nhar = min(nhar_ref, P.maxnhar);                   // row length of the flat arrays
  if(tid == 0) {
    P.o_f0[r] = f0 * (voicing ? 1 : 0); P.o_rd[r] = rd; P.o_nhar[r] = nhar;
    if(nhar > 0) {
      *lfs = lf_solve(lf_from_rd(rd, (float)(1.0 / (double)f0), 1.0f));
      double m, p; lf_spectrum(*lfs, (double)f0, &m, &p);
      *lf0 = (float)m;
    }
  }
  // undo the low-order IDCT, then the full-order inverse (coder.c:184-195)
  for(int j = tid; j < os; j += nth) mel[j] = (float)((double)src[3 + j] * 0.5 * (ns - 1) * 2.0 / os);
  for(int j = tid; j <= ob; j += nth) bap[j] = j == 0 ? (voicing ? 0.f : 1.f) : src[3 + os + j - 1];
  __syncthreads();
  coder_dct(mel, os, os, coef, part);
  if(tid == 0) coef[0] = (float)((double)coef[0] * 0.5);
  __syncthreads();
  coder_idct(coef, os, ns - 1, ns - 1, mel);
  for(int j = tid; j < ns - 1; j += nth) mel[j] = (float)((double)mel[j] * (2.0 / (ns - 1)));
  __syncthreads();
  if(tid == 0) mel[ns - 1] = mel[ns - 2];
  __syncthreads();
  // mel axis -> linear axis; band aperiodicity -> full aperiodicity; split the power (coder.c:200-219)
  for(int j = tid; j < ns; j += nth) {
    float psdj = coder_interp1(P.melaxis, mel, ns, P.faxis[j]);
    float ap = coder_interp1(P.apaxis, bap, ob + 1, P.faxis[j]);
    if(voicing) {
      const float fj = (float)j * P.fnyq / (float)ns;
      if(fj < 500) ap = (float)1e-3;
      else if(fj < 2000) ap = (float)(1e-3 + ((double)ap - 1e-3) * ((double)fj - 500) / 1500);
    }
    psdj = (float)exp(2.0 * (double)psdj);
    const float sum_psd = psdj;
    const float per_psd = (float)((double)sum_psd * (1.0 - (double)ap));
    float t = per_psd * f0; t = t * 4; t = t / 44100;
    fsp[j] = (float)sqrt((double)t);
    fap[j] = sum_psd * ap;
  }
  __syncthreads();
  // noise power -> scaled axis, log intensity (coder.c:223-225)
  for(int j = tid; j < P.npsd; j += nth)
    P.o_psd[r * P.npsd + j] = (float)(log((double)coder_interp1(P.faxis, fap, ns, P.psdaxis[j])) / 2.3025851 * 10.0);

  if(P.use_layer1) {
    float* vt = P.o_vtmagn + r * (size_t)ns;
    float* vs = P.o_vsphse + r * (size_t)P.maxnhar;
    if(nhar > 0) {
      // remove lip radiation and the source from the harmonic magnitude (coder.c:231-236), dB
      const float fstep = P.fnyq / (float)ns;
      const float lfmagnf0 = *lf0;
      for(int j = tid + 1; j < ns; j += nth) {
        double m, p; lf_spectrum(*lfs, (double)P.faxis[j], &m, &p);
        float v = fsp[j] / coder_lip_abs(P.lip_radius, fstep, j);
        v = v * P.faxis[j]; v = v / f0; v = v * lfmagnf0; v = v / (float)m;
        mel[j] = (float)(log((double)v) / 2.3025851 * 20.0);
      }
      __syncthreads();
      for(int j = tid; j < ns; j += nth) vt[j] = j == 0 ? mel[1] : mel[j];
      // source phases at the harmonics (coder.c:245-246): linspace(0, nhar f0, nhar + 1)[k + 1]
      for(int k = tid; k < nhar; k += nth) {
        const float freq = (float)(0.0 + ((double)((float)nhar_ref * f0) - 0.0) * (k + 1) / nhar_ref);
        double m, p; lf_spectrum(*lfs, (double)freq, &m, &p);
        vs[k] = (float)p;
      }
    } else {
      for(int j = tid; j < ns; j += nth) vt[j] = 0.f;
    }
    for(int k = nhar + tid; k < P.maxnhar; k += nth) vs[k] = 0.f;
  } else {
    float* oa = P.o_ampl + r * (size_t)P.maxnhar;
    float* op = P.o_phse + r * (size_t)P.maxnhar;
    if(nhar > 0) {
      for(int k = tid; k < nhar; k += nth) {
        const float freq = (float)(0.0 + ((double)((float)nhar_ref * f0) - 0.0) * (k + 1) / nhar_ref);
        const float a = coder_interp1(P.faxis, fsp, ns, freq);
        oa[k] = a;                                               // hm -> ampl (coder.c:257-258)
        double m, p; lf_spectrum(*lfs, (double)freq, &m, &p);
        lfm[k] = (float)m; vsp[k] = (float)p;
        am[k] = a / coder_lip_abs(P.lip_radius, f0, k);          // inverse lip radiation (coder.c:259)
      }
      __syncthreads();
      const float lfm0 = lfm[0];
      __syncthreads();
      for(int k = tid; k < nhar; k += nth) {
        const float vs_ampl = (float)((double)lfm[k] / (k + 1.0) / (double)lfm0);
        am[k] = am[k] / vs_ampl;                                 // vocal-tract magnitude (coder.c:264-267)
      }
      __syncthreads();
      block_harmonic_minphase(am, nhar, vtp, bufa, bufb, ha, P.tw, P.ntw);
      for(int k = tid; k < nhar; k += nth) {
        const float omega = (float)((double)f0 * (1.0 + k) * 2.0 * LLSM_PI);
        const float ph = vtp[k] + lip_arg(lip_response(P.lip_radius, omega));   // coder.c:270
        op[k] = ph + vsp[k];
      }
    }
    for(int k = nhar + tid; k < P.maxnhar; k += nth) { oa[k] = 0.f; op[k] = 0.f; }
  }
}
