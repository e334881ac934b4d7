// Layer-1 conversion kernels (BASELINE config 4 and the on-the-fly L1 -> L0 of PbP synthesis).
//
//   rd_fit_kernel        llsm_analyze_rd per frame (layer1.c:60-71): inverse lip radiation,
//                        llsm_spectral_glottal_fitting (dsputils.c:545-575) against the cached LF
//                        power spectra (dsputils.c:514-535)
//   rd_smooth_kernel     interp_in_blank + llsm_smoothing_filter along time (layer1.c:74-76,
//                        dsputils.c:578-604)
//   tolayer1_kernel      llsm_frame_tolayer1 (layer1.c:86-127): source / lip removal, minimum-phase
//                        vocal tract (llsm_harmonic_minphase dsputils.c:481-505), source phases,
//                        spectral envelope (llsm_harmonic_envelope dsputils.c:465-479 with
//                        llsm_harmonic_spectrum :432-453 and the cepstral smoother)
//   tolayer0_kernel      llsm_frame_tolayer0 (layer1.c:151-195)
#pragma once
#include "common.cuh"
#include "lf_model.cuh"

// ---- lip radiation response at angular frequency omega (dsputils.c:396-430), FP_TYPE arithmetic
__device__ __forceinline__ float2 lip_response(float radius, float omega) {
  const float Rr = (float)(128.0 / 9.0 / LLSM_PI / LLSM_PI);
  const float Lr = (float)(8.0 * (double)radius / 100.0 / 3.0 / LLSM_PI / 340.0);
  float ar = __fmul_rn(__fmul_rn(omega, Lr), Rr);      // numerator (real)
  float br = Rr, bi = __fmul_rn(omega, Lr);            // denominator
  float d = __fadd_rn(__fmul_rn(br, br), __fmul_rn(bi, bi));
  float qr = __fadd_rn(__fmul_rn(ar, br), 0.0f) / d;   // (a.re b.re + a.im b.im) / d, a.im = 0
  float qi = __fadd_rn(0.0f, -__fmul_rn(ar, bi)) / d;  // (a.im b.re - a.re b.im) / d
  return make_float2(-qi, qr);                          // times i
}
__device__ __forceinline__ float lip_abs(float2 r) {
  return (float)sqrt((double)__fadd_rn(__fmul_rn(r.x, r.x), __fmul_rn(r.y, r.y)));
}
__device__ __forceinline__ float lip_arg(float2 r) { return (float)atan2((double)r.y, (double)r.x); }

// ---- block-cooperative llsm_harmonic_minphase (dsputils.c:481-505) ---------------------------
// ampl[nhar] (shared or global) -> out[nhar] (shared). bufa / bufb: 2 x nfft float2 of shared memory,
// ha: (nhar + 1) floats of shared scratch. nfft = max(64, 2^(ceil(log2 nhar) + 2)).
__device__ __forceinline__ int minphase_nfft(int nhar) {
  int n = pow2_ceil(log2((double)nhar) + 2.0);
  return n > 64 ? n : 64;
}

__device__ void block_harmonic_minphase(const float* ampl, int nhar, float* out, float2* bufa, float2* bufb,
  float* ha, const float2* __restrict__ tw, int ntw) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int nfft = minphase_nfft(nhar);
  int lg = 0; while((1 << lg) < nfft) lg ++;
  const int ns = nfft / 2 + 1;
  // har_idx[i + 1] = (i + 1.0) / (nhar + 1.0) * nfft / 2.0 ; har_ampl[i + 1] = log(ampl[i] + 1e-10)
  for(int i = tid; i < nhar; i += nth) ha[i + 1] = (float)log((double)ampl[i] + 1e-10);
  __syncthreads();
  if(tid == 0) ha[0] = ha[1];
  __syncthreads();
  const float hlast = (float)(((double)nhar) / (nhar + 1.0) * nfft / 2.0);
  const float hprev = nhar >= 2 ? (float)(((double)nhar - 1.0) / (nhar + 1.0) * nfft / 2.0) : 0.f;
  const float x1 = __fadd_rn(__fmul_rn(hlast, 2.0f), -hprev);     // har_idx[nhar] * 2 - har_idx[nhar - 1]
  const double step = (double)x1 / (nhar + 1);
  for(int k = tid; k < ns; k += nth) {                            // interp1u onto the FFT grid
    double p = (double)(float)k / step;
    float v;
    if(! (p > 0)) v = ha[0];
    else if(p >= nhar) v = ha[nhar];
    else { int q = (int)p; double r = p - q; v = (float)((double)ha[q] + ((double)ha[q + 1] - (double)ha[q]) * r); }
    bufa[k] = make_float2(v, 0.f);
    if(k > 0 && k < nfft / 2) bufa[nfft - k] = make_float2(v, 0.f);
  }
  __syncthreads();
  float2* C = block_fft<true>(bufa, bufb, lg, tw, ntw);           // cepstrum * nfft
  float2* D = (C == bufa) ? bufb : bufa;
  for(int q = tid; q < nfft; q += nth) {
    float c = C[q].x / (float)nfft;
    if(q > 0 && q < nfft / 2) c *= 2.f; else if(q > nfft / 2) c = 0.f;
    D[q] = make_float2(c, 0.f);
  }
  __syncthreads();
  float2* Ph = block_fft<false>(D, C, lg, tw, ntw);               // phase = imaginary part
  // har_phse = interp1u(0, nfft/2 + 1, phase, nfft/2 + 1, har_idx, nhar + 1) then the shifted copy
  // har_phse[i - 1] = har_phse[i], i = 1 .. nhar - 1 (the last entry keeps index nhar - 1)
  for(int k = tid; k < nhar; k += nth) {
    int src = k <= nhar - 2 ? k + 1 : nhar - 1;
    if(nhar == 1) src = 0;
    float hx = (float)(((double)src) / (nhar + 1.0) * nfft / 2.0);
    if(src == 0) hx = 0.f;
    double p = (double)hx;                                         // step = 1
    float v;
    if(! (p > 0)) v = Ph[0].y;
    else if(p >= ns - 1) v = Ph[ns - 1].y;
    else { int q = (int)p; double r = p - q; v = (float)((double)Ph[q].y + ((double)Ph[q + 1].y - (double)Ph[q].y) * r); }
    out[k] = v;
  }
  __syncthreads();
}

// ---- L1 -> L0 -----------------------------------------------------------------------------------
struct ToLayer0Params {
  int nfrm; const int* nfrm_utt;
  const float* f0; const float* rd;        // [B][nfrm]
  const float* vtmagn; int nspec;          // [B][nfrm][nspec] dB
  const float* vsphse; const int* nvs;     // [B][nfrm][vs_stride], lengths [B][nfrm]
  int vs_stride;
  int maxnhar;                             // LLSM_CONF_MAXNHAR (cap) and row length of the outputs
  float fnyq, lip_radius;
  const float2* tw; int ntw;               // twiddles for the min-phase FFT (ntw >= 4 * 2^ceil(log2 maxnhar))
  int max_nfft;
  int* nhar_out; float* ampl; float* phse; // [B][nfrm](, [maxnhar])
};

#define L1_THREADS 128

__global__ void __launch_bounds__(L1_THREADS) tolayer0_kernel(ToLayer0Params P) {
  LLSM_DYN_SMEM(smem);
  float2* bufa = (float2*)smem;
  float2* bufb = bufa + P.max_nfft;
  float* ha = (float*)(bufb + P.max_nfft);       // [maxnhar + 2]
  float* vt = ha + P.maxnhar + 2;                // [maxnhar] vocal-tract amplitude
  float* vs = vt + P.maxnhar;                    // [maxnhar] source amplitude
  float* ph = vs + P.maxnhar;                    // [maxnhar] min-phase
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t r = (size_t)b * P.nfrm + i;
  const float f0 = P.f0[r];
  int nhar = 0;
  if(f0 != 0) {
    nhar = P.nvs[r];
    if(nhar > P.maxnhar) nhar = P.maxnhar;
    int cap = (int)(P.fnyq / f0);
    if(nhar > cap) nhar = cap;
  }
  if(nhar <= 0) {
    if(tid == 0) P.nhar_out[r] = 0;
    for(int k = tid; k < P.maxnhar; k += nth) { P.ampl[r * P.maxnhar + k] = 0; P.phse[r * P.maxnhar + k] = 0; }
    return;
  }
  const LfSolved lf = lf_solve(lf_from_rd(P.rd[r], (float)(1.0 / (double)f0), 1.0f));
  const float* env = P.vtmagn + r * (size_t)P.nspec;
  for(int k = tid; k < nhar; k += nth) {
    float freq = (float)((double)f0 * (k + 1.0));
    double m, p; lf_spectrum(lf, (double)freq, &m, &p);
    vs[k] = (float)m;
    // vt_ampl = exp(DB2LOG(interp1(linspace(0, fnyq, nspec), VTMAGN, freq)))   layer1.c:177-180
    float v;
    const int ns = P.nspec;
    float xlast = P.fnyq;
    if(! (freq > 0.f)) v = env[0];
    else if(freq >= xlast) v = env[ns - 1];
    else {
      int lo = 0, hi = ns - 1;
      while(hi - lo > 1) {
        int mid = (lo + hi) / 2;
        float xm = (float)(0.0 + ((double)P.fnyq - 0.0) * mid / (ns - 1));
        if(xm <= freq) lo = mid; else hi = mid;
      }
      float xl = (float)(((double)P.fnyq) * lo / (ns - 1)), xh = (float)(((double)P.fnyq) * hi / (ns - 1));
      double rr = ((double)freq - xl) / ((double)xh - xl);
      v = (float)((double)env[lo] + ((double)env[hi] - (double)env[lo]) * rr);
    }
    vt[k] = (float)exp((double)v * 2.3025851 / 20.0);
  }
  __syncthreads();
  const float vs0 = vs[0];
  __syncthreads();
  for(int k = tid; k < nhar; k += nth)
    vs[k] = k == 0 ? 1.0f : (float)((double)vs[k] / ((1.0 + k) * (double)vs0));   // layer1.c:174-175
  block_harmonic_minphase(vt, nhar, ph, bufa, bufb, ha, P.tw, P.ntw);
  for(int k = tid; k < nhar; k += nth) {
    float a = vt[k] * vs[k];
    float p = ph[k] + P.vsphse[r * P.vs_stride + k];
    float omega = (float)((double)f0 * (1.0 + k) * 2.0 * LLSM_PI);
    float2 ir = lip_response(P.lip_radius, omega);
    a = a * lip_abs(ir);
    p = p + lip_arg(ir);
    P.ampl[r * P.maxnhar + k] = a;
    P.phse[r * P.maxnhar + k] = p;
  }
  for(int k = nhar + tid; k < P.maxnhar; k += nth) { P.ampl[r * P.maxnhar + k] = 0; P.phse[r * P.maxnhar + k] = 0; }
  if(tid == 0) P.nhar_out[r] = nhar;
}

// ---- Rd fitting -----------------------------------------------------------------------------------
struct RdFitParams {
  int nfrm; const int* nfrm_utt;
  const float* f0; const int* nhar; const float* ampl; int maxnhar;
  float lip_radius;
  const float* model;       // [64][80] cached squared LF amplitudes / (j + 1)^2  (dsputils.c:526-531)
  const float* rd_list;     // [64] linspace(0.02, 3.0, 64)
  float* rd;                // [B][nfrm], 0 for unvoiced frames
};

#define RD_NCAND 64
#define RD_NHAR 80

__global__ void __launch_bounds__(RD_NCAND) rd_fit_kernel(RdFitParams P) {
  __shared__ float pw[RD_NHAR];
  __shared__ float dist[RD_NCAND];
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t r = (size_t)b * P.nfrm + i;
  const float f0 = P.f0[r];
  if(f0 == 0) { if(tid == 0) P.rd[r] = 0.f; return; }
  int nhar = P.nhar[r];
  int cap = (int)round(8000.0 / (double)f0);                       // layer1.c:65
  if(nhar > cap) nhar = cap;
  int nfit = nhar < RD_NHAR ? nhar : RD_NHAR;
  for(int k = tid; k < nfit; k += blockDim.x) {
    float a = P.ampl[r * P.maxnhar + k];
    float omega = (float)((double)f0 * (1.0 + k) * 2.0 * LLSM_PI);
    a = a / lip_abs(lip_response(P.lip_radius, omega));            // inverse lip filter, layer1.c:68
    pw[k] = a * a;
  }
  __syncthreads();
  {
    const float* md = P.model + tid * RD_NHAR;
    float gain = pw[0] / md[0];
    double acc = 0;
    for(int j = 0; j < nfit; j ++) {
      float pm = md[j] * gain;
      double q = (double)pw[j] / (double)pm;
      acc += q - log(q) - 1.0;
    }
    float is = nfit > 0 ? (float)(acc / nfit) : 0.f;               // itakura_saito returns FP_TYPE
    dist[tid] = (float)exp((double)is);
  }
  __syncthreads();
  if(tid == 0) {
    int v = 0;
    for(int c = 1; c < RD_NCAND; c ++) if(dist[c] < dist[v]) v = c;
    float rd = P.rd_list[v];
    if(v > 0 && v < RD_NCAND - 1) {                                // parabolic refinement, dsputils.c:568-572
      double a = dist[v - 1], bq = dist[v], c = dist[v + 1];
      double a1 = (a + c) * 0.5 - bq, a2 = (c - a) * 0.5;
      double x = a1 != 0 ? -a2 / (2.0 * a1) : 0;
      if(! (fabs(x) < 1.0)) x = 0;
      float pos = (float)(v + x);
      int ip = (int)pos;
      float fr = (float)fmod((double)pos, 1.0);
      rd = P.rd_list[ip] + (P.rd_list[ip + 1] - P.rd_list[ip]) * fr;
    }
    P.rd[r] = rd;
  }
}

// One CTA per utterance: fill unvoiced gaps by linear interpolation (hold at the ends), then the
// impulse-insensitive moving average of order round(0.02 / thop).
struct RdSmoothParams { int nfrm; const int* nfrm_utt; float* rd; int order; };

__global__ void __launch_bounds__(256) rd_smooth_kernel(RdSmoothParams P) {
  LLSM_DYN_SMEM(smem);
  float* x = (float*)smem;            // [nfrm] gap-filled
  float* raw = x + P.nfrm;            // [nfrm]
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int n = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  float* rd = P.rd + (size_t)b * P.nfrm;
  for(int i = tid; i < n; i += nth) raw[i] = rd[i];
  __syncthreads();
  for(int i = tid; i < n; i += nth) {
    float v = raw[i];
    if(v == 0.f) {                    // interp_in_blank(rd, nfrm, 0)
      int l = i - 1; while(l >= 0 && raw[l] == 0.f) l --;
      int u = i + 1; while(u < n && raw[u] == 0.f) u ++;
      if(l < 0 && u >= n) v = 0.f;
      else if(l < 0) v = raw[u];
      else if(u >= n) v = raw[l];
      else v = (float)((double)raw[l] + ((double)raw[u] - (double)raw[l]) * (i - l) / (u - l));
    }
    x[i] = v;
  }
  __syncthreads();
  const int order = P.order;
  for(int i = tid; i < n; i += nth) {
    float y;
    if(n < order || order <= 0) y = x[i];
    else if(i < order / 2) { double s = 0; for(int j = 0; j < order; j ++) s += x[j]; y = (float)(s / order); }
    else if(i >= n - order / 2) { double s = 0; for(int j = 0; j < order; j ++) s += x[n - order + j]; y = (float)(s / order); }
    else {
      int l = i - order / 2;
      double s = 0; for(int j = 0; j < order; j ++) s += x[l + j];
      float mean = (float)(s / order);
      int npos = 0, nneg = 0; float dt = 0.f;
      for(int j = l; j < l + order; j ++) {
        npos += x[j] >= mean; nneg += x[j] <= mean;
        float d = x[j] - mean; dt = dt + (d > 0.f ? d : 0.f);
      }
      y = mean + (float)(npos - nneg) * dt / (float)order / (float)order;
    }
    rd[i] = y;
  }
}

// ---- L0 -> L1 per frame ------------------------------------------------------------------------------
struct ToLayer1Params {
  int nfrm; const int* nfrm_utt;
  const float* f0; const float* rd; const int* nhar; const float* ampl; const float* phse; int maxnhar;
  float fnyq, lip_radius;
  int nfft, lg_nfft, nspec;                // envelope transform (llsm_chunk_tolayer1's nfft)
  const float2* tw; int ntw;               // twiddles, ntw >= max(nfft, min-phase nfft)
  int max_nfft;
  float* vtmagn; float* vsphse; int* nvs;  // [B][nfrm][nspec], [B][nfrm][maxnhar], [B][nfrm]
};

__device__ __forceinline__ float dirichlet(float M, float omega) {   // safe_aliased_sinc
  double d = sin(0.5 * (double)omega);
  if(fabs(d) < 1e-9) return M;
  return (float)(sin(0.5 * (double)M * (double)omega) / d);
}

__global__ void __launch_bounds__(L1_THREADS) tolayer1_kernel(ToLayer1Params P) {
  LLSM_DYN_SMEM(smem);
  float2* bufa = (float2*)smem;
  float2* bufb = bufa + P.max_nfft;
  float* ha = (float*)(bufb + P.max_nfft);       // [maxnhar + 2]
  float* am = ha + P.maxnhar + 2;                // [maxnhar] vocal-tract amplitudes
  float* ph = am + P.maxnhar;                    // [maxnhar] min-phase
  float* cm = ph + P.maxnhar;                    // [maxnhar] compressed amplitudes
  float* red = cm + P.maxnhar;                   // [L1_THREADS]
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t r = (size_t)b * P.nfrm + i;
  const float f0 = P.f0[r];
  const int nhar = f0 != 0 ? P.nhar[r] : 0;
  if(nhar <= 0) {                                // unvoiced: only RD is attached (layer1.c:143-145)
    if(tid == 0) P.nvs[r] = 0;
    for(int k = tid; k < P.nspec; k += nth) P.vtmagn[r * P.nspec + k] = 0;
    for(int k = tid; k < P.maxnhar; k += nth) P.vsphse[r * P.maxnhar + k] = 0;
    return;
  }
  const LfSolved lf = lf_solve(lf_from_rd(P.rd[r], (float)(1.0 / (double)f0), 1.0f));
  // source amplitudes (vs_ampl), inverse lip filter, division -> vocal-tract amplitudes
  for(int k = tid; k < nhar; k += nth) {
    float freq = (float)((double)f0 * (k + 1.0));
    double m, p; lf_spectrum(lf, (double)freq, &m, &p);
    ha[k] = (float)m;                            // raw |LF| for now
  }
  __syncthreads();
  const float vs0 = ha[0];
  __syncthreads();
  for(int k = tid; k < nhar; k += nth) {
    float vsk = k == 0 ? 1.0f : (float)((double)ha[k] / ((1.0 + k) * (double)vs0));
    float omega = (float)((double)f0 * (1.0 + k) * 2.0 * LLSM_PI);
    float2 ir = lip_response(P.lip_radius, omega);
    float a = P.ampl[r * P.maxnhar + k] / lip_abs(ir);
    a = a / vsk;
    am[k] = a;
    cm[k] = P.phse[r * P.maxnhar + k] - lip_arg(ir);      // phase after inverse lip filter (parked)
  }
  __syncthreads();
  block_harmonic_minphase(am, nhar, ph, bufa, bufb, ha, P.tw, P.ntw);
  for(int k = tid; k < nhar; k += nth) P.vsphse[r * P.maxnhar + k] = cm[k] - ph[k];   // layer1.c:110
  for(int k = nhar + tid; k < P.maxnhar; k += nth) P.vsphse[r * P.maxnhar + k] = 0;
  if(tid == 0) P.nvs[r] = nhar;
  __syncthreads();

  // ---- spectral envelope (dsputils.c:465-479): compress, paint Hann main lobes, cepstral smoothing
  float mx = -3.0e38f;
  for(int k = tid; k < nhar; k += nth) mx = fmaxf(mx, am[k]);
  red[tid] = mx;
  __syncthreads();
  for(int o = nth >> 1; o > 0; o >>= 1) { if(tid < o) red[tid] = fmaxf(red[tid], red[tid + o]); __syncthreads(); }
  const float peak = (float)log((double)red[0]);
  for(int k = tid; k < nhar; k += nth) {
    float lg = (float)log((double)am[k]) - peak;
    float c = lg > -10.f ? lg : (float)(((double)lg + 10.0) / 2 - 10.0);
    cm[k] = (float)exp((double)c);
  }
  __syncthreads();
  const int nfft = P.nfft, nX = P.nspec;
  const float f0n = (float)((double)(f0 / P.fnyq) / 2.0);          // f0 / fnyq / 2.0
  const int T = (int)(3.0 / (double)f0n);
  const int width = (int)ceil((double)__fmul_rn(f0n, (float)nfft) * 1.5);
  for(int j = tid; j < nX; j += nth) {
    // harmonics whose painted range [center - width, center + width] contains bin j
    float X = 0.f;
    int klo = (int)floor(((double)j - width - 1.0) / ((double)f0n * nfft)) - 1; if(klo < 0) klo = 0;
    int khi = (int)ceil(((double)j + width + 1.0) / ((double)f0n * nfft)); if(khi > nhar - 1) khi = nhar - 1;
    for(int k = klo; k <= khi; k ++) {
      float ifreq = (float)((double)f0n * (1.0 + k));
      int center = (int)round((double)__fmul_rn(ifreq, (float)nfft));
      int lo = center - width; if(lo < 0) lo = 0;
      int hi = center + width + 1; if(hi > nX) hi = nX;
      if(j < lo || j >= hi) continue;
      float omega = (float)((double)__fadd_rn((float)j / (float)nfft, -ifreq) * 2.0 * LLSM_PI);
      float resp = (float)(0.5 * (double)dirichlet((float)T, omega) +
                           0.25 * (double)dirichlet((float)T, (float)((double)omega - 2.0 * LLSM_PI / T)) +
                           0.25 * (double)dirichlet((float)T, (float)((double)omega + 2.0 * LLSM_PI / T)));
      X = fmaxf(X, resp * cm[k]);
    }
    X = X * f0n;
    // cepstral smoother input: log(max(S, 1e-10)), Hermitian extension
    float lg = (float)log((double)(X > 1e-10f ? X : 1e-10f));
    bufa[j] = make_float2(lg, 0.f);
    if(j > 0 && j < nfft / 2) bufa[nfft - j] = make_float2(lg, 0.f);
  }
  __syncthreads();
  float2* Cq = block_fft<true>(bufa, bufb, P.lg_nfft, P.tw, P.ntw);
  float2* D = (Cq == bufa) ? bufb : bufa;
  for(int q = tid; q <= nfft / 2; q += nth) {
    double xq = (double)f0n * q;
    double sinc = 1.0;
    if(q > 0) { double s, c; sincospi(xq, &s, &c); sinc = s / (LLSM_PI * xq); }
    double s2, c2; sincospi(2.0 * xq, &s2, &c2);
    float cv = (float)((double)Cq[q].x / nfft * sinc * (1.18 - 0.18 * c2));
    D[q] = make_float2(cv, 0.f);
    if(q > 0 && q < nfft / 2) D[nfft - q] = make_float2(cv, 0.f);
  }
  __syncthreads();
  float2* Ev = block_fft<false>(D, Cq, P.lg_nfft, P.tw, P.ntw);
  for(int j = tid; j < nX; j += nth) {
    float e = Ev[j].x;
    float dc = e > -10.f ? e : (float)(((double)e + 10.0) * 2 - 10.0);
    P.vtmagn[r * P.nspec + j] = (float)(((double)dc + (double)peak) / 2.3025851 * 20.0);   // LOG2DB
  }
}
