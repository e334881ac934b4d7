// Streaming synthesis (llsmrt.c): a batch of S independent streams advanced one frame per step.
// All streams share sampling rate and hop time, hence the same hop schedule and ring positions
// (host scalars, RtClock below); per-stream state lives in device arrays:
//   mod[S][nch][cap], sin[S][cap], noise[S][cap], exc[S][cap]   ring buffers (buffer.h:32-137)
//   tmpl[S][nch][ntemplate]                                     circular noise templates (llsmrt.c:80-107)
//   prev_psd[S][npsd]                                           previous frame's noise PSD (llsmrt.c:513-520)
//
//   rt_template_kernel   stretch_stationary_noise + llsm_get_circular_noise (dsputils.c:363-383, llsmrt.c:80-91)
//   rt_warmup_kernel     llsm_fill_excitation_buffers (llsmrt.c:149-155)
//   rt_feed_kernel       one llsm_rtsynth_buffer_feed (llsmrt.c:505-521) without the L1 pulse path:
//                        ring advance (:124-128), modulation components (:255-270), sinusoids (:273-291),
//                        excitation (:134-147), STFT noise shaping (:422-478), output read-out (:480-503),
//                        previous-NM update (:513-520)
#pragma once
#include "common.cuh"
#include "kernels_synth.cuh"

struct RtFeedParams {
  int S, nchannel, maxnhar, maxnhar_e, npsd, cap, ntemplate;
  // frame row of stream s: s * row_stride + row_off (blocks of frames are fed step by step)
  int row_stride, row_off;
  const float* f0; const int* nhar; const float* ampl; const float* phse;
  const float* psd; const float* psdres; const float* edc; const int* enhar; const float* eampl; const float* ephse;
  // rings + state
  float* mod; float* sin_; float* noise; float* exc; const float* tmpl; float* prev_psd;
  int has_prev;             // a previous NM exists (every stream is fed the same number of frames)
  unsigned chan_mask;
  // shared clock of this step (RtClock)
  int H, next_nhop;         // curr_nhop, next_nhop
  int cur_old, cur_new;     // sin / noise ring position before and after appendblank
  int mod_old, mod_new;     // modulation ring position
  int exc_old, exc_new;     // excitation ring position before / after the appendchunk of H samples
  int exc_cycle;            // template read position for this step
  int sin_pos;              // read lag of the sinusoid ring
  float cycle;              // fractional position (seconds), llsmrt.c:279
  float fs;
  int nfft, lg_nfft, nspec;
  float wsqr;               // float-accumulated sum of win^2 (llsmrt.c:428-430)
  const float* win;         // hanning(2 H)
  const int* psd_lo; const float* psd_r;   // interp1 plan of llsm_spectrum_from_envelope on nspec - 1 bins
  const float2* tw;         // [nfft]
  int use_iczt; float iczt_a, iczt_b;
  int skip_sin;             // 1: the L1 path owns the sinusoid ring this step (PbP engaged)
  float* out_p; float* out_ap; int out_stride, out_off;   // [S][out_stride], next_nhop samples at out_off
};

#define RT_THREADS 256

__device__ __forceinline__ int rt_wrap(int i, int cap) { i %= cap; return i < 0 ? i + cap : i; }

__global__ void __launch_bounds__(RT_THREADS) rt_feed_kernel(RtFeedParams P) {
  LLSM_DYN_SMEM(smem);
  const int nfft = P.nfft, nspec = P.nspec, npsd = P.npsd, cap = P.cap, H = P.H, nwin = 2 * P.H;
  float2* bufa = (float2*)smem;
  float2* bufb = bufa + nfft;
  float* pbuf = (float*)(bufb + nfft);             // [nspec]
  float* spsd = pbuf + nspec;                      // [npsd]
  float* sx = spsd + npsd;                         // [nwin] scratch (excitation chunk)
  float* wmax = sx + nwin + 2;                     // [32]
  float* ca = wmax + 32;                           // [maxnhar] a cos(phi')
  float* cb = ca + P.maxnhar;                      // [maxnhar] a sin(phi')
  float* ea = cb + P.maxnhar;                      // [nch * maxnhar_e] envelope a cos(psi)
  float* eb = ea + P.nchannel * P.maxnhar_e;       // [nch * maxnhar_e] envelope a sin(psi)
  const int s = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int nch = P.nchannel;
  float* sinr = P.sin_ + (size_t)s * cap;
  float* noiser = P.noise + (size_t)s * cap;
  float* excr = P.exc + (size_t)s * cap;
  float* modr = P.mod + (size_t)s * nch * cap;

  // ---- llsm_update_cycle: appendblank on the modulation, sinusoid and noise rings (llsmrt.c:124-127)
  for(int i = tid; i < H; i += nth) {
    sinr[rt_wrap(P.cur_old + i, cap)] = 0.f;
    noiser[rt_wrap(P.cur_old + i, cap)] = 0.f;
    for(int c = 0; c < nch; c ++) modr[(size_t)c * cap + rt_wrap(P.mod_old + i, cap)] = 0.f;
  }
  __syncthreads();

  const size_t fr = (size_t)s * P.row_stride + P.row_off;
  const float f0 = P.f0[fr];
  const float f0n = f0 / P.fs;
  for(int e = tid; e < nch * P.maxnhar_e; e += nth) {
    size_t ec = fr * nch + e / P.maxnhar_e;
    float a = P.eampl[ec * P.maxnhar_e + e % P.maxnhar_e], ph = P.ephse[ec * P.maxnhar_e + e % P.maxnhar_e];
    float sp, cp; sincosf(ph, &sp, &cp);
    ea[e] = a * cp; eb[e] = a * sp;
  }
  {
    float t = P.cycle * 2.0f;
    const float phase_shift = (float)((double)t * LLSM_PI * (double)f0);            // llsmrt.c:279
    int nhs = P.nhar[fr]; if(nhs > P.maxnhar) nhs = P.maxnhar;
    for(int k = tid; k < nhs; k += nth) {
      float a = P.ampl[fr * P.maxnhar + k];
      float ph = (float)((double)P.phse[fr * P.maxnhar + k] - (double)phase_shift * ((double)k + 1.0));
      float sp, cp; sincosf(ph, &sp, &cp);
      ca[k] = a * cp; cb[k] = a * sp;
    }
  }
  __syncthreads();
  // ---- modulation components (llsmrt.c:255-270): addchunk(mod[c], -nwin, nwin)
  for(int j = tid; j < nwin; j += nth) {
    const float wj = P.win[j];
    float2 z = unit_phasor_turns((double)f0n * (double)(j - H));
    for(int c = 0; c < nch; c ++) {
      const size_t ec = fr * nch + c;
      int ne = f0 > 0 ? P.enhar[ec] : 0;
      if(ne > P.maxnhar_e) ne = P.maxnhar_e;
      float acc = 0.f;
      float2 w = make_float2(1.f, 0.f);
      for(int k = 0; k < ne; k ++) {
        w = cmul(w, z);
        acc = fmaf(ea[c * P.maxnhar_e + k], w.x, fmaf(-eb[c * P.maxnhar_e + k], w.y, acc));
      }
      float v = acc + P.edc[ec];
      if(! (v > 1e-8f)) v = 1e-8f;
      v = v * wj;
      modr[(size_t)c * cap + rt_wrap(P.mod_new - nwin + j, cap)] += v;
    }
  }
  // ---- sinusoids (llsmrt.c:273-291): addchunk(sin, -nwin, nwin)
  int nh = P.nhar[fr];
  if(nh > nfft) nh = nfft;
  if(! P.skip_sin && f0 > 0 && nh > 0) {
    bool iczt = false;
    if(P.use_iczt) iczt = log((double)nwin) * (double)P.iczt_a < log((double)nh) - (double)P.iczt_b;
    if(iczt && nh > nwin - 1) nh = nwin - 1;
    const float omega0 = (float)(2.0 * LLSM_PI * (double)f0n);
    const double nu = iczt ? (double)omega0 / (2.0 * LLSM_PI) : (double)f0n;
    for(int n = tid; n <= H; n += nth) {                                               // sample pair j = H +- n
      float2 z = unit_phasor_turns(nu * (double)n);
      float2 w = make_float2(1.f, 0.f);
      float C = 0.f, Sn = 0.f;
      for(int k = 0; k < nh; k ++) {
        w = cmul(w, z);
        C = fmaf(ca[k], w.x, C);
        Sn = fmaf(cb[k], w.y, Sn);
      }
      if(n <= H - 1) sinr[rt_wrap(P.cur_new - nwin + H + n, cap)] += (C - Sn) * P.win[H + n];
      if(n >= 1) sinr[rt_wrap(P.cur_new - nwin + H - n, cap)] += (C + Sn) * P.win[H - n];
    }
  }
  __syncthreads();

  // ---- llsm_run_excitation_buffers(dst, H) (llsmrt.c:134-147): mod chunk at lag -H - H, templates,
  //      appendchunk(exc_mix, H)
  for(int i = tid; i < H; i += nth) {
    float x = 0.f;
    for(int c = 0; c < nch; c ++) {
      if(! ((P.chan_mask >> c) & 1u)) continue;      // templates of absent channels are zero
      float m = modr[(size_t)c * cap + rt_wrap(P.mod_new - 2 * H + i, cap)];
      float tv = P.tmpl[((size_t)s * nch + c) * P.ntemplate + (P.exc_cycle + i) % P.ntemplate];
      x = (float)((double)x + sqrt((double)m) * (double)tv);
    }
    excr[rt_wrap(P.exc_old + i, cap)] = x;           // forward(H) then writechunk(-H, H)
  }
  __syncthreads();

  // ---- feed_filter (llsmrt.c:422-478) with the previous frame's noise model
  bool do_filter = P.has_prev != 0;
  if(do_filter) {
    float mx = -3.0e38f;
    for(int j = tid; j < npsd; j += nth) { float v = P.prev_psd[(size_t)s * npsd + j]; spsd[j] = v; mx = fmaxf(mx, v); }
    for(int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if((tid & 31) == 0) wmax[tid >> 5] = mx;
    for(int j = tid; j < nfft; j += nth) {
      int i = j - nfft / 2 + H;                        // x_re[i - nhop + nfft / 2] = exc[-2H + i] * win[i]
      float v = 0.f;
      if(i >= 0 && i < nwin) v = excr[rt_wrap(P.exc_new - nwin + i, cap)] * P.win[i];
      bufa[j] = make_float2(v, 0.f);
    }
    __syncthreads();
    float peak = wmax[0];
    for(int w = 1; w < (nth >> 5); w ++) peak = fmaxf(peak, wmax[w]);
    if(peak < -100.f) do_filter = false;               // uniform
  }
  if(do_filter) {
    float2* X = block_fft<false>(bufa, bufb, P.lg_nfft, P.tw, nfft);
    float2* Y = (X == bufa) ? bufb : bufa;
    for(int k = tid; k < nspec; k += nth) { float2 v = X[k]; pbuf[k] = (v.x * v.x + v.y * v.y) / P.wsqr; }
    __syncthreads();
    for(int k = tid; k < nspec - 1; k += nth) {
      int l = max(0, k - 3), u = min(nspec - 1, k + 3);
      float sm = 0.f;
      for(int q = l; q <= u; q ++) sm += pbuf[q];
      float envk = sm / (float)(u - l + 1);
      int pl = P.psd_lo[k]; float pr = P.psd_r[k];
      float hdb = spsd[pl];
      if(pr != 0.f) hdb = hdb + (spsd[pl + 1] - hdb) * pr;
      float Hg = expf(hdb * (2.3025851f / 20.0f)) / sqrtf(envk * 44100.f / P.fs + 1e-8f);
      float2 v = X[k];
      v.x *= Hg; v.y *= Hg;
      Y[k] = v;
      if(k > 0) Y[nfft - k] = make_float2(v.x, -v.y);
      if(k == nspec - 2) Y[nspec - 1] = v;
    }
    __syncthreads();
    float2* T = block_fft<true>(Y, X, P.lg_nfft, P.tw, nfft);
    const float inv = 1.0f / (float)nfft;
    for(int j = tid; j < nfft; j += nth) {
      float v = T[j].x * inv;
      if(j < 16) v *= (float)j / 16.f;
      if(j >= nfft - 16) v = (float)((double)v * (1.0 - (double)((float)(nfft - 1 - j) / 16.f)));
      noiser[rt_wrap(P.cur_new - nfft + j, cap)] += v;          // addchunk(noise, -nfft, nfft)
    }
  }
  __syncthreads();

  // ---- feed_mix (llsmrt.c:480-503): next_nhop samples of the noise ring at lag -nfft and of the sinusoid
  //      ring at lag sin_pos
  for(int i = tid; i < P.next_nhop; i += nth) {
    P.out_ap[(size_t)s * P.out_stride + P.out_off + i] = noiser[rt_wrap(P.cur_new - nfft + i, cap)];
    P.out_p[(size_t)s * P.out_stride + P.out_off + i] = sinr[rt_wrap(P.cur_new + P.sin_pos + i, cap)];
  }
  // ---- previous noise model for the next step (llsmrt.c:513-520), residual folded in
  const double resbias = 0.375 / 2.3025851 * 10.0;
  for(int j = tid; j < npsd; j += nth) {
    float v = P.psd[fr * npsd + j];
    if(P.psdres) v = (float)((double)v + ((double)P.psdres[fr * npsd + j] - resbias));
    P.prev_psd[(size_t)s * npsd + j] = v;
  }
}

// ---- templates: stretch the 20000-sample band-limited template to ntemplate samples and close the loop
struct RtTemplateParams { int nseq, nt_src, src_stride, ntemplate; const float* colored; float* tmpl; };

__global__ void __launch_bounds__(256) rt_template_kernel(RtTemplateParams P) {
  const int seq = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j >= P.ntemplate) return;
  const float* x = P.colored + (size_t)seq * P.src_stride;
  const int nx_src = P.nt_src - 128;                   // template length before the 128-sample extension
  // llsm_get_circular_noise (llsmrt.c:80-91) over the stretched sequence y[0 .. ntemplate)
  const int overlap = 32;
  float y = stretched_value(x, stretch_index(nx_src, P.ntemplate, j));
  if(j < overlap) {
    float r = (float)j / (float)overlap;
    float y2 = stretched_value(x, stretch_index(nx_src, P.ntemplate, P.ntemplate - overlap + j));
    y = (float)((double)y * (1.0 - (double)r));
    y = y + y2 * r;
    float d = 2.0f * r; d = d * (r - 1.0f); d = d + 1.0f;
    y = (float)((double)y / sqrt((double)d));
  }
  P.tmpl[(size_t)seq * P.ntemplate + j] = y;
}

// ---- warm-up: modulation rings at 1e-5 (all but one slot), excitation ring filled from the templates
struct RtWarmParams { int S, nchannel, cap, ntemplate; unsigned chan_mask; float* mod; float* sin_; float* noise;
  float* exc; const float* tmpl; int hole; /* index left at zero in the modulation rings */ };

__global__ void __launch_bounds__(256) rt_warmup_kernel(RtWarmParams P) {
  const int s = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if(q >= P.cap) return;
  for(int c = 0; c < P.nchannel; c ++) P.mod[((size_t)s * P.nchannel + c) * P.cap + q] = q == P.hole ? 0.f : 1e-5f;
  P.sin_[(size_t)s * P.cap + q] = 0.f;
  P.noise[(size_t)s * P.cap + q] = 0.f;
  // 5 x llsm_run_excitation_buffers(ninternal / 5): exc[q] = sum_c sqrt(1e-5f) * tmpl_c[q % ntemplate]
  float x = 0.f;
  const int chunk = P.cap / 5;
  if(q < 5 * chunk) {
    for(int c = 0; c < P.nchannel; c ++) {
      if(! ((P.chan_mask >> c) & 1u)) continue;
      float tv = P.tmpl[((size_t)s * P.nchannel + c) * P.ntemplate + q % P.ntemplate];
      x = (float)((double)x + sqrt((double)1e-5f) * (double)tv);
    }
  }
  P.exc[(size_t)s * P.cap + q] = x;
}

// ---- llsm_rtsynth_buffer_clear (llsmrt.c:578-602): the modulation rings survive a clear and only get the
//      appendblank of the re-primed clock; sinusoid / noise / excitation rings restart from zero
struct RtClearParams { int S, nchannel, cap, mod_old, H; float* mod; float* sin_; float* noise; float* exc; };

__global__ void __launch_bounds__(256) rt_clear_kernel(RtClearParams P) {
  const int s = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if(q >= P.cap) return;
  P.sin_[(size_t)s * P.cap + q] = 0.f;
  P.noise[(size_t)s * P.cap + q] = 0.f;
  P.exc[(size_t)s * P.cap + q] = 0.f;
  int d = q - P.mod_old; if(d < 0) d += P.cap;
  if(d < P.H)
    for(int c = 0; c < P.nchannel; c ++) P.mod[((size_t)s * P.nchannel + c) * P.cap + q] = 0.f;
}
