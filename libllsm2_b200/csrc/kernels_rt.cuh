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
#include "kernels_pbp.cuh"

// ---- pulse-by-pulse streaming (use_l1): per-stream tracker state and the plan of one feed step --------
struct RtPbpState { float pulse; int state; int offset; };      // llsmrt.c:44,52-53: pulse, pbp_state, pbp_offset
struct RtPbpPlan {
  int do_sin;                 // feed the harmonic-model frame to the sinusoid ring
  int npulse, pulse_base, pre_rotate, pulse_size;
  int read_main, main_offset; // windowed copy pulse buffer -> sinusoid ring (llsmrt.c:395-403)
  int termination, term_offset;   // trapezoid catch-up copy (llsmrt.c:404-419)
  int overflow;               // more pulses than the plan holds
  PbpPulse pulses[PBP_MAXP];
};

#if defined(__CUDA_ARCH__)
#define RT_FMUL(a, b) __fmul_rn((a), (b))
#define RT_FADD(a, b) __fadd_rn((a), (b))
#else
#define RT_FMUL(a, b) ((a) * (b))
#define RT_FADD(a, b) ((a) + (b))
#endif

// The L1 part of llsm_update_cycle (llsmrt.c:114-116) and of llsm_rtsynth_buffer_feed_deterministic
// (llsmrt.c:305-392) that is sequential per stream: pulse tracker locked on the first source harmonic,
// HM <-> PbP onset / termination, the pulse list. `mod` is the per-pulse llsm_pbpeffect hook (identity on
// the device; the drop-in runs this on the host when a frame carries an effect).
template <class Mod>
LF_HD void rt_pbp_track(RtPbpState& st, RtPbpPlan& pl, int prev_nhop, int nhop, int sin_pos, float fs, int nspec,
  float f0, int has_l1, float rd, float vsphse0, int pbp_on, Mod& mod, int frame) {
  st.pulse = st.pulse - (float)prev_nhop;
  if(st.state && st.offset > sin_pos + nhop) st.offset -= prev_nhop;
  pl.do_sin = 0; pl.npulse = 0; pl.pulse_base = 0; pl.pre_rotate = 0; pl.pulse_size = 0;
  pl.read_main = 0; pl.main_offset = 0; pl.termination = 0; pl.term_offset = 0; pl.overflow = 0;
  if(! has_l1 || f0 == 0) return;                                  // llsmrt.c:312
  float len_period = fs / f0;
  const float t_period = (float)(1.0 / (double)f0);
  const LfModel sm = lf_from_rd(rd, t_period, 1.0f);
  float source_p0; {
    LfSolved ss = lf_solve(sm);
    double m, ph; lf_spectrum(ss, (double)f0, &m, &ph);
    source_p0 = (float)ph;
    source_p0 = (float)((double)source_p0 - 0.5 * LLSM_PI);
  }
  float p0; {
    double pv = (double)vsphse0;
    double q = pv - 2.0 * LLSM_PI * floor((pv + LLSM_PI) / (2.0 * LLSM_PI));
    if(q <= -LLSM_PI) q += 2.0 * LLSM_PI;
    p0 = (float)q;
  }
  float p0_dist; {
    double pv = (double)(p0 - source_p0);
    double q = pv - 2.0 * LLSM_PI * floor((pv + LLSM_PI) / (2.0 * LLSM_PI));
    if(q <= -LLSM_PI) q += 2.0 * LLSM_PI;
    p0_dist = (float)q;
  }
  if(p0_dist < 0) p0_dist = (float)((double)p0_dist + 2.0 * LLSM_PI);
  const float pulse_projected = (float)((double)(p0_dist / 2.0f) / LLSM_PI * (double)len_period);
  const int len_reset = (int)RT_FMUL((len_period > (float)nhop ? len_period : (float)nhop), 2.0f);
  if(pulse_projected - st.pulse > (float)len_reset) st.pulse = pulse_projected - (float)len_reset;
  int num_periods = (int)round((double)((pulse_projected - st.pulse) / len_period));
  if(num_periods > 0) len_period = (pulse_projected - st.pulse) / (float)num_periods;
  float mxs = RT_FMUL(len_period, 2.0f); if(! (mxs > (float)nspec)) mxs = (float)nspec;
  const int pulse_size = pow2_ceil(log2((double)mxs));
  int onset = 0, termination = 0;
  if(pbp_on && ! st.state) {                                       // llsmrt.c:343-350
    onset = 1; st.state = 1; st.offset = -nhop;
    pl.do_sin = 1;
  }
  if(! pbp_on && st.state) {                                       // llsmrt.c:351-355
    termination = 1; st.state = 0;
    num_periods = (int)((double)num_periods + ceil((double)((float)(-st.offset) / len_period)));
  }
  if(st.state || termination) {                                    // llsmrt.c:358-386
    const int period_begin = onset ? -2 : 0;
    const int num_pulses = num_periods - period_begin;
    const float nh2 = (float)(nhop * 2);
    const int pre_rotate = (int)(len_period < nh2 ? len_period : nh2);
    if(num_pulses > 0) {
      const int np = num_pulses < PBP_MAXP ? num_pulses : PBP_MAXP;
      for(int i = 0; i < num_pulses; i ++) {
        LfModel m = sm; float delta_t = 0.f;
        mod(m, delta_t, frame);
        if(i >= np) continue;
        float off = RT_FADD(st.pulse, RT_FMUL((float)(i + period_begin), len_period));
        off = RT_FADD(off, RT_FMUL(delta_t, fs));
        pl.pulses[i].T0 = m.T0; pl.pulses[i].te = m.te; pl.pulses[i].tp = m.tp; pl.pulses[i].ta = m.ta;
        pl.pulses[i].Ee = m.Ee; pl.pulses[i].offset = off;
      }
      const int pulse_base = (int)pl.pulses[0].offset;
      for(int i = 0; i < np; i ++) pl.pulses[i].offset = pl.pulses[i].offset - (float)pulse_base;
      pl.npulse = np; pl.overflow = num_pulses > PBP_MAXP;
      pl.pulse_base = pulse_base; pl.pre_rotate = pre_rotate; pl.pulse_size = pulse_size;
    }
  }
  if(! st.state) pl.do_sin = 1;                                    // llsmrt.c:387-391
  st.pulse = pulse_projected;
  if(st.state && st.offset <= sin_pos + nhop) { pl.read_main = 1; pl.main_offset = st.offset; }
  if(termination) { pl.termination = 1; pl.term_offset = st.offset; }
}

struct RtL1Params {
  const float* rd; const float* vtmagn; int nspec; const float* vsphse; const int* nvs; int vs_stride;
  const int* pbpsyn;                      // [S][nfeed] or NULL
  float* pulse_f; float* pulse_b;         // llsm_dualbuffer (buffer.h:146-209): forward / backward halves [S][cap]
  RtPbpState* state;                      // [S]   device tracker state
  const RtPbpPlan* plans;                 // [S]   plans made on the host (effect callbacks), or NULL
  int prev_nhop;
  float fnyq, lip_radius;
  int max_size, maxnhar_vs;               // pulse FFT size limit, harmonic rows of the pulse scratch
  const float2* tw_p; int ntw_p;
};

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
  float* out_p; float* out_ap; int out_stride, out_off;   // [S][out_stride], next_nhop samples at out_off
};

#define RT_THREADS 256

__device__ __forceinline__ int rt_wrap(int i, int cap) { i %= cap; return i < 0 ? i + cap : i; }

template <bool L1>
__global__ void __launch_bounds__(RT_THREADS) rt_feed_kernel(RtFeedParams P, RtL1Params Q) {
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
    if(L1) {                                           // llsm_dualbuffer_forward (buffer.h:186-192)
      const size_t at = (size_t)s * cap + rt_wrap(P.cur_old + i, cap);
      Q.pulse_b[at] = Q.pulse_f[at]; Q.pulse_f[at] = 0.f;
    }
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
  // ---- layer-1 path: tracker step, filtered pulses into the dual buffer (llsmrt.c:305-386)
  __shared__ RtPbpPlan plan;
  __shared__ LfSolved solved[PBP_MAXP];
  bool do_sin = true;
  if(L1) {
    if(tid == 0) {
      if(Q.plans) plan = Q.plans[s];
      else {
        RtPbpState st = Q.state[s];
        PbpNoEffect none;
        const int has = Q.nvs[fr] > 0;
        rt_pbp_track(st, plan, Q.prev_nhop, H, P.sin_pos, P.fs, Q.nspec, f0, has, Q.rd[fr], Q.vsphse[fr * Q.vs_stride],
          Q.pbpsyn ? (Q.pbpsyn[fr] == 1) : 0, none, 0);
        Q.state[s] = st;
      }
    }
    __syncthreads();
    do_sin = plan.do_sin != 0;
    if(plan.npulse > 0 && plan.pulse_size <= Q.max_size) {
      PulseSmem M;
      M.bufa = (float2*)smem; M.bufb = M.bufa + Q.max_size;
      M.ha = (float*)(M.bufb + Q.max_size); M.vta = M.ha + Q.maxnhar_vs + 2; M.vtp = M.vta + Q.maxnhar_vs;
      M.pre = M.vtp + Q.maxnhar_vs; M.pim = M.pre + Q.maxnhar_vs + 1; M.solved = solved;
      PulseFrame F;
      F.f0 = f0; F.rd = Q.rd[fr]; F.vtmagn = Q.vtmagn + fr * (size_t)Q.nspec; F.nspec = Q.nspec;
      F.vsphse = Q.vsphse + fr * (size_t)Q.vs_stride;
      F.nhar = Q.nvs[fr] < Q.maxnhar_vs ? Q.nvs[fr] : Q.maxnhar_vs;
      F.fs = P.fs; F.fnyq = Q.fnyq; F.lip_radius = Q.lip_radius; F.tw = Q.tw_p; F.ntw = Q.ntw_p;
      float2* T = pbp_make_pulse(F, plan.pulses, plan.npulse, plan.pre_rotate, plan.pulse_size, M);
      // llsm_dualbuffer_addchunk(buffer_pulse, pulse_base - pre_rotate - nhop, pulse_size, y) (llsmrt.c:379-380)
      const int off = plan.pulse_base - plan.pre_rotate - H;
      for(int k = tid; k < plan.pulse_size; k += nth) {
        const int o = off + k;
        const size_t at = (size_t)s * cap + rt_wrap(P.cur_new + o, cap);
        if(o < 0) Q.pulse_b[at] += T[k].x; else Q.pulse_f[at] += T[k].x;
      }
      __syncthreads();
      // the pulse scratch overlaps the harmonic coefficient tables: rebuild them
      for(int e = tid; e < nch * P.maxnhar_e; e += nth) {
        size_t ec = fr * nch + e / P.maxnhar_e;
        float a = P.eampl[ec * P.maxnhar_e + e % P.maxnhar_e], ph = P.ephse[ec * P.maxnhar_e + e % P.maxnhar_e];
        float sp, cp; sincosf(ph, &sp, &cp);
        ea[e] = a * cp; eb[e] = a * sp;
      }
      {
        float t = P.cycle * 2.0f;
        const float phase_shift = (float)((double)t * LLSM_PI * (double)f0);
        int nhs = P.nhar[fr]; if(nhs > P.maxnhar) nhs = P.maxnhar;
        for(int k = tid; k < nhs; k += nth) {
          float a = P.ampl[fr * P.maxnhar + k];
          float ph = (float)((double)P.phse[fr * P.maxnhar + k] - (double)phase_shift * ((double)k + 1.0));
          float sp, cp; sincosf(ph, &sp, &cp);
          ca[k] = a * cp; cb[k] = a * sp;
        }
      }
      __syncthreads();
    }
  }
  // ---- sinusoids (llsmrt.c:273-291): addchunk(sin, -nwin, nwin)
  int nh = P.nhar[fr];
  if(nh > nfft) nh = nfft;
  if(do_sin && f0 > 0 && nh > 0) {
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
  if(L1) {
    if(plan.read_main) {                               // llsmrt.c:395-403
      for(int i = tid; i < nwin; i += nth) {
        const int o = plan.main_offset + i;
        const size_t at = (size_t)s * cap + rt_wrap(P.cur_new + o, cap);
        float v = (o < 0 ? Q.pulse_b[at] : Q.pulse_f[at]) * P.win[i];
        sinr[rt_wrap(P.cur_new + o, cap)] += v;
      }
      __syncthreads();
    }
    if(plan.termination) {                             // llsmrt.c:404-419
      const int size = -H - plan.term_offset;
      for(int i = tid; i < size; i += nth) {
        const int o = plan.term_offset + i;
        const size_t at = (size_t)s * cap + rt_wrap(P.cur_new + o, cap);
        float v = o < 0 ? Q.pulse_b[at] : Q.pulse_f[at];
        if(i < H) v *= P.win[i];
        if(i >= size - H) v *= P.win[i - (size - H) + H];
        sinr[rt_wrap(P.cur_new + o, cap)] += v;
      }
      __syncthreads();
    }
  }

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
