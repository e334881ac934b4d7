// Pulse-by-pulse (PbP) synthesis from layer-1 parameters: llsm_synthesize_harmonics with
// use_l1 (layer0.c:148-287) and llsm_make_filtered_pulse (llsmutils.c:60-201).
//
//   pbp_prep_kernel    per voiced L1 frame: LF model of the frame, phase of its first harmonic
//                      (layer0.c:184-189)
//   pbp_track_kernel   per utterance, sequential in time: pulse tracker locked on the first source
//                      harmonic, pulse plan, HM <-> PbP switch ramp, which frames still need the
//                      harmonic model (layer0.c:164-261). No user callbacks on this path; the
//                      drop-in llsm_synthesize runs the same tracker on the host when a frame
//                      carries an llsm_pbpeffect.
//   pbp_pulse_kernel   per frame with pulses: filtered glottal pulses (LF spectra, phase-delta
//                      envelope, lip radiation, vocal-tract gain, inverse FFT, fades), overlap-add
//   pbp_mix_kernel     y = y_hm (1 - m) + y_pbp m  (layer0.c:281-282)
#pragma once
#include "common.cuh"
#include "lf_model.cuh"
#include "kernels_layer1.cuh"

#define PBP_MAXP 16          // pulses per frame the plan can hold

struct PbpPulse { float T0, te, tp, ta, Ee, offset; };

struct PbpPlan {             // per (utterance, frame)
  int* npulse;               // 0 = no pulses this frame
  int* pulse_base;           // (int) offsets[0]
  int* pre_rotate;           // (int) len_period
  float* len_period;         // re-quantised period (OLA index)
  int* pulse_size;
  PbpPulse* pulses;          // [B][F][PBP_MAXP]
  int* need_hm;              // harmonic-model frame is synthesised
};

// ---- per-frame LF phase ------------------------------------------------------------------------
struct PbpPrepParams {
  int nfrm; const int* nfrm_utt;
  const float* f0; const float* rd; const int* nvs;
  float* source_p0;          // [B][F] arg LF(f0) - pi / 2
};

__global__ void __launch_bounds__(128) pbp_prep_kernel(PbpPrepParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t r = (size_t)b * P.nfrm + i;
  const float f0 = P.f0[r];
  if(f0 == 0 || P.nvs[r] <= 0) { P.source_p0[r] = 0.f; return; }
  const float t_period = (float)(1.0 / (double)f0);
  LfSolved s = lf_solve(lf_from_rd(P.rd[r], t_period, 1.0f));
  double m, ph; lf_spectrum(s, (double)f0, &m, &ph);
  float p0 = (float)ph;
  p0 = (float)((double)p0 - 0.5 * LLSM_PI);              // layer0.c:189
  P.source_p0[r] = p0;
}

__device__ __forceinline__ float wrapf(float p) {        // (-pi, pi]
  double q = (double)p - 2.0 * LLSM_PI * floor(((double)p + LLSM_PI) / (2.0 * LLSM_PI));
  if(q <= -LLSM_PI) q += 2.0 * LLSM_PI;
  return (float)q;
}

// ---- sequential tracker --------------------------------------------------------------------------
struct PbpTrackParams {
  int nutt, nfrm; const int* nfrm_utt; const int* ny_utt; int ny, stride;
  const float* f0; const float* rd; const float* vsphse; const int* nvs; int vs_stride;
  const int* pbpsyn;         // [B][F] 1 = PbP requested (LLSM_FRAME_PBPSYN), may be NULL (all 0)
  const float* source_p0;
  const int* base_trunc;     // [F] (int)(i * thop * fs)
  float hop;                 // thop * fs (float product)
  float fs; int nspec;
  PbpPlan plan;
  float* y_mix;              // [B][stride], zero-initialised by the caller
};

// Tracker of one utterance. `mod(model, delta_t, frame)` is the per-pulse hook of llsm_pbpeffect
// (layer0.c:208-217): identity on the device, the user's callback in the host build of the drop-in API.
struct PbpNoEffect { LF_HD void operator()(LfModel&, float&, int) const {} };

template <class Mod>
LF_HD void pbp_track_utterance(const PbpTrackParams& P, int b, Mod& mod) {
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int ny = P.ny_utt ? P.ny_utt[b] : P.ny;
  float pulse_previous = 0.f;
  int pbp_periods = 0;
  const int thrd = 3;
  float switch_state = 0.f;
  int baseidx_prev = 0;
  float* ymix = P.y_mix + (size_t)b * P.stride;
  for(int i = 0; i < nf; i ++) {
    const size_t r = (size_t)b * P.nfrm + i;
    P.plan.npulse[r] = 0; P.plan.need_hm[r] = 0;
    const float f0 = P.f0[r];
    if(f0 == 0) continue;                               // layer0.c:172
    if(P.nvs[r] <= 0) continue;                         // no layer-1 members, layer0.c:180
    const int baseidx = P.base_trunc[i];
    const int pbp_on = P.pbpsyn ? (P.pbpsyn[r] == 1) : 0;
    float len_period = P.fs / f0;
    const float source_p0 = P.source_p0[r];
    float p0; {
      double pv = (double)P.vsphse[r * P.vs_stride];
      double q = pv - 2.0 * LLSM_PI * floor((pv + LLSM_PI) / (2.0 * LLSM_PI));
      if(q <= -LLSM_PI) q += 2.0 * LLSM_PI;
      p0 = (float)q;
    }
    float p0_dist; {
      double pv = (double)(p0 - source_p0);             // phase_diff(source_p0, p0) = wrap(p0 - source_p0)
      double q = pv - 2.0 * LLSM_PI * floor((pv + LLSM_PI) / (2.0 * LLSM_PI));
      if(q <= -LLSM_PI) q += 2.0 * LLSM_PI;
      p0_dist = (float)q;
    }
    if(p0_dist < 0) p0_dist = (float)((double)p0_dist + 2.0 * LLSM_PI);
    const float pulse_projected = (float)((double)baseidx + (double)(p0_dist / 2.0f) / LLSM_PI * (double)len_period);
    const int len_reset = (int)((len_period > P.hop ? len_period : P.hop) * 2.0f);
    if(pulse_projected - pulse_previous > (float)len_reset) pulse_previous = pulse_projected - (float)len_reset;
    const int num_periods = (int)round((double)((pulse_projected - pulse_previous) / len_period));
    len_period = (pulse_projected - pulse_previous) / (float)num_periods;      // unguarded, layer0.c:200
    if((pbp_on || pbp_periods > 0) && num_periods > 0) {
      float lp2 = len_period * 2.0f;
      float mxs = lp2 > (float)P.nspec ? lp2 : (float)P.nspec;
      const int pulse_size = pow2_ceil(log2((double)mxs));
      const float t_period = (float)(1.0 / (double)f0);
      const LfModel sm = lf_from_rd(P.rd[r], t_period, 1.0f);
      const int np = num_periods < PBP_MAXP ? num_periods : PBP_MAXP;
      PbpPulse* pl = P.plan.pulses + r * PBP_MAXP;
      int pulse_base = 0;
      for(int j = 0; j < num_periods; j ++) {
        LfModel m = sm; float delta_t = 0.f;
        mod(m, delta_t, i);                               // called once per pulse, in time order
        if(j >= np) continue;
        float off = pulse_previous + (float)j * len_period;
        off = off + delta_t * P.fs;                       // layer0.c:216
        pl[j].T0 = m.T0; pl[j].te = m.te; pl[j].tp = m.tp; pl[j].ta = m.ta; pl[j].Ee = m.Ee;
        pl[j].offset = off;
      }
      pulse_base = (int)pl[0].offset;                     // layer0.c:218-219
      for(int j = 0; j < np; j ++) pl[j].offset = pl[j].offset - (float)pulse_base;
      P.plan.npulse[r] = num_periods <= PBP_MAXP ? num_periods : -num_periods;   // < 0: overflow
      P.plan.pulse_base[r] = pulse_base;
      P.plan.pre_rotate[r] = (int)len_period;
      P.plan.len_period[r] = len_period;
      P.plan.pulse_size[r] = pulse_size;
      pbp_periods += pbp_on ? num_periods : -num_periods;
      if(pbp_periods > thrd) pbp_periods = thrd;
      if(pbp_periods < 0) pbp_periods = 0;
    }
    pulse_previous = pulse_projected;

    const float rate = (float)(1.0 / (double)(len_period < P.hop ? len_period : P.hop));
    int require_hm = 0;
    int jend = baseidx < ny ? baseidx : ny;             // y_mix has ny samples
    if(pbp_on && pbp_periods == thrd) {
      for(int j = baseidx_prev; j < baseidx; j ++) {
        if((double)switch_state < 1.0) { switch_state = switch_state + rate; require_hm = 1; }
        if(j < jend && j >= 0) ymix[j] = switch_state;
      }
    } else if(! pbp_on && pbp_periods == 0) {
      for(int j = baseidx_prev; j < baseidx; j ++) {
        if(switch_state > 0) { switch_state = switch_state - rate; require_hm = 1; }
        if(j < jend && j >= 0) ymix[j] = switch_state;
      }
    } else {
      for(int j = baseidx_prev; j < baseidx; j ++) if(j < jend && j >= 0) ymix[j] = switch_state;
    }
    baseidx_prev = baseidx;
    if(pbp_on && pbp_periods == thrd && ! require_hm) continue;
    P.plan.need_hm[r] = 1;
  }
}

__global__ void __launch_bounds__(32) pbp_track_kernel(PbpTrackParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= P.nutt) return;
  PbpNoEffect none;
  pbp_track_utterance(P, b, none);
}

// ---- pulses ------------------------------------------------------------------------------------------
struct PbpPulseParams {
  int nfrm; const int* nfrm_utt; const int* ny_utt; int ny, stride;
  const float* f0; const float* rd; const float* vtmagn; int nspec;
  const float* vsphse; const int* nvs; int vs_stride;
  float fs, fnyq, lip_radius;
  PbpPlan plan;
  const float2* tw; int ntw; int max_size; int max_mp; int maxnhar;
  float* y_pbp;              // [B][stride], zero-initialised; accumulated with atomics
};

#define PBP_THREADS 256

// linear interpolation on uniform-in-index float knots x_k = knot(k), clamped (oracle interp1)
__device__ __forceinline__ float interp_linspace(const float* y, int n, float xmax, float v) {
  if(! (v > 0.f)) return y[0];
  if(v >= xmax) return y[n - 1];
  int lo = 0, hi = n - 1;
  while(hi - lo > 1) {
    int mid = (lo + hi) / 2;
    float xm = (float)(((double)xmax) * mid / (n - 1));
    if(xm <= v) lo = mid; else hi = mid;
  }
  float xl = (float)(((double)xmax) * lo / (n - 1)), xh = (float)(((double)xmax) * hi / (n - 1));
  double rr = ((double)v - xl) / ((double)xh - xl);
  return (float)((double)y[lo] + ((double)y[hi] - (double)y[lo]) * rr);
}

// One frame's filtered pulse train (llsm_make_filtered_pulse, llsmutils.c:132-201) computed by the whole CTA:
// returns the buffer whose .x holds the `S` finished samples (inverse FFT scaled, fades applied).
// smem: bufa/bufb [S] float2, ha [maxnhar + 2], vta/vtp [maxnhar], pre/pim [maxnhar + 1]; solved [PBP_MAXP].
struct PulseFrame {
  float f0, rd; const float* vtmagn; int nspec; const float* vsphse; int nhar;
  float fs, fnyq, lip_radius; const float2* tw; int ntw;
};
struct PulseSmem { float2* bufa; float2* bufb; float* ha; float* vta; float* vtp; float* pre; float* pim; LfSolved* solved; };

__device__ float2* pbp_make_pulse(const PulseFrame& F, const PbpPulse* pulses, int npulse, int pre_rotate, int S,
  const PulseSmem& M) {
  const int tid = threadIdx.x, nth = blockDim.x;
  float2* bufa = M.bufa; float2* bufb = M.bufb;
  float* ha = M.ha; float* vta = M.vta; float* vtp = M.vtp; float* pre = M.pre; float* pim = M.pim;
  int lgS = 0; while((1 << lgS) < S) lgS ++;
  const int halfsize = S / 2 + 1;
  const float f0 = F.f0;
  const int nhar = F.nhar;
  const float* vtmagn = F.vtmagn;
  const float* vsphse = F.vsphse;

  // ---- vocal-tract amplitudes at the harmonics and their minimum phase (llsmutils.c:152-160)
  const LfSolved lf0 = lf_solve(lf_from_rd(F.rd, (float)(1.0 / (double)f0), 1.0f));
  for(int k = tid; k < nhar; k += nth) {
    float fh = (float)(k + 1) * f0;                                  // freq_har[i] = i * f0[0]
    float v = interp_linspace(vtmagn, F.nspec, F.fnyq, fh);
    vta[k] = (float)exp((double)v * 2.3025851 / 20.0);
  }
  __syncthreads();
  block_harmonic_minphase(vta, nhar, vtp, bufa, bufb, ha, F.tw, F.ntw);
  // ---- harmonic phase deltas (llsmutils.c:75-96)
  for(int k = tid; k <= nhar; k += nth) {
    if(k >= 1) {
      double m, p; lf_spectrum(lf0, (double)((float)k * f0), &m, &p);
      ha[k] = (float)p;                                              // raw LF phase of harmonic k
    }
  }
  __syncthreads();
  const float vsshift = (float)((double)vsphse[0] - ((double)ha[1] - 0.5 * LLSM_PI));
  for(int k = tid; k <= nhar; k += nth) {
    float ph = 0.f;
    if(k >= 1) {
      float lp = (float)((double)ha[k] - 0.5 * LLSM_PI);
      float d = vsphse[k - 1] - lp;
      d = d - vsshift * (float)k;
      ph = wrapf(d) + vtp[k - 1];
    }
    pre[k] = (float)cos((double)ph); pim[k] = (float)sin((double)ph);
  }
  double m0, p0u; lf_spectrum(lf0, (double)f0, &m0, &p0u);
  const float lfmagnf0 = (float)m0;
  if(tid < npulse) {
    const PbpPulse pu = pulses[tid];
    LfModel mm; mm.T0 = pu.T0; mm.te = pu.te; mm.tp = pu.tp; mm.ta = pu.ta; mm.Ee = pu.Ee;
    M.solved[tid] = lf_solve(mm);
  }
  __syncthreads();

  // ---- full-size spectrum: sum over pulses (llsmutils.c:108-125), lip, VT gain
  const float fhmax = (float)nhar * f0;
  for(int q = tid; q < halfsize; q += nth) {
    float re = 0.f, im = 0.f;
    if(q >= 1) {
      float fq = (float)q * F.fs; fq = fq / (float)S;                // freq_axis[i] = i * fs / size
      // interp1(freq_har, phse_{re,im}, nhar + 1, freq_axis): knots k * f0, clamped
      float dre, dim;
      if(! (fq > 0.f)) { dre = pre[0]; dim = pim[0]; }
      else if(fq >= fhmax) { dre = pre[nhar]; dim = pim[nhar]; }
      else {
        int lo = 0, hi = nhar;
        while(hi - lo > 1) { int mid = (lo + hi) / 2; if((float)mid * f0 <= fq) lo = mid; else hi = mid; }
        double rr = ((double)fq - (double)((float)lo * f0)) / ((double)((float)hi * f0) - (double)((float)lo * f0));
        dre = (float)((double)pre[lo] + ((double)pre[hi] - (double)pre[lo]) * rr);
        dim = (float)((double)pim[lo] + ((double)pim[hi] - (double)pim[lo]) * rr);
      }
      const float pdelta = (float)atan2((double)dim, (double)dre);
      for(int p = 0; p < npulse; p ++) {
        const float poff = pulses[p].offset;
        const LfSolved ls = M.solved[p];
        double mg, pg; lf_spectrum(ls, (double)fq, &mg, &pg);
        float lm = (float)mg, lph = (float)pg;
        float sc = F.fnyq / fq; sc = sc / lfmagnf0;
        lm = lm * sc;
        const float phase_shift = -poff - (float)pre_rotate;
        float t = phase_shift * (float)q; t = t * 2.0f;
        lph = (float)((double)lph + (double)t * LLSM_PI / (double)S);
        lph = (float)((double)lph + ((double)pdelta - 0.5 * LLSM_PI));
        re = (float)((double)re + (double)lm * cos((double)lph));
        im = (float)((double)im + (double)lm * sin((double)lph));
      }
    }
    // lip radiation: bin q is filtered at frequency (q + 1) * fs / size (dsputils.c:418-419)
    float fbase = F.fs / (float)S;
    float omega = (float)((double)fbase * (1.0 + q) * 2.0 * LLSM_PI);
    float2 ir = lip_response(F.lip_radius, omega);
    float yr = __fadd_rn(__fmul_rn(re, ir.x), -__fmul_rn(im, ir.y));
    float yi = __fadd_rn(__fmul_rn(re, ir.y), __fmul_rn(im, ir.x));
    // vocal-tract magnitude
    float fq0 = (float)q * F.fs; fq0 = fq0 / (float)S;
    float g = (float)exp((double)interp_linspace(vtmagn, F.nspec, F.fnyq, fq0) * 2.3025851 / 20.0);
    yr *= g; yi *= g;
    bufa[q] = make_float2(yr, yi);
    if(q > 0 && q < S / 2) bufa[S - q] = make_float2(yr, -yi);
  }
  __syncthreads();
  float2* T = block_fft<true>(bufa, bufb, lgS, F.tw, F.ntw);
  // ---- scale and fades (llsmutils.c:186-197)
  const int fadein = pre_rotate < 256 ? pre_rotate : 256;
  const int fadeout = S < 256 ? S : 256;
  const float invS = 1.0f / (float)S;
  for(int k = tid; k < S; k += nth) {
    float v = T[k].x * invS;
    if(k < fadein) v *= (float)k / (float)fadein;
    if(k >= S - fadeout) v *= (float)(S - k) / (float)fadeout;
    T[k].x = v;
  }
  __syncthreads();
  return T;
}

__global__ void __launch_bounds__(PBP_THREADS) pbp_pulse_kernel(PbpPulseParams P) {
  LLSM_DYN_SMEM(smem);
  __shared__ LfSolved solved[PBP_MAXP];              // LF model of every pulse of the frame
  PulseSmem M;
  M.bufa = (float2*)smem;                            // [max_size]
  M.bufb = M.bufa + P.max_size;                      // [max_size]
  M.ha = (float*)(M.bufb + P.max_size);              // [maxnhar + 2]
  M.vta = M.ha + P.maxnhar + 2;                      // [maxnhar]     harmonic VT amplitudes
  M.vtp = M.vta + P.maxnhar;                         // [maxnhar]     harmonic VT min-phase
  M.pre = M.vtp + P.maxnhar;                         // [maxnhar + 1] cos(phase delta)
  M.pim = M.pre + P.maxnhar + 1;                     // [maxnhar + 1] sin(phase delta)
  M.solved = solved;
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(i >= nf) return;
  const size_t r = (size_t)b * P.nfrm + i;
  const int npulse = P.plan.npulse[r];
  if(npulse <= 0) return;
  const int ny = P.ny_utt ? P.ny_utt[b] : P.ny;
  const int S = P.plan.pulse_size[r];
  if(S > P.max_size) return;
  PulseFrame F;
  F.f0 = P.f0[r]; F.rd = P.rd[r]; F.vtmagn = P.vtmagn + r * (size_t)P.nspec; F.nspec = P.nspec;
  F.vsphse = P.vsphse + r * (size_t)P.vs_stride;
  F.nhar = P.nvs[r] < P.maxnhar ? P.nvs[r] : P.maxnhar;
  F.fs = P.fs; F.fnyq = P.fnyq; F.lip_radius = P.lip_radius; F.tw = P.tw; F.ntw = P.ntw;
  const int pre_rotate = P.plan.pre_rotate[r];
  float2* T = pbp_make_pulse(F, P.plan.pulses + r * PBP_MAXP, npulse, pre_rotate, S, M);
  // ---- overlap-add (layer0.c:222-225)
  const float lenp = P.plan.len_period[r];
  const int pbase = P.plan.pulse_base[r];
  float* y = P.y_pbp + (size_t)b * P.stride;
  for(int k = tid; k < S; k += nth) {
    int idx = (int)__fadd_rn((float)(pbase + k), -lenp);             // int idx = pulse_base + k - len_period
    if(idx >= 0 && idx < ny) atomicAdd(y + idx, T[k].x);
  }
}

struct PbpMixParams { int ny, stride; const int* ny_utt; const float* y_hm; const float* y_pbp; float* y_mix_inout; };

__global__ void pbp_mix_kernel(PbpMixParams P) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  int ny = P.ny_utt ? P.ny_utt[b] : P.ny;
  if(n >= P.stride) return;
  size_t o = (size_t)b * P.stride + n;
  float m = P.y_mix_inout[o];
  float v = 0.f;
  if(n < ny) v = (float)((double)P.y_hm[o] * (1.0 - (double)m) + (double)(P.y_pbp[o] * m));   // layer0.c:282
  P.y_mix_inout[o] = v;
}
