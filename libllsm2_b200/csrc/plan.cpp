// Host-side plan construction (see plan.h). Compile WITHOUT fast-math / FMA contraction: the
// expressions below reproduce the reference's float arithmetic step by step.
#include "plan.h"
#include "cheby_table.h"
#include <cmath>
#include <cstring>
#include <algorithm>

int plan_output_length(int nfrm, float thop, float fs) {
  float v = (float)(nfrm + 1) * thop;   // layer0.c:643  round((nfrm + 1) * thop * fs)
  v = v * fs;
  return (int)round((double)v);
}

int plan_template_length(int ny) {     // dsputils.c:386-388
  return std::min(20000, ny) + 128;
}

void make_hanning(std::vector<float>& w, int n) {
  w.resize(n > 0 ? n : 0);
  for(int i = 0; i < n; i ++) w[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / n));
}

void make_blackman(std::vector<float>& w, int n) {
  w.resize(n > 0 ? n : 0);
  for(int i = 0; i < n; i ++)
    w[i] = (float)(0.42 - 0.5 * cos(2.0 * M_PI * i / n) + 0.08 * cos(4.0 * M_PI * i / n));
}

void build_twiddle(std::vector<float>& tw, int n) {
  tw.resize(2 * (size_t)n);
  for(int m = 0; m < n; m ++) {
    tw[2 * m] = (float)cos(2.0 * M_PI * m / n);
    tw[2 * m + 1] = (float)(-sin(2.0 * M_PI * m / n));
  }
}

static void pick_cheby(float cutoff, bool lowpass, double* b, double* a) {
  // dsputils.c:31-32: index = max(0, round(cutoff * 2.0 / step_freq - 1)), clamped to the table
  const float step_freq = (float)LLSM_CHEBY_STEP;
  double v = round((double)cutoff * 2.0 / (double)step_freq - 1);
  int index = (int)(v > 0 ? v : 0);
  if(index >= LLSM_CHEBY_NFILT) index = LLSM_CHEBY_NFILT - 1;
  for(int i = 0; i < LLSM_CHEBY_NCOEF; i ++) {
    // the model stores the coefficients as FP_TYPE (float)
    a[i] = (double)(float)(lowpass ? llsm_cheby_l_a[index][i] : llsm_cheby_h_a[index][i]);
    b[i] = (double)(float)(lowpass ? llsm_cheby_l_b[index][i] : llsm_cheby_h_b[index][i]);
  }
}

int select_chebyfilt(float c1, float c2, double b[2][5], double a[2][5]) {
  // dsputils.c:51-70: band-pass = high-pass(c1) followed by low-pass(c2)
  if(! (c1 > 0.0f)) c1 = 0.0f;
  if(c2 > 0.5f) c2 = 0.5f;
  if(c1 != 0 && c2 < 0.5f) {
    pick_cheby(c1, false, b[0], a[0]);
    pick_cheby(c2, true, b[1], a[1]);
    return 2;
  }
  if(c1 == 0) pick_cheby(c2, true, b[0], a[0]);
  else        pick_cheby(c1, false, b[0], a[0]);
  return 1;
}

void build_synth_plan(SynthPlan& p, int nfrm, float fs, float thop, int npsd, int nchannel,
  const float* chanfreq) {
  p.nfrm = nfrm; p.fs = fs; p.thop = thop; p.npsd = npsd; p.nchannel = nchannel;
  p.ny = plan_output_length(nfrm, thop, fs);

  // ---- harmonic OLA, layer0.c:121-129
  {
    float hop = thop * fs;
    p.n_hm = (int)(round((double)hop) * 2);
    p.hm_base.resize(nfrm); p.hm_frac.resize(nfrm); p.base_trunc.resize(nfrm);
    p.hop_f = hop;
    for(int i = 0; i < nfrm; i ++) {
      float rawidx = (float)i * thop;
      rawidx = rawidx * fs;
      p.base_trunc[i] = (int)rawidx;
      int baseidx = (int)round((double)rawidx);
      p.hm_base[i] = baseidx;
      p.hm_frac[i] = rawidx - (float)baseidx;
    }
    make_hanning(p.win_hm, p.n_hm);
  }

  // ---- noise envelope OLA, layer0.c:293,307
  {
    p.n_env = (int)round((double)thop * 2.0 * (double)fs);
    p.env_r.resize(nfrm); p.env_off.resize(nfrm);
    for(int i = 0; i < nfrm; i ++) {
      float r = (float)(i - 1) * thop;
      r = r * fs;
      p.env_r[i] = r;
      p.env_off[i] = (int)round((double)r);
    }
    // layer0.c:307 evaluates round((i - 1) * thop * fs + j) in float for every j: where the float sum
    // keeps the fraction of env_r[i] the indices are contiguous and the kernel takes a fast path
    p.env_contig.assign(nfrm, 1);
    for(int i = 0; i < nfrm; i ++)
      for(int j = 0; j < p.n_env; j ++) {
        float t = p.env_r[i] + (float)j;
        if((int)round((double)t) != p.env_off[i] + j) { p.env_contig[i] = 0; break; }
      }
    make_hanning(p.win_env, p.n_env);
  }

  // ---- noise shaping, layer0.c:559-603
  {
    float w2 = thop * fs;
    w2 = w2 * 2;
    p.n_ns = (int)round((double)w2);
    make_hanning(p.win_ns, p.n_ns);
    float wsqr = 0;
    for(int i = 0; i < p.n_ns; i ++) {
      float sq = p.win_ns[i] * p.win_ns[i];
      wsqr = wsqr + sq;
    }
    p.wsqr = wsqr;
    const int nfade = 16;
    p.nfft_ns = (int)pow(2.0, ceil(log2(p.n_ns * 1.2 + nfade * 2)));
    p.lg_nfft_ns = 0;
    while((1 << p.lg_nfft_ns) < p.nfft_ns) p.lg_nfft_ns ++;
    p.nspec_ns = p.nfft_ns / 2 + 1;

    // llsm_spectrum_from_envelope(src_axis, src_psd, npsd, nspec - 1, fs / 2.0): dsputils.c:308-316
    float fnyq = (float)((double)fs / 2.0);
    std::vector<float> xi(npsd);
    for(int i = 0; i < npsd; i ++)   // linspace(0, fnyq, npsd)
      xi[i] = npsd > 1 ? (float)(0.0 + ((double)fnyq - 0.0) * i / (npsd - 1)) : 0.0f;
    int nq = p.nspec_ns - 1;
    p.psd_lo.resize(nq); p.psd_r.resize(nq);
    for(int j = 0; j < nq; j ++) {
      float v = (float)j * fnyq;
      v = v / (float)nq;
      if(! (v > xi[0])) { p.psd_lo[j] = 0; p.psd_r[j] = 0; continue; }
      if(v >= xi[npsd - 1]) { p.psd_lo[j] = npsd - 1; p.psd_r[j] = 0; continue; }
      int lo = 0, hi = npsd - 1;
      while(hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if(xi[mid] <= v) lo = mid; else hi = mid;
      }
      p.psd_lo[j] = lo;
      p.psd_r[j] = (float)(((double)v - xi[lo]) / ((double)xi[hi] - xi[lo]));
    }
  }

  // ---- noise template + channel filters, layer0.c:540-544, dsputils.c:385-394
  p.ntemplate = std::min(20000, p.ny);
  p.nt = p.ntemplate + 128;
  p.chan.assign(nchannel, SynthPlan::ChanFilt());
  for(int c = 0; c < nchannel; c ++) {
    float fmin = c == 0 ? 0.0f : chanfreq[c - 1];
    float fmax = c == nchannel - 1 ? (float)((double)fs / 2.0) : chanfreq[c];
    SynthPlan::ChanFilt& cf = p.chan[c];
    memset(&cf, 0, sizeof(cf));
    if((double)fmin >= (double)fs / 2.0) {           // layer0.c:543 break
      for(int d = c; d < nchannel; d ++) p.chan[d].nstage = 0;
      break;
    }
    cf.nstage = select_chebyfilt(fmin / fs, fmax / fs, cf.b, cf.a);
  }
}

static void mat4_mul(const long double* x, const long double* y, long double* out) {
  long double t[16];
  for(int i = 0; i < 4; i ++) for(int j = 0; j < 4; j ++) {
    long double s = 0;
    for(int k = 0; k < 4; k ++) s += x[i * 4 + k] * y[k * 4 + j];
    t[i * 4 + j] = s;
  }
  for(int i = 0; i < 16; i ++) out[i] = t[i];
}

void build_iir_section(const double b[5], const double a[5], int L, int nlog, double* coef, double* mpow) {
  for(int i = 0; i < 5; i ++) coef[i] = b[i] / a[0];
  for(int i = 1; i < 5; i ++) coef[4 + i] = a[i] / a[0];
  // zero-input transition of DF2T: y = z0; z0' = z1 - a1 y; z1' = z2 - a2 y; z2' = z3 - a3 y; z3' = -a4 y
  long double A[16] = {0};
  A[0] = -coef[5]; A[1] = 1;
  A[4] = -coef[6]; A[6] = 1;
  A[8] = -coef[7]; A[11] = 1;
  A[12] = -coef[8];
  // M = A^L by binary exponentiation
  long double M[16], P[16];
  for(int i = 0; i < 16; i ++) { M[i] = (i % 5 == 0) ? 1 : 0; P[i] = A[i]; }
  for(int e = L; e > 0; e >>= 1) {
    if(e & 1) mat4_mul(M, P, M);
    mat4_mul(P, P, P);
  }
  for(int q = 0; q < nlog; q ++) {
    for(int i = 0; i < 16; i ++) mpow[q * 16 + i] = (double)M[i];
    mat4_mul(M, M, M);
  }
}
