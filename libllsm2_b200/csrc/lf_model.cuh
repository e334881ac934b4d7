// Liljencrants-Fant glottal flow-derivative model: Rd parameterisation (Fant 1995) and the
// closed-form spectrum of the two LF segments (Doval, d'Alessandro & Henrich 2006).
// Host + device (used by the layer-1 kernels and by host plans). The conventions are those of the
// CPU checker's ciglet restatement (lfmodel_from_rd / lfmodel_spectrum there), which the
// reference calls at layer1.c:100-101,172-173, llsmutils.c:75-76,114-115, layer0.c:186-188,
// dsputils.c:526-527: te / tp / ta relative to T0, Rap clamped to >= 1e-3, zero-net-flow alpha by
// bisection, everything in double, struct fields rounded to FP_TYPE (float).
#pragma once
#include "common.cuh"

#if defined(__CUDACC__) || defined(LLSM_EMU)
#define LF_HD __host__ __device__ __forceinline__
#else
#define LF_HD inline
#endif

struct LfModel { float T0, te, tp, ta, Ee; };
struct LfSolved { double te, tp, ta, wg, eps, alpha, E0, Ee, T0; };
struct LfCplx { double re, im; };

LF_HD LfCplx lf_c(double r, double i) { LfCplx c; c.re = r; c.im = i; return c; }
LF_HD LfCplx lf_add(LfCplx a, LfCplx b) { return lf_c(a.re + b.re, a.im + b.im); }
LF_HD LfCplx lf_sub(LfCplx a, LfCplx b) { return lf_c(a.re - b.re, a.im - b.im); }
LF_HD LfCplx lf_mul(LfCplx a, LfCplx b) { return lf_c(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
LF_HD LfCplx lf_scale(LfCplx a, double s) { return lf_c(a.re * s, a.im * s); }
LF_HD LfCplx lf_div(LfCplx a, LfCplx b) {
  double d = b.re * b.re + b.im * b.im;
  return lf_c((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
LF_HD LfCplx lf_exp(LfCplx a) { double e = exp(a.re); return lf_c(e * cos(a.im), e * sin(a.im)); }

LF_HD LfModel lf_from_rd(float rd, float T0, float Ee) {
  double Rd = rd;
  double Rap = (-1.0 + 4.8 * Rd) / 100.0;
  double Rkp = (22.4 + 11.8 * Rd) / 100.0;
  double Rgp = 1.0 / (4.0 * ((0.11 * Rd / (0.5 + 1.2 * Rkp)) - Rap) / Rkp);
  if(Rap < 1e-3) Rap = 1e-3;
  LfModel m;
  m.T0 = T0;
  m.tp = (float)(1.0 / (2.0 * Rgp));
  m.te = (float)((double)m.tp * (Rkp + 1.0));
  m.ta = (float)Rap;
  m.Ee = Ee;
  return m;
}

LF_HD double lf_netflow(const LfSolved& s, double alpha) {
  double ste = sin(s.wg * s.te), cte = cos(s.wg * s.te);
  double den = alpha * alpha + s.wg * s.wg;
  double A1 = s.Ee / (-ste) * ((alpha * ste - s.wg * cte) + s.wg * exp(-alpha * s.te)) / den;
  double d = 1.0 - s.te;
  double ex = exp(-s.eps * d);
  double A2 = -(s.Ee / (s.eps * s.ta)) * ((1.0 - ex) / s.eps - d * ex);
  return A1 + A2;
}

LF_HD LfSolved lf_solve(LfModel m) {
  LfSolved s;
  s.te = m.te; s.tp = m.tp; s.ta = m.ta; s.Ee = m.Ee; s.T0 = m.T0;
  if(s.ta < 1e-6) s.ta = 1e-6;
  if(s.te > 1.0 - 1e-6) s.te = 1.0 - 1e-6;
  s.wg = LLSM_PI / s.tp;
  double d = 1.0 - s.te;
  double eps = 1.0 / s.ta;
  for(int it = 0; it < 50; it ++) {
    double ex = exp(-eps * d);
    double f = eps * s.ta - 1.0 + ex;
    double fp = s.ta - d * ex;
    double step = f / fp;
    eps -= step;
    if(eps <= 0) eps = 1e-3;
    if(fabs(step) < 1e-13 * fabs(eps)) break;
  }
  s.eps = eps;
  double lo = -200.0, hi = 400.0;
  for(int it = 0; it < 100; it ++) {
    double mid = 0.5 * (lo + hi);
    if(lf_netflow(s, mid) > 0) lo = mid; else hi = mid;
  }
  s.alpha = 0.5 * (lo + hi);
  s.E0 = -s.Ee / (exp(s.alpha * s.te) * sin(s.wg * s.te));
  return s;
}

// Fourier transform of the flow derivative at `freq` Hz: magnitude and phase.
LF_HD void lf_spectrum(const LfSolved& s, double freq, double* magn, double* phase) {
  const double d = 1.0 - s.te;
  const double ste = sin(s.wg * s.te), cte = cos(s.wg * s.te);
  const double exd = exp(-s.eps * d);
  const double w = 2.0 * LLSM_PI * freq * s.T0;
  LfCplx sc = lf_c(s.alpha, -w);
  LfCplx num = lf_add(lf_mul(lf_exp(lf_scale(sc, s.te)), lf_sub(lf_scale(sc, ste), lf_c(s.wg * cte, 0))), lf_c(s.wg, 0));
  LfCplx den = lf_add(lf_mul(sc, sc), lf_c(s.wg * s.wg, 0));
  LfCplx P1 = lf_scale(lf_div(num, den), s.E0);
  LfCplx ew = lf_c(s.eps, w);
  LfCplx t1 = lf_div(lf_sub(lf_c(1, 0), lf_exp(lf_scale(ew, -d))), ew);
  LfCplx t2;
  if(fabs(w) > 1e-12) t2 = lf_scale(lf_div(lf_sub(lf_c(1, 0), lf_exp(lf_c(0, -w * d))), lf_c(0, w)), exd);
  else t2 = lf_c(exd * d, 0);
  LfCplx P2 = lf_scale(lf_mul(lf_exp(lf_c(0, -w * s.te)), lf_sub(t1, t2)), -(s.Ee / (s.eps * s.ta)));
  LfCplx X = lf_scale(lf_add(P1, P2), s.T0);
  *magn = sqrt(X.re * X.re + X.im * X.im);
  *phase = atan2(X.im, X.re);
}
