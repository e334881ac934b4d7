/*
  ciglet shim -- TEST INFRASTRUCTURE (oracle), not product code.

  libllsm2 takes every DSP primitive from the third-party library "ciglet"
  (Sleepwalking/ciglet; `#include <ciglet/ciglet.h>` in layer0.c:20, dsputils.c:20,
  llsmrt.c:24, llsmutils.h:26, frame.c:20, layer1.c:20). ciglet is neither vendored nor
  version-pinned by the reference (README.md:43-51) and is absent from this machine, so this
  header + ciglet.c restate, from the published definitions of each primitive, exactly the
  symbols the reference's hot path needs. Where ciglet's behaviour is not forced by a
  reference call site the choice made here is documented next to the function and mirrored
  by the CUDA kernels. PARITY UNPINNED at the 1e-4 level: the reference ships no golden
  vectors for any of these functions; the reference's own tolerance-level
  known-answer tests (test/test-dsputils.c) are re-run against this shim in tests/.

  Build-time knobs:
    FP_TYPE   storage type, set by the includer (the reference uses float).
    CIG_WORK  internal working precision of transforms/recurrences (default double for the
              parity build; the timing build uses float, like a real ciglet with FP_TYPE=float).
*/
#ifndef CIGLET_SHIM_H
#define CIGLET_SHIM_H

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

#ifndef FP_TYPE
#error "FP_TYPE must be defined by the includer (libllsm2 README.md:53)"
#endif
#ifndef CIG_WORK
#define CIG_WORK double
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846 /* constants.h:9-11 falls back to a wrong value otherwise */
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- small macros (used with mixed int/float arguments, e.g. dsputils.c:31) ---- */
#ifndef max
#define max(a, b) ((a) > (b) ? (a) : (b))
#endif
#ifndef min
#define min(a, b) ((a) < (b) ? (a) : (b))
#endif
#define linterp(a, b, r) ((a) + ((b) - (a)) * (r))

/* accuracy-tiered math of upstream ciglet: the shim uses exact libm for every tier */
#define exp_1(x) exp(x)
#define exp_2(x) exp(x)
#define exp_3(x) exp(x)
#define log_1(x) log(x)
#define log_2(x) log(x)
#define log_3(x) log(x)
#define cos_1(x) cos(x)
#define cos_2(x) cos(x)
#define cos_3(x) cos(x)
#define sin_1(x) sin(x)
#define sin_2(x) sin(x)
#define sin_3(x) sin(x)

/* ---- complex helpers (dsputils.c:402-409,420-428) ---- */
typedef struct { FP_TYPE real; FP_TYPE imag; } cplx;
static inline cplx c_cplx(FP_TYPE re, FP_TYPE im) { cplx r; r.real = re; r.imag = im; return r; }
static inline cplx c_mul(cplx a, cplx b) {
  return c_cplx(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real);
}
static inline cplx c_div(cplx a, cplx b) {
  FP_TYPE d = b.real * b.real + b.imag * b.imag;
  return c_cplx((a.real * b.real + a.imag * b.imag) / d, (a.imag * b.real - a.real * b.imag) / d);
}
static inline FP_TYPE c_abs(cplx a) { return sqrt(a.real * a.real + a.imag * a.imag); }
static inline FP_TYPE c_arg(cplx a) { return atan2(a.imag, a.real); }

/* ---- phase helpers: wrap to (-pi, pi]; phase_diff(a,b) = wrap(b - a) (layer0.c:190-192) ---- */
static inline FP_TYPE wrap(FP_TYPE p) {
  double q = p - 2.0 * M_PI * floor((p + M_PI) / (2.0 * M_PI));
  if(q <= -M_PI) q += 2.0 * M_PI;
  return q;
}
static inline FP_TYPE phase_diff(FP_TYPE a, FP_TYPE b) { return wrap(b - a); }

/* ---- mel scale (coder.c:65-70). Decision: the HTK / O'Shaughnessy curve mel = 1125 ln(1 + f / 700). ---- */
static inline FP_TYPE freq2mel(FP_TYPE f) { return 1125.0 * log(1.0 + f / 700.0); }
static inline FP_TYPE mel2freq(FP_TYPE m) { return 700.0 * (exp(m / 1125.0) - 1.0); }

/* ---- memory / vector helpers ---- */
void** malloc2d(size_t n, size_t m, size_t size);
#define free2d(ptr, n) cig_free2d((void**)(ptr), (n))
void cig_free2d(void** ptr, size_t n);
FP_TYPE* linspace(FP_TYPE a, FP_TYPE b, int n);   /* inclusive end points */
FP_TYPE* cumsum(FP_TYPE* x, int n);
FP_TYPE sumfp(FP_TYPE* x, int n);
FP_TYPE meanfp(FP_TYPE* x, int n);
FP_TYPE maxfp(FP_TYPE* x, int n);
int cig_find_peak(FP_TYPE* x, int lidx, int uidx, int orient); /* arg-max (orient>0) over [l,u] */
int find_minima(FP_TYPE* x, int lidx, int uidx);                /* arg-min over [l,u] */
FP_TYPE qifft(FP_TYPE* magn, int k, FP_TYPE* dst_pos);          /* parabolic peak refinement */
FP_TYPE* fetch_frame(FP_TYPE* x, int nx, int center, int nf);
FP_TYPE* interp1(FP_TYPE* xi, FP_TYPE* yi, int ni, FP_TYPE* xq, int nq);
FP_TYPE* interp1u(FP_TYPE x0, FP_TYPE x1, FP_TYPE* yi, int ni, FP_TYPE* xq, int nq);
FP_TYPE* interp_in_blank(FP_TYPE* x, int n, FP_TYPE blank);
FP_TYPE* moving_avg(FP_TYPE* x, int n, FP_TYPE halford);
void complete_symm(FP_TYPE* x, int n);
void complete_asymm(FP_TYPE* x, int n);
FP_TYPE itakura_saito(FP_TYPE* S, FP_TYPE* S0, int n);
FP_TYPE safe_aliased_sinc(FP_TYPE M, FP_TYPE omega);
FP_TYPE randn(FP_TYPE mu, FP_TYPE var);

/* ---- windows (periodic form, centre at index n/2) ---- */
FP_TYPE* hanning(int n);
FP_TYPE* blackman(int n);
#define hanning_2(n) hanning(n)

/* ---- transforms ---- */
void fft(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, int n, FP_TYPE* buffer);
void ifft(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, int n, FP_TYPE* buffer);
void czt(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, FP_TYPE omega0, int n);
void iczt(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, FP_TYPE omega0, int n);
FP_TYPE* gensins(FP_TYPE* freq, FP_TYPE* ampl, FP_TYPE* phse, int nsin, FP_TYPE fs, int n);
void cig_stft_forward(FP_TYPE* x, int nx, int* center, int* nwin, int nfrm, int nfft,
  char* window, int subt_mean, int optlv, FP_TYPE* norm_factor, FP_TYPE* weight_factor,
  FP_TYPE** Xmagn, FP_TYPE** Xphse);
FP_TYPE* cig_spec2env(FP_TYPE* S, int nfft, FP_TYPE f0, int nhar, FP_TYPE* Cout);
#define spec2env(S, nfft, f0, Cout) cig_spec2env(S, nfft, f0, 0, Cout)
FP_TYPE* minphase(FP_TYPE* logmagn, int nfft);

/* ---- filters ---- */
FP_TYPE* filtfilt(FP_TYPE* b, int nb, FP_TYPE* a, int na, FP_TYPE* x, int nx);
FP_TYPE* kalmanf1d(FP_TYPE* z, FP_TYPE* Q, FP_TYPE* R, int n, FP_TYPE* P_out, FP_TYPE* L_out);
FP_TYPE* kalmans1d(FP_TYPE* y, FP_TYPE* P, FP_TYPE* Q, int n);

/* ---- Liljencrants-Fant glottal model; te/tp/ta are relative to T0 (llsmutils.c:25-43) ---- */
typedef struct { FP_TYPE T0; FP_TYPE te; FP_TYPE tp; FP_TYPE ta; FP_TYPE Ee; } lfmodel;
lfmodel lfmodel_from_rd(FP_TYPE rd, FP_TYPE T0, FP_TYPE Ee);
FP_TYPE* lfmodel_spectrum(lfmodel model, FP_TYPE* freq, int nf, FP_TYPE* dst_phase);

/* ---- instantaneous-frequency detector (dsputils.c:79-87) ---- */
typedef struct { FP_TYPE fc; int nh; FP_TYPE* hr; FP_TYPE* hi; FP_TYPE* hdr; FP_TYPE* hdi; } ifdetector;
ifdetector* create_ifdetector(FP_TYPE fc, FP_TYPE fres);
FP_TYPE ifdetector_estimate(ifdetector* ifd, FP_TYPE* x, int nx);
void delete_ifdetector(ifdetector* dst);

#ifdef __cplusplus
}
#endif
#endif
