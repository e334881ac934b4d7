/*
  ciglet shim implementation -- TEST INFRASTRUCTURE (oracle), not product code.
  See ciglet/ciglet.h for scope. Every function below is a from-scratch restatement of the
  standard definition of the primitive; the "Decision:" notes record choices that the reference's
  call sites do not force and that the CUDA kernels mirror.
*/
#include "ciglet/ciglet.h"
#include <complex.h>

typedef CIG_WORK wk;

/* ------------------------------------------------------------------ memory / vectors */

void** malloc2d(size_t n, size_t m, size_t size) {
  void** ret = calloc(n > 0 ? n : 1, sizeof(void*));
  for(size_t i = 0; i < n; i ++) ret[i] = calloc(m > 0 ? m : 1, size);
  return ret;
}

void cig_free2d(void** ptr, size_t n) {
  if(ptr == NULL) return;
  for(size_t i = 0; i < n; i ++) free(ptr[i]);
  free(ptr);
}

FP_TYPE* linspace(FP_TYPE a, FP_TYPE b, int n) {
  FP_TYPE* y = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  for(int i = 0; i < n; i ++)
    y[i] = n > 1 ? (double)a + ((double)b - (double)a) * i / (n - 1) : a;
  return y;
}

FP_TYPE* cumsum(FP_TYPE* x, int n) {
  FP_TYPE* y = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  double acc = 0;
  for(int i = 0; i < n; i ++) { acc += x[i]; y[i] = acc; }
  return y;
}

FP_TYPE sumfp(FP_TYPE* x, int n) {
  double acc = 0;
  for(int i = 0; i < n; i ++) acc += x[i];
  return acc;
}

FP_TYPE meanfp(FP_TYPE* x, int n) {
  return n > 0 ? sumfp(x, n) / n : 0;
}

FP_TYPE maxfp(FP_TYPE* x, int n) {
  FP_TYPE m = x[0];
  for(int i = 1; i < n; i ++) if(x[i] > m) m = x[i];
  return m;
}

/* first arg-max / arg-min over the inclusive range (dsputils.c:136,566) */
int cig_find_peak(FP_TYPE* x, int lidx, int uidx, int orient) {
  int best = lidx;
  for(int i = lidx + 1; i <= uidx; i ++)
    if(orient > 0 ? x[i] > x[best] : x[i] < x[best]) best = i;
  return best;
}

int find_minima(FP_TYPE* x, int lidx, int uidx) {
  return cig_find_peak(x, lidx, uidx, -1);
}

/* Parabola through (k-1,k,k+1); returns the vertex value, *dst_pos = fractional index.
   Decision: offsets outside (-1,1) or a degenerate parabola fall back to the bin itself. */
FP_TYPE qifft(FP_TYPE* magn, int k, FP_TYPE* dst_pos) {
  double a = magn[k - 1], b = magn[k], c = magn[k + 1];
  double a1 = (a + c) * 0.5 - b;
  double a2 = (c - a) * 0.5;
  double x = a1 != 0 ? -a2 / (2.0 * a1) : 0;
  if(! (fabs(x) < 1.0)) x = 0;
  *dst_pos = k + x;
  return a1 * x * x + a2 * x + b;
}

/* y[j] = x[center + j - nf/2], zero outside [0,nx) (alignment forced by layer0.c:589-623) */
FP_TYPE* fetch_frame(FP_TYPE* x, int nx, int center, int nf) {
  FP_TYPE* y = calloc(nf > 0 ? nf : 1, sizeof(FP_TYPE));
  for(int j = 0; j < nf; j ++) {
    int idx = center + j - nf / 2;
    if(idx >= 0 && idx < nx) y[j] = x[idx];
  }
  return y;
}

/* Linear interpolation on sorted knots, clamped outside the knot range. */
FP_TYPE* interp1(FP_TYPE* xi, FP_TYPE* yi, int ni, FP_TYPE* xq, int nq) {
  FP_TYPE* y = calloc(nq > 0 ? nq : 1, sizeof(FP_TYPE));
  for(int q = 0; q < nq; q ++) {
    FP_TYPE v = xq[q];
    if(! (v > xi[0])) { y[q] = yi[0]; continue; }
    if(v >= xi[ni - 1]) { y[q] = yi[ni - 1]; continue; }
    int lo = 0, hi = ni - 1;              /* xi[lo] < v < xi[hi] */
    while(hi - lo > 1) {
      int mid = (lo + hi) / 2;
      if(xi[mid] <= v) lo = mid; else hi = mid;
    }
    double r = ((double)v - xi[lo]) / ((double)xi[hi] - xi[lo]);
    y[q] = yi[lo] + ((double)yi[hi] - yi[lo]) * r;
  }
  return y;
}

/* Uniform-grid interpolation. Decision: EXCLUSIVE end point, i.e. knot i sits
   at x0 + i * (x1 - x0) / ni -- both dsputils.c:495-499 call sites pass "last knot + one step"
   as x1. Clamped outside [knot 0, knot ni-1]. */
FP_TYPE* interp1u(FP_TYPE x0, FP_TYPE x1, FP_TYPE* yi, int ni, FP_TYPE* xq, int nq) {
  FP_TYPE* y = calloc(nq > 0 ? nq : 1, sizeof(FP_TYPE));
  double step = ((double)x1 - (double)x0) / ni;
  for(int q = 0; q < nq; q ++) {
    double p = ((double)xq[q] - (double)x0) / step;
    if(! (p > 0)) { y[q] = yi[0]; continue; }
    if(p >= ni - 1) { y[q] = yi[ni - 1]; continue; }
    int k = (int)p;
    double r = p - k;
    y[q] = yi[k] + ((double)yi[k + 1] - yi[k]) * r;
  }
  return y;
}

/* Fill runs equal to `blank` by linear interpolation between the neighbouring valid samples;
   hold the nearest valid sample at the ends (layer1.c:74). */
FP_TYPE* interp_in_blank(FP_TYPE* x, int n, FP_TYPE blank) {
  FP_TYPE* y = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  memcpy(y, x, n * sizeof(FP_TYPE));
  int prev = -1;
  for(int i = 0; i < n; i ++) {
    if(x[i] == blank) continue;
    if(prev < 0) {
      for(int j = 0; j < i; j ++) y[j] = x[i];
    } else if(i - prev > 1) {
      for(int j = prev + 1; j < i; j ++)
        y[j] = x[prev] + ((double)x[i] - x[prev]) * (j - prev) / (i - prev);
    }
    prev = i;
  }
  if(prev >= 0) for(int j = prev + 1; j < n; j ++) y[j] = x[prev];
  return y;
}

/* Centred moving average. Decision: the third argument is the HALF order (h = 3 -> 7 taps),
   the window shrinks at the edges (layer0.c:597, llsmrt.c:455). */
FP_TYPE* moving_avg(FP_TYPE* x, int n, FP_TYPE halford) {
  FP_TYPE* y = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  int h = (int)halford;
  for(int i = 0; i < n; i ++) {
    int lo = i - h < 0 ? 0 : i - h;
    int hi = i + h > n - 1 ? n - 1 : i + h;
    double acc = 0;
    for(int j = lo; j <= hi; j ++) acc += x[j];
    y[i] = acc / (hi - lo + 1);
  }
  return y;
}

/* Hermitian completion of a half spectrum stored in an n-point buffer */
void complete_symm(FP_TYPE* x, int n) {
  for(int i = 1; i < n / 2; i ++) x[n - i] = x[i];
}

void complete_asymm(FP_TYPE* x, int n) {
  for(int i = 1; i < n / 2; i ++) x[n - i] = -x[i];
}

/* mean Itakura-Saito divergence (dsputils.c:561) */
FP_TYPE itakura_saito(FP_TYPE* S, FP_TYPE* S0, int n) {
  double acc = 0;
  for(int i = 0; i < n; i ++) {
    double r = (double)S[i] / (double)S0[i];
    acc += r - log(r) - 1.0;
  }
  return n > 0 ? acc / n : 0;
}

/* Dirichlet kernel sin(M w / 2) / sin(w / 2), -> M at w -> 0 (dsputils.c:444-446) */
FP_TYPE safe_aliased_sinc(FP_TYPE M, FP_TYPE omega) {
  double d = sin(0.5 * (double)omega);
  if(fabs(d) < 1e-9) return M;
  return sin(0.5 * (double)M * (double)omega) / d;
}

/* Gaussian deviate from libc rand(); the second argument is the VARIANCE
   (test/verify-utils.h:51). Decision: one Box-Muller cosine branch per call, two rand() draws:
     u1 = (rand() + 1) / (RAND_MAX + 2), u2 = (rand() + 1) / (RAND_MAX + 2),
     z  = sqrt(-2 ln u1) cos(2 pi u2).
   The product's host-side template generator repeats exactly this sequence. */
FP_TYPE randn(FP_TYPE mu, FP_TYPE var) {
  double u1 = ((double)rand() + 1.0) / ((double)RAND_MAX + 2.0);
  double u2 = ((double)rand() + 1.0) / ((double)RAND_MAX + 2.0);
  double z = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
  return (double)mu + sqrt((double)var) * z;
}

/* ------------------------------------------------------------------ windows */
/* Decision: periodic (DFT-even) forms so that the centre is sample n/2 and Hann windows at hop
   n/2 sum to one (needed by the OLA at layer0.c:121-140). */

FP_TYPE* hanning(int n) {
  FP_TYPE* w = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  for(int i = 0; i < n; i ++) w[i] = 0.5 - 0.5 * cos(2.0 * M_PI * i / n);
  return w;
}

FP_TYPE* blackman(int n) {
  FP_TYPE* w = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  for(int i = 0; i < n; i ++)
    w[i] = 0.42 - 0.5 * cos(2.0 * M_PI * i / n) + 0.08 * cos(4.0 * M_PI * i / n);
  return w;
}

/* ------------------------------------------------------------------ FFT */
/* Unnormalised forward transform exp(-i 2 pi k n / N); the inverse is scaled by 1/N
   (layer0.c:615-624 uses the inverse output directly). n must be a power of two. */

#define CIG_MAXLOG2 24
static wk* tw_re[CIG_MAXLOG2 + 1];
static wk* tw_im[CIG_MAXLOG2 + 1];

static int ilog2(int n) {
  int l = 0;
  while((1 << l) < n) l ++;
  return l;
}

static void ensure_twiddle(int lg) {
  if(tw_re[lg] != NULL) return;
  int n = 1 << lg;
  int h = n / 2 > 0 ? n / 2 : 1;
  wk* re = malloc(h * sizeof(wk));
  wk* im = malloc(h * sizeof(wk));
  for(int i = 0; i < h; i ++) {
    re[i] = cos(2.0 * M_PI * i / n);
    im[i] = -sin(2.0 * M_PI * i / n);
  }
  tw_im[lg] = im;
  tw_re[lg] = re;
}

/* in-place radix-2 decimation-in-time on working-precision arrays; sign = -1 forward, +1 inverse */
static void fft_core(wk* re, wk* im, int n, int sign) {
  int lg = ilog2(n);
  ensure_twiddle(lg);
  for(int i = 1, j = 0; i < n; i ++) {
    int bit = n >> 1;
    for(; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if(i < j) {
      wk t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  const wk* twr = tw_re[lg];
  const wk* twi = tw_im[lg];
  for(int len = 2; len <= n; len <<= 1) {
    int half = len >> 1;
    int stride = n / len;
    for(int base = 0; base < n; base += len) {
      for(int k = 0; k < half; k ++) {
        wk wr = twr[k * stride];
        wk wi = sign < 0 ? twi[k * stride] : -twi[k * stride];
        int a = base + k, b = a + half;
        wk tr = re[b] * wr - im[b] * wi;
        wk ti = re[b] * wi + im[b] * wr;
        re[b] = re[a] - tr; im[b] = im[a] - ti;
        re[a] += tr;        im[a] += ti;
      }
    }
  }
}

static void fft_generic(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, int n, int sign) {
  wk* re = malloc(2 * n * sizeof(wk));
  wk* im = re + n;
  for(int i = 0; i < n; i ++) {
    re[i] = xr != NULL ? xr[i] : 0;
    im[i] = xi != NULL ? xi[i] : 0;
  }
  fft_core(re, im, n, sign);
  wk scale = sign > 0 ? (wk)1.0 / n : (wk)1.0;
  if(yr != NULL) for(int i = 0; i < n; i ++) yr[i] = re[i] * scale;
  if(yi != NULL) for(int i = 0; i < n; i ++) yi[i] = im[i] * scale;
  free(re);
}

/* `buffer` (2n scratch in upstream ciglet) is accepted and ignored; in-place use is allowed. */
void fft(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, int n, FP_TYPE* buffer) {
  (void)buffer;
  fft_generic(xr, xi, yr, yi, n, -1);
}

void ifft(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, int n, FP_TYPE* buffer) {
  (void)buffer;
  fft_generic(xr, xi, yr, yi, n, +1);
}

/* ------------------------------------------------------------------ chirp-z transforms */
/* czt : Y[k] = sum_{m<n} x[m] exp(-i w0 k m), k < n, unnormalised     (dsputils.c:156-164)
   iczt: y[t] = (1/n) sum_{k<n} X[k] exp(+i w0 k t), t < n             (dsputils.c:345-348)
   Both by Bluestein's identity k m = (k^2 + m^2 - (k-m)^2) / 2 with a power-of-two circular
   convolution; chirp phases are reduced in double before sin/cos. */
static void bluestein(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, double w0, int n,
  int sign, double scale) {
  int L = 1;
  while(L < 2 * n - 1) L <<= 1;
  wk* buf = calloc(6 * (size_t)L, sizeof(wk));
  wk* ar = buf, *ai = buf + L, *br = buf + 2 * L, *bi = buf + 3 * L;
  wk* cr = buf + 4 * L, *ci = buf + 5 * L; /* chirp exp(sign * i w0 m^2 / 2) */
  for(int m = 0; m < n; m ++) {
    double ph = fmod(0.5 * w0 * (double)m * (double)m, 2.0 * M_PI);
    cr[m] = cos(ph);
    ci[m] = sign * sin(ph);
  }
  for(int m = 0; m < n; m ++) {
    wk r = xr != NULL ? xr[m] : 0, i = xi != NULL ? xi[m] : 0;
    ar[m] = r * cr[m] - i * ci[m];
    ai[m] = r * ci[m] + i * cr[m];
  }
  br[0] = cr[0]; bi[0] = -ci[0];
  for(int m = 1; m < n; m ++) {
    br[m] = br[L - m] = cr[m];
    bi[m] = bi[L - m] = -ci[m];
  }
  fft_core(ar, ai, L, -1);
  fft_core(br, bi, L, -1);
  for(int k = 0; k < L; k ++) {
    wk r = ar[k] * br[k] - ai[k] * bi[k];
    wk i = ar[k] * bi[k] + ai[k] * br[k];
    ar[k] = r; ai[k] = i;
  }
  fft_core(ar, ai, L, +1);
  wk s = scale / L;
  for(int k = 0; k < n; k ++) {
    wk r = (ar[k] * cr[k] - ai[k] * ci[k]) * s;
    wk i = (ar[k] * ci[k] + ai[k] * cr[k]) * s;
    if(yr != NULL) yr[k] = r;
    if(yi != NULL) yi[k] = i;
  }
  free(buf);
}

void czt(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, FP_TYPE omega0, int n) {
  bluestein(xr, xi, yr, yi, omega0, n, -1, 1.0);
}

void iczt(FP_TYPE* xr, FP_TYPE* xi, FP_TYPE* yr, FP_TYPE* yi, FP_TYPE omega0, int n) {
  bluestein(xr, xi, yr, yi, omega0, n, +1, 1.0 / n);
}

/* y[j] = sum_s a_s cos(2 pi f_s / fs * (j - n/2) + phi_s): time origin at sample n/2, forced by the
   equivalence with the ICZT variant (dsputils.c:345-346 vs :333). Evaluated with a complex
   rotation recurrence re-seeded every 64 samples. */
FP_TYPE* gensins(FP_TYPE* freq, FP_TYPE* ampl, FP_TYPE* phse, int nsin, FP_TYPE fs, int n) {
  FP_TYPE* y = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  double* acc = calloc(n > 0 ? n : 1, sizeof(double));
  for(int s = 0; s < nsin; s ++) {
    double w = 2.0 * M_PI * (double)freq[s] / (double)fs;
    double rr = cos(w), ri = sin(w);
    double a = ampl[s];
    double cr = 0, ci = 0;
    for(int j = 0; j < n; j ++) {
      if((j & 63) == 0) {
        double ph = fmod(w * (double)(j - n / 2), 2.0 * M_PI) + (double)phse[s];
        cr = cos(ph); ci = sin(ph);
      }
      acc[j] += a * cr;
      double t = cr * rr - ci * ri;
      ci = cr * ri + ci * rr;
      cr = t;
    }
  }
  for(int j = 0; j < n; j ++) y[j] = acc[j];
  free(acc);
  return y;
}

/* ------------------------------------------------------------------ STFT
   Per frame: window named by `window` ("blackman" / "hanning") of length nwin[i], frame centred at
   center[i] (fetch_frame alignment), zero-phase placement (window centre at buffer index 0), FFT
   of size nfft, linear magnitude and phase in radians, bins 0..nfft/2.
   Decisions: magnitudes are NOT normalised; norm_factor[i] and weight_factor[i] (when non-NULL)
   both receive sum(w_i), so that llsm_compute_spectrogram's rescale (dsputils.c:100-113) nets
   |X| * 2 / sum(w_i) for periodic windows. A window longer than nfft is time-aliased into the
   buffer (exact sampling of the DTFT). subt_mean removes the frame mean before windowing; optlv
   (upstream's fast-math level) is ignored. */
void cig_stft_forward(FP_TYPE* x, int nx, int* center, int* nwin, int nfrm, int nfft,
  char* window, int subt_mean, int optlv, FP_TYPE* norm_factor, FP_TYPE* weight_factor,
  FP_TYPE** Xmagn, FP_TYPE** Xphse) {
  (void)optlv;
  wk* re = malloc(2 * (size_t)nfft * sizeof(wk));
  wk* im = re + nfft;
  int is_blackman = ! strcmp(window, "blackman");
  for(int i = 0; i < nfrm; i ++) {
    int n = nwin[i];
    FP_TYPE* w = is_blackman ? blackman(n) : hanning(n);
    FP_TYPE* f = fetch_frame(x, nx, center[i], n);
    double mean = 0;
    if(subt_mean) mean = meanfp(f, n);
    double wsum = 0;
    for(int j = 0; j < nfft; j ++) { re[j] = 0; im[j] = 0; }
    for(int j = 0; j < n; j ++) {
      int k = ((j - n / 2) % nfft + nfft) % nfft;
      re[k] += ((double)f[j] - mean) * w[j];
      wsum += w[j];
    }
    fft_core(re, im, nfft, -1);
    for(int k = 0; k <= nfft / 2; k ++) {
      if(Xmagn != NULL) Xmagn[i][k] = sqrt(re[k] * re[k] + im[k] * im[k]);
      if(Xphse != NULL) Xphse[i][k] = atan2(im[k], re[k]);
    }
    if(norm_factor != NULL) norm_factor[i] = wsum;
    if(weight_factor != NULL) weight_factor[i] = wsum;
    free(w); free(f);
  }
  free(re);
}

/* ------------------------------------------------------------------ cepstral routines */

/* Cepstral smoothing of a LINEAR magnitude spectrum (nfft/2+1 bins); f0 in cycles per sample.
   Returns the NATURAL-LOG magnitude envelope (layer0.c:342, dsputils.c:475).
   Decision (CheapTrick-style lifter): c = IDFT(log max(S, 1e-10));
     c[q] *= sinc(f0 q) * (1.18 - 0.18 cos(2 pi f0 q)),  sinc(x) = sin(pi x) / (pi x);
   envelope = Re DFT(c). Cout (optional) receives the liftered cepstrum, nfft/2+1 values. */
FP_TYPE* cig_spec2env(FP_TYPE* S, int nfft, FP_TYPE f0, int nhar, FP_TYPE* Cout) {
  (void)nhar;
  int ns = nfft / 2 + 1;
  wk* re = malloc(2 * (size_t)nfft * sizeof(wk));
  wk* im = re + nfft;
  for(int k = 0; k < ns; k ++) {
    double v = S[k] > 1e-10 ? S[k] : 1e-10;
    re[k] = log(v); im[k] = 0;
  }
  for(int k = 1; k < nfft / 2; k ++) { re[nfft - k] = re[k]; im[nfft - k] = 0; }
  fft_core(re, im, nfft, +1);
  for(int q = 0; q < ns; q ++) {
    double xq = (double)f0 * q;
    double sinc = q == 0 ? 1.0 : sin(M_PI * xq) / (M_PI * xq);
    double l = sinc * (1.18 - 0.18 * cos(2.0 * M_PI * xq));
    wk c = re[q] / nfft * l;
    re[q] = c; im[q] = 0;
    if(q > 0 && q < nfft / 2) { re[nfft - q] = c; im[nfft - q] = 0; }
    if(Cout != NULL) Cout[q] = c;
  }
  fft_core(re, im, nfft, -1);
  FP_TYPE* env = calloc(ns, sizeof(FP_TYPE));
  for(int k = 0; k < ns; k ++) env[k] = re[k];
  free(re);
  return env;
}

/* Phase (nfft/2+1 bins) of the minimum-phase system whose natural-log magnitude is `logmagn`
   (cepstral folding; dsputils.c:497): c = IDFT(logmagn), fold c[q] *= 2 for 0 < q < nfft/2,
   zero for q > nfft/2, phase = Im DFT(c). */
FP_TYPE* minphase(FP_TYPE* logmagn, int nfft) {
  int ns = nfft / 2 + 1;
  wk* re = malloc(2 * (size_t)nfft * sizeof(wk));
  wk* im = re + nfft;
  for(int k = 0; k < ns; k ++) { re[k] = logmagn[k]; im[k] = 0; }
  for(int k = 1; k < nfft / 2; k ++) { re[nfft - k] = re[k]; im[nfft - k] = 0; }
  fft_core(re, im, nfft, +1);
  for(int q = 0; q < nfft; q ++) {
    wk c = re[q] / nfft;
    if(q > 0 && q < nfft / 2) c *= 2;
    else if(q > nfft / 2) c = 0;
    re[q] = c; im[q] = 0;
  }
  fft_core(re, im, nfft, -1);
  FP_TYPE* ph = calloc(ns, sizeof(FP_TYPE));
  for(int k = 0; k < ns; k ++) ph[k] = im[k];
  free(re);
  return ph;
}

/* ------------------------------------------------------------------ filters */

/* Zero-phase IIR filtering: forward pass, time reversal, forward pass, time reversal
   (dsputils.c:67). Decision: direct form II transposed, ZERO initial state on both passes, no
   edge padding. a[0] is assumed non-zero and is normalised out. */
static void iir_df2t(const FP_TYPE* b, int nb, const FP_TYPE* a, int na, const wk* x, wk* y, int nx) {
  int order = (nb > na ? nb : na) - 1;
  wk z[16] = {0};
  wk bb[17] = {0}, aa[17] = {0};
  for(int i = 0; i < nb; i ++) bb[i] = (wk)b[i] / (wk)a[0];
  for(int i = 0; i < na; i ++) aa[i] = (wk)a[i] / (wk)a[0];
  for(int n = 0; n < nx; n ++) {
    wk xn = x[n];
    wk yn = bb[0] * xn + z[0];
    for(int i = 0; i < order - 1; i ++)
      z[i] = bb[i + 1] * xn + z[i + 1] - aa[i + 1] * yn;
    if(order > 0) z[order - 1] = bb[order] * xn - aa[order] * yn;
    y[n] = yn;
  }
}

FP_TYPE* filtfilt(FP_TYPE* b, int nb, FP_TYPE* a, int na, FP_TYPE* x, int nx) {
  FP_TYPE* y = calloc(nx > 0 ? nx : 1, sizeof(FP_TYPE));
  wk* t0 = malloc(2 * (size_t)(nx > 0 ? nx : 1) * sizeof(wk));
  wk* t1 = t0 + nx;
  for(int i = 0; i < nx; i ++) t0[i] = x[i];
  iir_df2t(b, nb, a, na, t0, t1, nx);
  for(int i = 0; i < nx; i ++) t0[i] = t1[nx - 1 - i];
  iir_df2t(b, nb, a, na, t0, t1, nx);
  for(int i = 0; i < nx; i ++) y[i] = t1[nx - 1 - i];
  free(t0);
  return y;
}

/* Scalar random-walk Kalman filter: x_t = x_{t-1} + N(0,Q_t), z_t = x_t + N(0,R_t)
   (layer0.c:378). Returns posterior means; P_out receives posterior variances; L_out (optional)
   the total log-likelihood. Decision: initial state x_0 = z_0 with variance R_0. */
FP_TYPE* kalmanf1d(FP_TYPE* z, FP_TYPE* Q, FP_TYPE* R, int n, FP_TYPE* P_out, FP_TYPE* L_out) {
  FP_TYPE* y = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  if(n <= 0) return y;
  double x = z[0], P = R[0], L = 0;
  y[0] = x;
  if(P_out != NULL) P_out[0] = P;
  for(int t = 1; t < n; t ++) {
    double Pp = P + Q[t];
    double S = Pp + R[t];
    double K = Pp / S;
    double e = (double)z[t] - x;
    L += -0.5 * (log(2.0 * M_PI * S) + e * e / S);
    x += K * e;
    P = (1.0 - K) * Pp;
    y[t] = x;
    if(P_out != NULL) P_out[t] = P;
  }
  if(L_out != NULL) *L_out = L;
  return y;
}

/* Rauch-Tung-Striebel smoother matching kalmanf1d (layer0.c:379): y = filtered means,
   P = filtered variances. */
FP_TYPE* kalmans1d(FP_TYPE* y, FP_TYPE* P, FP_TYPE* Q, int n) {
  FP_TYPE* s = calloc(n > 0 ? n : 1, sizeof(FP_TYPE));
  if(n <= 0) return s;
  double sn = y[n - 1];
  s[n - 1] = sn;
  for(int t = n - 2; t >= 0; t --) {
    double Pp = (double)P[t] + Q[t + 1];
    double C = (double)P[t] / Pp;
    sn = (double)y[t] + C * (sn - (double)y[t]);
    s[t] = sn;
  }
  return s;
}

/* ------------------------------------------------------------------ LF glottal model */

/* Fant (1995) Rd regression; te/tp/ta relative to T0 (consistent with llsmutils.c:25-43).
   Decision: Rap is clamped to >= 1e-3 (the raw regression is negative below Rd ~ 0.21). */
lfmodel lfmodel_from_rd(FP_TYPE rd, FP_TYPE T0, FP_TYPE Ee) {
  double Rd = rd;
  double Rap = (-1.0 + 4.8 * Rd) / 100.0;
  double Rkp = (22.4 + 11.8 * Rd) / 100.0;
  double Rgp = 1.0 / (4.0 * ((0.11 * Rd / (0.5 + 1.2 * Rkp)) - Rap) / Rkp);
  if(Rap < 1e-3) Rap = 1e-3;
  lfmodel m;
  m.T0 = T0;
  m.tp = 1.0 / (2.0 * Rgp);
  m.te = m.tp * (Rkp + 1.0);
  m.ta = Rap;
  m.Ee = Ee;
  return m;
}

typedef struct { double te, tp, ta, wg, eps, alpha, E0, Ee; } lf_solved;

/* open-phase + return-phase net flow as a function of alpha (normalised time, tc = 1) */
static double lf_netflow(const lf_solved* s, double alpha) {
  double ste = sin(s -> wg * s -> te), cte = cos(s -> wg * s -> te);
  /* A1 = Ee / (-sin(wg te)) * int_0^te exp(alpha (t - te)) sin(wg t) dt */
  double den = alpha * alpha + s -> wg * s -> wg;
  double A1 = s -> Ee / (-ste) * ((alpha * ste - s -> wg * cte) + s -> wg * exp(-alpha * s -> te)) / den;
  double d = 1.0 - s -> te;
  double ex = exp(-s -> eps * d);
  double A2 = -(s -> Ee / (s -> eps * s -> ta)) * ((1.0 - ex) / s -> eps - d * ex);
  return A1 + A2;
}

static lf_solved lf_solve(lfmodel m) {
  lf_solved s;
  s.te = m.te; s.tp = m.tp; s.ta = m.ta; s.Ee = m.Ee;
  if(s.ta < 1e-6) s.ta = 1e-6;
  if(s.te > 1.0 - 1e-6) s.te = 1.0 - 1e-6;
  s.wg = M_PI / s.tp;
  /* eps ta = 1 - exp(-eps (1 - te)) by Newton from 1 / ta */
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
  /* zero net flow: bisection on alpha (net flow decreases monotonically in alpha) */
  double lo = -200.0, hi = 400.0;
  for(int it = 0; it < 100; it ++) {
    double mid = 0.5 * (lo + hi);
    if(lf_netflow(& s, mid) > 0) lo = mid; else hi = mid;
  }
  s.alpha = 0.5 * (lo + hi);
  s.E0 = -s.Ee / (exp(s.alpha * s.te) * sin(s.wg * s.te));
  return s;
}

/* Magnitude (returned, malloc'd) and optional phase of the Fourier transform of the LF
   flow-DERIVATIVE waveform at `freq` (Hz): closed-form integrals of the two LF segments
   (Doval, d'Alessandro & Henrich 2006). The callers integrate to flow by /f and -pi/2
   (llsmutils.c:114-122, layer0.c:186-189). */
FP_TYPE* lfmodel_spectrum(lfmodel model, FP_TYPE* freq, int nf, FP_TYPE* dst_phase) {
  FP_TYPE* magn = calloc(nf > 0 ? nf : 1, sizeof(FP_TYPE));
  lf_solved s = lf_solve(model);
  double T0 = model.T0;
  double d = 1.0 - s.te;
  double ste = sin(s.wg * s.te), cte = cos(s.wg * s.te);
  double exd = exp(-s.eps * d);
  for(int i = 0; i < nf; i ++) {
    double w = 2.0 * M_PI * (double)freq[i] * T0; /* radians per normalised time unit */
    double complex sc = s.alpha - I * w;
    double complex P1 = s.E0 * (cexp(sc * s.te) * (sc * ste - s.wg * cte) + s.wg) /
                        (sc * sc + s.wg * s.wg);
    double complex ew = s.eps + I * w;
    double complex t1 = (1.0 - cexp(-ew * d)) / ew;
    double complex t2 = fabs(w) > 1e-12 ? exd * (1.0 - cexp(-I * w * d)) / (I * w) : exd * d;
    double complex P2 = -(s.Ee / (s.eps * s.ta)) * cexp(-I * w * s.te) * (t1 - t2);
    double complex X = T0 * (P1 + P2);
    magn[i] = cabs(X);
    if(dst_phase != NULL) dst_phase[i] = carg(X);
  }
  return magn;
}

/* ------------------------------------------------------------------ IF detector */
/* Windowed complex-demodulation instantaneous-frequency estimator (Flanagan's phase-vocoder
   derivative form). fc, fres and the returned frequency are in cycles per sample
   (dsputils.c:79-83). Decision: Hann window of nh = 2 * round(2 / fres) + 1 samples (4 periods of
   fres); with y = sum x w e^{-j wc m}, yd = sum x w' e^{-j wc m}:
     f = fc - Im(yd / y) / (2 pi). */
ifdetector* create_ifdetector(FP_TYPE fc, FP_TYPE fres) {
  ifdetector* d = malloc(sizeof(ifdetector));
  int half = (int)round(2.0 / (double)fres);
  if(half < 2) half = 2;
  d -> fc = fc;
  d -> nh = 2 * half + 1;
  d -> hr = calloc(d -> nh, sizeof(FP_TYPE));
  d -> hi = calloc(d -> nh, sizeof(FP_TYPE));
  d -> hdr = calloc(d -> nh, sizeof(FP_TYPE));
  d -> hdi = calloc(d -> nh, sizeof(FP_TYPE));
  double L = 2.0 * half + 2.0; /* Hann support: zero at m = +-(half + 1) */
  for(int j = 0; j < d -> nh; j ++) {
    double m = j - d -> nh / 2;
    double w = 0.5 + 0.5 * cos(2.0 * M_PI * m / L);
    double wd = -0.5 * (2.0 * M_PI / L) * sin(2.0 * M_PI * m / L);
    double ph = fmod(2.0 * M_PI * (double)fc * m, 2.0 * M_PI);
    d -> hr[j] = w * cos(ph);  d -> hi[j] = -w * sin(ph);
    d -> hdr[j] = wd * cos(ph); d -> hdi[j] = -wd * sin(ph);
  }
  return d;
}

FP_TYPE ifdetector_estimate(ifdetector* ifd, FP_TYPE* x, int nx) {
  int n = nx < ifd -> nh ? nx : ifd -> nh;
  double yr = 0, yi = 0, dr = 0, di = 0;
  for(int j = 0; j < n; j ++) {
    yr += (double)x[j] * ifd -> hr[j];  yi += (double)x[j] * ifd -> hi[j];
    dr += (double)x[j] * ifd -> hdr[j]; di += (double)x[j] * ifd -> hdi[j];
  }
  double den = yr * yr + yi * yi;
  if(den < 1e-30) return ifd -> fc;
  double im = (di * yr - dr * yi) / den; /* Im(yd / y) */
  return (double)ifd -> fc - im / (2.0 * M_PI);
}

void delete_ifdetector(ifdetector* dst) {
  if(dst == NULL) return;
  free(dst -> hr); free(dst -> hi); free(dst -> hdr); free(dst -> hdi);
  free(dst);
}

/* ---- ddct (coder.c:27,149-152,189-192): Ooura's in-place cosine transform, by its published definition
   (fft4g.c header): isgn = -1  C[k] = sum_{j<n} a[j] cos(pi (j + 1/2) k / n)   (DCT)
                     isgn = +1  C[k] = sum_{j<n} a[j] cos(pi j (k + 1/2) / n)   (inverse DCT without scale;
   "a[0] *= 0.5; ddct(n, 1, a); a[j] *= 2 / n" inverts isgn = -1, which is how coder.c uses the pair).
   Direct O(n^2) summation in double: test infrastructure, sizes <= 1024. ---- */
void ddct(int n, int isgn, FP_TYPE* a) {
  double* c = calloc(n > 0 ? n : 1, sizeof(double));
  for(int k = 0; k < n; k ++) {
    double acc = 0;
    if(isgn < 0) for(int j = 0; j < n; j ++) acc += (double)a[j] * cos(M_PI * (j + 0.5) * k / n);
    else         for(int j = 0; j < n; j ++) acc += (double)a[j] * cos(M_PI * j * (k + 0.5) / n);
    c[k] = acc;
  }
  for(int k = 0; k < n; k ++) a[k] = c[k];
  free(c);
}
