// Host-side plans: every index, window and table the kernels need that depends only on the model
// configuration (fs, thop, nfrm, npsd ...), evaluated once on the host with EXACTLY the
// float/double expression order of the reference (the frame positions sit on
// float knife edges, so they are never recomputed on the device).
#pragma once
#include <vector>
#include <cstdint>

struct SynthPlan {
  // configuration echo
  int nfrm = 0; float fs = 0, thop = 0; int npsd = 0, nchannel = 0;
  // output length, layer0.c:643
  int ny = 0;
  // harmonic OLA (layer0.c:121-139)
  int n_hm = 0;                       // nwin = round(thop * fs) * 2
  std::vector<int> hm_base;           // round(i * thop * fs)            [nfrm]
  std::vector<float> hm_frac;         // rawidx - baseidx (float)        [nfrm]
  std::vector<float> win_hm;          // hanning(n_hm)
  std::vector<int> base_trunc;        // (int)(i * thop * fs), the PbP path's positions (layer0.c:173)
  float hop_f = 0;                    // thop * fs as a float product
  // noise envelope OLA (layer0.c:293-309)
  int n_env = 0;                      // round(thop * 2.0 * fs) in double
  std::vector<float> env_r;           // (float)((i - 1) * thop * fs)    [nfrm]
  std::vector<int> env_off;           // round(env_r[i])                 [nfrm]
  std::vector<int> env_contig;        // 1 when round(env_r[i] + j) == env_off[i] + j for every j
  std::vector<float> win_env;         // hanning(n_env)
  // noise shaping STFT (layer0.c:559-603)
  int n_ns = 0;                       // round(thop * fs * 2) in float
  int nfft_ns = 0, lg_nfft_ns = 0, nspec_ns = 0;
  float wsqr = 0;                     // float-accumulated sum of squares of win_ns
  std::vector<float> win_ns;          // hanning(n_ns)
  std::vector<int> psd_lo;            // interp1 lower knot for bin j    [nspec_ns - 1]
  std::vector<float> psd_r;           // interp1 ratio for bin j         [nspec_ns - 1]
  // noise template (dsputils.c:385-394)
  int ntemplate = 0;                  // min(20000, ny)
  int nt = 0;                         // ntemplate + 128
  // channel filters (layer0.c:541-544, dsputils.c:28-70): per channel up to two IIR sections
  struct ChanFilt { int nstage; double b[2][5]; double a[2][5]; };
  std::vector<ChanFilt> chan;         // [nchannel]; nstage == 0 -> channel absent (fmin >= fs/2)
};

// ny for a given frame count (layer0.c:643); shared by the C ABI helper.
int plan_output_length(int nfrm, float thop, float fs);
int plan_template_length(int ny);
void build_synth_plan(SynthPlan& p, int nfrm, float fs, float thop, int npsd, int nchannel,
  const float* chanfreq);

// Full-circle twiddle table exp(-2 pi i m / n), float2 interleaved, built in double.
void build_twiddle(std::vector<float>& tw, int n);

// hanning(n) / blackman(n) as the oracle's windows (periodic, double evaluation, float store)
void make_hanning(std::vector<float>& w, int n);
void make_blackman(std::vector<float>& w, int n);

// Chebyshev sub-band filter selection (dsputils.c:28-70): fills up to two {b,a} sections for the
// band [c1, c2] (cycles per sample); returns the number of sections.
int select_chebyfilt(float c1, float c2, double b[2][5], double a[2][5]);

// Tables of the chunk-parallel IIR kernel (kernels_iir.cuh) for one filter section {b, a}:
// coef[9] = {b0..b4, a1..a4} / a0 ; mpow[nlog][16] = (A^L)^(2^q), A = zero-input state transition
// of the direct-form-II-transposed realisation.
void build_iir_section(const double b[5], const double a[5], int L, int nlog, double* coef, double* mpow);
