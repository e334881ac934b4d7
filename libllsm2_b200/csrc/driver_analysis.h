// Launch orchestration of the layer-0 analysis path (llsm_analyze, layer0.c:478-511).
#pragma once
#include "driver.h"
#include "kernels_analysis.cuh"
#include <cmath>

struct AnaKey {
  int nfrm, npsd, nchannel; float fs, thop; float cf[LLSM_B200_MAXCHANNEL];
  bool operator<(const AnaKey& o) const { return memcmp(this, &o, sizeof(AnaKey)) < 0; }
};

// host plan of the analysis path: sizes and tables fixed by the configuration
struct AnaPlan {
  int nwin = 0, nfft = 0, lg_nfft = 0, nspec = 0, nfft_s = 0, lg_nfft_s = 0;
  float win_power = 0, std_norm = 0, std_norm_blackman = 0;
  std::vector<float> win_psd;
  std::vector<int> ip_k; std::vector<float> ip_r;
  struct ChanFilt { int nstage; double b[2][5]; double a[2][5]; };
  ChanFilt chan[LLSM_B200_MAXCHANNEL];
  unsigned use_x_mask = 0;
  // Blackman windows of the harmonic estimator (dsputils.c:190-199) for every even length 2 h, h = 1 .. bw_cap, stored
  // from the centre outwards: bwin[bw_off[h] + n] = w(h +- n) = 0.42 + 0.5 cos(2 pi n / 2h) + 0.08 cos(4 pi n / 2h),
  // n = 0 .. h (the periodic window is symmetric about its centre sample); bw_sum[h] = (FP_TYPE) sum of the 2 h
  // taps. Offsets are multiples of four floats so that a window is a 16-byte aligned bulk-copy source.
  int bw_cap = 0;
  std::vector<float> bwin, bw_sum; std::vector<int> bw_off;
};

static inline int fpad_host(int j) { return j + (j >> 4); }
static inline int noise_spec_variant() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_NS_VARIANT"); v = e ? atoi(e) : 1; }
  return v;
}
static inline int ilog2_ceil(int n) { int l = 0; while((1 << l) < n) l ++; return l; }

static inline void build_ana_plan(AnaPlan& p, float fs, float thop, int npsd, int nchannel,
  const float* chanfreq) {
  // layer0.c:320-325
  float t = thop * 4; t = t * fs;
  p.nwin = (int)round((double)t);
  p.nfft = (int)pow(2.0, ceil(log2((double)p.nwin)));
  p.lg_nfft = ilog2_ceil(p.nfft);
  p.nspec = p.nfft / 2 + 1;
  p.nfft_s = (int)pow(2.0, ceil(log2(0.03 * (double)fs)));
  p.lg_nfft_s = ilog2_ceil(p.nfft_s);
  // blackman(nwin) and its float-accumulated power (dsputils.c:248-258)
  make_blackman(p.win_psd, p.nwin);
  float wp = 0;
  for(int i = 0; i < p.nwin; i ++) { float sq = p.win_psd[i] * p.win_psd[i]; wp = wp + sq; }
  p.win_power = wp;
  // standard normaliser of llsm_compute_spectrogram (dsputils.c:100-105): sum of hanning(1024)
  std::vector<float> h; make_hanning(h, 1024);
  double ws = 0; for(int i = 0; i < 1024; i ++) ws += h[i];
  float sn = (float)ws; sn = sn * 0.5f;
  p.std_norm = sn;
  std::vector<float> bw; make_blackman(bw, 1024);
  double wb = 0; for(int i = 0; i < 1024; i ++) wb += bw[i];
  float snb = (float)wb; snb = snb * 0.5f;
  p.std_norm_blackman = snb;
  // interp1u(0, fs / 2, v, nspec, linspace(0, fs / 2, npsd), npsd) (layer0.c:388-396), exclusive end
  float x1 = (float)((double)fs / 2.0);
  double step = ((double)x1 - 0.0) / p.nspec;
  p.ip_k.resize(npsd); p.ip_r.resize(npsd);
  for(int j = 0; j < npsd; j ++) {
    float xq = npsd > 1 ? (float)(0.0 + ((double)x1 - 0.0) * j / (npsd - 1)) : 0.0f;
    double pos = ((double)xq - 0.0) / step;
    if(! (pos > 0)) { p.ip_k[j] = 0; p.ip_r[j] = 0; continue; }
    if(pos >= p.nspec - 1) { p.ip_k[j] = p.nspec - 1; p.ip_r[j] = 0; continue; }
    int k = (int)pos;
    p.ip_k[j] = k; p.ip_r[j] = (float)(pos - k);
  }
  // window table of the harmonic estimator
  p.bw_cap = (int)ceil((double)fs / mma_min_f0() * 2.0) + 4;
  p.bw_off.assign(p.bw_cap + 1, 0); p.bw_sum.assign(p.bw_cap + 1, 0.f);
  {
    size_t total = 0;
    for(int hh = 1; hh <= p.bw_cap; hh ++) { p.bw_off[hh] = (int)total; total += (size_t)((hh + 1 + 3) & ~3); }
    p.bwin.assign(total + 4, 0.f);
    for(int hh = 1; hh <= p.bw_cap; hh ++) {
      const int n2 = 2 * hh;
      float* w = &p.bwin[p.bw_off[hh]];
      for(int n = 0; n <= hh; n ++)
        w[n] = (float)(0.42 + 0.5 * cos(2.0 * M_PI * n / n2) + 0.08 * cos(4.0 * M_PI * n / n2));
      double acc = 0;                                     // sumfp over m = 0 .. 2h - 1, m = h + n | h - n
      for(int m = 0; m < n2; m ++) acc += (double)w[m < hh ? hh - m : m - hh];
      p.bw_sum[hh] = (float)acc;
    }
  }
  // channel filters (layer0.c:434-440)
  p.use_x_mask = 0;
  for(int c = 0; c < LLSM_B200_MAXCHANNEL; c ++) p.chan[c].nstage = 0;
  for(int c = 0; c < nchannel; c ++) {
    float fmin = c == 0 ? 0.0f : chanfreq[c - 1];
    float fmax = c == nchannel - 1 ? (float)((double)fs / 2.0) : chanfreq[c];
    if((double)fmin > 6000.0) p.use_x_mask |= 1u << c;
    p.chan[c].nstage = select_chebyfilt(fmin / fs, fmax / fs, p.chan[c].b, p.chan[c].a);
  }
}

struct AnaPlanDev {
  AnaPlan h;
  float *win_psd = nullptr, *ip_r = nullptr; int* ip_k = nullptr;
  float *bwin = nullptr, *bw_sum = nullptr; int* bw_off = nullptr;
  float2 *tw_s = nullptr, *tw_p = nullptr, *tw_pp = nullptr;   // tw_pp: 8192-point table for HMPP
  // chunk-parallel IIR tables of the sub-band filters, valid for sequences of iir_nx samples
  DevBuf iir_coef, iir_mpow; int iir_nx = -1, iir_L = 0; int nchannel = 0;
  std::vector<double> h_coef, h_mpow;
  // shared-memory resident variant (kernels_iir_smem.cuh): cluster size (0: not applicable), chunk length, tables
  DevBuf iis_mpow, iis_wts; int iis_cs = 0, iis_L = 0;
  std::vector<double> h_mpow9, h_wts;
  std::vector<void*> owned;
  template <class T> int up(T** dst, const std::vector<T>& src, cudaStream_t st) {
    void* d = nullptr;
    if(dev_alloc(&d, src.size() * sizeof(T)) != 0) return -1;
    owned.push_back(d);
    if(! src.empty() && dev_upload(d, src.data(), src.size() * sizeof(T), st) != 0) return -1;
    *dst = (T*)d; return 0;
  }
  int build(float fs, float thop, int npsd, int nchannel, const float* chanfreq, cudaStream_t st) {
    build_ana_plan(h, fs, thop, npsd, nchannel, chanfreq);
    this->nchannel = nchannel;
    std::vector<float> tws, twp;
    build_twiddle(tws, h.nfft_s); build_twiddle(twp, h.nfft);
    int rc = 0;
    rc |= up(&win_psd, h.win_psd, st); rc |= up(&ip_k, h.ip_k, st); rc |= up(&ip_r, h.ip_r, st);
    rc |= up(&bwin, h.bwin, st); rc |= up(&bw_sum, h.bw_sum, st); rc |= up(&bw_off, h.bw_off, st);
    float* a = nullptr; rc |= up(&a, tws, st); tw_s = (float2*)a;
    float* b = nullptr; rc |= up(&b, twp, st); tw_p = (float2*)b;
    std::vector<float> twpp; build_twiddle(twpp, 8192);
    float* c2 = nullptr; rc |= up(&c2, twpp, st); tw_pp = (float2*)c2;
    if(dev_sync(st) != 0) rc = -1;
    return rc;
  }
  int ensure_iir(int nx, cudaStream_t st) {
    if(nx == iir_nx) return 0;
    if(dev_sync(st) != 0) return -1;           // previous tables may still be in use / in flight
    iir_L = ((nx + IIR_NT - 1) / IIR_NT + IIR_T - 1) & ~(IIR_T - 1);
    h_coef.assign((size_t)LLSM_B200_MAXCHANNEL * 2 * 9, 0.0);
    h_mpow.assign((size_t)LLSM_B200_MAXCHANNEL * 2 * IIR_NLOG * 16, 0.0);
    for(int c = 0; c < nchannel; c ++)
      for(int s2 = 0; s2 < h.chan[c].nstage; s2 ++)
        build_iir_section(h.chan[c].b[s2], h.chan[c].a[s2], iir_L, IIR_NLOG,
          &h_coef[((size_t)c * 2 + s2) * 9], &h_mpow[((size_t)c * 2 + s2) * IIR_NLOG * 16]);
    if(iir_coef.reserve(h_coef.size() * 8) || iir_mpow.reserve(h_mpow.size() * 8)) return -1;
    if(dev_upload(iir_coef.p, h_coef.data(), h_coef.size() * 8, st) ||
       dev_upload(iir_mpow.p, h_mpow.data(), h_mpow.size() * 8, st)) return -1;
    iir_smem_geometry(nx, iis_cs, iis_L);
    if(iis_cs > 0) {
      std::vector<double> coef2(9);
      h_mpow9.assign((size_t)LLSM_B200_MAXCHANNEL * 2 * IIS_NLOG * 16, 0.0); h_wts.assign((size_t)LLSM_B200_MAXCHANNEL * 2 * iis_L * 4, 0.0);
      for(int c = 0; c < nchannel; c ++)
        for(int s2 = 0; s2 < h.chan[c].nstage; s2 ++)
          build_iir_smem_section(h.chan[c].b[s2], h.chan[c].a[s2], iis_L, coef2.data(),
            &h_mpow9[((size_t)c * 2 + s2) * IIS_NLOG * 16], &h_wts[((size_t)c * 2 + s2) * iis_L * 4]);
      if(iis_mpow.reserve(h_mpow9.size() * 8) || iis_wts.reserve(h_wts.size() * 8)) return -1;
      if(dev_upload(iis_mpow.p, h_mpow9.data(), h_mpow9.size() * 8, st) || dev_upload(iis_wts.p, h_wts.data(), h_wts.size() * 8, st)) return -1;
    }
    if(dev_sync(st) != 0) return -1;
    iir_nx = nx;
    return 0;
  }
  void release() { for(void* p : owned) dev_free(p); owned.clear(); iir_coef.release(); iir_mpow.release(); iis_mpow.release(); iis_wts.release(); iir_nx = -1; }
};

struct AnaScratch {
  DevBuf x_sin, x_res, ce, env, lpsd, res, filt, pvar, nfft_utt;
  DevBuf long_list;      // frames the staged kernels leave to the general ones: [0] = count, [1 ..] = b * nfrm + i
  void release() { nfft_utt.release(); long_list.release(); x_sin.release(); x_res.release(); ce.release(); env.release(); lpsd.release(); res.release(); filt.release(); pvar.release(); }
};

// one FFT size per utterance for the peak-picking method (llsm_get_fftsize, dsputils.c:318-326)
static inline int run_utt_fftsize(AnaScratch& sc, const llsm_b200_conf& conf, const llsm_b200_aoptions& opt,
  const float* f0, const int* nfrm_utt, cudaStream_t st, LaunchCounter* lc) {
  if(sc.nfft_utt.reserve((size_t)conf.nutt * 4) != 0) return LLSM_B200_ENOMEM;
  MinF0Params M; memset(&M, 0, sizeof(M));
  M.nfrm = conf.nfrm; M.nfrm_utt = nfrm_utt; M.f0 = f0; M.fs = conf.fs; M.rel_winsize = opt.rel_winsize;
  M.nfft_utt = sc.nfft_utt.as<int>();
  LLSM_LAUNCH(utt_fftsize_kernel, dim3(conf.nutt), dim3(128), 0, st, M);
  if(lc) lc->n ++;
  return 0;
}

// llsm_harmonic_analysis (dsputils.c:175-228) of nsig signals per utterance ([B][nsig][sstride]) at the frame centres
// of the plan: CZT (direct DFT at the harmonics) or peak picking; optionally the short-time means (edc_o).
static inline int run_harmonic_pass(const SynthPlanDev& sp, AnaPlanDev& ap, AnaScratch& sc, const llsm_b200_conf& conf,
  const llsm_b200_aoptions& opt, const float* f0, const int* nfrm_utt, const float* sig, int nsig, int nx, int sstride,
  int maxnhar, int* nhar_o, float* ampl_o, float* phse_o, float* edc_o, bool prep_fftsize, cudaStream_t st,
  LaunchCounter* lc) {
  const int B = conf.nutt, F = conf.nfrm;
  const int max_half = (int)ceil((double)conf.fs / 20.0 * (double)opt.rel_winsize / 4.0 * 2.0) + 4;
  if(opt.hm_method == 1) {
    HarmDftParams H; memset(&H, 0, sizeof(H));
    H.nfrm = F; H.nfrm_utt = nfrm_utt; H.sig = sig; H.nsig = nsig; H.nx = nx; H.xstride = sstride;
    H.f0 = f0; H.center = sp.hm_base; H.fs = conf.fs; H.rel_winsize = opt.rel_winsize;
    H.maxnhar = maxnhar; H.nhar_out = nhar_o; H.ampl = ampl_o; H.phse = phse_o; H.max_half = max_half;
    H.edc = edc_o; H.thop = conf.thop;
    if(sc.long_list.reserve(((size_t)B * F + 1) * 4) != 0) return LLSM_B200_ENOMEM;
    H.bwin = ap.bwin; H.bw_off = ap.bw_off; H.bw_sum = ap.bw_sum; H.bw_cap = ap.h.bw_cap;
    H.long_list = sc.long_list.as<int>();
    H.hop_max = 1;
    for(int i = 1; i < F; i ++) H.hop_max = std::max(H.hop_max, sp.h.hm_base[i] - sp.h.hm_base[i - 1]);
    if(launch_harmonic_dft(H, B, st) != 0) return LLSM_B200_ERANGE;
  } else {
    if(prep_fftsize) { int rc = run_utt_fftsize(sc, conf, opt, f0, nfrm_utt, st, lc); if(rc) return rc; }
    HarmPpParams H; memset(&H, 0, sizeof(H));
    H.nfrm = F; H.nfrm_utt = nfrm_utt; H.sig = sig; H.nsig = nsig; H.nx = nx; H.xstride = sstride;
    H.f0 = f0; H.center = sp.hm_base; H.nfft_utt = sc.nfft_utt.as<int>(); H.fs = conf.fs;
    H.rel_winsize = opt.rel_winsize; H.std_norm = ap.h.std_norm_blackman; H.maxnhar = maxnhar;
    H.nhar_out = nhar_o; H.ampl = ampl_o; H.phse = phse_o; H.tw = ap.tw_pp; H.ntw = 8192; H.max_nfft = 8192;
    H.bwin = ap.bwin; H.bw_off = ap.bw_off; H.bw_cap = ap.h.bw_cap;
    // The transform size is per utterance (device data); the shared-memory footprint decides how many CTAs share an SM
    // (8192 points: one, 2048: seven). One launch per size tier, each serving the utterances of its tier (the others'
    // CTAs leave at once), instead of one launch sized for the largest transform.
    int lo = 0;
    for(int tier = 2048; tier <= 8192; tier <<= 1) {
      H.min_nfft = lo; H.max_nfft = tier;
      size_t smem = (size_t)tier * 16 + 16;
#ifndef LLSM_EMU
      cudaFuncSetAttribute(harmonic_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
      LLSM_LAUNCH(harmonic_pp_kernel, dim3(F, B * nsig), dim3(HP_THREADS), smem, st, H);
      if(lc && tier > 2048) lc->n ++;
      lo = tier;
    }
  }
  if(lc) lc->n ++;
  return 0;
}

// Second stream of the analysis: after the residual the noise-PSD chain (spectra -> Kalman / RTS -> output) and the
// sub-band envelope chain (IIR -> envelope harmonics) are independent; on two streams the HBM-bound smoother runs beside
// the FP64-bound filter and the output stage beside the envelope kernel. Not used while per-kernel timing is recorded.
struct AnaFork {
#ifndef LLSM_EMU
  cudaStream_t st2 = nullptr; cudaEvent_t fork = nullptr, join = nullptr;
#else
  int unused = 0;
#endif
};
static inline int ana_overlap_enabled() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_ANA_OVERLAP"); v = e ? atoi(e) : 1; }
  return v;
}

// x: [B][xstride] device; fr: device output arrays; x_res_out optional [B][xstride]
static inline int run_analyze_l0(const SynthPlanDev& sp, AnaPlanDev& ap, AnaScratch& sc,
  const llsm_b200_conf& conf, const llsm_b200_aoptions& opt, const float* x, int nx, int xstride,
  const llsm_b200_frames_out& fr, const int* nfrm_utt, float* x_res_out, cudaStream_t st,
  LaunchCounter* lc, const AnaFork* fork = nullptr) {
  const int B = conf.nutt, F = conf.nfrm, nch = conf.nchannel;
  const AnaPlan& h = ap.h;
  if(opt.hm_method != 0 && opt.hm_method != 1) return LLSM_B200_EINVAL;
  if(h.nfft > 8192 || h.nfft_s > 8192) return LLSM_B200_ERANGE;
  const size_t BF = (size_t)B * F;
  const int cst = (nx + 3) & ~3;     // row stride of the sub-band signals (16-byte aligned rows)
  if(sc.x_sin.reserve((size_t)B * nx * 4) || sc.x_res.reserve((size_t)B * nx * 4) ||
     sc.ce.reserve((size_t)B * nch * cst * 4) || sc.env.reserve(BF * h.nspec * 4) ||
     sc.lpsd.reserve(BF * h.nspec * 4) || sc.res.reserve(BF * h.nspec * 4) ||
     sc.filt.reserve(BF * h.nspec * 4) || sc.pvar.reserve(BF * h.nspec * 4)) return LLSM_B200_ENOMEM;
  if(ap.ensure_iir(nx, st) != 0) return LLSM_B200_ENOMEM;
  float* x_res = x_res_out ? x_res_out : sc.x_res.as<float>();
  const int rstride = x_res_out ? xstride : nx;

  lc_mark(lc, st, "start");
  // 1. F0 refinement (dsputils.c:72-94)
  if(opt.f0_refine) {
    RefineParams R; memset(&R, 0, sizeof(R));
    R.nfrm = F; R.nfrm_utt = nfrm_utt; R.x = x; R.nx = nx; R.xstride = xstride;
    R.center = sp.hm_base; R.fs = conf.fs; R.f0 = fr.f0;
    LLSM_LAUNCH(refine_f0_kernel, dim3((F + RF_WARPS - 1) / RF_WARPS, B), dim3(32 * RF_WARPS), 0, st, R);
    if(lc) lc->n ++;
    lc_mark(lc, st, "refine_f0");
  }

  // 2. harmonic analysis of x (dsputils.c:175-228)
  auto harmonic_pass = [&](const float* sig, int nsig, int sstride, int maxnhar, int* nhar_o, float* ampl_o, float* phse_o,
                           float* edc_o, cudaStream_t hs) -> int {
    return run_harmonic_pass(sp, ap, sc, conf, opt, fr.f0, nfrm_utt, sig, nsig, nx, sstride, maxnhar, nhar_o, ampl_o,
      phse_o, edc_o, false, hs, lc);
  };
  if(opt.hm_method == 0) { int rc = run_utt_fftsize(sc, conf, opt, fr.f0, nfrm_utt, st, lc); if(rc) return rc; }
  { int rc = harmonic_pass(x, 1, xstride, conf.maxnhar, fr.nhar, fr.ampl, fr.phse, nullptr, st); if(rc) return rc; }
  lc_mark(lc, st, opt.hm_method == 1 ? "harmonic_czt" : "harmonic_pp");

  // 3. residual: x - resynthesised sinusoids (layer0.c:498-501; options == NULL, ny = nx)
  {
    llsm_b200_frames fin; memset(&fin, 0, sizeof(fin));
    fin.nfrm_utt = nfrm_utt; fin.f0 = fr.f0; fin.nhar = fr.nhar; fin.ampl = fr.ampl; fin.phse = fr.phse;
    // Which bank: the tensor-core one for the CZT estimator (into x_sin, then one subtraction pass), the direct FP32
    // summation for peak picking, with the subtraction riding in its write-once output (kernels_synth.cuh: launch_hm_bank
    // has the measurements); LLSM_RESIDUAL_TC = 0 / 1 forces either. (The subtraction inside the tensor-core bank's tile
    // emit measured 3.52 ms against 1.94 ms for bank + subtraction pass: the four read-back warps are that kernel's
    // critical path and would wait on the global read of x.)
    const int rtc = residual_tc_enabled();
    const bool tc = (rtc >= 0 ? rtc != 0 : opt.hm_method == 1) && bank_tc_enabled() && conf.maxnhar >= 24;
    int rc = tc ? run_harmonics(sp, conf, fin, nullptr, nullptr, sc.x_sin.as<float>(), nx, nx, nx, st, lc, 0, 0, nullptr, 0, true)
                : run_harmonics(sp, conf, fin, nullptr, nullptr, x_res, nx, nx, rstride, st, lc, 0, 0, x, xstride);
    if(rc != 0) return rc;
    if(tc) {
      LLSM_LAUNCH(residual_kernel, dim3((nx + 255) / 256, B), dim3(256), 0, st,
        x, (const float*)sc.x_sin.as<float>(), x_res, nx, xstride, nx, rstride);
      if(lc) lc->n ++;
    }
    lc_mark(lc, st, "residual_bank");
  }

  // 4. noise PSD (layer0.c:318-415): spectra, Kalman / RTS smoother, output stage
  auto noise_spec_stage = [&](cudaStream_t s) {
    NoiseSpecParams N; memset(&N, 0, sizeof(N));
    N.nfrm = F; N.nfrm_utt = nfrm_utt; N.x = x; N.xstride = xstride; N.x_res = x_res; N.rstride = rstride;
    N.nx = nx; N.f0 = fr.f0; N.center = sp.hm_base; N.fs = conf.fs;
    N.nwin = h.nwin; N.nfft = h.nfft; N.lg_nfft = h.lg_nfft; N.nspec = h.nspec;
    N.nfft_s = h.nfft_s; N.lg_nfft_s = h.lg_nfft_s;
    N.win_psd = ap.win_psd; N.win_power = h.win_power; N.std_norm = h.std_norm;
    N.tw_s = ap.tw_s; N.tw_p = ap.tw_p;
    N.env = sc.env.as<float>(); N.lpsd = sc.lpsd.as<float>();
    if(h.nfft_s == 2048 && h.nfft == 1024 && noise_spec_variant() == 1) {
      // register-resident transforms, one warp per frame pair
      const size_t smem = noise_spec_warp_smem();
#ifndef LLSM_EMU
      cudaFuncSetAttribute(noise_spec_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
      const int npair = (F + 1) / 2;
      LLSM_LAUNCH(noise_spec_warp_kernel, dim3((npair + NSW_WARPS - 1) / NSW_WARPS, B), dim3(NSW_THREADS), smem, s, N);
    } else {
      size_t smem = (size_t)(fpad_host(std::max(h.nfft, h.nfft_s)) + 1) * 16 + 16;
#ifndef LLSM_EMU
      cudaFuncSetAttribute(noise_spec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
      LLSM_LAUNCH(noise_spec_kernel, dim3((F + 1) / 2, B), dim3(NS_THREADS), smem, s, N);
    }
    if(lc) lc->n ++;
    lc_mark(lc, s, "noise_spec");
  };
  auto kalman_stage = [&](cudaStream_t s) {
    KalmanParams K; memset(&K, 0, sizeof(K));
    K.nfrm = F; K.nspec = h.nspec; K.nfrm_utt = nfrm_utt;
    K.env = sc.env.as<float>(); K.lpsd = sc.lpsd.as<float>(); K.res = sc.res.as<float>();
    K.filt = sc.filt.as<float>(); K.pvar = sc.pvar.as<float>();
    LLSM_LAUNCH(noise_kalman_kernel, dim3((h.nspec + 127) / 128, B), dim3(128), 0, s, K);
    if(lc) lc->n ++;
    lc_mark(lc, s, "noise_kalman");
  };
  auto psd_out_stage = [&](cudaStream_t s) {
    PsdOutParams O; memset(&O, 0, sizeof(O));
    O.nfrm = F; O.nspec = h.nspec; O.npsd = conf.npsd; O.nfrm_utt = nfrm_utt;
    O.lpsd = sc.lpsd.as<float>(); O.res = sc.res.as<float>(); O.ip_k = ap.ip_k; O.ip_r = ap.ip_r;
    O.fs = conf.fs; O.psd = fr.psd; O.psdres = fr.psdres;
    LLSM_LAUNCH(noise_psd_out_kernel, dim3(F, B), dim3(128), 0, s, O);
    if(lc) lc->n ++;
    lc_mark(lc, s, "noise_psd_out");
  };

  // 5. noise envelope per channel (layer0.c:417-469): sub-band IIR, then harmonics + short-time means of the envelopes
  auto iir_stage = [&](cudaStream_t s) {
    IirParams I; memset(&I, 0, sizeof(I));
    I.nchannel = nch; I.n = nx; I.L = ap.iir_L; I.y = sc.ce.as<float>(); I.ystride = cst; I.vec_ok = 1;
    I.src_a = x_res; I.sa_stride = rstride; I.src_b = x; I.sb_stride = xstride;
    I.src_b_mask = h.use_x_mask; I.src_per_utt = 1;
    I.coef = ap.iir_coef.as<double>(); I.mpow = ap.iir_mpow.as<double>();
    for(int c = 0; c < nch; c ++) I.nstage[c] = h.chan[c].nstage;
    I.square = 1;
    bool done = false;
    if(ap.iis_cs > 0 && iir_variant() == 1) {      // sequences resident in (distributed) shared memory
      IirSmemParams Q; Q.base = I; Q.mpow = ap.iis_mpow.as<double>(); Q.wts = ap.iis_wts.as<double>(); Q.L = ap.iis_L;
      done = launch_iir_smem(Q, B * nch, ap.iis_cs, s, s != st) == 0;
    }
    if(! done) LLSM_LAUNCH(iir_filtfilt_kernel, dim3(B * nch), dim3(IIR_NT), 0, s, I);
    if(lc) lc->n ++;
    lc_mark(lc, s, "subband_iir");
  };
  // CZT pass: the short-time means ride along in the same kernel; peak picking keeps the separate kernel
  const bool dc_fused = conf.maxnhar_e > 0 && opt.hm_method == 1;
  auto envelope_stage = [&](cudaStream_t s) -> int {
    if(conf.maxnhar_e > 0) {
      int rc = harmonic_pass(sc.ce.as<float>(), nch, cst, conf.maxnhar_e, fr.enhar, fr.eampl, fr.ephse,
        dc_fused ? fr.edc : nullptr, s);
      if(rc) return rc;
      lc_mark(lc, s, "envelope_harmonics");
    }
    if(dc_fused) return 0;
    DcParams D; memset(&D, 0, sizeof(D));
    D.nfrm = F; D.nchannel = nch; D.nfrm_utt = nfrm_utt; D.ce = sc.ce.as<float>(); D.cstride = cst; D.nx = nx;
    D.f0 = fr.f0; D.center = sp.hm_base; D.fs = conf.fs; D.thop = conf.thop; D.edc = fr.edc;
    LLSM_LAUNCH(frame_dc_kernel, dim3(F, B * nch), dim3(128), 0, s, D);
    if(lc) lc->n ++;
    lc_mark(lc, s, "frame_dc");
    return 0;
  };

#ifndef LLSM_EMU
  if(fork && fork->st2 && dc_fused && ana_overlap_enabled() && ! (lc && lc->ev)) {
    // The second stream joins in after the spectra: the sub-band filter's grid is persistent (one cluster slot per SM),
    // so all of it is dispatched at once and the smoother's CTAs, launched right behind it, share the SMs with it -- a
    // grid larger than the machine would keep the block scheduler to itself until its last wave (measured: two full-size
    // kernels on two streams take exactly the sum of their times).
    noise_spec_stage(st);
    cudaEventRecord(fork->fork, st);
    cudaStreamWaitEvent(fork->st2, fork->fork, 0);
    iir_stage(fork->st2);
    kalman_stage(st);
    const int rc = envelope_stage(fork->st2);
    psd_out_stage(st);
    cudaEventRecord(fork->join, fork->st2);
    cudaStreamWaitEvent(st, fork->join, 0);
    return rc;
  }
#else
  (void)fork;
#endif
  noise_spec_stage(st); kalman_stage(st); psd_out_stage(st);
  iir_stage(st);
  return envelope_stage(st);
}
