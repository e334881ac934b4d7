// Launch orchestration shared by the CUDA library (api.cu) and the CPU thread emulation used by the
// unit tests (tests/emu/emu_api.cpp). "Device" memory is cudaMalloc'd in the product and plain
// malloc'd under LLSM_EMU.
#pragma once
#include "../../include/llsm_b200.h"
#include "plan.h"
#include "kernels_synth.cuh"
#include "kernels_iir_smem.cuh"
#include <vector>
#include <string>
#include <map>
#include <cstring>

#ifdef LLSM_EMU
static inline int dev_alloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : -1; }
static inline void dev_free(void* p) { free(p); }
static inline int dev_upload(void* d, const void* h, size_t n, cudaStream_t) { memcpy(d, h, n); return 0; }
static inline int dev_download(void* h, const void* d, size_t n, cudaStream_t) { memcpy(h, d, n); return 0; }
static inline int dev_sync(cudaStream_t) { return 0; }
static inline int dev_memset(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline const char* dev_last_error() { return nullptr; }
#else
static inline int dev_alloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess ? 0 : -1; }
static inline void dev_free(void* p) { if(p) cudaFree(p); }
static inline int dev_upload(void* d, const void* h, size_t n, cudaStream_t s) {
  return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) == cudaSuccess ? 0 : -1;
}
static inline int dev_download(void* h, const void* d, size_t n, cudaStream_t s) {
  return cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) == cudaSuccess ? 0 : -1;
}
static inline int dev_sync(cudaStream_t s) { return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -1; }
static inline int dev_memset(void* d, int v, size_t n, cudaStream_t s) { return cudaMemsetAsync(d, v, n, s) == cudaSuccess ? 0 : -1; }
static inline const char* dev_last_error() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
#endif

// growable device scratch buffer
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  int reserve(size_t n) {
    if(n <= cap) return 0;
    dev_free(p); p = nullptr; cap = 0;
    if(dev_alloc(&p, n) != 0) return -1;
    cap = n; return 0;
  }
  void release() { dev_free(p); p = nullptr; cap = 0; }
  template <class T> T* as() { return (T*)p; }
};

// device-resident copy of a SynthPlan
struct SynthPlanDev {
  SynthPlan h;
  int *hm_base = nullptr, *env_off = nullptr, *psd_lo = nullptr, *env_contig = nullptr, *base_trunc = nullptr;
  float* zero_frac = nullptr;
  float *hm_frac = nullptr, *win_hm = nullptr, *env_r = nullptr, *win_env = nullptr,
        *win_ns = nullptr, *psd_r = nullptr;
  float2* tw_ns = nullptr;
  double *iir_coef = nullptr, *iir_mpow = nullptr;   // template filters, chunk length iir_L
  int iir_L = 0;
  // shared-memory resident variant (kernels_iir_smem.cuh): cluster size (0: not applicable) and its tables
  int iis_cs = 0, iis_L = 0; double *iis_mpow = nullptr, *iis_wts = nullptr;
  std::vector<void*> owned;
  template <class T> int up(T** dst, const std::vector<T>& src, cudaStream_t st) {
    void* d = nullptr;
    if(dev_alloc(&d, src.size() * sizeof(T)) != 0) return -1;
    owned.push_back(d);
    if(! src.empty() && dev_upload(d, src.data(), src.size() * sizeof(T), st) != 0) return -1;
    *dst = (T*)d;
    return 0;
  }
  int build(int nfrm, float fs, float thop, int npsd, int nchannel, const float* chanfreq,
    cudaStream_t st) {
    build_synth_plan(h, nfrm, fs, thop, npsd, nchannel, chanfreq);
    std::vector<float> tw;
    build_twiddle(tw, h.nfft_ns);
    int rc = 0;
    rc |= up(&hm_base, h.hm_base, st); rc |= up(&hm_frac, h.hm_frac, st);
    rc |= up(&win_hm, h.win_hm, st);   rc |= up(&env_r, h.env_r, st);
    rc |= up(&env_off, h.env_off, st); rc |= up(&win_env, h.win_env, st);
    rc |= up(&env_contig, h.env_contig, st); rc |= up(&base_trunc, h.base_trunc, st);
    { std::vector<float> z(nfrm, 0.0f); rc |= up(&zero_frac, z, st); if(dev_sync(st) != 0) rc = -1; }
    rc |= up(&win_ns, h.win_ns, st);   rc |= up(&psd_lo, h.psd_lo, st);
    rc |= up(&psd_r, h.psd_r, st);
    float* twd = nullptr; rc |= up(&twd, tw, st); tw_ns = (float2*)twd;
    iir_L = ((h.nt + IIR_NT - 1) / IIR_NT + IIR_T - 1) & ~(IIR_T - 1);
    std::vector<double> coef((size_t)LLSM_B200_MAXCHANNEL * 2 * 9, 0.0), mpow((size_t)LLSM_B200_MAXCHANNEL * 2 * IIR_NLOG * 16, 0.0);
    for(int c = 0; c < nchannel; c ++)
      for(int s2 = 0; s2 < h.chan[c].nstage; s2 ++)
        build_iir_section(h.chan[c].b[s2], h.chan[c].a[s2], iir_L, IIR_NLOG,
          &coef[((size_t)c * 2 + s2) * 9], &mpow[((size_t)c * 2 + s2) * IIR_NLOG * 16]);
    rc |= up(&iir_coef, coef, st); rc |= up(&iir_mpow, mpow, st);
    iir_smem_geometry(h.nt, iis_cs, iis_L);
    std::vector<double> mpow9, wts;
    if(iis_cs > 0) {
      std::vector<double> coef2(9);
      mpow9.assign((size_t)LLSM_B200_MAXCHANNEL * 2 * IIS_NLOG * 16, 0.0); wts.assign((size_t)LLSM_B200_MAXCHANNEL * 2 * iis_L * 4, 0.0);
      for(int c = 0; c < nchannel; c ++)
        for(int s2 = 0; s2 < h.chan[c].nstage; s2 ++)
          build_iir_smem_section(h.chan[c].b[s2], h.chan[c].a[s2], iis_L, coef2.data(),
            &mpow9[((size_t)c * 2 + s2) * IIS_NLOG * 16], &wts[((size_t)c * 2 + s2) * iis_L * 4]);
      rc |= up(&iis_mpow, mpow9, st); rc |= up(&iis_wts, wts, st);
    }
    if(dev_sync(st) != 0) rc = -1;   // the host vectors must outlive the async copies
    return rc;
  }
  void release() { for(void* p : owned) dev_free(p); owned.clear(); }
};

// output samples a noise-shaper CTA owns (its eight warps take 16 frames a round: the default keeps the frames that touch
// a segment within a whole number of rounds at a 5 ms hop)
static inline int shape_seg() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_SHAPE_SEG"); v = e ? atoi(e) : 6016; if(v < 1024) v = 1024; if(v > 16384) v = 16384; }
  return v;
}
static inline int iir_variant() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_IIR_VARIANT"); v = e ? atoi(e) : 1; }
  return v;
}

struct SynthScratch {
  DevBuf colored, y_exc, ny_utt;
};

// launch counter of a context; when `ev` is set (llsm_b200_set_kernel_timing) every pipeline also records a named
// event after each of its kernels so that their times can be read back live (bench.py's per-kernel rooflines)
#define LLSM_KT_MAX 48
struct LaunchCounter {
  long long n = 0;
#ifndef LLSM_EMU
  cudaEvent_t* ev = nullptr;      // [LLSM_KT_MAX] or NULL
  const char* names[LLSM_KT_MAX];
  int mark = 0;
#endif
};
// name == "start": the interval that ends at this mark is not a kernel of ours (it is skipped when the times are read)
static inline void lc_mark(LaunchCounter* lc, cudaStream_t st, const char* name) {
#ifndef LLSM_EMU
  if(lc && lc->ev && lc->mark < LLSM_KT_MAX) { lc->names[lc->mark] = name; cudaEventRecord(lc->ev[lc->mark ++], st); }
#else
  (void)lc; (void)st; (void)name;
#endif
}

// Harmonic component only. frames/y_sin are device pointers. has_options = 0 reproduces the
// options == NULL call of the analysis residual (layer0.c:498).
static inline int run_harmonics(const SynthPlanDev& pd, const llsm_b200_conf& conf,
  const llsm_b200_frames& fr, const llsm_b200_soptions* opt, const int* ny_utt_dev,
  float* y_sin, int ny_valid, int nsamp, int stride, cudaStream_t st, LaunchCounter* lc,
  int frame_lo = 0, int frame_hi = 0, const float* sub_from = nullptr, int sub_stride = 0, bool residual_tc = false) {
  BankParams P;
  memset(&P, 0, sizeof(P));
  P.nfrm = conf.nfrm; P.maxnhar = conf.maxnhar;
  P.nfrm_utt = fr.nfrm_utt; P.ny_utt = ny_utt_dev;
  P.f0 = fr.f0; P.nhar = fr.nhar; P.ampl = fr.ampl; P.phse = fr.phse;
  P.hm_base = pd.hm_base; P.hm_frac = pd.hm_frac; P.win = pd.win_hm;
  P.n_hm = pd.h.n_hm; P.ny = ny_valid; P.nsamp = nsamp; P.stride = stride;
  P.fs = conf.fs; P.hop = pd.h.hop_f;
  P.has_options = opt != nullptr;
  if(opt) { P.use_iczt = opt->use_iczt; P.iczt_a = opt->iczt_param_a; P.iczt_b = opt->iczt_param_b; }
  P.y_sin = y_sin; P.frame_lo = frame_lo; P.frame_hi = frame_hi;
  P.sub_from = sub_from; P.sub_stride = sub_stride; P.residual_tc = residual_tc ? 1 : 0;
  if(launch_hm_bank(P, conf.nutt, conf.nfrm, st) != 0) return LLSM_B200_ERANGE;
  if(lc) lc->n += 1;
  return 0;
}

// Noise component + final mix (layer0.c:652-659): templates, excitation, shaping, y = y_sin + y_noise.
// out.y_sin must already hold the harmonic component.
static inline int run_noise_part(const SynthPlanDev& pd, SynthScratch& sc, const llsm_b200_conf& conf,
  const llsm_b200_frames& fr, const llsm_b200_soptions& opt, const llsm_b200_output& out,
  const int* ny_utt_dev, cudaStream_t st, LaunchCounter* lc, int frame_lo = 0, int frame_hi = 0, int utt_base = 0);

// Full layer-0 synthesis on device pointers (llsm_synthesize, layer0.c:636-664).
static inline int run_synth_l0(const SynthPlanDev& pd, SynthScratch& sc, const llsm_b200_conf& conf,
  const llsm_b200_frames& fr, const llsm_b200_soptions& opt, const llsm_b200_output& out,
  const int* ny_utt_dev, cudaStream_t st, LaunchCounter* lc, int frame_lo = 0, int frame_hi = 0, int utt_base = 0) {
  const SynthPlan& h = pd.h;
  const int nch = conf.nchannel;
  if(nch < 1 || nch > LLSM_B200_MAXCHANNEL) return LLSM_B200_EINVAL;
  if(out.stride < h.ny) return LLSM_B200_EINVAL;

  // 1. harmonic component
  lc_mark(lc, st, "start");
  int rc = run_harmonics(pd, conf, fr, &opt, ny_utt_dev, out.y_sin, h.ny, out.stride, out.stride,
    st, lc, frame_lo, frame_hi);
  if(rc != 0) return rc;
  lc_mark(lc, st, "hm_bank");
  return run_noise_part(pd, sc, conf, fr, opt, out, ny_utt_dev, st, lc, frame_lo, frame_hi, utt_base);
}

static inline int run_noise_part(const SynthPlanDev& pd, SynthScratch& sc, const llsm_b200_conf& conf,
  const llsm_b200_frames& fr, const llsm_b200_soptions& opt, const llsm_b200_output& out,
  const int* ny_utt_dev, cudaStream_t st, LaunchCounter* lc, int frame_lo, int frame_hi, int utt_base) {
  const SynthPlan& h = pd.h;
  const int B = conf.nutt, nch = conf.nchannel;
  const int tstride = (h.nt + 3) & ~3;
  if(sc.colored.reserve((size_t)B * nch * tstride * 4) != 0) return LLSM_B200_ENOMEM;
  if(sc.y_exc.reserve((size_t)B * out.stride * 4) != 0) return LLSM_B200_ENOMEM;

  // 2. band-limited noise templates (dsputils.c:385-394): white fill, chunk-parallel filtfilt
  unsigned mask = 0;
  {
    WhiteParams W; memset(&W, 0, sizeof(W));
    W.nseq = B * nch; W.nt = h.nt; W.ostride = tstride; W.white = opt.white; W.seed = opt.seed; W.out = sc.colored.as<float>();
    W.seq_base = utt_base * nch;
    LLSM_LAUNCH(white_fill_kernel, dim3((h.nt / 4 + 256) / 256, B * nch), dim3(256), 0, st, W);
    if(lc) lc->n += 1;
    lc_mark(lc, st, "white_fill");
    IirParams I; memset(&I, 0, sizeof(I));
    I.nchannel = nch; I.n = h.nt; I.L = pd.iir_L; I.y = sc.colored.as<float>(); I.ystride = tstride; I.vec_ok = 1;
    I.coef = pd.iir_coef; I.mpow = pd.iir_mpow;
    for(int c = 0; c < nch; c ++) {
      I.nstage[c] = h.chan[c].nstage;
      if(h.chan[c].nstage > 0) mask |= 1u << c;
    }
    bool done = false;
    if(pd.iis_cs > 0 && iir_variant() == 1) {      // sequences resident in (distributed) shared memory
      IirSmemParams Q; Q.base = I; Q.mpow = pd.iis_mpow; Q.wts = pd.iis_wts; Q.L = pd.iis_L;
      done = launch_iir_smem(Q, B * nch, pd.iis_cs, st) == 0;
    }
    if(! done) LLSM_LAUNCH(iir_filtfilt_kernel, dim3(B * nch), dim3(IIR_NT), 0, st, I);
    if(lc) lc->n += 1;
    lc_mark(lc, st, "iir_filtfilt");
  }

  // 3. excitation
  ExcParams E;
  memset(&E, 0, sizeof(E));
  E.nfrm = conf.nfrm; E.nchannel = nch; E.maxnhar_e = conf.maxnhar_e;
  E.nfrm_utt = fr.nfrm_utt; E.ny_utt = ny_utt_dev;
  E.f0 = fr.f0; E.edc = fr.edc; E.enhar = fr.enhar; E.eampl = fr.eampl; E.ephse = fr.ephse;
  E.env_r = pd.env_r; E.env_off = pd.env_off; E.env_contig = pd.env_contig; E.win_env = pd.win_env; E.n_env = h.n_env;
  E.ny = h.ny; E.nsamp = out.stride; E.stride = out.stride; E.fs = conf.fs;
  E.has_options = 1; E.use_iczt = opt.use_iczt; E.iczt_a = opt.iczt_param_a; E.iczt_b = opt.iczt_param_b;
  E.colored = sc.colored.as<float>(); E.nt = h.nt; E.ntemplate = h.ntemplate; E.tstride = tstride;
  E.chan_mask = mask; E.hop = h.hop_f;
  if(frame_hi > 0) {                 // excitation is only needed under the owned frames' windows
    E.samp_lo = h.hm_base[frame_lo] - h.n_ns / 2 - 2;
    E.samp_hi = h.hm_base[frame_hi - 1] + h.n_ns / 2 + 2;
    if(E.samp_lo < 0) E.samp_lo = 0;
  }
  E.y_exc = sc.y_exc.as<float>();
  if(launch_noise_excitation(E, B, st) != 0) return LLSM_B200_ERANGE;
  if(lc) lc->n += 1;
  lc_mark(lc, st, "noise_excitation");

  // 4. shaping + mix
  ShapeParams S;
  memset(&S, 0, sizeof(S));
  S.nfrm = conf.nfrm; S.npsd = conf.npsd;
  S.nfrm_utt = fr.nfrm_utt; S.ny_utt = ny_utt_dev;
  S.psd = fr.psd; S.psdres = fr.psdres;
  S.center = pd.hm_base; S.win = pd.win_ns;
  S.n_ns = h.n_ns; S.nfft = h.nfft_ns; S.lg_nfft = h.lg_nfft_ns; S.nspec = h.nspec_ns;
  S.wsqr = h.wsqr; S.fs = conf.fs;
  S.psd_lo = pd.psd_lo; S.psd_r = pd.psd_r; S.tw = pd.tw_ns;
  S.y_exc = sc.y_exc.as<float>(); S.stride_exc = out.stride;
  S.y_sin = out.y_sin; S.y_noise = out.y_noise; S.y = out.y;
  S.ny = h.ny; S.nsamp = out.stride; S.stride = out.stride;
  S.seg = shape_seg(); S.frame_lo = frame_lo; S.frame_hi = frame_hi;
  if(h.nfft_ns > 8192) return LLSM_B200_ERANGE;
  if(launch_noise_shape(S, B, st) != 0) return LLSM_B200_ERANGE;
  if(lc) lc->n += 1;
  lc_mark(lc, st, "noise_shape");
  return 0;
}
