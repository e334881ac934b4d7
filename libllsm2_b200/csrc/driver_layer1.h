// Launch orchestration of the layer-1 conversions (layer1.c).
#pragma once
#include "driver.h"
#include "kernels_layer1.cuh"

struct L1PlanDev {
  float *model = nullptr, *rd_list = nullptr;    // cached glottal model (dsputils.c:514-535)
  float2* tw = nullptr; int ntw = 0;
  std::vector<void*> owned;
  template <class T> int up(T** dst, const std::vector<T>& src, cudaStream_t st) {
    void* d = nullptr;
    if(dev_alloc(&d, src.size() * sizeof(T)) != 0) return -1;
    owned.push_back(d);
    if(! src.empty() && dev_upload(d, src.data(), src.size() * sizeof(T), st) != 0) return -1;
    *dst = (T*)d; return 0;
  }
  int build(cudaStream_t st) {
    std::vector<float> rd(RD_NCAND), model((size_t)RD_NCAND * RD_NHAR);
    for(int i = 0; i < RD_NCAND; i ++) rd[i] = (float)(0.02 + (3.0 - 0.02) * i / (RD_NCAND - 1));   // linspace(0.02, 3.0, 64)
    const float f0 = 200.0f;
    for(int i = 0; i < RD_NCAND; i ++) {
      LfSolved s = lf_solve(lf_from_rd(rd[i], (float)(1.0 / (double)f0), 1.0f));
      for(int j = 0; j < RD_NHAR; j ++) {
        float freq = (float)((double)f0 * (1.0 + j));
        double m, p; lf_spectrum(s, (double)freq, &m, &p);
        float v = (float)m;
        v = (float)((double)v / (j + 1.0));
        model[(size_t)i * RD_NHAR + j] = v * v;
      }
    }
    ntw = 8192;
    std::vector<float> twh; build_twiddle(twh, ntw);
    int rc = up(&model_dev(), model, st) | up(&rd_list, rd, st);
    float* t = nullptr; rc |= up(&t, twh, st); tw = (float2*)t;
    if(dev_sync(st) != 0) rc = -1;
    return rc;
  }
  float*& model_dev() { return model; }
  void release() { for(void* p : owned) dev_free(p); owned.clear(); }
};

static inline int l1_minphase_nfft(int nhar) {
  int n = (int)pow(2.0, ceil(log2((double)(nhar > 1 ? nhar : 1)) + 2.0));
  return n > 64 ? n : 64;
}

static inline int run_tolayer1(const L1PlanDev& lp, const llsm_b200_conf& conf, const llsm_b200_frames& fr,
  int nfft, const llsm_b200_layer1& out, cudaStream_t st, LaunchCounter* lc) {
  const int B = conf.nutt, F = conf.nfrm;
  int lg = 0; while((1 << lg) < nfft) lg ++;
  if((1 << lg) != nfft || nfft < 64 || nfft > lp.ntw) return LLSM_B200_ERANGE;
  int mp_nfft = l1_minphase_nfft(conf.maxnhar);
  int max_nfft = nfft > mp_nfft ? nfft : mp_nfft;
  if(max_nfft > lp.ntw) return LLSM_B200_ERANGE;

  RdFitParams R; memset(&R, 0, sizeof(R));
  R.nfrm = F; R.nfrm_utt = fr.nfrm_utt; R.f0 = fr.f0; R.nhar = fr.nhar; R.ampl = fr.ampl; R.maxnhar = conf.maxnhar;
  R.lip_radius = conf.lip_radius; R.model = lp.model; R.rd_list = lp.rd_list; R.rd = out.rd;
  LLSM_LAUNCH(rd_fit_kernel, dim3(F, B), dim3(RD_NCAND), 0, st, R);
  if(lc) lc->n ++;

  RdSmoothParams S; memset(&S, 0, sizeof(S));
  S.nfrm = F; S.nfrm_utt = fr.nfrm_utt; S.rd = out.rd;
  S.order = (int)round(0.02 / (double)conf.thop);                  // layer1.c:75-76
  size_t ssm = (size_t)F * 8 + 16;
  if(ssm > 200 * 1024) return LLSM_B200_ERANGE;
#ifndef LLSM_EMU
  if(ssm > 48 * 1024) cudaFuncSetAttribute(rd_smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
#endif
  LLSM_LAUNCH(rd_smooth_kernel, dim3(B), dim3(256), ssm, st, S);
  if(lc) lc->n ++;

  ToLayer1Params T; memset(&T, 0, sizeof(T));
  T.nfrm = F; T.nfrm_utt = fr.nfrm_utt; T.f0 = fr.f0; T.rd = out.rd; T.nhar = fr.nhar; T.ampl = fr.ampl;
  T.phse = fr.phse; T.maxnhar = conf.maxnhar; T.fnyq = (float)((double)conf.fs / 2.0); T.lip_radius = conf.lip_radius;
  T.nfft = nfft; T.lg_nfft = lg; T.nspec = nfft / 2 + 1; T.tw = lp.tw; T.ntw = lp.ntw; T.max_nfft = max_nfft;
  T.vtmagn = out.vtmagn; T.vsphse = out.vsphse; T.nvs = out.nvs;
  size_t smem = (size_t)max_nfft * 16 + ((size_t)conf.maxnhar * 4 + 2 + L1_THREADS) * 4 + 16;
  if(smem > 200 * 1024) return LLSM_B200_ERANGE;
#ifndef LLSM_EMU
  cudaFuncSetAttribute(tolayer1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  LLSM_LAUNCH(tolayer1_kernel, dim3(F, B), dim3(L1_THREADS), smem, st, T);
  if(lc) lc->n ++;
  return 0;
}

static inline int run_tolayer0(const L1PlanDev& lp, const llsm_b200_conf& conf, const int* nfrm_utt,
  const float* f0, const llsm_b200_layer1& in, int* nhar, float* ampl, float* phse, cudaStream_t st,
  LaunchCounter* lc) {
  const int B = conf.nutt, F = conf.nfrm;
  int mp_nfft = l1_minphase_nfft(conf.maxnhar);
  if(mp_nfft > lp.ntw) return LLSM_B200_ERANGE;
  ToLayer0Params P; memset(&P, 0, sizeof(P));
  P.nfrm = F; P.nfrm_utt = nfrm_utt; P.f0 = f0; P.rd = in.rd; P.vtmagn = in.vtmagn; P.nspec = in.nspec;
  P.vsphse = in.vsphse; P.nvs = in.nvs; P.vs_stride = conf.maxnhar; P.maxnhar = conf.maxnhar;
  P.fnyq = (float)((double)conf.fs / 2.0); P.lip_radius = conf.lip_radius;
  P.tw = lp.tw; P.ntw = lp.ntw; P.max_nfft = mp_nfft;
  P.nhar_out = nhar; P.ampl = ampl; P.phse = phse;
  size_t smem = (size_t)mp_nfft * 16 + ((size_t)conf.maxnhar * 4 + 2) * 4 + 16;
  if(smem > 200 * 1024) return LLSM_B200_ERANGE;
#ifndef LLSM_EMU
  cudaFuncSetAttribute(tolayer0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  LLSM_LAUNCH(tolayer0_kernel, dim3(F, B), dim3(L1_THREADS), smem, st, P);
  if(lc) lc->n ++;
  return 0;
}
