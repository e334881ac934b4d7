// Launch orchestration of the frame coder (coder.c): the axes of llsm_create_coder (coder.c:46-74) are built on the
// host with the reference's expression types and cached per (fs, npsd, nspec, order_bap).
#pragma once
#include "driver_layer1.h"
#include "kernels_coder.cuh"

struct CoderPlanDev {
  float fs = 0; int npsd = 0, nspec = 0, order_bap = 0;
  float *psdaxis = nullptr, *faxis = nullptr, *melaxis = nullptr, *apaxis = nullptr;
  std::vector<void*> owned;
  template <class T> int up(T** dst, const std::vector<T>& src, cudaStream_t st) {
    void* d = nullptr;
    if(dev_alloc(&d, src.size() * sizeof(T)) != 0) return -1;
    owned.push_back(d);
    if(! src.empty() && dev_upload(d, src.data(), src.size() * sizeof(T), st) != 0) return -1;
    *dst = (T*)d; return 0;
  }
  static float lin(float a, float b, int i, int n) {          // linspace of the oracle's ciglet shim
    return n > 1 ? (float)((double)a + ((double)b - (double)a) * i / (n - 1)) : a;
  }
  int build(float fs_, int npsd_, int nspec_, int order_bap_, cudaStream_t st) {
    fs = fs_; npsd = npsd_; nspec = nspec_; order_bap = order_bap_;
    const float fnyq = (float)((double)fs / 2.0);
    const int nfull = (nspec - 1) * 2;
    std::vector<float> pa(npsd), fa(nspec), ma(nspec), aa(order_bap + 1);
    for(int i = 0; i < npsd; i ++) pa[i] = lin(0, fnyq, i, npsd);
    for(int i = 0; i < nspec; i ++) { float t = fnyq * 2; t = t * i; fa[i] = t / nfull; }           // coder.c:62-63
    const float mel_ceil = (float)(1125.0 * log(1.0 + (double)fnyq / 700.0));                       // freq2mel
    const float mel_floor = (float)(1125.0 * log(1.0 + (double)50.0f / 700.0));
    for(int i = 0; i < nspec; i ++) {
      float t = (mel_ceil - mel_floor) * i; t = t / nspec; t = mel_floor + t;                        // coder.c:69-70
      ma[i] = (float)(700.0 * (exp((double)t / 1125.0) - 1.0));                                     // mel2freq
    }
    for(int i = 0; i <= order_bap; i ++) aa[i] = lin(0, fnyq, i, order_bap + 1);
    int rc = up(&psdaxis, pa, st) | up(&faxis, fa, st) | up(&melaxis, ma, st) | up(&apaxis, aa, st);
    if(dev_sync(st) != 0) rc = -1;
    return rc;
  }
  void release() { for(void* p : owned) dev_free(p); owned.clear(); }
};

static inline void coder_common(CoderParams& P, const CoderPlanDev& cp, const llsm_b200_conf& conf, const int* nfrm_utt,
  int order_spec) {
  memset(&P, 0, sizeof(P));
  P.nfrm = conf.nfrm; P.nfrm_utt = nfrm_utt; P.npsd = conf.npsd; P.nspec = cp.nspec; P.order_spec = order_spec;
  P.order_bap = cp.order_bap; P.maxnhar = conf.maxnhar;
  P.fnyq = (float)((double)conf.fs / 2.0); P.lip_radius = conf.lip_radius;
  P.psdaxis = cp.psdaxis; P.faxis = cp.faxis; P.melaxis = cp.melaxis; P.apaxis = cp.apaxis;
}

static inline int run_coder_encode(const CoderPlanDev& cp, const llsm_b200_conf& conf, const int* nfrm_utt,
  const float* f0, const float* psd, const float* rd, const float* vtmagn, int order_spec, float* enc,
  cudaStream_t st, LaunchCounter* lc) {
  if(order_spec < 1 || order_spec > cp.nspec - 1 || cp.order_bap < 1 || cp.order_bap >= CODER_THREADS) return LLSM_B200_ERANGE;
  CoderParams P; coder_common(P, cp, conf, nfrm_utt, order_spec);
  P.f0 = f0; P.psd = psd; P.rd = rd; P.vtmagn = vtmagn; P.enc = enc;
  size_t smem = (size_t)4 * order_spec * 8 + (size_t)2 * cp.nspec * 4 + (size_t)(order_spec + 2) * 4 + sizeof(LfSolved) + 64;
  if(smem > 200 * 1024) return LLSM_B200_ERANGE;
#ifndef LLSM_EMU
  cudaFuncSetAttribute(coder_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  LLSM_LAUNCH(coder_encode_kernel, dim3(conf.nfrm, conf.nutt), dim3(CODER_THREADS), smem, st, P);
  if(lc) lc->n ++;
  return 0;
}

static inline int run_coder_decode(const CoderPlanDev& cp, const L1PlanDev& lp, const llsm_b200_conf& conf,
  const int* nfrm_utt, const float* enc, int order_spec, int use_layer1, float* f0, float* rd, float* psd, int* nhar,
  float* ampl, float* phse, float* vtmagn, float* vsphse, cudaStream_t st, LaunchCounter* lc) {
  if(order_spec < 1 || order_spec > cp.nspec - 1 || cp.order_bap < 1) return LLSM_B200_ERANGE;
  const int mp_nfft = use_layer1 ? 0 : l1_minphase_nfft(conf.maxnhar);
  if(mp_nfft > lp.ntw) return LLSM_B200_ERANGE;
  CoderParams P; coder_common(P, cp, conf, nfrm_utt, order_spec);
  P.enc_in = enc; P.use_layer1 = use_layer1; P.o_f0 = f0; P.o_rd = rd; P.o_psd = psd; P.o_nhar = nhar;
  P.o_ampl = ampl; P.o_phse = phse; P.o_vtmagn = vtmagn; P.o_vsphse = vsphse;
  P.tw = lp.tw; P.ntw = lp.ntw; P.max_nfft = mp_nfft;
  size_t smem = (size_t)mp_nfft * 16 + (size_t)4 * order_spec * 8 + (size_t)3 * cp.nspec * 4
    + (size_t)(order_spec + cp.order_bap + 2) * 4 + ((size_t)conf.maxnhar * 5 + 2) * 4 + sizeof(LfSolved) + 64;
  if(smem > 200 * 1024) return LLSM_B200_ERANGE;
#ifndef LLSM_EMU
  cudaFuncSetAttribute(coder_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  LLSM_LAUNCH(coder_decode_kernel, dim3(conf.nfrm, conf.nutt), dim3(CODER_THREADS), smem, st, P);
  if(lc) lc->n ++;
  return 0;
}
