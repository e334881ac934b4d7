// Launch orchestration of layer-1 (pulse-by-pulse) synthesis: llsm_synthesize with use_l1
// (layer0.c:148-287 for the deterministic part, then the same noise path as layer 0).
#pragma once
#include "driver.h"
#include "driver_layer1.h"
#include "kernels_pbp.cuh"

struct PbpScratch {
  DevBuf npulse, pulse_base, pre_rotate, len_period, pulse_size, pulses, need_hm, source_p0, y_hm, y_pbp,
         hm_nhar, hm_ampl, hm_phse;
  void release() {
    DevBuf* all[] = {&npulse, &pulse_base, &pre_rotate, &len_period, &pulse_size, &pulses, &need_hm, &source_p0,
                     &y_hm, &y_pbp, &hm_nhar, &hm_ampl, &hm_phse};
    for(DevBuf* b : all) b->release();
  }
};

#ifdef LLSM_EMU
static inline int dev_zero(void* p, size_t n, cudaStream_t) { memset(p, 0, n); return 0; }
#else
static inline int dev_zero(void* p, size_t n, cudaStream_t s) { return cudaMemsetAsync(p, 0, n, s) == cudaSuccess ? 0 : -1; }
#endif

// l1: layer-1 members (device); pbpsyn [B][F] or NULL; fr: frames with the noise model and,
// optionally, stored harmonic models (fr.nhar/ampl/phse all non-NULL) -- otherwise the HM frames
// are derived from layer 1 (llsm_frame_tolayer0 on the fly, layer0.c:264-265).
static inline int run_synth_l1(const SynthPlanDev& pd, const L1PlanDev& lp, SynthScratch& sc, PbpScratch& ps,
  const llsm_b200_conf& conf, const llsm_b200_frames& fr, const llsm_b200_layer1& l1, const int* pbpsyn,
  const llsm_b200_soptions& opt, const llsm_b200_output& out, const int* ny_utt_dev, cudaStream_t st,
  LaunchCounter* lc, bool plan_ready = false) {
  const SynthPlan& h = pd.h;
  const int B = conf.nutt, F = conf.nfrm;
  const size_t BF = (size_t)B * F;
  if(out.stride < h.ny) return LLSM_B200_EINVAL;
  const size_t obytes = (size_t)B * out.stride * 4;
  if(ps.npulse.reserve(BF * 4) || ps.pulse_base.reserve(BF * 4) || ps.pre_rotate.reserve(BF * 4) ||
     ps.len_period.reserve(BF * 4) || ps.pulse_size.reserve(BF * 4) || ps.pulses.reserve(BF * PBP_MAXP * sizeof(PbpPulse)) ||
     ps.need_hm.reserve(BF * 4) || ps.source_p0.reserve(BF * 4) || ps.y_hm.reserve(obytes) || ps.y_pbp.reserve(obytes))
    return LLSM_B200_ENOMEM;

  // harmonic models: stored or derived
  llsm_b200_frames hmf = fr;
  if(! (fr.nhar && fr.ampl && fr.phse)) {
    if(ps.hm_nhar.reserve(BF * 4) || ps.hm_ampl.reserve(BF * conf.maxnhar * 4) || ps.hm_phse.reserve(BF * conf.maxnhar * 4))
      return LLSM_B200_ENOMEM;
    int rc = run_tolayer0(lp, conf, fr.nfrm_utt, fr.f0, l1, ps.hm_nhar.as<int>(), ps.hm_ampl.as<float>(),
      ps.hm_phse.as<float>(), st, lc);
    if(rc != 0) return rc;
    hmf.nhar = ps.hm_nhar.as<int>(); hmf.ampl = ps.hm_ampl.as<float>(); hmf.phse = ps.hm_phse.as<float>();
  }

  PbpPlan plan;
  plan.npulse = ps.npulse.as<int>(); plan.pulse_base = ps.pulse_base.as<int>(); plan.pre_rotate = ps.pre_rotate.as<int>();
  plan.len_period = ps.len_period.as<float>(); plan.pulse_size = ps.pulse_size.as<int>();
  plan.pulses = ps.pulses.as<PbpPulse>(); plan.need_hm = ps.need_hm.as<int>();

  if(! plan_ready) {
  PbpPrepParams Q; memset(&Q, 0, sizeof(Q));
  Q.nfrm = F; Q.nfrm_utt = fr.nfrm_utt; Q.f0 = fr.f0; Q.rd = l1.rd; Q.nvs = l1.nvs; Q.source_p0 = ps.source_p0.as<float>();
  LLSM_LAUNCH(pbp_prep_kernel, dim3((F + 127) / 128, B), dim3(128), 0, st, Q);
  if(lc) lc->n ++;

  if(dev_zero(out.y_sin, obytes, st) || dev_zero(ps.y_pbp.p, obytes, st) || dev_zero(ps.npulse.p, BF * 4, st) ||
     dev_zero(ps.need_hm.p, BF * 4, st)) return LLSM_B200_ECUDA;

  PbpTrackParams T; memset(&T, 0, sizeof(T));
  T.nutt = B; T.nfrm = F; T.nfrm_utt = fr.nfrm_utt; T.ny_utt = ny_utt_dev; T.ny = h.ny; T.stride = out.stride;
  T.f0 = fr.f0; T.rd = l1.rd; T.vsphse = l1.vsphse; T.nvs = l1.nvs; T.vs_stride = conf.maxnhar; T.pbpsyn = pbpsyn;
  T.source_p0 = ps.source_p0.as<float>(); T.base_trunc = pd.base_trunc; T.hop = h.hop_f; T.fs = conf.fs;
  T.nspec = l1.nspec; T.plan = plan; T.y_mix = out.y_sin;
  LLSM_LAUNCH(pbp_track_kernel, dim3((B + 31) / 32), dim3(32), 0, st, T);
  if(lc) lc->n ++;
  } else {
    // plan, need_hm and the switch ramp (in out.y_sin) were produced by the host tracker and uploaded
    if(dev_zero(ps.y_pbp.p, obytes, st)) return LLSM_B200_ECUDA;
  }

  // harmonic-model frames at truncated positions, no sub-sample phase correction (layer0.c:173,263-277)
  {
    BankParams P; memset(&P, 0, sizeof(P));
    P.nfrm = F; P.maxnhar = conf.maxnhar; P.nfrm_utt = fr.nfrm_utt; P.ny_utt = ny_utt_dev;
    P.f0 = hmf.f0; P.nhar = hmf.nhar; P.ampl = hmf.ampl; P.phse = hmf.phse;
    P.hm_base = pd.base_trunc; P.hm_frac = pd.zero_frac; P.win = pd.win_hm; P.n_hm = h.n_hm;
    P.ny = h.ny; P.nsamp = out.stride; P.stride = out.stride; P.fs = conf.fs; P.hop = h.hop_f;
    P.has_options = 1; P.use_iczt = opt.use_iczt; P.iczt_a = opt.iczt_param_a; P.iczt_b = opt.iczt_param_b;
    P.frame_mask = plan.need_hm; P.y_sin = ps.y_hm.as<float>();
    if(launch_hm_bank(P, B, F, st) != 0) return LLSM_B200_ERANGE;
    if(lc) lc->n ++;
  }

  // pulses
  {
    PbpPulseParams U; memset(&U, 0, sizeof(U));
    U.nfrm = F; U.nfrm_utt = fr.nfrm_utt; U.ny_utt = ny_utt_dev; U.ny = h.ny; U.stride = out.stride;
    U.f0 = fr.f0; U.rd = l1.rd; U.vtmagn = l1.vtmagn; U.nspec = l1.nspec; U.vsphse = l1.vsphse; U.nvs = l1.nvs;
    U.vs_stride = conf.maxnhar; U.fs = conf.fs; U.fnyq = (float)((double)conf.fs / 2.0); U.lip_radius = conf.lip_radius;
    U.plan = plan; U.tw = lp.tw; U.ntw = lp.ntw; U.maxnhar = conf.maxnhar; U.y_pbp = ps.y_pbp.as<float>();
    int mp = l1_minphase_nfft(conf.maxnhar);
    U.max_size = 4096 > mp ? 4096 : mp; U.max_mp = mp;
    if(U.max_size > lp.ntw) return LLSM_B200_ERANGE;
    size_t smem = (size_t)U.max_size * 16 + ((size_t)conf.maxnhar * 5 + 6) * 4 + 16;
    if(smem > 200 * 1024) return LLSM_B200_ERANGE;
#ifndef LLSM_EMU
    cudaFuncSetAttribute(pbp_pulse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    LLSM_LAUNCH(pbp_pulse_kernel, dim3(F, B), dim3(PBP_THREADS), smem, st, U);
    if(lc) lc->n ++;
  }

  PbpMixParams M; memset(&M, 0, sizeof(M));
  M.ny = h.ny; M.stride = out.stride; M.ny_utt = ny_utt_dev; M.y_hm = ps.y_hm.as<float>(); M.y_pbp = ps.y_pbp.as<float>();
  M.y_mix_inout = out.y_sin;
  LLSM_LAUNCH(pbp_mix_kernel, dim3((out.stride + 255) / 256, B), dim3(256), 0, st, M);
  if(lc) lc->n ++;

  return run_noise_part(pd, sc, conf, fr, opt, out, ny_utt_dev, st, lc);
}
