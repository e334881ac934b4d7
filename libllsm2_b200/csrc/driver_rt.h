// Streaming synthesis driver (llsmrt.c): shared hop clock on the host, per-stream state on the device.
#pragma once
#include "driver.h"
#include "driver_layer1.h"
#include "kernels_rt.cuh"
#include <cmath>
#include <map>

struct RtWindow { float* win = nullptr; float wsqr = 0; };

struct RtBatch {
  // configuration
  int S = 0, nch = 0, npsd = 0, maxnhar = 0, maxnhar_e = 0;
  float fs = 0, thop = 0;
  int use_iczt = 1, use_l1 = 0; float iczt_a = 0.275f, iczt_b = 2.26f, lip_radius = 1.5f;
  int ntemplate = 0, cap = 0, nfft = 0, lg_nfft = 0, nspec = 0;
  unsigned chan_mask = 0;
  // clock (llsmrt.c:40-53)
  float cycle = 0;
  int curr_nhop = 0, next_nhop = 0, exc_cycle = 0, sin_pos = 0;
  int cur = 0, mod_cur = 0, exc_cur = 0;   // ring positions: sin & noise, modulation, excitation
  int has_prev = 0;
  // device state
  DevBuf mod, sinb, noise, exc, tmpl, prev_psd, out_p, out_ap, colored, iir_coef, iir_mpow;
  int* psd_lo = nullptr; float* psd_r = nullptr; float2* tw = nullptr;
  std::map<int, RtWindow> wins;
  // layer-1 (pulse-by-pulse) streaming
  int nspec_l1 = 0, max_pulse = 4096, host_tracker = 0, prev_nhop = 0;
  DevBuf pulse_f, pulse_b, pstate, plans, hm_nhar, hm_ampl, hm_phse;
  std::vector<RtPbpState> hstate;          // tracker state when the effect callbacks keep it on the host
  std::vector<RtPbpPlan> hplans;
  std::vector<void*> owned;
  DevBuf stage[12];

  template <class T> int up(T** dst, const std::vector<T>& src, cudaStream_t st) {
    void* d = nullptr;
    if(dev_alloc(&d, src.size() * sizeof(T)) != 0) return -1;
    owned.push_back(d);
    if(! src.empty() && dev_upload(d, src.data(), src.size() * sizeof(T), st) != 0) return -1;
    if(dev_sync(st) != 0) return -1;
    *dst = (T*)d; return 0;
  }
  void release() {
    DevBuf* all[] = {&mod, &sinb, &noise, &exc, &tmpl, &prev_psd, &out_p, &out_ap, &colored, &iir_coef, &iir_mpow,
      &pulse_f, &pulse_b, &pstate, &plans, &hm_nhar, &hm_ampl, &hm_phse};
    for(DevBuf* b : all) b->release();
    for(auto& s : stage) s.release();
    for(void* p : owned) dev_free(p);
    owned.clear(); wins.clear();
  }

  RtWindow* window(int H, cudaStream_t st) {
    auto it = wins.find(H);
    if(it != wins.end()) return &it->second;
    std::vector<float> w; make_hanning(w, 2 * H);          // hanning_2(nwin), llsmrt.c:119-120
    float wsqr = 0;
    for(int i = 0; i < 2 * H; i ++) { float sq = w[i] * w[i]; wsqr = wsqr + sq; }   // llsmrt.c:428-430
    RtWindow rw; rw.wsqr = wsqr;
    if(up(&rw.win, w, st) != 0) return nullptr;
    wins[H] = rw;
    return &wins[H];
  }

  // llsm_update_cycle (llsmrt.c:110-129): clock part; the ring advance is done by the feed kernel
  void update_cycle(int* prev_out) {
    int prev_nhop = curr_nhop;
    cycle = cycle + thop;
    float cf = cycle * fs;
    curr_nhop = (int)floor((double)cf);
    cycle = cycle - (float)prev_nhop / fs;
    this->prev_nhop = prev_nhop;             // the per-stream pulse / pbp_offset updates run in the feed kernel
    float nf = (cycle + thop) * fs;
    next_nhop = (int)floor((double)nf);
    if(prev_out) *prev_out = prev_nhop;
  }
};

// create: sizes, tables, templates, warm-up (llsm_create_rtsynth_buffer, llsmrt.c:157-223)
static inline int rt_create(RtBatch& R, const llsm_b200_conf& conf, const llsm_b200_soptions& opt, int use_l1,
  cudaStream_t st, LaunchCounter* lc) {
  R.S = conf.nutt; R.nch = conf.nchannel; R.npsd = conf.npsd; R.maxnhar = conf.maxnhar; R.maxnhar_e = conf.maxnhar_e;
  R.fs = conf.fs; R.thop = conf.thop; R.use_iczt = opt.use_iczt; R.iczt_a = opt.iczt_param_a; R.iczt_b = opt.iczt_param_b;
  R.use_l1 = use_l1;
  R.ntemplate = (int)conf.fs;                              // ret -> ntemplate = options -> fs
  R.cap = (int)((double)conf.fs * 0.2);                    // ninternal
  float t = conf.thop * conf.fs;
  R.nfft = pow2_ceil(log2((double)t * 2.2 + 32));          // llsmrt.c:181 (also llsm_b200_rt_fft_size)          // llsmrt.c:181
  R.lg_nfft = 0; while((1 << R.lg_nfft) < R.nfft) R.lg_nfft ++;
  R.nspec = R.nfft / 2 + 1;
  if(R.cap < 2 * R.nfft || R.nfft > 8192 || R.nch < 1 || R.nch > LLSM_B200_MAXCHANNEL) return LLSM_B200_ERANGE;
  const int S = R.S, nch = R.nch;
  // interp1 plan of llsm_spectrum_from_envelope(psd_axis, psd, npsd, nspec - 1, fs / 2) (llsmrt.c:456-457)
  {
    float fnyq = (float)((double)conf.fs / 2.0);
    std::vector<float> xi(R.npsd);
    for(int i = 0; i < R.npsd; i ++) xi[i] = R.npsd > 1 ? (float)(((double)fnyq) * i / (R.npsd - 1)) : 0.0f;
    int nq = R.nspec - 1;
    std::vector<int> lo(nq); std::vector<float> rr(nq);
    for(int j = 0; j < nq; j ++) {
      float v = (float)j * fnyq; v = v / (float)nq;
      if(! (v > xi[0])) { lo[j] = 0; rr[j] = 0; continue; }
      if(v >= xi[R.npsd - 1]) { lo[j] = R.npsd - 1; rr[j] = 0; continue; }
      int a = 0, b = R.npsd - 1;
      while(b - a > 1) { int mid = (a + b) / 2; if(xi[mid] <= v) a = mid; else b = mid; }
      lo[j] = a; rr[j] = (float)(((double)v - xi[a]) / ((double)xi[b] - xi[a]));
    }
    std::vector<float> tw; build_twiddle(tw, R.nfft);
    float* twd = nullptr;
    if(R.up(&R.psd_lo, lo, st) || R.up(&R.psd_r, rr, st) || R.up(&twd, tw, st)) return LLSM_B200_ENOMEM;
    R.tw = (float2*)twd;
  }
  const size_t ring = (size_t)S * R.cap * 4;
  if(R.mod.reserve(ring * nch) || R.sinb.reserve(ring) || R.noise.reserve(ring) || R.exc.reserve(ring) ||
     R.tmpl.reserve((size_t)S * nch * R.ntemplate * 4) || R.prev_psd.reserve((size_t)S * R.npsd * 4))
    return LLSM_B200_ENOMEM;
  // ---- clock: curr_nhop = 1; update_cycle; cycle = 0; sin_pos (llsmrt.c:213-217)
  R.cycle = 0; R.exc_cycle = 0; R.has_prev = 0;
  R.curr_nhop = 1;
  R.update_cycle(nullptr);
  R.cycle = 0;
  R.sin_pos = -R.curr_nhop * 2 - R.nfft / 2;
  R.cur = R.curr_nhop % R.cap;                             // appendblank(curr_nhop) on empty rings
  // ---- templates (llsm_make_exc_template, llsmrt.c:93-107)
  const int nsrc = (R.ntemplate < 20000 ? R.ntemplate : 20000) + 128;
  const int tstride = (nsrc + 3) & ~3;
  if(R.colored.reserve((size_t)S * nch * tstride * 4)) return LLSM_B200_ENOMEM;
  std::vector<double> coef((size_t)LLSM_B200_MAXCHANNEL * 2 * 9, 0.0), mpow((size_t)LLSM_B200_MAXCHANNEL * 2 * IIR_NLOG * 16, 0.0);
  const int L = ((nsrc + IIR_NT - 1) / IIR_NT + IIR_T - 1) & ~(IIR_T - 1);
  IirParams I; memset(&I, 0, sizeof(I));
  R.chan_mask = 0;
  for(int c = 0; c < nch; c ++) {
    float fmin = c == 0 ? 0.0f : conf.chanfreq[c - 1];
    float fmax = c == nch - 1 ? (float)((double)conf.fs / 2.0) : conf.chanfreq[c];
    if((double)fmin >= (double)conf.fs / 2.0) break;       // llsmrt.c:99
    double b[2][5], a[2][5];
    int ns = select_chebyfilt(fmin / conf.fs, fmax / conf.fs, b, a);
    for(int s2 = 0; s2 < ns; s2 ++)
      build_iir_section(b[s2], a[s2], L, IIR_NLOG, &coef[((size_t)c * 2 + s2) * 9], &mpow[((size_t)c * 2 + s2) * IIR_NLOG * 16]);
    I.nstage[c] = ns;
    R.chan_mask |= 1u << c;
  }
  if(R.iir_coef.reserve(coef.size() * 8) || R.iir_mpow.reserve(mpow.size() * 8)) return LLSM_B200_ENOMEM;
  if(dev_upload(R.iir_coef.p, coef.data(), coef.size() * 8, st) || dev_upload(R.iir_mpow.p, mpow.data(), mpow.size() * 8, st) ||
     dev_sync(st)) return LLSM_B200_ECUDA;
  WhiteParams W; memset(&W, 0, sizeof(W));
  W.nseq = S * nch; W.nt = nsrc; W.ostride = tstride; W.white = opt.white; W.seed = opt.seed; W.out = R.colored.as<float>();
  LLSM_LAUNCH(white_fill_kernel, dim3((nsrc / 4 + 256) / 256, S * nch), dim3(256), 0, st, W);
  I.nchannel = nch; I.n = nsrc; I.L = L; I.y = R.colored.as<float>(); I.ystride = tstride; I.vec_ok = 1;
  I.coef = R.iir_coef.as<double>(); I.mpow = R.iir_mpow.as<double>();
  LLSM_LAUNCH(iir_filtfilt_kernel, dim3(S * nch), dim3(IIR_NT), 0, st, I);
  RtTemplateParams T; memset(&T, 0, sizeof(T));
  T.nseq = S * nch; T.nt_src = nsrc; T.src_stride = tstride; T.ntemplate = R.ntemplate;
  T.colored = R.colored.as<float>(); T.tmpl = R.tmpl.as<float>();
  LLSM_LAUNCH(rt_template_kernel, dim3((R.ntemplate + 255) / 256, S * nch), dim3(256), 0, st, T);
  // ---- warm-up (llsm_fill_excitation_buffers, llsmrt.c:149-155)
  RtWarmParams Wm; memset(&Wm, 0, sizeof(Wm));
  Wm.S = S; Wm.nchannel = nch; Wm.cap = R.cap; Wm.ntemplate = R.ntemplate; Wm.chan_mask = R.chan_mask;
  Wm.mod = R.mod.as<float>(); Wm.sin_ = R.sinb.as<float>(); Wm.noise = R.noise.as<float>(); Wm.exc = R.exc.as<float>();
  Wm.tmpl = R.tmpl.as<float>();
  R.mod_cur = (R.cur + R.cap - 1) % R.cap;                 // ninternal - 1 appends after the blank
  Wm.hole = (R.mod_cur + 0) % R.cap;                       // the slot the appends never reach
  LLSM_LAUNCH(rt_warmup_kernel, dim3((R.cap + 255) / 256, S), dim3(256), 0, st, Wm);
  const int chunk = R.cap / 5;
  R.exc_cur = (5 * chunk) % R.cap;
  R.exc_cycle = (5 * chunk) % R.ntemplate;
  if(lc) lc->n += 4;
  if(use_l1) {
    R.lip_radius = conf.lip_radius;
    if(R.pulse_f.reserve(ring) || R.pulse_b.reserve(ring) || R.pstate.reserve((size_t)S * sizeof(RtPbpState)) ||
       R.plans.reserve((size_t)S * sizeof(RtPbpPlan))) return LLSM_B200_ENOMEM;
    if(dev_memset(R.pulse_f.p, 0, ring, st) || dev_memset(R.pulse_b.p, 0, ring, st)) return LLSM_B200_ECUDA;
    RtPbpState s0; s0.pulse = -1.0f; s0.state = 0; s0.offset = 0;      // pulse after the ctor's update_cycle
    R.hstate.assign(S, s0); R.hplans.resize(S);
    if(dev_upload(R.pstate.p, R.hstate.data(), (size_t)S * sizeof(RtPbpState), st) || dev_sync(st)) return LLSM_B200_ECUDA;
  }
  return 0;
}

// one feed step on device frame arrays (row length 1 per stream); outputs next_nhop samples per stream
struct RtL1Feed {                         // layer-1 inputs of a feed step (all device pointers)
  const llsm_b200_layer1* l1 = nullptr; const int* pbpsyn = nullptr; const L1PlanDev* lp = nullptr;
  const RtPbpPlan* plans = nullptr;       // host-made plans of this step, [S]
};

static inline int rt_feed(RtBatch& R, const llsm_b200_frames& fr, int row_stride, int row_off,
  float* out_p, float* out_ap, int out_stride, int out_off, cudaStream_t st, LaunchCounter* lc,
  const RtL1Feed* L = nullptr) {
  int prev = 0;
  R.update_cycle(&prev);
  const int H = R.curr_nhop;
  if(H < 1 || 2 * H > R.nfft || out_off + R.next_nhop > out_stride) return LLSM_B200_ERANGE;
  RtWindow* w = R.window(H, st);
  if(! w) return LLSM_B200_ENOMEM;
  RtFeedParams P; memset(&P, 0, sizeof(P));
  P.S = R.S; P.nchannel = R.nch; P.maxnhar = R.maxnhar; P.maxnhar_e = R.maxnhar_e; P.npsd = R.npsd; P.cap = R.cap;
  P.ntemplate = R.ntemplate; P.row_stride = row_stride; P.row_off = row_off;
  P.f0 = fr.f0; P.nhar = fr.nhar; P.ampl = fr.ampl; P.phse = fr.phse; P.psd = fr.psd; P.psdres = fr.psdres;
  P.edc = fr.edc; P.enhar = fr.enhar; P.eampl = fr.eampl; P.ephse = fr.ephse;
  P.mod = R.mod.as<float>(); P.sin_ = R.sinb.as<float>(); P.noise = R.noise.as<float>(); P.exc = R.exc.as<float>();
  P.tmpl = R.tmpl.as<float>(); P.prev_psd = R.prev_psd.as<float>(); P.has_prev = R.has_prev; P.chan_mask = R.chan_mask;
  P.H = H; P.next_nhop = R.next_nhop;
  P.cur_old = R.cur; P.cur_new = (R.cur + H) % R.cap;
  P.mod_old = R.mod_cur; P.mod_new = (R.mod_cur + H) % R.cap;
  P.exc_old = R.exc_cur; P.exc_new = (R.exc_cur + H) % R.cap;
  P.exc_cycle = R.exc_cycle; P.sin_pos = R.sin_pos; P.cycle = R.cycle; P.fs = R.fs;
  P.nfft = R.nfft; P.lg_nfft = R.lg_nfft; P.nspec = R.nspec; P.wsqr = w->wsqr; P.win = w->win;
  P.psd_lo = R.psd_lo; P.psd_r = R.psd_r; P.tw = R.tw;
  P.use_iczt = R.use_iczt; P.iczt_a = R.iczt_a; P.iczt_b = R.iczt_b;
  P.out_p = out_p; P.out_ap = out_ap; P.out_stride = out_stride; P.out_off = out_off;
  size_t smem = (size_t)R.nfft * 16 + ((size_t)R.nspec + R.npsd + 2 * H + 2 + 32 + 2 * R.maxnhar + 2 * R.nch * R.maxnhar_e) * 4 + 16;
  RtL1Params Q; memset(&Q, 0, sizeof(Q));
  if(R.use_l1) {
    if(L == nullptr || L->l1 == nullptr || L->lp == nullptr) return LLSM_B200_EINVAL;
    Q.rd = L->l1->rd; Q.vtmagn = L->l1->vtmagn; Q.nspec = L->l1->nspec; Q.vsphse = L->l1->vsphse; Q.nvs = L->l1->nvs;
    Q.vs_stride = R.maxnhar; Q.pbpsyn = L->pbpsyn;
    Q.pulse_f = R.pulse_f.as<float>(); Q.pulse_b = R.pulse_b.as<float>();
    Q.state = R.pstate.as<RtPbpState>(); Q.plans = L->plans; Q.prev_nhop = R.prev_nhop;
    Q.fnyq = (float)((double)R.fs / 2.0); Q.lip_radius = R.lip_radius;
    Q.max_size = R.max_pulse; Q.maxnhar_vs = R.maxnhar; Q.tw_p = L->lp->tw; Q.ntw_p = L->lp->ntw;
    size_t ps = (size_t)R.max_pulse * 16 + ((size_t)R.maxnhar * 5 + 8) * 4 + 16;
    if(ps > smem) smem = ps;
  }
  if(smem > 200 * 1024) return LLSM_B200_ERANGE;
  if(R.use_l1) {
    auto kfn = rt_feed_kernel<true>;
#ifndef LLSM_EMU
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    LLSM_LAUNCH(kfn, dim3(R.S), dim3(RT_THREADS), smem, st, P, Q);
  } else {
    auto kfn = rt_feed_kernel<false>;
#ifndef LLSM_EMU
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    LLSM_LAUNCH(kfn, dim3(R.S), dim3(RT_THREADS), smem, st, P, Q);
  }
  if(lc) lc->n ++;
  R.cur = P.cur_new; R.mod_cur = P.mod_new; R.exc_cur = P.exc_new;
  R.exc_cycle = (R.exc_cycle + H) % R.ntemplate;
  R.has_prev = 1;
  return 0;
}

// samples the next `nfeed` feeds will produce in total, without advancing the clock
static inline int rt_peek_output(const RtBatch& R, int nfeed) {
  RtBatch c; c.cycle = R.cycle; c.curr_nhop = R.curr_nhop; c.thop = R.thop; c.fs = R.fs;
  int total = 0;
  for(int i = 0; i < nfeed; i ++) { c.update_cycle(nullptr); total += c.next_nhop; }
  return total;
}

// llsm_rtsynth_buffer_clear (llsmrt.c:578-602); prev_nm and the modulation rings are kept, as there
static inline int rt_clear(RtBatch& R, cudaStream_t st, LaunchCounter* lc) {
  R.curr_nhop = 1;
  R.update_cycle(nullptr);                      // with the stale cycle, as the reference does
  RtClearParams C; memset(&C, 0, sizeof(C));
  C.S = R.S; C.nchannel = R.nch; C.cap = R.cap; C.mod_old = R.mod_cur; C.H = R.curr_nhop;
  C.mod = R.mod.as<float>(); C.sin_ = R.sinb.as<float>(); C.noise = R.noise.as<float>(); C.exc = R.exc.as<float>();
  LLSM_LAUNCH(rt_clear_kernel, dim3((R.cap + 255) / 256, R.S), dim3(256), 0, st, C);
  if(lc) lc->n ++;
  R.mod_cur = (R.mod_cur + R.curr_nhop) % R.cap;
  R.cur = R.curr_nhop % R.cap;
  R.exc_cur = 0;
  R.cycle = 0; R.exc_cycle = 0;
  R.sin_pos = -R.curr_nhop * 2 - R.nfft / 2;
  if(R.use_l1) {
    const size_t ring = (size_t)R.S * R.cap * 4;
    RtPbpState s0; s0.pulse = 0.f; s0.state = 0; s0.offset = 0;
    R.hstate.assign(R.S, s0);
    if(dev_memset(R.pulse_f.p, 0, ring, st) || dev_memset(R.pulse_b.p, 0, ring, st) ||
       dev_upload(R.pstate.p, R.hstate.data(), (size_t)R.S * sizeof(RtPbpState), st) || dev_sync(st)) return LLSM_B200_ECUDA;
  }
  return 0;
}

// ---- block feed with layer-1 members ------------------------------------------------------------------
// Host view of the few scalars the pulse tracker needs, for streams whose frames carry effect callbacks.
struct RtHostTrack {
  const float* f0 = nullptr; const int* nvs = nullptr; const float* rd = nullptr;
  const float* vsphse = nullptr; int vs_stride = 0; const int* pbpsyn = nullptr;   // host arrays [S][nfeed]
};

// Feed nfeed frames per stream ([S][nfeed][..] device arrays). fr.ampl == NULL: the harmonic model of
// every frame is derived from its layer-1 members first (llsm_frame_tolayer0, llsmrt.c:346-347,388-389).
// ht != NULL: tracker on the host with `mod` called once per pulse in time order, plans uploaded per step.
template <class Mod>
static inline int rt_feed_block_l1(RtBatch& R, llsm_b200_frames fr, int nfeed, const llsm_b200_layer1& l1,
  const int* pbpsyn, const L1PlanDev& lp, const RtHostTrack* ht, Mod* mods,
  float* out_p, float* out_ap, int out_stride, int* nout, cudaStream_t st, LaunchCounter* lc) {
  const int S = R.S;
  if(fr.ampl == nullptr) {
    const size_t rows = (size_t)S * nfeed;
    if(R.hm_nhar.reserve(rows * 4) || R.hm_ampl.reserve(rows * R.maxnhar * 4) || R.hm_phse.reserve(rows * R.maxnhar * 4))
      return LLSM_B200_ENOMEM;
    llsm_b200_conf c; memset(&c, 0, sizeof(c));
    c.nutt = S; c.nfrm = nfeed; c.maxnhar = R.maxnhar; c.fs = R.fs; c.thop = R.thop; c.lip_radius = R.lip_radius;
    int rc = run_tolayer0(lp, c, nullptr, fr.f0, l1, R.hm_nhar.as<int>(), R.hm_ampl.as<float>(), R.hm_phse.as<float>(), st, lc);
    if(rc) return rc;
    fr.nhar = R.hm_nhar.as<int>(); fr.ampl = R.hm_ampl.as<float>(); fr.phse = R.hm_phse.as<float>();
  }
  RtL1Feed L; L.l1 = &l1; L.pbpsyn = pbpsyn; L.lp = &lp;
  int off = 0;
  for(int i = 0; i < nfeed; i ++) {
    if(ht != nullptr) {
      // peek this step's clock, run the trackers, ship the plans
      RtBatch c; c.cycle = R.cycle; c.curr_nhop = R.curr_nhop; c.thop = R.thop; c.fs = R.fs;
      int prev = 0; c.update_cycle(&prev);
      for(int s = 0; s < S; s ++) {
        const size_t r = (size_t)s * nfeed + i;
        rt_pbp_track(R.hstate[s], R.hplans[s], prev, c.curr_nhop, R.sin_pos, R.fs, l1.nspec, ht->f0[r], ht->nvs[r] > 0,
          ht->rd[r], ht->vsphse[r * ht->vs_stride], ht->pbpsyn ? (ht->pbpsyn[r] == 1) : 0, mods[s], i);
      }
      if(dev_sync(st) || dev_upload(R.plans.p, R.hplans.data(), (size_t)S * sizeof(RtPbpPlan), st)) return LLSM_B200_ECUDA;
      L.plans = R.plans.as<RtPbpPlan>();
    }
    int rc = rt_feed(R, fr, nfeed, i, out_p, out_ap, out_stride, off, st, lc, &L);
    if(rc) return rc;
    off += R.next_nhop;
  }
  if(ht != nullptr && dev_sync(st)) return LLSM_B200_ECUDA;      // hplans is reused by the next call
  if(nout) *nout = off;
  return 0;
}
