// C entry points of the CPU thread emulation of the kernels -- TEST INFRASTRUCTURE ONLY.
// Same launch orchestration (driver.h) and kernel sources as the CUDA library, compiled by g++
// with -DLLSM_EMU; "device" pointers are host pointers here. Loaded only by tests/.
#include "cuda_emu.h"
#include "../../libllsm2_b200/csrc/driver.h"

extern "C" {

int emu_synthesize_l0(const llsm_b200_conf* conf, const llsm_b200_frames* fr,
  const llsm_b200_soptions* opt, const llsm_b200_output* out) {
  SynthPlanDev pd;
  if(pd.build(conf->nfrm, conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq, nullptr) != 0)
    return -100;
  SynthScratch sc;
  DevBuf nyb;
  const int* ny_utt = nullptr;
  if(fr->nfrm_utt) {
    nyb.reserve(conf->nutt * sizeof(int));
    LLSM_LAUNCH(ny_utt_kernel, dim3((conf->nutt + 63) / 64), dim3(64), 0, nullptr,
      fr->nfrm_utt, conf->nutt, conf->thop, conf->fs, nyb.as<int>());
    ny_utt = nyb.as<int>();
  }
  int rc = run_synth_l0(pd, sc, *conf, *fr, *opt, *out, ny_utt, nullptr, nullptr);
  sc.colored.release(); sc.y_exc.release(); nyb.release();
  pd.release();
  return rc;
}

int emu_synthesize_harmonics(const llsm_b200_conf* conf, const llsm_b200_frames* fr,
  const llsm_b200_soptions* opt_or_null, float* y_sin, int nsamp, int stride) {
  SynthPlanDev pd;
  if(pd.build(conf->nfrm, conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq, nullptr) != 0)
    return -100;
  int rc = run_harmonics(pd, *conf, *fr, opt_or_null, nullptr, y_sin, nsamp, nsamp, stride, nullptr, nullptr);
  pd.release();
  return rc;
}

// plan introspection for the host-logic tests
int emu_plan_query(int nfrm, float fs, float thop, int npsd, int* out_ints, int* hm_base, int* env_off) {
  SynthPlan p;
  float cf[8] = {2000, 4000, 8000, 0, 0, 0, 0, 0};
  build_synth_plan(p, nfrm, fs, thop, npsd, 4, cf);
  out_ints[0] = p.ny; out_ints[1] = p.n_hm; out_ints[2] = p.n_env; out_ints[3] = p.n_ns;
  out_ints[4] = p.nfft_ns; out_ints[5] = p.ntemplate; out_ints[6] = p.nt;
  if(hm_base) memcpy(hm_base, p.hm_base.data(), nfrm * sizeof(int));
  if(env_off) memcpy(env_off, p.env_off.data(), nfrm * sizeof(int));
  return 0;
}

}

#include "../../libllsm2_b200/csrc/driver_analysis.h"
extern "C" int emu_analyze_l0(const llsm_b200_conf* conf, const llsm_b200_aoptions* opt, const float* x,
  int nx, int xstride, const llsm_b200_frames_out* fr, float* x_res) {
  SynthPlanDev sp; AnaPlanDev ap; AnaScratch sc;
  if(sp.build(conf->nfrm, conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq, nullptr) != 0) return -100;
  if(ap.build(conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq, nullptr) != 0) return -101;
  int rc = run_analyze_l0(sp, ap, sc, *conf, *opt, x, nx, xstride, *fr, nullptr, x_res, nullptr, nullptr);
  sc.release(); ap.release(); sp.release();
  return rc;
}

#include "../../libllsm2_b200/csrc/driver_layer1.h"
extern "C" int emu_tolayer1(const llsm_b200_conf* conf, const llsm_b200_frames* fr, int nfft, const llsm_b200_layer1* out) {
  L1PlanDev lp; if(lp.build(nullptr) != 0) return -100;
  int rc = run_tolayer1(lp, *conf, *fr, nfft, *out, nullptr, nullptr);
  lp.release(); return rc;
}
extern "C" int emu_tolayer0(const llsm_b200_conf* conf, const float* f0, const llsm_b200_layer1* in, int* nhar,
  float* ampl, float* phse) {
  L1PlanDev lp; if(lp.build(nullptr) != 0) return -100;
  int rc = run_tolayer0(lp, *conf, nullptr, f0, *in, nhar, ampl, phse, nullptr, nullptr);
  lp.release(); return rc;
}

#include "../../libllsm2_b200/csrc/driver_pbp.h"
extern "C" int emu_synthesize_l1(const llsm_b200_conf* conf, const llsm_b200_frames* fr, const llsm_b200_layer1* l1,
  const int* pbpsyn, const llsm_b200_soptions* opt, const llsm_b200_output* out) {
  SynthPlanDev pd; L1PlanDev lp; SynthScratch sc; PbpScratch ps;
  if(pd.build(conf->nfrm, conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq, nullptr) != 0) return -100;
  if(lp.build(nullptr) != 0) return -101;
  DevBuf nyb;
  const int* ny_utt = nullptr;
  if(fr->nfrm_utt) {                                   // ragged batch: per-utterance output lengths, as the library does
    nyb.reserve(conf->nutt * sizeof(int));
    LLSM_LAUNCH(ny_utt_kernel, dim3((conf->nutt + 63) / 64), dim3(64), 0, nullptr,
      fr->nfrm_utt, conf->nutt, conf->thop, conf->fs, nyb.as<int>());
    ny_utt = nyb.as<int>();
  }
  int rc = run_synth_l1(pd, lp, sc, ps, *conf, *fr, *l1, pbpsyn, *opt, *out, ny_utt, nullptr, nullptr);
  sc.colored.release(); sc.y_exc.release(); ps.release(); lp.release(); pd.release(); nyb.release();
  return rc;
}

extern "C" int emu_synthesize_l0_shard(const llsm_b200_conf* conf, const llsm_b200_frames* fr,
  const llsm_b200_soptions* opt, const llsm_b200_output* out, int frame_lo, int frame_hi) {
  SynthPlanDev pd;
  if(pd.build(conf->nfrm, conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq, nullptr) != 0) return -100;
  SynthScratch sc;
  int rc = run_synth_l0(pd, sc, *conf, *fr, *opt, *out, nullptr, nullptr, nullptr, frame_lo, frame_hi);
  sc.colored.release(); sc.y_exc.release(); pd.release();
  return rc;
}

#include "../../libllsm2_b200/csrc/driver_rt.h"
// whole-utterance streaming run on the emulator: feed nfrm frames, collect the streamed samples
extern "C" int emu_rtsynth(const llsm_b200_conf* conf, const llsm_b200_frames* fr, const llsm_b200_soptions* opt,
  float* out_p, float* out_ap, int out_cap, int* nout, int* latency, int clear_at) {
  RtBatch R;
  int rc = rt_create(R, *conf, *opt, 0, nullptr, nullptr);
  if(rc != 0) return rc;
  *latency = -R.sin_pos - R.curr_nhop;
  int total = 0;
  for(int i = 0; i < conf->nfrm; i ++) {
    if(i == clear_at) rt_clear(R, nullptr, nullptr);
    rc = rt_feed(R, *fr, conf->nfrm, i, out_p, out_ap, out_cap, total, nullptr, nullptr);
    if(rc != 0) { R.release(); return rc; }
    total += R.next_nhop;
  }
  *nout = total;
  R.release();
  return 0;
}

// layer-1 streaming: frames + L1 members ([S][F][..] host arrays = "device" arrays of the emulator)
extern "C" int emu_rtsynth_l1(const llsm_b200_conf* conf, const llsm_b200_frames* fr, const llsm_b200_layer1* l1,
  const int* pbpsyn, const llsm_b200_soptions* opt, float* out_p, float* out_ap, int out_cap, int* nout, int* latency,
  int host_tracker, int block) {
  RtBatch R;
  int rc = rt_create(R, *conf, *opt, 1, nullptr, nullptr);
  if(rc != 0) return rc;
  R.host_tracker = host_tracker;
  *latency = -R.sin_pos - R.curr_nhop;
  L1PlanDev lp;
  if(lp.build(nullptr) != 0) return -1;
  const int S = conf->nutt, F = conf->nfrm;
  int total = 0;
  std::vector<float> bp((size_t)S * 8192), bap((size_t)S * 8192);
  for(int i0 = 0; i0 < F; i0 += block) {
    const int k = F - i0 < block ? F - i0 : block;
    // slice [S][k] rows out of the [S][F] arrays
    auto cut = [&](const void* src, size_t rowbytes, std::vector<char>& dst) -> void* {
      if(src == nullptr) return nullptr;
      dst.resize((size_t)S * k * rowbytes);
      for(int s = 0; s < S; s ++)
        memcpy(dst.data() + (size_t)s * k * rowbytes, (const char*)src + ((size_t)s * F + i0) * rowbytes, (size_t)k * rowbytes);
      return dst.data();
    };
    std::vector<char> b[16];
    const size_t nch = conf->nchannel;
    llsm_b200_frames f; memset(&f, 0, sizeof(f));
    f.f0 = (const float*)cut(fr->f0, 4, b[0]); f.nhar = (const int*)cut(fr->nhar, 4, b[1]);
    f.ampl = (const float*)cut(fr->ampl, 4 * conf->maxnhar, b[2]); f.phse = (const float*)cut(fr->phse, 4 * conf->maxnhar, b[3]);
    f.psd = (const float*)cut(fr->psd, 4 * conf->npsd, b[4]); f.psdres = (const float*)cut(fr->psdres, 4 * conf->npsd, b[5]);
    f.edc = (const float*)cut(fr->edc, 4 * nch, b[6]); f.enhar = (const int*)cut(fr->enhar, 4 * nch, b[7]);
    f.eampl = (const float*)cut(fr->eampl, 4 * nch * conf->maxnhar_e, b[8]); f.ephse = (const float*)cut(fr->ephse, 4 * nch * conf->maxnhar_e, b[9]);
    llsm_b200_layer1 l = *l1;
    l.rd = (float*)cut(l1->rd, 4, b[10]); l.vtmagn = (float*)cut(l1->vtmagn, 4 * (size_t)l1->nspec, b[11]);
    l.vsphse = (float*)cut(l1->vsphse, 4 * conf->maxnhar, b[12]); l.nvs = (int*)cut(l1->nvs, 4, b[13]);
    const int* pb = (const int*)cut(pbpsyn, 4, b[14]);
    RtHostTrack ht; ht.f0 = f.f0; ht.nvs = l.nvs; ht.rd = l.rd; ht.vsphse = l.vsphse; ht.vs_stride = conf->maxnhar; ht.pbpsyn = pb;
    std::vector<PbpNoEffect> mods(S);
    int n = 0;
    rc = rt_feed_block_l1(R, f, k, l, pb, lp, host_tracker ? &ht : nullptr, mods.data(), bp.data(), bap.data(), 8192, &n,
      nullptr, nullptr);
    if(rc != 0) { R.release(); lp.release(); return rc; }
    for(int s = 0; s < S; s ++)
      for(int q = 0; q < n && total + q < out_cap; q ++) {
        out_p[(size_t)s * out_cap + total + q] = bp[(size_t)s * 8192 + q];
        out_ap[(size_t)s * out_cap + total + q] = bap[(size_t)s * 8192 + q];
      }
    total += n;
  }
  *nout = total;
  R.release(); lp.release();
  return 0;
}

#include "../../libllsm2_b200/csrc/kernels_phase.cuh"
extern "C" int emu_phase_op(const llsm_b200_conf* conf, const int* nfrm_utt, const llsm_b200_frames_out* fr,
  const llsm_b200_layer1* l1, int mode, int arg) {
  std::vector<float> theta((size_t)conf->nutt * conf->nfrm);
  PhaseParams P; memset(&P, 0, sizeof(P));
  P.nutt = conf->nutt; P.nfrm = conf->nfrm; P.maxnhar = conf->maxnhar; P.maxnhar_e = conf->maxnhar_e; P.nchannel = conf->nchannel;
  P.nfrm_utt = nfrm_utt; P.f0 = fr->f0; P.nhar = fr->nhar; P.phse = fr->phse; P.enhar = fr->enhar; P.ephse = fr->ephse;
  if(l1) { P.vsphse = l1->vsphse; P.nvs = l1->nvs; }
  P.thop = conf->thop; P.mode = mode; P.arg = arg; P.theta = theta.data();
  return run_phase_op(P, nullptr, nullptr);
}

#include "../../libllsm2_b200/csrc/driver_coder.h"
extern "C" int emu_coder_encode(const llsm_b200_conf* conf, const float* f0, const float* psd, const llsm_b200_layer1* l1,
  int order_spec, int order_bap, float* enc) {
  CoderPlanDev cp; if(cp.build(conf->fs, conf->npsd, l1->nspec, order_bap, nullptr) != 0) return -100;
  int rc = run_coder_encode(cp, *conf, nullptr, f0, psd, l1->rd, l1->vtmagn, order_spec, enc, nullptr, nullptr);
  cp.release(); return rc;
}
extern "C" int emu_coder_decode(const llsm_b200_conf* conf, const float* enc, int order_spec, int order_bap, int use_layer1,
  const llsm_b200_frames_out* out, const llsm_b200_layer1* l1) {
  CoderPlanDev cp; if(cp.build(conf->fs, conf->npsd, l1->nspec, order_bap, nullptr) != 0) return -100;
  L1PlanDev lp; if(lp.build(nullptr) != 0) return -101;
  int rc = run_coder_decode(cp, lp, *conf, nullptr, enc, order_spec, use_layer1, out->f0, l1->rd, out->psd, out->nhar,
    out->ampl, out->phse, l1->vtmagn, l1->vsphse, nullptr, nullptr);
  cp.release(); lp.release(); return rc;
}

#include "../../libllsm2_b200/csrc/kernels_stretch.cuh"
extern "C" int emu_frames_stretch(const llsm_b200_conf* conf, const llsm_b200_frames* src, const llsm_b200_layer1* sl,
  int nfrm_new, const int* base, const float* ratio, const int* residx, int map_per_utt,
  const llsm_b200_frames_out* dst, const llsm_b200_layer1* dl) {
  StretchParams P; memset(&P, 0, sizeof(P));
  P.nutt = conf->nutt; P.nfrm = conf->nfrm; P.nfrm_new = nfrm_new; P.maxnhar = conf->maxnhar; P.maxnhar_e = conf->maxnhar_e;
  P.npsd = conf->npsd; P.nchannel = conf->nchannel; P.nspec = sl->nspec; P.map_per_utt = map_per_utt;
  P.base = base; P.ratio = ratio; P.residx = residx;
  P.f0 = src->f0; P.rd = sl->rd; P.vtmagn = sl->vtmagn; P.vsphse = sl->vsphse; P.nvs = sl->nvs;
  P.psd = src->psd; P.psdres = src->psdres; P.edc = src->edc; P.enhar = src->enhar; P.eampl = src->eampl; P.ephse = src->ephse;
  P.nhar = src->nhar; P.ampl = src->ampl; P.phse = src->phse;
  P.o_f0 = dst->f0; P.o_rd = dl->rd; P.o_vtmagn = dl->vtmagn; P.o_vsphse = dl->vsphse; P.o_nvs = dl->nvs;
  P.o_psd = dst->psd; P.o_psdres = dst->psdres; P.o_edc = dst->edc; P.o_enhar = dst->enhar; P.o_eampl = dst->eampl;
  P.o_ephse = dst->ephse; P.o_nhar = dst->nhar; P.o_ampl = dst->ampl; P.o_phse = dst->phse;
  return run_frames_stretch(P, nullptr, nullptr);
}

// ---- halo exchange of frame-range shards: the two kernels around the all-gather (kernels_halo.cuh) ----
#include "../../libllsm2_b200/csrc/kernels_halo.cuh"
// spos: [world + 1] owned-range boundaries; strips: this rank's [B][2][2][halo]
extern "C" int emu_halo_pack(int nutt, int ny, int stride, int halo, int rank, int world, const int* spos,
  float* y_sin, float* y_noise, float* strips) {
  HaloParams P; memset(&P, 0, sizeof(P));
  P.nutt = nutt; P.ny = ny; P.stride = stride; P.halo = halo; P.rank = rank; P.world = world; P.spos = spos;
  P.y_sin = y_sin; P.y_noise = y_noise; P.strips = strips;
  run_halo_pack(P, nullptr);
  return 0;
}
// gathered: [world][B][2][2][halo]
extern "C" int emu_halo_add(int nutt, int ny, int stride, int halo, int rank, int world, const int* spos,
  float* y_sin, float* y_noise, float* y, float* gathered) {
  HaloParams P; memset(&P, 0, sizeof(P));
  P.nutt = nutt; P.ny = ny; P.stride = stride; P.halo = halo; P.rank = rank; P.world = world; P.spos = spos;
  P.y_sin = y_sin; P.y_noise = y_noise; P.y = y; P.strips = gathered;
  run_halo_add(P, spos[rank + 1] - spos[rank], nullptr);
  return 0;
}

// Zero-phase sub-band filter, two implementations side by side (test_emu_iir.py): the streaming kernel
// (kernels_iir.cuh) and the shared-memory resident one (kernels_iir_smem.cuh, single-CTA instance) on the same input.
// src: [nutt][sstride] (one row per utterance, read by every channel's first section), y_*: [nutt * nch][ystride].
extern "C" int emu_iir_both(int nutt, int nch, int n, float fs, const float* chanfreq, const float* src, int sstride,
  int square, float* y_stream, float* y_smem, int ystride) {
  AnaPlan h;
  build_ana_plan(h, fs, 0.005f, 64, nch, chanfreq);
  const int L = ((n + IIR_NT - 1) / IIR_NT + IIR_T - 1) & ~(IIR_T - 1);
  std::vector<double> coef((size_t)LLSM_B200_MAXCHANNEL * 2 * 9, 0.0), mpow((size_t)LLSM_B200_MAXCHANNEL * 2 * IIR_NLOG * 16, 0.0);
  int cs = 0, Ls = 0;
  iir_smem_geometry(n, cs, Ls);
  if(cs != 1) return -2;
  std::vector<double> coef2(9), mpow9((size_t)LLSM_B200_MAXCHANNEL * 2 * IIS_NLOG * 16, 0.0), wts((size_t)LLSM_B200_MAXCHANNEL * 2 * Ls * 4, 0.0);
  for(int c = 0; c < nch; c ++)
    for(int s2 = 0; s2 < h.chan[c].nstage; s2 ++) {
      build_iir_section(h.chan[c].b[s2], h.chan[c].a[s2], L, IIR_NLOG, &coef[((size_t)c * 2 + s2) * 9], &mpow[((size_t)c * 2 + s2) * IIR_NLOG * 16]);
      build_iir_smem_section(h.chan[c].b[s2], h.chan[c].a[s2], Ls, coef2.data(), &mpow9[((size_t)c * 2 + s2) * IIS_NLOG * 16],
        &wts[((size_t)c * 2 + s2) * Ls * 4]);
    }
  IirParams I; memset(&I, 0, sizeof(I));
  I.nchannel = nch; I.n = n; I.L = L; I.ystride = ystride; I.vec_ok = (ystride & 3) == 0;
  I.src_a = src; I.sa_stride = sstride; I.src_per_utt = 1; I.coef = coef.data(); I.mpow = mpow.data();
  for(int c = 0; c < nch; c ++) I.nstage[c] = h.chan[c].nstage;
  I.square = square;
  I.y = y_stream;
  LLSM_LAUNCH(iir_filtfilt_kernel, dim3(nutt * nch), dim3(IIR_NT), 0, nullptr, I);
  IirSmemParams Q; Q.base = I; Q.base.y = y_smem; Q.mpow = mpow9.data(); Q.wts = wts.data(); Q.L = Ls;
  return launch_iir_smem(Q, nutt * nch, cs, nullptr);
}
