// C ABI of libllsm2_b200.so (include/llsm_b200.h): context, plan cache, host/device entry points.
// CUDA only -- there is no CPU path; every compute entry fails with LLSM_B200_ENODEVICE when no
// device is present.
#include "driver.h"
#include "driver_analysis.h"
#include "driver_layer1.h"
#include "driver_pbp.h"
#include "driver_rt.h"
#include "kernels_phase.cuh"
#include "kernels_stretch.cuh"
#include "driver_coder.h"
#include "kernels_halo.cuh"
#include <dlfcn.h>
#include <cstdio>
#include <cstdarg>
#include <mutex>
#include <memory>

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct NcclUniqueId { char internal[128]; };   // ncclUniqueId (nccl.h): passed by value to ncclCommInitRank

struct PlanKey {
  int nfrm, npsd, nchannel; float fs, thop; float cf[LLSM_B200_MAXCHANNEL];
  bool operator<(const PlanKey& o) const { return memcmp(this, &o, sizeof(PlanKey)) < 0; }
};

struct llsm_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::map<PlanKey, std::unique_ptr<SynthPlanDev>> plans;
  std::map<AnaKey, std::unique_ptr<AnaPlanDev>> aplans;
  SynthScratch scratch;
  AnaScratch ascratch;
  AnaFork afork;
  std::unique_ptr<L1PlanDev> l1plan;
  std::unique_ptr<CoderPlanDev> coderplan;
  PbpScratch pbp;
  DevBuf ny_utt, phase_theta;
  // frame-range sharding: NCCL communicator (ours or the caller's) and the strip buffers of the halo exchange
  void* comm = nullptr; bool own_comm = false; int comm_rank = 0, comm_world = 1;
  DevBuf halo_send, halo_recv, halo_pos;
  DevBuf stage[24];          // device staging for the *_host entry points
  // copy / compute pipeline of synthesize_l0_host: two slots of input and output staging
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  DevBuf pin[2][12], pout[2][12];
  LaunchCounter lc;
  cudaEvent_t kt_ev[LLSM_KT_MAX] = {};
  // recursive: the *_host entry points hold it across their staging + the device entry they call, and a user
  // llsm_pbpeffect callback running under it may call back into llsm_* functions of the same context
  std::recursive_mutex mtx;
};
typedef std::lock_guard<std::recursive_mutex> CtxLock;

static PlanKey make_key(const llsm_b200_conf* c) {
  PlanKey k; memset(&k, 0, sizeof(k));
  k.nfrm = c->nfrm; k.npsd = c->npsd; k.nchannel = c->nchannel; k.fs = c->fs; k.thop = c->thop;
  for(int i = 0; i + 1 < c->nchannel && i < LLSM_B200_MAXCHANNEL; i ++) k.cf[i] = c->chanfreq[i];
  return k;
}

static int check_conf(const llsm_b200_conf* c) {
  if(c == nullptr) return fail(LLSM_B200_EINVAL, "conf is NULL");
  if(c->nutt < 1 || c->nfrm < 1 || c->maxnhar < 1 || c->maxnhar_e < 0 || c->npsd < 2)
    return fail(LLSM_B200_EINVAL, "bad sizes: nutt %d nfrm %d maxnhar %d maxnhar_e %d npsd %d",
      c->nutt, c->nfrm, c->maxnhar, c->maxnhar_e, c->npsd);
  if(c->nchannel < 1 || c->nchannel > LLSM_B200_MAXCHANNEL)
    return fail(LLSM_B200_EINVAL, "nchannel %d outside [1, %d]", c->nchannel, LLSM_B200_MAXCHANNEL);
  if(! (c->fs > 0) || ! (c->thop > 0)) return fail(LLSM_B200_EINVAL, "fs / thop must be positive");
  return 0;
}

static SynthPlanDev* get_plan(llsm_b200_ctx* ctx, const llsm_b200_conf* conf) {
  PlanKey k = make_key(conf);
  auto it = ctx->plans.find(k);
  if(it != ctx->plans.end()) return it->second.get();
  // one plan per distinct (nfrm, configuration): a service that synthesises variable-length utterances would grow the
  // cache without bound, so it is emptied (after the stream has drained: plans may be in use) when it reaches the cap
  if(ctx->plans.size() >= 64) {
    cudaStreamSynchronize(ctx->stream);
    for(auto& kv : ctx->plans) kv.second->release();
    ctx->plans.clear();
  }
  std::unique_ptr<SynthPlanDev> p(new SynthPlanDev());
  if(p->build(conf->nfrm, conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq,
       ctx->stream) != 0) { p->release(); return nullptr; }
  SynthPlanDev* raw = p.get();
  ctx->plans[k] = std::move(p);
  return raw;
}

static int cuda_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if(e != cudaSuccess) return fail(LLSM_B200_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

extern "C" {

const char* llsm_b200_last_error(void) { return g_err; }

llsm_b200_ctx* llsm_b200_create(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess || n <= 0) {
    fail(LLSM_B200_ENODEVICE, "no CUDA device (%s); libllsm2_b200 has no CPU path",
      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    cudaGetLastError();
    return nullptr;
  }
  if(device < 0 || device >= n) { fail(LLSM_B200_EINVAL, "device %d of %d", device, n); return nullptr; }
  if(cudaSetDevice(device) != cudaSuccess) { cuda_ok("cudaSetDevice"); return nullptr; }
  llsm_b200_ctx* ctx = new llsm_b200_ctx();
  ctx->device = device;
  if(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cuda_ok("cudaStreamCreate"); delete ctx; return nullptr;
  }
  ctx->own_stream = true;
  return ctx;
}

int llsm_b200_comm_destroy(llsm_b200_ctx* ctx);

void llsm_b200_destroy(llsm_b200_ctx* ctx) {
  if(ctx == nullptr) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  llsm_b200_comm_destroy(ctx);
  ctx->halo_send.release(); ctx->halo_recv.release(); ctx->halo_pos.release();
  for(auto& kv : ctx->plans) kv.second->release();
  for(auto& kv : ctx->aplans) kv.second->release();
  ctx->scratch.colored.release(); ctx->scratch.y_exc.release(); ctx->scratch.ny_utt.release();
  ctx->ascratch.release();
  if(ctx->l1plan) ctx->l1plan->release();
  ctx->pbp.release();
  ctx->ny_utt.release();
  for(auto& s : ctx->stage) s.release();
  for(int k = 0; k < 2; k ++) {
    for(auto& b : ctx->pin[k]) b.release();
    for(auto& b : ctx->pout[k]) b.release();
    if(ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
    if(ctx->ev_comp[k]) cudaEventDestroy(ctx->ev_comp[k]);
    if(ctx->ev_out[k]) cudaEventDestroy(ctx->ev_out[k]);
  }
  for(auto& e : ctx->kt_ev) if(e) cudaEventDestroy(e);
  ctx->phase_theta.release();
  if(ctx->coderplan) ctx->coderplan->release();
  if(ctx->afork.st2) cudaStreamDestroy(ctx->afork.st2);
  if(ctx->afork.fork) cudaEventDestroy(ctx->afork.fork);
  if(ctx->afork.join) cudaEventDestroy(ctx->afork.join);
  if(ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if(ctx->s_out) cudaStreamDestroy(ctx->s_out);
  if(ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int llsm_b200_set_stream(llsm_b200_ctx* ctx, void* cuda_stream) {
  if(ctx == nullptr) return fail(LLSM_B200_EINVAL, "ctx is NULL");
  if(ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return 0;
}

int llsm_b200_synchronize(llsm_b200_ctx* ctx) {
  if(ctx == nullptr) return fail(LLSM_B200_EINVAL, "ctx is NULL");
  if(cudaStreamSynchronize(ctx->stream) != cudaSuccess) return cuda_ok("cudaStreamSynchronize");
  return 0;
}

long long llsm_b200_launch_count(const llsm_b200_ctx* ctx) { return ctx ? ctx->lc.n : 0; }

int llsm_b200_output_length(int nfrm, float thop, float fs) { return plan_output_length(nfrm, thop, fs); }
int llsm_b200_template_length(int ny) { return plan_template_length(ny); }

static const int* ragged_lengths(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt_dev) {
  if(nfrm_utt_dev == nullptr) return nullptr;
  if(ctx->ny_utt.reserve(conf->nutt * sizeof(int)) != 0) return nullptr;
  ny_utt_kernel<<<(conf->nutt + 127) / 128, 128, 0, ctx->stream>>>(nfrm_utt_dev, conf->nutt,
    conf->thop, conf->fs, ctx->ny_utt.as<int>());
  ctx->lc.n += 1;
  return ctx->ny_utt.as<int>();
}

// body shared by the device entry and the host pipeline; the caller holds ctx->mtx
static int synth_l0_impl(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_frames* fr,
  const llsm_b200_soptions* opt, const llsm_b200_output* out, int utt_base) {
  SynthPlanDev* pd = get_plan(ctx, conf);
  if(pd == nullptr) return fail(LLSM_B200_ENOMEM, "could not build the synthesis plan");
  if(out->stride < pd->h.ny) return fail(LLSM_B200_EINVAL, "stride %d < ny %d", out->stride, pd->h.ny);
  const int* ny_utt = ragged_lengths(ctx, conf, fr->nfrm_utt);
  if(fr->nfrm_utt && ! ny_utt) return fail(LLSM_B200_ENOMEM, "ny_utt");
  int rc = run_synth_l0(*pd, ctx->scratch, *conf, *fr, *opt, *out, ny_utt, ctx->stream, &ctx->lc, 0, 0, utt_base);
  if(rc != 0) return fail(rc, "synthesis launch failed (code %d): size outside supported range?", rc);
  return cuda_ok("synthesize_l0");
}

int llsm_b200_synthesize_l0(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* fr, const llsm_b200_soptions* opt, const llsm_b200_output* out) {
  if(ctx == nullptr) return fail(LLSM_B200_ENODEVICE, "no context (no CUDA device?)");
  int rc = check_conf(conf); if(rc) return rc;
  if(! fr || ! opt || ! out) return fail(LLSM_B200_EINVAL, "NULL argument");
  if(! fr->f0 || ! fr->nhar || ! fr->ampl || ! fr->phse || ! fr->psd || ! fr->edc || ! fr->enhar ||
     ! fr->eampl || ! fr->ephse) return fail(LLSM_B200_EINVAL, "a required frame array is NULL");
  if(! out->y_sin || ! out->y_noise) return fail(LLSM_B200_EINVAL, "y_sin and y_noise are required");
  CtxLock lk(ctx->mtx);
  cudaSetDevice(ctx->device);
  return synth_l0_impl(ctx, conf, fr, opt, out, 0);
}

int llsm_b200_synthesize_l0_shard(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* fr, const llsm_b200_soptions* opt, const llsm_b200_output* out,
  int frame_lo, int frame_hi) {
  if(ctx == nullptr) return fail(LLSM_B200_ENODEVICE, "no context (no CUDA device?)");
  int rc = check_conf(conf); if(rc) return rc;
  if(! fr || ! opt || ! out) return fail(LLSM_B200_EINVAL, "NULL argument");
  if(! fr->f0 || ! fr->nhar || ! fr->ampl || ! fr->phse || ! fr->psd || ! fr->edc || ! fr->enhar ||
     ! fr->eampl || ! fr->ephse) return fail(LLSM_B200_EINVAL, "a required frame array is NULL");
  if(! out->y_sin || ! out->y_noise) return fail(LLSM_B200_EINVAL, "y_sin and y_noise are required");
  CtxLock lk(ctx->mtx);
  cudaSetDevice(ctx->device);
  SynthPlanDev* pd = get_plan(ctx, conf);
  if(pd == nullptr) return fail(LLSM_B200_ENOMEM, "could not build the synthesis plan");
  if(out->stride < pd->h.ny) return fail(LLSM_B200_EINVAL, "stride %d < ny %d", out->stride, pd->h.ny);
  const int* ny_utt = ragged_lengths(ctx, conf, fr->nfrm_utt);
  if(fr->nfrm_utt && ! ny_utt) return fail(LLSM_B200_ENOMEM, "ny_utt");
  if(frame_lo < 0 || frame_hi > conf->nfrm || frame_lo >= frame_hi) return fail(LLSM_B200_EINVAL, "bad frame range [%d, %d)", frame_lo, frame_hi);
  rc = run_synth_l0(*pd, ctx->scratch, *conf, *fr, *opt, *out, ny_utt, ctx->stream, &ctx->lc, frame_lo, frame_hi);
  if(rc != 0) return fail(rc, "synthesis launch failed (code %d): size outside supported range?", rc);
  return cuda_ok("synthesize_l0");
}

int llsm_b200_halo_length(const llsm_b200_conf* conf) {
  if(conf == nullptr) return -1;
  SynthPlan p; build_synth_plan(p, 2, conf->fs, conf->thop, conf->npsd, conf->nchannel, conf->chanfreq);
  int h = p.n_hm / 2 + 1;
  if(p.nfft_ns / 2 + 1 > h) h = p.nfft_ns / 2 + 1;
  return h;
}

int llsm_b200_frame_position(int i, float thop, float fs) {
  float r = (float)i * thop; r = r * fs;
  return (int)round((double)r);
}

int llsm_b200_synthesize_harmonics(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* fr, const llsm_b200_soptions* opt, float* y_sin, int nsamp, int stride) {
  if(ctx == nullptr) return fail(LLSM_B200_ENODEVICE, "no context (no CUDA device?)");
  int rc = check_conf(conf); if(rc) return rc;
  if(! fr || ! fr->f0 || ! fr->nhar || ! fr->ampl || ! fr->phse || ! y_sin)
    return fail(LLSM_B200_EINVAL, "NULL argument");
  if(stride < nsamp) return fail(LLSM_B200_EINVAL, "stride < nsamp");
  CtxLock lk(ctx->mtx);
  cudaSetDevice(ctx->device);
  SynthPlanDev* pd = get_plan(ctx, conf);
  if(pd == nullptr) return fail(LLSM_B200_ENOMEM, "could not build the synthesis plan");
  const int* ny_utt = nullptr;   // ragged rows: every sample below nsamp is valid here
  rc = run_harmonics(*pd, *conf, *fr, opt, ny_utt, y_sin, nsamp, nsamp, stride, ctx->stream, &ctx->lc);
  if(rc != 0) return fail(rc, "harmonic bank launch failed (window too long?)");
  return cuda_ok("synthesize_harmonics");
}

// ---- host-buffer variants: copy in, run, copy out, synchronise ----
struct Up { DevBuf* buf; const void* src; size_t bytes; };

static int upload_all(llsm_b200_ctx* ctx, std::vector<Up>& ups) {
  for(auto& u : ups) {
    if(u.src == nullptr) continue;
    if(u.buf->reserve(u.bytes) != 0) return fail(LLSM_B200_ENOMEM, "device staging (%zu bytes)", u.bytes);
    if(cudaMemcpyAsync(u.buf->p, u.src, u.bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      return cuda_ok("H2D copy");
  }
  return 0;
}

// Host buffers in, host buffers out. The batch is cut into slices of utterances that flow through a
// three-stage pipeline -- H2D on a copy stream, the kernels on the context's stream, D2H on a second copy
// stream -- with two staging slots, so that PCIe traffic in both directions overlaps the kernels (the
// step is PCIe-bound: ~3 KB in and ~5 KB out per frame). Pinned host memory is needed for real overlap;
// pageable memory works, serialised by the driver. Results do not depend on the slicing.
static int pipeline_ready(llsm_b200_ctx* ctx) {
  if(ctx->s_in) return 0;
  if(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) != cudaSuccess ||
     cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) != cudaSuccess) return cuda_ok("pipeline streams");
  for(int k = 0; k < 2; k ++)
    if(cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming) != cudaSuccess ||
       cudaEventCreateWithFlags(&ctx->ev_comp[k], cudaEventDisableTiming) != cudaSuccess ||
       cudaEventCreateWithFlags(&ctx->ev_out[k], cudaEventDisableTiming) != cudaSuccess) return cuda_ok("pipeline events");
  return 0;
}

int llsm_b200_synthesize_l0_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* fr, const llsm_b200_soptions* opt, const llsm_b200_output* out) {
  if(ctx == nullptr) return fail(LLSM_B200_ENODEVICE, "no context (no CUDA device?)");
  int rc = check_conf(conf); if(rc) return rc;
  if(! fr || ! opt || ! out) return fail(LLSM_B200_EINVAL, "NULL argument");
  if(! fr->f0 || ! fr->nhar || ! fr->ampl || ! fr->phse || ! fr->psd || ! fr->edc || ! fr->enhar ||
     ! fr->eampl || ! fr->ephse) return fail(LLSM_B200_EINVAL, "a required frame array is NULL");
  CtxLock lk(ctx->mtx);
  cudaSetDevice(ctx->device);
  rc = pipeline_ready(ctx); if(rc) return rc;
  const int B = conf->nutt, F = conf->nfrm;
  const size_t nch = conf->nchannel;
  const int ny = plan_output_length(F, conf->thop, conf->fs);
  if(out->stride < ny) return fail(LLSM_B200_EINVAL, "stride %d < ny %d", out->stride, ny);
  const int nt = plan_template_length(ny);
  // slice size: at least ~8 MB of output per slice, at most 8 slices (measured on B200 / PCIe 5: 1 slice
  // 7.2 M frames/s, 4: 11.8 M, 8: 12.5 M, 16: 11.5 M on the 1024 x 400-frame workload)
  int nslice = 1;
  {
    const char* e = getenv("LLSM_B200_HOST_SLICES");
    size_t total = (size_t)B * out->stride * 4;
    nslice = e ? atoi(e) : (int)(total / (8u << 20));
    if(nslice > 8 && ! e) nslice = 8;
    if(nslice > 64) nslice = 64;
    if(nslice > B) nslice = B;
    if(nslice < 1) nslice = 1;
  }
  const int Bs = (B + nslice - 1) / nslice;
  // row sizes (bytes per utterance) of the eleven frame arrays and the template
  const size_t row[12] = {4, (size_t)F * 4, (size_t)F * 4, (size_t)F * conf->maxnhar * 4, (size_t)F * conf->maxnhar * 4,
    (size_t)F * conf->npsd * 4, (size_t)F * conf->npsd * 4, (size_t)F * nch * 4, (size_t)F * nch * 4,
    (size_t)F * nch * conf->maxnhar_e * 4, (size_t)F * nch * conf->maxnhar_e * 4, nch * nt * 4};
  const void* src[12] = {fr->nfrm_utt, fr->f0, fr->nhar, fr->ampl, fr->phse, fr->psd, fr->psdres, fr->edc, fr->enhar,
    fr->eampl, fr->ephse, opt->white};
  float* dsth[3] = {out->y, out->y_sin, out->y_noise};
  const size_t orow = (size_t)out->stride * 4;
  for(int k = 0; k < 2 && k < nslice; k ++) {
    for(int a = 0; a < 12; a ++)
      if(src[a] && ctx->pin[k][a].reserve(row[a] * Bs) != 0) return fail(LLSM_B200_ENOMEM, "device input staging");
    for(int a = 0; a < 3; a ++)
      if(ctx->pout[k][a].reserve(orow * Bs) != 0) return fail(LLSM_B200_ENOMEM, "device output staging");
  }
  for(int c = 0, b0 = 0; b0 < B; c ++, b0 += Bs) {
    const int k = c & 1, bc = (B - b0 < Bs) ? B - b0 : Bs;
    DevBuf* in = ctx->pin[k]; DevBuf* ob = ctx->pout[k];
    if(c >= 2) cudaStreamWaitEvent(ctx->s_in, ctx->ev_comp[k], 0);          // slot inputs consumed
    for(int a = 0; a < 12; a ++)
      if(src[a] && cudaMemcpyAsync(in[a].p, (const char*)src[a] + row[a] * b0, row[a] * bc, cudaMemcpyHostToDevice,
           ctx->s_in) != cudaSuccess) return cuda_ok("H2D copy");
    cudaEventRecord(ctx->ev_in[k], ctx->s_in);
    cudaStreamWaitEvent(ctx->stream, ctx->ev_in[k], 0);
    if(c >= 2) cudaStreamWaitEvent(ctx->stream, ctx->ev_out[k], 0);         // slot outputs drained
    llsm_b200_conf cs = *conf; cs.nutt = bc;
    llsm_b200_frames d;
    d.nfrm_utt = fr->nfrm_utt ? in[0].as<int>() : nullptr;
    d.f0 = in[1].as<float>(); d.nhar = in[2].as<int>(); d.ampl = in[3].as<float>(); d.phse = in[4].as<float>();
    d.psd = in[5].as<float>(); d.psdres = fr->psdres ? in[6].as<float>() : nullptr;
    d.edc = in[7].as<float>(); d.enhar = in[8].as<int>(); d.eampl = in[9].as<float>(); d.ephse = in[10].as<float>();
    llsm_b200_soptions o = *opt;
    o.white = opt->white ? in[11].as<float>() : nullptr;
    llsm_b200_output od;
    od.y = ob[0].as<float>(); od.y_sin = ob[1].as<float>(); od.y_noise = ob[2].as<float>(); od.stride = out->stride;
    rc = synth_l0_impl(ctx, &cs, &d, &o, &od, b0);
    if(rc) { cudaDeviceSynchronize(); return rc; }
    cudaEventRecord(ctx->ev_comp[k], ctx->stream);
    cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[k], 0);
    for(int a = 0; a < 3; a ++)
      if(dsth[a] && cudaMemcpyAsync((char*)dsth[a] + orow * b0, ob[a].p, orow * bc, cudaMemcpyDeviceToHost,
           ctx->s_out) != cudaSuccess) return cuda_ok("D2H copy");
    cudaEventRecord(ctx->ev_out[k], ctx->s_out);
  }
  if(cudaStreamSynchronize(ctx->s_out) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
     cudaStreamSynchronize(ctx->s_in) != cudaSuccess) return cuda_ok("synchronize");
  return cuda_ok("synthesize_l0_host");
}

// Per-kernel timing (bench.py): while enabled, the analysis and synthesis pipelines record a named CUDA event after
// each of their kernels; every call with enable != 0 restarts the list. llsm_b200_kernel_timing_read waits for the last
// event and returns the duration (ms) and name of every kernel interval recorded since.
int llsm_b200_set_kernel_timing(llsm_b200_ctx* ctx, int enable) {
  if(ctx == nullptr) return fail(LLSM_B200_ENODEVICE, "no context (no CUDA device?)");
  CtxLock lk(ctx->mtx);
  cudaSetDevice(ctx->device);
  if(enable && ! ctx->kt_ev[0])
    for(int i = 0; i < LLSM_KT_MAX; i ++)
      if(cudaEventCreate(&ctx->kt_ev[i]) != cudaSuccess) return cuda_ok("kernel timing events");
  ctx->lc.ev = enable ? ctx->kt_ev : nullptr;
  ctx->lc.mark = 0;
  return 0;
}
int llsm_b200_kernel_timing_read(llsm_b200_ctx* ctx, int max, const char** names, float* ms) {
  if(ctx == nullptr || names == nullptr || ms == nullptr) { fail(LLSM_B200_EINVAL, "NULL argument"); return -1; }
  CtxLock lk(ctx->mtx);
  cudaSetDevice(ctx->device);
  if(! ctx->lc.ev || ctx->lc.mark < 2) { fail(LLSM_B200_EINVAL, "no timed step recorded"); return -1; }
  if(cudaEventSynchronize(ctx->kt_ev[ctx->lc.mark - 1]) != cudaSuccess) { cuda_ok("kernel timing"); return -1; }
  int n = 0;
  for(int i = 1; i < ctx->lc.mark && n < max; i ++) {
    if(strcmp(ctx->lc.names[i], "start") == 0) continue;
    if(cudaEventElapsedTime(&ms[n], ctx->kt_ev[i - 1], ctx->kt_ev[i]) != cudaSuccess) { cuda_ok("kernel timing"); return -1; }
    names[n ++] = ctx->lc.names[i];
  }
  return n;
}

#include "api_analysis.inc"
#include "api_layer1.inc"
#include "api_rt.inc"
#include "api_blob.inc"
#include "api_halo.inc"

} // extern "C"
