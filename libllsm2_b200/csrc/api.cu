// C ABI of libllsm2_b200.so (include/llsm_b200.h): context, plan cache, host/device entry points.
// CUDA only -- there is no CPU path; every compute entry fails with LLSM_B200_ENODEVICE when no
// device is present.
#include "driver.h"
#include "driver_analysis.h"
#include "driver_layer1.h"
#include "driver_pbp.h"
#include "driver_rt.h"
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
  std::unique_ptr<L1PlanDev> l1plan;
  PbpScratch pbp;
  DevBuf ny_utt;
  DevBuf stage[24];          // device staging for the *_host entry points
  LaunchCounter lc;
  std::mutex mtx;
};

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

void llsm_b200_destroy(llsm_b200_ctx* ctx) {
  if(ctx == nullptr) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for(auto& kv : ctx->plans) kv.second->release();
  for(auto& kv : ctx->aplans) kv.second->release();
  ctx->scratch.colored.release(); ctx->scratch.y_exc.release(); ctx->scratch.ny_utt.release();
  ctx->ascratch.release();
  if(ctx->l1plan) ctx->l1plan->release();
  ctx->pbp.release();
  ctx->ny_utt.release();
  for(auto& s : ctx->stage) s.release();
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

int llsm_b200_synthesize_l0(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* fr, const llsm_b200_soptions* opt, const llsm_b200_output* out) {
  if(ctx == nullptr) return fail(LLSM_B200_ENODEVICE, "no context (no CUDA device?)");
  int rc = check_conf(conf); if(rc) return rc;
  if(! fr || ! opt || ! out) return fail(LLSM_B200_EINVAL, "NULL argument");
  if(! fr->f0 || ! fr->nhar || ! fr->ampl || ! fr->phse || ! fr->psd || ! fr->edc || ! fr->enhar ||
     ! fr->eampl || ! fr->ephse) return fail(LLSM_B200_EINVAL, "a required frame array is NULL");
  if(! out->y_sin || ! out->y_noise) return fail(LLSM_B200_EINVAL, "y_sin and y_noise are required");
  std::lock_guard<std::mutex> lk(ctx->mtx);
  cudaSetDevice(ctx->device);
  SynthPlanDev* pd = get_plan(ctx, conf);
  if(pd == nullptr) return fail(LLSM_B200_ENOMEM, "could not build the synthesis plan");
  if(out->stride < pd->h.ny) return fail(LLSM_B200_EINVAL, "stride %d < ny %d", out->stride, pd->h.ny);
  const int* ny_utt = ragged_lengths(ctx, conf, fr->nfrm_utt);
  if(fr->nfrm_utt && ! ny_utt) return fail(LLSM_B200_ENOMEM, "ny_utt");
  rc = run_synth_l0(*pd, ctx->scratch, *conf, *fr, *opt, *out, ny_utt, ctx->stream, &ctx->lc);
  if(rc != 0) return fail(rc, "synthesis launch failed (code %d): size outside supported range?", rc);
  return cuda_ok("synthesize_l0");
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
  std::lock_guard<std::mutex> lk(ctx->mtx);
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
  std::lock_guard<std::mutex> lk(ctx->mtx);
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

int llsm_b200_synthesize_l0_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* fr, const llsm_b200_soptions* opt, const llsm_b200_output* out) {
  if(ctx == nullptr) return fail(LLSM_B200_ENODEVICE, "no context (no CUDA device?)");
  int rc = check_conf(conf); if(rc) return rc;
  if(! fr || ! opt || ! out) return fail(LLSM_B200_EINVAL, "NULL argument");
  cudaSetDevice(ctx->device);
  const size_t BF = (size_t)conf->nutt * conf->nfrm;
  const size_t nch = conf->nchannel;
  const int ny = plan_output_length(conf->nfrm, conf->thop, conf->fs);
  if(out->stride < ny) return fail(LLSM_B200_EINVAL, "stride %d < ny %d", out->stride, ny);
  const int nt = plan_template_length(ny);
  DevBuf* s = ctx->stage;
  std::vector<Up> ups = {
    {&s[0], fr->nfrm_utt, (size_t)conf->nutt * 4}, {&s[1], fr->f0, BF * 4}, {&s[2], fr->nhar, BF * 4},
    {&s[3], fr->ampl, BF * conf->maxnhar * 4}, {&s[4], fr->phse, BF * conf->maxnhar * 4},
    {&s[5], fr->psd, BF * conf->npsd * 4}, {&s[6], fr->psdres, BF * conf->npsd * 4},
    {&s[7], fr->edc, BF * nch * 4}, {&s[8], fr->enhar, BF * nch * 4},
    {&s[9], fr->eampl, BF * nch * conf->maxnhar_e * 4}, {&s[10], fr->ephse, BF * nch * conf->maxnhar_e * 4},
    {&s[11], opt->white, (size_t)conf->nutt * nch * nt * 4},
  };
  rc = upload_all(ctx, ups); if(rc) return rc;
  const size_t obytes = (size_t)conf->nutt * out->stride * 4;
  for(int i = 12; i < 15; i ++)
    if(s[i].reserve(obytes) != 0) return fail(LLSM_B200_ENOMEM, "device output staging");
  llsm_b200_frames d;
  d.nfrm_utt = fr->nfrm_utt ? s[0].as<int>() : nullptr;
  d.f0 = s[1].as<float>(); d.nhar = s[2].as<int>(); d.ampl = s[3].as<float>(); d.phse = s[4].as<float>();
  d.psd = s[5].as<float>(); d.psdres = fr->psdres ? s[6].as<float>() : nullptr;
  d.edc = s[7].as<float>(); d.enhar = s[8].as<int>(); d.eampl = s[9].as<float>(); d.ephse = s[10].as<float>();
  llsm_b200_soptions o = *opt;
  o.white = opt->white ? s[11].as<float>() : nullptr;
  llsm_b200_output od;
  od.y = s[12].as<float>(); od.y_sin = s[13].as<float>(); od.y_noise = s[14].as<float>();
  od.stride = out->stride;
  rc = llsm_b200_synthesize_l0(ctx, conf, &d, &o, &od);
  if(rc) return rc;
  if(out->y && cudaMemcpyAsync(out->y, od.y, obytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    return cuda_ok("D2H y");
  if(out->y_sin && cudaMemcpyAsync(out->y_sin, od.y_sin, obytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    return cuda_ok("D2H y_sin");
  if(out->y_noise && cudaMemcpyAsync(out->y_noise, od.y_noise, obytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    return cuda_ok("D2H y_noise");
  if(cudaStreamSynchronize(ctx->stream) != cudaSuccess) return cuda_ok("synchronize");
  return 0;
}

#include "api_analysis.inc"
#include "api_layer1.inc"
#include "api_rt.inc"

} // extern "C"
