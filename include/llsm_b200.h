/*
  llsm_b200.h -- C ABI of the B200 hot path (batched, structure-of-arrays).

  This is the FFI boundary of libllsm2_b200.so for the layer-0 analysis / synthesis loops of
  libllsm2. Plain pointers and sizes only. The reference API is one utterance at a time over
  pointer-chasing containers (llsm.h:310-313, :336-339); the entry points below take the same
  quantities as flat arrays for a whole batch of utterances, and the drop-in llsm_synthesize /
  llsm_analyze of include/llsm.h are thin packers around them (batch = 1).

  Every entry point returns 0 on success and a negative code on failure;
  llsm_b200_last_error() describes the last failure of the calling thread. There is no CPU
  fallback: without a CUDA device every compute call fails with LLSM_B200_ENODEVICE.

  Array layout (row-major, B = nutt utterances, F = nfrm frames per utterance):
    f0      [B][F]                 Hz, 0 = unvoiced                (llsm.h:98  LLSM_FRAME_F0)
    nhar    [B][F]                 harmonics in use                (llsm.h:134-138 llsm_hmframe.nhar)
    ampl    [B][F][maxnhar]        linear amplitude                (llsm_hmframe.ampl)
    phse    [B][F][maxnhar]        radians                         (llsm_hmframe.phse)
    psd     [B][F][npsd]           dB                              (llsm.h:157-165 llsm_nmframe.psd)
    psdres  [B][F][npsd]           dB residual, optional (NULL)    (llsm.h:101 LLSM_FRAME_PSDRES)
    edc     [B][F][nchannel]       envelope mean                   (llsm_nmframe.edc)
    enhar   [B][F][nchannel]       envelope harmonics in use       (llsm_nmframe.eenv[c]->nhar)
    eampl   [B][F][nchannel][maxnhar_e]                            (llsm_nmframe.eenv[c]->ampl)
    ephse   [B][F][nchannel][maxnhar_e]                            (llsm_nmframe.eenv[c]->phse)
  Waveforms: [B][nsamp] with a common row stride.
*/
#ifndef LLSM_B200_H
#define LLSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLSM_B200_MAXCHANNEL 8

#define LLSM_B200_OK         0
#define LLSM_B200_EINVAL    -1   /* inconsistent sizes / NULL required pointer */
#define LLSM_B200_ENODEVICE -2   /* no CUDA device, or library built without its kernels */
#define LLSM_B200_ECUDA     -3   /* a CUDA runtime call or kernel failed */
#define LLSM_B200_ENOMEM    -4
#define LLSM_B200_ERANGE    -5   /* size outside what the kernels support */

typedef struct llsm_b200_ctx llsm_b200_ctx; /* device, stream, scratch, cached plans */

/* Model configuration shared by a batch: the LLSM_CONF_* entries the path needs (llsm.h:115-128)
   plus the output sampling rate of llsm_soptions (llsm.h:290-299). */
typedef struct {
  int   nutt;                               /* B */
  int   nfrm;                               /* F (row stride of the frame arrays) */
  int   maxnhar;                            /* LLSM_CONF_MAXNHAR / row stride of ampl, phse */
  int   maxnhar_e;                          /* LLSM_CONF_MAXNHAR_E */
  int   npsd;                               /* LLSM_CONF_NPSD */
  int   nchannel;                           /* LLSM_CONF_NCHANNEL (<= LLSM_B200_MAXCHANNEL) */
  float fs;                                 /* sampling rate, Hz (fnyq = fs / 2) */
  float thop;                               /* LLSM_CONF_THOP, seconds */
  float chanfreq[LLSM_B200_MAXCHANNEL];     /* LLSM_CONF_CHANFREQ, nchannel - 1 used */
  float lip_radius;                         /* LLSM_CONF_LIPRADIUS */
} llsm_b200_conf;

typedef struct {
  const int*   nfrm_utt;  /* [B] frames actually present per utterance, NULL = all nfrm */
  const float* f0;
  const int*   nhar;
  const float* ampl;
  const float* phse;
  const float* psd;
  const float* psdres;    /* may be NULL */
  const float* edc;
  const int*   enhar;
  const float* eampl;
  const float* ephse;
} llsm_b200_frames;

/* Writable twin of llsm_b200_frames, filled by analysis. */
typedef struct {
  float* f0;              /* in/out: refined when f0_refine is set (layer0.c:487-488) */
  int*   nhar;
  float* ampl;
  float* phse;
  float* psd;
  float* psdres;
  float* edc;
  int*   enhar;
  float* eampl;
  float* ephse;
} llsm_b200_frames_out;

/* Synthesis options: llsm_soptions (llsm.h:290-299) + the noise source.
   white == NULL : the white-noise templates are drawn on the device (Philox + Box-Muller) from
                   `seed` -- statistically equivalent to, but not bit-equal with, the reference.
   white != NULL : [B][nchannel][ntemplate] host-drawn N(0,1) templates, ntemplate =
                   llsm_b200_template_length(); the drop-in llsm_synthesize fills this with the
                   reference's randn() sequence so outputs match the reference build. */
typedef struct {
  int      use_iczt;
  float    iczt_param_a;
  float    iczt_param_b;
  const float* white;
  uint64_t seed;
} llsm_b200_soptions;

typedef struct {
  float* y;               /* [B][stride] y_sin + y_noise            (llsm.h:246-252 llsm_output.y) */
  float* y_sin;           /* [B][stride]                                            (.y_sin)        */
  float* y_noise;         /* [B][stride]                                            (.y_noise)      */
  int    stride;          /* >= llsm_b200_output_length() */
} llsm_b200_output;

/* Analysis options: llsm_aoptions (llsm.h:260-272); sizes come from llsm_b200_conf. */
typedef struct {
  int   f0_refine;
  int   hm_method;        /* 0 = LLSM_AOPTION_HMPP, 1 = LLSM_AOPTION_HMCZT */
  float rel_winsize;
} llsm_b200_aoptions;

/* ---- context ---- */
llsm_b200_ctx* llsm_b200_create(int device);            /* NULL on failure */
void           llsm_b200_destroy(llsm_b200_ctx* ctx);
const char*    llsm_b200_last_error(void);
int            llsm_b200_set_stream(llsm_b200_ctx* ctx, void* cuda_stream); /* cudaStream_t */
int            llsm_b200_synchronize(llsm_b200_ctx* ctx);
/* number of kernels this library has launched on ctx since creation (bench.py gpu_launches) */
long long      llsm_b200_launch_count(const llsm_b200_ctx* ctx);

/* Per-kernel timing: while enabled, the analysis and synthesis pipelines record a named CUDA event after each of their
   kernels (every call with enable != 0 restarts the list); llsm_b200_kernel_timing_read waits for the last one and
   returns how many kernel intervals were recorded (<= max; -1 on error), their names (static strings) and durations
   in ms, in launch order. */
int llsm_b200_set_kernel_timing(llsm_b200_ctx* ctx, int enable);
int llsm_b200_kernel_timing_read(llsm_b200_ctx* ctx, int max, const char** names, float* ms);

/* ---- size helpers (host only, exact replicas of the reference's float expressions) ---- */
int llsm_b200_output_length(int nfrm, float thop, float fs);    /* layer0.c:643 */
int llsm_b200_template_length(int ny);                          /* dsputils.c:386-388 */

/* ---- layer-0 synthesis: replaces the frame loops of llsm_synthesize (layer0.c:636-664:
        llsm_synthesize_harmonics_l0 :117-146, llsm_synthesize_noise_excitation :535-555,
        llsm_synthesize_noise_envelope :289-316, llsm_filter_noise :557-634) ---- */
/* all pointers are DEVICE pointers on ctx's device */
int llsm_b200_synthesize_l0(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, const llsm_b200_soptions* opt, const llsm_b200_output* out);
/* all pointers are HOST pointers; copies in, runs, copies out, synchronises */
int llsm_b200_synthesize_l0_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, const llsm_b200_soptions* opt, const llsm_b200_output* out);

/* Only the harmonic component (y_sin): the harmonic-bank kernel alone, device pointers.
   Used for the residual resynthesis of analysis (layer0.c:498, options == NULL there) and by the
   benchmark's roofline leg. nsamp = row length of y_sin (layer0.c:118 `ny`). */
int llsm_b200_synthesize_harmonics(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, const llsm_b200_soptions* opt_or_null,
  float* y_sin, int nsamp, int stride);

/* ---- layer-0 analysis: replaces the frame loops of llsm_analyze (layer0.c:478-511:
        llsm_refine_f0 dsputils.c:72-94, llsm_harmonic_analysis :175-228, residual layer0.c:498-501,
        llsm_analyze_noise_psd :318-415, llsm_analyze_noise_envelope :417-469) ----
   x: [B][xstride] waveforms of nx samples; x_res (optional, may be NULL) receives x - x_sin. */
int llsm_b200_analyze_l0(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_aoptions* opt, const float* x, int nx, int xstride,
  const llsm_b200_frames_out* frames, float* x_res);
int llsm_b200_analyze_l0_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_aoptions* opt, const float* x, int nx, int xstride,
  const llsm_b200_frames_out* frames, float* x_res);


/* Analysis -> [chunk phase operations] -> synthesis in one call, HOST waveforms in and out; the chunk (every frame
   array) stays in HBM. The chain the reference's end-to-end test runs and times (test/test-layer0-anasynth.c:40-46:
   llsm_analyze then llsm_synthesize on the chunk; :62-66 with llsm_chunk_phasesync_rps + llsm_chunk_phasepropagate in
   between) -- the "layer0 analysis+synthesis" metric of BASELINE.json end to end.
     x [B][xstride], f0 [B][F] (read only; the refined track is returned in f0_refined when it is not NULL)
     phase_ops   bit 0: llsm_chunk_phasesync_rps(chunk, 0), bit 1: llsm_chunk_phasepropagate(chunk, +1), in that order
     sopt->white HOST pointer [B][nchannel][llsm_b200_template_length(ny)] or NULL (device generator)
     out         HOST pointers, stride >= llsm_b200_output_length(); any of y / y_sin / y_noise may be NULL */
int llsm_b200_anasynth_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_aoptions* aopt,
  const llsm_b200_soptions* sopt, const float* x, int nx, int xstride, const float* f0, float* f0_refined,
  int phase_ops, const llsm_b200_output* out);

/* Slice sizes the call above uses for a batch of nutt utterances x nfrm frames moving bytes_per_utt host bytes (in + out)
   per utterance: small slices at both ends (the first upload and the last download overlap with nothing), few large
   ones between (every kernel of a slice ends in a tail). Writes up to cap_sizes entries (each >= 1, sum = nutt) and
   returns their number; host arithmetic only. LLSM_B200_HOST_SLICES / LLSM_B200_HOST_SLICE_LIST override. */
int llsm_b200_host_slice_plan(int nutt, int nfrm, size_t bytes_per_utt, int* out_sizes, int cap_sizes);


/* ---- per-frame routines of the reference's dsputils.h that its tests call directly (test/test-dsputils.c:80,
        test/test-harmonic.c:40-43), as batch entries ----
   llsm_b200_harmonic_analysis    llsm_harmonic_analysis (dsputils.c:175-228): nhar / ampl / phse of x at the frame
                                  centres round(i thop fs) for a KNOWN f0 [B][F] (no refinement, no noise model);
                                  opt->hm_method 0 = peak picking, 1 = CZT. Device pointers; _host: host pointers.
   llsm_b200_refine_f0_host       llsm_refine_f0 (dsputils.c:72-94): f0 [B][F] refined in place (host pointers)
   llsm_b200_harmonic_frames_host llsm_synthesize_harmonic_frame (iczt = 0) / _iczt (iczt = 1) (dsputils.c:328-351) for
                                  nfrm frames: unwindowed y [nfrm][nx], f0n in cycles per sample (host pointers) */
int llsm_b200_harmonic_analysis(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_aoptions* opt,
  const float* x, int nx, int xstride, const float* f0, int* nhar, float* ampl, float* phse);
int llsm_b200_harmonic_analysis_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_aoptions* opt,
  const float* x, int nx, int xstride, const float* f0, int* nhar, float* ampl, float* phse);
int llsm_b200_refine_f0_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const float* x, int nx, int xstride, float* f0);
int llsm_b200_harmonic_frames_host(llsm_b200_ctx* ctx, int nfrm, int maxnhar, const int* nhar, const float* f0n,
  const float* ampl, const float* phse, int nx, int iczt, float* y);


/* ---- layer-1 members (llsm.h:105-108): RD, VTMAGN, VSPHSE per frame, flat ----
     rd      [B][F]            Rd glottal parameter (every frame, smoothed track)
     vtmagn  [B][F][nspec]     vocal-tract magnitude response, dB (voiced frames)
     vsphse  [B][F][maxnhar]   vocal-source harmonic phases (voiced frames)
     nvs     [B][F]            length of the VSPHSE vector (0 when unvoiced)            */
typedef struct {
  float* rd;
  float* vtmagn;
  float* vsphse;
  int*   nvs;
  int    nspec;          /* LLSM_CONF_NSPEC = nfft / 2 + 1 */
} llsm_b200_layer1;

/* ---- flat serialisation of a batch, host memory ----
   The reference has no on-disk / wire format (llsm_chunk only lives in memory, llsm.h:310-313). A blob is the
   structure-of-arrays batch in one relocatable buffer: 256-byte header (magic "LLSMB200", version, the conf, which
   arrays are present, their offsets) followed by the arrays of llsm_b200_frames, 64-byte aligned. unpack returns
   pointers INTO the blob (no copy). Absent optional arrays (nfrm_utt, psdres; or nhar / ampl / phse of a layer-1
   batch) are NULL on both sides. */
size_t llsm_b200_frames_blob_size(const llsm_b200_conf* conf, const llsm_b200_frames* frames);
int llsm_b200_frames_pack(const llsm_b200_conf* conf, const llsm_b200_frames* frames, void* blob, size_t size);
int llsm_b200_frames_unpack(const void* blob, size_t size, llsm_b200_conf* conf, llsm_b200_frames* frames);

/* ---- frame coder: coder.c:46-292 for a batch, device pointers ----
   A frame <-> a vector of order_spec + order_bap + 3 numbers: [voicing, f0, Rd, order_spec mel-cepstral
   coefficients of the total power spectrum, order_bap band aperiodicities] (llsm_create_coder(conf, order_spec,
   order_bap), llsm_coder_encode, llsm_coder_decode_layer0 / _layer1). nspec = LLSM_CONF_NSPEC (nfft / 2 + 1 of the
   layer-1 conversion). Encoding reads f0, psd and, for voiced frames, rd and vtmagn; enc is [B][F][dim].
   Decoding writes f0, rd, psd, nhar and either ampl / phse (use_layer1 = 0: llsm_coder_decode_layer0) or
   layer1->vtmagn / vsphse (use_layer1 = 1; layer1->nvs receives nhar). Rows of ampl / phse / vsphse are
   conf->maxnhar long: a decoded frame with more harmonics than that (f0 < fnyq / maxnhar) is cut there. */
int llsm_b200_coder_dimension(int order_spec, int order_bap);
int llsm_b200_coder_encode(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const float* f0, const float* psd, const llsm_b200_layer1* layer1, int order_spec, int order_bap, float* enc);
int llsm_b200_coder_decode(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const float* enc, int order_spec, int order_bap, int use_layer1,
  const llsm_b200_frames_out* out, const llsm_b200_layer1* layer1);
/* the same through host buffers (copies in and out, synchronises) */
int llsm_b200_coder_encode_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const float* f0, const float* psd, const llsm_b200_layer1* layer1, int order_spec, int order_bap, float* enc);
int llsm_b200_coder_decode_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const float* enc, int order_spec, int order_bap, int use_layer1,
  const llsm_b200_frames_out* out, const llsm_b200_layer1* layer1);

/* ---- chunk phase utilities, in place on device arrays ----
   What real use runs between analysis and synthesis (test/test-layer0-anasynth.c:62-63, test-llsmrt.c:88,112):
     llsm_chunk_phasepropagate(chunk, sign)      layer0.c:694-706  theta_i = cumsum(f0)_i thop sign 2 pi
     llsm_chunk_phasesync_rps(chunk, l1_based)   layer0.c:687-692  theta_i = -phse_i[0]  (-VSPHSE_i[0] when l1_based)
   followed by llsm_frame_phaseshift (frame.c:152-166): phse[k], the envelope phases ephse[c][k] and VSPHSE[k]
   become wrap(. + theta_i (k + 1)). Reads frames->{f0, nhar, enhar}; rewrites frames->{phse, ephse} and, when
   layer1 != NULL, layer1->vsphse (lengths layer1->nvs). */
int llsm_b200_chunk_phasepropagate(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const llsm_b200_frames_out* frames, const llsm_b200_layer1* layer1, int sign);
int llsm_b200_chunk_phasesync_rps(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const llsm_b200_frames_out* frames, const llsm_b200_layer1* layer1, int layer1_based);

/* ---- frame interpolation / time-stretch of a layer-1 batch, device arrays ----
   The reference keeps this step in a demo (test/demo-stretch.c), between llsm_chunk_tolayer1 +
   llsm_chunk_phasepropagate(-1) and llsm_chunk_tolayer0 + llsm_chunk_phasepropagate(+1):
     interp_llsm_frame  test/demo-stretch.c:50-129  f0, Rd, VSPHSE (circular), VTMAGN (dB, faded in / out at voicing
                                                    boundaries, floored at -80), voiced / unvoiced cases
     interp_nmframe     test/demo-stretch.c:16-44   psd, edc, envelope harmonics
     the frame loop     test/demo-stretch.c:169-185 out[i] = copy(frames[base[i]]) blended towards frames[base[i] + 1]
                                                    with weight ratio[i]; PSDRES copied from frames[residx[i]]
   conf describes the SOURCE batch (nfrm frames); the destination arrays hold nfrm_new frames per utterance with the
   same row strides. base / ratio / residx are device arrays of nfrm_new entries (map_per_utt = 0: one map for the
   batch) or [B][nfrm_new] (map_per_utt = 1); residx == NULL takes PSDRES from frame base[i]. base is clamped to
   [0, nfrm - 2], residx to [0, nfrm - 1]. psdres and the layer-0 harmonics (nhar, ampl, phse: carried over unchanged
   from frame base[i]) are optional on both sides. Unvoiced output frames get nvs = 0 and zeroed vtmagn / vsphse rows. */
int llsm_b200_frames_stretch(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_frames* src,
  const llsm_b200_layer1* src_layer1, int nfrm_new, const int* base, const float* ratio, const int* residx,
  int map_per_utt, const llsm_b200_frames_out* dst, const llsm_b200_layer1* dst_layer1);
/* host helper: the uniform map of test/demo-stretch.c:170-175 in the reference's float arithmetic
   (mapped = (float)i * nfrm / nfrm_new; base = (int)mapped; ratio = mapped - base; base = min(base, nfrm - 2)).
   residx (optional) receives the unclamped base, to which the caller adds its jitter (:173-174). HOST pointers. */
int llsm_b200_stretch_map(int nfrm, int nfrm_new, int* base, float* ratio, int* residx);

/* llsm_chunk_tolayer1 (layer1.c:129-149) for a batch: Rd track (glottal fitting + smoothing),
   vocal-tract envelope and source phases of every voiced frame. Device pointers. Reads
   frames->{nfrm_utt, f0, nhar, ampl, phse}. nfft as in the reference call (power of two). */
int llsm_b200_tolayer1(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, int nfft, const llsm_b200_layer1* out);
/* llsm_chunk_tolayer0 / llsm_frame_tolayer0 (layer1.c:151-201): harmonic model from the layer-1
   members. Writes nhar / ampl / phse ([B][F], [B][F][maxnhar]). Device pointers. */
int llsm_b200_tolayer0(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const float* f0, const llsm_b200_layer1* in, int* nhar, float* ampl, float* phse);


/* ---- layer-1 synthesis: llsm_synthesize with use_l1 = 1 (layer0.c:148-287 deterministic part: pulse
        tracker, llsm_make_filtered_pulse llsmutils.c:132-201, HM <-> PbP cross-fade; then the layer-0
        noise path). Device pointers.
   l1      layer-1 members (in); pbpsyn [B][F] (LLSM_FRAME_PBPSYN flags, NULL = none set)
   frames  noise model (psd, psdres, edc, enhar, eampl, ephse) and f0 are required; when nhar / ampl /
           phse are all non-NULL they are the stored harmonic models, otherwise the harmonic frames are
           derived from layer 1 on the fly (layer0.c:264-265).
   User llsm_pbpeffect callbacks are host functions and are not part of this entry point. */
int llsm_b200_synthesize_l1(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, const llsm_b200_layer1* l1, const int* pbpsyn,
  const llsm_b200_soptions* opt, const llsm_b200_output* out);


/* Per-pulse hook of llsm_pbpeffect (llsm.h:190-197, called at layer0.c:208-217): the library calls it
   once per glottal pulse, in time order, on the host. It receives the glottal-flow parameters of the
   pulse (llsm_gfm fields) and delta_t, may modify them, and returns non-zero when the frame carries an
   effect (the reference then round-trips the model through llsm_gfm even if nothing changed). */
typedef int (*llsm_b200_pulse_hook)(void* user, int utt, int frame, float* Fa, float* Rk, float* Rg,
  float* T0, float* Ee, float* delta_t);
/* Host-buffer form of llsm_b200_synthesize_l1. hook == NULL: everything runs on the device. With a
   hook the (cheap, strictly sequential) pulse tracker runs on the host so that user callbacks are
   honoured in order; pulse generation, harmonic frames, noise and mixing stay on the device. */
int llsm_b200_synthesize_l1_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, const llsm_b200_layer1* l1, const int* pbpsyn,
  const llsm_b200_soptions* opt, const llsm_b200_output* out, llsm_b200_pulse_hook hook, void* user);
/* Host-buffer forms of the layer-1 conversions. */
int llsm_b200_tolayer1_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, int nfft, const llsm_b200_layer1* out);
int llsm_b200_tolayer0_host(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const int* nfrm_utt,
  const float* f0, const llsm_b200_layer1* in, int* nhar, float* ampl, float* phse);


/* ---- frame-range sharding of one batch across GPUs ----
   Only frames [frame_lo, frame_hi) of every utterance contribute: the outputs are PARTIAL sums over the
   full-length rows (zero where the shard does not reach). Summing the partial y_sin / y_noise of all
   shards gives the unsharded result; since a frame reaches at most llsm_b200_halo_length() samples
   beyond its centre, only boundary strips have to be exchanged (libllsm2_b200/parallel.py does it with
   one NCCL all-gather). The frame arrays must hold valid data for [frame_lo - 3, frame_hi + 3). */
int llsm_b200_synthesize_l0_shard(llsm_b200_ctx* ctx, const llsm_b200_conf* conf,
  const llsm_b200_frames* frames, const llsm_b200_soptions* opt, const llsm_b200_output* out,
  int frame_lo, int frame_hi);
int llsm_b200_halo_length(const llsm_b200_conf* conf);              /* samples a frame reaches past its centre */
int llsm_b200_frame_position(int i, float thop, float fs);          /* round(i * thop * fs), layer0.c:127-128 */

/* The exchange that completes frame-range shards (north_star: "a single NCCL all-gather of the overlap-add boundary
   samples"). world ranks, rank r owning the frames [frame_edges[r], frame_edges[r + 1]) (frame_edges: HOST array of
   world + 1 ascending entries, 0 ... nfrm) and the output samples [llsm_b200_shard_position(edge r), ..(edge r + 1)).
   After llsm_b200_synthesize_l0_shard, llsm_b200_halo_exchange packs what this rank's frames add outside its own
   sample range into two strips of llsm_b200_halo_length() samples per component (kernel), all-gathers the strips
   (ncclAllGather on ctx's stream), adds the other ranks' strips into the owned range and rewrites y = y_sin + y_noise
   there (kernel). On return the owned sample range of out->{y_sin, y_noise, y} is complete and equal to the unsharded
   result; samples outside it still hold this rank's partial sums. Device pointers.
   The communicator belongs to the context: either created by the library from an ncclUniqueId that the caller moves
   from rank 0 to the other ranks by its own means (llsm_b200_comm_unique_id on rank 0, llsm_b200_comm_init on every
   rank: collective), or the caller's own ncclComm_t (llsm_b200_comm_attach; it must come from the libnccl.so.2 this
   process has loaded). NCCL is bound at run time; without it these calls fail with LLSM_B200_ENODEVICE. */
#define LLSM_B200_COMM_ID_BYTES 128
int llsm_b200_comm_unique_id(void* id /* LLSM_B200_COMM_ID_BYTES */);
int llsm_b200_comm_init(llsm_b200_ctx* ctx, const void* id, int rank, int world);
int llsm_b200_comm_attach(llsm_b200_ctx* ctx, void* nccl_comm, int rank, int world);
int llsm_b200_comm_destroy(llsm_b200_ctx* ctx);
int llsm_b200_shard_position(const llsm_b200_conf* conf, int frame_edge);   /* 0, frame positions, ny at nfrm */
int llsm_b200_halo_exchange(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_output* out,
  const int* frame_edges);

/* ---- streaming synthesis: llsm_rtsynth_buffer_* (llsmrt.h:33-56, llsmrt.c:157-602) --------------------
   One llsm_b200_rt advances conf->nutt independent streams that share fs and thop (hence one hop
   schedule: llsm_update_cycle, llsmrt.c:110-129). Every stream owns its ring buffers, circular noise
   templates and previous noise model in HBM; one kernel launch per fed frame serves all streams.
   Replaces: llsm_create_rtsynth_buffer (:157), _feed (:505), _fetch/_fetch_decomposed (:523/:545, here the
   samples of a feed are returned by the feed itself), _getlatency (:568), _clear (:578), delete (:225).
   conf->nfrm is ignored. opt->white: N(0,1) templates [nutt][nchannel][llsm_b200_rt_template_length(fs)]
   (device pointer, or host pointer with white_on_host = 1) or NULL for the device generator (opt->seed). */
typedef struct llsm_b200_rt llsm_b200_rt;
int llsm_b200_rt_fft_size(float fs, float thop);                     /* llsmrt.c:181; feed keeps min(nhar, nfft) harmonics */
int llsm_b200_rt_template_length(float fs);                          /* min(20000, (int)fs) + 128 */
int llsm_b200_rt_create(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_soptions* opt,
  int white_on_host, llsm_b200_rt** out);
void llsm_b200_rt_destroy(llsm_b200_rt* rt);
int llsm_b200_rt_latency(const llsm_b200_rt* rt);                    /* llsmrt.c:568-571 */
int llsm_b200_rt_output_length(const llsm_b200_rt* rt, int nfeed);   /* samples the next nfeed feeds produce */
int llsm_b200_rt_clear(llsm_b200_rt* rt);                            /* llsmrt.c:578-602 */
/* Feed nfeed consecutive frames per stream (frames: [nutt][nfeed][..]) and receive the samples they
   release, periodic and aperiodic parts apart (fetch_decomposed): out_p/out_ap [nutt][out_stride],
   *nout samples per stream. Device pointers; the _host variant copies in and out and synchronises. */
int llsm_b200_rt_feed(llsm_b200_rt* rt, const llsm_b200_frames* frames, int nfeed,
  float* out_p, float* out_ap, int out_stride, int* nout);
int llsm_b200_rt_feed_host(llsm_b200_rt* rt, const llsm_b200_frames* frames, int nfeed,
  float* out_p, float* out_ap, int out_stride, int* nout);

/* Streams synthesised from layer-1 members (soptions.use_l1 = 1, llsmrt.c:305-419): per frame the pulse
   tracker locked on the first source harmonic, HM <-> pulse-by-pulse onset (two periods early) and
   termination (trapezoid catch-up), filtered glottal pulses overlap-added in a forward/backward pulse
   buffer (llsm_dualbuffer, buffer.h:146-209). nspec = LLSM_CONF_NSPEC. host_tracker = 1 keeps the
   (tiny, sequential) tracker on the host so that llsm_pbpeffect callbacks run there, once per pulse, in time
   order (feed_l1_host only); 0 runs it inside the feed kernel.
   frames->ampl == NULL: the harmonic model of each frame is derived from layer 1 (llsm_frame_tolayer0,
   llsmrt.c:346-347,388-389). pbpsyn: [nutt][nfeed] LLSM_FRAME_PBPSYN flags or NULL. */
int llsm_b200_rt_create_l1(llsm_b200_ctx* ctx, const llsm_b200_conf* conf, const llsm_b200_soptions* opt,
  int white_on_host, int nspec, int host_tracker, llsm_b200_rt** out);
int llsm_b200_rt_feed_l1(llsm_b200_rt* rt, const llsm_b200_frames* frames, const llsm_b200_layer1* layer1,
  const int* pbpsyn, int nfeed, float* out_p, float* out_ap, int out_stride, int* nout);
int llsm_b200_rt_feed_l1_host(llsm_b200_rt* rt, const llsm_b200_frames* frames, const llsm_b200_layer1* layer1,
  const int* pbpsyn, int nfeed, llsm_b200_pulse_hook hook, void* user,
  float* out_p, float* out_ap, int out_stride, int* nout);

#ifdef __cplusplus
}
#endif
#endif
