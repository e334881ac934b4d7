/*
  Reference harness -- TEST INFRASTRUCTURE (oracle), not product code.

  Compiled together with the UNMODIFIED reference sources (taken where they lie under
  /root/reference; see oracle/Makefile) and the ciglet shim into oracle/_ref/libllsm2_ref*.so.
  It only drives the reference's public API (llsm.h) with flat "structure of arrays" buffers so
  that the Python tests and bench.py can feed the very same numbers to the reference build and to
  the CUDA library. Layout of the flat buffers = include/llsm_b200.h (one utterance per call).
*/
#define _POSIX_C_SOURCE 199309L
#include <time.h>
#include <ciglet/ciglet.h>
#include "llsm.h"
#include "llsmrt.h"
#include "dsputils.h"
#include "llsmutils.h"

/* ---- chunk construction from flat arrays (layer 0) ---- */
static llsm_chunk* chunk_from_soa(int nfrm, float fs, float thop, int maxnhar, int maxnhar_e,
  int npsd, int nchannel, const float* chanfreq, float lip_radius,
  const float* f0, const int* nhar, const float* ampl, const float* phse,
  const float* psd, const float* psdres, const float* edc, const int* enhar,
  const float* eampl, const float* ephse) {
  llsm_aoptions* opt = llsm_create_aoptions();
  opt -> thop = thop;
  opt -> maxnhar = maxnhar;
  opt -> maxnhar_e = maxnhar_e;
  opt -> npsd = npsd;
  opt -> nchannel = nchannel;
  free(opt -> chanfreq);
  opt -> chanfreq = calloc(nchannel > 1 ? nchannel - 1 : 1, sizeof(FP_TYPE));
  for(int c = 0; c < nchannel - 1; c ++) opt -> chanfreq[c] = chanfreq[c];
  opt -> lip_radius = lip_radius;
  llsm_container* conf = llsm_aoptions_toconf(opt, fs / 2.0);
  ((int*)llsm_container_get(conf, LLSM_CONF_NFRM))[0] = nfrm;
  llsm_chunk* chunk = llsm_create_chunk(conf, 1);
  llsm_delete_container(conf);
  llsm_delete_aoptions(opt);

  for(int i = 0; i < nfrm; i ++) {
    llsm_container* frame = chunk -> frames[i];
    ((FP_TYPE*)llsm_container_get(frame, LLSM_FRAME_F0))[0] = f0[i];
    if(f0[i] > 0 && nhar != NULL) {
      llsm_hmframe* hm = llsm_create_hmframe(nhar[i]);
      memcpy(hm -> ampl, ampl + (size_t)i * maxnhar, nhar[i] * sizeof(FP_TYPE));
      memcpy(hm -> phse, phse + (size_t)i * maxnhar, nhar[i] * sizeof(FP_TYPE));
      llsm_container_attach(frame, LLSM_FRAME_HM, hm, llsm_delete_hmframe, llsm_copy_hmframe);
    }
    llsm_nmframe* nm = llsm_container_get(frame, LLSM_FRAME_NM);
    if(psd != NULL) memcpy(nm -> psd, psd + (size_t)i * npsd, npsd * sizeof(FP_TYPE));
    for(int c = 0; c < nchannel; c ++) {
      if(edc != NULL) nm -> edc[c] = edc[(size_t)i * nchannel + c];
      if(enhar != NULL) {
        int ne = enhar[(size_t)i * nchannel + c];
        llsm_hmframe* e = llsm_create_hmframe(ne);
        size_t off = ((size_t)i * nchannel + c) * maxnhar_e;
        memcpy(e -> ampl, eampl + off, ne * sizeof(FP_TYPE));
        memcpy(e -> phse, ephse + off, ne * sizeof(FP_TYPE));
        llsm_copy_hmframe_inplace(nm -> eenv[c], e);
        llsm_delete_hmframe(e);
      }
    }
    if(psdres != NULL) {
      FP_TYPE* res = llsm_create_fparray(npsd);
      memcpy(res, psdres + (size_t)i * npsd, npsd * sizeof(FP_TYPE));
      llsm_container_attach(frame, LLSM_FRAME_PSDRES, res, llsm_delete_fparray,
        llsm_copy_fparray);
    }
  }
  return chunk;
}

int ref_output_length(int nfrm, float thop, float fs) {
  return round((nfrm + 1) * thop * fs); /* same float expression as layer0.c:643 */
}

/* llsm_synthesize on one utterance; srand(seed) first so that the noise template is reproducible
   (the template is drawn inside llsm_synthesize, layer0.c:544 -> dsputils.c:353-361).
   Returns ny, or -1 if the reference returned NULL. */
int ref_synthesize_soa(int nfrm, float fs, float thop, int maxnhar, int maxnhar_e, int npsd,
  int nchannel, const float* chanfreq, float lip_radius, int use_iczt,
  const float* f0, const int* nhar, const float* ampl, const float* phse,
  const float* psd, const float* psdres, const float* edc, const int* enhar,
  const float* eampl, const float* ephse, unsigned seed,
  float* y, float* y_sin, float* y_noise) {
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nchannel,
    chanfreq, lip_radius, f0, nhar, ampl, phse, psd, psdres, edc, enhar, eampl, ephse);
  llsm_soptions* sopt = llsm_create_soptions(fs);
  sopt -> use_iczt = use_iczt;
  srand(seed);
  llsm_output* out = llsm_synthesize(sopt, chunk);
  int ny = -1;
  if(out != NULL) {
    ny = out -> ny;
    if(y != NULL) memcpy(y, out -> y, ny * sizeof(float));
    if(y_sin != NULL) memcpy(y_sin, out -> y_sin, ny * sizeof(float));
    if(y_noise != NULL) memcpy(y_noise, out -> y_noise, ny * sizeof(float));
    llsm_delete_output(out);
  }
  llsm_delete_soptions(sopt);
  llsm_delete_chunk(chunk);
  return ny;
}

/* The white-noise draw of one llsm_synthesize call, channel by channel, exactly as
   llsm_generate_white_noise consumes rand() (dsputils.c:353-361, :385-388): used by the tests to
   hand the CUDA path the same template. dst: [nchannel][min(20000,ny)+128]. */
int ref_draw_white_noise(int ny, int nchannel, unsigned seed, float* dst) {
  int nt = min(20000, ny) + 128;
  srand(seed);
  for(int c = 0; c < nchannel; c ++) {
    FP_TYPE* w = llsm_generate_white_noise(nt);
    memcpy(dst + (size_t)c * nt, w, nt * sizeof(float));
    free(w);
  }
  return nt;
}

/* llsm_analyze on one utterance, results unpacked to flat arrays. f0 is in/out (refinement,
   layer0.c:487-488). hm_method: 0 = peak picking, 1 = CZT. Returns 0 on success. */
int ref_analyze_soa(const float* x, int nx, float fs, float* f0, int nfrm, float thop,
  int maxnhar, int maxnhar_e, int npsd, int nchannel, const float* chanfreq,
  int f0_refine, int hm_method, float rel_winsize,
  int* nhar, float* ampl, float* phse, float* psd, float* psdres, float* edc, int* enhar,
  float* eampl, float* ephse, float* x_res) {
  llsm_aoptions* opt = llsm_create_aoptions();
  opt -> thop = thop;
  opt -> maxnhar = maxnhar;
  opt -> maxnhar_e = maxnhar_e;
  opt -> npsd = npsd;
  opt -> nchannel = nchannel;
  free(opt -> chanfreq);
  opt -> chanfreq = calloc(nchannel > 1 ? nchannel - 1 : 1, sizeof(FP_TYPE));
  for(int c = 0; c < nchannel - 1; c ++) opt -> chanfreq[c] = chanfreq[c];
  opt -> f0_refine = f0_refine;
  opt -> hm_method = hm_method;
  opt -> rel_winsize = rel_winsize;
  FP_TYPE* xap = NULL;
  FP_TYPE* xcopy = malloc(nx * sizeof(FP_TYPE));
  memcpy(xcopy, x, nx * sizeof(FP_TYPE));
  llsm_chunk* chunk = llsm_analyze(opt, xcopy, nx, fs, f0, nfrm, & xap);
  free(xcopy);
  llsm_delete_aoptions(opt);
  if(chunk == NULL) return -1;
  if(x_res != NULL) memcpy(x_res, xap, nx * sizeof(float));
  free(xap);
  for(int i = 0; i < nfrm; i ++) {
    llsm_container* frame = chunk -> frames[i];
    llsm_hmframe* hm = llsm_container_get(frame, LLSM_FRAME_HM);
    llsm_nmframe* nm = llsm_container_get(frame, LLSM_FRAME_NM);
    FP_TYPE* res = llsm_container_get(frame, LLSM_FRAME_PSDRES);
    int nh = (f0[i] > 0 && hm != NULL) ? hm -> nhar : 0;
    nhar[i] = nh;
    for(int k = 0; k < maxnhar; k ++) {
      ampl[(size_t)i * maxnhar + k] = k < nh ? hm -> ampl[k] : 0;
      phse[(size_t)i * maxnhar + k] = k < nh ? hm -> phse[k] : 0;
    }
    memcpy(psd + (size_t)i * npsd, nm -> psd, npsd * sizeof(float));
    if(psdres != NULL) {
      if(res != NULL) memcpy(psdres + (size_t)i * npsd, res, npsd * sizeof(float));
      else memset(psdres + (size_t)i * npsd, 0, npsd * sizeof(float));
    }
    for(int c = 0; c < nchannel; c ++) {
      edc[(size_t)i * nchannel + c] = nm -> edc[c];
      int ne = f0[i] > 0 ? nm -> eenv[c] -> nhar : 0;
      enhar[(size_t)i * nchannel + c] = ne;
      size_t off = ((size_t)i * nchannel + c) * maxnhar_e;
      for(int k = 0; k < maxnhar_e; k ++) {
        eampl[off + k] = k < ne ? nm -> eenv[c] -> ampl[k] : 0;
        ephse[off + k] = k < ne ? nm -> eenv[c] -> phse[k] : 0;
      }
    }
  }
  llsm_delete_chunk(chunk);
  return 0;
}

/* timing helper for bench.py --impl reference: nrep x llsm_synthesize on the same chunk */
double ref_time_synthesize_soa(int nrep, int nfrm, float fs, float thop, int maxnhar,
  int maxnhar_e, int npsd, int nchannel, const float* chanfreq, float lip_radius, int use_iczt,
  const float* f0, const int* nhar, const float* ampl, const float* phse,
  const float* psd, const float* psdres, const float* edc, const int* enhar,
  const float* eampl, const float* ephse) {
  struct timespec t0, t1;
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nchannel,
    chanfreq, lip_radius, f0, nhar, ampl, phse, psd, psdres, edc, enhar, eampl, ephse);
  llsm_soptions* sopt = llsm_create_soptions(fs);
  sopt -> use_iczt = use_iczt;
  clock_gettime(CLOCK_MONOTONIC, & t0);
  for(int r = 0; r < nrep; r ++) {
    llsm_output* out = llsm_synthesize(sopt, chunk);
    llsm_delete_output(out);
  }
  clock_gettime(CLOCK_MONOTONIC, & t1);
  llsm_delete_soptions(sopt);
  llsm_delete_chunk(chunk);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- layer 1 (layer1.c) ---- */
int ref_tolayer1_soa(int nfrm, float fs, float thop, int maxnhar, float lip_radius, int nfft,
  const float* f0, const int* nhar, const float* ampl, const float* phse,
  float* rd, float* vtmagn, float* vsphse, int* nvs) {
  float cf[3] = {2000, 4000, 8000};
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, 4, 8, 4, cf, lip_radius,
    f0, nhar, ampl, phse, NULL, NULL, NULL, NULL, NULL, NULL);
  llsm_chunk_tolayer1(chunk, nfft);
  int nspec = nfft / 2 + 1;
  for(int i = 0; i < nfrm; i ++) {
    FP_TYPE* r = llsm_container_get(chunk -> frames[i], LLSM_FRAME_RD);
    FP_TYPE* vt = llsm_container_get(chunk -> frames[i], LLSM_FRAME_VTMAGN);
    FP_TYPE* vs = llsm_container_get(chunk -> frames[i], LLSM_FRAME_VSPHSE);
    rd[i] = r != NULL ? r[0] : -1;
    memset(vtmagn + (size_t)i * nspec, 0, nspec * sizeof(float));
    memset(vsphse + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
    nvs[i] = 0;
    if(vt != NULL) memcpy(vtmagn + (size_t)i * nspec, vt, nspec * sizeof(float));
    if(vs != NULL) {
      nvs[i] = llsm_fparray_length(vs);
      memcpy(vsphse + (size_t)i * maxnhar, vs, nvs[i] * sizeof(float));
    }
  }
  llsm_delete_chunk(chunk);
  return 0;
}

int ref_tolayer0_soa(int nfrm, float fs, float thop, int maxnhar, float lip_radius, int nspec,
  const float* f0, const float* rd, const float* vtmagn, const float* vsphse, const int* nvs,
  int* nhar_out, float* ampl, float* phse) {
  float cf[3] = {2000, 4000, 8000};
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, 4, 8, 4, cf, lip_radius,
    f0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
  llsm_container_attach(chunk -> conf, LLSM_CONF_NSPEC, llsm_create_int(nspec), llsm_delete_int, llsm_copy_int);
  for(int i = 0; i < nfrm; i ++) {
    llsm_container* fr = chunk -> frames[i];
    llsm_container_attach(fr, LLSM_FRAME_HM, NULL, NULL, NULL);
    llsm_container_attach(fr, LLSM_FRAME_RD, llsm_create_fp(rd[i]), llsm_delete_fp, llsm_copy_fp);
    if(f0[i] > 0) {
      FP_TYPE* vt = llsm_create_fparray(nspec);
      memcpy(vt, vtmagn + (size_t)i * nspec, nspec * sizeof(float));
      FP_TYPE* vs = llsm_create_fparray(nvs[i]);
      memcpy(vs, vsphse + (size_t)i * maxnhar, nvs[i] * sizeof(float));
      llsm_container_attach(fr, LLSM_FRAME_VTMAGN, vt, llsm_delete_fparray, llsm_copy_fparray);
      llsm_container_attach(fr, LLSM_FRAME_VSPHSE, vs, llsm_delete_fparray, llsm_copy_fparray);
    }
  }
  llsm_chunk_tolayer0(chunk);
  for(int i = 0; i < nfrm; i ++) {
    llsm_hmframe* hm = llsm_container_get(chunk -> frames[i], LLSM_FRAME_HM);
    int n = hm != NULL ? hm -> nhar : 0;
    nhar_out[i] = n;
    memset(ampl + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
    memset(phse + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
    if(n > 0) {
      memcpy(ampl + (size_t)i * maxnhar, hm -> ampl, n * sizeof(float));
      memcpy(phse + (size_t)i * maxnhar, hm -> phse, n * sizeof(float));
    }
  }
  llsm_delete_chunk(chunk);
  return 0;
}

/* layer-1 synthesis (test/test-layer1-anasynth.c:29-56 pattern): L0 frames -> llsm_chunk_tolayer1 ->
   optional removal of the HM members -> PBPSYN flags -> llsm_synthesize with use_l1 = 1.
   The layer-1 members are also returned so that the device path can be fed the same numbers. */
int ref_synthesize_l1_soa(int nfrm, float fs, float thop, int maxnhar, int maxnhar_e, int npsd,
  int nchannel, const float* chanfreq, float lip_radius, int nfft, int remove_hm, const int* pbpsyn,
  const float* f0, const int* nhar, const float* ampl, const float* phse,
  const float* psd, const float* psdres, const float* edc, const int* enhar,
  const float* eampl, const float* ephse, unsigned seed,
  float* rd, float* vtmagn, float* vsphse, int* nvs,
  float* y, float* y_sin, float* y_noise) {
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nchannel,
    chanfreq, lip_radius, f0, nhar, ampl, phse, psd, psdres, edc, enhar, eampl, ephse);
  llsm_chunk_tolayer1(chunk, nfft);
  int nspec = nfft / 2 + 1;
  for(int i = 0; i < nfrm; i ++) {
    llsm_container* fr = chunk -> frames[i];
    FP_TYPE* r = llsm_container_get(fr, LLSM_FRAME_RD);
    FP_TYPE* vt = llsm_container_get(fr, LLSM_FRAME_VTMAGN);
    FP_TYPE* vs = llsm_container_get(fr, LLSM_FRAME_VSPHSE);
    rd[i] = r != NULL ? r[0] : 0;
    memset(vtmagn + (size_t)i * nspec, 0, nspec * sizeof(float));
    memset(vsphse + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
    nvs[i] = 0;
    if(vt != NULL) memcpy(vtmagn + (size_t)i * nspec, vt, nspec * sizeof(float));
    if(vs != NULL) { nvs[i] = llsm_fparray_length(vs); memcpy(vsphse + (size_t)i * maxnhar, vs, nvs[i] * sizeof(float)); }
    if(remove_hm) llsm_container_attach(fr, LLSM_FRAME_HM, NULL, NULL, NULL);
    if(pbpsyn != NULL && pbpsyn[i])
      llsm_container_attach(fr, LLSM_FRAME_PBPSYN, llsm_create_int(1), llsm_delete_int, llsm_copy_int);
  }
  llsm_soptions* sopt = llsm_create_soptions(fs);
  sopt -> use_l1 = 1;
  srand(seed);
  llsm_output* out = llsm_synthesize(sopt, chunk);
  int ny = -1;
  if(out != NULL) {
    ny = out -> ny;
    memcpy(y, out -> y, ny * sizeof(float));
    memcpy(y_sin, out -> y_sin, ny * sizeof(float));
    memcpy(y_noise, out -> y_noise, ny * sizeof(float));
    llsm_delete_output(out);
  }
  llsm_delete_soptions(sopt);
  llsm_delete_chunk(chunk);
  return ny;
}

#include "llsmrt.h"

/* Streaming synthesis of one utterance through llsm_rtsynth_buffer_* (llsmrt.c): srand(seed), create,
   then feed every frame and drain the output ring after each feed. use_l1 != 0 converts the chunk to
   layer 1 first and sets use_l1 (pulse-by-pulse streaming). Returns the number of samples fetched;
   *latency = llsm_rtsynth_buffer_getlatency. */
int ref_rtsynth_soa(int nfrm, float fs, float thop, int maxnhar, int maxnhar_e, int npsd,
  int nchannel, const float* chanfreq, float lip_radius, int use_iczt, int use_l1,
  const float* f0, const int* nhar, const float* ampl, const float* phse,
  const float* psd, const float* psdres, const float* edc, const int* enhar,
  const float* eampl, const float* ephse, unsigned seed,
  float* y_p, float* y_ap, int cap, int* latency, int clear_at, const int* pbpsyn, int remove_hm) {
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nchannel,
    chanfreq, lip_radius, f0, nhar, ampl, phse, psd, psdres, edc, enhar, eampl, ephse);
  llsm_soptions* sopt = llsm_create_soptions(fs);
  sopt -> use_iczt = use_iczt;
  if(use_l1) {
    llsm_chunk_tolayer1(chunk, 2048);
    for(int i = 0; i < nfrm; i ++) {
      if(remove_hm) llsm_container_attach(chunk -> frames[i], LLSM_FRAME_HM, NULL, NULL, NULL);
      if(pbpsyn != NULL && pbpsyn[i])
        llsm_container_attach(chunk -> frames[i], LLSM_FRAME_PBPSYN, llsm_create_int(1), llsm_delete_int, llsm_copy_int);
    }
    sopt -> use_l1 = 1;
  }
  srand(seed);
  llsm_rtsynth_buffer* rt = llsm_create_rtsynth_buffer(sopt, chunk -> conf, 4096);
  *latency = llsm_rtsynth_buffer_getlatency(rt);
  int n = 0;
  for(int i = 0; i < nfrm; i ++) {
    if(i == clear_at) llsm_rtsynth_buffer_clear(rt);
    llsm_rtsynth_buffer_feed(rt, chunk -> frames[i]);
    FP_TYPE p, ap;
    while(llsm_rtsynth_buffer_fetch_decomposed(rt, & p, & ap)) {
      if(n < cap) { y_p[n] = p; y_ap[n] = ap; }
      n ++;
    }
  }
  llsm_delete_rtsynth_buffer(rt);
  llsm_delete_soptions(sopt);
  llsm_delete_chunk(chunk);
  return n < cap ? n : cap;
}

/* timing helper for bench.py: nrep x llsm_analyze on one utterance (x, f0 are copied per run because
   llsm_analyze refines f0 in place). Returns seconds inside llsm_analyze. */
double ref_time_analyze(int nrep, const float* x, int nx, float fs, const float* f0, int nfrm, float thop,
  int maxnhar, int maxnhar_e, int npsd, int nchannel, const float* chanfreq, int hm_method) {
  struct timespec t0, t1;
  double tot = 0;
  llsm_aoptions* opt = llsm_create_aoptions();
  opt -> thop = thop; opt -> maxnhar = maxnhar; opt -> maxnhar_e = maxnhar_e; opt -> npsd = npsd;
  opt -> nchannel = nchannel;
  free(opt -> chanfreq);
  opt -> chanfreq = calloc(nchannel > 1 ? nchannel - 1 : 1, sizeof(FP_TYPE));
  for(int c = 0; c < nchannel - 1; c ++) opt -> chanfreq[c] = chanfreq[c];
  opt -> hm_method = hm_method;
  FP_TYPE* xc = malloc(nx * sizeof(FP_TYPE));
  FP_TYPE* fc = malloc(nfrm * sizeof(FP_TYPE));
  for(int r = 0; r < nrep; r ++) {
    memcpy(xc, x, nx * sizeof(FP_TYPE));
    memcpy(fc, f0, nfrm * sizeof(FP_TYPE));
    clock_gettime(CLOCK_MONOTONIC, & t0);
    llsm_chunk* chunk = llsm_analyze(opt, xc, nx, fs, fc, nfrm, NULL);
    clock_gettime(CLOCK_MONOTONIC, & t1);
    tot += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if(chunk != NULL) llsm_delete_chunk(chunk);
  }
  free(xc); free(fc);
  llsm_delete_aoptions(opt);
  return tot;
}

/* timing helper for bench.py: the chain of test/test-layer0-anasynth.c:40-46 -- llsm_analyze, then llsm_synthesize on the
   chunk it returned -- nrep times on one utterance. t[0] += seconds inside llsm_analyze, t[1] += seconds inside
   llsm_synthesize. Optionally returns the last output (y, ny_cap samples). Returns ny. */
int ref_time_anasynth(int nrep, const float* x, int nx, float fs, const float* f0, int nfrm, float thop,
  int maxnhar, int maxnhar_e, int npsd, int nchannel, const float* chanfreq, int hm_method, double* t,
  float* y, int ny_cap) {
  struct timespec t0, t1;
  int ny = 0;
  llsm_aoptions* opt = llsm_create_aoptions();
  opt -> thop = thop; opt -> maxnhar = maxnhar; opt -> maxnhar_e = maxnhar_e; opt -> npsd = npsd;
  opt -> nchannel = nchannel;
  free(opt -> chanfreq);
  opt -> chanfreq = calloc(nchannel > 1 ? nchannel - 1 : 1, sizeof(FP_TYPE));
  for(int c = 0; c < nchannel - 1; c ++) opt -> chanfreq[c] = chanfreq[c];
  opt -> hm_method = hm_method;
  llsm_soptions* sopt = llsm_create_soptions(fs);
  FP_TYPE* xc = malloc(nx * sizeof(FP_TYPE));
  FP_TYPE* fc = malloc(nfrm * sizeof(FP_TYPE));
  for(int r = 0; r < nrep; r ++) {
    memcpy(xc, x, nx * sizeof(FP_TYPE));
    memcpy(fc, f0, nfrm * sizeof(FP_TYPE));
    clock_gettime(CLOCK_MONOTONIC, & t0);
    llsm_chunk* chunk = llsm_analyze(opt, xc, nx, fs, fc, nfrm, NULL);
    clock_gettime(CLOCK_MONOTONIC, & t1);
    t[0] += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if(chunk == NULL) continue;
    clock_gettime(CLOCK_MONOTONIC, & t0);
    llsm_output* out = llsm_synthesize(sopt, chunk);
    clock_gettime(CLOCK_MONOTONIC, & t1);
    t[1] += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if(out != NULL) {
      ny = out -> ny;
      if(y != NULL) for(int i = 0; i < ny && i < ny_cap; i ++) y[i] = out -> y[i];
      llsm_delete_output(out);
    }
    llsm_delete_chunk(chunk);
  }
  free(xc); free(fc);
  llsm_delete_aoptions(opt);
  llsm_delete_soptions(sopt);
  return ny;
}

/* chunk phase utilities (layer0.c:687-706): op 0 = llsm_chunk_phasepropagate(chunk, arg), op 1 =
   llsm_chunk_phasesync_rps(chunk, arg). phse / ephse (and vsphse when given) are rewritten in place. */
int ref_phase_op_soa(int nfrm, float fs, float thop, int maxnhar, int maxnhar_e, int nchannel,
  const float* f0, const int* nhar, const float* ampl, float* phse, const int* enhar, const float* eampl,
  float* ephse, float* vsphse, const int* nvs, int op, int arg) {
  float cf[7] = {2000, 4000, 8000, 12000, 14000, 16000, 18000};
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, maxnhar_e, 8, nchannel, cf, 0.015f,
    f0, nhar, ampl, phse, NULL, NULL, NULL, enhar, eampl, ephse);
  if(vsphse != NULL)
    for(int i = 0; i < nfrm; i ++) if(nvs[i] > 0) {
      FP_TYPE* vs = llsm_create_fparray(nvs[i]);
      memcpy(vs, vsphse + (size_t)i * maxnhar, nvs[i] * sizeof(float));
      llsm_container_attach(chunk -> frames[i], LLSM_FRAME_VSPHSE, vs, llsm_delete_fparray, llsm_copy_fparray);
    }
  if(op == 0) llsm_chunk_phasepropagate(chunk, arg);
  else llsm_chunk_phasesync_rps(chunk, arg);
  for(int i = 0; i < nfrm; i ++) {
    llsm_hmframe* hm = llsm_container_get(chunk -> frames[i], LLSM_FRAME_HM);
    if(hm != NULL) memcpy(phse + (size_t)i * maxnhar, hm -> phse, hm -> nhar * sizeof(float));
    llsm_nmframe* nm = llsm_container_get(chunk -> frames[i], LLSM_FRAME_NM);
    for(int c = 0; c < nchannel; c ++)
      memcpy(ephse + ((size_t)i * nchannel + c) * maxnhar_e, nm -> eenv[c] -> phse, nm -> eenv[c] -> nhar * sizeof(float));
    FP_TYPE* vs = llsm_container_get(chunk -> frames[i], LLSM_FRAME_VSPHSE);
    if(vs != NULL && vsphse != NULL) memcpy(vsphse + (size_t)i * maxnhar, vs, llsm_fparray_length(vs) * sizeof(float));
  }
  llsm_delete_chunk(chunk);
  return 0;
}

/* ---- coder (coder.c:46-292): frame <-> fixed-dimension vector of order_spec + order_bap + 3 numbers ----
   Encoding reads F0, the noise PSD and (voiced frames) the layer-1 members RD, VTMAGN. */
int ref_coder_encode_soa(int nfrm, float fs, float thop, int maxnhar, int npsd, int nchannel, int maxnhar_e,
  float lip_radius, int nspec, int order_spec, int order_bap,
  const float* f0, const float* psd, const float* rd, const float* vtmagn, float* enc) {
  float cf[7] = {2000, 4000, 8000, 12000, 14000, 16000, 18000};
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nchannel, cf, lip_radius,
    f0, NULL, NULL, NULL, psd, NULL, NULL, NULL, NULL, NULL);
  llsm_container_attach(chunk -> conf, LLSM_CONF_NSPEC, llsm_create_int(nspec), llsm_delete_int, llsm_copy_int);
  llsm_coder* coder = llsm_create_coder(chunk -> conf, order_spec, order_bap);
  int dim = order_spec + order_bap + 3;
  for(int i = 0; i < nfrm; i ++) {
    llsm_container* fr = chunk -> frames[i];
    llsm_container_attach(fr, LLSM_FRAME_RD, llsm_create_fp(rd[i]), llsm_delete_fp, llsm_copy_fp);
    if(f0[i] > 0) {
      FP_TYPE* vt = llsm_create_fparray(nspec);
      memcpy(vt, vtmagn + (size_t)i * nspec, nspec * sizeof(float));
      llsm_container_attach(fr, LLSM_FRAME_VTMAGN, vt, llsm_delete_fparray, llsm_copy_fparray);
    }
    FP_TYPE* e = llsm_coder_encode(coder, fr);
    memcpy(enc + (size_t)i * dim, e, dim * sizeof(float));
    free(e);
  }
  llsm_delete_coder(coder);
  llsm_delete_chunk(chunk);
  return 0;
}

/* Decoding: llsm_coder_decode_layer1 (use_layer1 = 1: RD, VTMAGN, VSPHSE) or _layer0 (HM). Rows of ampl / phse /
   vsphse are maxnhar long; harmonics beyond that are dropped (nhar_out reports the reference's count). */
int ref_coder_decode_soa(int nfrm, float fs, float thop, int maxnhar, int npsd, int nchannel, int maxnhar_e,
  float lip_radius, int nspec, int order_spec, int order_bap, int use_layer1, const float* enc,
  float* f0, float* rd, float* psd, int* nhar_out, float* ampl, float* phse, float* vtmagn, float* vsphse) {
  float cf[7] = {2000, 4000, 8000, 12000, 14000, 16000, 18000};
  llsm_aoptions* opt = llsm_create_aoptions();
  opt -> thop = thop; opt -> maxnhar = maxnhar; opt -> maxnhar_e = maxnhar_e; opt -> npsd = npsd; opt -> nchannel = nchannel;
  free(opt -> chanfreq);
  opt -> chanfreq = calloc(nchannel > 1 ? nchannel - 1 : 1, sizeof(FP_TYPE));
  for(int c = 0; c < nchannel - 1; c ++) opt -> chanfreq[c] = cf[c];
  opt -> lip_radius = lip_radius;
  llsm_container* conf = llsm_aoptions_toconf(opt, fs / 2.0);
  llsm_container_attach(conf, LLSM_CONF_NSPEC, llsm_create_int(nspec), llsm_delete_int, llsm_copy_int);
  llsm_coder* coder = llsm_create_coder(conf, order_spec, order_bap);
  int dim = order_spec + order_bap + 3;
  for(int i = 0; i < nfrm; i ++) {
    FP_TYPE* e = calloc(dim, sizeof(FP_TYPE));
    memcpy(e, enc + (size_t)i * dim, dim * sizeof(float));
    llsm_container* fr = use_layer1 ? llsm_coder_decode_layer1(coder, e) : llsm_coder_decode_layer0(coder, e);
    free(e);
    f0[i] = ((FP_TYPE*)llsm_container_get(fr, LLSM_FRAME_F0))[0];
    rd[i] = ((FP_TYPE*)llsm_container_get(fr, LLSM_FRAME_RD))[0];
    llsm_nmframe* nm = llsm_container_get(fr, LLSM_FRAME_NM);
    memcpy(psd + (size_t)i * npsd, nm -> psd, npsd * sizeof(float));
    nhar_out[i] = 0;
    if(use_layer1) {
      FP_TYPE* vt = llsm_container_get(fr, LLSM_FRAME_VTMAGN);
      FP_TYPE* vs = llsm_container_get(fr, LLSM_FRAME_VSPHSE);
      memset(vtmagn + (size_t)i * nspec, 0, nspec * sizeof(float));
      memset(vsphse + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
      if(vt != NULL) memcpy(vtmagn + (size_t)i * nspec, vt, nspec * sizeof(float));
      if(vs != NULL) {
        int n = llsm_fparray_length(vs);
        nhar_out[i] = n;
        memcpy(vsphse + (size_t)i * maxnhar, vs, (n < maxnhar ? n : maxnhar) * sizeof(float));
      }
    } else {
      llsm_hmframe* hm = llsm_container_get(fr, LLSM_FRAME_HM);
      memset(ampl + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
      memset(phse + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
      if(hm != NULL) {
        int n = hm -> nhar;
        nhar_out[i] = n;
        memcpy(ampl + (size_t)i * maxnhar, hm -> ampl, (n < maxnhar ? n : maxnhar) * sizeof(float));
        memcpy(phse + (size_t)i * maxnhar, hm -> phse, (n < maxnhar ? n : maxnhar) * sizeof(float));
      }
    }
    llsm_delete_container(fr);
  }
  llsm_delete_coder(coder);
  llsm_delete_container(conf);
  llsm_delete_aoptions(opt);
  return 0;
}

/* ---- frame interpolation / time-stretch: the reference keeps it in a demo (test/demo-stretch.c:6-129), as file-static
   functions. The demo is compiled here from where it lies -- its main() renamed and never called, its pitch-tracker /
   wav-file includes satisfied by oracle/ciglet-shim/nebula.h -- so that interp_llsm_frame itself is the oracle. ---- */
#define main ref_demo_stretch_main_unused
#include "test/demo-stretch.c"
#undef main

/* out[i] = interp_llsm_frame(copy(frames[base[i]]), frames[base[i] + 1], ratio[i]) with PSDRES from frames[residx[i]]
   (the loop of test/demo-stretch.c:169-185 for a given map). Rows of the layer-1 arrays of unvoiced frames are zero. */
int ref_stretch_soa(int nfrm, float fs, float thop, int maxnhar, int maxnhar_e, int npsd, int nchannel, float lip_radius,
  int nspec, const float* f0, const float* rd, const float* vtmagn, const float* vsphse, const int* nvs,
  const float* psd, const float* psdres, const float* edc, const int* enhar, const float* eampl, const float* ephse,
  int nfrm_new, const int* base, const float* ratio, const int* residx,
  float* o_f0, float* o_rd, float* o_vtmagn, float* o_vsphse, int* o_nvs,
  float* o_psd, float* o_psdres, float* o_edc, int* o_enhar, float* o_eampl, float* o_ephse) {
  float cf[7] = {2000, 4000, 8000, 12000, 14000, 16000, 18000};
  llsm_chunk* chunk = chunk_from_soa(nfrm, fs, thop, maxnhar, maxnhar_e, npsd, nchannel, cf, lip_radius,
    f0, NULL, NULL, NULL, psd, psdres, edc, enhar, eampl, ephse);
  for(int i = 0; i < nfrm; i ++) {
    llsm_container* fr = chunk -> frames[i];
    llsm_container_attach(fr, LLSM_FRAME_RD, llsm_create_fp(rd[i]), llsm_delete_fp, llsm_copy_fp);
    if(f0[i] > 0) {
      FP_TYPE* vt = llsm_create_fparray(nspec);
      memcpy(vt, vtmagn + (size_t)i * nspec, nspec * sizeof(float));
      llsm_container_attach(fr, LLSM_FRAME_VTMAGN, vt, llsm_delete_fparray, llsm_copy_fparray);
      FP_TYPE* vs = llsm_create_fparray(nvs[i]);
      memcpy(vs, vsphse + (size_t)i * maxnhar, nvs[i] * sizeof(float));
      llsm_container_attach(fr, LLSM_FRAME_VSPHSE, vs, llsm_delete_fparray, llsm_copy_fparray);
    }
  }
  for(int i = 0; i < nfrm_new; i ++) {
    llsm_container* fr = llsm_copy_container(chunk -> frames[base[i]]);
    interp_llsm_frame(fr, chunk -> frames[base[i] + 1], ratio[i]);
    o_f0[i] = ((FP_TYPE*)llsm_container_get(fr, LLSM_FRAME_F0))[0];
    o_rd[i] = ((FP_TYPE*)llsm_container_get(fr, LLSM_FRAME_RD))[0];
    FP_TYPE* vt = llsm_container_get(fr, LLSM_FRAME_VTMAGN);
    FP_TYPE* vs = llsm_container_get(fr, LLSM_FRAME_VSPHSE);
    memset(o_vtmagn + (size_t)i * nspec, 0, nspec * sizeof(float));
    memset(o_vsphse + (size_t)i * maxnhar, 0, maxnhar * sizeof(float));
    o_nvs[i] = 0;
    if(vt != NULL) memcpy(o_vtmagn + (size_t)i * nspec, vt, nspec * sizeof(float));
    if(vs != NULL) {
      o_nvs[i] = llsm_fparray_length(vs);
      memcpy(o_vsphse + (size_t)i * maxnhar, vs, o_nvs[i] * sizeof(float));
    }
    llsm_nmframe* nm = llsm_container_get(fr, LLSM_FRAME_NM);
    memcpy(o_psd + (size_t)i * npsd, nm -> psd, npsd * sizeof(float));
    if(psdres != NULL && o_psdres != NULL) {
      FP_TYPE* res = llsm_container_get(chunk -> frames[residx[i]], LLSM_FRAME_PSDRES);
      memcpy(o_psdres + (size_t)i * npsd, res, npsd * sizeof(float));
    }
    for(int c = 0; c < nchannel; c ++) {
      size_t oc = (size_t)i * nchannel + c;
      o_edc[oc] = nm -> edc[c];
      o_enhar[oc] = nm -> eenv[c] -> nhar;
      memset(o_eampl + oc * maxnhar_e, 0, maxnhar_e * sizeof(float));
      memset(o_ephse + oc * maxnhar_e, 0, maxnhar_e * sizeof(float));
      memcpy(o_eampl + oc * maxnhar_e, nm -> eenv[c] -> ampl, nm -> eenv[c] -> nhar * sizeof(float));
      memcpy(o_ephse + oc * maxnhar_e, nm -> eenv[c] -> phse, nm -> eenv[c] -> nhar * sizeof(float));
    }
    llsm_delete_container(fr);
  }
  llsm_delete_chunk(chunk);
  return 0;
}
