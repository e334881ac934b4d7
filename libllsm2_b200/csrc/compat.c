/*
  Host side of the drop-in API (include/llsm.h): the container / frame / chunk data model in plain
  C, and llsm_analyze / llsm_synthesize as packers around the batched CUDA entry points of
  include/llsm_b200.h. Behavioural contract = the reference's container.c, frame.c and the
  option / chunk helpers of layer0.c (cited per function); written against that contract, not
  copied from it.
*/
#include "../../include/llsm.h"
#include "../../include/llsm_b200.h"
#include <stdlib.h>
#include <pthread.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static __thread char g_compat_err[256] = "";
static void set_err(const char* msg) { snprintf(g_compat_err, sizeof(g_compat_err), "%s", msg); }
const char* llsm_last_error(void) {
  return g_compat_err[0] ? g_compat_err : llsm_b200_last_error();
}

/* ------------------------------------------------------------------ boxed values ---------------- */
/* fparray: an int length lives in the 4 bytes before the returned pointer (container.c:36-40). */
FP_TYPE* llsm_create_fp(FP_TYPE x) { FP_TYPE* p = malloc(sizeof(FP_TYPE)); *p = x; return p; }
int* llsm_create_int(int x) { int* p = malloc(sizeof(int)); *p = x; return p; }
FP_TYPE* llsm_create_fparray(int size) {
  int* raw = calloc(1, sizeof(int) + sizeof(FP_TYPE) * (size_t)(size > 0 ? size : 0));
  raw[0] = size;
  return (FP_TYPE*)(raw + 1);
}
FP_TYPE* llsm_copy_fp(FP_TYPE* src) { return llsm_create_fp(src[0]); }
int* llsm_copy_int(int* src) { return llsm_create_int(src[0]); }
int llsm_fparray_length(FP_TYPE* src) { return ((int*)src)[-1]; }
FP_TYPE* llsm_copy_fparray(FP_TYPE* src) {
  int n = llsm_fparray_length(src);
  FP_TYPE* dst = llsm_create_fparray(n);
  if(n > 0) memcpy(dst, src, sizeof(FP_TYPE) * (size_t)n);
  return dst;
}
void llsm_delete_fp(FP_TYPE* dst) { free(dst); }
void llsm_delete_int(int* dst) { free(dst); }
void llsm_delete_fparray(FP_TYPE* dst) { if(dst != NULL) free((int*)dst - 1); }

/* ------------------------------------------------------------------ container ------------------- */
llsm_container* llsm_create_container(int nmember) {
  llsm_container* c = malloc(sizeof(llsm_container));
  size_t n = (size_t)(nmember > 0 ? nmember : 0);
  c -> nmember = (int)n;
  c -> members = calloc(n ? n : 1, sizeof(void*));
  c -> destructors = calloc(n ? n : 1, sizeof(llsm_fdestructor));
  c -> copyctors = calloc(n ? n : 1, sizeof(llsm_fcopy));
  return c;
}

static void container_grow(llsm_container* c, int n) {
  if(n <= c -> nmember) return;
  c -> members = realloc(c -> members, sizeof(void*) * (size_t)n);
  c -> destructors = realloc(c -> destructors, sizeof(llsm_fdestructor) * (size_t)n);
  c -> copyctors = realloc(c -> copyctors, sizeof(llsm_fcopy) * (size_t)n);
  for(int i = c -> nmember; i < n; i ++) {
    c -> members[i] = NULL; c -> destructors[i] = NULL; c -> copyctors[i] = NULL;
  }
  c -> nmember = n;
}

void* llsm_container_get(llsm_container* src, int index) {
  if(src == NULL || index < 0 || index >= src -> nmember) return NULL;
  return src -> members[index];
}

/* removing calls the member's destructor when one was attached (container.c:148-156) */
void llsm_container_remove(llsm_container* dst, int index) {
  if(index < 0 || index >= dst -> nmember || dst -> members[index] == NULL) return;
  if(dst -> destructors[index] != NULL) dst -> destructors[index](dst -> members[index]);
  dst -> members[index] = NULL;
  dst -> destructors[index] = NULL;
  dst -> copyctors[index] = NULL;
}

/* attach replaces (and destroys) the previous member; attaching NULL is the "remove" idiom
   (container.c:126-146, test/test-llsmrt.c:93) */
void llsm_container_attach_(llsm_container* dst, int index, void* ptr,
  llsm_fdestructor dtor, llsm_fcopy copyctor) {
  container_grow(dst, index + 1);
  llsm_container_remove(dst, index);
  dst -> members[index] = ptr;
  dst -> destructors[index] = dtor;
  dst -> copyctors[index] = copyctor;
}

/* deep copy where a copy constructor exists, aliasing otherwise -- and an aliased member is NOT
   owned by the copy (container.c:82-93; aliasing is pinned by test/test-structs.c:34-35) */
llsm_container* llsm_copy_container(llsm_container* src) {
  llsm_container* c = llsm_create_container(src -> nmember);
  for(int i = 0; i < src -> nmember; i ++) {
    c -> copyctors[i] = src -> copyctors[i];
    if(src -> copyctors[i] != NULL) {
      c -> members[i] = src -> copyctors[i](src -> members[i]);
      c -> destructors[i] = src -> destructors[i];
    } else
      c -> members[i] = src -> members[i];
  }
  return c;
}

/* in-place variant: empties dst first; unlike llsm_copy_container, shallow-copied members keep
   the source's destructor (container.c:95-107) */
void llsm_copy_container_inplace(llsm_container* dst, llsm_container* src) {
  for(int i = 0; i < dst -> nmember; i ++) llsm_container_remove(dst, i);
  for(int i = 0; i < src -> nmember; i ++) {
    if(src -> members[i] == NULL) continue;
    void* m = src -> copyctors[i] != NULL ? src -> copyctors[i](src -> members[i]) : src -> members[i];
    llsm_container_attach_(dst, i, m, src -> destructors[i], src -> copyctors[i]);
  }
}

void llsm_delete_container(llsm_container* dst) {
  if(dst == NULL) return;
  for(int i = 0; i < dst -> nmember; i ++)
    if(dst -> destructors[i] != NULL) dst -> destructors[i](dst -> members[i]);
  free(dst -> members); free(dst -> destructors); free(dst -> copyctors);
  free(dst);
}

/* ------------------------------------------------------------------ hm / nm frames -------------- */
static FP_TYPE wrap_phase(double p) {           /* (-pi, pi] */
  double q = p - 2.0 * M_PI * floor((p + M_PI) / (2.0 * M_PI));
  if(q <= -M_PI) q += 2.0 * M_PI;
  return (FP_TYPE)q;
}

llsm_hmframe* llsm_create_hmframe(int nhar) {
  llsm_hmframe* h = malloc(sizeof(llsm_hmframe));
  size_t n = (size_t)(nhar > 0 ? nhar : 0);
  h -> nhar = nhar;
  h -> ampl = calloc(n ? n : 1, sizeof(FP_TYPE));
  h -> phse = calloc(n ? n : 1, sizeof(FP_TYPE));
  return h;
}

void llsm_copy_hmframe_inplace(llsm_hmframe* dst, llsm_hmframe* src) {
  size_t bytes = sizeof(FP_TYPE) * (size_t)(src -> nhar > 0 ? src -> nhar : 0);
  if(dst -> nhar < src -> nhar) {
    dst -> ampl = realloc(dst -> ampl, bytes);
    dst -> phse = realloc(dst -> phse, bytes);
  }
  if(bytes) { memcpy(dst -> ampl, src -> ampl, bytes); memcpy(dst -> phse, src -> phse, bytes); }
  dst -> nhar = src -> nhar;
}

llsm_hmframe* llsm_copy_hmframe(llsm_hmframe* src) {
  llsm_hmframe* h = llsm_create_hmframe(src -> nhar);
  llsm_copy_hmframe_inplace(h, src);
  return h;
}

void llsm_delete_hmframe(llsm_hmframe* dst) {
  if(dst == NULL) return;
  free(dst -> ampl); free(dst -> phse); free(dst);
}

/* phase of harmonic k (1-based) advances by k * theta, wrapped (frame.c:57-60) */
void llsm_hmframe_phaseshift(llsm_hmframe* dst, FP_TYPE theta) {
  for(int k = 0; k < dst -> nhar; k ++)
    dst -> phse[k] = wrap_phase((FP_TYPE)(dst -> phse[k] + theta * (k + 1.0)));
}

/* noise-equivalent power of each harmonic, a^2 / 2, optionally in dB (frame.c:62-69) */
FP_TYPE* llsm_hmframe_harpsd(llsm_hmframe* src, int db_scale) {
  FP_TYPE* psd = calloc(src -> nhar > 0 ? src -> nhar : 1, sizeof(FP_TYPE));
  for(int k = 0; k < src -> nhar; k ++) {
    psd[k] = src -> ampl[k] * src -> ampl[k] * 0.5;
    if(db_scale) psd[k] = 10.0 * log10(psd[k]);
  }
  return psd;
}

/* defaults: psd -120 dB, edc 1e-5 (frame.c:71-90) */
llsm_nmframe* llsm_create_nmframe(int nchannel, int nhar_e, int npsd) {
  llsm_nmframe* n = malloc(sizeof(llsm_nmframe));
  n -> nchannel = nchannel; n -> npsd = npsd;
  n -> eenv = calloc(nchannel > 0 ? nchannel : 1, sizeof(llsm_hmframe*));
  n -> edc = calloc(nchannel > 0 ? nchannel : 1, sizeof(FP_TYPE));
  n -> psd = calloc(npsd > 0 ? npsd : 1, sizeof(FP_TYPE));
  for(int j = 0; j < npsd; j ++) n -> psd[j] = -120.0;
  for(int c = 0; c < nchannel; c ++) { n -> eenv[c] = llsm_create_hmframe(nhar_e); n -> edc[c] = 1e-5; }
  return n;
}

void llsm_copy_nmframe_inplace(llsm_nmframe* dst, llsm_nmframe* src) {
  if(dst -> npsd < src -> npsd) dst -> psd = realloc(dst -> psd, sizeof(FP_TYPE) * (size_t)src -> npsd);
  memcpy(dst -> psd, src -> psd, sizeof(FP_TYPE) * (size_t)src -> npsd);
  dst -> npsd = src -> npsd;
  if(dst -> nchannel < src -> nchannel) {
    dst -> edc = realloc(dst -> edc, sizeof(FP_TYPE) * (size_t)src -> nchannel);
    dst -> eenv = realloc(dst -> eenv, sizeof(llsm_hmframe*) * (size_t)src -> nchannel);
    for(int c = dst -> nchannel; c < src -> nchannel; c ++) dst -> eenv[c] = llsm_create_hmframe(0);
  } else
    for(int c = src -> nchannel; c < dst -> nchannel; c ++) llsm_delete_hmframe(dst -> eenv[c]);
  for(int c = 0; c < src -> nchannel; c ++) {
    dst -> edc[c] = src -> edc[c];
    llsm_copy_hmframe_inplace(dst -> eenv[c], src -> eenv[c]);
  }
  dst -> nchannel = src -> nchannel;
}

llsm_nmframe* llsm_copy_nmframe(llsm_nmframe* src) {
  llsm_nmframe* n = llsm_create_nmframe(src -> nchannel, 0, src -> npsd);
  llsm_copy_nmframe_inplace(n, src);
  return n;
}

void llsm_delete_nmframe(llsm_nmframe* dst) {
  if(dst == NULL) return;
  for(int c = 0; c < dst -> nchannel; c ++) llsm_delete_hmframe(dst -> eenv[c]);
  free(dst -> eenv); free(dst -> edc); free(dst -> psd); free(dst);
}

/* ------------------------------------------------------------------ effects --------------------- */
llsm_pbpeffect* llsm_create_pbpeffect(llsm_fgfm modifier, void* info) {
  llsm_pbpeffect* e = malloc(sizeof(llsm_pbpeffect));
  e -> modifier = modifier; e -> info = info;
  return e;
}
llsm_pbpeffect* llsm_copy_pbpeffect(llsm_pbpeffect* src) {
  return llsm_create_pbpeffect(src -> modifier, src -> info);
}
void llsm_delete_pbpeffect(llsm_pbpeffect* dst) { free(dst); }

/* ------------------------------------------------------------------ frames ---------------------- */
static void* copy_one_fp(void* p) { return llsm_create_fp(((FP_TYPE*)p)[0]); }

/* a frame = container {F0 = 0, HM(nhar), NM(nchannel, nhar_e, npsd)} (frame.c:137-150) */
llsm_container* llsm_create_frame(int nhar, int nchannel, int nhar_e, int npsd) {
  llsm_container* f = llsm_create_container(3);
  llsm_container_attach(f, LLSM_FRAME_F0, llsm_create_fp(0), free, copy_one_fp);
  llsm_container_attach(f, LLSM_FRAME_HM, llsm_create_hmframe(nhar), llsm_delete_hmframe, llsm_copy_hmframe);
  llsm_container_attach(f, LLSM_FRAME_NM, llsm_create_nmframe(nchannel, nhar_e, npsd),
    llsm_delete_nmframe, llsm_copy_nmframe);
  return f;
}

/* HM, every noise-envelope model and the source phases rotate together (frame.c:152-166) */
void llsm_frame_phaseshift(llsm_container* dst, FP_TYPE theta) {
  llsm_hmframe* hm = llsm_container_get(dst, LLSM_FRAME_HM);
  llsm_nmframe* nm = llsm_container_get(dst, LLSM_FRAME_NM);
  FP_TYPE* vs = llsm_container_get(dst, LLSM_FRAME_VSPHSE);
  if(hm != NULL) llsm_hmframe_phaseshift(hm, theta);
  if(nm != NULL) for(int c = 0; c < nm -> nchannel; c ++) llsm_hmframe_phaseshift(nm -> eenv[c], theta);
  if(vs != NULL) {
    int n = llsm_fparray_length(vs);
    for(int k = 0; k < n; k ++) vs[k] = wrap_phase((FP_TYPE)(vs[k] + theta * (k + 1.0)));
  }
}

/* relative phase shift: subtract the first harmonic's phase (frame.c:168-178) */
void llsm_frame_phasesync_rps(llsm_container* dst, int layer1_based) {
  llsm_hmframe* hm = llsm_container_get(dst, LLSM_FRAME_HM);
  FP_TYPE* vs = llsm_container_get(dst, LLSM_FRAME_VSPHSE);
  FP_TYPE ref = 0;
  if(layer1_based && vs != NULL && llsm_fparray_length(vs) > 0) ref = vs[0];
  else if(hm != NULL && hm -> nhar > 0) ref = hm -> phse[0];
  llsm_frame_phaseshift(dst, -ref);
}

int llsm_frame_checklayer0(llsm_container* src) {      /* frame.c:211-218 */
  FP_TYPE* f0 = llsm_container_get(src, LLSM_FRAME_F0);
  if(f0 == NULL || llsm_container_get(src, LLSM_FRAME_NM) == NULL) return 0;
  if(f0[0] != 0 && llsm_container_get(src, LLSM_FRAME_HM) == NULL) return 0;
  return 1;
}

int llsm_frame_checklayer1(llsm_container* src) {      /* frame.c:220-229 */
  FP_TYPE* f0 = llsm_container_get(src, LLSM_FRAME_F0);
  if(f0 == NULL || llsm_container_get(src, LLSM_FRAME_RD) == NULL ||
     llsm_container_get(src, LLSM_FRAME_NM) == NULL) return 0;
  if(f0[0] > 0 && (llsm_container_get(src, LLSM_FRAME_VTMAGN) == NULL ||
     llsm_container_get(src, LLSM_FRAME_VSPHSE) == NULL)) return 0;
  return 1;
}

int llsm_conf_checklayer0(llsm_container* src) {       /* layer0.c:513-523 */
  const int need[] = {LLSM_CONF_NFRM, LLSM_CONF_THOP, LLSM_CONF_NPSD, LLSM_CONF_FNYQ,
    LLSM_CONF_NCHANNEL, LLSM_CONF_CHANFREQ};
  for(size_t i = 0; i < sizeof(need) / sizeof(need[0]); i ++)
    if(llsm_container_get(src, need[i]) == NULL) return 0;
  return 1;
}

int llsm_conf_checklayer1(llsm_container* src) {
  return llsm_conf_checklayer0(src) && llsm_container_get(src, LLSM_CONF_NSPEC) != NULL &&
    llsm_container_get(src, LLSM_CONF_LIPRADIUS) != NULL;
}

/* needs the deprecated NOSWARP entry that llsm_aoptions_toconf never creates (frame.c:186-188):
   always NULL for configurations made by this API, as in the reference */
FP_TYPE* llsm_frame_compute_snr(llsm_container* src, llsm_container* conf, int as_aperiodicity) {
  (void)src; (void)conf; (void)as_aperiodicity;
  return NULL;
}

/* ------------------------------------------------------------------ layer 1 --------------------- */
static llsm_b200_ctx* shared_ctx(void);

static int conf_to_b200(llsm_container* conf, llsm_b200_conf* c, float fs_or_zero) {
  memset(c, 0, sizeof(*c));
  int* nfrm = llsm_container_get(conf, LLSM_CONF_NFRM);
  int* npsd = llsm_container_get(conf, LLSM_CONF_NPSD);
  int* nch = llsm_container_get(conf, LLSM_CONF_NCHANNEL);
  FP_TYPE* thop = llsm_container_get(conf, LLSM_CONF_THOP);
  FP_TYPE* fnyq = llsm_container_get(conf, LLSM_CONF_FNYQ);
  FP_TYPE* lip = llsm_container_get(conf, LLSM_CONF_LIPRADIUS);
  FP_TYPE* cf = llsm_container_get(conf, LLSM_CONF_CHANFREQ);
  int* mne = llsm_container_get(conf, LLSM_CONF_MAXNHAR_E);
  if(nfrm == NULL || thop == NULL || fnyq == NULL) return 0;
  c -> nutt = 1; c -> nfrm = *nfrm; c -> npsd = npsd != NULL ? *npsd : 2; c -> nchannel = nch != NULL ? *nch : 1;
  c -> thop = *thop; c -> fs = fs_or_zero > 0 ? fs_or_zero : (float)(*fnyq * 2.0);
  c -> lip_radius = lip != NULL ? *lip : 1.5f;
  c -> maxnhar = 1; c -> maxnhar_e = mne != NULL ? *mne : 0;
  if(c -> nchannel < 1 || c -> nchannel > LLSM_B200_MAXCHANNEL) return 0;
  if(cf != NULL) for(int i = 0; i < c -> nchannel - 1; i ++) c -> chanfreq[i] = cf[i];
  return 1;
}

/* L0 -> L1 for a whole chunk (layer1.c:129-149): attaches NSPEC to the conf, RD to every frame,
   VTMAGN / VSPHSE to the voiced ones. Returns silently on failure, like the reference. */
void llsm_chunk_tolayer1(llsm_chunk* dst, int nfft) {
  g_compat_err[0] = 0;
  llsm_b200_conf c;
  if(dst == NULL || ! conf_to_b200(dst -> conf, & c, 0) ||
     llsm_container_get(dst -> conf, LLSM_CONF_LIPRADIUS) == NULL) return;
  for(int i = 0; i < c.nfrm; i ++) if(! llsm_frame_checklayer0(dst -> frames[i])) return;   /* layer1.c:34-36 */
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return;
  int maxnhar = 1;
  for(int i = 0; i < c.nfrm; i ++) {
    llsm_hmframe* hm = llsm_container_get(dst -> frames[i], LLSM_FRAME_HM);
    if(hm != NULL && hm -> nhar > maxnhar) maxnhar = hm -> nhar;
  }
  c.maxnhar = maxnhar;
  const size_t F = (size_t)c.nfrm; const int nspec = nfft / 2 + 1;
  float* f0 = calloc(F, 4); int* nhar = calloc(F, 4); float* ampl = calloc(F * maxnhar, 4); float* phse = calloc(F * maxnhar, 4);
  float* rd = calloc(F, 4); float* vt = calloc(F * nspec, 4); float* vs = calloc(F * maxnhar, 4); int* nvs = calloc(F, 4);
  for(int i = 0; i < c.nfrm; i ++) {
    FP_TYPE* pf0 = llsm_container_get(dst -> frames[i], LLSM_FRAME_F0);
    llsm_hmframe* hm = llsm_container_get(dst -> frames[i], LLSM_FRAME_HM);
    f0[i] = pf0[0];
    if(pf0[0] != 0 && hm != NULL) {
      nhar[i] = hm -> nhar;
      memcpy(ampl + (size_t)i * maxnhar, hm -> ampl, 4 * (size_t)hm -> nhar);
      memcpy(phse + (size_t)i * maxnhar, hm -> phse, 4 * (size_t)hm -> nhar);
    }
  }
  llsm_b200_frames fr; memset(& fr, 0, sizeof(fr));
  fr.f0 = f0; fr.nhar = nhar; fr.ampl = ampl; fr.phse = phse;
  llsm_b200_layer1 l1 = {rd, vt, vs, nvs, nspec};
  if(llsm_b200_tolayer1_host(ctx, & c, & fr, nfft, & l1) == 0) {
    llsm_container_attach(dst -> conf, LLSM_CONF_NSPEC, llsm_create_int(nspec), llsm_delete_int, llsm_copy_int);
    for(int i = 0; i < c.nfrm; i ++) {
      llsm_container* f = dst -> frames[i];
      llsm_container_attach(f, LLSM_FRAME_RD, llsm_create_fp(rd[i]), llsm_delete_fp, llsm_copy_fp);
      if(f0[i] == 0) continue;
      FP_TYPE* a = llsm_create_fparray(nspec);
      memcpy(a, vt + (size_t)i * nspec, 4 * (size_t)nspec);
      FP_TYPE* p = llsm_create_fparray(nvs[i]);
      memcpy(p, vs + (size_t)i * maxnhar, 4 * (size_t)nvs[i]);
      llsm_container_attach(f, LLSM_FRAME_VTMAGN, a, llsm_delete_fparray, llsm_copy_fparray);
      llsm_container_attach(f, LLSM_FRAME_VSPHSE, p, llsm_delete_fparray, llsm_copy_fparray);
    }
  }
  free(f0); free(nhar); free(ampl); free(phse); free(rd); free(vt); free(vs); free(nvs);
}

/* L1 -> L0 for `n` frames sharing `conf` (layer1.c:151-195): attaches an HM member to every voiced
   frame that passes the layer-1 check. */
static void frames_tolayer0(llsm_container** frames, int n, llsm_container* conf) {
  FP_TYPE* fnyq = llsm_container_get(conf, LLSM_CONF_FNYQ);
  FP_TYPE* lip = llsm_container_get(conf, LLSM_CONF_LIPRADIUS);
  int* nspec = llsm_container_get(conf, LLSM_CONF_NSPEC);
  int* cap = llsm_container_get(conf, LLSM_CONF_MAXNHAR);
  if(fnyq == NULL || lip == NULL || nspec == NULL || n < 1) return;    /* layer1.c:40-46 */
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return;
  int maxlen = 1;
  for(int i = 0; i < n; i ++) {
    FP_TYPE* vs = llsm_container_get(frames[i], LLSM_FRAME_VSPHSE);
    if(vs != NULL && llsm_fparray_length(vs) > maxlen) maxlen = llsm_fparray_length(vs);
  }
  int maxnhar = maxlen;
  if(cap != NULL && *cap < maxnhar) maxnhar = *cap > 0 ? *cap : 1;     /* rows only need min(len, MAXNHAR) */
  llsm_b200_conf c; memset(& c, 0, sizeof(c));
  c.nutt = 1; c.nfrm = n; c.maxnhar = maxnhar; c.maxnhar_e = 0; c.npsd = 2; c.nchannel = 1;
  c.fs = (float)(*fnyq * 2.0); c.thop = 0.005f; c.lip_radius = *lip;
  const size_t F = (size_t)n;
  float* f0 = calloc(F, 4); float* rd = calloc(F, 4); float* vt = calloc(F * *nspec, 4); float* vsb = calloc(F * maxnhar, 4);
  int* nvs = calloc(F, 4); int* nhar = calloc(F, 4); float* ampl = calloc(F * maxnhar, 4); float* phse = calloc(F * maxnhar, 4);
  for(int i = 0; i < n; i ++) {
    if(! llsm_frame_checklayer1(frames[i])) continue;
    FP_TYPE* pf0 = llsm_container_get(frames[i], LLSM_FRAME_F0);
    FP_TYPE* prd = llsm_container_get(frames[i], LLSM_FRAME_RD);
    FP_TYPE* pvt = llsm_container_get(frames[i], LLSM_FRAME_VTMAGN);
    FP_TYPE* pvs = llsm_container_get(frames[i], LLSM_FRAME_VSPHSE);
    if(pf0[0] == 0) continue;
    f0[i] = pf0[0]; rd[i] = prd[0];
    memcpy(vt + (size_t)i * *nspec, pvt, 4 * (size_t)*nspec);
    int len = llsm_fparray_length(pvs);
    nvs[i] = len;                                           /* the kernel applies min(len, MAXNHAR, fnyq / f0) */
    memcpy(vsb + (size_t)i * maxnhar, pvs, 4 * (size_t)(len < maxnhar ? len : maxnhar));
  }
  llsm_b200_layer1 l1 = {rd, vt, vsb, nvs, *nspec};
  if(llsm_b200_tolayer0_host(ctx, & c, NULL, f0, & l1, nhar, ampl, phse) == 0) {
    for(int i = 0; i < n; i ++) {
      if(f0[i] == 0) continue;
      llsm_hmframe* hm = llsm_create_hmframe(nhar[i]);
      memcpy(hm -> ampl, ampl + (size_t)i * maxnhar, 4 * (size_t)nhar[i]);
      memcpy(hm -> phse, phse + (size_t)i * maxnhar, 4 * (size_t)nhar[i]);
      llsm_container_attach(frames[i], LLSM_FRAME_HM, hm, llsm_delete_hmframe, llsm_copy_hmframe);
    }
  }
  free(f0); free(rd); free(vt); free(vsb); free(nvs); free(nhar); free(ampl); free(phse);
}

void llsm_frame_tolayer0(llsm_container* dst, llsm_container* conf) {
  g_compat_err[0] = 0;
  if(dst == NULL || conf == NULL) return;
  frames_tolayer0(& dst, 1, conf);
}

void llsm_chunk_tolayer0(llsm_chunk* dst) {
  g_compat_err[0] = 0;
  int* nfrm = llsm_container_get(dst -> conf, LLSM_CONF_NFRM);
  if(nfrm == NULL) return;
  frames_tolayer0(dst -> frames, *nfrm, dst -> conf);
}

/* ------------------------------------------------------------------ coder ----------------------- */
/* llsm_create_coder .. llsm_coder_decode_layer{0,1} (coder.c:46-292): one frame per call through the batched device
   coder (llsm_b200_coder_*_host). NULL when no CUDA device is usable. */
typedef struct { llsm_b200_conf c; int order_spec, order_bap, nspec; } compat_coder;

llsm_coder* llsm_create_coder(llsm_container* conf, int order_spec, int order_bap) {
  g_compat_err[0] = 0;
  if(conf == NULL) return NULL;
  FP_TYPE* fnyq = llsm_container_get(conf, LLSM_CONF_FNYQ);
  int* nch = llsm_container_get(conf, LLSM_CONF_NCHANNEL);
  int* mne = llsm_container_get(conf, LLSM_CONF_MAXNHAR_E);
  int* npsd = llsm_container_get(conf, LLSM_CONF_NPSD);
  int* nspec = llsm_container_get(conf, LLSM_CONF_NSPEC);
  FP_TYPE* lip = llsm_container_get(conf, LLSM_CONF_LIPRADIUS);
  if(fnyq == NULL || nch == NULL || mne == NULL || npsd == NULL || nspec == NULL || lip == NULL) return NULL;
  compat_coder* r = calloc(1, sizeof(compat_coder));
  r -> c.nutt = 1; r -> c.nfrm = 1; r -> c.maxnhar = 1; r -> c.maxnhar_e = *mne; r -> c.npsd = *npsd;
  r -> c.nchannel = *nch; r -> c.fs = (float)(*fnyq * 2.0); r -> c.thop = 0.005f; r -> c.lip_radius = *lip;
  r -> order_spec = order_spec; r -> order_bap = order_bap; r -> nspec = *nspec;
  return r;
}
void llsm_delete_coder(llsm_coder* dst) { free(dst); }

FP_TYPE* llsm_coder_encode(llsm_coder* c_, llsm_container* src) {
  g_compat_err[0] = 0;
  compat_coder* c = (compat_coder*)c_;
  if(c == NULL || src == NULL) return NULL;
  FP_TYPE* f0 = llsm_container_get(src, LLSM_FRAME_F0);
  llsm_nmframe* nm = llsm_container_get(src, LLSM_FRAME_NM);
  FP_TYPE* rd = llsm_container_get(src, LLSM_FRAME_RD);
  FP_TYPE* vt = llsm_container_get(src, LLSM_FRAME_VTMAGN);
  if(f0 == NULL || nm == NULL || nm -> npsd != c -> c.npsd) return NULL;
  if(f0[0] > 0 && (rd == NULL || vt == NULL)) return NULL;
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return NULL;
  float rdv = rd != NULL ? rd[0] : 0;
  float* vtz = NULL;
  if(vt == NULL) { vtz = calloc(c -> nspec, 4); vt = vtz; }
  llsm_b200_layer1 l1 = {& rdv, vt, NULL, NULL, c -> nspec};
  FP_TYPE* enc = calloc(c -> order_spec + c -> order_bap + 3, sizeof(FP_TYPE));
  int rc = llsm_b200_coder_encode_host(ctx, & c -> c, NULL, f0, nm -> psd, & l1, c -> order_spec, c -> order_bap, enc);
  free(vtz);
  if(rc != 0) { snprintf(g_compat_err, sizeof(g_compat_err), "%s", llsm_b200_last_error()); free(enc); return NULL; }
  return enc;
}

static llsm_container* coder_decode(compat_coder* c, FP_TYPE* src, int use_layer1) {
  g_compat_err[0] = 0;
  if(c == NULL || src == NULL) return NULL;
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return NULL;
  const float fnyq = (float)((double)c -> c.fs / 2.0);
  const int voicing = src[0] > 0.5;
  const float f0c = (float)(src[1] > 20.0 ? (double)src[1] : 20.0);           /* coder.c:170 */
  const int nhar = voicing ? (int)(fnyq / f0c) : 0;                           /* coder.c:172 */
  llsm_b200_conf cc = c -> c; cc.maxnhar = nhar > 0 ? nhar : 1;
  float f0 = 0, rd = 0; int nh = 0;
  float* psd = calloc(cc.npsd, 4); float* a = calloc(cc.maxnhar, 4); float* p = calloc(cc.maxnhar, 4);
  float* vt = calloc(c -> nspec, 4);
  llsm_b200_frames_out o; memset(& o, 0, sizeof(o));
  o.f0 = & f0; o.psd = psd; o.nhar = & nh; o.ampl = a; o.phse = p;
  llsm_b200_layer1 l1 = {& rd, vt, p, NULL, c -> nspec};
  llsm_container* ret = NULL;
  if(llsm_b200_coder_decode_host(ctx, & cc, NULL, src, c -> order_spec, c -> order_bap, use_layer1, & o, & l1) == 0) {
    ret = llsm_create_frame(nhar, cc.nchannel, cc.maxnhar_e, cc.npsd);
    llsm_nmframe* nm = llsm_container_get(ret, LLSM_FRAME_NM);
    memcpy(nm -> psd, psd, 4 * (size_t)cc.npsd);
    llsm_container_attach(ret, LLSM_FRAME_RD, llsm_create_fp(rd), llsm_delete_fp, llsm_copy_fp);
    llsm_container_attach(ret, LLSM_FRAME_F0, llsm_create_fp(f0), llsm_delete_fp, llsm_copy_fp);
    if(nhar > 0 && use_layer1) {
      llsm_container_remove(ret, LLSM_FRAME_HM);
      FP_TYPE* v = llsm_create_fparray(c -> nspec); memcpy(v, vt, 4 * (size_t)c -> nspec);
      FP_TYPE* s = llsm_create_fparray(nhar); memcpy(s, p, 4 * (size_t)nhar);
      llsm_container_attach(ret, LLSM_FRAME_VTMAGN, v, llsm_delete_fparray, llsm_copy_fparray);
      llsm_container_attach(ret, LLSM_FRAME_VSPHSE, s, llsm_delete_fparray, llsm_copy_fparray);
    }
    if(nhar > 0 && ! use_layer1) {
      llsm_hmframe* hm = llsm_create_hmframe(nhar);
      memcpy(hm -> ampl, a, 4 * (size_t)nhar); memcpy(hm -> phse, p, 4 * (size_t)nhar);
      llsm_container_attach(ret, LLSM_FRAME_HM, hm, llsm_delete_hmframe, llsm_copy_hmframe);
    }
  } else snprintf(g_compat_err, sizeof(g_compat_err), "%s", llsm_b200_last_error());
  free(psd); free(a); free(p); free(vt);
  return ret;
}
llsm_container* llsm_coder_decode_layer1(llsm_coder* c, FP_TYPE* src) { return coder_decode((compat_coder*)c, src, 1); }
llsm_container* llsm_coder_decode_layer0(llsm_coder* c, FP_TYPE* src) { return coder_decode((compat_coder*)c, src, 0); }

/* ------------------------------------------------------------------ options --------------------- */
llsm_aoptions* llsm_create_aoptions(void) {            /* defaults of layer0.c:27-43 */
  llsm_aoptions* o = malloc(sizeof(llsm_aoptions));
  o -> thop = 0.005; o -> maxnhar = 100; o -> maxnhar_e = 4; o -> npsd = 256; o -> nchannel = 4;
  o -> chanfreq = calloc(3, sizeof(FP_TYPE));
  o -> chanfreq[0] = 2000.0; o -> chanfreq[1] = 4000.0; o -> chanfreq[2] = 8000.0;
  o -> lip_radius = 1.5; o -> f0_refine = 1; o -> hm_method = LLSM_AOPTION_HMCZT; o -> rel_winsize = 4.0;
  return o;
}

void llsm_delete_aoptions(llsm_aoptions* dst) {
  if(dst == NULL) return;
  free(dst -> chanfreq); free(dst);
}

llsm_container* llsm_aoptions_toconf(llsm_aoptions* src, FP_TYPE fnyq) {   /* layer0.c:51-76 */
  llsm_container* c = llsm_create_container(10);
  llsm_container_attach(c, LLSM_CONF_NFRM, llsm_create_int(0), llsm_delete_int, llsm_copy_int);
  llsm_container_attach(c, LLSM_CONF_THOP, llsm_create_fp(src -> thop), llsm_delete_fp, llsm_copy_fp);
  llsm_container_attach(c, LLSM_CONF_MAXNHAR, llsm_create_int(src -> maxnhar), llsm_delete_int, llsm_copy_int);
  llsm_container_attach(c, LLSM_CONF_MAXNHAR_E, llsm_create_int(src -> maxnhar_e), llsm_delete_int, llsm_copy_int);
  llsm_container_attach(c, LLSM_CONF_NPSD, llsm_create_int(src -> npsd), llsm_delete_int, llsm_copy_int);
  llsm_container_attach(c, LLSM_CONF_FNYQ, llsm_create_fp(fnyq), llsm_delete_fp, llsm_copy_fp);
  llsm_container_attach(c, LLSM_CONF_NCHANNEL, llsm_create_int(src -> nchannel), llsm_delete_int, llsm_copy_int);
  llsm_container_attach(c, LLSM_CONF_LIPRADIUS, llsm_create_fp(src -> lip_radius), llsm_delete_fp, llsm_copy_fp);
  FP_TYPE* cf = llsm_create_fparray(src -> nchannel - 1);
  for(int i = 0; i < src -> nchannel - 1; i ++) cf[i] = src -> chanfreq[i];
  llsm_container_attach(c, LLSM_CONF_CHANFREQ, cf, llsm_delete_fparray, llsm_copy_fparray);
  return c;
}

llsm_soptions* llsm_create_soptions(FP_TYPE fs) {      /* layer0.c:78-87 */
  llsm_soptions* o = malloc(sizeof(llsm_soptions));
  o -> fs = fs; o -> use_iczt = 1; o -> use_l1 = 0; o -> iczt_param_a = 0.275; o -> iczt_param_b = 2.26;
  return o;
}
void llsm_delete_soptions(llsm_soptions* dst) { free(dst); }

void llsm_delete_output(llsm_output* dst) {
  if(dst == NULL) return;
  free(dst -> y); free(dst -> y_sin); free(dst -> y_noise); free(dst);
}

/* ------------------------------------------------------------------ chunks ---------------------- */
llsm_chunk* llsm_create_chunk(llsm_container* conf, int init_frames) {      /* container.c:158-174 */
  int* nfrm = llsm_container_get(conf, LLSM_CONF_NFRM);
  int* nchannel = llsm_container_get(conf, LLSM_CONF_NCHANNEL);
  int* npsd = llsm_container_get(conf, LLSM_CONF_NPSD);
  if(nchannel == NULL || npsd == NULL) return NULL;
  llsm_chunk* ck = malloc(sizeof(llsm_chunk));
  ck -> conf = llsm_copy_container(conf);
  ck -> frames = NULL;
  if(nfrm != NULL) {
    ck -> frames = calloc(*nfrm > 0 ? *nfrm : 1, sizeof(llsm_container*));
    if(init_frames)
      for(int i = 0; i < *nfrm; i ++) ck -> frames[i] = llsm_create_frame(0, *nchannel, 0, *npsd);
  }
  return ck;
}

llsm_chunk* llsm_copy_chunk(llsm_chunk* src) {
  llsm_chunk* ck = llsm_create_chunk(src -> conf, 0);
  int* nfrm = llsm_container_get(src -> conf, LLSM_CONF_NFRM);
  if(ck != NULL && nfrm != NULL)
    for(int i = 0; i < *nfrm; i ++) ck -> frames[i] = llsm_copy_container(src -> frames[i]);
  return ck;
}

void llsm_delete_chunk(llsm_chunk* dst) {
  if(dst == NULL) return;
  int* nfrm = llsm_container_get(dst -> conf, LLSM_CONF_NFRM);
  if(nfrm != NULL && dst -> frames != NULL)
    for(int i = 0; i < *nfrm; i ++) llsm_delete_container(dst -> frames[i]);
  llsm_delete_container(dst -> conf);
  free(dst -> frames);
  free(dst);
}

FP_TYPE* llsm_chunk_getf0(llsm_chunk* src, int* dst_nfrm) {                 /* layer0.c:674-685 */
  int* nfrm = llsm_container_get(src -> conf, LLSM_CONF_NFRM);
  if(nfrm == NULL) return NULL;
  FP_TYPE* f0 = calloc(*nfrm > 0 ? *nfrm : 1, sizeof(FP_TYPE));
  *dst_nfrm = *nfrm;
  for(int i = 0; i < *nfrm; i ++) {
    FP_TYPE* v = llsm_container_get(src -> frames[i], LLSM_FRAME_F0);
    if(v != NULL) f0[i] = v[0];
  }
  return f0;
}

void llsm_chunk_phasesync_rps(llsm_chunk* dst, int layer1_based) {          /* layer0.c:687-692 */
  int* nfrm = llsm_container_get(dst -> conf, LLSM_CONF_NFRM);
  if(nfrm == NULL) return;
  for(int i = 0; i < *nfrm; i ++) llsm_frame_phasesync_rps(dst -> frames[i], layer1_based);
}

/* add (sign = +1) or remove (-1) the running integral of F0, 2 pi thop cumsum(f0) (layer0.c:694-706;
   the cumulative sum is accumulated in double and stored as FP_TYPE like the oracle's cumsum) */
void llsm_chunk_phasepropagate(llsm_chunk* dst, int sign) {
  int nfrm = 0;
  FP_TYPE* f0 = llsm_chunk_getf0(dst, & nfrm);
  FP_TYPE* thop = llsm_container_get(dst -> conf, LLSM_CONF_THOP);
  if(thop == NULL || f0 == NULL) { free(f0); return; }
  double acc = 0;
  for(int i = 0; i < nfrm; i ++) {
    acc += f0[i];
    FP_TYPE d = (FP_TYPE)acc;
    d = (FP_TYPE)(d * (*thop * sign * 2.0 * M_PI));
    llsm_frame_phaseshift(dst -> frames[i], d);
  }
  free(f0);
}

/* ------------------------------------------------------------------ device context -------------- */
static llsm_b200_ctx* g_ctx = NULL;
static pthread_once_t g_ctx_once = PTHREAD_ONCE_INIT;
static void shared_ctx_init(void) {
  const char* e = getenv("LLSM_B200_DEVICE");
  g_ctx = llsm_b200_create(e != NULL ? atoi(e) : 0);
}
/* one process-wide device context, created once however many threads make the first call; the context itself
   serialises the calls that go through it (the reference's functions are reentrant: INTEGRATION.md section 4) */
static llsm_b200_ctx* shared_ctx(void) {
  pthread_once(&g_ctx_once, shared_ctx_init);
  return g_ctx;
}

/* N(0, var) from libc rand(): one Box-Muller cosine branch, two rand() draws per value -- the draw
   sequence of the reference's llsm_generate_white_noise (dsputils.c:353-361) under the oracle's
   randn(); keeps y_noise reproducible against the CPU reference for a given srand() state. */
static FP_TYPE host_randn(void) {
  double u1 = ((double)rand() + 1.0) / ((double)RAND_MAX + 2.0);
  double u2 = ((double)rand() + 1.0) / ((double)RAND_MAX + 2.0);
  return (FP_TYPE)(0.0 + sqrt(1.0) * (sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2)));
}

static int same_conf(llsm_container* a, llsm_container* b) {
  const int ints[] = {LLSM_CONF_NFRM, LLSM_CONF_NPSD, LLSM_CONF_NCHANNEL};
  for(int i = 0; i < 3; i ++)
    if(*(int*)llsm_container_get(a, ints[i]) != *(int*)llsm_container_get(b, ints[i])) return 0;
  if(*(FP_TYPE*)llsm_container_get(a, LLSM_CONF_THOP) != *(FP_TYPE*)llsm_container_get(b, LLSM_CONF_THOP)) return 0;
  FP_TYPE* ca = llsm_container_get(a, LLSM_CONF_CHANFREQ); FP_TYPE* cb = llsm_container_get(b, LLSM_CONF_CHANFREQ);
  int nch = *(int*)llsm_container_get(a, LLSM_CONF_NCHANNEL);
  for(int c = 0; c < nch - 1; c ++) if(ca[c] != cb[c]) return 0;
  return 1;
}

static int chunk_ok(llsm_chunk* src) {                 /* layer0.c:525-533 */
  if(src == NULL || ! llsm_conf_checklayer0(src -> conf)) return 0;
  int nfrm = *(int*)llsm_container_get(src -> conf, LLSM_CONF_NFRM);
  for(int i = 0; i < nfrm; i ++)
    if(! llsm_frame_checklayer0(src -> frames[i]) && ! llsm_frame_checklayer1(src -> frames[i])) return 0;
  return 1;
}

static FP_TYPE host_randn(void);
/* ------------------------------------------------------------------ layer-1 synthesis ----------- */
/* adapter between the C ABI's per-pulse hook and llsm_pbpeffect (llsm.h:190-197) */
typedef struct { llsm_chunk** chunks; } hook_env;
static int pulse_hook(void* user, int utt, int frame, float* Fa, float* Rk, float* Rg, float* T0, float* Ee,
  float* delta_t) {
  hook_env* env = user;
  llsm_container* fr = env -> chunks[utt] -> frames[frame];
  llsm_pbpeffect* eff = llsm_container_get(fr, LLSM_FRAME_PBPEFF);
  if(eff == NULL) return 0;
  llsm_gfm g = {*Fa, *Rk, *Rg, *T0, *Ee};
  eff -> modifier(& g, delta_t, eff -> info, fr);                   /* layer0.c:212 */
  *Fa = g.Fa; *Rk = g.Rk; *Rg = g.Rg; *T0 = g.T0; *Ee = g.Ee;
  return 1;
}

static int synthesize_l1_batch(llsm_b200_ctx* ctx, llsm_soptions* options, llsm_chunk** src, int n,
  llsm_output** dst) {
  llsm_container* conf = src[0] -> conf;
  llsm_b200_conf c;
  if(! conf_to_b200(conf, & c, options -> fs)) { set_err("unsupported configuration"); return -1; }
  c.nutt = n;
  FP_TYPE* cf = llsm_container_get(conf, LLSM_CONF_CHANFREQ);
  /* frames that will need a harmonic model but carry none: derive it now (layer0.c:264-265) */
  int any_eff = 0, nspec = 0;
  for(int b = 0; b < n; b ++) {
    int cnt = 0;
    llsm_container** todo = calloc(c.nfrm > 0 ? c.nfrm : 1, sizeof(llsm_container*));
    for(int i = 0; i < c.nfrm; i ++) {
      llsm_container* f = src[b] -> frames[i];
      FP_TYPE* pf0 = llsm_container_get(f, LLSM_FRAME_F0);
      FP_TYPE* vt = llsm_container_get(f, LLSM_FRAME_VTMAGN);
      if(llsm_container_get(f, LLSM_FRAME_PBPEFF) != NULL) any_eff = 1;
      if(vt != NULL && llsm_fparray_length(vt) > nspec) nspec = llsm_fparray_length(vt);
      if(pf0[0] != 0 && llsm_frame_checklayer1(f) && llsm_container_get(f, LLSM_FRAME_HM) == NULL) todo[cnt ++] = f;
    }
    if(cnt > 0) frames_tolayer0(todo, cnt, src[b] -> conf);
    free(todo);
  }
  if(nspec < 2) { int* ns = llsm_container_get(conf, LLSM_CONF_NSPEC); nspec = ns != NULL ? *ns : 2; }
  int maxnhar = 1, maxnhar_e = 1, has_res = 0;
  for(int b = 0; b < n; b ++) for(int i = 0; i < c.nfrm; i ++) {
    llsm_container* f = src[b] -> frames[i];
    llsm_hmframe* hm = llsm_container_get(f, LLSM_FRAME_HM);
    llsm_nmframe* nm = llsm_container_get(f, LLSM_FRAME_NM);
    FP_TYPE* vs = llsm_container_get(f, LLSM_FRAME_VSPHSE);
    if(hm != NULL && hm -> nhar > maxnhar) maxnhar = hm -> nhar;
    if(vs != NULL && llsm_fparray_length(vs) > maxnhar) maxnhar = llsm_fparray_length(vs);
    if(nm -> nchannel != c.nchannel || nm -> npsd != c.npsd) { set_err("frame / conf size mismatch"); return -1; }
    for(int ch = 0; ch < c.nchannel; ch ++) if(nm -> eenv[ch] -> nhar > maxnhar_e) maxnhar_e = nm -> eenv[ch] -> nhar;
    if(llsm_container_get(f, LLSM_FRAME_PSDRES) != NULL) has_res = 1;
  }
  if(maxnhar > 2048) maxnhar = 2048;
  c.maxnhar = maxnhar; c.maxnhar_e = maxnhar_e;

  const size_t BF = (size_t)n * c.nfrm;
  float* f0 = calloc(BF, 4); int* nhar = calloc(BF, 4);
  float* ampl = calloc(BF * maxnhar, 4); float* phse = calloc(BF * maxnhar, 4);
  float* psd = calloc(BF * c.npsd, 4); float* psdres = has_res ? calloc(BF * c.npsd, 4) : NULL;
  float* edc = calloc(BF * c.nchannel, 4); int* enhar = calloc(BF * c.nchannel, 4);
  float* eampl = calloc(BF * c.nchannel * maxnhar_e, 4); float* ephse = calloc(BF * c.nchannel * maxnhar_e, 4);
  float* rd = calloc(BF, 4); float* vtm = calloc(BF * nspec, 4); float* vsp = calloc(BF * maxnhar, 4);
  int* nvs = calloc(BF, 4); int* pbp = calloc(BF, 4);
  const double resbias = 0.375 / 2.3025851 * 10.0;
  for(int b = 0; b < n; b ++) for(int i = 0; i < c.nfrm; i ++) {
    size_t r = (size_t)b * c.nfrm + i;
    llsm_container* f = src[b] -> frames[i];
    FP_TYPE* pf0 = llsm_container_get(f, LLSM_FRAME_F0);
    llsm_hmframe* hm = llsm_container_get(f, LLSM_FRAME_HM);
    llsm_nmframe* nm = llsm_container_get(f, LLSM_FRAME_NM);
    FP_TYPE* res = llsm_container_get(f, LLSM_FRAME_PSDRES);
    FP_TYPE* prd = llsm_container_get(f, LLSM_FRAME_RD);
    FP_TYPE* pvt = llsm_container_get(f, LLSM_FRAME_VTMAGN);
    FP_TYPE* pvs = llsm_container_get(f, LLSM_FRAME_VSPHSE);
    int* psyn = llsm_container_get(f, LLSM_FRAME_PBPSYN);
    f0[r] = pf0[0];
    if(pf0[0] != 0 && hm != NULL) {
      int k = hm -> nhar < maxnhar ? hm -> nhar : maxnhar;
      nhar[r] = k;
      memcpy(ampl + r * maxnhar, hm -> ampl, 4 * (size_t)k);
      memcpy(phse + r * maxnhar, hm -> phse, 4 * (size_t)k);
    }
    if(pf0[0] != 0 && prd != NULL && pvt != NULL && pvs != NULL && llsm_fparray_length(pvt) == nspec) {   /* layer0.c:180 */
      int len = llsm_fparray_length(pvs);
      rd[r] = prd[0];
      memcpy(vtm + r * nspec, pvt, 4 * (size_t)nspec);
      nvs[r] = len < maxnhar ? len : maxnhar;
      memcpy(vsp + r * maxnhar, pvs, 4 * (size_t)nvs[r]);
    }
    pbp[r] = psyn != NULL && psyn[0] == 1;
    memcpy(psd + r * c.npsd, nm -> psd, 4 * (size_t)c.npsd);
    if(has_res) for(int j = 0; j < c.npsd; j ++) psdres[r * c.npsd + j] = res != NULL ? res[j] : (float)resbias;
    for(int ch = 0; ch < c.nchannel; ch ++) {
      size_t e = r * c.nchannel + ch;
      edc[e] = nm -> edc[ch];
      int k = nm -> eenv[ch] -> nhar;
      enhar[e] = k;
      memcpy(eampl + e * maxnhar_e, nm -> eenv[ch] -> ampl, 4 * (size_t)k);
      memcpy(ephse + e * maxnhar_e, nm -> eenv[ch] -> phse, 4 * (size_t)k);
    }
  }
  const int ny = llsm_b200_output_length(c.nfrm, c.thop, c.fs);
  const int nt = llsm_b200_template_length(ny);
  float* white = malloc(sizeof(float) * (size_t)n * c.nchannel * nt);
  for(int b = 0; b < n; b ++) for(int ch = 0; ch < c.nchannel; ch ++) {
    FP_TYPE fmin = ch == 0 ? 0 : cf[ch - 1];
    float* w = white + ((size_t)b * c.nchannel + ch) * nt;
    if(fmin >= c.fs / 2.0) { memset(w, 0, sizeof(float) * nt); continue; }
    int ntemplate = ny < 20000 ? ny : 20000;
    int ndraw = (ntemplate + 128) < 20000 ? (ntemplate + 128) : 20000;
    for(int j = 0; j < ndraw; j ++) w[j] = host_randn();
    for(int j = ndraw; j < nt; j ++) w[j] = w[(j - ndraw) % ndraw];
  }
  float* y = malloc(sizeof(float) * (size_t)n * ny);
  float* ys = malloc(sizeof(float) * (size_t)n * ny);
  float* yn = malloc(sizeof(float) * (size_t)n * ny);
  llsm_b200_frames fr = {NULL, f0, nhar, ampl, phse, psd, psdres, edc, enhar, eampl, ephse};
  llsm_b200_layer1 l1 = {rd, vtm, vsp, nvs, nspec};
  llsm_b200_soptions so; memset(& so, 0, sizeof(so));
  so.use_iczt = options -> use_iczt; so.iczt_param_a = options -> iczt_param_a; so.iczt_param_b = options -> iczt_param_b;
  so.white = white;
  llsm_b200_output out = {y, ys, yn, ny};
  hook_env env = {src};
  int rc = llsm_b200_synthesize_l1_host(ctx, & c, & fr, & l1, pbp, & so, & out, any_eff ? pulse_hook : NULL, & env);
  if(rc == 0) {
    for(int b = 0; b < n; b ++) {
      llsm_output* o = malloc(sizeof(llsm_output));
      o -> ny = ny; o -> fs = options -> fs;
      o -> y = malloc(sizeof(FP_TYPE) * (size_t)ny); o -> y_sin = malloc(sizeof(FP_TYPE) * (size_t)ny);
      o -> y_noise = malloc(sizeof(FP_TYPE) * (size_t)ny);
      memcpy(o -> y, y + (size_t)b * ny, 4 * (size_t)ny);
      memcpy(o -> y_sin, ys + (size_t)b * ny, 4 * (size_t)ny);
      memcpy(o -> y_noise, yn + (size_t)b * ny, 4 * (size_t)ny);
      dst[b] = o;
    }
  }
  free(f0); free(nhar); free(ampl); free(phse); free(psd); free(psdres); free(edc); free(enhar); free(eampl);
  free(ephse); free(rd); free(vtm); free(vsp); free(nvs); free(pbp); free(white); free(y); free(ys); free(yn);
  return rc;
}

/* ------------------------------------------------------------------ synthesis ------------------- */
int llsm_synthesize_batch(llsm_soptions* options, llsm_chunk** src, int n, llsm_output** dst) {
  g_compat_err[0] = 0;
  for(int b = 0; b < n; b ++) dst[b] = NULL;
  if(options == NULL || n < 1) { set_err("bad arguments"); return -1; }
  for(int b = 0; b < n; b ++) {
    if(! chunk_ok(src[b])) { set_err("chunk fails the layer-0 integrity check"); return -1; }
    if(b > 0 && ! same_conf(src[0] -> conf, src[b] -> conf)) {
      set_err("llsm_synthesize_batch: chunks must share one configuration"); return -1;
    }
  }
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return -1;
  if(options -> use_l1) return synthesize_l1_batch(ctx, options, src, n, dst);

  llsm_container* conf = src[0] -> conf;
  llsm_b200_conf c; memset(& c, 0, sizeof(c));
  c.nutt = n;
  c.nfrm = *(int*)llsm_container_get(conf, LLSM_CONF_NFRM);
  c.npsd = *(int*)llsm_container_get(conf, LLSM_CONF_NPSD);
  c.nchannel = *(int*)llsm_container_get(conf, LLSM_CONF_NCHANNEL);
  c.thop = *(FP_TYPE*)llsm_container_get(conf, LLSM_CONF_THOP);
  c.fs = options -> fs;
  FP_TYPE* lip = llsm_container_get(conf, LLSM_CONF_LIPRADIUS);
  c.lip_radius = lip != NULL ? *lip : 1.5f;
  FP_TYPE* cf = llsm_container_get(conf, LLSM_CONF_CHANFREQ);
  if(c.nchannel < 1 || c.nchannel > LLSM_B200_MAXCHANNEL || c.nfrm < 1) { set_err("unsupported configuration"); return -1; }
  for(int i = 0; i < c.nchannel - 1; i ++) c.chanfreq[i] = cf[i];

  /* row lengths = the largest harmonic counts present */
  int maxnhar = 1, maxnhar_e = 1, has_res = 0;
  for(int b = 0; b < n; b ++) for(int i = 0; i < c.nfrm; i ++) {
    llsm_container* f = src[b] -> frames[i];
    llsm_hmframe* hm = llsm_container_get(f, LLSM_FRAME_HM);
    llsm_nmframe* nm = llsm_container_get(f, LLSM_FRAME_NM);
    if(hm != NULL && hm -> nhar > maxnhar) maxnhar = hm -> nhar;
    if(nm -> nchannel != c.nchannel || nm -> npsd != c.npsd) { set_err("frame / conf size mismatch"); return -1; }
    for(int ch = 0; ch < c.nchannel; ch ++) if(nm -> eenv[ch] -> nhar > maxnhar_e) maxnhar_e = nm -> eenv[ch] -> nhar;
    if(llsm_container_get(f, LLSM_FRAME_PSDRES) != NULL) has_res = 1;
  }
  if(maxnhar > 2048) maxnhar = 2048;                    /* layer0.c:119 */
  c.maxnhar = maxnhar; c.maxnhar_e = maxnhar_e;

  const size_t BF = (size_t)n * c.nfrm;
  float* f0 = calloc(BF, 4); int* nhar = calloc(BF, 4);
  float* ampl = calloc(BF * maxnhar, 4); float* phse = calloc(BF * maxnhar, 4);
  float* psd = calloc(BF * c.npsd, 4); float* psdres = has_res ? calloc(BF * c.npsd, 4) : NULL;
  float* edc = calloc(BF * c.nchannel, 4); int* enhar = calloc(BF * c.nchannel, 4);
  float* eampl = calloc(BF * c.nchannel * maxnhar_e, 4); float* ephse = calloc(BF * c.nchannel * maxnhar_e, 4);
  const double resbias = 0.375 / 2.3025851 * 10.0;
  for(int b = 0; b < n; b ++) for(int i = 0; i < c.nfrm; i ++) {
    size_t r = (size_t)b * c.nfrm + i;
    llsm_container* f = src[b] -> frames[i];
    FP_TYPE* pf0 = llsm_container_get(f, LLSM_FRAME_F0);
    llsm_hmframe* hm = llsm_container_get(f, LLSM_FRAME_HM);
    llsm_nmframe* nm = llsm_container_get(f, LLSM_FRAME_NM);
    FP_TYPE* res = llsm_container_get(f, LLSM_FRAME_PSDRES);
    f0[r] = pf0[0];
    if(pf0[0] != 0 && hm != NULL) {
      int k = hm -> nhar < maxnhar ? hm -> nhar : maxnhar;
      nhar[r] = k;
      memcpy(ampl + r * maxnhar, hm -> ampl, 4 * (size_t)k);
      memcpy(phse + r * maxnhar, hm -> phse, 4 * (size_t)k);
    }
    memcpy(psd + r * c.npsd, nm -> psd, 4 * (size_t)c.npsd);
    if(has_res) {
      /* a frame without PSDRES adds nothing (layer0.c:599-601): encode as the bias itself */
      for(int j = 0; j < c.npsd; j ++) psdres[r * c.npsd + j] = res != NULL ? res[j] : (float)resbias;
    }
    for(int ch = 0; ch < c.nchannel; ch ++) {
      size_t e = r * c.nchannel + ch;
      edc[e] = nm -> edc[ch];
      int k = nm -> eenv[ch] -> nhar;
      enhar[e] = k;
      memcpy(eampl + e * maxnhar_e, nm -> eenv[ch] -> ampl, 4 * (size_t)k);
      memcpy(ephse + e * maxnhar_e, nm -> eenv[ch] -> phse, 4 * (size_t)k);
    }
  }

  const int ny = llsm_b200_output_length(c.nfrm, c.thop, c.fs);
  const int nt = llsm_b200_template_length(ny);
  float* white = malloc(sizeof(float) * (size_t)n * c.nchannel * nt);
  for(int b = 0; b < n; b ++) for(int ch = 0; ch < c.nchannel; ch ++) {
    /* llsm_synthesize_noise_excitation stops drawing at the first channel above Nyquist (layer0.c:543) */
    FP_TYPE fmin = ch == 0 ? 0 : cf[ch - 1];
    float* w = white + ((size_t)b * c.nchannel + ch) * nt;
    if(fmin >= c.fs / 2.0) { memset(w, 0, sizeof(float) * nt); continue; }
    int ntemplate = ny < 20000 ? ny : 20000;
    int ndraw = (ntemplate + 128) < 20000 ? (ntemplate + 128) : 20000;   /* dsputils.c:355-359 */
    for(int j = 0; j < ndraw; j ++) w[j] = host_randn();
    for(int j = ndraw; j < nt; j ++) w[j] = w[(j - ndraw) % ndraw];
  }

  float* y = malloc(sizeof(float) * (size_t)n * ny);
  float* ys = malloc(sizeof(float) * (size_t)n * ny);
  float* yn = malloc(sizeof(float) * (size_t)n * ny);
  llsm_b200_frames fr = {NULL, f0, nhar, ampl, phse, psd, psdres, edc, enhar, eampl, ephse};
  llsm_b200_soptions so; memset(& so, 0, sizeof(so));
  so.use_iczt = options -> use_iczt; so.iczt_param_a = options -> iczt_param_a; so.iczt_param_b = options -> iczt_param_b;
  so.white = white; so.seed = 0;
  llsm_b200_output out = {y, ys, yn, ny};
  int rc = llsm_b200_synthesize_l0_host(ctx, & c, & fr, & so, & out);
  if(rc == 0) {
    for(int b = 0; b < n; b ++) {
      llsm_output* o = malloc(sizeof(llsm_output));
      o -> ny = ny; o -> fs = options -> fs;
      o -> y = malloc(sizeof(FP_TYPE) * (size_t)ny); o -> y_sin = malloc(sizeof(FP_TYPE) * (size_t)ny);
      o -> y_noise = malloc(sizeof(FP_TYPE) * (size_t)ny);
      memcpy(o -> y, y + (size_t)b * ny, 4 * (size_t)ny);
      memcpy(o -> y_sin, ys + (size_t)b * ny, 4 * (size_t)ny);
      memcpy(o -> y_noise, yn + (size_t)b * ny, 4 * (size_t)ny);
      dst[b] = o;
    }
  }
  free(f0); free(nhar); free(ampl); free(phse); free(psd); free(psdres); free(edc); free(enhar);
  free(eampl); free(ephse); free(white); free(y); free(ys); free(yn);
  return rc;
}

llsm_output* llsm_synthesize(llsm_soptions* options, llsm_chunk* src) {     /* layer0.c:636-664 */
  llsm_output* out = NULL;
  if(llsm_synthesize_batch(options, & src, 1, & out) != 0) return NULL;
  return out;
}

/* ------------------------------------------------------------------ analysis -------------------- */
llsm_chunk* llsm_analyze(llsm_aoptions* options, FP_TYPE* x, int nx, FP_TYPE fs, FP_TYPE* f0,
  int nfrm, FP_TYPE** x_ap) {                                               /* layer0.c:478-511 */
  g_compat_err[0] = 0;
  if(options == NULL || x == NULL || f0 == NULL || nx < 1 || nfrm < 1) { set_err("bad arguments"); return NULL; }
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return NULL;
  llsm_b200_conf c; memset(& c, 0, sizeof(c));
  c.nutt = 1; c.nfrm = nfrm; c.maxnhar = options -> maxnhar > 0 ? options -> maxnhar : 1;
  c.maxnhar_e = options -> maxnhar_e; c.npsd = options -> npsd; c.nchannel = options -> nchannel;
  c.fs = fs; c.thop = options -> thop; c.lip_radius = options -> lip_radius;
  if(c.nchannel < 1 || c.nchannel > LLSM_B200_MAXCHANNEL) { set_err("unsupported nchannel"); return NULL; }
  for(int i = 0; i < c.nchannel - 1; i ++) c.chanfreq[i] = options -> chanfreq[i];
  llsm_b200_aoptions ao = {options -> f0_refine, options -> hm_method, options -> rel_winsize};

  const size_t F = (size_t)nfrm;
  int mne = c.maxnhar_e > 0 ? c.maxnhar_e : 1;
  int* nhar = calloc(F, 4); float* ampl = calloc(F * c.maxnhar, 4); float* phse = calloc(F * c.maxnhar, 4);
  float* psd = calloc(F * c.npsd, 4); float* psdres = calloc(F * c.npsd, 4);
  float* edc = calloc(F * c.nchannel, 4); int* enhar = calloc(F * c.nchannel, 4);
  float* eampl = calloc(F * c.nchannel * mne, 4); float* ephse = calloc(F * c.nchannel * mne, 4);
  float* xres = malloc(sizeof(float) * (size_t)nx);
  llsm_b200_frames_out fo = {f0, nhar, ampl, phse, psd, psdres, edc, enhar, eampl, ephse};
  int rc = llsm_b200_analyze_l0_host(ctx, & c, & ao, x, nx, nx, & fo, xres);   /* f0 refined in place */
  llsm_chunk* ret = NULL;
  if(rc == 0) {
    llsm_container* conf = llsm_aoptions_toconf(options, fs / 2.0);
    ((int*)llsm_container_get(conf, LLSM_CONF_NFRM))[0] = nfrm;
    ret = llsm_create_chunk(conf, 1);
    llsm_delete_container(conf);
    for(int i = 0; i < nfrm; i ++) {
      llsm_container* f = ret -> frames[i];
      ((FP_TYPE*)llsm_container_get(f, LLSM_FRAME_F0))[0] = f0[i];
      llsm_nmframe* nm = llsm_container_get(f, LLSM_FRAME_NM);
      if(f0[i] != 0) {
        llsm_hmframe* hm = llsm_create_hmframe(nhar[i]);
        memcpy(hm -> ampl, ampl + (size_t)i * c.maxnhar, 4 * (size_t)nhar[i]);
        memcpy(hm -> phse, phse + (size_t)i * c.maxnhar, 4 * (size_t)nhar[i]);
        llsm_container_attach(f, LLSM_FRAME_HM, hm, llsm_delete_hmframe, llsm_copy_hmframe);
      }
      memcpy(nm -> psd, psd + (size_t)i * c.npsd, 4 * (size_t)c.npsd);
      FP_TYPE* res = llsm_create_fparray(c.npsd);
      memcpy(res, psdres + (size_t)i * c.npsd, 4 * (size_t)c.npsd);
      llsm_container_attach(f, LLSM_FRAME_PSDRES, res, llsm_delete_fparray, llsm_copy_fparray);
      for(int ch = 0; ch < c.nchannel; ch ++) {
        size_t e = (size_t)i * c.nchannel + ch;
        nm -> edc[ch] = edc[e];
        if(f0[i] == 0) continue;                        /* layer0.c:452 */
        llsm_hmframe* eh = llsm_create_hmframe(enhar[e]);
        memcpy(eh -> ampl, eampl + e * mne, 4 * (size_t)enhar[e]);
        memcpy(eh -> phse, ephse + e * mne, 4 * (size_t)enhar[e]);
        llsm_copy_hmframe_inplace(nm -> eenv[ch], eh);
        llsm_delete_hmframe(eh);
      }
    }
    if(x_ap != NULL) { *x_ap = xres; xres = NULL; }
  }
  free(nhar); free(ampl); free(phse); free(psd); free(psdres); free(edc); free(enhar); free(eampl);
  free(ephse); free(xres);
  return ret;
}

/* ------------------------------------------------------------------ dsputils.h ------------------- */
#include "../../include/dsputils.h"

static void dsp_conf(llsm_b200_conf* c, int nfrm, int maxnhar, float fs, float thop) {
  memset(c, 0, sizeof(*c));
  c -> nutt = 1; c -> nfrm = nfrm; c -> maxnhar = maxnhar > 0 ? maxnhar : 1; c -> maxnhar_e = 0; c -> npsd = 2;
  c -> nchannel = 1; c -> fs = fs; c -> thop = thop; c -> lip_radius = 1.5f;
}

void llsm_refine_f0(FP_TYPE* x, int nx, FP_TYPE fs, FP_TYPE* f0, int nfrm, FP_TYPE thop) {
  g_compat_err[0] = 0;
  if(x == NULL || f0 == NULL || nx < 1 || nfrm < 1) return;
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return;
  llsm_b200_conf c; dsp_conf(& c, nfrm, 1, fs, thop);
  if(llsm_b200_refine_f0_host(ctx, & c, x, nx, nx, f0) != 0) set_err(llsm_b200_last_error());
}

int llsm_get_fftsize(FP_TYPE* f0, int nfrm, FP_TYPE fs, FP_TYPE rel_winsize) {   /* dsputils.c:318-326 */
  FP_TYPE minf0 = 1000;
  for(int i = 0; i < nfrm; i ++) if(f0[i] > 0 && f0[i] < minf0) minf0 = f0[i];
  FP_TYPE t = fs / minf0; t = t * rel_winsize; t = t / 2;
  int max_winsize = (int)round((double)t) * 2;
  int n = 1;
  while(n < max_winsize) n *= 2;
  return n;
}

void llsm_harmonic_analysis(FP_TYPE* x, int nx, FP_TYPE fs, FP_TYPE* f0, int nfrm, FP_TYPE thop, FP_TYPE rel_winsize,
  int maxnhar, int method, int* dst_nhar, FP_TYPE** dst_ampl, FP_TYPE** dst_phse) {
  g_compat_err[0] = 0;
  if(x == NULL || f0 == NULL || nx < 1 || nfrm < 1 || maxnhar < 1 || dst_nhar == NULL || dst_ampl == NULL || dst_phse == NULL) return;
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) return;
  llsm_b200_conf c; dsp_conf(& c, nfrm, maxnhar, fs, thop);
  llsm_b200_aoptions ao = {0, method == LLSM_AOPTION_HMPP ? 0 : 1, rel_winsize};
  int* nhar = calloc((size_t)nfrm, sizeof(int));
  float* ampl = calloc((size_t)nfrm * maxnhar, sizeof(float));
  float* phse = calloc((size_t)nfrm * maxnhar, sizeof(float));
  if(llsm_b200_harmonic_analysis_host(ctx, & c, & ao, x, nx, nx, f0, nhar, ampl, phse) == 0) {
    for(int i = 0; i < nfrm; i ++) {
      if(! (f0[i] > 0) || nhar[i] < 0) continue;           /* unvoiced frames are left untouched (dsputils.c:186-192) */
      dst_nhar[i] = nhar[i];
      dst_ampl[i] = calloc(nhar[i] > 0 ? nhar[i] : 1, sizeof(FP_TYPE));
      dst_phse[i] = calloc(nhar[i] > 0 ? nhar[i] : 1, sizeof(FP_TYPE));
      memcpy(dst_ampl[i], ampl + (size_t)i * maxnhar, sizeof(FP_TYPE) * (size_t)nhar[i]);
      memcpy(dst_phse[i], phse + (size_t)i * maxnhar, sizeof(FP_TYPE) * (size_t)nhar[i]);
    }
  } else set_err(llsm_b200_last_error());
  free(nhar); free(ampl); free(phse);
}

static FP_TYPE* harmonic_frame(FP_TYPE* ampl, FP_TYPE* phse, int nhar, FP_TYPE f0, int nx, int iczt) {
  g_compat_err[0] = 0;
  if(nx < 1 || nhar < 0 || (nhar > 0 && (ampl == NULL || phse == NULL))) return NULL;
  FP_TYPE* y = calloc((size_t)nx, sizeof(FP_TYPE));
  if(nhar == 0) return y;
  llsm_b200_ctx* ctx = shared_ctx();
  if(ctx == NULL) { free(y); return NULL; }
  if(llsm_b200_harmonic_frames_host(ctx, 1, nhar, & nhar, & f0, ampl, phse, nx, iczt, y) != 0) {
    set_err(llsm_b200_last_error()); free(y); return NULL;
  }
  return y;
}
FP_TYPE* llsm_synthesize_harmonic_frame(FP_TYPE* ampl, FP_TYPE* phse, int nhar, FP_TYPE f0, int nx) {
  return harmonic_frame(ampl, phse, nhar, f0, nx, 0);
}
FP_TYPE* llsm_synthesize_harmonic_frame_iczt(FP_TYPE* ampl, FP_TYPE* phse, int nhar, FP_TYPE f0, int nx) {
  return harmonic_frame(ampl, phse, nhar, f0, nx, 1);
}

#include "compat_rt.inc"
