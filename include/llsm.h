/*
  llsm.h -- drop-in C API of libllsm2_b200 for the layer-0 analysis / synthesis path.

  Source-compatible with the public interface of libllsm2 2.1.0 (reference llsm.h): same type
  layouts, same index constants, same function names and ownership rules, so existing callers
  relink against libllsm2_b200.so unchanged. Written from the interface contract, not copied:
  each block cites the reference declaration it has to stay compatible with.

  What runs where
    * llsm_analyze / llsm_synthesize (HM and use_l1 pulse-by-pulse) / llsm_chunk_tolayer1 /
      llsm_chunk_tolayer0 / llsm_frame_tolayer0: packed to flat arrays and executed by the CUDA kernels
      behind include/llsm_b200.h (batch of one). No CPU fallback: they return NULL (or silently, for the
      void functions, as the reference does on bad input) when no CUDA device is usable
      (llsm_last_error() tells why). llsm_pbpeffect callbacks run on the host, once per pulse, in order.
    * containers, frames, chunks, option structs, phase utilities: plain host C (bookkeeping).
    * llsm_coder_encode / llsm_coder_decode_layer0 / _layer1: one frame per call through the batched device coder
      (include/llsm_b200.h has the batch form). llsm_frame_compute_snr is outside the path (returns NULL).

  FP_TYPE must be float (the device kernels compute in FP32 like the reference's default build).
*/
#ifndef LLSM_H
#define LLSM_H

#ifndef FP_TYPE
#define FP_TYPE float
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LLSM_VERSION_STRING   "2.1.0-b200"
#define LLSM_VERSION_MAJOR    2
#define LLSM_VERSION_MINOR    1
#define LLSM_VERSION_REVISION 0

/* ---- generic container (reference llsm.h:30-92) ------------------------------------------- */
typedef void  (*llsm_fdestructor)(void*);
typedef void* (*llsm_fcopy)(void*);

typedef struct {
  void** members;                 /* nmember slots, NULL = absent */
  llsm_fdestructor* destructors;  /* NULL = not owned */
  llsm_fcopy* copyctors;          /* NULL = shallow copy */
  int nmember;
} llsm_container;

llsm_container* llsm_create_container(int nmember);
llsm_container* llsm_copy_container(llsm_container* src);
void  llsm_copy_container_inplace(llsm_container* dst, llsm_container* src);
void  llsm_delete_container(llsm_container* dst);
void* llsm_container_get(llsm_container* src, int index);
void  llsm_container_attach_(llsm_container* dst, int index, void* ptr,
        llsm_fdestructor dtor, llsm_fcopy copyctor);
void  llsm_container_remove(llsm_container* dst, int index);
#define llsm_container_attach(dst, index, ptr, dtor, copyctor) \
  llsm_container_attach_((dst), (index), (ptr), (llsm_fdestructor)(dtor), (llsm_fcopy)(copyctor))

/* boxed scalars and length-prefixed arrays (reference llsm.h:38-47) */
FP_TYPE* llsm_create_fp(FP_TYPE x);
int*     llsm_create_int(int x);
FP_TYPE* llsm_create_fparray(int size);
FP_TYPE* llsm_copy_fp(FP_TYPE* src);
int*     llsm_copy_int(int* src);
FP_TYPE* llsm_copy_fparray(FP_TYPE* src);
void     llsm_delete_fp(FP_TYPE* dst);
void     llsm_delete_int(int* dst);
void     llsm_delete_fparray(FP_TYPE* dst);
int      llsm_fparray_length(FP_TYPE* src);

/* ---- member indices (reference llsm.h:98-128) --------------------------------------------- */
enum {
  LLSM_FRAME_F0 = 0, LLSM_FRAME_HM = 1, LLSM_FRAME_NM = 2, LLSM_FRAME_PSDRES = 3,
  LLSM_FRAME_PBPEFF = 8, LLSM_FRAME_PBPSYN = 9, LLSM_FRAME_RD = 10, LLSM_FRAME_VTMAGN = 11,
  LLSM_FRAME_VSPHSE = 12
};
enum {
  LLSM_CONF_NFRM = 0, LLSM_CONF_THOP = 1, LLSM_CONF_MAXNHAR = 2, LLSM_CONF_MAXNHAR_E = 3,
  LLSM_CONF_NPSD = 4, LLSM_CONF_NOSWARP = 5, LLSM_CONF_FNYQ = 6, LLSM_CONF_NCHANNEL = 7,
  LLSM_CONF_CHANFREQ = 8, LLSM_CONF_NSPEC = 10, LLSM_CONF_LIPRADIUS = 11
};

/* ---- harmonic / noise model frames (reference llsm.h:134-174) ------------------------------ */
typedef struct { FP_TYPE* ampl; FP_TYPE* phse; int nhar; } llsm_hmframe;
typedef struct {
  llsm_hmframe** eenv;   /* per-channel envelope harmonics */
  FP_TYPE* edc;          /* per-channel envelope mean */
  FP_TYPE* psd;          /* dB */
  int npsd;
  int nchannel;
} llsm_nmframe;

llsm_hmframe* llsm_create_hmframe(int nhar);
llsm_hmframe* llsm_copy_hmframe(llsm_hmframe* src);
void     llsm_copy_hmframe_inplace(llsm_hmframe* dst, llsm_hmframe* src);
void     llsm_delete_hmframe(llsm_hmframe* dst);
void     llsm_hmframe_phaseshift(llsm_hmframe* dst, FP_TYPE theta);
FP_TYPE* llsm_hmframe_harpsd(llsm_hmframe* src, int db_scale);
llsm_nmframe* llsm_create_nmframe(int nchannel, int nhar_e, int npsd);
llsm_nmframe* llsm_copy_nmframe(llsm_nmframe* src);
void     llsm_copy_nmframe_inplace(llsm_nmframe* dst, llsm_nmframe* src);
void     llsm_delete_nmframe(llsm_nmframe* dst);

/* ---- glottal-flow model and pulse-by-pulse effects (reference llsm.h:181-208) -------------- */
typedef struct { FP_TYPE Fa; FP_TYPE Rk; FP_TYPE Rg; FP_TYPE T0; FP_TYPE Ee; } llsm_gfm;
typedef void (*llsm_fgfm)(llsm_gfm* dst, FP_TYPE* delta_t, void* info, llsm_container* src_frame);
typedef struct { llsm_fgfm modifier; void* info; } llsm_pbpeffect;
llsm_pbpeffect* llsm_create_pbpeffect(llsm_fgfm modifier, void* info);
llsm_pbpeffect* llsm_copy_pbpeffect(llsm_pbpeffect* src);
void llsm_delete_pbpeffect(llsm_pbpeffect* dst);

/* ---- frames (reference llsm.h:217-243) ------------------------------------------------------ */
llsm_container* llsm_create_frame(int nhar, int nchannel, int nhar_e, int npsd);
void     llsm_frame_tolayer0(llsm_container* dst, llsm_container* conf);
void     llsm_frame_phaseshift(llsm_container* dst, FP_TYPE theta);
void     llsm_frame_phasesync_rps(llsm_container* dst, int layer1_based);
FP_TYPE* llsm_frame_compute_snr(llsm_container* src, llsm_container* conf, int as_aperiodicity);
int      llsm_frame_checklayer0(llsm_container* src);
int      llsm_frame_checklayer1(llsm_container* src);
int      llsm_conf_checklayer0(llsm_container* src);
int      llsm_conf_checklayer1(llsm_container* src);

/* ---- synthesis result (reference llsm.h:246-255) ------------------------------------------- */
typedef struct { int ny; FP_TYPE fs; FP_TYPE* y; FP_TYPE* y_sin; FP_TYPE* y_noise; } llsm_output;
void llsm_delete_output(llsm_output* dst);

/* ---- options (reference llsm.h:260-304) ------------------------------------------------------ */
typedef struct {
  FP_TYPE thop; int maxnhar; int maxnhar_e; int npsd; int nchannel; FP_TYPE* chanfreq;
  FP_TYPE lip_radius; int f0_refine; int hm_method; FP_TYPE rel_winsize;
} llsm_aoptions;
#define LLSM_AOPTION_HMPP  0
#define LLSM_AOPTION_HMCZT 1
llsm_aoptions*  llsm_create_aoptions(void);
void            llsm_delete_aoptions(llsm_aoptions* dst);
llsm_container* llsm_aoptions_toconf(llsm_aoptions* src, FP_TYPE fnyq);

typedef struct { FP_TYPE fs; int use_iczt; int use_l1; FP_TYPE iczt_param_a; FP_TYPE iczt_param_b; } llsm_soptions;
llsm_soptions* llsm_create_soptions(FP_TYPE fs);
void           llsm_delete_soptions(llsm_soptions* dst);

/* ---- chunks and the two pipelines (reference llsm.h:310-339) --------------------------------- */
typedef struct { llsm_container* conf; llsm_container** frames; } llsm_chunk;
llsm_chunk* llsm_create_chunk(llsm_container* conf, int init_frames);
llsm_chunk* llsm_copy_chunk(llsm_chunk* src);
void     llsm_delete_chunk(llsm_chunk* dst);
void     llsm_chunk_tolayer1(llsm_chunk* dst, int nfft);
void     llsm_chunk_tolayer0(llsm_chunk* dst);
void     llsm_chunk_phasesync_rps(llsm_chunk* dst, int layer1_based);
void     llsm_chunk_phasepropagate(llsm_chunk* dst, int sign);
FP_TYPE* llsm_chunk_getf0(llsm_chunk* src, int* dst_nfrm);

llsm_chunk*  llsm_analyze(llsm_aoptions* options, FP_TYPE* x, int nx, FP_TYPE fs, FP_TYPE* f0,
               int nfrm, FP_TYPE** x_ap);
llsm_output* llsm_synthesize(llsm_soptions* options, llsm_chunk* src);

/* ---- coder (reference llsm.h:342-362, coder.c) ---------------------------------------------------
   A frame <-> a vector of order_spec + order_bap + 3 numbers: [voicing, f0, Rd, order_spec mel-cepstral numbers of
   the total power spectrum, order_bap band aperiodicities]. The conf needs FNYQ, NCHANNEL, MAXNHAR_E, NPSD, NSPEC
   and LIPRADIUS (i.e. a layer-1 chunk's conf). Encoding reads F0, NM.psd and, for voiced frames, RD and VTMAGN. */
typedef void llsm_coder;
llsm_coder*     llsm_create_coder(llsm_container* conf, int order_spec, int order_bap);
void            llsm_delete_coder(llsm_coder* dst);
FP_TYPE*        llsm_coder_encode(llsm_coder* c, llsm_container* src);          /* caller frees */
llsm_container* llsm_coder_decode_layer1(llsm_coder* c, FP_TYPE* src);          /* RD, VTMAGN, VSPHSE, no HM */
llsm_container* llsm_coder_decode_layer0(llsm_coder* c, FP_TYPE* src);          /* RD, HM */

/* ---- extensions of this library ---------------------------------------------------------------- */
/* Batched forms of the two pipelines: n utterances that share one configuration go through the
   kernels in a single launch sequence. Results are exactly those of n single calls. */
int llsm_synthesize_batch(llsm_soptions* options, llsm_chunk** src, int n, llsm_output** dst);
/* Reason of the last NULL / failure returned by this library on the calling thread. */
const char* llsm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
