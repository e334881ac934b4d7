/*
  dsputils.h -- the per-frame DSP routines of libllsm2 that applications and the reference's own tests call directly
  (the reference installs a header of this name: /root/reference/dsputils.h:25-132, makefile:135). This drop-in
  exports the ones on the analysis / synthesis hot path, each running on the GPU as a batch of one call:

    llsm_refine_f0                       dsputils.h:29-30   test/ -- used through llsm_analyze
    llsm_harmonic_analysis               dsputils.h:53-55   test/test-dsputils.c:80 (chirp known-answer test)
    llsm_get_fftsize                     dsputils.h:78      (host arithmetic)
    llsm_synthesize_harmonic_frame       dsputils.h:81-82   test/test-harmonic.c:18,42,58
    llsm_synthesize_harmonic_frame_iczt  dsputils.h:85-86   test/test-harmonic.c:25,40

  Same signatures, ownership (results are malloc'd, the caller frees) and error behaviour (silent return / NULL).
  Not exported: the remaining helpers of the reference header (spectrogram, peak picking on a caller-made spectrum,
  warped-frequency utilities, glottal-model cache ...); they exist only inside the kernels here.
*/
#ifndef LLSM_DSPUTILS_H
#define LLSM_DSPUTILS_H

#include "llsm.h"

#ifdef __cplusplus
extern "C" {
#endif

void llsm_refine_f0(FP_TYPE* x, int nx, FP_TYPE fs, FP_TYPE* f0, int nfrm, FP_TYPE thop);

/* dst_nhar[i] / dst_ampl[i] / dst_phse[i] are written for voiced frames only (f0[i] > 0); dst_ampl[i] and dst_phse[i]
   are calloc'd arrays of dst_nhar[i] numbers that the caller frees. method: LLSM_AOPTION_HMPP / LLSM_AOPTION_HMCZT. */
void llsm_harmonic_analysis(FP_TYPE* x, int nx, FP_TYPE fs, FP_TYPE* f0, int nfrm, FP_TYPE thop, FP_TYPE rel_winsize,
  int maxnhar, int method, int* dst_nhar, FP_TYPE** dst_ampl, FP_TYPE** dst_phse);

int llsm_get_fftsize(FP_TYPE* f0, int nfrm, FP_TYPE fs, FP_TYPE rel_winsize);

/* f0 in cycles per sample; nx samples centred at nx / 2; the caller frees the result */
FP_TYPE* llsm_synthesize_harmonic_frame(FP_TYPE* ampl, FP_TYPE* phse, int nhar, FP_TYPE f0, int nx);
FP_TYPE* llsm_synthesize_harmonic_frame_iczt(FP_TYPE* ampl, FP_TYPE* phse, int nhar, FP_TYPE f0, int nx);

#ifdef __cplusplus
}
#endif
#endif
