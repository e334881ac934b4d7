/*
  llsmrt.h -- drop-in streaming synthesis API of libllsm2_b200.

  Source-compatible with reference llsmrt.h:26-56 (libllsm2 2.1.0): same opaque handle, same eight
  entry points, same blocking rules (feed waits while the output ring cannot take the next hop, fetch
  never waits). Behind it one llsm_b200_rt stream (include/llsm_b200.h) runs on the GPU: every feed is
  one kernel launch that advances the rings held in device memory and hands back the hop's samples,
  which this layer keeps in two host rings for the per-sample fetch calls.

  No CPU fallback: llsm_create_rtsynth_buffer returns NULL without a CUDA device (llsm_last_error()).
*/
#ifndef LLSM_LLSMRT_H
#define LLSM_LLSMRT_H

#include "llsm.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void llsm_rtsynth_buffer;

/* reference llsmrt.c:157-223. capacity_samples sizes the two output rings. */
llsm_rtsynth_buffer* llsm_create_rtsynth_buffer(llsm_soptions* options, llsm_container* conf,
  int capacity_samples);
/* reference llsmrt.c:225-253 */
void llsm_delete_rtsynth_buffer(llsm_rtsynth_buffer* dst);
/* -sin_pos - curr_nhop, reference llsmrt.c:568-571 */
int llsm_rtsynth_buffer_getlatency(llsm_rtsynth_buffer* src);
/* samples waiting in the output rings, reference llsmrt.c:573-576 */
int llsm_rtsynth_buffer_numoutput(llsm_rtsynth_buffer* src);
/* one frame in, next_nhop samples out; blocks while the output ring is full (reference llsmrt.c:505-521) */
void llsm_rtsynth_buffer_feed(llsm_rtsynth_buffer* dst, llsm_container* frame);
/* one sample out (periodic + aperiodic); 1 on success, 0 when empty (reference llsmrt.c:523-543) */
int llsm_rtsynth_buffer_fetch(llsm_rtsynth_buffer* src, FP_TYPE* dst);
/* same with the two parts apart (reference llsmrt.c:545-566) */
int llsm_rtsynth_buffer_fetch_decomposed(llsm_rtsynth_buffer* src, FP_TYPE* dst_p, FP_TYPE* dst_ap);
/* restart the clock and the signal rings (reference llsmrt.c:578-602) */
void llsm_rtsynth_buffer_clear(llsm_rtsynth_buffer* dst);

#ifdef __cplusplus
}
#endif
#endif
