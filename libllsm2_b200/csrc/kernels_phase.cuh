// Chunk phase utilities on the batched structure-of-arrays frames (what real use runs
// between analysis and synthesis, test/test-layer0-anasynth.c:62-63, test-llsmrt.c:88,112):
//   llsm_chunk_phasepropagate (layer0.c:694-706): theta_i = cumsum(f0)_i * (thop * sign * 2 pi), then
//   llsm_chunk_phasesync_rps  (layer0.c:687-692, frame.c:168-178): theta_i = -phse_i[0] (or -vsphse_i[0]), then
//   llsm_frame_phaseshift     (frame.c:152-166, :57-60): phse[k], eenv phases, VSPHSE[k] <- wrap(. + theta (k + 1)).
// Everything is evaluated in the reference's precision: the running sum and wrap() in double, the products as
// C promotes them, one float rounding where the reference stores a FP_TYPE.
#pragma once
#include "common.cuh"

struct PhaseParams {
  int nutt, nfrm, maxnhar, maxnhar_e, nchannel;
  const int* nfrm_utt;      // [B] or NULL
  const float* f0;          // [B][F]
  const int* nhar;          // [B][F]
  float* phse;              // [B][F][maxnhar]            in place
  const int* enhar;         // [B][F][nch]
  float* ephse;             // [B][F][nch][maxnhar_e]     in place
  float* vsphse;            // [B][F][maxnhar] or NULL    in place
  const int* nvs;           // [B][F] or NULL
  float thop;
  int mode;                 // 0: phasepropagate(sign = arg), 1: phasesync_rps(layer1_based = arg)
  int arg;
  float* theta;             // [B][F] scratch: the shift of every frame
};

// wrap() of the oracle's ciglet shim: the argument is a FP_TYPE, the reduction runs in double
// (explicit round-to-nearest products and sums: the reference is built without FMA contraction)
__device__ __forceinline__ float phase_wrap(float p) {
  const double fl = floor(__dadd_rn((double)p, LLSM_PI) / (2.0 * LLSM_PI));
  double q = __dadd_rn((double)p, -__dmul_rn(2.0 * LLSM_PI, fl));
  if(q <= -LLSM_PI) q = __dadd_rn(q, 2.0 * LLSM_PI);
  return (float)q;
}
// phse + theta * (k + 1.0), rounded to FP_TYPE as the argument of wrap() (frame.c:59)
__device__ __forceinline__ float phase_shifted(float p, float theta, int k) {
  return phase_wrap((float)__dadd_rn((double)p, __dmul_rn((double)theta, (double)k + 1.0)));
}

// one thread per utterance: the frame shifts (a running sum along time for the propagation)
__global__ void phase_theta_kernel(PhaseParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= P.nutt) return;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const size_t row = (size_t)b * P.nfrm;
  if(P.mode == 0) {
    const double rhs = __dmul_rn((double)__fmul_rn(P.thop, (float)P.arg) * 2.0, LLSM_PI);   // *thop * sign * 2.0 * M_PI
    double acc = 0;                                                         // cumsum
    for(int i = 0; i < nf; i ++) {
      acc += (double)P.f0[row + i];
      const float d = (float)acc;
      P.theta[row + i] = (float)__dmul_rn((double)d, rhs);
    }
  } else {
    for(int i = 0; i < nf; i ++) {
      float ref = 0.f;
      if(P.arg && P.vsphse && P.nvs && P.nvs[row + i] > 0) ref = P.vsphse[(row + i) * (size_t)P.maxnhar];
      else if(P.nhar[row + i] > 0) ref = P.phse[(row + i) * (size_t)P.maxnhar];
      P.theta[row + i] = -ref;
    }
  }
}

// one CTA per frame: shift every phase vector of the frame
__global__ void phase_shift_kernel(PhaseParams P) {
  const int f = blockIdx.x, b = blockIdx.y;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  if(f >= nf) return;
  const size_t fi = (size_t)b * P.nfrm + f;
  const float theta = P.theta[fi];
  const int nh = min(P.nhar[fi], P.maxnhar);
  for(int k = threadIdx.x; k < nh; k += blockDim.x) {
    float* p = P.phse + fi * P.maxnhar + k;
    *p = phase_shifted(*p, theta, k);
  }
  for(int e = threadIdx.x; e < P.nchannel * P.maxnhar_e; e += blockDim.x) {
    const int c = e / P.maxnhar_e, k = e - c * P.maxnhar_e;
    if(k < P.enhar[fi * P.nchannel + c]) {
      float* p = P.ephse + (fi * P.nchannel + c) * P.maxnhar_e + k;
      *p = phase_shifted(*p, theta, k);
    }
  }
  if(P.vsphse && P.nvs) {
    const int nv = min(P.nvs[fi], P.maxnhar);
    for(int k = threadIdx.x; k < nv; k += blockDim.x) {
      float* p = P.vsphse + fi * P.maxnhar + k;
      *p = phase_shifted(*p, theta, k);
    }
  }
}

// returns 0; theta is a [B][F] float scratch buffer on the device
static inline int run_phase_op(PhaseParams P, cudaStream_t st, LaunchCounter* lc) {
  LLSM_LAUNCH(phase_theta_kernel, dim3((P.nutt + 63) / 64), dim3(64), 0, st, P);
  LLSM_LAUNCH(phase_shift_kernel, dim3(P.nfrm, P.nutt), dim3(128), 0, st, P);
  if(lc) lc->n += 2;
  return 0;
}
