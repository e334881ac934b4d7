// Frame interpolation / time-stretch of a layer-1 batch: the step between
// llsm_chunk_tolayer1 + phasepropagate(-1) and llsm_chunk_tolayer0 + phasepropagate(+1) in concatenative and
// time-scaling use. The reference keeps it in a demo, not in the library:
//   test/demo-stretch.c:6-14    linterpc           circular interpolation of two phases
//   test/demo-stretch.c:16-44   interp_nmframe     noise model: psd, edc, envelope harmonics
//   test/demo-stretch.c:50-129  interp_llsm_frame  f0, Rd, VSPHSE, VTMAGN with the voiced / unvoiced cases
//   test/demo-stretch.c:169-185 the frame loop     out[i] = copy(frames[base]) blended towards frames[base + 1],
//                                                  PSDRES taken from a third frame
// One CTA per (output frame, utterance); every array row is contiguous, so the gather reads two source rows and writes
// one with coalesced accesses: HBM-bound, 2 rows in + 1 row out per output frame.
// Arithmetic follows the reference operation by operation: linterp in FP_TYPE = float without contraction, cos / sin /
// atan2 / log in double (the oracle's ciglet shim maps cos_2, sin_2, log_2 to libm), one float rounding per store.
#pragma once
#include "common.cuh"

struct StretchParams {
  int nutt, nfrm, nfrm_new, maxnhar, maxnhar_e, npsd, nchannel, nspec;
  int map_per_utt;            // 0: base / ratio / residx are [nfrm_new], 1: [B][nfrm_new]
  const int* base;            // source frame index (clamped to [0, nfrm - 2])
  const float* ratio;         // weight of frame base + 1
  const int* residx;          // frame the PSDRES row is taken from, NULL = base (clamped to [0, nfrm - 1])
  // source rows [B][nfrm]
  const float *f0, *rd, *vtmagn, *vsphse, *psd, *psdres, *edc, *eampl, *ephse, *ampl, *phse;
  const int *nvs, *enhar, *nhar;
  // destination rows [B][nfrm_new]
  float *o_f0, *o_rd, *o_vtmagn, *o_vsphse, *o_psd, *o_psdres, *o_edc, *o_eampl, *o_ephse, *o_ampl, *o_phse;
  int *o_nvs, *o_enhar, *o_nhar;
};

// ciglet's linterp macro on floats: a + (b - a) * r, every operation rounded (built without FMA contraction)
__device__ __forceinline__ float st_linterp(float a, float b, float r) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), r));
}
// test/demo-stretch.c:6-14
__device__ __forceinline__ float st_linterpc(float a, float b, float r) {
  double sa, ca, sb, cb;
  sincos((double)a, &sa, &ca); sincos((double)b, &sb, &cb);
  const float ax = (float)ca, ay = (float)sa, bx = (float)cb, by = (float)sb;
  const float cx = st_linterp(ax, bx, r), cy = st_linterp(ay, by, r);
  return (float)atan2((double)cy, (double)cx);
}
// mag2db(max(EPS, x)) with x a double (test/demo-stretch.c:46-47,119,122)
__device__ __forceinline__ float st_fade(double x) {
  const double m = x > 1e-8 ? x : 1e-8;            // max(EPS, x) is (EPS > x ? EPS : x)
  return (float)__dmul_rn(log(m), 20.0 / 2.3025851);
}

#define STRETCH_THREADS 256

__global__ void __launch_bounds__(STRETCH_THREADS) frames_stretch_kernel(StretchParams P) {
  const int i = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const size_t mi = P.map_per_utt ? (size_t)b * P.nfrm_new + i : (size_t)i;
  int base = P.base[mi];
  base = base < 0 ? 0 : (base > P.nfrm - 2 ? P.nfrm - 2 : base);
  int res = P.residx ? P.residx[mi] : base;
  res = res < 0 ? 0 : (res > P.nfrm - 1 ? P.nfrm - 1 : res);
  const float ratio = P.ratio[mi];
  const size_t d = (size_t)b * P.nfrm + base, s = d + 1, o = (size_t)b * P.nfrm_new + i;

  const float d_f0 = P.f0[d], s_f0 = P.f0[s];
  const bool dv = d_f0 > 0, sv = s_f0 > 0;
  const int dn = dv ? P.nvs[d] : 0, sn = sv ? P.nvs[s] : 0;     // VSPHSE is only attached to voiced frames
  const int nmax = dn > sn ? dn : sn, nmin = dn < sn ? dn : sn;

  if(tid == 0) {                                                // test/demo-stretch.c:75-97
    float f0, rd;
    if(dv && sv)      { f0 = st_linterp(d_f0, s_f0, ratio); rd = st_linterp(P.rd[d], P.rd[s], ratio); }
    else if(sv)       { f0 = s_f0; rd = P.rd[s]; }
    else if(dv)       { f0 = d_f0; rd = P.rd[d]; }
    else              { f0 = 0.f;  rd = 1.0f; }
    P.o_f0[o] = f0; P.o_rd[o] = rd;
    P.o_nvs[o] = (dv && sv) ? nmax : (sv ? sn : dn);
  }

  // ---- source phases and vocal-tract magnitudes (test/demo-stretch.c:92-125)
  {
    const float* dp = P.vsphse + d * P.maxnhar; const float* sp = P.vsphse + s * P.maxnhar;
    float* op = P.o_vsphse + o * P.maxnhar;
    const float* dm = P.vtmagn + d * P.nspec; const float* sm = P.vtmagn + s * P.nspec;
    float* om = P.o_vtmagn + o * P.nspec;
    if(dv && sv) {
      for(int k = tid; k < P.maxnhar; k += STRETCH_THREADS) {
        float v = 0.f;                                          // a fresh, zeroed array (container.c:36-40)
        if(k < nmin) v = st_linterpc(dp[k], sp[k], ratio);
        else if(k < nmax && dn < sn) v = sp[k];
        op[k] = v;
      }
      for(int k = tid; k < P.nspec; k += STRETCH_THREADS) {
        const float v = st_linterp(dm[k], sm[k], ratio);
        om[k] = v < -80.f ? -80.f : v;                          // max(-80, .)
      }
    } else if(sv || dv) {
      const float* vp = sv ? sp : dp; const float* vm = sv ? sm : dm;
      const int vn = sv ? sn : dn;
      const float fade = sv ? st_fade((double)ratio) : st_fade(1.0 - (double)ratio);
      for(int k = tid; k < P.maxnhar; k += STRETCH_THREADS) op[k] = k < vn ? vp[k] : 0.f;
      for(int k = tid; k < P.nspec; k += STRETCH_THREADS) {
        const float v = __fadd_rn(vm[k], fade);
        om[k] = v < -80.f ? -80.f : v;
      }
    } else {                                                    // unvoiced: neither array is attached
      for(int k = tid; k < P.maxnhar; k += STRETCH_THREADS) op[k] = 0.f;
      for(int k = tid; k < P.nspec; k += STRETCH_THREADS) om[k] = 0.f;
    }
  }

  // ---- noise model (test/demo-stretch.c:16-44)
  for(int k = tid; k < P.npsd; k += STRETCH_THREADS)
    P.o_psd[o * P.npsd + k] = st_linterp(P.psd[d * P.npsd + k], P.psd[s * P.npsd + k], ratio);
  if(P.o_psdres && P.psdres)                                    // test/demo-stretch.c:180-183
    for(int k = tid; k < P.npsd; k += STRETCH_THREADS)
      P.o_psdres[o * P.npsd + k] = P.psdres[((size_t)b * P.nfrm + res) * P.npsd + k];
  for(int c = tid; c < P.nchannel; c += STRETCH_THREADS) {
    const size_t dc = d * P.nchannel + c, sc = s * P.nchannel + c, oc = o * P.nchannel + c;
    P.o_edc[oc] = st_linterp(P.edc[dc], P.edc[sc], ratio);
    P.o_enhar[oc] = P.enhar[dc] > P.enhar[sc] ? P.enhar[dc] : P.enhar[sc];
  }
  for(int q = tid; q < P.nchannel * P.maxnhar_e; q += STRETCH_THREADS) {     // all channels in one pass
    const int c = q / P.maxnhar_e, k = q - c * P.maxnhar_e;
    const size_t dc = d * P.nchannel + c, sc = s * P.nchannel + c, oc = o * P.nchannel + c;
    const int de = P.enhar[dc], se = P.enhar[sc];
    const int emin = de < se ? de : se, emax = de > se ? de : se;
    float a = 0.f, p = 0.f;
    if(k < emin) {
      a = st_linterp(P.eampl[dc * P.maxnhar_e + k], P.eampl[sc * P.maxnhar_e + k], ratio);
      p = st_linterpc(P.ephse[dc * P.maxnhar_e + k], P.ephse[sc * P.maxnhar_e + k], ratio);
    } else if(k < emax) {                                       // the longer frame's own harmonics, unscaled
      const size_t e = (se > de ? sc : dc) * P.maxnhar_e + k;
      a = P.eampl[e]; p = P.ephse[e];
    }
    P.o_eampl[oc * P.maxnhar_e + k] = a; P.o_ephse[oc * P.maxnhar_e + k] = p;
  }

  // ---- the layer-0 harmonics ride along unchanged from frame `base` (llsm_copy_container, :176); they are
  //      recomputed by llsm_chunk_tolayer0 afterwards
  if(P.o_ampl && P.ampl && P.o_phse && P.phse && P.o_nhar && P.nhar) {
    const int nh = P.nhar[d];
    if(tid == 0) P.o_nhar[o] = nh;
    for(int k = tid; k < P.maxnhar; k += STRETCH_THREADS) {
      P.o_ampl[o * P.maxnhar + k] = k < nh ? P.ampl[d * P.maxnhar + k] : 0.f;
      P.o_phse[o * P.maxnhar + k] = k < nh ? P.phse[d * P.maxnhar + k] : 0.f;
    }
  }
}

static inline int run_frames_stretch(const StretchParams& P, cudaStream_t st, LaunchCounter* lc) {
  LLSM_LAUNCH(frames_stretch_kernel, dim3(P.nfrm_new, P.nutt), dim3(STRETCH_THREADS), 0, st, P);
  if(lc) lc->n += 1;
  return 0;
}
