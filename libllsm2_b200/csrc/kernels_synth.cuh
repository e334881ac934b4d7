// Layer-0 synthesis kernels for B200 (sm_100a).
//
//   hm_bank_ola_kernel   llsm_synthesize_harmonics_l0 (layer0.c:117-146) + the per-frame harmonic
//                        generator behind it (llsmutils.c:45-58, dsputils.c:328-351), Hann window
//                        and overlap-add fused; every output sample is written exactly once.
//   noise_template_kernel  llsm_generate_bandlimited_noise's chebyfilt (dsputils.c:385-394,51-70)
//   noise_excitation_kernel llsm_synthesize_noise_envelope (layer0.c:289-316) for all channels +
//                        the modulation/mix loop of llsm_synthesize_noise_excitation (:547-550)
//                        + stretch_stationary_noise (dsputils.c:363-383) as an index map
//   noise_shape_kernel   llsm_filter_noise (layer0.c:557-634) + the final y = y_sin + y_noise
//                        (layer0.c:657-659)
#pragma once
#include "common.cuh"
#include "kernels_iir.cuh"

// ------------------------------------------------------------------------------------------
// Harmonic bank + Hann + OLA
// ------------------------------------------------------------------------------------------
struct BankParams {
  int nfrm;                 // frame-array row length
  int maxnhar;              // ampl/phse row length
  const int* nfrm_utt;      // [B] or NULL
  const int* ny_utt;        // [B] or NULL: per-utterance valid output length
  const float* f0;          // [B][nfrm]
  const int* nhar;          // [B][nfrm]
  const float* ampl;        // [B][nfrm][maxnhar]
  const float* phse;
  const int* hm_base;       // [nfrm]  round(i * thop * fs)
  const float* hm_frac;     // [nfrm]  rawidx - baseidx
  const float* win;         // [n_hm]
  int n_hm;                 // window length (even)
  int ny;                   // valid output length when ny_utt == NULL
  int nsamp;                // samples to write per row (>= ny; the tail is zero-filled)
  int stride;               // output row stride
  float fs;
  int has_options;          // 0: options == NULL (always sinusoid bank, llsmutils.c:48-49)
  int use_iczt; float iczt_a, iczt_b;
  const int* frame_mask;    // [B][nfrm] optional: 0 = skip the frame (PbP path, layer0.c:261)
  int frame_lo, frame_hi;   // frames [lo, hi) contribute (frame-range sharding); hi <= 0 means all
  int npass;                // frame slots per CTA = warps * npass
  float hop;                // thop * fs (float product): hm_base[f] = round(f * hop)
  int residual_tc;          // options == NULL only: the tensor-core bank may serve this call (see launch_hm_bank)
  const float* sub_from; int sub_stride;   // optional [B][sub_stride]: write sub_from - y instead of y (analysis residual, layer0.c:500-501)
  float* y_sin;             // [B][stride]
  double iczt_nh;           // tensor-core bank: exp(log(n_hm) a + b), the harmonic count above which the ICZT branch is taken
};

#define BANK_KC 64          // harmonics staged per chunk

// NP packed float2 slots (two sample pairs each, FFMA2 / FMUL2) followed by NS scalar sample pairs
// (plain FFMA) per lane: lane l owns sample offsets n = l + 32 q, q < 2 NP + NS.
template <int NP, int NS, int NTHR, int MINB>
__global__ void __launch_bounds__(NTHR, MINB) hm_bank_ola_kernel(BankParams P) {
  LLSM_DYN_SMEM(smem);
  const int NW = blockDim.x >> 5;
  const int nslot = NW * P.npass;
  const int F = nslot - 2;                 // tiles owned by this CTA
  const int N = P.n_hm, H = N >> 1;
  const int npad = (N + 3) & ~3;
  float* fb = (float*)smem;                                   // [nslot][npad]
  float4* coef_all = (float4*)(fb + (size_t)nslot * npad);    // [NW][BANK_KC]
  int* sb = (int*)(coef_all + NW * BANK_KC);                  // [nslot] frame position
  int* sv = sb + nslot;                                       // [nslot] slot holds a frame

  const int b = blockIdx.y, seg = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int ny_b = P.ny_utt ? P.ny_utt[b] : P.ny;
  const int t0 = seg * F;
  // owned output range [start, end)
  const int start = t0 == 0 ? 0 : (t0 < nf ? P.hm_base[t0] : P.nsamp);
  const int end = (t0 + F < nf) ? P.hm_base[t0 + F] : P.nsamp;
  if(start >= end) return;                 // uniform per CTA

  float4* coef = coef_all + warp * BANK_KC;
  const size_t row = (size_t)b * P.nfrm;

  for(int q = 0; q < P.npass; q ++) {
    const int s = q * NW + warp;
    const int f = t0 - 1 + s;
    const bool inrange = f >= 0 && f < nf;
    float f0 = 0; int nh = 0;
    if(inrange) { f0 = P.f0[row + f]; nh = P.nhar[row + f]; }
    if(nh > 2048) nh = 2048;               // layer0.c:119,130
    bool voiced = inrange && f0 > 0 && nh > 0;        // layer0.c:125 (f0 == 0 skips the frame)
    if(voiced && P.frame_mask) voiced = P.frame_mask[row + f] != 0;
    if(P.frame_hi > 0 && (f < P.frame_lo || f >= P.frame_hi)) voiced = false;
    if(lane == 0) {
      sb[s] = inrange ? P.hm_base[f] : (f < 0 ? -(1 << 28) : (1 << 28));
      sv[s] = voiced ? 1 : 0;
    }
    if(! voiced) continue;                 // warp-uniform

    // ---- per-frame scalars, in the reference's precision (layer0.c:127-134, llsmutils.c:45-58,
    //      dsputils.c:338-348)
    const float f0n = f0 / P.fs;                                   // f0[i] / fs
    const float omega0 = (float)(2.0 * LLSM_PI * (double)f0n);     // FP_TYPE omega0 = 2 pi f0
    bool iczt = false;
    if(P.has_options && P.use_iczt)
      iczt = log((double)N) * (double)P.iczt_a < log((double)nh) - (double)P.iczt_b;
    if(iczt && nh > N - 1) nh = N - 1;     // the ICZT branch transforms bins 0..nx-1 only
    const double nu = iczt ? (double)omega0 / (2.0 * LLSM_PI) : (double)f0n;   // turns / sample
    const float frac = P.hm_frac[f];
    const float corr = (float)((double)(frac * 2.0f) * LLSM_PI / (double)P.fs * (double)f0);

    // ---- rotation seeds z = e^{i 2 pi nu n} for this lane's sample offsets n
    float2 zr[NP > 0 ? NP : 1], zi[NP > 0 ? NP : 1], nzi[NP > 0 ? NP : 1], wr[NP > 0 ? NP : 1],
           wi[NP > 0 ? NP : 1], C[NP > 0 ? NP : 1], S[NP > 0 ? NP : 1];
    float szr[NS > 0 ? NS : 1], szi[NS > 0 ? NS : 1], swr[NS > 0 ? NS : 1], swi[NS > 0 ? NS : 1],
          sC[NS > 0 ? NS : 1], sS[NS > 0 ? NS : 1];
#pragma unroll
    for(int m = 0; m < NS; m ++) {
      float2 a = unit_phasor_turns(nu * (double)(lane + 32 * (2 * NP + m)));
      szr[m] = a.x; szi[m] = a.y; swr[m] = 1.f; swi[m] = 0.f; sC[m] = 0.f; sS[m] = 0.f;
    }
#pragma unroll
    for(int m = 0; m < NP; m ++) {
      int nA = lane + 64 * m, nB = nA + 32;
      float2 a = unit_phasor_turns(nu * (double)nA);
      float2 bq = unit_phasor_turns(nu * (double)nB);
      zr[m] = make_float2(a.x, bq.x);  zi[m] = make_float2(a.y, bq.y);
      nzi[m] = make_float2(-a.y, -bq.y);
      wr[m] = make_float2(1.f, 1.f);   wi[m] = make_float2(0.f, 0.f);
      C[m] = make_float2(0.f, 0.f);    S[m] = make_float2(0.f, 0.f);
    }

    const float* ampl = P.ampl + (row + f) * (size_t)P.maxnhar;
    const float* phse = P.phse + (row + f) * (size_t)P.maxnhar;
    for(int kc = 0; kc < nh; kc += BANK_KC) {
      __syncwarp();
      // stage a_k cos(phi'_k), a_k sin(phi'_k) with phi'_k = phse[k] - corr (k+1)  (layer0.c:132)
      for(int kk = lane; kk < BANK_KC; kk += 32) {
        int k = kc + kk;
        float ca = 0.f, sa = 0.f;
        if(k < nh) {
          float a = ampl[k];
          float ph = (float)((double)phse[k] - (double)corr * ((double)k + 1.0));
          float s, c; sincosf(ph, &s, &c);
          ca = a * c; sa = a * s;
        }
        coef[kk] = make_float4(ca, ca, sa, sa);
      }
      __syncwarp();
      const int kn = min(BANK_KC, nh - kc);
      for(int kk = 0; kk < kn; kk ++) {
        const float4 cf = coef[kk];
        const float2 aa = make_float2(cf.x, cf.y), bb = make_float2(cf.z, cf.w);
#pragma unroll
        for(int m = 0; m < NP; m ++) {
          float2 t1 = fmul2(wr[m], zr[m]);
          float2 t2 = fmul2(wi[m], zr[m]);
          float2 nwr = ffma2(wi[m], nzi[m], t1);   // wr zr - wi zi
          float2 nwi = ffma2(wr[m], zi[m], t2);    // wi zr + wr zi
          wr[m] = nwr; wi[m] = nwi;
          C[m] = ffma2(aa, nwr, C[m]);             // sum a cos(phi) cos(k w n)
          S[m] = ffma2(bb, nwi, S[m]);             // sum a sin(phi) sin(k w n)
        }
#pragma unroll
        for(int m = 0; m < NS; m ++) {
          float nwr = fmaf(-swi[m], szi[m], swr[m] * szr[m]);
          float nwi = fmaf(swr[m], szi[m], swi[m] * szr[m]);
          swr[m] = nwr; swi[m] = nwi;
          sC[m] = fmaf(cf.x, nwr, sC[m]);
          sS[m] = fmaf(cf.z, nwi, sS[m]);
        }
      }
    }

    // ---- window and park the frame in shared memory: sample j = H +- n
    float* dst = fb + (size_t)s * npad;
#pragma unroll
    for(int m = 0; m < NP; m ++) {
#pragma unroll
      for(int h = 0; h < 2; h ++) {
        int n = lane + 64 * m + 32 * h;
        float c = h ? C[m].y : C[m].x, sn = h ? S[m].y : S[m].x;
        if(n <= H - 1) dst[H + n] = (c - sn) * P.win[H + n];
        if(n >= 1 && n <= H) dst[H - n] = (c + sn) * P.win[H - n];
      }
    }
#pragma unroll
    for(int m = 0; m < NS; m ++) {
      int n = lane + 32 * (2 * NP + m);
      if(n <= H - 1) dst[H + n] = (sC[m] - sS[m]) * P.win[H + n];
      if(n >= 1 && n <= H) dst[H - n] = (sC[m] + sS[m]) * P.win[H - n];
    }
  }
  __syncthreads();

  // ---- overlap-add: gather the (<= 3) frames covering each owned sample, ascending frame order
  //      as in the reference's sequential y[idx] += yi[j] (layer0.c:135-140)
  float* yrow = P.y_sin + (size_t)b * P.stride;
  const float inv_hop = 1.0f / P.hop;
  for(int idx = start + (int)threadIdx.x; idx < end; idx += blockDim.x) {
    float acc = 0.f;
    if(idx < ny_b) {
      // first slot whose window end (sb + H) is beyond idx: frame positions are round(f * hop), so the
      // hop arithmetic gives it up to one frame; start one slot early (early frames fail j < N below)
      int lo = (int)((float)(idx - H) * inv_hop) - (t0 - 1) - 1;
      if(lo < 0) lo = 0;
      for(int s = lo; s < nslot; s ++) {
        int j = idx - sb[s] + H;
        if(j < 0) break;
        if(sv[s] && j < N) acc += fb[(size_t)s * npad + j];
      }
    }
    yrow[idx] = P.sub_from ? P.sub_from[(size_t)b * P.sub_stride + idx] - acc : acc;
  }
}

static inline size_t bank_smem_bytes(int nwarps, int npass, int n_hm) {
  int nslot = nwarps * npass;
  size_t npad = (n_hm + 3) & ~3;
  return (size_t)nslot * npad * 4 + (size_t)nwarps * BANK_KC * 16 + (size_t)nslot * 8 + 16;
}

// returns 0 on success, -1 when the window is too long for the specialisations below
template <int NP, int NS, int NTHR, int MINB>
static inline void launch_hm_bank_t(const BankParams& P, dim3 grid, size_t smem, cudaStream_t st) {
  auto kfn = hm_bank_ola_kernel<NP, NS, NTHR, MINB>;
#ifndef LLSM_EMU
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  LLSM_LAUNCH(kfn, grid, dim3(NTHR), smem, st, P);
}

#include "kernels_bank_tc.cuh"

// LLSM_BANK_TC=0 forces the CUDA-core bank (hm_bank_ola_kernel); default: tensor-core bank where it applies
static inline int bank_tc_enabled() {
  const char* e = getenv("LLSM_BANK_TC");     // read per launch: the tests switch between the two banks
  return e ? atoi(e) : 1;
}

static inline int bank_variant() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_BANK_VARIANT"); v = e ? atoi(e) : 0; }
  return v;
}

// LLSM_RESIDUAL_TC: 1 = the analysis residual always on the tensor-core bank, 0 = never, unset (-1) = the caller decides
// (driver_analysis.h: yes for the CZT estimator, no for peak picking)
static inline int residual_tc_enabled() {
  static int v = -2;
  if(v == -2) { const char* e = getenv("LLSM_RESIDUAL_TC"); v = e ? atoi(e) : -1; }
  return v;
}

// returns 0 on success, -1 when the window is too long for the specialisations below
static inline int launch_hm_bank(BankParams P, int nutt, int nfrm_max, cudaStream_t st) {
#ifndef LLSM_EMU
  // many harmonics: operand generation + tcgen05 GEMM (kernels_bank_tc.cuh, ~1e-7 of the frame amplitude, i.e. ~5e-9
  // absolute on speech-level frames). The analysis residual (options == NULL, x - x_sin: a small difference of large
  // numbers) takes it only when the caller allows (P.residual_tc): 5e-9 passes the residual's 1e-6 RMS bar and every bar
  // of the CZT estimator's outputs, but it is 1e-5 of the residual, which the sub-band envelope phases of the
  // peak-picking method feel (measured on B200: 1.8e-4 rad on envelope harmonics of 5e-5 amplitude) -- that method keeps
  // the direct summation, at 1.4 ms more per 409 600 frames.
  // Few harmonics or long windows: direct FP32 summation below
  if(bank_tc_enabled() && (P.has_options || P.residual_tc) && P.maxnhar >= 24 &&
     launch_hm_bank_tc(P, nutt, nfrm_max, st) == 0) return 0;
#endif
  const int NTHR = 256, NW = NTHR / 32;
  P.npass = 4;                              // 32 frame slots, 30 tiles per CTA
  const int F = NW * P.npass - 2;
  int nseg = (std::max(nfrm_max, 1) + F - 1) / F;
  dim3 grid(nseg, nutt);
  size_t smem = bank_smem_bytes(NW, P.npass, P.n_hm);
  if(smem > 220 * 1024) return -1;
  int pairs = (P.n_hm / 2 + 1 + 31) / 32;   // sample pairs per lane
  if(pairs <= 2)  { launch_hm_bank_t<1, 0, NTHR, 3>(P, grid, smem, st); return 0; }
  if(pairs <= 4)  { launch_hm_bank_t<2, 0, NTHR, 3>(P, grid, smem, st); return 0; }
  if(pairs <= 6)  { launch_hm_bank_t<3, 0, NTHR, 3>(P, grid, smem, st); return 0; }
  if(pairs <= 8) {
    switch(bank_variant()) {                // packed : scalar split of the 8 pairs (tuning knob)
      case 1: launch_hm_bank_t<0, 8, NTHR, 3>(P, grid, smem, st); break;
      case 2: launch_hm_bank_t<2, 4, NTHR, 3>(P, grid, smem, st); break;
      case 3: launch_hm_bank_t<3, 2, NTHR, 3>(P, grid, smem, st); break;
      case 4: launch_hm_bank_t<1, 6, NTHR, 3>(P, grid, smem, st); break;
      case 5: launch_hm_bank_t<2, 4, NTHR, 2>(P, grid, smem, st); break;
      default: launch_hm_bank_t<4, 0, NTHR, 3>(P, grid, smem, st); break;
    }
    return 0;
  }
  if(pairs <= 12) { launch_hm_bank_t<6, 0, NTHR, 1>(P, grid, smem, st); return 0; }
  if(pairs <= 16) { launch_hm_bank_t<8, 0, NTHR, 1>(P, grid, smem, st); return 0; }
  return -1;
}

// ------------------------------------------------------------------------------------------
// Noise template: white -> band-limited ("coloured") template per (utterance, channel)
// ------------------------------------------------------------------------------------------
struct ChanFiltDev { int nstage; double b[2][5]; double a[2][5]; };

// ------------------------------------------------------------------------------------------
// Noise excitation: per-channel envelope (harmonic bank of <= maxnhar_e terms + DC, Hann, OLA),
// modulation of the stretched band-limited templates, channel mix.
// ------------------------------------------------------------------------------------------
struct ExcParams {
  int nfrm, nchannel, maxnhar_e;
  const int* nfrm_utt; const int* ny_utt;
  const float* f0; const float* edc; const int* enhar; const float* eampl; const float* ephse;
  const float* env_r;       // [nfrm] (float)((i - 1) * thop * fs)
  const int* env_off;       // [nfrm] round(env_r)
  const int* env_contig;    // [nfrm] indices of the frame are off + j exactly
  const float* win_env;     // [n_env]
  int n_env;
  int ny, nsamp, stride;
  float fs;
  int has_options, use_iczt; float iczt_a, iczt_b;
  const float* colored;     // [B][nchannel][nt]
  int nt, ntemplate, tstride;   // tstride: row stride of colored
  unsigned chan_mask;       // bit c set when channel c exists (fmin < fs / 2)
  float hop;                // thop * fs (float product), spacing of env_off
  int samp_lo, samp_hi;     // only samples [lo, hi) are needed (frame-range sharding); hi <= 0 means all
  int tile_nq;              // tiles of EXC_THREADS samples per CTA (set by the launcher)
  float* y_exc;             // [B][stride]
};

#define EXC_THREADS 256                 // (512 measured slower: 3.68 ms vs 2.68 ms at C2)
#define EXC_FCHUNK 32                   // frame slots staged at a time
#define EXC_NQ_MAX 16                   // a CTA owns up to EXC_NQ_MAX tiles of EXC_THREADS consecutive samples

// stretch_stationary_noise (dsputils.c:363-383) as a closed-form index map: output position p reads
// template index `base`, cross-faded with template index `ii` (>= 0) at the 128-sample seams.
struct StretchIdx { int base, ii; };
__device__ __forceinline__ StretchIdx stretch_index(int nx, int ny, int p) {
  const int overlap = 128;
  const int period = nx - overlap;
  StretchIdx s; s.ii = -1;
  if(p < nx) {
    s.base = p;
    if(ny > nx && p >= nx - overlap) s.ii = p - (nx - overlap);
  } else {
    // copy segment after the first template: (p - nx) / period by a float reciprocal and one correction
    // (exact: p - nx < 2^24 and the estimate is within one of the quotient)
    int q = (int)((float)(p - nx) * (1.0f / (float)period));
    if(q * period > p - nx) q --;
    else if((q + 1) * period <= p - nx) q ++;
    int head = nx + q * period;
    int i = p - head;
    s.base = i + overlap;
    if(i >= period - overlap && head + period <= ny) s.ii = i - (period - overlap);
  }
  return s;
}
__device__ __forceinline__ float stretched_value(const float* __restrict__ x, StretchIdx s) {
  float base = x[s.base];
  if(s.ii >= 0) {
    const float r = (float)s.ii / 128.0f;
    float y = (float)((double)base * (1.0 - (double)r));
    y = y + x[s.ii] * r;
    float d = 2.0f * r; d = d * (r - 1.0f); d = d + 1.0f;
    base = (float)((double)y / sqrt((double)d));
  }
  return base;
}

// A CTA owns P.tile_nq consecutive tiles of EXC_THREADS samples (a thread: one sample of each tile, one after the other --
// the two-samples-at-once variant ran out of registers: 4.9 ms against 3.1 ms). The parameters of every frame that reaches
// the CTA's samples are staged ONCE for all tiles: with one 256-sample tile per CTA (the first version) the staging -- a
// dependent chain of global reads and a sincosf per envelope harmonic, then a block barrier -- was repeated for every 1.2
// frames of output and its stalls (barrier 2.6, long scoreboard 2.6 per issued instruction, profiles/r2t) were the kernel.
template <int MAXCH>
__global__ void __launch_bounds__(EXC_THREADS) noise_excitation_kernel(ExcParams P) {
  LLSM_DYN_SMEM(smem);
  const int b = blockIdx.y;
  const int nq_all = P.tile_nq > 0 ? P.tile_nq : 1;
  const int span = nq_all * EXC_THREADS;
  const int p0 = blockIdx.x * span;
  if(P.samp_hi > 0 && (p0 + span <= P.samp_lo || p0 >= P.samp_hi)) {   // CTA outside the shard (uniform)
    for(int q = 0; q < nq_all; q ++) {
      const int p = p0 + q * EXC_THREADS + (int)threadIdx.x;
      if(p < P.nsamp) P.y_exc[(size_t)blockIdx.y * P.stride + p] = 0.f;
    }
    return;
  }
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int ny_b = P.ny_utt ? P.ny_utt[b] : P.ny;
  const int nch = P.nchannel, mne = P.maxnhar_e;
  // (tried: the channels of one harmonic interleaved, two 16-byte reads per harmonic instead of four 8-byte ones --
  //  6 % slower on B200, 2.83 against 2.68 ms: the broadcast reads cost the same wavefronts either way)
  const int fstride = 4 + nch * (2 + 2 * mne);     // floats per staged frame
  float* fr = (float*)smem;                        // [EXC_FCHUNK][fstride]
  const size_t row = (size_t)b * P.nfrm;

  // frames that can reach [p0, p0 + span): env_off + n_env + 1 > p0 and env_off - 1 <= pend, with
  // env_off[i] = round((i - 1) * hop): bounds from the hop arithmetic (one frame of slack), then made exact
  // with a couple of table reads. One thread does the search for the CTA.
  int* srange = (int*)(fr + EXC_FCHUNK * fstride);
  const float inv_hop = 1.0f / P.hop;
  if(threadIdx.x == 0) {
    const int pend = p0 + span - 1;
    int a = (int)floorf((float)(p0 - P.n_env - 1) * inv_hop);
    int bb = (int)floorf((float)(pend + 1) * inv_hop) + 3;
    if(a < 0) a = 0;
    if(bb > nf) bb = nf;
    if(a > bb) a = bb;
    while(a < bb && P.env_off[a] + P.n_env + 1 <= p0) a ++;
    while(bb > a && P.env_off[bb - 1] - 1 > pend) bb --;
    srange[0] = a; srange[1] = bb;
  }
  __syncthreads();
  const int ia = srange[0], ib = srange[1];
  const int half = P.n_env / 2;
  const bool single = ib - ia <= EXC_FCHUNK;       // every frame of the CTA fits the staging slots (the launcher sees to it)

  // ---- stage the parameters of frames [i0, i0 + nfc) (slot e -> frame e / 64, field e % 64 when a frame fits 64 fields)
  auto stage = [&](int i0, int nfc) {
    const bool pow2 = fstride <= 64;
    for(int e = threadIdx.x; e < nfc * (pow2 ? 64 : fstride); e += blockDim.x) {
      int fi, q;
      if(pow2) { fi = e >> 6; q = e & 63; if(q >= fstride) continue; }
      else { fi = e / fstride; q = e - fi * fstride; }
      int i = i0 + fi;
      float f0 = P.f0[row + i];
      float v;
      if(q == 0) v = f0 > 0 ? f0 / P.fs : 0.f;               // f0[i] / fs (0 when unvoiced)
      else if(q == 1) v = P.env_r[i];
      else if(q == 2) v = __int_as_float(P.env_off[i]);
      else if(q == 3) v = __int_as_float(P.env_contig[i]);
      else {
        int qq = q - 4, c = 0;
        const int per = 2 + 2 * mne;
        while(qq >= per) { qq -= per; c ++; }                  // at most nchannel - 1 steps
        const int w = qq;
        size_t ec = (row + i) * nch + c;
        int nh = f0 > 0 ? P.enhar[ec] : 0;                   // layer0.c:298 unvoiced -> 0 harmonics
        if(nh > mne) nh = mne;
        if(w == 0) v = P.edc[ec];
        else if(w == 1) v = __int_as_float(nh);
        else {
          int k = (w - 2) >> 1;
          if(k < nh) {
            float a = P.eampl[ec * mne + k], ph = P.ephse[ec * mne + k];
            float s, co; sincosf(ph, &s, &co);
            v = ((w - 2) & 1) ? a * s : a * co;
          } else v = 0.f;
        }
      }
      fr[fi * fstride + q] = v;
    }
  };
  // ---- envelope contributions of the staged frames [i0, i0 + nfc) to sample p (frames ascend in position)
  auto accumulate = [&](int p, int i0, int nfc, float (&env)[MAXCH]) {
    int fi0 = (int)floorf((float)(p - P.n_env - 2) * inv_hop) - i0;   // a frame before the first that can reach p
    if(fi0 < 0) fi0 = 0;
    for(int fi = fi0; fi < nfc; fi ++) {
      const float* F = fr + fi * fstride;
      const float f0n = F[0], r = F[1];
      const int off = __float_as_int(F[2]);
      if(off - 1 > p) break;                                 // this frame and every later one start beyond p
      const unsigned rel = (unsigned)(p - off + 1);
      if(rel > (unsigned)(P.n_env + 1)) continue;            // frame cannot reach this sample
      const int contig = __float_as_int(F[3]);
      for(int dj = contig ? 0 : -1; dj <= (contig ? 0 : 1); dj ++) {
        int j = p - off + dj;
        if(j < 0 || j >= P.n_env) continue;
        if(! contig) {
          float tpos = __fadd_rn(r, (float)j);               // (i - 1) * thop * fs + j in float
          int idx = (int)roundf(tpos);                       // layer0.c:307
          if(idx != p) continue;
        }
        const float wj = P.win_env[j];
        float2 z = unit_phasor_small(f0n * (float)(j - half));   // at most ~one turn across the window
        float2 w = make_float2(1.f, 0.f);
        float hs[MAXCH];
#pragma unroll
        for(int c = 0; c < MAXCH; c ++) hs[c] = 0.f;
        for(int k = 0; k < mne; k ++) {
          w = cmul(w, z);
#pragma unroll
          for(int c = 0; c < MAXCH; c ++) if(c < nch) {
            const float2 ab = ((const float2*)(F + 4 + c * (2 + 2 * mne) + 2))[k];
            hs[c] = fmaf(ab.x, w.x, fmaf(-ab.y, w.y, hs[c]));
          }
        }
#pragma unroll
        for(int c = 0; c < MAXCH; c ++) if(c < nch) {
          float v = hs[c] + F[4 + c * (2 + 2 * mne)];
          if(! (v > 1e-8f)) v = 1e-8f;                       // layer0.c:304
          env[c] += v * wj;                                  // layer0.c:306,309
        }
      }
    }
  };

  if(single) { stage(ia, ib - ia); __syncthreads(); }
#pragma unroll 1
  for(int q = 0; q < nq_all; q ++) {
    const int p = p0 + q * EXC_THREADS + (int)threadIdx.x;
    float env[MAXCH];
#pragma unroll
    for(int c = 0; c < MAXCH; c ++) env[c] = 0.f;
    const bool livep = p < P.nsamp && p < ny_b;
    if(single) {
      if(livep) accumulate(p, ia, ib - ia, env);
    } else {
      for(int i0 = ia; i0 < ib; i0 += EXC_FCHUNK) {           // (more frames than slots: tiny hops)
        const int nfc = min(EXC_FCHUNK, ib - i0);
        __syncthreads();
        stage(i0, nfc);
        __syncthreads();
        if(livep) accumulate(p, i0, nfc, env);
      }
    }
    if(p < P.nsamp) {
      float y = 0.f;
      if(p < ny_b) {
        const StretchIdx si = stretch_index(P.ntemplate, ny_b, p);
#pragma unroll
        for(int c = 0; c < MAXCH; c ++) if(c < nch && ((P.chan_mask >> c) & 1u)) {
          const float* tp = P.colored + ((size_t)b * nch + c) * P.tstride;
          float x = stretched_value(tp, si);
          x = x * sqrtf(env[c]);                             // layer0.c:548
          y += x;                                            // layer0.c:549
        }
      }
      P.y_exc[(size_t)b * P.stride + p] = y;
    }
  }
}

static inline int exc_tile_nq() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_EXC_NQ"); v = e ? atoi(e) : EXC_NQ_MAX; if(v < 1) v = 1; if(v > EXC_NQ_MAX) v = EXC_NQ_MAX; }
  return v;
}

static inline size_t exc_smem_bytes(int nchannel, int maxnhar_e) {
  return (size_t)EXC_FCHUNK * (4 + nchannel * (2 + 2 * maxnhar_e)) * 4 + 32;
}

static inline int launch_noise_excitation(const ExcParams& Pin, int nutt, cudaStream_t st) {
  ExcParams P = Pin;
  // tiles per CTA: as many as keep the frames reaching a CTA's samples within the staging slots
  int nq = exc_tile_nq();
  while(nq > 1 && (float)(nq * EXC_THREADS + P.n_env) / P.hop + 3.0f > (float)EXC_FCHUNK) nq --;
  P.tile_nq = nq;
  dim3 grid((P.nsamp + nq * EXC_THREADS - 1) / (nq * EXC_THREADS), nutt), block(EXC_THREADS);
  size_t smem = exc_smem_bytes(P.nchannel, P.maxnhar_e);
  if(smem > 200 * 1024) return -1;
#ifndef LLSM_EMU
  if(smem > 48 * 1024) {
    cudaFuncSetAttribute(noise_excitation_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(noise_excitation_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
#endif
  if(P.nchannel <= 4) LLSM_LAUNCH(noise_excitation_kernel<4>, grid, block, smem, st, P);
  else                LLSM_LAUNCH(noise_excitation_kernel<8>, grid, block, smem, st, P);
  return 0;
}

// ------------------------------------------------------------------------------------------
// Noise shaping: per frame STFT -> PSD -> gain from the model PSD -> ISTFT -> OLA, + final mix.
// One CTA owns a run of output samples of one utterance and walks, in ascending order, every
// frame whose nfft-long output touches the run, accumulating in shared memory: no atomics, each
// output sample is written once, and the summation order equals the reference's frame loop.
// ------------------------------------------------------------------------------------------
struct ShapeParams {
  int nfrm, npsd;
  const int* nfrm_utt; const int* ny_utt;
  const float* psd; const float* psdres;    // [B][nfrm][npsd]; psdres may be NULL
  const int* center;        // [nfrm] round(i * thop * fs)
  const float* win;         // [n_ns]
  int n_ns, nfft, lg_nfft, nspec;
  float wsqr, fs;
  const int* psd_lo; const float* psd_r;    // [nspec - 1] interpolation plan
  const float2* tw;         // [nfft] exp(-2 pi i m / nfft)
  const float* y_exc;       // [B][stride_exc]
  int stride_exc;
  const float* y_sin;       // [B][stride] (may be NULL: y = y_noise)
  float* y_noise; float* y; // [B][stride]; y may be NULL
  int ny, nsamp, stride;
  int seg;                  // output samples owned by a CTA
  int frame_lo, frame_hi;   // frames [lo, hi) are filtered (frame-range sharding); hi <= 0 means all
};

#define SHAPE_THREADS 256

// Two consecutive frames share one complex FFT: frame A rides in the real part, frame B in the
// imaginary part; their spectra are separated with the Hermitian symmetry, filtered with their own
// gains, re-packed (both filtered spectra are Hermitian, so the inverse transform returns A's
// output in the real part and B's in the imaginary part).
__global__ void __launch_bounds__(SHAPE_THREADS) noise_shape_kernel(ShapeParams P) {
  LLSM_DYN_SMEM(smem);
  const int nfft = P.nfft, nspec = P.nspec, npsd = P.npsd;
  float2* bufa = (float2*)smem;                    // [nfft]
  float2* bufb = bufa + nfft;                      // [nfft]
  float* acc = (float*)(bufb + nfft);              // [seg]
  float* spsd = acc + P.seg;                       // [2][npsd]
  float* pbuf = spsd + 2 * npsd;                   // [2][nspec]
  float* wmax = pbuf + 2 * nspec;                  // [2][32]

  const int b = blockIdx.y;
  const int oa = blockIdx.x * P.seg;
  const int ob = min(oa + P.seg, P.nsamp);
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int ny_b = P.ny_utt ? P.ny_utt[b] : P.ny;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int half = nfft >> 1, hw = P.n_ns / 2;
  const size_t row = (size_t)b * P.nfrm;
  const float* exc = P.y_exc + (size_t)b * P.stride_exc;

  for(int i = tid; i < P.seg; i += nth) acc[i] = 0.f;

  // frames with [center - half, center + half) intersecting [oa, ob)
  int lo = 0, hi = nf;
  while(lo < hi) { int mid = (lo + hi) >> 1; if(P.center[mid] + half > oa) hi = mid; else lo = mid + 1; }
  const int ia = lo;
  lo = ia; hi = nf;
  while(lo < hi) { int mid = (lo + hi) >> 1; if(P.center[mid] - half >= ob) hi = mid; else lo = mid + 1; }
  int ib = lo;
  int ia_ = ia;
  if(P.frame_hi > 0) { if(ia_ < P.frame_lo) ia_ = P.frame_lo; if(ib > P.frame_hi) ib = P.frame_hi; }
  const double resbias = 0.375 / 2.3025851 * 10.0;   // LOG2IN(LOGRESBIAS), constants.h:5,12
  const float inv = 1.0f / (float)nfft;

  for(int i = ia_; i < ib; i += 2) {
    const bool hasB = i + 1 < ib;
    __syncthreads();
    // ---- model PSD (+ residual) and its peak for both frames (layer0.c:584-585,598-601)
    for(int h = 0; h < 2; h ++) {
      float mx = -3.0e38f;
      if(h == 0 || hasB) {
        const float* psd = P.psd + (row + i + h) * (size_t)npsd;
        const float* res = P.psdres ? P.psdres + (row + i + h) * (size_t)npsd : nullptr;
        for(int j = tid; j < npsd; j += nth) {
          float v = psd[j];
          mx = fmaxf(mx, v);
          if(res) v = (float)((double)v + ((double)res[j] - resbias));
          spsd[h * npsd + j] = v;
        }
      }
      for(int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if((tid & 31) == 0) wmax[h * 32 + (tid >> 5)] = mx;
    }
    // ---- windowed frames, centred in the FFT buffer (layer0.c:588-592): A real, B imaginary
    const int cA = P.center[i], cB = hasB ? P.center[i + 1] : 0;
    for(int j = tid; j < nfft; j += nth) {
      int jj = j - half + hw;
      float va = 0.f, vb = 0.f;
      if(jj >= 0 && jj < P.n_ns) {
        float w = P.win[jj];
        int ia2 = cA + jj - hw;
        if(ia2 >= 0 && ia2 < ny_b) va = exc[ia2] * w;
        if(hasB) { int ib2 = cB + jj - hw; if(ib2 >= 0 && ib2 < ny_b) vb = exc[ib2] * w; }
      }
      bufa[j] = make_float2(va, vb);
    }
    __syncthreads();
    float pkA = wmax[0], pkB = wmax[32];
    for(int w = 1; w < (nth >> 5); w ++) { pkA = fmaxf(pkA, wmax[w]); pkB = fmaxf(pkB, wmax[32 + w]); }
    const bool doA = ! (pkA < -100.f);              // -100 dB floor, layer0.c:585
    const bool doB = hasB && ! (pkB < -100.f);
    if(! doA && ! doB) continue;                    // uniform

    float2* X = block_fft<false>(bufa, bufb, P.lg_nfft, P.tw, nfft);
    float2* Y = (X == bufa) ? bufb : bufa;

    // ---- separate the two spectra; PSDs (dsputils.c:237-244). A_k stays in X[k], B_k in X[nfft-k].
    for(int k = tid; k <= half; k += nth) {
      float2 zk = X[k];
      float2 A, B;
      if(k == 0 || k == half) { A = make_float2(zk.x, 0.f); B = make_float2(zk.y, 0.f); X[k] = make_float2(zk.x, zk.y); }
      else {
        float2 zn = X[nfft - k];
        A = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        B = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
        X[k] = A; X[nfft - k] = B;
      }
      pbuf[k] = (A.x * A.x + A.y * A.y) / P.wsqr;
      pbuf[nspec + k] = (B.x * B.x + B.y * B.y) / P.wsqr;
    }
    __syncthreads();
    // ---- gains: target PSD / smoothed measured PSD (layer0.c:597-612); re-pack W = A' + i B'
    for(int k = tid; k < nspec - 1; k += nth) {
      int l = max(0, k - 3), u = min(nspec - 1, k + 3);
      float smA = 0.f, smB = 0.f;
      for(int q = l; q <= u; q ++) { smA += pbuf[q]; smB += pbuf[nspec + q]; }
      const float cnt = (float)(u - l + 1);
      int pl = P.psd_lo[k]; float pr = P.psd_r[k];
      float hA = spsd[pl], hB = spsd[npsd + pl];
      if(pr != 0.f) { hA = hA + (spsd[pl + 1] - hA) * pr; hB = hB + (spsd[npsd + pl + 1] - hB) * pr; }
      float HA = doA ? expf(hA * (2.3025851f / 20.0f)) / sqrtf(smA / cnt * 44100.f / P.fs + 1e-8f) : 0.f;
      float HB = doB ? expf(hB * (2.3025851f / 20.0f)) / sqrtf(smB / cnt * 44100.f / P.fs + 1e-8f) : 0.f;
      float2 A, B;
      if(k == 0) { float2 z0 = X[0]; A = make_float2(z0.x * HA, 0.f); B = make_float2(z0.y * HB, 0.f); }
      else { float2 a0 = X[k], b0 = X[nfft - k]; A = make_float2(a0.x * HA, a0.y * HA); B = make_float2(b0.x * HB, b0.y * HB); }
      Y[k] = make_float2(A.x - B.y, A.y + B.x);
      if(k > 0) Y[nfft - k] = make_float2(A.x + B.y, B.x - A.y);   // conj(A') + i conj(B')
      if(k == nspec - 2) Y[half] = make_float2(A.x, B.x);            // Nyquist bin copies bin nspec-2
    }
    __syncthreads();
    float2* T = block_fft<true>(Y, X, P.lg_nfft, P.tw, nfft);
    // ---- scale, fades (layer0.c:616-619), accumulate in frame order (layer0.c:620-624)
    for(int h = 0; h < 2; h ++) {
      if(h == 1) __syncthreads();
      if(h == 0 ? ! doA : ! doB) continue;
      const int center = h == 0 ? cA : cB;
      for(int j = tid; j < nfft; j += nth) {
        float v = (h == 0 ? T[j].x : T[j].y) * inv;
        if(j < 16) v *= (float)j / 16.f;
        if(j >= nfft - 16) v = (float)((double)v * (1.0 - (double)((float)(nfft - 1 - j) / 16.f)));
        int idx = center + j - half;
        if(idx >= oa && idx < ob && idx < ny_b) acc[idx - oa] += v;
      }
    }
  }
  __syncthreads();
  for(int i = oa + tid; i < ob; i += nth) {
    float v = i < ny_b ? acc[i - oa] : 0.f;
    size_t o = (size_t)b * P.stride + i;
    P.y_noise[o] = v;
    if(P.y) P.y[o] = (P.y_sin ? P.y_sin[o] : 0.f) + v;       // layer0.c:658-659
  }
}

static inline size_t shape_smem_bytes(int nfft, int seg, int npsd, int nspec) {
  return (size_t)nfft * 16 + (size_t)(seg + 2 * npsd + 2 * nspec + 64) * 4 + 16;
}

static inline int launch_noise_shape_block(const ShapeParams& P, int nutt, cudaStream_t st) {
  dim3 grid((P.nsamp + P.seg - 1) / P.seg, nutt), block(SHAPE_THREADS);
  size_t smem = shape_smem_bytes(P.nfft, P.seg, P.npsd, P.nspec);
  if(smem > 220 * 1024) return -1;
#ifndef LLSM_EMU
  cudaFuncSetAttribute(noise_shape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
  LLSM_LAUNCH(noise_shape_kernel, grid, block, smem, st, P);
  return 0;
}

// per-utterance output length round((nfrm_b + 1) * thop * fs) with the reference's float
// products (layer0.c:643), for ragged batches
__global__ void ny_utt_kernel(const int* nfrm_utt, int nutt, float thop, float fs, int* ny_utt) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= nutt) return;
  float v = __fmul_rn((float)(nfrm_utt[b] + 1), thop);
  v = __fmul_rn(v, fs);
  ny_utt[b] = (int)round((double)v);
}

// ------------------------------------------------------------------------------------------
// Noise shaping, warp-per-frame-pair version for nfft = 1024: every warp filters two consecutive
// frames with one register-resident complex FFT / IFFT (warp_fft.cuh); a round = 8 warps = 16
// frames, after which the 16 frame outputs (parked in shared memory) are added to the CTA's output
// run in ascending frame order by sample-owning threads. Two block barriers per 16 frames.
// ------------------------------------------------------------------------------------------
#include "warp_fft.cuh"

#define SHW_WARPS 8
#define SHW_THREADS (SHW_WARPS * 32)

__global__ void __launch_bounds__(SHW_THREADS, 2) noise_shape_warp_kernel(ShapeParams P) {
  LLSM_DYN_SMEM(smem);
  const int NF = 1024, HALF = 512, NSPEC = 513;
  float2* tw2 = (float2*)smem;                                   // [1024]
  char* wbase = (char*)(tw2 + 1024);
  float* acc = (float*)(wbase + SHW_WARPS * WFFT_SCRATCH_BYTES); // [seg]
  int* fcen = (int*)(acc + P.seg);                               // [2 * SHW_WARPS] frame centres of the round
  int* fval = fcen + 2 * SHW_WARPS;                              // [2 * SHW_WARPS] frame produced output
  float2* z0 = (float2*)(fval + 2 * SHW_WARPS) + 32 * (threadIdx.x >> 5);   // [SHW_WARPS][32]

  const int b = blockIdx.y;
  const int oa = blockIdx.x * P.seg;
  const int ob = min(oa + P.seg, P.nsamp);
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int ny_b = P.ny_utt ? P.ny_utt[b] : P.ny;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npsd = P.npsd, hw = P.n_ns / 2;
  const size_t row = (size_t)b * P.nfrm;
  const float* exc = P.y_exc + (size_t)b * P.stride_exc;
  float2* scratch = (float2*)(wbase + warp * WFFT_SCRATCH_BYTES);
  float* slot = (float*)scratch;                                 // [2][1024] outputs of the pair
  float* pbuf = (float*)scratch;                                 // [2][NSPEC] PSDs (between FFTs)
  float* spsd = pbuf + 2 * NSPEC + 2;                            // [2][npsd]  (npsd <= 512)

  wfft_build_tw2(tw2, P.tw);
  for(int i = tid; i < P.seg; i += blockDim.x) acc[i] = 0.f;

  int lo = 0, hi = nf;
  while(lo < hi) { int mid = (lo + hi) >> 1; if(P.center[mid] + HALF > oa) hi = mid; else lo = mid + 1; }
  const int ia = lo;
  lo = ia; hi = nf;
  while(lo < hi) { int mid = (lo + hi) >> 1; if(P.center[mid] - HALF >= ob) hi = mid; else lo = mid + 1; }
  int ib = lo;
  int ia_ = ia;
  if(P.frame_hi > 0) { if(ia_ < P.frame_lo) ia_ = P.frame_lo; if(ib > P.frame_hi) ib = P.frame_hi; }
  const double resbias = 0.375 / 2.3025851 * 10.0;
  const float inv = 1.0f / 1024.0f;
  __syncthreads();

  const float psc = 44100.f / P.fs, psc7 = psc / 7.f;
  const float inv_wsqr = 1.0f / P.wsqr;
  for(int r0 = ia_; r0 < ib; r0 += 2 * SHW_WARPS) {
    const int i = r0 + 2 * warp;                 // this warp's pair (i, i + 1)
    const bool hasA = i < ib, hasB = i + 1 < ib;
    bool doA = false, doB = false;
    const int cA = hasA ? P.center[i] : 0, cB = hasB ? P.center[i + 1] : 0;
    if(hasA) {
      // ---- peaks of the model PSDs (layer0.c:584-585)
      float mxA = -3.0e38f, mxB = -3.0e38f;
      const float* psdA = P.psd + (row + i) * (size_t)npsd;
      const float* psdB = P.psd + (row + i + (hasB ? 1 : 0)) * (size_t)npsd;
      for(int j0 = 0; j0 < npsd; j0 += 256) {
        float pa[8], pb[8];
#pragma unroll
        for(int u = 0; u < 8; u ++) {
          const int j = j0 + lane + 32 * u;
          pa[u] = j < npsd ? psdA[j] : -3.0e38f;
          pb[u] = (hasB && j < npsd) ? psdB[j] : -3.0e38f;
        }
#pragma unroll
        for(int u = 0; u < 8; u ++) { mxA = fmaxf(mxA, pa[u]); mxB = fmaxf(mxB, pb[u]); }
      }
      for(int o = 16; o > 0; o >>= 1) {
        mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, o));
        mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, o));
      }
      doA = ! (mxA < -100.f);
      doB = hasB && ! (mxB < -100.f);
    }
    // ---- warm the caches: the model PSDs of this pair are needed only after the first FFT, the excitation
    //      of the pair this warp takes next round right at its start (one 128-byte line per lane and array)
    if(hasA) {
      const int nline = (npsd * 4 + 127) >> 7;
      for(int h = 0; h < (hasB ? 2 : 1); h ++) {
        if(lane < nline) {
          prefetch_l2(P.psd + (row + i + h) * (size_t)npsd + lane * 32);
          if(P.psdres) prefetch_l2(P.psdres + (row + i + h) * (size_t)npsd + lane * 32);
        }
      }
      const int inext = i + 2 * SHW_WARPS;
      if(inext < ib) {
        const int c0 = P.center[inext] - hw;
        for(int q = lane * 32; q < P.n_ns + 256; q += 32 * 32) {
          const int at = c0 + q;
          if(at >= 0 && at < ny_b) prefetch_l2(exc + at);
        }
      }
    }
    if(doA || doB) {                             // warp-uniform
      float2 x[32];
      const int plane = (32 - lane) & 31;
#pragma unroll 1
      for(int pass = 0; pass < 2; pass ++) {     // 0: analysis FFT, 1: synthesis IFFT (shared FFT code)
        if(pass == 0) {
          // ---- load the two windowed frames (layer0.c:588-592) through shared memory (rolled loop:
          //      small code), then pick element j = lane + 32 r into register r
          //      (eight iterations in flight: the loop is bound by the latency of the excitation reads)
          //      (rows that lie wholly outside the analysis window are zero-filled without any read)
          const int row_lo = (HALF - hw) >> 5, row_hi = (HALF - hw + P.n_ns - 1) >> 5;
#pragma unroll 1
          for(int r0 = 0; r0 < 32; r0 += 8) {
            if(r0 + 7 < row_lo || r0 > row_hi) {             // warp-uniform
#pragma unroll
              for(int u = 0; u < 8; u ++) scratch[lane + 32 * (r0 + u)] = make_float2(0.f, 0.f);
              continue;
            }
            float va[8], vb[8], wv[8];
#pragma unroll
            for(int u = 0; u < 8; u ++) {
              const int jj = lane + 32 * (r0 + u) - HALF + hw;
              const bool in = jj >= 0 && jj < P.n_ns;
              const int ia2 = cA + jj - hw, ib2 = cB + jj - hw;
              wv[u] = in ? P.win[jj] : 0.f;
              va[u] = (in && doA && ia2 >= 0 && ia2 < ny_b) ? exc[ia2] : 0.f;
              vb[u] = (in && doB && ib2 >= 0 && ib2 < ny_b) ? exc[ib2] : 0.f;
            }
#pragma unroll
            for(int u = 0; u < 8; u ++) scratch[lane + 32 * (r0 + u)] = make_float2(va[u] * wv[u], vb[u] * wv[u]);
          }
          __syncwarp();
#pragma unroll
          for(int r = 0; r < 32; r ++) x[r] = scratch[lane + 32 * r];
          __syncwarp();
        }
        warp_fft1024(x, scratch, tw2, lane, pass);
        if(pass == 0) {
          // ---- spectra of the two frames from Z = FFT(a + i b): with the partner bin Zn = Z[N - m],
          //      A = (Z + conj Zn) / 2, B = (Z - conj Zn) / (2i). Bin m = lane + 32 k1; its partner sits
          //      in lane (32 - lane) % 32, register 31 - k1 (lane 0: its own register (32 - k1) % 32,
          //      read back from a small shared copy).
          if(lane == 0) {
#pragma unroll
            for(int k1 = 0; k1 < 32; k1 ++) z0[k1] = x[k1];
          }
          __syncwarp();
          // PSDs of bins 0..512 into shared memory (dsputils.c:237-244)
#pragma unroll
          for(int k1 = 0; k1 <= 16; k1 ++) {
            float px = __shfl_sync(0xffffffffu, x[31 - k1].x, plane);
            float py = __shfl_sync(0xffffffffu, x[31 - k1].y, plane);
            float2 zn = lane == 0 ? z0[(32 - k1) & 31] : make_float2(px, py);
            float2 zk = x[k1];
            int m = lane + 32 * k1;
            if(m <= HALF) {
              float ax = 0.5f * (zk.x + zn.x), ay = 0.5f * (zk.y - zn.y);
              float bx = 0.5f * (zk.y + zn.y), by = 0.5f * (zn.x - zk.x);
              pbuf[m] = (ax * ax + ay * ay) * inv_wsqr;
              pbuf[NSPEC + m] = (bx * bx + by * by) * inv_wsqr;
            }
          }
          // model PSD (+ residual) (layer0.c:598-601)
          for(int h = 0; h < 2; h ++) {
            if(h == 1 && ! hasB) break;
            const float* psd = P.psd + (row + i + h) * (size_t)npsd;
            const float* res = P.psdres ? P.psdres + (row + i + h) * (size_t)npsd : nullptr;
            for(int j0 = 0; j0 < npsd; j0 += 256) {          // eight reads per array in flight
              float pv[8], rv[8];
#pragma unroll
              for(int u = 0; u < 8; u ++) {
                const int j = j0 + lane + 32 * u;
                pv[u] = j < npsd ? psd[j] : 0.f;
                rv[u] = (res && j < npsd) ? res[j] : 0.f;
              }
#pragma unroll
              for(int u = 0; u < 8; u ++) {
                const int j = j0 + lane + 32 * u;
                float v = pv[u];
                if(res) v = (float)((double)v + ((double)rv[u] - resbias));
                if(j < npsd) spsd[h * npsd + j] = v;
              }
            }
          }
          __syncwarp();
          // ---- gains (layer0.c:597-605), in place over the PSDs: block t = bins 32 t .. 32 t + 31 is
          //      overwritten one iteration late, after block t + 1 has read its 3-bin margin
          {
            float hAp = 0.f, hBp = 0.f;
#pragma unroll 1
            for(int t = 0; t <= 16; t ++) {
              const int kk = lane + 32 * t;
              float HA = 0.f, HB = 0.f;
              if(t < 16) {
                int l = max(0, kk - 3), u = min(NSPEC - 1, kk + 3);
                float smA = 0.f, smB = 0.f;
                if(kk >= 3 && kk + 3 <= NSPEC - 1) {         // interior bins: fixed 7-tap sum, same order
                  const float* pa = pbuf + kk - 3; const float* pb = pbuf + NSPEC + kk - 3;
#pragma unroll
                  for(int q = 0; q < 7; q ++) { smA += pa[q]; smB += pb[q]; }
                } else {
                  for(int q = l; q <= u; q ++) { smA += pbuf[q]; smB += pbuf[NSPEC + q]; }
                }
                const float rc = (u - l == 6) ? psc7 : psc / (float)(u - l + 1);
                int pl = P.psd_lo[kk]; float pr = P.psd_r[kk];
                float hA = spsd[pl], hB = hasB ? spsd[npsd + pl] : 0.f;
                if(pr != 0.f) { hA = hA + (spsd[pl + 1] - hA) * pr; if(hasB) hB = hB + (spsd[npsd + pl + 1] - hB) * pr; }
                // 10^(h / 20) as 2^(h log2(10) / 20): one ex2 instead of expf's range reduction
                if(doA) HA = exp2f(hA * 0.16609640474f) * rsqrtf(smA * rc + 1e-8f);
                if(doB) HB = exp2f(hB * 0.16609640474f) * rsqrtf(smB * rc + 1e-8f);
              }
              __syncwarp();
              if(t >= 1) { pbuf[kk - 32] = hAp; pbuf[NSPEC + kk - 32] = hBp; }
              hAp = HA; hBp = HB;
            }
            __syncwarp();
          }
          // ---- re-packing W = HA A + i HB B; registers k1 and 31 - k1 are rewritten together so that
          //      every partner is read before it is overwritten
          float nyqA = 0.f, nyqB = 0.f;          // Re of the scaled bin 511 (lane 31, register 15)
#pragma unroll
          for(int k1 = 0; k1 < 16; k1 ++) {
            float p1x = __shfl_sync(0xffffffffu, x[31 - k1].x, plane), p1y = __shfl_sync(0xffffffffu, x[31 - k1].y, plane);
            float p2x = __shfl_sync(0xffffffffu, x[k1].x, plane), p2y = __shfl_sync(0xffffffffu, x[k1].y, plane);
#pragma unroll
            for(int half2 = 0; half2 < 2; half2 ++) {
              const int reg = half2 == 0 ? k1 : 31 - k1;
              float2 zn = half2 == 0 ? (lane == 0 ? z0[(32 - k1) & 31] : make_float2(p1x, p1y))
                                     : (lane == 0 ? z0[(k1 + 1) & 31] : make_float2(p2x, p2y));
              float2 zk = x[reg];
              int m = lane + 32 * reg;
              int kk = m <= HALF ? m : 1024 - m;
              float HA = kk < NSPEC - 1 ? pbuf[kk] : 0.f, HB = kk < NSPEC - 1 ? pbuf[NSPEC + kk] : 0.f;
              float2 a = make_float2(0.5f * (zk.x + zn.x) * HA, 0.5f * (zk.y - zn.y) * HA);
              float2 bq = make_float2(0.5f * (zk.y + zn.y) * HB, 0.5f * (zn.x - zk.x) * HB);
              if(m == 0) { a.y = 0.f; bq.y = 0.f; }
              if(reg == 15) { nyqA = a.x; nyqB = bq.x; }
              x[reg] = make_float2(a.x - bq.y, a.y + bq.x);
            }
          }
          // Nyquist bin (lane 0, register 16) copies the scaled bin 511 held by lane 31, register 15
          nyqA = __shfl_sync(0xffffffffu, nyqA, 31); nyqB = __shfl_sync(0xffffffffu, nyqB, 31);
          if(lane == 0) x[16] = make_float2(nyqA, nyqB);
          __syncwarp();
        } else {
          // ---- scale and park both outputs in the warp's slot; fades (layer0.c:616-619) touch only the
          //      first and last 16 samples
#pragma unroll
          for(int r = 0; r < 32; r ++) {
            int j = lane + 32 * r;
            slot[j] = x[r].x * inv; slot[1024 + j] = x[r].y * inv;
          }
          __syncwarp();
          if(lane < 16) {
            float g = (float)lane / 16.f;
            slot[lane] *= g; slot[1024 + lane] *= g;
            int j = NF - 16 + lane;
            double g2 = 1.0 - (double)((float)(NF - 1 - j) / 16.f);
            slot[j] = (float)((double)slot[j] * g2); slot[1024 + j] = (float)((double)slot[1024 + j] * g2);
          }
        }
      }
    }
    // slot table of the round: centre of every frame (ascending) and whether it produced output
    if(lane == 0) {
      fcen[2 * warp] = hasA ? cA : (1 << 30); fcen[2 * warp + 1] = hasB ? cB : (1 << 30);
      fval[2 * warp] = doA; fval[2 * warp + 1] = doB;
    }
    __syncthreads();
    // ---- add the round's frames to the output run, ascending frame order (layer0.c:620-624): slot by slot,
    //      every thread owning the run positions i = tid (mod blockDim) -- the same thread adds all frames to a
    //      given sample, so the order is the reference's and no barrier is needed between slots
    {
      const int nslot = min(2 * SHW_WARPS, ib - r0);
      const int run = min(ob, ny_b) - oa;                       // valid positions of the run
      for(int f = 0; f < nslot; f ++) {
        if(! fval[f]) continue;
        const float* src = (const float*)(wbase + (f >> 1) * WFFT_SCRATCH_BYTES) + (f & 1) * 1024;
        const int off = fcen[f] - HALF - oa;                    // run position of the frame's sample 0
        const int lo = max(off, 0), hi = min(off + NF, run);
        for(int i = lo + ((tid - lo) & (SHW_THREADS - 1)); i < hi; i += SHW_THREADS) acc[i] += src[i - off];
      }
    }
    __syncthreads();
  }
  // write-out (layer0.c:658-659); the harmonic part is read eight rows ahead: the loop is latency-bound
  for(int i0 = oa + tid; i0 < ob; i0 += 8 * SHW_THREADS) {
    float ys[8];
#pragma unroll
    for(int u = 0; u < 8; u ++) {
      const int i = i0 + u * SHW_THREADS;
      ys[u] = (P.y && P.y_sin && i < ob) ? P.y_sin[(size_t)b * P.stride + i] : 0.f;
    }
#pragma unroll
    for(int u = 0; u < 8; u ++) {
      const int i = i0 + u * SHW_THREADS;
      if(i < ob) {
        const float v = i < ny_b ? acc[i - oa] : 0.f;
        const size_t o = (size_t)b * P.stride + i;
        P.y_noise[o] = v;
        if(P.y) P.y[o] = ys[u] + v;
      }
    }
  }
}

static inline size_t shape_warp_smem_bytes(int seg) {
  return (size_t)1024 * 8 + (size_t)SHW_WARPS * WFFT_SCRATCH_BYTES + (size_t)seg * 4 + 4 * SHW_WARPS * 4 +
         (size_t)SHW_WARPS * 32 * 8 + 16;
}

static inline int shape_variant() {
  static int v = -1;
  if(v < 0) { const char* e = getenv("LLSM_SHAPE_VARIANT"); v = e ? atoi(e) : 1; }
  return v;
}

static inline int launch_noise_shape(const ShapeParams& P, int nutt, cudaStream_t st) {
  if(shape_variant() == 1 && P.nfft == 1024 && P.npsd <= 512 && P.n_ns <= 1024) {
    dim3 grid((P.nsamp + P.seg - 1) / P.seg, nutt), block(SHW_THREADS);
    size_t smem = shape_warp_smem_bytes(P.seg);
#ifndef LLSM_EMU
    cudaFuncSetAttribute(noise_shape_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    LLSM_LAUNCH(noise_shape_warp_kernel, grid, block, smem, st, P);
    return 0;
  }
  return launch_noise_shape_block(P, nutt, st);
}
