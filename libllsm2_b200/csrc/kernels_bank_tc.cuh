// Harmonic bank + Hann + overlap-add on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as hm_bank_ola_kernel (kernels_synth.cuh): llsm_synthesize_harmonics_l0
// (layer0.c:117-146) with the per-frame generator behind it (llsmutils.c:45-58, dsputils.c:328-351);
// every output sample is written exactly once, frames are summed in ascending order.
//
// Factorisation. A frame is y[n] = sum_k a_k cos(k w n + phi_k), n in [-H, H). Write n = 8 p + q
// (q = 0..7) and use cos(-x) = cos(x):
//     y[ 8 p + q] = U[p][q] - V[p][q],   y[-8 p + q] = U[p][q] + V[p][q],   p = 0..31
//     U = Ar Br,  V = Ai Bi,   A[p][k] = e^{i k w 8 p}  (32 x K),   B[k][q] = a_k e^{i (k w q + phi_k)}  (K x 8)
// so a frame costs (32 + 8) K generated phasors and two small real GEMMs instead of 2 H K sinusoid
// samples. Four frames of one utterance share an M = 128 tcgen05.mma: frame g occupies tensor-memory
// lanes 32 g .. 32 g + 31 (rows of A) and columns 8 g .. 8 g + 7 of the N = 32 operand B; only the four
// diagonal 32 x 8 blocks of the product are read back (the tensor pipe has the slack: the kernel is bound
// by generating the operands). Accuracy: each operand is split x = hi + lo with hi = x truncated to TF32
// (what the hardware reads) and lo = x - hi, and three products hi hi + lo hi + hi lo are accumulated in
// FP32 -- 2^-21 relative per term, as good as the FP32 recurrences of the CUDA-core kernel.
//
// A is written from registers straight into tensor memory (tcgen05.st), B goes through shared memory
// (K-major core matrices, no swizzle), the accumulators live in tensor memory. One CTA per SM (all 512
// tensor-memory columns: a three-deep ring of A operands, 3 x 128 columns, and two sets of U / V accumulators,
// 2 x 64), 21 warps, warp specialised:
//   warps  0..7   A rows: two interleaved harmonic recurrences per thread, 16 harmonics of a 32-harmonic chunk,
//                 hi / lo split, tcgen05.st; afterwards the next group's coefficients a_k e^{i phi_k} (cfull[])
//   warps  8..15  B columns, two teams of four on alternate chunks, into a three-deep 16 KB ring
//   warps 16..19  read-back: tcgen05.ld of the finished accumulators (dfull[] / dempty[]), window, park the
//                 frame, overlap-add and write every output tile whose frames are complete
//   warp  20      issues the 24 MMAs of a chunk from one elected lane of the converged warp -- the operands are
//                 then warp-uniform and the MMAs leave at the tensor pipe's own rate, 16 cycles each; issued
//                 behind a plain `lane == 0` branch they cost 45 (tools/tc_rate.cu) -- and frees the ring with
//                 tcgen05.commit (empty[])
// full[] counts the eight A warps and the four B warps of a chunk. Phasor seeds: the turn count is reduced
// exactly in 64-bit fixed point, sine and cosine come from the special-function unit.
// BTC_STAMP is a tracing hook (tools/btc_trace.cu); it compiles to nothing in the library.
#pragma once
#ifndef LLSM_EMU
#include "tcgen05.cuh"
#ifndef BTC_STAMP
#define BTC_STAMP(who, tag) do { } while(0)
#define BTC_STAMP_DECL
#endif

#define BTC_GEN_WARPS 16
#define BTC_GEN_THREADS (BTC_GEN_WARPS * 32)
#define BTC_OUT_WARPS 4      // read-back / overlap-add warps (one per tensor-memory lane quadrant)
#define BTC_THREADS (BTC_GEN_THREADS + 32 * BTC_OUT_WARPS + 32)   // + the MMA-issuing warp
#define BTC_KC 32           // harmonics per MMA chunk (columns of each A tile)
#define BTC_SLOT 512        // floats per parked frame: sample n lives at index n + 256
#define BTC_NBUF 3          // operand ring depth: 3 x 128 tensor-memory columns of A, 3 x 16 KB of B
#define BTC_NSLOT 64        // frame slots per CTA (16 groups of 4), 62 owned output tiles
#define BTC_CST 512         // harmonics per frame in the coefficient buffers (more: CUDA-core bank)
#define BTC_MAXH 248        // largest half window: n = 8 p + q, p < 32

struct BtcFrame { unsigned long long nufix; float corr; int nh; };   // nufix = nu 2^64 (turns per sample and harmonic)

// e^{i 2 pi t nu}: the product is reduced exactly in 64-bit fixed point (the turn count wraps), the sine and
// cosine come from the special-function unit (absolute error 2^-21.4: below the TF32 split's 2^-21 per term)
__device__ __forceinline__ float2 btc_phasor(unsigned long long nufix, unsigned t) {
  const unsigned long long p = nufix * (unsigned long long)t;
  const float a = (float)(int)(unsigned)(p >> 32) * 1.4629180792671596e-9f;     // 2 pi / 2^32
  return make_float2(__cosf(a), __sinf(a));
}

__device__ __forceinline__ void btc_split2(float2 x, uint32_t& h0, uint32_t& h1, uint32_t& l0, uint32_t& l1) {
  h0 = __float_as_uint(x.x) & 0xffffe000u; h1 = __float_as_uint(x.y) & 0xffffe000u;
  float2 lo = ffma2(make_float2(__uint_as_float(h0), __uint_as_float(h1)), make_float2(-1.f, -1.f), x);
  l0 = __float_as_uint(lo.x); l1 = __float_as_uint(lo.y);
}
// (wr + i wi) (zr + i zi) on two packed elements
__device__ __forceinline__ void btc_rot2(float2& wr, float2& wi, float2 zr, float2 zi, float2 nzi) {
  const float2 t1 = fmul2(wr, zr), t2 = fmul2(wi, zr);
  const float2 nr = ffma2(wi, nzi, t1), ni = ffma2(wr, zi, t2);
  wr = nr; wi = ni;
}

// Overlap-add of one owned tile [base_f, base_f+1) by one warp: the frames f - 1 .. f + 2 can reach it
// (H = round(hop), positions round(f hop)); ascending frame order as in layer0.c:135-140.
__device__ __forceinline__ void btc_emit_tile(int s, int lane, int t0, int nf, int ny_b, int N, int H, int nsamp,
                                              const int* sb, const int* sv, const float* fb, float* yrow) {
  const int f = t0 - 1 + s;
  if(f >= nf) return;
  const int lo = f == 0 ? 0 : sb[s];
  const int hi = f + 1 < nf ? sb[s + 1] : nsamp;
  int off[4]; const float* src[4];
#pragma unroll
  for(int c = 0; c < 4; c ++) {
    const int sc = min(s - 1 + c, BTC_NSLOT - 1);
    const bool ok = (s - 1 + c < BTC_NSLOT) && sv[sc];
    off[c] = ok ? sb[sc] - H : (1 << 29);                     // j = idx - off; invalid slots fail j < N
    src[c] = fb + (size_t)sc * BTC_SLOT + 256 - H;
  }
#pragma unroll 2
  for(int idx = lo + lane; idx < hi; idx += 32) {
    float acc = 0.f;
    if(idx < ny_b) {
#pragma unroll
      for(int c = 0; c < 4; c ++) {
        const int j = idx - off[c];
        if((unsigned)j < (unsigned)N) acc += src[c][j];
      }
    }
    yrow[idx] = acc;
  }
}

// 24 MMAs of one chunk: D_U += Arh Brh + Arl Brh + Arh Brl, D_V += Aih Bih + Ail Bih + Aih Bil, four K = 8 slabs.
// BF (operand buffer) is a compile-time constant so that every descriptor is base + constant.
template <int BF>
__device__ __forceinline__ void btc_issue_chunk(uint32_t tbase, uint64_t bbase, uint32_t idesc, uint32_t d_u, uint32_t d_v, int c) {
  const uint32_t a0 = tbase + 128 * BF;
#pragma unroll
  for(int ks = 0; ks < 4; ks ++) {
    const uint64_t brh = bbase + (uint64_t)((16384 * BF + 256 * ks) >> 4), brl = brh + (4096 >> 4);
    const uint64_t bih = brh + (8192 >> 4), bil = brh + (12288 >> 4);
    tc::mma_tf32_ts(d_u, a0 + 8 * ks, brh, idesc, (c | ks) ? 1u : 0u);
    tc::mma_tf32_ts(d_v, a0 + 64 + 8 * ks, bih, idesc, (c | ks) ? 1u : 0u);
    tc::mma_tf32_ts(d_u, a0 + 32 + 8 * ks, brh, idesc, 1u);
    tc::mma_tf32_ts(d_v, a0 + 96 + 8 * ks, bih, idesc, 1u);
    tc::mma_tf32_ts(d_u, a0 + 8 * ks, brl, idesc, 1u);
    tc::mma_tf32_ts(d_v, a0 + 64 + 8 * ks, bil, idesc, 1u);
  }
}

__global__ void __launch_bounds__(BTC_THREADS, 1) hm_bank_tc_kernel(BankParams P) {
  extern __shared__ __align__(1024) char smem_tc[];
  float* fb = (float*)smem_tc;                               // [64][512]
  float* swin = fb + BTC_NSLOT * BTC_SLOT;                   // [512]  window at index n + 256
  float* bt = swin + BTC_SLOT;                               // [3 buffers][4 tiles Brh, Brl, Bih, Bil][1024]
  float* Cr = bt + BTC_NBUF * 4096;                          // [2 groups][4 frames][512]
  float* Ci = Cr + 2 * 4 * BTC_CST;                          // [2][4][512]
  BtcFrame* finfo = (BtcFrame*)(Ci + 2 * 4 * BTC_CST);       // [64]
  int* sb = (int*)(finfo + BTC_NSLOT);                       // [64] frame position
  int* sv = sb + BTC_NSLOT;                                  // [64] slot holds a voiced frame
  uint64_t* full = (uint64_t*)(sv + BTC_NSLOT);              // [3] operands of a chunk are in place
  uint64_t* empty = full + BTC_NBUF;                         // [3] the MMAs that read them have completed
  uint64_t* dfull = empty + BTC_NBUF;                               // [2] a group's accumulators are complete
  uint64_t* dempty = dfull + 2;                              // [2] ... and have been read back
  uint64_t* cfull = dempty + 2;                              // [2] a group's coefficients are staged
  uint32_t* tbase_s = (uint32_t*)(cfull + 2);
  int* tail_s = (int*)(tbase_s + 1);                         // first tile not yet written when the pipeline drains

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, seg = blockIdx.x;
  BTC_STAMP_DECL
  const int F = BTC_NSLOT - 2;
  const int N = P.n_hm, H = N >> 1;
  const int nf = P.nfrm_utt ? P.nfrm_utt[b] : P.nfrm;
  const int ny_b = P.ny_utt ? P.ny_utt[b] : P.ny;
  const int t0 = seg * F;
  const int start = t0 == 0 ? 0 : (t0 < nf ? P.hm_base[t0] : P.nsamp);
  const int end = (t0 + F < nf) ? P.hm_base[t0 + F] : P.nsamp;
  if(start >= end) return;                                   // uniform per CTA
  if(nf <= 0) {                                              // an empty utterance owns no tile: its row is silence
    float* yrow = P.y_sin + (size_t)b * P.stride;
    for(int i = tid; i < P.nsamp; i += BTC_THREADS) yrow[i] = 0.f;
    return;
  }
  const size_t row = (size_t)b * P.nfrm;

  // ---- set-up: tensor memory, barriers, window, per-frame scalars ----
  if(warp == 20) BTC_STAMP(2, 40);
  if(warp == 0) tc::tmem_alloc(tbase_s, 512);
  if(tid == 32) {
    for(int i = 0; i < BTC_NBUF; i ++) { tc::mbar_init(full + i, 12); tc::mbar_init(empty + i, 1); }
    for(int i = 0; i < 2; i ++) { tc::mbar_init(dfull + i, 1); tc::mbar_init(dempty + i, BTC_OUT_WARPS); tc::mbar_init(cfull + i, 8); }
    tc::fence_mbar_init();
  }
  for(int i = tid; i < BTC_SLOT; i += BTC_THREADS) {
    int j = i - 256 + H;
    swin[i] = (j >= 0 && j < N) ? P.win[j] : 0.f;
  }
  if(tid >= 64 && tid < 64 + BTC_NSLOT) {
    const int s = tid - 64, f = t0 - 1 + s;
    const bool inrange = f >= 0 && f < nf;
    float f0 = 0; int nh = 0;
    if(inrange) { f0 = P.f0[row + f]; nh = P.nhar[row + f]; }
    if(nh > 2048) nh = 2048;                                 // layer0.c:119,130
    bool voiced = inrange && f0 > 0 && nh > 0;               // layer0.c:125
    if(voiced && P.frame_mask) voiced = P.frame_mask[row + f] != 0;
    if(P.frame_hi > 0 && (f < P.frame_lo || f >= P.frame_hi)) voiced = false;
    sb[s] = inrange ? P.hm_base[f] : (f < 0 ? -(1 << 28) : (1 << 28));
    sv[s] = voiced ? 1 : 0;
    BtcFrame fi; fi.nufix = 0; fi.corr = 0; fi.nh = 0;
    if(voiced) {
      // per-frame scalars in the reference's precision (layer0.c:127-134, llsmutils.c:45-58, dsputils.c:338-348)
      const float f0n = f0 / P.fs;
      const float omega0 = (float)(2.0 * LLSM_PI * (double)f0n);
      bool iczt = false;
      if(P.has_options && P.use_iczt) {
        // log(N) a < log(nh) - b (llsmutils.c:51-53) <=> nh > exp(log(N) a + b), precomputed on the host; the two
        // logarithms are only evaluated when nh sits on the threshold
        iczt = (double)nh > P.iczt_nh;
        if(fabs((double)nh - P.iczt_nh) < 1e-6 * P.iczt_nh)
          iczt = log((double)N) * (double)P.iczt_a < log((double)nh) - (double)P.iczt_b;
      }
      if(iczt && nh > N - 1) nh = N - 1;
      fi.nufix = __double2ull_rn((iczt ? (double)omega0 / (2.0 * LLSM_PI) : (double)f0n) * 18446744073709551616.0);
      const float frac = P.hm_frac[f];
      fi.corr = (float)((double)(frac * 2.0f) * LLSM_PI / (double)P.fs * (double)f0);
      fi.nh = nh;
    }
    finfo[s] = fi;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tbase_s;
  if(warp == 20) BTC_STAMP(2, 41);
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);       // warp-uniform for the compiler

  if(uwarp == BTC_GEN_WARPS + BTC_OUT_WARPS) {
    // =========================== MMA issue: one warp, one elected lane ===========================
    // (the warp stays converged and the MMA operands are warp-uniform, so the elected lane's tcgen05.mma take
    //  them from uniform registers: one small MMA issues every 16 cycles, the tensor pipe's own rate)
    const uint32_t idesc = tc::idesc_tf32(128, 32, false);
    const uint64_t bbase = tc::smem_desc(tc::smem_u32(bt), 128, 1024);
    int bf = 0, gi = 0; uint32_t ph = 0;                      // ring position and its phase parity
    for(int grp = 0; grp < BTC_NSLOT / 4; grp ++) {
      int nhmax = 0;
#pragma unroll
      for(int g = 0; g < 4; g ++) nhmax = max(nhmax, finfo[4 * grp + g].nh);
      if(nhmax == 0) continue;
      const int nck = (nhmax + BTC_KC - 1) / BTC_KC;
      const uint32_t d_u = tbase + 128 * BTC_NBUF + 64 * (gi & 1), d_v = d_u + 32;
      tc::mbar_wait(dempty + (gi & 1), ((gi >> 1) & 1) ^ 1);     // the accumulators' previous contents have been read
      for(int c = 0; c < nck; c ++) {
        tc::mbar_wait(full + bf, ph);
        tc::fence_after_sync();
        BTC_STAMP(2, 30);
        if(tc::elect_one()) {
          if(bf == 0)      btc_issue_chunk<0>(tbase, bbase, idesc, d_u, d_v, c);
          else if(bf == 1) btc_issue_chunk<1>(tbase, bbase, idesc, d_u, d_v, c);
          else             btc_issue_chunk<2>(tbase, bbase, idesc, d_u, d_v, c);
          tc::mma_commit(empty + bf);
          if(c == nck - 1) tc::mma_commit(dfull + (gi & 1));
        }
        __syncwarp();
        BTC_STAMP(2, 31);
        if(++ bf == BTC_NBUF) { bf = 0; ph ^= 1; }
      }
      gi ++;
    }
  } else if(uwarp < 8) {
    // ========== A warps: rows of e^{i k theta} into tensor memory; coefficients of the next group ==========
    const int qd = warp & 3, hh = warp >> 2, pp = lane;          // frame (lane quadrant), half chunk, row
    const uint32_t tlane = tbase + ((uint32_t)(32 * qd) << 16) + 16 * hh;
    int bf = 0, gi = 0; uint32_t ph = 0;
    const int gC = tid >> 6, kC = tid & 63;                      // staging: frame, harmonic (mod 64)
    // a_k cos(phi'_k), a_k sin(phi'_k), phi'_k = phse[k] - corr (k + 1)  (layer0.c:132): group grp -> buffer cb.
    // The first 128 harmonics come from the registers loaded a group earlier.
    float pf_a[2], pf_p[2];
    auto prefetch = [&](int grp) {
#pragma unroll
      for(int it = 0; it < 2; it ++) {
        const int k = kC + 64 * it;
        pf_a[it] = 0.f; pf_p[it] = 0.f;
        if(k < finfo[4 * grp + gC].nh) { const size_t o = (row + t0 - 1 + 4 * grp + gC) * (size_t)P.maxnhar + k; pf_a[it] = P.ampl[o]; pf_p[it] = P.phse[o]; }
      }
    };
    auto stage = [&](int grp, int cb, int nhm) {
      const BtcFrame fi = finfo[4 * grp + gC];
      float* cr = Cr + (cb * 4 + gC) * BTC_CST; float* cim = Ci + (cb * 4 + gC) * BTC_CST;
      for(int k = kC, it = 0; k < nhm; k += 64, it ++) {
        float a = 0.f, phv = 0.f;
        if(it < 2) { a = it ? pf_a[1] : pf_a[0]; phv = it ? pf_p[1] : pf_p[0]; }
        else if(k < fi.nh) { const size_t o = (row + t0 - 1 + 4 * grp + gC) * (size_t)P.maxnhar + k; a = P.ampl[o]; phv = P.phse[o]; }
        const float ph2 = (float)((double)phv - (double)fi.corr * ((double)k + 1.0));
        float sn, cs; __sincosf(ph2, &sn, &cs);
        cr[k] = a * cs; cim[k] = a * sn;
      }
      __syncwarp();
      if(lane == 0) tc::mbar_arrive(cfull + cb);
    };
    auto next_group = [&](int from, int& nhm) {
      for(int g2 = from; g2 < BTC_NSLOT / 4; g2 ++) {
        int m = 0;
#pragma unroll
        for(int g = 0; g < 4; g ++) m = max(m, finfo[4 * g2 + g].nh);
        if(m > 0) { nhm = m; return g2; }
      }
      nhm = 0; return -1;
    };
    int nhmax = 0;
    int grp = next_group(0, nhmax);
    if(grp >= 0) { prefetch(grp); stage(grp, 0, ((nhmax + BTC_KC - 1) / BTC_KC) * BTC_KC); }
    while(grp >= 0) {
      const int s0 = 4 * grp;
      const int nck = (nhmax + BTC_KC - 1) / BTC_KC;
      int nh_next = 0;
      const int grp_next = next_group(grp + 1, nh_next);
      if(grp_next >= 0) prefetch(grp_next);                      // in flight while this group's chunks are generated
      float2 Wr, Wi, z2, z18;
      {
        const unsigned long long nfA = finfo[s0 + qd].nufix;
        const unsigned tA = 8u * (unsigned)pp, kka = 16u * (unsigned)hh + 1u;
        const float2 z1 = btc_phasor(nfA, tA);
        z2 = cmul(z1, z1); z18 = btc_phasor(nfA, 18u * tA);
        const float2 wa = btc_phasor(nfA, kka * tA), wb = cmul(wa, z1);
        Wr = make_float2(wa.x, wb.x); Wi = make_float2(wa.y, wb.y);
      }
      const float2 z2r = make_float2(z2.x, z2.x), z2i = make_float2(z2.y, z2.y), nz2i = make_float2(-z2.y, -z2.y);
      const float2 z18r = make_float2(z18.x, z18.x), z18i = make_float2(z18.y, z18.y), nz18i = make_float2(-z18.y, -z18.y);
      for(int c = 0; c < nck; c ++) {
        tc::mbar_wait(empty + bf, ph ^ 1);
        tc::fence_after_sync();
        const uint32_t ta = tlane + 128 * bf;
#pragma unroll
        for(int hf = 0; hf < 2; hf ++) {
          uint32_t arh[8], arl[8], aih[8], ail[8];
#pragma unroll
          for(int i = 0; i < 4; i ++) {
            btc_split2(Wr, arh[2 * i], arh[2 * i + 1], arl[2 * i], arl[2 * i + 1]);
            btc_split2(Wi, aih[2 * i], aih[2 * i + 1], ail[2 * i], ail[2 * i + 1]);
            if(hf == 1 && i == 3) btc_rot2(Wr, Wi, z18r, z18i, nz18i);     // to the next chunk's (k + 32, k + 33)
            else btc_rot2(Wr, Wi, z2r, z2i, nz2i);
          }
          tc::tmem_st8(ta + 8 * hf, arh); tc::tmem_st8(ta + 32 + 8 * hf, arl);
          tc::tmem_st8(ta + 64 + 8 * hf, aih); tc::tmem_st8(ta + 96 + 8 * hf, ail);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if(lane == 0) tc::mbar_arrive(full + bf);
        __syncwarp();
        if(++ bf == BTC_NBUF) { bf = 0; ph ^= 1; }
      }
      // the next group's coefficients go where the previous group's were: all MMAs of that group must be complete
      // (then the B warps have read its coefficients for the last time)
      gi ++;
      if(grp_next >= 0) {
        if(gi >= 2) tc::mbar_wait(dfull + (gi & 1), ((gi - 2) >> 1) & 1);
        stage(grp_next, gi & 1, ((nh_next + BTC_KC - 1) / BTC_KC) * BTC_KC);
      }
      grp = grp_next; nhmax = nh_next;
    }
  } else if(uwarp < BTC_GEN_WARPS) {
    // ====== B warps: columns a_k e^{i (k w q + phi_k)}; two teams of four warps take alternate chunks ======
    const int team = (warp - 8) >> 2, tT = tid - 256 - 128 * team, wB = warp - 8;
    const int qB = tT & 7, jB = (tT >> 3) & 7, g0 = tT >> 6;     // fine time index, harmonic quad, frames g0 and g0 + 2
    int bf = 0, gi = 0, ci = 0; uint32_t ph = 0;
    for(int grp = 0; grp < BTC_NSLOT / 4; grp ++) {
      const int s0 = 4 * grp;
      int nhmax = 0;
#pragma unroll
      for(int g = 0; g < 4; g ++) nhmax = max(nhmax, finfo[s0 + g].nh);
      if(nhmax == 0) continue;                                  // uniform: nothing voiced in the group
      const int nck = (nhmax + BTC_KC - 1) / BTC_KC;
      const int cb = gi & 1;
      const int c_first = (team ^ ci) & 1;                      // this team's first chunk of the group
      if(wB == 0) BTC_STAMP(1, 1);
      float2 w[2], rho[2], rho64[2], r2r[2], r2i[2];
#pragma unroll
      for(int m = 0; m < 2; m ++) {
        const unsigned long long nfB = finfo[s0 + g0 + 2 * m].nufix;
        rho[m] = btc_phasor(nfB, (unsigned)qB); rho64[m] = btc_phasor(nfB, 64u * (unsigned)qB);
        const float2 rho2 = cmul(rho[m], rho[m]);
        r2r[m] = make_float2(rho2.x, rho2.x); r2i[m] = make_float2(rho2.y, rho2.y);
        w[m] = btc_phasor(nfB, (unsigned)((32 * c_first + 4 * jB + 1) * qB));
      }
      tc::mbar_wait(cfull + cb, (gi >> 1) & 1);                  // the group's coefficients are staged
      if(wB == 0) BTC_STAMP(1, 3);
      for(int c = 0; c < nck; c ++, ci ++) {
        if(((ci ^ team) & 1) == 0) {
          if(wB == 0) BTC_STAMP(1, 4);
          tc::mbar_wait(empty + bf, ph ^ 1);
          if(wB == 0) BTC_STAMP(1, 10);
#pragma unroll
          for(int m = 0; m < 2; m ++) {
            const int gB = g0 + 2 * m;
            const float2 e1 = cmul(w[m], rho[m]);
            float2 Er = make_float2(w[m].x, e1.x), Ei = make_float2(w[m].y, e1.y);
            const float4 cr = *(const float4*)(Cr + (cb * 4 + gB) * BTC_CST + c * BTC_KC + 4 * jB);
            const float4 ci4 = *(const float4*)(Ci + (cb * 4 + gB) * BTC_CST + c * BTC_KC + 4 * jB);
            const float2 cr01 = make_float2(cr.x, cr.y), cr23 = make_float2(cr.z, cr.w);
            const float2 ci01 = make_float2(ci4.x, ci4.y), ci23 = make_float2(ci4.z, ci4.w);
            const float2 br01 = ffma2(Ei, make_float2(-ci4.x, -ci4.y), fmul2(Er, cr01)), bi01 = ffma2(Ei, cr01, fmul2(Er, ci01));
            btc_rot2(Er, Ei, r2r[m], r2i[m], make_float2(-r2i[m].x, -r2i[m].y));
            const float2 br23 = ffma2(Ei, make_float2(-ci4.z, -ci4.w), fmul2(Er, cr23)), bi23 = ffma2(Ei, cr23, fmul2(Er, ci23));
            uint4 rh, rl, ih, il;
            btc_split2(br01, rh.x, rh.y, rl.x, rl.y); btc_split2(br23, rh.z, rh.w, rl.z, rl.w);
            btc_split2(bi01, ih.x, ih.y, il.x, il.y); btc_split2(bi23, ih.z, ih.w, il.z, il.w);
            float* d = bt + bf * 4096 + jB * 32 + gB * 256 + qB * 4;   // floats: k-quad 128 B, frame 1 KB, q 16 B
            *(uint4*)(d) = rh; *(uint4*)(d + 1024) = rl; *(uint4*)(d + 2048) = ih; *(uint4*)(d + 3072) = il;
            w[m] = cmul(w[m], rho64[m]);
          }
          tc::fence_smem_to_async();
          __syncwarp();
          if(lane == 0) tc::mbar_arrive(full + bf);
          __syncwarp();
          if(wB == 0) BTC_STAMP(1, 13);
        }
        if(++ bf == BTC_NBUF) { bf = 0; ph ^= 1; }
      }
      gi ++;
    }
  } else {
    // ============ read-back warps: accumulators -> window -> parked frames -> overlap-add -> y_sin ============
    const int qd = warp & 3, pp = lane, wO = warp - BTC_GEN_WARPS;
    const uint32_t tlane = tbase + ((uint32_t)(32 * qd) << 16);
    float* yrow = P.y_sin + (size_t)b * P.stride;
    int gi = 0, next_tile = 1;
    auto emit_tiles = [&](int first, int last) {
      for(int s = first + ((wO - first) & 3); s <= last; s += BTC_OUT_WARPS)
        btc_emit_tile(s, lane, t0, nf, ny_b, N, H, P.nsamp, sb, sv, fb, yrow);
    };
    for(int grp = 0; grp < BTC_NSLOT / 4; grp ++) {
      int nhmax = 0;
#pragma unroll
      for(int g = 0; g < 4; g ++) nhmax = max(nhmax, finfo[4 * grp + g].nh);
      if(nhmax == 0) continue;
      const int db = gi & 1;
      if(wO == 0) BTC_STAMP(0, 20);
      tc::mbar_wait(dfull + db, (gi >> 1) & 1);
      tc::fence_after_sync();
      if(wO == 0) BTC_STAMP(0, 21);
      uint32_t u[8], v[8];
      const uint32_t td = tlane + 128 * BTC_NBUF + 64 * db + 8 * qd;
      tc::tmem_ld8(td, u); tc::tmem_ld8(td + 32, v);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      __syncwarp();
      if(lane == 0) tc::mbar_arrive(dempty + db);
      const int s = 4 * grp + qd;
      if(sv[s]) {
        // y[8 pp + q] = U - V, y[-8 pp + q] = U + V
        float* dp = fb + (size_t)s * BTC_SLOT + 256 + 8 * pp;
        const float* wp = swin + 256 + 8 * pp;
        const float4 w0 = *(const float4*)(wp), w1 = *(const float4*)(wp + 4);
        float o[8];
#pragma unroll
        for(int q = 0; q < 8; q ++) o[q] = __uint_as_float(u[q]) - __uint_as_float(v[q]);
        *(float4*)(dp) = make_float4(o[0] * w0.x, o[1] * w0.y, o[2] * w0.z, o[3] * w0.w);
        *(float4*)(dp + 4) = make_float4(o[4] * w1.x, o[5] * w1.y, o[6] * w1.z, o[7] * w1.w);
        if(pp > 0) {
          float* dn = fb + (size_t)s * BTC_SLOT + 256 - 8 * pp;
          const float* wn = swin + 256 - 8 * pp;
          const float4 x0 = *(const float4*)(wn), x1 = *(const float4*)(wn + 4);
#pragma unroll
          for(int q = 0; q < 8; q ++) o[q] = __uint_as_float(u[q]) + __uint_as_float(v[q]);
          *(float4*)(dn) = make_float4(o[0] * x0.x, o[1] * x0.y, o[2] * x0.z, o[3] * x0.w);
          *(float4*)(dn + 4) = make_float4(o[4] * x1.x, o[5] * x1.y, o[6] * x1.z, o[7] * x1.w);
        }
      }
      tc::named_bar_sync(2, 32 * BTC_OUT_WARPS);
      if(wO == 0) BTC_STAMP(0, 22);
      {   // the frames of this and of all earlier groups are parked and visible
        const int last = min(F, 4 * grp + 1);
        if(last >= next_tile) { emit_tiles(next_tile, last); next_tile = last + 1; }
      }
      if(wO == 0) BTC_STAMP(0, 23);
      gi ++;
    }
    if(tid == BTC_GEN_THREADS) *tail_s = next_tile;             // the tiles left over are shared by all warps below
  }
  if(warp == 20) BTC_STAMP(2, 42);
  tc::fence_before_sync();
  __syncthreads();
  if(warp == 20) BTC_STAMP(2, 43);
  if(warp == 0) tc::tmem_dealloc(tbase, 512);
  for(int s = *tail_s + warp; s <= F; s += BTC_THREADS / 32)
    btc_emit_tile(s, lane, t0, nf, ny_b, N, H, P.nsamp, sb, sv, fb, P.y_sin + (size_t)b * P.stride);
}

static inline size_t bank_tc_smem_bytes() {
  return (size_t)(BTC_NSLOT * BTC_SLOT + BTC_SLOT + BTC_NBUF * 4096 + 2 * 8 * BTC_CST) * 4 + BTC_NSLOT * (sizeof(BtcFrame) + 8) + 128;
}

// tensor-core path when the window fits the 32 x 8 sample grid; returns -1 otherwise
static inline int launch_hm_bank_tc(BankParams P, int nutt, int nfrm_max, cudaStream_t st) {
  P.iczt_nh = exp(log((double)P.n_hm) * (double)P.iczt_a + (double)P.iczt_b);
  if((P.n_hm >> 1) > BTC_MAXH || (P.n_hm & 1) || P.maxnhar > BTC_CST) return -1;
  const int F = BTC_NSLOT - 2;
  const int nseg = (std::max(nfrm_max, 1) + F - 1) / F;
  const size_t smem = bank_tc_smem_bytes();
  // (per launch: the attribute is per device, and a process may drive several)
  if(cudaFuncSetAttribute(hm_bank_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return -1;                     // not enough shared memory on this device: CUDA-core bank
  }
  hm_bank_tc_kernel<<<dim3(nseg, nutt), dim3(BTC_THREADS), smem, st>>>(P);
  return 0;
}
#endif  // LLSM_EMU
