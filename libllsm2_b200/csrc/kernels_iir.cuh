// Chunk-parallel zero-phase IIR filtering (filtfilt with zero initial state, order 4).
//
// Replaces the sequential recurrences of chebyfilt (dsputils.c:51-70) on the noise templates
// (dsputils.c:385-394) and on whole utterances (llsm_subband_energy, dsputils.c:230-235).
// One CTA per sequence; the sequence is cut into IIR_NT chunks of L samples (zero-padded to
// IIR_NT * L). For each direction of each section:
//   A. every thread filters its chunk from a zero state and keeps the final state f_c;
//   B. the incoming state of every chunk follows from the affine recurrence s_{c+1} = M s_c + f_c,
//      M = A^L (host-built, double), solved with a Kogge-Stone scan using M^(2^q);
//   C. every thread filters its chunk again from its true incoming state and writes in place.
// Mathematically identical to the sequential filter; arithmetic in double.
// (Tried and dropped: requesting the next four samples one iteration ahead -- 6.1 ms instead of 1.14 ms at C2;
//  replacing the second recurrence by y = y_zero_state + h_i . s_c with host-built impulse rows h_i = e0^T A^i in
//  shared memory, i.e. four independent FMAs per sample instead of nine dependent ones -- 1.87 ms instead of 1.17:
//  the extra store of the zero-state output and the eight-byte broadcast loads cost more than the recurrence.)
#pragma once
#include "common.cuh"

#define IIR_NT 128
#define IIR_NLOG 7        // log2(IIR_NT)

struct IirParams {
  int nchannel, n, L;
  float* y; int ystride;              // [nseq][ystride] in/out (section 0 may read elsewhere)
  const float* src_a; int sa_stride;  // optional stage-0 input (per utterance, sequence / nchannel)
  const float* src_b; int sb_stride;  // alternative stage-0 input
  unsigned src_b_mask;                // bit c: channel c reads src_b; both NULL -> read y
  int src_per_utt;                    // 1: src rows are per utterance (seq / nchannel), 0: per sequence
  const double* coef;                 // [nchannel][2][9]
  const double* mpow;                 // [nchannel][2][IIR_NLOG][16]
  int nstage[8];
  int square;                         // square the result (dsputils.c:232-233)
  int vec_ok;                         // y rows are 16-byte aligned (ystride % 4 == 0) and L % 4 == 0
};

struct IirCoef { double b0, b1, b2, b3, b4, a1, a2, a3, a4; };

__device__ __forceinline__ void iir_step(const IirCoef& cf, double xn, double& z0, double& z1,
  double& z2, double& z3, double& yn) {
  yn = cf.b0 * xn + z0;
  z0 = cf.b1 * xn + z1 - cf.a1 * yn;
  z1 = cf.b2 * xn + z2 - cf.a2 * yn;
  z2 = cf.b3 * xn + z3 - cf.a3 * yn;
  z3 = cf.b4 * xn - cf.a4 * yn;
}

// Memory access: thread t owns the samples [t L, (t + 1) L), so a warp reading "its" next sample touches 32 different
// cache lines (the first version did exactly that and was latency-bound: 32 sectors per request, 12 GB of DRAM
// traffic for 1.8 GB of data, 42 stall cycles per issued instruction on the scoreboard). The chunks therefore move
// through a shared-memory tile of IIR_NT rows x IIR_T samples: every row of the tile is loaded / stored by one warp
// as one 128-byte line, and a thread walks its own row (row stride IIR_T + 1: conflict-free).
#define IIR_T 32

__global__ void __launch_bounds__(IIR_NT) iir_filtfilt_kernel(IirParams P) {
  __shared__ double fs[IIR_NT][4];
  __shared__ double cfs[9];
  __shared__ double mp[IIR_NLOG][16];
  __shared__ float tile[IIR_NT][IIR_T + 1];
  const int seq = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = seq % P.nchannel;
  float* y = P.y + (size_t)seq * P.ystride;
  const int n = P.n, L = P.L;                      // L is a multiple of IIR_T
  const int nst = P.nstage[c];
  if(nst == 0) {
    for(int i = tid; i < n; i += blockDim.x) y[i] = 0.f;
    return;
  }
  const float* src0 = nullptr;
  if(P.src_a || P.src_b) {
    bool useb = (P.src_b_mask >> c) & 1u;
    size_t r = P.src_per_utt ? (size_t)(seq / P.nchannel) : (size_t)seq;
    src0 = useb ? P.src_b + r * P.sb_stride : P.src_a + r * P.sa_stride;
  }
  const int ntile = L / IIR_T;
  for(int st = 0; st < nst; st ++) {
    __syncthreads();
    if(tid < 9) cfs[tid] = P.coef[((size_t)c * 2 + st) * 9 + tid];
    for(int i = tid; i < IIR_NLOG * 16; i += blockDim.x)
      mp[i / 16][i % 16] = P.mpow[((size_t)c * 2 + st) * IIR_NLOG * 16 + i];
    __syncthreads();
    IirCoef cf;
    cf.b0 = cfs[0]; cf.b1 = cfs[1]; cf.b2 = cfs[2]; cf.b3 = cfs[3]; cf.b4 = cfs[4];
    cf.a1 = cfs[5]; cf.a2 = cfs[6]; cf.a3 = cfs[7]; cf.a4 = cfs[8];
    for(int dir = 0; dir < 2; dir ++) {
      const float* in = (st == 0 && dir == 0 && src0) ? src0 : y;
      // tile q (processing order) starts at sample offset `off` of every chunk
      auto load_tile = [&](int q) {
        const int off = dir == 0 ? q * IIR_T : L - IIR_T - q * IIR_T;
#pragma unroll 4
        for(int r = warp; r < IIR_NT; r += IIR_NT / 32) {
          const int idx = r * L + off + lane;
          tile[r][lane] = idx < n ? in[idx] : 0.f;
        }
      };
      // ---- A: zero-state pass over the chunk, keeps the final state
      double z0 = 0, z1 = 0, z2 = 0, z3 = 0, yn;
      for(int q = 0; q < ntile; q ++) {
        load_tile(q);
        __syncthreads();
#pragma unroll 8
        for(int e = 0; e < IIR_T; e ++) iir_step(cf, (double)tile[tid][dir == 0 ? e : IIR_T - 1 - e], z0, z1, z2, z3, yn);
        __syncthreads();
      }
      const int ord = dir == 0 ? tid : IIR_NT - 1 - tid;   // position in processing order
      fs[ord][0] = z0; fs[ord][1] = z1; fs[ord][2] = z2; fs[ord][3] = z3;
      __syncthreads();
      // ---- B: inclusive scan of s_{c+1} = M s_c + f_c
      for(int q = 0, o = 1; o < IIR_NT; q ++, o <<= 1) {
        double v0 = fs[ord][0], v1 = fs[ord][1], v2 = fs[ord][2], v3 = fs[ord][3];
        if(ord >= o) {
          const double* u = fs[ord - o];
          const double* M = mp[q];
          v0 += M[0] * u[0] + M[1] * u[1] + M[2] * u[2] + M[3] * u[3];
          v1 += M[4] * u[0] + M[5] * u[1] + M[6] * u[2] + M[7] * u[3];
          v2 += M[8] * u[0] + M[9] * u[1] + M[10] * u[2] + M[11] * u[3];
          v3 += M[12] * u[0] + M[13] * u[1] + M[14] * u[2] + M[15] * u[3];
        }
        __syncthreads();
        fs[ord][0] = v0; fs[ord][1] = v1; fs[ord][2] = v2; fs[ord][3] = v3;
        __syncthreads();
      }
      // ---- C: true pass, in place
      if(ord == 0) { z0 = z1 = z2 = z3 = 0; }
      else { z0 = fs[ord - 1][0]; z1 = fs[ord - 1][1]; z2 = fs[ord - 1][2]; z3 = fs[ord - 1][3]; }
      const bool last = P.square && st == nst - 1 && dir == 1;
      for(int q = 0; q < ntile; q ++) {
        load_tile(q);
        __syncthreads();
#pragma unroll 8
        for(int e = 0; e < IIR_T; e ++) {
          const int k = dir == 0 ? e : IIR_T - 1 - e;
          iir_step(cf, (double)tile[tid][k], z0, z1, z2, z3, yn);
          const float yf = (float)yn;
          tile[tid][k] = last ? yf * yf : yf;
        }
        __syncthreads();
        const int off = dir == 0 ? q * IIR_T : L - IIR_T - q * IIR_T;
#pragma unroll 4
        for(int r = warp; r < IIR_NT; r += IIR_NT / 32) {
          const int idx = r * L + off + lane;
          if(idx < n) y[idx] = tile[r][lane];
        }
        __syncthreads();
      }
    }
  }
}

// white-noise fill for the templates: copy the host-drawn template or draw on the device
struct WhiteParams { int nseq, nt, ostride; const float* white; unsigned long long seed; float* out;
  int seq_base; /* generator sequence number of row 0 (batches processed in slices draw the same noise) */ };

#ifndef LLSM_PHILOX_DEFINED
#define LLSM_PHILOX_DEFINED
// Philox4x32-10 counter-based generator (Salmon et al. 2011), used only when no host template is
// supplied (throughput mode; the reference draws from libc rand(), dsputils.c:353-361).
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
  unsigned k0, unsigned k1, unsigned* out) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for(int r = 0; r < 10; r ++) {
    unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif

// four N(0,1) samples per thread from one Philox block (two Box-Muller pairs)
__global__ void __launch_bounds__(256) white_fill_kernel(WhiteParams P) {
  const int seq = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;       // group of 4 samples
  const int n0 = q * 4;
  if(n0 >= P.nt) return;
  float* o = P.out + (size_t)seq * P.ostride;
  if(P.white) {
    const float* w = P.white + (size_t)seq * P.nt;
    for(int i = 0; i < 4 && n0 + i < P.nt; i ++) o[n0 + i] = w[n0 + i];
    return;
  }
  unsigned r[4];
  philox4x32_10((unsigned)q, (unsigned)(seq + P.seq_base), 0x6c6c736du, 0u, (unsigned)P.seed, (unsigned)(P.seed >> 32), r);
  float v[4];
  for(int h = 0; h < 2; h ++) {
    float u1 = ((float)r[2 * h] + 1.0f) * (1.0f / 4294967808.0f);
    float u2 = (float)r[2 * h + 1] * (1.0f / 4294967296.0f);
    float rad = sqrtf(-2.0f * logf(u1));
    float s, c; sincospif(2.0f * u2, &s, &c);
    v[2 * h] = rad * c; v[2 * h + 1] = rad * s;
  }
  for(int i = 0; i < 4 && n0 + i < P.nt; i ++) o[n0 + i] = v[i];
}
