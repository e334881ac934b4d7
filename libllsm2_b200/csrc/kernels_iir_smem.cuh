// Zero-phase order-4 IIR (filtfilt, zero initial state) with the WHOLE SEQUENCE RESIDENT IN SHARED MEMORY.
//
// Same operation as iir_filtfilt_kernel (kernels_iir.cuh: chebyfilt dsputils.c:51-70 on the noise templates and on
// whole utterances, llsm_subband_energy dsputils.c:230-235), same chunk-parallel mathematics, different data
// movement: that kernel streams every sequence through L2 / HBM twelve times (two sections x two directions x
// {zero-state pass, true pass + write}: 11 GB of DRAM traffic for 1.8 GB of data at BASELINE configs[1],
// profiles/r2c). Here a thread-block CLUSTER of CS CTAs (CS = 1, 2, 4, 8) holds one sequence in its distributed
// shared memory -- rank r keeps the samples [r 256 L, (r + 1) 256 L) as floats, brought in by bulk asynchronous
// copies (cp.async.bulk) -- all eight passes run on shared memory, and the sequence is written once.
//   * 256 chunks of L samples per CTA, L odd: thread t walks s[t L + e], a conflict-free stride.
//   * zero-state pass: only the chunk's final state is wanted, z_L = sum_e (A^(L-1-e) B) x[e]: four independent FMAs
//     per sample against host-built weights (broadcast from shared memory) instead of the nine-FMA dependent recurrence.
//   * incoming states: Kogge-Stone scan of s' = M s + f inside the CTA (M = A^L, host-built powers), the carry between
//     the CTAs of a cluster from their totals read through distributed shared memory (one cluster barrier per pass,
//     totals double-buffered), spread to a chunk by the binary decomposition of its index.
//   * true pass: the recurrence in place, in double, rounded to float where the streaming kernel rounds.
// (Tried and dropped: accumulating the next pass's chunk end states inside the true pass, whose samples it is producing
//  -- one loop less per pass, but 3.29 ms against 2.80 ms at C2: the extra FMAs and the conversion back to double sit in
//  the recurrence's issue stream, and the FP64 / conversion pipes, not the latency, are what the kernel runs into.)
// The g++ -DLLSM_EMU build (tests/emu) compiles the CS = 1 instance only (no clusters on the CPU emulation).
#pragma once
#include "kernels_iir.cuh"
#include "bulk_copy.cuh"
#include <cstdio>
#include <cstdlib>
#ifndef LLSM_EMU
#include <cooperative_groups.h>
#endif

#define IIS_NT 256
#define IIS_LOGNT 8         // log2(IIS_NT)
#define IIS_NLOG (IIS_LOGNT + 1)   // M^(L 2^q), q = 0 .. IIS_LOGNT (the last: a whole CTA)
#define IIS_LMAX 199        // 256 x 199 floats = 204 KB of samples per CTA

struct IirSmemParams {
  IirParams base;           // sequence geometry, sources, stages (coef: [nchannel][2][9])
  const double* mpow;       // [nchannel][2][IIS_NLOG][16]
  const double* wts;        // [nchannel][2][L][4]: A^j B, j = 0 .. L - 1
  int L;                    // chunk length (odd)
  int nseq;                 // sequences; the grid is persistent: cluster c serves sequences c, c + gridDim / CS, ...
};

static inline size_t iir_smem_bytes(int L) {
  return 64 + (size_t)IIS_NT * 4 * 8 + 2 * 4 * 8 + 16 * 8 + IIS_NLOG * 16 * 8 + (size_t)L * 4 * 8 + (size_t)IIS_NT * L * 4 + 16;
}

template <int CS>
__global__ void __launch_bounds__(IIS_NT, 2) iir_smem_kernel(IirSmemParams Q) {
  LLSM_DYN_SMEM(smem);
  const IirParams& P = Q.base;
  const int L = Q.L;
  bulk_bar_t* bar = (bulk_bar_t*)smem;                     // (64 bytes reserved)
  double* fs = (double*)(smem + 64);                       // [IIS_NT][4] chunk states
  double* tot = fs + IIS_NT * 4;                           // [2][4] this CTA's total, double-buffered across passes
  double* cfs = tot + 8;                                   // [16] (9 used)
  double* mp = cfs + 16;                                   // [IIS_NLOG][16]
  double* wt = mp + IIS_NLOG * 16;                         // [L][4]
  float* s = (float*)(wt + (size_t)L * 4);                 // [IIS_NT * L] this CTA's samples
  const int tid = threadIdx.x;
  const int rank = blockIdx.x % CS;
  const int n = P.n;
  const int part = IIS_NT * L;                             // samples per CTA
  const int g0 = rank * part;                              // first sample of this CTA
  const int mine = max(0, min(n - g0, part));              // real samples held here
#ifndef LLSM_EMU
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
#endif
  if(tid == 0) bulk_bar_init(bar);
  __syncthreads();
  int phase = 0, round = 0;
  // persistent: the grid is sized to the machine, so every CTA is dispatched at once and a kernel launched behind this
  // one on another stream (the analysis runs the HBM-bound Kalman smoother there) shares the SMs with it
  for(int seq = blockIdx.x / CS; seq < Q.nseq; seq += gridDim.x / CS) {
  const int c = seq % P.nchannel;
  float* y = P.y + (size_t)seq * P.ystride;
  const int nst = P.nstage[c];
  if(nst == 0) {                                           // channel absent: silence (uniform over the cluster)
    for(int i = tid; i < mine; i += IIS_NT) y[g0 + i] = 0.f;
    continue;
  }
  const float* src = y;
  if(P.src_a || P.src_b) {
    const bool useb = (P.src_b_mask >> c) & 1u;
    const size_t r = P.src_per_utt ? (size_t)(seq / P.nchannel) : (size_t)seq;
    src = useb ? P.src_b + r * P.sb_stride : P.src_a + r * P.sa_stride;
  }
  // ---- the sequence part: bulk copies when the row is 16-byte aligned, plain loads otherwise; zero padding behind it
  const bool aligned = (((uintptr_t)(src + g0)) & 15) == 0;
  const int nbulk = aligned ? (mine & ~3) : 0;
#ifndef LLSM_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy accesses above, async-proxy writes below
#endif
  __syncthreads();                                          // the previous sequence has left shared memory
  if(tid == 0) {
    bulk_expect(bar, (uint32_t)(nbulk * 4));
    for(int o = 0; o < nbulk; o += 8192)
      bulk_g2s(s + o, src + g0 + o, (uint32_t)(min(8192, nbulk - o) * 4), bar);
  }
  for(int i = nbulk + tid; i < part; i += IIS_NT) s[i] = i < mine ? src[g0 + i] : 0.f;
  bulk_wait(bar, (uint32_t)(round & 1));
  round ++;
  __syncthreads();

  for(int st = 0; st < nst; st ++) {
    __syncthreads();
    if(tid < 9) cfs[tid] = P.coef[((size_t)c * 2 + st) * 9 + tid];
    for(int i = tid; i < IIS_NLOG * 16; i += IIS_NT) mp[i] = Q.mpow[((size_t)c * 2 + st) * IIS_NLOG * 16 + i];
    for(int i = tid; i < L * 4; i += IIS_NT) wt[i] = Q.wts[((size_t)c * 2 + st) * L * 4 + i];
    __syncthreads();
    IirCoef cf;
    cf.b0 = cfs[0]; cf.b1 = cfs[1]; cf.b2 = cfs[2]; cf.b3 = cfs[3]; cf.b4 = cfs[4];
    cf.a1 = cfs[5]; cf.a2 = cfs[6]; cf.a3 = cfs[7]; cf.a4 = cfs[8];
    for(int dir = 0; dir < 2; dir ++, phase ++) {
      float* mych = s + (size_t)tid * L;
      const int ord = dir == 0 ? tid : IIS_NT - 1 - tid;     // position of the chunk in processing order, in the CTA
      const int ordc = dir == 0 ? rank : CS - 1 - rank;      // position of the CTA in processing order
      // ---- A: final state of the chunk from a zero state: z = sum_e w[L - 1 - e] x[e] (e in processing order)
      double z0 = 0, z1 = 0, z2 = 0, z3 = 0;
      {
        double u0 = 0, u1 = 0, u2 = 0, u3 = 0;               // (two accumulator sets: shorter dependency chains)
        int e = 0;
        for(; e + 1 < L; e += 2) {
          const double xa = (double)mych[dir == 0 ? e : L - 1 - e], xb = (double)mych[dir == 0 ? e + 1 : L - 2 - e];
          const double* wa = wt + (size_t)(L - 1 - e) * 4; const double* wb = wa - 4;
          z0 = fma(wa[0], xa, z0); z1 = fma(wa[1], xa, z1); z2 = fma(wa[2], xa, z2); z3 = fma(wa[3], xa, z3);
          u0 = fma(wb[0], xb, u0); u1 = fma(wb[1], xb, u1); u2 = fma(wb[2], xb, u2); u3 = fma(wb[3], xb, u3);
        }
        if(e < L) {
          const double xa = (double)mych[dir == 0 ? e : L - 1 - e];
          const double* wa = wt + (size_t)(L - 1 - e) * 4;
          z0 = fma(wa[0], xa, z0); z1 = fma(wa[1], xa, z1); z2 = fma(wa[2], xa, z2); z3 = fma(wa[3], xa, z3);
        }
        z0 += u0; z1 += u1; z2 += u2; z3 += u3;
      }
      fs[ord * 4 + 0] = z0; fs[ord * 4 + 1] = z1; fs[ord * 4 + 2] = z2; fs[ord * 4 + 3] = z3;
      __syncthreads();
      // ---- B: inclusive scan of s_{c+1} = M s_c + f_c over the CTA's chunks
      for(int q = 0, o = 1; o < IIS_NT; q ++, o <<= 1) {
        double v0 = fs[ord * 4], v1 = fs[ord * 4 + 1], v2 = fs[ord * 4 + 2], v3 = fs[ord * 4 + 3];
        if(ord >= o) {
          const double* u = fs + (size_t)(ord - o) * 4;
          const double* M = mp + q * 16;
          v0 += M[0] * u[0] + M[1] * u[1] + M[2] * u[2] + M[3] * u[3];
          v1 += M[4] * u[0] + M[5] * u[1] + M[6] * u[2] + M[7] * u[3];
          v2 += M[8] * u[0] + M[9] * u[1] + M[10] * u[2] + M[11] * u[3];
          v3 += M[12] * u[0] + M[13] * u[1] + M[14] * u[2] + M[15] * u[3];
        }
        __syncthreads();
        fs[ord * 4] = v0; fs[ord * 4 + 1] = v1; fs[ord * 4 + 2] = v2; fs[ord * 4 + 3] = v3;
        __syncthreads();
      }
      // ---- carry into this CTA from the CTAs before it in processing order
      double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#ifndef LLSM_EMU
      if(CS > 1) {
        double* tp = tot + (phase & 1) * 4;
        if(tid < 4) tp[tid] = fs[(IIS_NT - 1) * 4 + tid];
        cluster.sync();
        const double* M = mp + IIS_LOGNT * 16;               // M^(IIS_NT L): a whole CTA
        for(int j = 0; j < ordc; j ++) {
          const int rj = dir == 0 ? j : CS - 1 - j;
          const double* tr = cluster.map_shared_rank(tp, rj);
          const double t0 = tr[0], t1 = tr[1], t2 = tr[2], t3 = tr[3];
          const double n0 = M[0] * c0 + M[1] * c1 + M[2] * c2 + M[3] * c3 + t0;
          const double n1 = M[4] * c0 + M[5] * c1 + M[6] * c2 + M[7] * c3 + t1;
          const double n2 = M[8] * c0 + M[9] * c1 + M[10] * c2 + M[11] * c3 + t2;
          const double n3 = M[12] * c0 + M[13] * c1 + M[14] * c2 + M[15] * c3 + t3;
          c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        }
      }
#endif
      // incoming state of this chunk: the scan's value before it plus M^(ord L) carry
      if(ord == 0) { z0 = 0; z1 = 0; z2 = 0; z3 = 0; }
      else { z0 = fs[(ord - 1) * 4]; z1 = fs[(ord - 1) * 4 + 1]; z2 = fs[(ord - 1) * 4 + 2]; z3 = fs[(ord - 1) * 4 + 3]; }
      if(CS > 1 && ordc > 0) {
        for(int q = 0; q < IIS_LOGNT; q ++) {
          if((ord >> q) & 1) {
            const double* M = mp + q * 16;
            const double n0 = M[0] * c0 + M[1] * c1 + M[2] * c2 + M[3] * c3;
            const double n1 = M[4] * c0 + M[5] * c1 + M[6] * c2 + M[7] * c3;
            const double n2 = M[8] * c0 + M[9] * c1 + M[10] * c2 + M[11] * c3;
            const double n3 = M[12] * c0 + M[13] * c1 + M[14] * c2 + M[15] * c3;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
          }
        }
        z0 += c0; z1 += c1; z2 += c2; z3 += c3;
      }
      // ---- C: true pass, in place
      // (samples behind the end of the sequence stay zero: a pass must not see the previous pass ring on into the padding
      //  -- filtfilt filters exactly n samples each way)
      const bool last = P.square && st == nst - 1 && dir == 1;
      const int kend = mine - tid * L;                        // chunk positions k >= kend lie behind the end
      double yn;
#pragma unroll 4
      for(int e = 0; e < L; e ++) {
        const int k = dir == 0 ? e : L - 1 - e;
        iir_step(cf, (double)mych[k], z0, z1, z2, z3, yn);
        const float yf = (float)yn;
        mych[k] = k < kend ? (last ? yf * yf : yf) : 0.f;
      }
      __syncthreads();                                       // (fs is rewritten by the next pass)
    }
  }
  // ---- the sequence goes back to global memory once
  if((((uintptr_t)(y + g0)) & 15) == 0) {
    const int n4 = mine >> 2;
    for(int i = tid; i < n4; i += IIS_NT) ((float4*)(y + g0))[i] = ((const float4*)s)[i];
    for(int i = (n4 << 2) + tid; i < mine; i += IIS_NT) y[g0 + i] = s[i];
  } else {
    for(int i = tid; i < mine; i += IIS_NT) y[g0 + i] = s[i];
  }
  }
#ifndef LLSM_EMU
  if(CS > 1) cluster.sync();                                 // no CTA leaves while its totals may still be read
#endif
}

// chunk length and cluster size for sequences of n samples; cs = 0: too long for eight CTAs (streaming kernel)
static inline void iir_smem_geometry(int n, int& cs, int& L) {
  cs = 0; L = 0;
  static int lmax = -1;
  if(lmax < 0) { const char* e = getenv("LLSM_IIR_LMAX"); lmax = e ? atoi(e) : IIS_LMAX; if(lmax < 1 || lmax > IIS_LMAX) lmax = IIS_LMAX; }
  for(int c2 = 1; c2 <= 8; c2 <<= 1) {
    int l = (n + c2 * IIS_NT - 1) / (c2 * IIS_NT);
    if(l < 1) l = 1;
    l |= 1;                                                  // odd: conflict-free chunk stride
    if(l <= lmax || (c2 == 8 && l <= IIS_LMAX)) { cs = c2; L = l; return; }
  }
}

// host tables of one section: coef[9], mpow[IIS_NLOG][16] (via build_iir_section) and wts[L][4] = A^j B
static inline void build_iir_smem_section(const double b[5], const double a[5], int L, double* coef, double* mpow, double* wts) {
  build_iir_section(b, a, L, IIS_NLOG, coef, mpow);
  // state update of iir_step: z' = A z + B x with y = b0 x + z0
  const double b0 = coef[0];
  long double A[16] = {0}, Bv[4];
  A[0] = -coef[5]; A[1] = 1; A[4] = -coef[6]; A[6] = 1; A[8] = -coef[7]; A[11] = 1; A[12] = -coef[8];
  for(int i = 0; i < 4; i ++) Bv[i] = (long double)coef[1 + i] - (long double)coef[5 + i] * b0;
  long double v[4] = {Bv[0], Bv[1], Bv[2], Bv[3]};
  for(int j = 0; j < L; j ++) {
    for(int i = 0; i < 4; i ++) wts[(size_t)j * 4 + i] = (double)v[i];
    long double nv[4];
    for(int i = 0; i < 4; i ++) nv[i] = A[i * 4] * v[0] + A[i * 4 + 1] * v[1] + A[i * 4 + 2] * v[2] + A[i * 4 + 3] * v[3];
    for(int i = 0; i < 4; i ++) v[i] = nv[i];
  }
}

// launch on nseq sequences; returns -1 when the configuration does not fit (caller falls back to the streaming kernel)
static inline int launch_iir_smem(const IirSmemParams& Q, int nseq, int cs, cudaStream_t st, bool persistent = false) {
  const size_t smem = iir_smem_bytes(Q.L);
  if(getenv("LLSM_IIR_TRACE")) fprintf(stderr, "iir_smem: nseq %d n %d cluster %d L %d smem %zu\n", nseq, Q.base.n, cs, Q.L, smem);
#ifdef LLSM_EMU
  if(cs != 1) return -1;
  IirSmemParams Q2 = Q; Q2.nseq = nseq;
  LLSM_LAUNCH(iir_smem_kernel<1>, dim3(std::min(nseq, 3)), dim3(IIS_NT), smem, st, Q2);
  return 0;
#else
  if(smem > 227 * 1024) return -1;
  IirSmemParams Q2 = Q; Q2.nseq = nseq;
  int dev = 0, sms = 148;
  if(cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  // persistent (a machine-sized grid walking the sequences) only where another stream is to share the SMs: with a
  // CTA per sequence the hardware overlaps one CTA's load and store with its neighbour's arithmetic (templates: 0.58 ms
  // against 0.77 ms persistent)
  int nclu = persistent ? sms * per_sm / cs : nseq;
  if(nclu > nseq) nclu = nseq;
  if(nclu < 1) nclu = 1;
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(nclu * cs)); cfg.blockDim = dim3(IIS_NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaErrorInvalidValue;
#define LLSM_IIS_GO(CSV) { auto kfn = iir_smem_kernel<CSV>; \
    if(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1; } \
    e = cudaLaunchKernelEx(&cfg, kfn, Q2); }
  if(cs == 1) LLSM_IIS_GO(1) else if(cs == 2) LLSM_IIS_GO(2) else if(cs == 4) LLSM_IIS_GO(4) else if(cs == 8) LLSM_IIS_GO(8)
#undef LLSM_IIS_GO
  if(e != cudaSuccess) { cudaGetLastError(); return -1; }
  return 0;
#endif
}
