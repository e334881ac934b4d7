// Shared device helpers for the libllsm2 B200 kernels (sm_100a).
//
// The same sources compile in two modes:
//   * nvcc, -gencode arch=compute_100a,code=sm_100a : the product;
//   * g++ -DLLSM_EMU (tests/emu/) : CPU thread emulation used only by the CPU-side unit tests.
#pragma once

#ifdef LLSM_EMU
#include "cuda_emu.h"
typedef void* cudaStream_t;
#define LLSM_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define LLSM_DYN_SMEM(name) extern __shared__ __align__(16) char name[]
#define LLSM_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#define LLSM_PI 3.14159265358979323846
#define LLSM_WARP 32

// ---- packed FP32x2 arithmetic (Blackwell FFMA2 / FMUL2: two FP32 lanes per issue slot) ----
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
#else
  return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
#else
  return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#endif
}

// e^{i 2 pi u}: u (turns) is reduced in double, the sine/cosine are evaluated in float.
// Phases on this path reach ~1e3 rad (harmonic number x sample offset), where a float argument
// would lose 1e-4 rad; reducing the product in double first keeps the seed error at ~1e-7 rad.
__device__ __forceinline__ float2 unit_phasor_turns(double u) {
  u -= rint(u);                 // [-0.5, 0.5]
  float s, c;
  sincospif(2.0f * (float)u, &s, &c);
  return make_float2(c, s);
}

// e^{i 2 pi u} for a small turn count (|u| of a few turns at most, e.g. an envelope harmonic over one window):
// the product is formed in float, reduced with one rounding and evaluated by the special-function unit
// (absolute error ~5e-7: the envelopes it feeds are compared at 1e-4).
__device__ __forceinline__ float2 unit_phasor_small(float u) {
#ifdef LLSM_EMU
  u -= rintf(u);
  return make_float2(cosf(6.2831853071795865f * u), sinf(6.2831853071795865f * u));
#else
  u -= rintf(u);
  const float a = 6.2831853071795865f * u;
  return make_float2(__cosf(a), __sinf(a));
#endif
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// ---- block-wide Stockham FFT in shared memory -----------------------------------------------
// n = 2^lg complex points in `a` (second buffer `b`, same size); all threads of the block take
// part; radix-4 passes plus one radix-2 pass when lg is odd. `tw` is the full-circle table
// tw[m] = exp(-2 pi i m / ntw) (ntw >= n, power of two, built on the host in double).
// Returns the buffer holding the result. Forward: X[k] = sum x[m] e^{-2 pi i k m / n};
// inverse (INV): conjugate twiddles, unscaled. Every pass ends with __syncthreads().
template <bool INV>
__device__ __forceinline__ float2 fft_tw(const float2* __restrict__ tw, int idx) {
  float2 w = tw[idx];
  if(INV) w.y = -w.y;
  return w;
}

// Bank swizzle for the FFT buffers (SWZ): element j lives at j ^ (((j >> 4) & 3) * 5). The radix-4 passes
// write with strides 4 p (p = 1, 4): without the swizzle 8 lanes of a warp hit the same bank pair; with it
// every pass touches 16 distinct bank pairs per half-warp. It permutes the low 4 index bits inside each
// 64-element block, so buffers keep their size. Callers index the buffers through fsw() as well.
__device__ __forceinline__ int fsw(int j) { return j ^ (((j >> 4) & 3) * 5); }

template <bool INV, bool SWZ = false>
__device__ float2* block_fft(float2* a, float2* b, int lg, const float2* __restrict__ tw, int ntw) {
#define FIX(j) (SWZ ? fsw(j) : (j))
  const int n = 1 << lg;
  const int tid = threadIdx.x, nth = blockDim.x;
  int p = 1; // butterflies already combined
  int rem = lg;
  int sh = 0; while((4 << sh) < ntw) sh ++;    // twiddle index step of a radix-4 pass: ntw / (4 p) = 1 << sh
  while(rem >= 2) {
    const int t = n >> 2;
    for(int i = tid; i < t; i += nth) {
      int k = i & (p - 1);
      float2 u0 = a[FIX(i)], u1 = a[FIX(i + t)], u2 = a[FIX(i + 2 * t)], u3 = a[FIX(i + 3 * t)];
      if(p > 1) {                               // the first pass has unit twiddles (uniform branch)
        const int m = k << sh;                  // angle -2 pi k / (4p)
        u1 = cmul(u1, fft_tw<INV>(tw, m));
        u2 = cmul(u2, fft_tw<INV>(tw, 2 * m));
        u3 = cmul(u3, fft_tw<INV>(tw, 3 * m));
      }
      float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y);
      float2 v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
      float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
      float2 v3 = make_float2(u1.x - u3.x, u1.y - u3.y);
      // forward: multiply v3 by -i; inverse: by +i
      float2 v3r = INV ? make_float2(-v3.y, v3.x) : make_float2(v3.y, -v3.x);
      int j = ((i - k) << 2) + k;
      b[FIX(j)]         = make_float2(v0.x + v2.x, v0.y + v2.y);
      b[FIX(j + p)]     = make_float2(v1.x + v3r.x, v1.y + v3r.y);
      b[FIX(j + 2 * p)] = make_float2(v0.x - v2.x, v0.y - v2.y);
      b[FIX(j + 3 * p)] = make_float2(v1.x - v3r.x, v1.y - v3r.y);
    }
    __syncthreads();
    float2* sw = a; a = b; b = sw;
    p <<= 2; rem -= 2; sh -= 2;
  }
  if(rem == 1) {
    const int t = n >> 1;
    for(int i = tid; i < t; i += nth) {
      int k = i & (p - 1);
      float2 u0 = a[FIX(i)];
      float2 u1 = a[FIX(i + t)];
      if(p > 1) u1 = cmul(u1, fft_tw<INV>(tw, k << (sh + 1)));   // ntw / (2 p)
      int j = ((i - k) << 1) + k;
      b[FIX(j)]     = make_float2(u0.x + u1.x, u0.y + u1.y);
      b[FIX(j + p)] = make_float2(u0.x - u1.x, u0.y - u1.y);
    }
    __syncthreads();
    float2* sw = a; a = b; b = sw;
  }
  return a;
#undef FIX
}

// ---- block-wide Stockham FFT, radix 8 (register butterflies) ---------------------------------------
// Same contract as block_fft (n = 2^lg points in `a`, ping-pong buffer `b`, full-circle table tw[m] =
// exp(-2 pi i m / ntw), returns the buffer holding the result, every pass ends with __syncthreads()), but a thread
// carries a whole radix-8 butterfly in registers: lg / 3 passes (+ one radix-4 or radix-2 pass) instead of lg / 2,
// i.e. 4 shared-memory round trips for 2048 points instead of 6, and a third fewer twiddle products.
// Buffers are indexed through fpad(): one float2 of padding after every 16 makes the stride-8 scatter of the
// first pass conflict-free (16 lanes -> 16 distinct bank pairs); a buffer holds fpad(n) = n + n / 16 float2.
__device__ __forceinline__ int fpad(int j) { return j + (j >> 4); }

template <bool INV>
__device__ __forceinline__ float2 mul_w8(float2 v, int m) {      // v * W8^m, W8 = exp(-+ 2 pi i / 8), m = 1, 2, 3
  const float h = 0.70710678118654752f;
  if(m == 2) return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
  if(m == 1) return INV ? make_float2((v.x - v.y) * h, (v.x + v.y) * h) : make_float2((v.x + v.y) * h, (v.y - v.x) * h);
  return INV ? make_float2(-(v.x + v.y) * h, (v.x - v.y) * h) : make_float2((v.y - v.x) * h, -(v.x + v.y) * h);
}

// 8-point DFT in registers: u[j] <- sum_q u[q] W8^{jq} (decimation in frequency, outputs written to natural order)
template <bool INV>
__device__ __forceinline__ void dft8_regs(float2 (&u)[8]) {
  float2 a[8];
#pragma unroll
  for(int q = 0; q < 4; q ++) {
    a[q] = make_float2(u[q].x + u[q + 4].x, u[q].y + u[q + 4].y);
    a[q + 4] = make_float2(u[q].x - u[q + 4].x, u[q].y - u[q + 4].y);
  }
  a[5] = mul_w8<INV>(a[5], 1); a[6] = mul_w8<INV>(a[6], 2); a[7] = mul_w8<INV>(a[7], 3);
#pragma unroll
  for(int h = 0; h < 2; h ++) {                          // 4-point DFT of a[4h .. 4h + 3] -> X[h + 2m]
    const float2 c0 = a[4 * h], c1 = a[4 * h + 1], c2 = a[4 * h + 2], c3 = a[4 * h + 3];
    const float2 e0 = make_float2(c0.x + c2.x, c0.y + c2.y), e1 = make_float2(c0.x - c2.x, c0.y - c2.y);
    const float2 o0 = make_float2(c1.x + c3.x, c1.y + c3.y);
    const float2 o1 = mul_w8<INV>(make_float2(c1.x - c3.x, c1.y - c3.y), 2);
    u[h]     = make_float2(e0.x + o0.x, e0.y + o0.y);
    u[h + 4] = make_float2(e0.x - o0.x, e0.y - o0.y);
    u[h + 2] = make_float2(e1.x + o1.x, e1.y + o1.y);
    u[h + 6] = make_float2(e1.x - o1.x, e1.y - o1.y);
  }
}

// One Stockham pass of radix R (8, 4 or 2) from buffer a to buffer b; p = product of the radices already applied.
template <int R, bool INV>
__device__ __forceinline__ void fft_pass(const float2* a, float2* b, int n, int p, int lg_p,
  const float2* __restrict__ tw, int lg_ntw) {
  const int t = n / R;
  constexpr int LR = R == 8 ? 3 : (R == 4 ? 2 : 1);
  const int sh = lg_ntw - lg_p - LR;                     // twiddle index step: ntw / (p R)
  for(int i = threadIdx.x; i < t; i += blockDim.x) {
    const int k = i & (p - 1);
    float2 u[R];
#pragma unroll
    for(int q = 0; q < R; q ++) u[q] = a[fpad(i + q * t)];
    if(p > 1) {                                          // the first pass has unit twiddles (uniform branch)
      const int m = k << sh;
#pragma unroll
      for(int q = 1; q < R; q ++) u[q] = cmul(u[q], fft_tw<INV>(tw, q * m));
    }
    const int j = ((i - k) << LR) + k;
    if(R == 8) {
      float2 v[8];
#pragma unroll
      for(int q = 0; q < 8; q ++) v[q] = u[q < R ? q : 0];
      dft8_regs<INV>(v);
#pragma unroll
      for(int q = 0; q < 8; q ++) b[fpad(j + q * p)] = v[q];
    } else if(R == 4) {
      const float2 v0 = make_float2(u[0].x + u[2 % R].x, u[0].y + u[2 % R].y), v1 = make_float2(u[0].x - u[2 % R].x, u[0].y - u[2 % R].y);
      const float2 v2 = make_float2(u[1].x + u[3 % R].x, u[1].y + u[3 % R].y);
      const float2 v3 = mul_w8<INV>(make_float2(u[1].x - u[3 % R].x, u[1].y - u[3 % R].y), 2);
      b[fpad(j)]         = make_float2(v0.x + v2.x, v0.y + v2.y);
      b[fpad(j + p)]     = make_float2(v1.x + v3.x, v1.y + v3.y);
      b[fpad(j + 2 * p)] = make_float2(v0.x - v2.x, v0.y - v2.y);
      b[fpad(j + 3 * p)] = make_float2(v1.x - v3.x, v1.y - v3.y);
    } else {
      b[fpad(j)]     = make_float2(u[0].x + u[1].x, u[0].y + u[1].y);
      b[fpad(j + p)] = make_float2(u[0].x - u[1].x, u[0].y - u[1].y);
    }
  }
  __syncthreads();
}

template <bool INV>
__device__ float2* block_fft8(float2* a, float2* b, int lg, const float2* __restrict__ tw, int lg_ntw) {
  const int n = 1 << lg;
  int lg_p = 0;
  while(lg - lg_p >= 3) {
    fft_pass<8, INV>(a, b, n, 1 << lg_p, lg_p, tw, lg_ntw);
    float2* s = a; a = b; b = s; lg_p += 3;
  }
  if(lg - lg_p == 2) {
    fft_pass<4, INV>(a, b, n, 1 << lg_p, lg_p, tw, lg_ntw);
    float2* s = a; a = b; b = s;
  } else if(lg - lg_p == 1) {
    fft_pass<2, INV>(a, b, n, 1 << lg_p, lg_p, tw, lg_ntw);
    float2* s = a; a = b; b = s;
  }
  return a;
}

// request a line into L2 ahead of its use (no register, no stall)
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef LLSM_EMU
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#else
  (void)p;
#endif
}

// ---- warp-level tensor-core MMA, TF32 operands, FP32 accumulate -------------------------------------
// D (16 x 8) += A (16 x 8, row) * B (8 x 8, col). Fragment ownership (g = lane / 4, t = lane % 4):
//   a0 = A[g][t], a1 = A[g + 8][t], a2 = A[g][t + 4], a3 = A[g + 8][t + 4]
//   b0 = B[t][g], b1 = B[t + 4][g]
//   d0 = D[g][2t], d1 = D[g][2t + 1], d2 = D[g + 8][2t], d3 = D[g + 8][2t + 1]
// Operands are FP32 bit patterns whose low 13 mantissa bits the tensor core ignores.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
#ifdef LLSM_EMU
  static float sa[64][16][8], sb[64][8][8];           // per warp of the (single) running CTA
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 63, g = lane >> 2, t = lane & 3;
  auto tf = [](float v) { unsigned u; memcpy(&u, &v, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; };
  __syncwarp();
  sa[w][g][t] = tf(a[0]); sa[w][g + 8][t] = tf(a[1]); sa[w][g][t + 4] = tf(a[2]); sa[w][g + 8][t + 4] = tf(a[3]);
  sb[w][t][g] = tf(b[0]); sb[w][t + 4][g] = tf(b[1]);
  __syncwarp();
  for(int k = 0; k < 8; k ++) {
    d[0] += sa[w][g][k] * sb[w][k][2 * t];     d[1] += sa[w][g][k] * sb[w][k][2 * t + 1];
    d[2] += sa[w][g + 8][k] * sb[w][k][2 * t]; d[3] += sa[w][g + 8][k] * sb[w][k][2 * t + 1];
  }
#else
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
    : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
    : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
      "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
#endif
}
// x = hi + lo with hi exactly representable in TF32 (truncation) and lo the exact remainder
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}

// ---- warp-level tensor-core MMA, FP16 operands, FP32 accumulate ---------------------------------------
// D (16 x 8) += A (16 x 16, row) * B (16 x 8, col). Fragment ownership (g = lane / 4, t = lane % 4), two halves per
// register, the lower column / row index in the low 16 bits:
//   a0 = A[g][2t, 2t+1], a1 = A[g+8][2t, 2t+1], a2 = A[g][2t+8, 2t+9], a3 = A[g+8][2t+8, 2t+9]
//   b0 = B[2t, 2t+1][g], b1 = B[2t+8, 2t+9][g];   d as for the TF32 form.
// Twice the multiply-adds of the TF32 instruction per issue slot of the tensor pipe.
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
#ifdef LLSM_EMU
  _Float16 a = (_Float16)lo_elem, b = (_Float16)hi_elem;
  uint16_t ua, ub; memcpy(&ua, &a, 2); memcpy(&ub, &b, 2);
  return (uint32_t)ua | ((uint32_t)ub << 16);
#else
  uint32_t d; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem)); return d;
#endif
}
__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
#ifdef LLSM_EMU
  static float sa[64][16][16], sb[64][16][8];           // per warp of the (single) running CTA
  const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & 63, g = lane >> 2, t = lane & 3;
  auto lo = [](uint32_t v) { uint16_t u = (uint16_t)(v & 0xffffu); _Float16 h; memcpy(&h, &u, 2); return (float)h; };
  auto hi = [](uint32_t v) { uint16_t u = (uint16_t)(v >> 16); _Float16 h; memcpy(&h, &u, 2); return (float)h; };
  __syncwarp();
  sa[w][g][2 * t] = lo(a[0]); sa[w][g][2 * t + 1] = hi(a[0]); sa[w][g + 8][2 * t] = lo(a[1]); sa[w][g + 8][2 * t + 1] = hi(a[1]);
  sa[w][g][2 * t + 8] = lo(a[2]); sa[w][g][2 * t + 9] = hi(a[2]); sa[w][g + 8][2 * t + 8] = lo(a[3]); sa[w][g + 8][2 * t + 9] = hi(a[3]);
  sb[w][2 * t][g] = lo(b[0]); sb[w][2 * t + 1][g] = hi(b[0]); sb[w][2 * t + 8][g] = lo(b[1]); sb[w][2 * t + 9][g] = hi(b[1]);
  __syncwarp();
  for(int k = 0; k < 16; k ++) {
    d[0] += sa[w][g][k] * sb[w][k][2 * t];     d[1] += sa[w][g][k] * sb[w][k][2 * t + 1];
    d[2] += sa[w][g + 8][k] * sb[w][k][2 * t]; d[3] += sa[w][g + 8][k] * sb[w][k][2 * t + 1];
  }
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
    : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
    : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#endif
}
// x = hi + lo / 2048: hi on FP16's 11-bit mantissa grid (exactly convertible while |x| >= 2^-14), lo the exact
// remainder scaled into FP16's normal range; products hi hi + (hi lo + lo hi) / 2048 carry 22 bits, as 3xTF32 does
#define F16_LO_SCALE 2048.0f
__device__ __forceinline__ void f16_split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = (x - hi) * F16_LO_SCALE;
}

// 2^ceil(e) as an exact integer (device pow() is only accurate to an ulp, and the reference
// truncates pow(2, ceil(log2(x))) to int: layer0.c:201, dsputils.c:485)
__host__ __device__ __forceinline__ int pow2_ceil(double e) {
  int k = (int)ceil(e);
  return k <= 0 ? 1 : (k >= 30 ? (1 << 30) : (1 << k));
}
