// Warp-level 1024-point complex FFT held in registers (32 lanes x 32 points per lane).
//
// Cooley-Tukey split N = 32 x 32: with n = n1 + 32 n2 (n1 = lane) and k = k2 + 32 k1,
//   X[k2 + 32 k1] = sum_{n1} W32^{n1 k1} [ W1024^{n1 k2} sum_{n2} x[n1 + 32 n2] W32^{n2 k2} ].
// Pass 1: each lane runs a 32-point FFT over n2 entirely in registers; the result is multiplied
// by W1024^{lane k2}; a padded shared-memory transpose hands lane k2 the 32 values over n1; pass 2
// is another in-register 32-point FFT. Input and output use the same distribution: lane l holds
// element l + 32 r in register r. No block-level synchronisation, two __syncwarp per transform.
#pragma once
#include "common.cuh"

#define WFFT_ROW 33                       // padded row (float2 units) of the transpose scratch
#define WFFT_SCRATCH_BYTES (32 * WFFT_ROW * 8)

__device__ __forceinline__ constexpr int wfft_brev5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// cos(2 pi m / 32), sin(2 pi m / 32), m = 0..15
__device__ __forceinline__ constexpr float wfft_c32(int m) {
  constexpr float t[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
    0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
    0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
    -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
  return t[m];
}
__device__ __forceinline__ constexpr float wfft_s32(int m) {
  constexpr float t[16] = {0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
    0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
    1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
    0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};
  return t[m];
}

// In-register radix-2 decimation-in-frequency FFT of 32 points; the output lands in bit-reversed
// order: v[brev5(k)] = X[k]. INV: conjugate twiddles, unscaled.
template <bool INV>
__device__ __forceinline__ void fft32_regs(float2 (&v)[32]) {
#pragma unroll
  for(int h = 16; h >= 1; h >>= 1) {
#pragma unroll
    for(int base = 0; base < 32; base += 2 * h) {
#pragma unroll
      for(int j = 0; j < h; j ++) {
        const int m = j * (16 / h);              // twiddle exponent on the 32-point circle
        const float c = wfft_c32(m), s = INV ? wfft_s32(m) : -wfft_s32(m);
        float2 a = v[base + j], b = v[base + j + h];
        v[base + j] = make_float2(a.x + b.x, a.y + b.y);
        float dx = a.x - b.x, dy = a.y - b.y;
        if(m == 0) v[base + j + h] = make_float2(dx, dy);
        else if(m == 8) v[base + j + h] = INV ? make_float2(-dy, dx) : make_float2(dy, -dx);
        else v[base + j + h] = make_float2(dx * c - dy * s, dx * s + dy * c);
      }
    }
  }
}

// Twiddle table for pass 1 -> pass 2: tw2[k2 * 32 + lane] = exp(-2 pi i lane k2 / 1024), built once
// per CTA in shared memory from the full-circle table tw1024[m] = exp(-2 pi i m / 1024).
__device__ __forceinline__ void wfft_build_tw2(float2* tw2, const float2* __restrict__ tw1024) {
  for(int e = threadIdx.x; e < 1024; e += blockDim.x) {
    int k2 = e >> 5, l = e & 31;
    tw2[e] = tw1024[(l * k2) & 1023];
  }
}

// x[r] = element (lane + 32 r) on entry and on exit. `scratch` = WFFT_SCRATCH_BYTES of shared
// memory private to the warp. Forward transform; an inverse (unscaled) transform is obtained by
// conjugating the data before and after (`inv` != 0). The two 32-point passes share one copy of
// the butterfly code (rolled 2-iteration loop) to keep the kernel inside the instruction cache.
__device__ __forceinline__ void warp_fft1024(float2 (&x)[32], float2* scratch, const float2* tw2, int lane, int inv) {
  if(inv) {
#pragma unroll
    for(int r = 0; r < 32; r ++) x[r].y = -x[r].y;
  }
#pragma unroll 1
  for(int stage = 0; stage < 2; stage ++) {
    fft32_regs<false>(x);                    // x[brev(k)] = DFT32(x)[k]
    if(stage == 0) {
#pragma unroll
      for(int k2 = 0; k2 < 32; k2 ++) {
        float2 a = x[wfft_brev5(k2)];
        float2 w = tw2[k2 * 32 + lane];
        scratch[lane * WFFT_ROW + k2] = make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
      }
      __syncwarp();
#pragma unroll
      for(int n1 = 0; n1 < 32; n1 ++) x[n1] = scratch[n1 * WFFT_ROW + lane];
      __syncwarp();
    }
  }
  // natural order (a compile-time register permutation); undo the conjugation for the inverse
  float2 t[32];
#pragma unroll
  for(int k1 = 0; k1 < 32; k1 ++) t[k1] = x[wfft_brev5(k1)];
#pragma unroll
  for(int k1 = 0; k1 < 32; k1 ++) x[k1] = make_float2(t[k1].x, inv ? -t[k1].y : t[k1].y);
}
