// CUDA-on-threads emulation -- TEST INFRASTRUCTURE ONLY.
//
// Lets the *same* kernel sources under libllsm2_b200/csrc/ be compiled by g++ (-DLLSM_EMU) and run
// on the CPU, one OS thread per CUDA thread, with real barriers for __syncthreads/__syncwarp and
// exchange-array shuffles. The container this repo is developed in has no GPU, so the `-m "not gpu"`
// tests use this to check kernel logic (indexing, synchronisation, numerics) against the oracle.
// It is never linked into, imported by, or reachable from the product library or package: the
// product has no CPU path and fails loudly without the CUDA build (see libllsm2_b200/_lib.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <thread>
#include <vector>
#include <functional>
#include <algorithm>
#include <pthread.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 r = {a, b}; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r = {a, b, c, d}; return r; }
static inline double2 make_double2(double a, double b) { double2 r = {a, b}; return r; }
static inline int2 make_int2(int a, int b) { int2 r = {a, b}; return r; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = {a, b, c, d}; return r; }

namespace emu {
struct BlockCtx {
  pthread_barrier_t bar;
  std::vector<pthread_barrier_t> wbar;
  std::vector<uint64_t> xch;   // shuffle exchange, one slot per thread
  std::vector<char> smem;
  int nthreads;
};
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local BlockCtx* t_ctx;
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
}

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
#define LLSM_DYN_SMEM(name) char* name = emu::t_ctx->smem.data()

static inline void __syncthreads() { pthread_barrier_wait(&emu::t_ctx->bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  pthread_barrier_wait(&emu::t_ctx->wbar[emu::t_threadIdx.x / 32]);
}
template <class T> static inline T emu_shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::BlockCtx* c = emu::t_ctx;
  int tid = emu::t_threadIdx.x, w0 = tid & ~31;
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  c->xch[tid] = raw;
  __syncwarp();
  int src = w0 + (src_lane & 31);
  if(src >= c->nthreads) src = tid;
  uint64_t got = c->xch[src];
  __syncwarp();
  T out; memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_shfl_from(v, lane); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return emu_shfl_from(v, (int)(emu::t_threadIdx.x & 31) ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  int l = (int)(emu::t_threadIdx.x & 31) + d; return emu_shfl_from(v, l > 31 ? (int)(emu::t_threadIdx.x & 31) : l);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  int l = (int)(emu::t_threadIdx.x & 31) - d; return emu_shfl_from(v, l < 0 ? (int)(emu::t_threadIdx.x & 31) : l);
}
float atomicAdd(float* p, float v); // mutex-serialised (cuda_emu.cpp)
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// device math used by the kernels
static inline void sincospif(float x, float* s, float* c) {
  double r = std::fmod((double)x, 2.0); *s = (float)std::sin(M_PI * r); *c = (float)std::cos(M_PI * r);
}
static inline void sincospi(double x, double* s, double* c) {
  double r = std::fmod(x, 2.0); *s = std::sin(M_PI * r); *c = std::cos(M_PI * r);
}
static inline float cospif(float x) { double r = std::fmod((double)x, 2.0); return (float)std::cos(M_PI * r); }
static inline float sinpif(float x) { double r = std::fmod((double)x, 2.0); return (float)std::sin(M_PI * r); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline double cospi(double x) { return cos(3.14159265358979323846 * x); }
static inline double sinpi(double x) { return sin(3.14159265358979323846 * x); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline float __ldg(const float* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline float2 __ldg(const float2* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __int2float_rn(int i) { return (float)i; }
static inline int __float2int_rn(float f) { return (int)std::nearbyint(f); }
using std::min; using std::max;
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
#define __shared__ static   /* blocks run one at a time in the emulator */
