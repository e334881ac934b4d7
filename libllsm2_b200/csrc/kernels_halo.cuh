// Frame-range sharding across GPUs: the two kernels around the one collective.
//
// Rank r synthesises the frames [edge[r], edge[r + 1]) of every utterance into full-length rows
// (llsm_b200_synthesize_l0_shard): partial sums, non-zero at most `halo` samples beyond the centres of its first and
// last frame. It OWNS the output samples [s_r, s_{r+1}), s_r = round(edge[r] thop fs) (the frame positions of
// layer0.c:127-128; s_0 = 0, s_world = ny). What its frames add outside that range is packed into two strips per
// component (y_sin, y_noise):
//     strip[b][0][c][j] = partial_c[b][s_r     - halo + j]      (spill into the ranks to the left)
//     strip[b][1][c][j] = partial_c[b][s_{r+1}        + j]      (spill into the ranks to the right), j < halo,
// ONE all-gather (ncclAllGather over NVLink) hands every rank every strip, and the edge-add kernel completes the owned
// range: for every other rank q, the part of q's strips that falls inside [s_r, s_{r+1}) is added (ascending q: fixed
// order), then y = y_sin + y_noise (layer0.c:657-659) on the owned range.
#pragma once
#include "common.cuh"

struct HaloParams {
  int nutt, ny, stride, halo, rank, world;
  const int* spos;            // [world + 1] owned-range boundaries s_q (device)
  float* y_sin; float* y_noise; float* y;     // [B][stride]; y may be NULL
  float* strips;              // pack: this rank's [B][2][2][halo]; add: gathered [world][B][2][2][halo]
};

__global__ void __launch_bounds__(256) halo_pack_kernel(HaloParams P) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;             // (side, component, j)
  if(e >= 4 * P.halo) return;
  const int side = e / (2 * P.halo), c = (e / P.halo) & 1, j = e % P.halo;
  const int n = side == 0 ? P.spos[P.rank] - P.halo + j : P.spos[P.rank + 1] + j;
  const float* src = (c == 0 ? P.y_sin : P.y_noise) + (size_t)b * P.stride;
  // the first rank owns from sample 0 and the last one to ny: nothing of theirs lies beyond those ends
  const bool in = n >= 0 && n < P.ny && ! (side == 0 && P.rank == 0) && ! (side == 1 && P.rank == P.world - 1);
  P.strips[(size_t)b * 4 * P.halo + e] = in ? src[n] : 0.f;
}

__global__ void __launch_bounds__(256) halo_add_kernel(HaloParams P) {
  const int b = blockIdx.y;
  const int sa = P.spos[P.rank], sb = P.spos[P.rank + 1];
  const int n = sa + blockIdx.x * blockDim.x + threadIdx.x;
  if(n >= sb) return;
  const size_t o = (size_t)b * P.stride + n;
  float vs = P.y_sin[o], vn = P.y_noise[o];
  const size_t per_rank = (size_t)P.nutt * 4 * P.halo;
  for(int q = 0; q < P.world; q ++) {
    if(q == P.rank) continue;
    const float* st = P.strips + (size_t)q * per_rank + (size_t)b * 4 * P.halo;
    const int jl = n - (P.spos[q] - P.halo);                       // index in q's left strip
    if(jl >= 0 && jl < P.halo) { vs += st[jl]; vn += st[P.halo + jl]; }
    const int jr = n - P.spos[q + 1];                              // index in q's right strip
    if(jr >= 0 && jr < P.halo) { vs += st[2 * P.halo + jr]; vn += st[3 * P.halo + jr]; }
  }
  P.y_sin[o] = vs; P.y_noise[o] = vn;
  if(P.y) P.y[o] = vs + vn;
}

static inline void run_halo_pack(const HaloParams& P, cudaStream_t st) {
  LLSM_LAUNCH(halo_pack_kernel, dim3((4 * P.halo + 255) / 256, P.nutt), dim3(256), 0, st, P);
}
// own = number of samples this rank owns (host copy of spos[rank + 1] - spos[rank])
static inline void run_halo_add(const HaloParams& P, int own, cudaStream_t st) {
  if(own <= 0) return;
  LLSM_LAUNCH(halo_add_kernel, dim3((own + 255) / 256, P.nutt), dim3(256), 0, st, P);
}
