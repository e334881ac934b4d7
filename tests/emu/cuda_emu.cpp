// CUDA-on-threads emulation runtime -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local BlockCtx* t_ctx = nullptr;

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  int nthreads = (int)(block.x * block.y * block.z);
  int nwarps = (nthreads + 31) / 32;
  for(unsigned bz = 0; bz < grid.z; bz ++)
  for(unsigned by = 0; by < grid.y; by ++)
  for(unsigned bx = 0; bx < grid.x; bx ++) {
    BlockCtx ctx;
    ctx.nthreads = nthreads;
    ctx.smem.assign(smem_bytes + 64, 0);
    ctx.xch.assign(nthreads, 0);
    ctx.wbar.resize(nwarps);
    pthread_barrier_init(&ctx.bar, nullptr, nthreads);
    for(int w = 0; w < nwarps; w ++) {
      int cnt = std::min(32, nthreads - w * 32);
      pthread_barrier_init(&ctx.wbar[w], nullptr, cnt);
    }
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for(int t = 0; t < nthreads; t ++) {
      th.emplace_back([&, t]() {
        t_ctx = &ctx;
        t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        t_blockIdx = dim3(bx, by, bz);
        t_blockDim = block;
        t_gridDim = grid;
        body();
      });
    }
    for(auto& x : th) x.join();
    pthread_barrier_destroy(&ctx.bar);
    for(int w = 0; w < nwarps; w ++) pthread_barrier_destroy(&ctx.wbar[w]);
  }
}
}

static pthread_mutex_t g_atomic_mtx = PTHREAD_MUTEX_INITIALIZER;
float atomicAdd(float* p, float v) {
  pthread_mutex_lock(&g_atomic_mtx);
  float old = *p; *p = old + v;
  pthread_mutex_unlock(&g_atomic_mtx);
  return old;
}
