/* TEST INFRASTRUCTURE. Stand-in for the pitch tracker header that test/demo-stretch.c includes (the demo calls
   pyin_* and ciglet's wavread / wavwrite from its main()). oracle/ref_harness.c compiles the demo only for its
   file-static interpolation functions; main() is renamed and never called, these stubs only let it compile and
   link. Nothing here is an algorithm of the reference. */
#ifndef ORACLE_NEBULA_STUB_H
#define ORACLE_NEBULA_STUB_H
#include <stddef.h>
#ifndef FP_TYPE
#define FP_TYPE float
#endif
typedef struct { FP_TYPE fmin, fmax; int trange, bias, nf; } pyin_config;
static inline pyin_config pyin_init(int nhop) { pyin_config c = {0, 0, 0, 0, 0}; (void)nhop; return c; }
static inline FP_TYPE* pyin_analyze(pyin_config c, FP_TYPE* x, int nx, int fs, int* nfrm) {
  (void)c; (void)x; (void)nx; (void)fs; *nfrm = 0; return NULL;
}
static inline FP_TYPE* wavread(const char* path, int* fs, int* nbit, int* nx) {
  (void)path; *fs = 0; *nbit = 0; *nx = 0; return NULL;
}
static inline void wavwrite(FP_TYPE* y, int ny, int fs, int nbit, const char* path) {
  (void)y; (void)ny; (void)fs; (void)nbit; (void)path;
}
#endif
