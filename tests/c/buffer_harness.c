/* Test harness for buffer.h (ours: include/buffer.h; the reference's: /root/reference/buffer.h -- same API).
   Compiled twice by tests/test_buffer.py with a different include path; exports
     buffer_equivalences()   the reference's own ring-buffer test (test/test-structs.c:168-214): pairs of call
                             sequences that must give identical results, exact equality; returns 0 when all hold
     buffer_script(...)      a seeded pseudo-random sequence of every ring / dual-buffer operation; every value read is
                             appended to `out` -- two implementations of the contract must produce identical logs. */
#define FP_TYPE float
#include <stdlib.h>
typedef void (*llsm_fdestructor)(void*);
#define LLSM_H                 /* the typedef above stands in for llsm.h */
#include "buffer.h"

static unsigned lcg(unsigned* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

int buffer_equivalences(void) {
  llsm_ringbuffer* rb1 = llsm_create_ringbuffer(4096);
  float x[100], y[200];
  unsigned s = 7;
  int bad = 0;
  for(int i = 0; i < 100; i ++) x[i] = (float)(lcg(& s) % 10000) / 10000.0f - 0.5f;
  for(int i = 0; i < 100; i ++) {
    if(i % 2 == 0) llsm_ringbuffer_appendchunk(rb1, 100, x);
    else { llsm_ringbuffer_forward(rb1, 100); llsm_ringbuffer_writechunk(rb1, -100, 100, x); }
    if(i % 3 == 0) llsm_ringbuffer_appendblank(rb1, 100);
    else for(int j = 0; j < 100; j ++) llsm_ringbuffer_append(rb1, 0);
    if(i % 4 == 0) for(int j = 0; j < 200; j ++) y[j] = llsm_ringbuffer_read(rb1, j - 200);
    else llsm_ringbuffer_readchunk(rb1, -200, 200, y);
    for(int j = 0; j < 100; j ++) bad += y[j] != x[j];
    for(int j = 100; j < 200; j ++) bad += y[j] != 0;
    if(i <= 5) continue;
    llsm_ringbuffer_readchunk(rb1, -1300, 200, y);
    for(int j = 0; j < 100; j ++) bad += y[j] != 0;
    for(int j = 100; j < 200; j ++) bad += y[j] != x[j - 100];
  }
  llsm_delete_ringbuffer(rb1);
  return bad;
}

static int g_destroyed = 0;
static void count_free(void* p) { g_destroyed ++; free(p); }

int buffer_script(unsigned seed, int nops, int capacity, float* out, int nout) {
  llsm_ringbuffer* rb = llsm_create_ringbuffer(capacity);
  llsm_dualbuffer* db = llsm_create_dualbuffer(capacity);
  llsm_vringbuffer* vb = llsm_create_vringbuffer(8, count_free);
  float* tmp = malloc(sizeof(float) * capacity);
  unsigned s = seed;
  int n = 0;
  g_destroyed = 0;
  for(int op = 0; op < nops && n + capacity + 8 < nout; op ++) {
    int kind = lcg(& s) % 12;
    int size = 1 + lcg(& s) % (capacity / 3);
    int lag = -size - (int)(lcg(& s) % (capacity - size));            /* lag + size <= 0, lag > -capacity */
    for(int i = 0; i < size; i ++) tmp[i] = (float)(lcg(& s) % 2001) / 1000.0f - 1.0f;
    switch(kind) {
      case 0: llsm_ringbuffer_append(rb, tmp[0]); break;
      case 1: llsm_ringbuffer_forward(rb, size); break;
      case 2: llsm_ringbuffer_writechunk(rb, lag, size, tmp); break;
      case 3: llsm_ringbuffer_addchunk(rb, lag, size, tmp); break;
      case 4: llsm_ringbuffer_appendchunk(rb, size, tmp); break;
      case 5: llsm_ringbuffer_appendblank(rb, size); break;
      case 6: llsm_ringbuffer_write(rb, lag, tmp[0]); out[n ++] = llsm_ringbuffer_read(rb, lag); break;
      case 7: llsm_ringbuffer_readchunk(rb, lag, size, out + n); n += size; break;
      case 8: { int off = (int)(lcg(& s) % (2 * size + 1)) - size;    /* straddles the present */
                llsm_dualbuffer_addchunk(db, off, size, tmp); break; }
      case 9: llsm_dualbuffer_forward(db, size); break;
      case 10: { int off = (int)(lcg(& s) % (2 * size + 1)) - size;
                 llsm_dualbuffer_readchunk(db, off, size, out + n); n += size; break; }
      default: { float* obj = malloc(sizeof(float)); *obj = tmp[0];
                 if(lcg(& s) % 2) llsm_vringbuffer_append(vb, obj); else llsm_vringbuffer_write(vb, -1 - (int)(lcg(& s) % 8), obj);
                 float* back = llsm_vringbuffer_read(vb, -1 - (int)(lcg(& s) % 8));
                 out[n ++] = back != NULL ? *back : -9.0f; break; }
    }
  }
  out[n ++] = (float)rb -> curr; out[n ++] = (float)db -> curr; out[n ++] = (float)g_destroyed;
  llsm_delete_ringbuffer(rb); llsm_delete_dualbuffer(db); llsm_delete_vringbuffer(vb);
  out[n ++] = (float)g_destroyed;
  free(tmp);
  return n;
}
