/*
  buffer.h -- sample ring buffers of the streaming synthesizer (the reference installs a header of this name with the
  same types and static-inline functions: /root/reference/buffer.h:32-36,146-151,218-223; used by llsmrt.c and by
  applications that keep their own look-back buffers). Written from the contract, not from the reference's text:

    llsm_ringbuffer   circular FP_TYPE array; `curr` is the next write position; every index is a LAG relative to it
                      (negative: -1 is the newest sample). read / write one sample, append (write at curr, advance),
                      forward (advance only), readchunk / writechunk / addchunk(lag, size, ptr) over lag .. lag + size
                      (must end at or before 0), appendchunk = forward + writechunk(-size), appendblank = forward +
                      zero fill                                                    (test/test-structs.c:168-214 pins them)
    llsm_dualbuffer   a pair of rings sharing `curr`: samples at negative offsets live in the backward ring, samples at
                      offsets >= 0 in the forward ring (overlap-add ahead of the present); forward(size) retires
                      `size` samples from the forward into the backward ring and zeroes their forward slots
    llsm_vringbuffer  ring of owned pointers with a destructor

  Differences from a sample-by-sample implementation: chunk operations are done as at most two contiguous runs
  (memcpy / vector loops) instead of a modulo per sample.
  FP_TYPE must be defined by the includer (as for llsm.h).
*/
#ifndef LLSM_BUFFER_H
#define LLSM_BUFFER_H

#include <assert.h>
#include <stdlib.h>
#include <string.h>

#ifndef LLSM_H
typedef void (*llsm_fdestructor)(void*);
#endif

typedef struct {
  FP_TYPE* data;
  int capacity;
  int curr;
} llsm_ringbuffer;

/* position of lag `idx` (negative) relative to curr, in [0, capacity) */
static inline int llsm_ring_at_(int curr, int capacity, int idx) {
  int p = (curr + idx) % capacity;
  return p < 0 ? p + capacity : p;
}

static inline llsm_ringbuffer* llsm_create_ringbuffer(int capacity) {
  assert(capacity > 0);
  llsm_ringbuffer* r = (llsm_ringbuffer*)malloc(sizeof(llsm_ringbuffer));
  r -> data = (FP_TYPE*)calloc((size_t)capacity, sizeof(FP_TYPE));
  r -> capacity = capacity;
  r -> curr = 0;
  return r;
}

static inline void llsm_delete_ringbuffer(llsm_ringbuffer* dst) {
  if(dst == NULL) return;
  free(dst -> data);
  free(dst);
}

static inline FP_TYPE llsm_ringbuffer_read(llsm_ringbuffer* src, int idx) {
  assert(idx < 0 && idx >= -src -> capacity);
  return src -> data[llsm_ring_at_(src -> curr, src -> capacity, idx)];
}

static inline void llsm_ringbuffer_write(llsm_ringbuffer* dst, int idx, FP_TYPE x) {
  assert(idx < 0 && idx >= -dst -> capacity);
  dst -> data[llsm_ring_at_(dst -> curr, dst -> capacity, idx)] = x;
}

static inline void llsm_ringbuffer_append(llsm_ringbuffer* dst, FP_TYPE x) {
  dst -> data[dst -> curr] = x;
  dst -> curr = dst -> curr + 1 == dst -> capacity ? 0 : dst -> curr + 1;
}

static inline void llsm_ringbuffer_forward(llsm_ringbuffer* dst, int size) {
  dst -> curr = llsm_ring_at_(dst -> curr, dst -> capacity, size);
}

/* the run lag .. lag + size as (first position, length of the part before the wrap) */
static inline int llsm_ring_run_(int curr, int capacity, int lag, int size, int* first) {
  *first = llsm_ring_at_(curr, capacity, lag);
  int head = capacity - *first;
  return head < size ? head : size;
}

static inline void llsm_ringbuffer_readchunk(llsm_ringbuffer* src, int lag, int size, FP_TYPE* dst) {
  assert(size > 0);
  assert(lag + size <= 0);
  assert(lag > -src -> capacity);
  int first, head = llsm_ring_run_(src -> curr, src -> capacity, lag, size, & first);
  memcpy(dst, src -> data + first, sizeof(FP_TYPE) * (size_t)head);
  if(size > head) memcpy(dst + head, src -> data, sizeof(FP_TYPE) * (size_t)(size - head));
}

static inline void llsm_ringbuffer_writechunk(llsm_ringbuffer* dst, int lag, int size, FP_TYPE* src) {
  assert(size > 0);
  assert(lag + size <= 0);
  assert(lag >= -dst -> capacity);
  int first, head = llsm_ring_run_(dst -> curr, dst -> capacity, lag, size, & first);
  memcpy(dst -> data + first, src, sizeof(FP_TYPE) * (size_t)head);
  if(size > head) memcpy(dst -> data, src + head, sizeof(FP_TYPE) * (size_t)(size - head));
}

static inline void llsm_ringbuffer_addchunk(llsm_ringbuffer* dst, int lag, int size, FP_TYPE* src) {
  assert(size > 0);
  assert(lag + size <= 0);
  assert(lag >= -dst -> capacity);
  int first, head = llsm_ring_run_(dst -> curr, dst -> capacity, lag, size, & first);
  for(int i = 0; i < head; i ++) dst -> data[first + i] += src[i];
  for(int i = head; i < size; i ++) dst -> data[i - head] += src[i];
}

static inline void llsm_ringbuffer_appendchunk(llsm_ringbuffer* dst, int size, FP_TYPE* src) {
  assert(size > 0);
  assert(size <= dst -> capacity);
  llsm_ringbuffer_forward(dst, size);
  llsm_ringbuffer_writechunk(dst, -size, size, src);
}

static inline void llsm_ringbuffer_appendblank(llsm_ringbuffer* dst, int size) {
  assert(size > 0);
  assert(size <= dst -> capacity);
  llsm_ringbuffer_forward(dst, size);
  int first, head = llsm_ring_run_(dst -> curr, dst -> capacity, -size, size, & first);
  memset(dst -> data + first, 0, sizeof(FP_TYPE) * (size_t)head);
  if(size > head) memset(dst -> data, 0, sizeof(FP_TYPE) * (size_t)(size - head));
}

typedef struct {
  FP_TYPE* data_frwd;
  FP_TYPE* data_bkwd;
  int capacity;
  int curr;
} llsm_dualbuffer;

static inline llsm_dualbuffer* llsm_create_dualbuffer(int capacity) {
  assert(capacity > 0);
  llsm_dualbuffer* r = (llsm_dualbuffer*)malloc(sizeof(llsm_dualbuffer));
  r -> data_frwd = (FP_TYPE*)calloc((size_t)capacity, sizeof(FP_TYPE));
  r -> data_bkwd = (FP_TYPE*)calloc((size_t)capacity, sizeof(FP_TYPE));
  r -> capacity = capacity;
  r -> curr = 0;
  return r;
}

static inline void llsm_delete_dualbuffer(llsm_dualbuffer* dst) {
  if(dst == NULL) return;
  free(dst -> data_frwd);
  free(dst -> data_bkwd);
  free(dst);
}

/* samples of the run offset .. offset + size that lie in the past (offset + i < 0) */
static inline int llsm_dual_past_(int offset, int size) {
  int past = offset > 0 ? 0 : -offset;
  return past > size ? size : past;
}

static inline void llsm_dualbuffer_readchunk(llsm_dualbuffer* src, int offset, int size, FP_TYPE* dst) {
  assert(size > 0);
  assert(size < src -> capacity);
  const int past = llsm_dual_past_(offset, size);
  for(int i = 0; i < size; i ++) {
    const FP_TYPE* ring = i < past ? src -> data_bkwd : src -> data_frwd;
    dst[i] = ring[llsm_ring_at_(src -> curr, src -> capacity, offset + i)];
  }
}

static inline void llsm_dualbuffer_forward(llsm_dualbuffer* dst, int size) {
  for(int i = 0; i < size; i ++) {
    dst -> data_bkwd[dst -> curr] = dst -> data_frwd[dst -> curr];
    dst -> data_frwd[dst -> curr] = 0;
    dst -> curr = dst -> curr + 1 == dst -> capacity ? 0 : dst -> curr + 1;
  }
}

static inline void llsm_dualbuffer_addchunk(llsm_dualbuffer* dst, int offset, int size, FP_TYPE* src) {
  assert(size > 0);
  assert(size < dst -> capacity);
  const int past = llsm_dual_past_(offset, size);
  for(int i = 0; i < size; i ++) {
    FP_TYPE* ring = i < past ? dst -> data_bkwd : dst -> data_frwd;
    ring[llsm_ring_at_(dst -> curr, dst -> capacity, offset + i)] += src[i];
  }
}

typedef struct {
  void** data;
  int capacity;
  int curr;
  llsm_fdestructor destructor;
} llsm_vringbuffer;

static inline llsm_vringbuffer* llsm_create_vringbuffer(int capacity, llsm_fdestructor destructor) {
  llsm_vringbuffer* r = (llsm_vringbuffer*)malloc(sizeof(llsm_vringbuffer));
  r -> data = (void**)calloc((size_t)capacity, sizeof(void*));
  r -> capacity = capacity;
  r -> curr = 0;
  r -> destructor = destructor;
  return r;
}

static inline void llsm_delete_vringbuffer(llsm_vringbuffer* dst) {
  if(dst == NULL) return;
  for(int i = 0; i < dst -> capacity; i ++)
    if(dst -> data[i] != NULL) dst -> destructor(dst -> data[i]);
  free(dst -> data);
  free(dst);
}

static inline void* llsm_vringbuffer_read(llsm_vringbuffer* src, int idx) {
  assert(idx < 0 && idx >= -src -> capacity);
  return src -> data[llsm_ring_at_(src -> curr, src -> capacity, idx)];
}

static inline void llsm_vringbuffer_write(llsm_vringbuffer* dst, int idx, void* x) {
  assert(idx < 0 && idx >= -dst -> capacity);
  void** slot = & dst -> data[llsm_ring_at_(dst -> curr, dst -> capacity, idx)];
  if(*slot != NULL) dst -> destructor(*slot);
  *slot = x;
}

static inline void llsm_vringbuffer_append(llsm_vringbuffer* dst, void* x) {
  void** slot = & dst -> data[dst -> curr];
  if(*slot != NULL) dst -> destructor(*slot);
  *slot = x;
  dst -> curr = dst -> curr + 1 == dst -> capacity ? 0 : dst -> curr + 1;
}

#endif
