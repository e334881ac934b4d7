/* Growl-like llsm_pbpeffect for bench.py's C3 leg (BASELINE configs[2]): the modifier of the reference's own test
   (test/test-pbpeffects.c:70-85: LFO-modulated oscillator on Fa / Rk / Ee, jitter on the pulse period), written against
   the per-pulse hook of llsm_b200.h with one state per stream and a private generator instead of ciglet's randn.
   Built by bench.py with gcc at run time; host code by contract (the callback is the caller's). */
#include <math.h>

typedef struct { int period_count; float osc; unsigned rng; } growl_state;

static float unit_uniform(unsigned* s) { *s = *s * 1664525u + 1013904223u; return ((*s >> 8) + 1.0f) / 16777217.0f; }

int growl_hook(void* user, int utt, int frame, float* Fa, float* Rk, float* Rg, float* T0, float* Ee, float* delta_t) {
  growl_state* st = (growl_state*)user + utt;
  (void)frame; (void)Rg;
  st->period_count ++;
  const float lfo = sinf(st->period_count * 2.0f * 3.14159265f / 50.0f);
  st->osc += 2.0f * 3.14159265f / (6.0f + lfo);
  const float osc = sinf(st->osc);
  const float u1 = unit_uniform(&st->rng), u2 = unit_uniform(&st->rng);
  const float gauss = sqrtf(-2.0f * logf(u1)) * cosf(2.0f * 3.14159265f * u2);
  *delta_t = *T0 * 0.01f * gauss;
  *Fa *= 1.0f - osc * 0.5f;
  *Rk *= 1.0f + osc * 0.3f;
  *Ee *= 1.0f - osc * 0.5f;
  return 1;
}
