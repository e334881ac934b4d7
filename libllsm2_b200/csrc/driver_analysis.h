// Launch orchestration of the layer-0 analysis path (llsm_analyze, layer0.c:478-511).
#pragma once
#include "driver.h"

struct AnaKey {
  int nfrm, nx; float fs, thop;
  bool operator<(const AnaKey& o) const { return memcmp(this, &o, sizeof(AnaKey)) < 0; }
};
struct AnaPlanDev { void release() {} };
struct AnaScratch { void release() {} };
