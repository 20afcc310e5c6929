// Host-side state shared by the tcgen05 kernels' launchers (kernels_tc.cu, kernels_gconv.cu).
#pragma once
#include <cuda.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcContext {
  EncodeTiledFn encode;
  int num_sms;
  int* fault_dev;
  int device;
  long long* trace_dev;   // debug: per-CTA %globaltimer stamps of the last launch (NULL unless enabled)
  bool splitk;            // split-K for the residual GEMMs (off in deterministic mode: slices reduce-add in arrival order)
};
