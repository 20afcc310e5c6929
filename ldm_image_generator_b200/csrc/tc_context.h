// Host-side state shared by the tcgen05 kernels' launchers (kernels_tc.cu, kernels_gconv.cu).
#pragma once
#include <cuda.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Debug / tuning knobs, read from the environment ONCE (first use), never on the launch path.
struct TcKnobs {
  int grid, stages, dbg, splits, force_cg, no_splitk, mlp_dbg, gconv_dbg, attn_dbg, ast, astp;
};
const TcKnobs& tc_knobs();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: one bit per CUDA device per kernel instance.
struct PerDeviceOnce {
  unsigned long long done[4] = {0, 0, 0, 0};
  bool need(int device) const { return device < 0 || device >= 256 || !((done[device >> 6] >> (device & 63)) & 1ull); }
  void mark(int device) { if (device >= 0 && device < 256) done[device >> 6] |= 1ull << (device & 63); }
};

struct TcContext {
  EncodeTiledFn encode;
  int num_sms;
  // fault word of the pipeline watchdogs: fault_dev[0] = first fault code (device memory, atomicCAS); fault_dev[2..3] = address of
  // a host-mapped mirror the faulting thread also writes, so the host can poll it without synchronising (fault_host)
  int* fault_dev;
  volatile int* fault_host;
  int device;
  long long* trace_dev;   // debug: per-CTA %globaltimer stamps of the last launch (NULL unless enabled)
  bool splitk;            // split-K for the residual GEMMs (off in deterministic mode: slices reduce-add in arrival order)
};
