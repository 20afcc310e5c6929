// Shared definitions for the libldmb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

typedef __nv_bfloat16 bf16;

// ---- A-operand addressing
enum { AM_ROWS = 0,    // A is a row-major [M, K] matrix (1x1 convolution on NHWC activations)
       AM_CONV3 = 1 }; // A is an NHWC tensor; K = 9*cC, k = tap*cC + c, 3x3 window, zero padding 1
// ---- epilogues
enum { EPI_STORE = 0,      // out(T)[m,n]   = act(acc + bias) (+ res)
       EPI_STORE_F32 = 1,  // out(f32)[m,n] = act(acc + bias)
       EPI_ACCUM_F32 = 2,  // out(f32)[m,n] += acc + bias           (residual stream)
       EPI_REGLU = 3,      // out(T)[m,j]   = (acc_a + bias_a) * relu(acc_b + bias_b)   (modules.py:15)
       EPI_CONVT = 4,      // ConvTranspose2d(k=2,s=2): column n = (dy*2+dx)*Cout + co scattered to pixel (2h+dy, 2w+dx)
       EPI_UPADD = 5 };    // out(f32)[2x2 pixels of the nearest-upsampled row m, n] += acc + bias  (unet.py:85,100-101: the 1x1
                           // conv commutes with nearest Upsample(2), so the conv runs at low resolution and its result is
                           // replicated into the skip-connected residual stream); ctH, ctW = low-resolution height, width
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

// One GEMM-shaped launch: acc[M, N] = A[M, K] * W[N, K]^T, then an epilogue.  Used by both the
// CUDA-core (validation) kernel and the tcgen05 kernel so that they are interchangeable.
struct GemmDesc {
  const void* A; long long lda;        // AM_ROWS: row stride (elements). AM_CONV3: channels per pixel of the NHWC tensor
  int amode; int cH, cW, cC;           // AM_CONV3: image height/width and channels per (group of the) convolution
  const void* W; long long ldw;        // weights [rows, K], row stride in elements
  const float* bias;                   // indexed like the weight rows
  void* out; long long ldo;
  const void* res; long long ldr;      // optional extra addend in T (EPI_STORE: after the activation; EPI_ACCUM_F32: another branch)
  int M, N, K;                         // N counts accumulator columns
  int epi, act; float slope;
  // expert selection (RandomMoE, modules.py:34-36): which rows of the stacked expert weights a slot uses
  int sel;                             // 0 none; 1 slot = n / sel_span (a|b GEMM); 2 slot = k / sel_span (c GEMM);
                                       // 3 like 2 with every slot in order: slot q uses rows q * sel_stride (per-image plans)
  int sel_span;
  int sel_rows[4];                     // first weight row (and bias index) of slot 0..3
  // Device-side plan entry {skip, e1, e2, -} of the SwinBlock this launch belongs to (NULL: use sel_rows as given).
  // Lets one captured CUDA graph serve every step although the Python-RNG decisions differ per step:
  // slot q > 0 uses rows (1 + e_q) * sel_stride, slot 3 (attention out_proj) rows 5 * sel_stride; skip => kernel exits.
  const int* plan; int sel_stride;
  int glu_chunk;                       // EPI_REGLU: a and b columns interleaved in chunks of this many columns
  // EPI_REGLU with per-image expert decisions: output column block q = j / mask_span (q >= 1: expert q - 1) of row m is
  // written as zero unless the image m / mask_rows drew that expert (word per image: skip | e1 << 8 | e2 << 16)
  const int* mask_plan; int mask_rows, mask_span;
  // grid.z batching (grouped convolution groups, per-block FiLM projections)
  int batch; long long a_koff_b, w_row_b, out_off_b, bias_off_b;
  int ctH, ctW, ctC;                   // EPI_CONVT: input height, width, output channels
  int max_ctas;                        // > 0: cap of the persistent tcgen05 grid (SM partitioning: another kernel runs on the other SMs)
};

// Per-call parameters that change every step; kernels read them from device memory so the launch sequence is static.
struct StepParams {
  const float* x_in; float* out; const float* noise;
  float c_eps_in, c_div, c_x0, c_eps_out, sigma;
  int final_step, ddim_enabled;
};

// Does row m of a per-image-masked ReGLU output keep column j?  (modules.py:35-36: general + the two drawn experts)
__device__ __forceinline__ bool expert_kept(const GemmDesc& d, int m, int j) {
  const int q = j / d.mask_span;
  if (q == 0) return true;
  const int w = d.mask_plan[(m < d.M ? m : d.M - 1) / d.mask_rows];
  return q - 1 == ((w >> 8) & 0xff) || q - 1 == ((w >> 16) & 0xff);
}

// Resolve the device-side plan into sel_rows; returns true when the block is skipped (stochastic depth, unet.py:39-40).
__device__ __forceinline__ bool resolve_plan(GemmDesc& d) {
  if (d.plan == nullptr) return false;
  if (d.plan[0] != 0) return true;
  if (d.sel == 1 || d.sel == 2) {
    d.sel_rows[0] = 0;
    d.sel_rows[1] = (1 + d.plan[1]) * d.sel_stride;
    d.sel_rows[2] = (1 + d.plan[2]) * d.sel_stride;
    d.sel_rows[3] = 5 * d.sel_stride;
  }
  return false;
}

// A pipeline watchdog fired: record the first fault code in device memory and mirror it to the host-mapped word whose
// address sits at fault[2..3] (tc_context.h), so the host sees it without synchronising.
__device__ __forceinline__ void report_fault(int* fault, int code) {
  if (atomicCAS(fault, 0, code) == 0) {
    volatile int* host = *reinterpret_cast<volatile int* const*>(fault + 2);
    if (host != nullptr) { *host = code; __threadfence_system(); }
  }
}

// ---- programmatic dependent launch (PDL): every kernel of the library is launched with the stream-serialization
// attribute and starts with pdl_wait(), so its launch latency / prologue overlaps the tail of the previous kernel.
// griddepcontrol.wait blocks until the preceding kernel has completed and its writes are visible; it is a no-op
// for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
extern bool g_ldmb_pdl;   // host toggle (LDMB_NO_PDL=1 disables), defined in kernels_simt.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_ldmb_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}

// Weight row / bias index of accumulator column n (for sel != 2).
__device__ __forceinline__ int wrow_of_col(const GemmDesc& d, int z, int n) {
  int row = n;
  if (d.sel == 1) row = d.sel_rows[n / d.sel_span] + n % d.sel_span;
  return row + (int)(z * d.w_row_b);
}
// Bias of accumulator column n (sums the per-slot biases for sel == 2).
__device__ __forceinline__ float bias_of_col(const GemmDesc& d, int z, int n) {
  if (d.bias == nullptr) return 0.f;
  const float* b = d.bias + z * d.bias_off_b;
  if (d.sel == 2 || d.sel == 3) {
    float s = 0.f;
    for (int q = 0; q < d.K / d.sel_span; ++q) s += b[(d.sel == 3 ? q * d.sel_stride : d.sel_rows[q]) + n];
    return s;
  }
  if (d.sel == 1) return b[d.sel_rows[n / d.sel_span] + n % d.sel_span];
  return b[n];
}
// EPI_CONVT destination offset (elements) of (input pixel m, accumulator column n).
__device__ __forceinline__ long long convt_offset(const GemmDesc& d, int m, int n) {
  const int HW = d.ctH * d.ctW;
  const int b = m / HW, r = m % HW, h = r / d.ctW, w = r % d.ctW;
  const int q = n / d.ctC, co = n % d.ctC, dy = q >> 1, dx = q & 1;
  return (((long long)b * (2 * d.ctH) + 2 * h + dy) * (2 * d.ctW) + 2 * w + dx) * d.ctC + co;
}
