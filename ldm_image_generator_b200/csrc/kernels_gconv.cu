// Grouped 3x3 convolution of the SwinBlock (unet.py:30,44; groups of 32 channels) on tcgen05, reading every
// activation ONCE: x += conv3x3_grouped(xm) + bias.
//
// The implicit-GEMM conv in kernels_tc.cu issues 9 shifted TMA box loads per tile (9x the activation bytes from L2,
// which is what bounded it).  Here a tile's (TH+2) x (TW+2) halo patch of a 64-channel slice (= a pair of groups,
// block-diagonal 64->64 weights) is loaded by ONE 4-d TMA box (out-of-bounds pixels zero-filled = the conv padding)
// into 128-byte-swizzled shared memory, rows = patch pixels in row-major order with pitch TW+2.  Output row i of the
// MMA tile is patch position i (y = i / (TW+2), x = i % (TW+2)), so tap (dy,dx) reads the SAME smem patch at row
// offset dy*(TW+2)+dx: nine UMMA descriptor start addresses into one buffer, no re-load.  Rows with x >= TW or
// y >= TH are garbage and dropped by the epilogue.  The 128-byte swizzle is a function of the shared-memory ADDRESS
// bits (both for the TMA write and the UMMA read), so a start address that is a whole number of 128-byte rows into
// the 1024-byte swizzle atom needs nothing else: the descriptor's base-offset field stays 0 (measured on B200:
// setting it to (addr >> 7) & 7 gives wrong results, leaving it 0 is exact).
// The slice's 9 x [64 x 64] weight tiles stay resident in shared memory for all tiles of a persistent CTA.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2..5 epilogue (TMEM lane quadrant each: tcgen05.ld -> + bias ->
// red.global.add.v4.f32 into the fp32 residual stream).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tc_context.h"

namespace {

constexpr int kSlice = 64;                 // channels per CTA slice: two groups of 32
constexpr int kWTile = kSlice * 128;       // one tap's [64 co x 64 ci] bf16 tile, 128 B rows
constexpr int kWBytes = 9 * kWTile;        // 72 KB resident weights
constexpr int kThreads = 64 + 4 * 32;

struct GconvGeom {
  int B, H, W, C;
  int TW, TH, TB, pitch;                   // tile: TB images x TH rows x TW columns; patch pitch = TW + 2
  int w_tiles, h_tiles, b_tiles;           // spatial tiling
  int a_bytes, stage_bytes, stages;
  int out_rows, out_half, out_bytes;       // staged residual update: TB*TH*TW compact rows x 2 halves of 32 fp32 channels, each half
                                           // 1024-aligned (swizzle phase); out_bytes = 2 * out_half (0: red.global path)
  int dbg;                                 // debug experiments: 1 = no reds, 2 = one tap only
  int sp_lo, sp_hi;                        // this launch covers the spatial tiles [sp_lo, sp_hi) of every channel slice (a conv split over two launches)
  // dense mode (VAE 3x3 conv with 64 input and 64 output channels, vae.py:57-58): out(bf16) = act(acc + bias) (+ res)
  int dense; float slope; bf16* out; const bf16* res;
};

__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* fault, int code) {
  if (ptx::mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (true) {
    if (ptx::mbar_try_wait(bar, parity)) return true;
    if (*s_abort) return false;
    if (clock64() - t0 > 3000000000LL) {
      *s_abort = 1;
      report_fault(fault, code);
      return false;
    }
  }
}

__device__ __forceinline__ void trace_stamp(long long* trace, int slot) {     // debug: same 16-slot layout as kernels_tc.cu
  if (trace != nullptr) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 16 + slot] = t;
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
gconv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                  float* __restrict__ x,
                  const float* __restrict__ bias, const GconvGeom g, const int* __restrict__ plan, int* fault, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* wts = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = wts + kWBytes;
  uint8_t* outbuf = stages + g.stages * g.stage_bytes;            // 2 x out_bytes, 1024-aligned (stage_bytes is a multiple of 1024)
  uint64_t* full = reinterpret_cast<uint64_t*>(outbuf + 2 * g.out_bytes);
  uint64_t* empty = full + 4;
  uint64_t* tfull = empty + 4;
  uint64_t* tempty = tfull + 2;
  uint64_t* wbar = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) trace_stamp(trace, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 4); }
    ptx::mbar_init(wbar, 1);
    *s_abort = 0;
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmW);
    if (g.out_bytes) ptx::prefetch_tensormap(&tmX);
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 128); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // plan and weights are older than the previous kernel: read / requested before waiting on it
  const bool skip_block = plan != nullptr && plan[0] != 0;       // stochastic depth (unet.py:39-40)
  const bool is_producer = warp == 0 && lane == 0;
  if (!is_producer) pdl_wait();

  // persistent CTA: fixed channel slice z (weights stay resident), strided over the spatial tiles
  const int nz = g.C / kSlice;
  const int z = blockIdx.x % nz;
  const int n_sp = g.b_tiles * g.h_tiles * g.w_tiles;
  const int sp0 = g.sp_lo + blockIdx.x / nz, sp_step = gridDim.x / nz;
  const int sp_end = skip_block ? 0 : (g.sp_hi < n_sp ? g.sp_hi : n_sp);

  if (warp == 0) {
    if (lane == 0 && sp0 < sp_end) {
      ptx::mbar_arrive_expect_tx(wbar, kWBytes);
      for (int tap = 0; tap < 9; ++tap) ptx::tma_load_2d(wts + tap * kWTile, &tmW, wbar, tap * kSlice, z * kSlice);
      trace_stamp(trace, 1);
      pdl_wait();
      trace_stamp(trace, 2);
      uint32_t stage = 0, phase = 0;
      for (int sp = sp0; sp < sp_end; sp += sp_step) {
        const int wt = sp % g.w_tiles, ht = (sp / g.w_tiles) % g.h_tiles, bt = sp / (g.w_tiles * g.h_tiles);
        if (!wait_bar(&empty[stage], phase ^ 1, s_abort, fault, 11)) break;
        ptx::mbar_arrive_expect_tx(&full[stage], g.a_bytes);
        ptx::tma_load_4d(stages + stage * g.stage_bytes, &tmA, &full[stage], z * kSlice, wt * g.TW - 1, ht * g.TH - 1, bt * g.TB);
        if (sp == sp0) trace_stamp(trace, 3);
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
      }
    } else if (lane == 0) {
      pdl_wait();
    }
    __syncwarp();
  } else if (warp == 1) {
    if (sp0 < sp_end) {   // the whole warp runs the loop (uniform control flow: descriptors stay in uniform registers); one elected lane issues
      const bool issuer = ptx::elect_one();
      constexpr uint32_t idesc = ptx::idesc_bf16(128, kSlice);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      bool ok = wait_bar(wbar, 0, s_abort, fault, 12);
      if (issuer) trace_stamp(trace, 10);
      for (int sp = sp0; ok && sp < sp_end; sp += sp_step) {
        if (!wait_bar(&tempty[as], aphase ^ 1, s_abort, fault, 13)) break;
        if (!wait_bar(&full[stage], phase, s_abort, fault, 14)) break;
        if (sp == sp0 && issuer) trace_stamp(trace, 4);
        ptx::tc_fence_after();
        // The MMA thread is issue-bound (ncu: ~21 SASS instructions per tcgen05.mma, most of them building the two
        // shared-memory descriptors and moving them to uniform registers): the descriptors of a tile differ only in their
        // 14-bit start-address field, so build one per tile / per kernel and ADD the (tap, k) offsets in 16-byte units.
        const uint64_t a_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(stages + stage * g.stage_bytes));
        const uint64_t w_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(wts));
        const uint32_t d_tmem = tmem_base + as * kSlice;
        const uint32_t pitch8 = (uint32_t)g.pitch * 8u;            // one patch row = 128 B = 8 descriptor units
        constexpr uint32_t idesc32 = ptx::idesc_bf16(128, 32);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if ((g.dbg & 2) && tap > 0) break;
          const uint64_t a_tap = a_desc0 + (uint64_t)((g.dbg & 4) ? (tap / 3) * 64u : (tap / 3) * pitch8 + (tap % 3) * 8u);   // dbg 4: 8-row aligned offsets (wrong results)
          const uint64_t w_tap = w_desc0 + (uint64_t)(tap * (kWTile / 16));
          if (g.dense || (g.dbg & 8)) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (issuer) ptx::umma_f16(d_tmem, a_tap + 2 * k, w_tap + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
          } else {
            // grouped: the slice's [64 out x 64 in] tap matrix is block-diagonal (two groups of 32 channels), so k-steps
            // 0-1 (inputs of group 0) only feed output columns 0..31 and k-steps 2-3 columns 32..63: N = 32 MMAs on
            // the matching 32 weight rows
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t grp = k >> 1;
              if (issuer) ptx::umma_f16(d_tmem + grp * 32, a_tap + 2 * k, w_tap + grp * (32 * 128 / 16) + 2 * k, idesc32, (tap | (k & 1)) != 0 ? 1u : 0u);
            }
          }
        }
        if (issuer) { ptx::umma_commit(&empty[stage]); ptx::umma_commit(&tfull[as]); }
        __syncwarp();
        if (++stage == (uint32_t)g.stages) { stage = 0; phase ^= 1; }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
      if (issuer) trace_stamp(trace, 5);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int i = q * 32 + lane;                   // output row = patch position
    const int per_img = (g.TH + 2) * g.pitch;
    const int img = i / per_img, r = i % per_img, yy = r / g.pitch, xx = r % g.pitch;
    const float* bz = bias + z * kSlice;
    uint32_t as = 0, aphase = 0;
    const bool staged = g.out_bytes != 0;
    // staged residual update: the tile's valid rows are packed ([image][y][x], 128-byte rows of 32 fp32 channels, two halves,
    // 128B-swizzled) and added to x by TWO bulk tensor reduce-adds per tile instead of 16 red.global.add.v4 per pixel
    const int cr = (img * g.TH + yy) * g.TW + xx;                  // compact row of this thread's pixel
    const bool in_tile = img < g.TB && yy < g.TH && xx < g.TW;
    const uint32_t sw = static_cast<uint32_t>(cr & 7);
    int n_done = 0;
    for (int sp = sp0; sp < sp_end; sp += sp_step, ++n_done) {
      const int wt = sp % g.w_tiles, ht = (sp / g.w_tiles) % g.h_tiles, bt = sp / (g.w_tiles * g.h_tiles);
      const int b = bt * g.TB + img, hh = ht * g.TH + yy, ww = wt * g.TW + xx;
      const bool valid = img < g.TB && b < g.B && yy < g.TH && hh < g.H && xx < g.TW && ww < g.W;
      const long long pix = (((long long)b * g.H + hh) * g.W + ww) * g.C + z * kSlice;
      float* orow = x + pix;
      if (!wait_bar(&tfull[as], aphase, s_abort, fault, 15)) break;
      if (threadIdx.x == 64) trace_stamp(trace, sp == sp0 ? 6 : 7);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kSlice;
      uint8_t* ob = outbuf + (n_done & 1) * g.out_bytes;
      if (staged && !g.dense) {
        if (threadIdx.x == 64) ptx::bulk_wait_read<1>();           // the reduce that last read this buffer (two tiles ago) is done with it
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_row + half * 32, v);
        ptx::tmem_ld_wait();
        if (valid && g.dense) {
          // act(v) = max(v,0) + slope*min(v,0); one full 64-byte half row per thread
          uint32_t pk[16];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bz + half * 32 + 4 * u));
            float t0 = __uint_as_float(v[4 * u]) + bb.x, t1 = __uint_as_float(v[4 * u + 1]) + bb.y;
            float t2 = __uint_as_float(v[4 * u + 2]) + bb.z, t3 = __uint_as_float(v[4 * u + 3]) + bb.w;
            t0 = fmaxf(t0, 0.f) + g.slope * fminf(t0, 0.f); t1 = fmaxf(t1, 0.f) + g.slope * fminf(t1, 0.f);
            t2 = fmaxf(t2, 0.f) + g.slope * fminf(t2, 0.f); t3 = fmaxf(t3, 0.f) + g.slope * fminf(t3, 0.f);
            if (g.res != nullptr) {
              const uint2 rr = __ldg(reinterpret_cast<const uint2*>(g.res + pix + half * 32 + 4 * u));
              const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&rr.x), r23 = *reinterpret_cast<const __nv_bfloat162*>(&rr.y);
              t0 += __low2float(r01); t1 += __high2float(r01); t2 += __low2float(r23); t3 += __high2float(r23);
            }
            __nv_bfloat162 p01 = __floats2bfloat162_rn(t0, t1), p23 = __floats2bfloat162_rn(t2, t3);
            pk[2 * u] = *reinterpret_cast<uint32_t*>(&p01); pk[2 * u + 1] = *reinterpret_cast<uint32_t*>(&p23);
          }
          uint4* op = reinterpret_cast<uint4*>(g.out + pix + half * 32);
#pragma unroll
          for (int u = 0; u < 4; ++u) op[u] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        } else if (staged && !g.dense) {
          if (in_tile) {
            const uint32_t srow = ptx::smem_u32(ob) + half * g.out_half + cr * 128;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bz + half * 32 + 4 * u));
              ptx::st_shared_v4(srow + ((u ^ sw) << 4), __float_as_uint(__uint_as_float(v[4 * u]) + bb.x), __float_as_uint(__uint_as_float(v[4 * u + 1]) + bb.y),
                                __float_as_uint(__uint_as_float(v[4 * u + 2]) + bb.z), __float_as_uint(__uint_as_float(v[4 * u + 3]) + bb.w));
            }
          }
        } else if (valid && !(g.dbg & 1)) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bz + half * 32 + 4 * u));
            red_add_v4(orow + half * 32 + 4 * u, __uint_as_float(v[4 * u]) + bb.x, __uint_as_float(v[4 * u + 1]) + bb.y,
                       __uint_as_float(v[4 * u + 2]) + bb.z, __uint_as_float(v[4 * u + 3]) + bb.w);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[as]);
      if (staged && !g.dense) {
        ptx::fence_proxy_async();                                  // generic-proxy staging writes -> the bulk reduce's async-proxy reads
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (threadIdx.x == 64 && !(g.dbg & 1)) {                   // rows / images beyond the tensor are clipped by the tensor map
          ptx::tma_reduce_add_4d(&tmX, ob, z * kSlice, wt * g.TW, ht * g.TH, bt * g.TB);
          ptx::tma_reduce_add_4d(&tmX, ob + g.out_half, z * kSlice + 32, wt * g.TW, ht * g.TH, bt * g.TB);
          ptx::bulk_commit();
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (staged && !g.dense && threadIdx.x == 64) ptx::bulk_wait_read<0>();     // the staging buffers stay valid until they have been read
  }
  if (threadIdx.x == 64) trace_stamp(trace, 8);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 128);
  if (threadIdx.x == 0) trace_stamp(trace, 9);
}


}  // namespace

static int g_gconv_mode = getenv("LDMB_GCONV_HALO") ? atoi(getenv("LDMB_GCONV_HALO")) : 1;   // debug: 0 = generic 9-tap-load kernel

bool gconv_halo_supported(int B, int H, int W, int C) {
  return g_gconv_mode != 0 && C % kSlice == 0 && C / kSlice <= 74 && B >= 1 && H >= 1 && W >= 1;
}

// x fp32 [B,H,W,C] += conv3x3(xm bf16 [B,H,W,C], grouped by 32) + bias; w packed [C/64][64][9*64] block-diagonal pairs.
static cudaError_t launch_halo_common(TcContext* ctx, const void* xm, const void* w, const float* bias, float* x, int B, int H, int W,
                                      int C, const int* plan, cudaStream_t st, int dense, float slope, bf16* out, const bf16* res,
                                      int max_ctas = 0, int part = 0, int split_permille = 0);

cudaError_t launch_gconv_halo(TcContext* ctx, const void* xm, const void* w, const float* bias, float* x, int B, int H, int W,
                              int C, const int* plan, cudaStream_t st, int max_ctas, int part, int split_permille) {
  return launch_halo_common(ctx, xm, w, bias, x, B, H, W, C, plan, st, 0, 0.f, nullptr, nullptr, max_ctas, part, split_permille);
}

bool conv64_halo_supported(int C, int N) { return g_gconv_mode != 0 && C == kSlice && N == kSlice; }

// Dense 3x3 convolution with 64 input and 64 output channels (the highest-resolution VAE level, vae.py:57-58):
// out bf16 [B,H,W,64] = act(conv(in) + bias) (+ res); w packed [64][9*64] tap-major -- the layout of one grouped-conv slice.
cudaError_t launch_conv64_halo(TcContext* ctx, const void* in, const void* w, const float* bias, void* out, const void* res, int B,
                               int H, int W, float slope, cudaStream_t st) {
  return launch_halo_common(ctx, in, w, bias, nullptr, B, H, W, kSlice, nullptr, st, 1, slope, static_cast<bf16*>(out),
                            static_cast<const bf16*>(res));
}

static cudaError_t launch_halo_common(TcContext* ctx, const void* xm, const void* w, const float* bias, float* x, int B, int H, int W,
                                      int C, const int* plan, cudaStream_t st, int dense, float slope, bf16* out, const bf16* res,
                                      int max_ctas, int part, int split_permille) {
  GconvGeom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.H = H; g.W = W; g.C = C;
  g.dense = dense; g.slope = slope; g.out = out; g.res = res;
  g.TW = W < 128 ? W : 128;
  g.pitch = g.TW + 2;
  g.TH = (128 - g.TW) / g.pitch + 1;                       // largest TH with (TH-1)*pitch + TW <= 128
  g.TB = 1;
  if (g.TH >= H) {
    g.TH = H;
    const int per_img = (H + 2) * g.pitch, last = (H - 1) * g.pitch + g.TW;   // rows used by one image
    if (g.TW == W) g.TB = (128 - last) / per_img + 1;
    if (g.TB > B) g.TB = B;
  }
  if ((long long)(g.TW + 2) > 256 || (g.TH + 2) > 256 || g.TB > 256) return cudaErrorNotSupported;
  g.w_tiles = (W + g.TW - 1) / g.TW; g.h_tiles = (H + g.TH - 1) / g.TH; g.b_tiles = (B + g.TB - 1) / g.TB;
  const int box_rows = g.TB * (g.TH + 2) * g.pitch;
  g.a_bytes = box_rows * 128;
  int need_rows = 128 + 2 * g.pitch + 2;                   // rows the nine descriptors can touch
  if (need_rows < box_rows) need_rows = box_rows;
  g.stage_bytes = ((need_rows * 128 + 1023) / 1024) * 1024;
  // staged residual update (grouped mode): compact rows x 2 halves, double buffered
  g.out_rows = g.TB * g.TH * g.TW;
  g.out_bytes = 0; g.out_half = 0;
  static const bool stage_out = getenv("LDMB_GCONV_RED") == nullptr;             // debug: LDMB_GCONV_RED=1 = red.global.add epilogue
  if (!dense && stage_out && g.out_rows <= 128 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 && C % 4 == 0)
  {
    g.out_half = ((g.out_rows * 128 + 1023) / 1024) * 1024;
    g.out_bytes = 2 * g.out_half;
  }
  const int avail = 232448 - 1024 - kWBytes - 256 - 2 * g.out_bytes;
  g.stages = avail / g.stage_bytes;
  if (g.stages > 4) g.stages = 4;
  if (g.stages < 1) return cudaErrorNotSupported;

  g.dbg = tc_knobs().gconv_dbg;
  CUtensorMap tmA, tmW;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
    const cuuint32_t box[4] = {kSlice, (cuuint32_t)g.pitch, (cuuint32_t)(g.TH + 2), (cuuint32_t)g.TB};
    if (ctx->encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xm), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)(9 * kSlice), (cuuint64_t)C};
    const cuuint64_t gstr[1] = {(cuuint64_t)(9 * kSlice) * 2};
    const cuuint32_t box[2] = {kSlice, kSlice};
    if (ctx->encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  CUtensorMap tmX;
  memset(&tmX, 0, sizeof(tmX));
  if (g.out_bytes) {
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)C * 4 * W, (cuuint64_t)C * 4 * W * H};
    const cuuint32_t box[4] = {32, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.TB};
    if (ctx->encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const int smem = 1024 + kWBytes + g.stages * g.stage_bytes + 2 * g.out_bytes + 256;
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(gconv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr.mark(ctx->device);
  }
  const int nz = C / kSlice;
  const int n_sp_all = g.b_tiles * g.h_tiles * g.w_tiles;
  // part 1 = the first split_permille / 1000 of the spatial tiles, part 2 = the rest (one conv as two launches: the first on the SMs a
  // concurrent kernel leaves idle, the second on the whole machine)
  const int sp_split = (int)((long long)n_sp_all * split_permille / 1000);
  g.sp_lo = part == 2 ? sp_split : 0;
  g.sp_hi = part == 1 ? sp_split : n_sp_all;
  const int n_sp = g.sp_hi - g.sp_lo;
  if (n_sp <= 0) return cudaSuccess;
  int per_z = ctx->num_sms / nz;
  if (max_ctas > 0 && max_ctas / nz < per_z) per_z = max_ctas / nz;
  if (per_z > n_sp) per_z = n_sp;
  if (per_z < 1) per_z = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(per_z * nz); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_ldmb_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gconv_halo_kernel, tmA, tmW, tmX, x, bias, g, plan, ctx->fault_dev, ctx->trace_dev);
}
