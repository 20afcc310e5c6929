// ChannelNorm + FiLM + grouped 3x3 convolution of one SwinBlock in ONE kernel, for the UNet levels where a whole image
// (plus its zero border) fits one 128-row tcgen05 tile (8x8 and 4x4 feature maps: 24 of the 36 blocks of the default UNet):
//     xm = norm(x) * mul + bias      (modules.py:23-25, unet.py:22)   -> bf16, written for the block's GEMMs
//     x += conv3x3_grouped(xm) + b   (unet.py:30,44)                  -> in place, plain stores
// At these levels the separate norm / conv kernels are pure launch + ramp latency (5 us and 9 us for 8 MB of activations),
// and the conv's CTAs took the SMs of the GEMM it ran beside.  Here:
//   * a CLUSTER of CL CTAs covers the C channels of an image, S = C / (64 CL) 64-channel slices (pairs of conv groups) per CTA,
//     with the slice's 9 tap matrices resident in shared memory (as in kernels_gconv.cu);
//   * the per-pixel statistics need all C channels: every CTA reduces its own slice (mean, M2 over 64 channels, 16 lanes per
//     pixel) and writes the partials into the shared memory of all CTAs of the cluster (DSMEM), one cluster barrier, then
//     each CTA merges the CL S partials (Chan's parallel variance) -- x is read from global exactly once;
//   * the normalised, modulated bf16 values go to global (xm) and, 128B-swizzled, into the zero-bordered halo patch the nine
//     row-shifted UMMA descriptors read (the gconv trick: tap (dy,dx) = row offset dy*(W+2)+dx into one patch);
//   * the epilogue thread of a patch row adds accumulator + bias to x with plain loads / stores: an image's slice of x is
//     read and written by this CTA only, so there is no reduction traffic and no second writer during the kernel.
#include <cuda.h>
#include <math.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tc_context.h"

namespace {

constexpr int kSlice = 64;
constexpr int kWTile = kSlice * 128;       // one tap's [64 co x 64 ci] bf16 tile (block-diagonal pair of groups), 128 B rows
constexpr int kWBytes = 9 * kWTile;        // 72 KB per slice
constexpr int kThreads = 256;
constexpr int kNV = 16;                    // (pixel, slice) units per thread per round: 16 half-warps x 16 = 256 units

struct NcGeom {
  int B, H, W, C, HW;
  int pitch, per_img, TB, n_tiles;         // patch pitch W + 2, patch rows per image, images per MMA tile, tiles = ceil(B / TB)
  int S, CL, n_clusters, R;                // slices per CTA, CTAs per cluster, clusters in the grid, MMA tiles per round
  int NU;                                  // units per round = R * TB * HW * S  (<= 256)
  int patch_bytes, slots;                  // bytes per (tile, slice) patch buffer; pixel slots per round = R * TB * HW
  int s_shift, hw_shift, w_shift;          // log2 of S, HW, W (all powers of two: the unit decode is shifts and masks)
  float eps;
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

__device__ __forceinline__ float half_warp_sum(float v) {      // over the 16 lanes that share a pixel
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
normconv_kernel(const __grid_constant__ CUtensorMap tmW, float* __restrict__ x, const float* __restrict__ film,
                const int* __restrict__ t_index, bf16* __restrict__ xm, const float* __restrict__ bias, const NcGeom g,
                const int* __restrict__ plan, int* fault, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* wts = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* patches = wts + g.S * kWBytes;
  float2* stats = reinterpret_cast<float2*>(patches + g.R * g.S * g.patch_bytes);      // [2][CL * S][slots]
  uint64_t* wbar = reinterpret_cast<uint64_t*>(stats + 2 * g.CL * g.S * g.slots);
  uint64_t* mma_done = wbar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(tmem_slot + 1);
  auto stamp = [&](int slot) {     // debug (ldmb_debug_tc_trace): %globaltimer at the phase boundaries, thread 0 of every CTA
    if (trace != nullptr && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace[blockIdx.x * 16 + slot] = t; }
  };
  stamp(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = g.CL > 1 ? ptx::cluster_ctarank() : 0u;
  const int cl = blockIdx.x / g.CL;

  if (threadIdx.x == 0) {
    ptx::mbar_init(wbar, 1);
    ptx::mbar_init(mma_done, 1);
    *s_abort = 0;
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmW);
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  // the halo patches: borders (and rows the descriptors reach beyond the last image) must read as zero = the conv padding
  for (int i = threadIdx.x; i < g.R * g.S * g.patch_bytes / 16; i += kThreads)
    ptx::st_shared_v4(ptx::smem_u32(patches) + i * 16, 0u, 0u, 0u, 0u);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  stamp(1);
  const uint32_t tmem_base = *tmem_slot;
  const bool skip_block = plan != nullptr && plan[0] != 0;       // stochastic depth (unet.py:39-40)
  // weights are older than the previous kernel: requested before waiting on it
  if (threadIdx.x == 0 && !skip_block) {
    ptx::mbar_arrive_expect_tx(wbar, g.S * kWBytes);
    for (int s = 0; s < g.S; ++s)
      for (int tap = 0; tap < 9; ++tap)
        ptx::tma_load_2d(wts + s * kWBytes + tap * kWTile, &tmW, wbar, tap * kSlice, ((int)rank * g.S + s) * kSlice);
  }
  if (g.CL > 1) ptx::cluster_sync();         // every CTA of the cluster is resident before anyone writes into its smem
  stamp(2);
  pdl_wait();
  stamp(3);

  const int tiles_per_pass = g.n_clusters * g.R;
  const int rounds = skip_block ? 0 : (g.n_tiles + tiles_per_pass - 1) / tiles_per_pass;
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const float inv_c1 = 1.f / (float)(g.C - 1);
  const int n_part = g.CL * g.S;
  bool ok = true;

  for (int r = 0; r < rounds; ++r) {
    const int tile0 = (r * g.n_clusters + cl) * g.R;
    float2* st_cur = stats + (r & 1) * n_part * g.slots;
    // ---------------- phase A: this CTA's slice(s) of x, partial statistics -> every CTA of the cluster
    float4 v[kNV];
    int pix[kNV];               // global pixel index (image * HW + p) of the unit, -1: none
#pragma unroll
    for (int k = 0; k < kNV; ++k) {
      const int u = hw + 16 * k;
      const int s = u & (g.S - 1), pp = u >> g.s_shift, p = pp & (g.HW - 1), ti = pp >> g.hw_shift;
      const int tile = tile0 + ti / g.TB, b = tile * g.TB + ti % g.TB;
      const bool valid = u < g.NU && tile < g.n_tiles && b < g.B;
      pix[k] = valid ? b * g.HW + p : -1;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) v[k] = *reinterpret_cast<const float4*>(x + (long long)pix[k] * g.C + ((int)rank * g.S + s) * kSlice + l16 * 4);
    }
#pragma unroll
    for (int k = 0; k < kNV; ++k) {
      const int u = hw + 16 * k;
      const float mean = half_warp_sum((v[k].x + v[k].y) + (v[k].z + v[k].w)) * (1.f / kSlice);
      const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
      const float m2 = half_warp_sum((dx * dx + dy * dy) + (dz * dz + dw * dw));
      if (u < g.NU && l16 < g.CL) {           // lane j of the half-warp delivers the partial to CTA j
        const int s = u & (g.S - 1), slot = u >> g.s_shift;
        const uint32_t local = ptx::smem_u32(st_cur + ((int)rank * g.S + s) * g.slots + slot);
        st_cluster_f2(g.CL > 1 ? mapa(local, (uint32_t)l16) : local, mean, m2);
      }
    }
    if (r == 0) stamp(4);
    if (g.CL > 1) ptx::cluster_sync(); else __syncthreads();
    if (r == 0) stamp(5);
    // ---------------- phase B: merge the partials, normalise + FiLM, -> xm (global) and the halo patch (smem)
    // (plain C++ loads / stores, no asm barriers: the FiLM loads of later units are free to issue ahead of earlier units' stores)
    float mean_k[kNV], rs_k[kNV];
#pragma unroll
    for (int k = 0; k < kNV; ++k) {
      const int u = hw + 16 * k;
      const int slot = u >> g.s_shift;
      float2 part = make_float2(0.f, 0.f);
      if (u < g.NU && l16 < n_part) part = st_cur[l16 * g.slots + slot];
      const float mean = half_warp_sum(part.x) / (float)n_part;
      const float dm = l16 < n_part ? part.x - mean : 0.f;
      const float m2 = half_warp_sum(part.y + (float)kSlice * dm * dm);       // Chan et al.: M2 = sum M2_i + n_i (mean_i - mean)^2
      mean_k[k] = mean;
      rs_k[k] = 1.f / sqrtf(m2 * inv_c1 + g.eps);                             // unbiased variance (modules.py:24)
    }
#pragma unroll
    for (int k0 = 0; k0 < kNV; k0 += 4) {
      float4 mu[4], bi[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk, u = hw + 16 * k;
        if (pix[k] >= 0) {
          const int s = u & (g.S - 1), p = (u >> g.s_shift) & (g.HW - 1);
          const int b = pix[k] >> g.hw_shift, c0 = ((int)rank * g.S + s) * kSlice + l16 * 4;
          const int trow = t_index != nullptr ? t_index[b] : 0;
          const float* fr = film + ((long long)trow * g.HW + p) * 2 * g.C + c0;
          mu[kk] = __ldg(reinterpret_cast<const float4*>(fr));
          bi[kk] = __ldg(reinterpret_cast<const float4*>(fr + g.C));
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk, u = hw + 16 * k;
        if (pix[k] >= 0) {
          const int s = u & (g.S - 1), slot = u >> g.s_shift, p = slot & (g.HW - 1), ti = slot >> g.hw_shift;
          const int c0 = ((int)rank * g.S + s) * kSlice + l16 * 4;
          const float mean = mean_k[k], rs = rs_k[k];
          const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf((v[k].x - mean) * rs, mu[kk].x, bi[kk].x), fmaf((v[k].y - mean) * rs, mu[kk].y, bi[kk].y));
          const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf((v[k].z - mean) * rs, mu[kk].z, bi[kk].z), fmaf((v[k].w - mean) * rs, mu[kk].w, bi[kk].w));
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(xm + (long long)pix[k] * g.C + c0) = pk;
          // patch row of pixel (y, x) of image ib in its tile: ib * per_img + (y + 1) * pitch + (x + 1); 8 channels = one 16-byte chunk
          const int prow = (ti % g.TB) * g.per_img + ((p >> g.w_shift) + 1) * g.pitch + ((p & (g.W - 1)) + 1);
          uint8_t* dst = patches + ((ti / g.TB) * g.S + s) * g.patch_bytes + prow * 128 + (((l16 >> 1) ^ (prow & 7)) << 4) + (l16 & 1) * 8;
          *reinterpret_cast<uint2*>(dst) = pk;
        }
      }
    }
    if (r == 0) stamp(6);
    ptx::fence_proxy_async();          // generic-proxy smem writes -> the tensor core's async-proxy reads
    ptx::tc_fence_before();
    __syncthreads();
    if (r == 0) stamp(7);
    // ---------------- phase C: the nine taps of every (tile, slice) of the round, one warp issues
    if (warp == 0) {
      ptx::tc_fence_after();
      if (r == 0) ok = wait_bar(wbar, 0, s_abort, fault, 51);
      const bool issuer = ptx::elect_one();
      constexpr uint32_t idesc32 = ptx::idesc_bf16(128, 32);
      const uint32_t pitch8 = (uint32_t)g.pitch * 8u;            // one patch row = 128 B = 8 descriptor units
      for (int j = 0; ok && j < g.R; ++j) {
        if (tile0 + j >= g.n_tiles) break;
        for (int s = 0; s < g.S; ++s) {
          const uint64_t a_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(patches + (j * g.S + s) * g.patch_bytes));
          const uint64_t w_desc0 = ptx::smem_desc_sw128(ptx::smem_u32(wts + s * kWBytes));
          const uint32_t d_tmem = tmem_base + (j * g.S + s) * kSlice;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t a_tap = a_desc0 + (uint64_t)((tap / 3) * pitch8 + (tap % 3) * 8u);
            const uint64_t w_tap = w_desc0 + (uint64_t)(tap * (kWTile / 16));
            // the slice's [64 out x 64 in] tap matrix is block-diagonal (two groups of 32 channels): k-steps 0-1 feed output
            // columns 0..31, k-steps 2-3 columns 32..63 -- N = 32 MMAs on the matching 32 weight rows (kernels_gconv.cu)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t grp = k >> 1;
              if (issuer) ptx::umma_f16(d_tmem + grp * 32, a_tap + 2 * k, w_tap + grp * (32 * 128 / 16) + 2 * k, idesc32, (tap | (k & 1)) != 0 ? 1u : 0u);
            }
          }
        }
      }
      if (issuer) ptx::umma_commit(mma_done);
      __syncwarp();
    }
    // ---------------- phase D: x += accumulator + bias for the valid patch rows (plain stores: single owner)
    if (!wait_bar(mma_done, r & 1, s_abort, fault, 52)) break;
    if (r == 0) stamp(8);
    ptx::tc_fence_after();
    {
      const int q = warp & 3, half = warp >> 2;                  // TMEM lane quadrant, 32-column half of the slice
      const int i = q * 32 + lane;                               // patch row
      const int ib = i / g.per_img, rr = i % g.per_img, yy = rr / g.pitch, xx = rr % g.pitch;
      const bool row_ok = ib < g.TB && yy < g.H && xx < g.W;
      // two (tile, slice) accumulators at a time: both rows of x are requested before either accumulator is read
      const int n_acc = g.R * g.S;
      for (int a0 = 0; a0 < n_acc; a0 += 2) {
        float4 cur[2][8];
        float* xr[2];
        bool valid[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = a0 + e, j = a >> g.s_shift, s = a & (g.S - 1);
          const int tile = tile0 + j, b = tile * g.TB + ib;
          valid[e] = a < n_acc && row_ok && tile < g.n_tiles && b < g.B;
          const int c0 = ((int)rank * g.S + s) * kSlice + half * 32;
          xr[e] = x + ((long long)(b * g.HW + yy * g.W + xx)) * g.C + c0;
          if (valid[e]) {
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[e][u] = *reinterpret_cast<const float4*>(xr[e] + 4 * u);
          }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = a0 + e, s = a & (g.S - 1);
          if (a >= n_acc) break;                                   // uniform
          uint32_t acc[32];
          ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * kSlice + half * 32, acc);
          ptx::tmem_ld_wait();
          if (valid[e]) {
            const float* bz = bias + ((int)rank * g.S + s) * kSlice + half * 32;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bz + 4 * u));
              float4 c4 = cur[e][u];
              c4.x += __uint_as_float(acc[4 * u]) + bb.x; c4.y += __uint_as_float(acc[4 * u + 1]) + bb.y;
              c4.z += __uint_as_float(acc[4 * u + 2]) + bb.z; c4.w += __uint_as_float(acc[4 * u + 3]) + bb.w;
              *reinterpret_cast<float4*>(xr[e] + 4 * u) = c4;
            }
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();               // TMEM and the patches are free for the next round
    ptx::tc_fence_after();
    if (r == 0) stamp(9);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (g.CL > 1) ptx::cluster_sync();         // no CTA leaves while a peer may still write statistics into its smem
  stamp(10);
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

bool plan_geom(int num_sms, int B, int H, int W, int C, NcGeom& g, int& smem) {
  memset(&g, 0, sizeof(g));
  g.B = B; g.H = H; g.W = W; g.C = C; g.HW = H * W;
  if (C % kSlice) return false;
  auto log2i = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
  g.hw_shift = log2i(H * W); g.w_shift = log2i(W);
  if (g.hw_shift < 0 || g.w_shift < 0) return false;        // power-of-two feature maps (8x8, 4x4, ...)
  const int nsl = C / kSlice;
  g.CL = nsl >= 8 ? 8 : nsl;
  if (g.CL != 1 && g.CL != 2 && g.CL != 4 && g.CL != 8) return false;
  if (nsl % g.CL) return false;
  g.S = nsl / g.CL;
  if (g.S > 2) return false;                                // 72 KB of resident tap matrices per slice
  g.s_shift = g.S == 2 ? 1 : 0;
  g.pitch = W + 2;
  g.per_img = (H + 2) * g.pitch;
  const int last = (H - 1) * g.pitch + W;                   // MMA rows one image needs (its valid output rows end here)
  if (last > 128) return false;                             // the image must fit one 128-row tile
  g.TB = (128 - last) / g.per_img + 1;
  if (g.TB > B) g.TB = B;
  g.n_tiles = (B + g.TB - 1) / g.TB;
  g.n_clusters = num_sms / g.CL;
  if (g.n_clusters < 1) return false;
  // rows the nine descriptors can touch: 128 + the largest tap offset; at least every image's full patch
  int rows = 128 + 2 * g.pitch + 2;
  if (rows < g.TB * g.per_img) rows = g.TB * g.per_img;
  g.patch_bytes = ((rows * 128 + 1023) / 1024) * 1024;
  // tiles per round: TMEM (64 columns per tile and slice), 256 (pixel, slice) units, shared memory
  for (g.R = 8 / g.S > 4 ? 4 : 8 / g.S; g.R >= 1; --g.R) {
    g.NU = g.R * g.TB * g.HW * g.S;
    g.slots = g.R * g.TB * g.HW;
    smem = 1024 + g.S * kWBytes + g.R * g.S * g.patch_bytes + 2 * g.CL * g.S * g.slots * 8 + 64;
    if (g.NU <= 16 * kNV && smem <= 232448) break;
  }
  if (g.R < 1) return false;
  const int need = (g.n_tiles + g.R - 1) / g.R;             // clusters that have work
  if (g.n_clusters > need) g.n_clusters = need;
  return true;
}

}  // namespace

// Off in the UNet step by default (LDMB_NORMCONV=1 turns it on): as it stands one round of the kernel takes ~28 us at the level-2
// shape -- its phases run back to back on 8 warps with 255 registers each, nothing hides their latency -- against 5 us + 9 us for
// the separate kernels.  The kernel-level entry point (ldmb_normconv) and its parity tests always run it.
static const bool g_normconv_in_step = getenv("LDMB_NORMCONV") != nullptr && atoi(getenv("LDMB_NORMCONV")) != 0;

bool normconv_in_step() { return g_normconv_in_step; }

bool normconv_supported(int B, int H, int W, int C) {
  NcGeom g;
  int smem;
  return B >= 1 && H >= 1 && W >= 1 && (long long)B * H * W < (1LL << 30) && plan_geom(148, B, H, W, C, g, smem);
}

// xm bf16 [B,H,W,C] = FiLM(ChannelNorm(x)); x fp32 [B,H,W,C] += conv3x3(xm, groups of 32) + bias.
// film fp32 [n_t][HW][2C] (mul | bias), row of image b = t_index[b] (NULL: 0); w packed [C/64][64][9*64] block-diagonal pairs.
cudaError_t launch_normconv(TcContext* ctx, float* x, const float* film, const int* t_index, void* xm, const void* w, const float* bias,
                            int B, int H, int W, int C, float eps, const int* plan, cudaStream_t st) {
  NcGeom g;
  int smem;
  if (!plan_geom(ctx->num_sms, B, H, W, C, g, smem)) return cudaErrorNotSupported;
  g.eps = eps;
  CUtensorMap tmW;
  {
    const cuuint32_t ones[2] = {1, 1};
    const cuuint64_t gdim[2] = {(cuuint64_t)(9 * kSlice), (cuuint64_t)C};
    const cuuint64_t gstr[1] = {(cuuint64_t)(9 * kSlice) * 2};
    const cuuint32_t box[2] = {kSlice, kSlice};
    if (ctx->encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(normconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr.mark(ctx->device);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g.n_clusters * g.CL); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (g.CL > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = g.CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
  cfg.attrs = at; cfg.numAttrs = na;
  // Clusters beyond what the GPU holds at once would run as a second wave of the whole kernel: cap the grid at the co-resident
  // count (a cluster needs CL free SMs in ONE GPC) and let the rounds loop cover the rest.  Cached per (cluster size, smem).
  static int max_active[16][4] = {};
  const int key = g.CL == 8 ? 3 : (g.CL == 4 ? 2 : (g.CL == 2 ? 1 : 0));
  int& cached = max_active[ctx->device & 15][key];
  if (cached == 0) {
    int n = 0;
    cfg.dynamicSmemBytes = 232448;
    if (g.CL > 1 && cudaOccupancyMaxActiveClusters(&n, normconv_kernel, &cfg) == cudaSuccess && n > 0) cached = n;
    else cached = ctx->num_sms / g.CL;
    cfg.dynamicSmemBytes = smem;
  }
  if (g.n_clusters > cached) {
    g.n_clusters = cached;
    cfg.gridDim = dim3(g.n_clusters * g.CL);
  }
  if (g_ldmb_pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; cfg.numAttrs = na; }
  return cudaLaunchKernelEx(&cfg, normconv_kernel, tmW, x, film, t_index, static_cast<bf16*>(xm), bias, g, plan, ctx->fault_dev, ctx->trace_dev);
}
