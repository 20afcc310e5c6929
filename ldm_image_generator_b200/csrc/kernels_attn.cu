// Window attention core on the warp-level tensor path (bf16 mma.sync m16n8k16, fp32 softmax), sm_100a.
//
// attention.py:13-85 + torch MHA as called there: per (image, window, head) S = (q/sqrt(32)) k^T + key bias,
// softmax over the keys, O = P v.  L = window tokens (36 for the 6x6 windows, 16 for the global 4x4 case) and
// d = 32, so one (window, head) problem is 2 x 36x36x32 MACs: far too small for a tcgen05 tile (M = 128), and
// at 0.3 % of the step's FLOPs the kernel is bound by moving qkv (3 M C bf16 in, M C bf16 out), not by math.
// One CTA = one (image, window) x 4 heads, one warp per head:
//   * q/k/v rows of the window are gathered with 16-byte loads into padded shared memory (index arithmetic
//     replaces the reference's pad / roll / split / concat copies, see window_attention_kernel in kernels_simt.cu)
//   * S: ldmatrix A (q) / B (k) fragments, MT x 2MT x 2 mma; scale, key bias (-inf mask or float bias), row softmax
//     with quad shuffles; P stays in registers and is re-used as the A fragment of the second product
//   * O: ldmatrix.trans B (v) fragments, MT x 4 x MT mma; normalised, staged through shared memory and written
//     with 16-byte coalesced stores (pad tokens are cropped, attention.py:56).
#include "kernels.h"

#include <math.h>

namespace {

constexpr int kD = 32;          // head_dim (unet.py:26)
constexpr int kHC = 4;          // heads per CTA = warps per CTA
constexpr int kRow = kHC * kD;  // channels per CTA
constexpr int kLd = kRow + 8;   // padded smem row (bf16 elements): 272 B, conflict-free for ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void ptx_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool ptx_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(parity) : "memory");
  return ok != 0;
}
// 1-d bulk async copy global -> shared (TMA), completion bytes on an mbarrier; 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void ptx_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src), "r"(bytes),
                 "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))) : "memory");
}

// MT = ceil(L / 16) query/key tiles of 16 (L <= 16 * MT)
template <int MT>
__global__ void __launch_bounds__(32 * kHC) window_attention_mma_kernel(
    const bf16* __restrict__ qkv, const bf16* __restrict__ xm, const float* __restrict__ b_in, bf16* __restrict__ att,
    long long ldo, int H, int W, int C, int wh, int ww, int shift, int Hp, int Wp, const int* __restrict__ skip) {
  pdl_wait();
  if (skip != nullptr && *skip != 0) return;
  constexpr int LP = 16 * MT, NT = 2 * MT;
  __shared__ __align__(16) bf16 Qs[LP * kLd];
  __shared__ __align__(16) bf16 Ks[LP * kLd];
  __shared__ __align__(16) bf16 Vs[LP * kLd];
  __shared__ float kb[LP];
  __shared__ long long rowm[LP];
  __shared__ __align__(16) bf16 padkv[2][kRow];   // k / v row of a pad token (x = 0 -> in_proj bias), this head chunk
  __shared__ __align__(8) uint64_t gbar;
  const int L = wh * ww;
  const int nww = Wp / ww, nwin = (Hp / wh) * nww;
  const int b = blockIdx.x / nwin, win = blockIdx.x % nwin;
  const int wi = win / nww, wj = win % nww;
  const int ch0 = blockIdx.y * kRow;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&gbar))) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < LP) {
    long long rm = -2;          // -2: tile padding (no token), -1: window slot outside the image (pad token)
    float bias = -INFINITY;
    if (tid < L) {
      const int si = tid / ww, sj = tid % ww;
      const int pi = (wi * wh + si - shift + Hp) % Hp, pj = (wj * ww + sj - shift + Wp) % Wp;
      const bool pad = pi >= H || pj >= W;
      rm = pad ? -1 : ((long long)b * H + pi) * W + pj;
      if (shift == 0) {
        bias = pad ? -INFINITY : 0.f;                       // bool key_padding_mask (attention.py:27-35)
      } else {                                              // float "mask" = rolled activation, channel 0 (attention.py:40)
        const int qi = (pi - shift + Hp) % Hp, qj = (pj - shift + Wp) % Wp;
        bias = (qi < H && qj < W) ? __bfloat162float(xm[(((long long)b * H + qi) * W + qj) * C]) : 0.f;
      }
    }
    rowm[tid] = rm;
    kb[tid] = bias;
  }
  for (int i = tid; i < 2 * kRow; i += 32 * kHC)            // once per CTA (the per-piece conversion in the fill loop below
    padkv[i / kRow][i % kRow] = __float2bfloat16_rn(b_in[(1 + i / kRow) * C + ch0 + i % kRow]);   // was 42 % of the stall samples at level 2)
  __syncthreads();
  // ---- gather q | k | v rows of this (window, head chunk): one 256-byte bulk async copy (TMA, 1-d) per (token, matrix)
  //      issued by one thread each and tracked by an mbarrier -- a handful of instructions instead of a 14-iteration
  //      per-thread loop of 16-byte copies (the kernel was instruction-bound in that loop).
  {
    const int n_real = __syncthreads_count(tid < LP && rowm[tid < LP ? tid : 0] >= 0);
    if (tid == 0) ptx_mbar_expect_tx(&gbar, static_cast<uint32_t>(n_real) * 3u * (kRow * 2));
    for (int it = tid; it < 3 * LP; it += 32 * kHC) {
      const int tok = it / 3, which = it % 3;
      const long long m = rowm[tok];
      bf16* dst = (which == 0 ? Qs : (which == 1 ? Ks : Vs)) + tok * kLd;
      if (m >= 0) {
        ptx_bulk_g2s(dst, qkv + m * 3 * C + (long long)which * C + ch0, kRow * 2, &gbar);
      } else {
        const bool bias_fill = m == -1 && which != 0;         // pad token: x = 0 -> k, v = in_proj bias (attention.py:19-23)
        const uint4* src = reinterpret_cast<const uint4*>(padkv[which == 2 ? 1 : 0]);
#pragma unroll
        for (int piece = 0; piece < kRow / 8; ++piece)
          *reinterpret_cast<uint4*>(dst + piece * 8) = bias_fill ? src[piece] : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    while (!ptx_mbar_try_wait(&gbar, 0)) {}
    __syncthreads();
  }

  const int g = lane >> 2, t4 = lane & 3;
  const int hc = warp * kD;                                 // this warp's head: columns [hc, hc + 32)
  // One 16-query tile at a time (K / V fragments are re-read from shared memory per tile): keeps the live state to
  // one S row-block + one O row-block so several CTAs fit an SM's register file.
  const float scale_l2 = 0.17677669529663687f * 1.4426950408889634f;   // 1/sqrt(32) * log2(e)
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    // a tile whose 16 query slots are all padding (windows on the padded border, tile padding): its outputs would be
    // cropped (attention.py:56) -- nothing to compute.  Same decision in every warp of the CTA.
    if (!__any_sync(0xffffffffu, lane < 16 && rowm[mt * 16 + (lane & 15)] >= 0)) continue;
    // ---- S = q k^T
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = hc + ks * 16 + (lane >> 4) * 8;
      ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(Qs + row * kLd + col)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t kf[4];
      const int row = nt * 8 + (lane & 7), col = hc + (lane >> 3) * 8;
      ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(Ks + row * kLd + col)), kf[0], kf[1], kf[2], kf[3]);
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      mma_bf16(s[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kf[0], kf[1]);
      mma_bf16(s[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kf[2], kf[3]);
    }
    // ---- softmax over keys (rows g and g+8 of the tile; a row is spread over the 4 lanes of a quad)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float b0 = kb[nt * 8 + 2 * t4] * 1.4426950408889634f, b1 = kb[nt * 8 + 2 * t4 + 1] * 1.4426950408889634f;
      s[nt][0] = fmaf(s[nt][0], scale_l2, b0); s[nt][1] = fmaf(s[nt][1], scale_l2, b1);
      s[nt][2] = fmaf(s[nt][2], scale_l2, b0); s[nt][3] = fmaf(s[nt][3], scale_l2, b1);
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mx0); s[nt][1] = exp2f(s[nt][1] - mx0);
      s[nt][2] = exp2f(s[nt][2] - mx1); s[nt][3] = exp2f(s[nt][3] - mx1);
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    // ---- O = P v (P re-used from the S accumulators as A fragments)
    float o[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < MT; ++kt) {
      const uint32_t p0 = pack2(s[2 * kt][0], s[2 * kt][1]), p1 = pack2(s[2 * kt][2], s[2 * kt][3]);
      const uint32_t p2 = pack2(s[2 * kt + 1][0], s[2 * kt + 1][1]), p3 = pack2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {                       // pairs of 8-wide d tiles
        uint32_t vf[4];
        const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = hc + (dp * 2 + (lane >> 4)) * 8;
        ldsm_x4_t(static_cast<uint32_t>(__cvta_generic_to_shared(Vs + row * kLd + col)), vf[0], vf[1], vf[2], vf[3]);
        mma_bf16(o[dp * 2], p0, p1, p2, p3, vf[0], vf[1]);
        mma_bf16(o[dp * 2 + 1], p0, p1, p2, p3, vf[2], vf[3]);
      }
    }
    // ---- stage this head's output over its own (consumed) q rows of this tile
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      bf16* q0 = Qs + (mt * 16 + g) * kLd + hc + dt * 8 + 2 * t4;
      *reinterpret_cast<uint32_t*>(q0) = pack2(o[dt][0] * inv0, o[dt][1] * inv0);
      *reinterpret_cast<uint32_t*>(q0 + 8 * kLd) = pack2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < L * (kRow / 8); idx += 32 * kHC) {
    const int piece = idx % (kRow / 8), tok = idx / (kRow / 8);
    const long long m = rowm[tok];
    if (m < 0) continue;                                    // outputs at pad positions are cropped (attention.py:56)
    *reinterpret_cast<uint4*>(att + m * ldo + ch0 + piece * 8) = *reinterpret_cast<const uint4*>(Qs + tok * kLd + piece * 8);
  }
}

}  // namespace

bool window_attention_mma_supported(int C, int head_dim, int win_h, int win_w, long long ldo) {
  return head_dim == kD && (C / kD) % kHC == 0 && C % kRow == 0 && win_h * win_w <= 48 && ldo % 8 == 0 && C % 8 == 0;
}

cudaError_t launch_window_attention_mma(const void* qkv, const void* xm, const float* b_in, void* att, long long ldo, int B,
                                        int H, int W, int C, int win_h, int win_w, int shift, const int* skip,
                                        cudaStream_t st) {
  const int Hp = (H + win_h - 1) / win_h * win_h, Wp = (W + win_w - 1) / win_w * win_w;
  const int L = win_h * win_w;
  dim3 grid(B * (Hp / win_h) * (Wp / win_w), C / kRow);
  if (grid.y > 65535u) return cudaErrorNotSupported;
  const bf16* q = static_cast<const bf16*>(qkv);
  const bf16* x = static_cast<const bf16*>(xm);
  bf16* o = static_cast<bf16*>(att);
  if (L <= 16) launch_k((window_attention_mma_kernel<1>), grid, 32 * kHC, 0, st, q, x, b_in, o, ldo, H, W, C, win_h, win_w, shift, Hp, Wp, skip);
  else if (L <= 32) launch_k((window_attention_mma_kernel<2>), grid, 32 * kHC, 0, st, q, x, b_in, o, ldo, H, W, C, win_h, win_w, shift, Hp, Wp, skip);
  else launch_k((window_attention_mma_kernel<3>), grid, 32 * kHC, 0, st, q, x, b_in, o, ldo, H, W, C, win_h, win_w, shift, Hp, Wp, skip);
  return cudaGetLastError();
}
