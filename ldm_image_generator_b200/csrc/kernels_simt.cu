// CUDA-core kernels of libldmb200: the fp32 validation GEMM/convolution, and every
// HBM-bound (norm / elementwise / gather) kernel of the sampling path.
// Reference lines cited are under /root/reference.
#include "kernels.h"

#include <math.h>
#include <stdlib.h>

bool g_ldmb_pdl = getenv("LDMB_NO_PDL") == nullptr;

namespace {

// =====================================================================================
// Generic GEMM / implicit-GEMM convolution on CUDA cores (64x64x16 tiles, 4x4 per thread).
// The LDMB_FP32_VALIDATE implementation of every contraction, and the bf16 fallback for
// toy channel counts that cannot be tiled for tcgen05.
// =====================================================================================
template <typename T, int AMODE, bool GLU>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDesc dparam) {
  pdl_wait();
  constexpr int BM = 64, BN = 64, BK = 16, NW = GLU ? 2 : 1;
  GemmDesc d = dparam;
  if (resolve_plan(d)) return;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[NW][BK][BN + 4];
  const int z = blockIdx.z;
  const int m0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int NL = GLU ? d.N / 2 : d.N;          // logical output columns
  const int G = d.glu_chunk;
  const T* __restrict__ A = reinterpret_cast<const T*>(d.A);
  const T* __restrict__ W = reinterpret_cast<const T*>(d.W);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[NW][4][4];
#pragma unroll
  for (int w = 0; w < NW; ++w)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[w][i][j] = 0.f;

  for (int k0 = 0; k0 < d.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256, r = idx >> 4, kk = idx & 15;
      const int m = m0 + r, k = k0 + kk;
      float v = 0.f;
      if (m < d.M && k < d.K) {
        if (AMODE == AM_ROWS) {
          v = to_f(A[(long long)m * d.lda + z * d.a_koff_b + k]);
        } else {
          const int tap = k / d.cC, c = k % d.cC;
          const int hw = d.cH * d.cW, b = m / hw, rem = m % hw;
          const int hh = rem / d.cW + tap / 3 - 1, ww = rem % d.cW + tap % 3 - 1;
          if (hh >= 0 && hh < d.cH && ww >= 0 && ww < d.cW)
            v = to_f(A[(((long long)b * d.cH + hh) * d.cW + ww) * d.lda + z * d.a_koff_b + c]);
        }
      }
      As[kk][r] = v;
    }
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256, c = idx >> 4, kk = idx & 15;
        const int j = j0 + c, k = k0 + kk;
        float v = 0.f;
        if (j < NL && k < d.K) {
          const int n = GLU ? (j / G) * 2 * G + j % G + w * G : j;
          long long row; int col = k;
          if (d.sel == 2) { row = d.sel_rows[k / d.sel_span] + n + z * d.w_row_b; col = k % d.sel_span; }
          else if (d.sel == 3) { row = (long long)(k / d.sel_span) * d.sel_stride + n + z * d.w_row_b; col = k % d.sel_span; }
          else row = wrow_of_col(d, z, n);
          v = to_f(W[row * d.ldw + col]);
        }
        Ws[w][kk][c] = v;
      }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[NW][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int j = 0; j < 4; ++j) b[w][j] = Ws[w][kk][tx * 4 + j];
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[w][i][j] = fmaf(a[i], b[w][j], acc[w][i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= d.M) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + tx * 4 + jj;
      if (j >= NL) continue;
      if (GLU) {
        const int na = (j / G) * 2 * G + j % G, nb = na + G;
        const float va = acc[0][i][jj] + bias_of_col(d, z, na);
        const float vb = acc[NW - 1][i][jj] + bias_of_col(d, z, nb);
        const bool keep = d.mask_plan == nullptr || expert_kept(d, m, j);
        reinterpret_cast<T*>(d.out)[(long long)m * d.ldo + z * d.out_off_b + j] = from_f<T>(keep ? va * fmaxf(vb, 0.f) : 0.f);
      } else {
        float v = acc[0][i][jj] + bias_of_col(d, z, j);
        if (d.epi == EPI_ACCUM_F32) {
          float* o = reinterpret_cast<float*>(d.out) + (long long)m * d.ldo + z * d.out_off_b + j;
          if (d.res) v += to_f(reinterpret_cast<const T*>(d.res)[(long long)m * d.ldr + j]);
          *o += v;
        } else if (d.epi == EPI_STORE_F32) {
          reinterpret_cast<float*>(d.out)[(long long)m * d.ldo + z * d.out_off_b + j] = apply_act(v, d.act, d.slope);
        } else if (d.epi == EPI_UPADD) {
          const int HW = d.ctH * d.ctW, b = m / HW, r = m % HW, hh = r / d.ctW, ww = r % d.ctW;
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<float*>(d.out)[(((long long)b * 2 * d.ctH + 2 * hh + (q >> 1)) * 2 * d.ctW + 2 * ww + (q & 1)) * d.ldo + j] += v;
        } else if (d.epi == EPI_CONVT) {
          reinterpret_cast<T*>(d.out)[convt_offset(d, m, j)] = from_f<T>(apply_act(v, d.act, d.slope));
        } else {
          v = apply_act(v, d.act, d.slope);
          if (d.res) v += to_f(reinterpret_cast<const T*>(d.res)[(long long)m * d.ldr + j]);
          reinterpret_cast<T*>(d.out)[(long long)m * d.ldo + z * d.out_off_b + j] = from_f<T>(v);
        }
      }
    }
  }
}

template <typename T>
cudaError_t gemm_simt_dispatch(const GemmDesc& d, cudaStream_t s) {
  const bool glu = d.epi == EPI_REGLU;
  const int NL = glu ? d.N / 2 : d.N;
  dim3 grid((NL + 63) / 64, (d.M + 63) / 64, d.batch > 0 ? d.batch : 1);
  if (grid.y > 65535u) return cudaErrorInvalidConfiguration;
  if (d.amode == AM_ROWS) {
    if (glu) launch_k((gemm_simt_kernel<T, AM_ROWS, true>), grid, 256, 0, s, d);
    else launch_k((gemm_simt_kernel<T, AM_ROWS, false>), grid, 256, 0, s, d);
  } else {
    if (glu) return cudaErrorNotSupported;
    launch_k((gemm_simt_kernel<T, AM_CONV3, false>), grid, 256, 0, s, d);
  }
  return cudaGetLastError();
}

// =====================================================================================
// weight repack
// =====================================================================================
struct Repack4 { int dims[4]; long long ss[4]; long long ds[4]; };
template <typename T>
__global__ void repack_kernel(const float* __restrict__ src, T* __restrict__ dst, Repack4 r, long long total) {
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int i3 = t % r.dims[3]; t /= r.dims[3];
    const int i2 = t % r.dims[2]; t /= r.dims[2];
    const int i1 = t % r.dims[1]; t /= r.dims[1];
    const int i0 = (int)t;
    dst[i0 * r.ds[0] + i1 * r.ds[1] + i2 * r.ds[2] + i3 * r.ds[3]] =
        from_f<T>(src[i0 * r.ss[0] + i1 * r.ss[1] + i2 * r.ss[2] + i3 * r.ss[3]]);
  }
}

// =====================================================================================
// ChannelNorm + FiLM: one warp per pixel, the row lives in registers (C <= 2048), 16-byte loads.
// Algorithmic bytes per pixel: 4C read (fp32 residual) + sizeof(T)*C written (+ the L2-resident FiLM row).
// =====================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> struct Pack4;
template <> struct Pack4<float> {
  static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  }
};
template <> struct Pack4<bf16> {
  static __device__ __forceinline__ void store(bf16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// One warp normalises the same pixel p of ROWS consecutive images: the FiLM row (mul | bias, fp32, 8C bytes -- twice
// the bytes of the x row itself) depends on (t, p) only, so it is loaded once and re-used for the ROWS images when
// they share a timestep (always, in DDPM.sample).  The ROWS x rows are independent 16-byte loads in flight.
template <typename T, int MAXV, int ROWS>
__global__ void __launch_bounds__(256) norm_film_kernel(const float* __restrict__ x, const float* __restrict__ film,
                                                        const int* __restrict__ t_index, T* __restrict__ out,
                                                        int M, int C, int HW, float eps, const int* __restrict__ skip) {
  pdl_wait();
  if (skip != nullptr && *skip != 0) return;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const int B = M / HW;
  const int groups = (B + ROWS - 1) / ROWS;
  const float inv_c = 1.f / (float)C, inv_c1 = 1.f / (float)(C - 1);
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < groups * HW; item += warps_per_grid) {
    const int p = item % HW, b0 = (item / HW) * ROWS;
    float4 v[ROWS][MAXV];
    float sum[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      sum[r] = 0.f;
      const float* xr = x + ((long long)(b0 + r) * HW + p) * C;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = i * 128 + lane * 4;
        if (b0 + r < B && c < C) {
          v[r][i] = __ldg(reinterpret_cast<const float4*>(xr + c));
          sum[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
        }
      }
    }
    float4 mu[MAXV], bi[MAXV];
    int t_cur = -1;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (b0 + r >= B) break;                          // warp-uniform
      const int ti = t_index ? t_index[b0 + r] : 0;
      if (ti != t_cur) {                               // warp-uniform
        t_cur = ti;
        const float* fr = film + ((long long)ti * HW + p) * 2 * C;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
          const int c = i * 128 + lane * 4;
          if (c < C) {
            mu[i] = __ldg(reinterpret_cast<const float4*>(fr + c));
            bi[i] = __ldg(reinterpret_cast<const float4*>(fr + C + c));
          }
        }
      }
      const float mean = warp_sum(sum[r]) * inv_c;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < C) {
          v[r][i].x -= mean; v[r][i].y -= mean; v[r][i].z -= mean; v[r][i].w -= mean;
          sq += (v[r][i].x * v[r][i].x + v[r][i].y * v[r][i].y) + (v[r][i].z * v[r][i].z + v[r][i].w * v[r][i].w);
        }
      }
      const float rs = 1.f / sqrtf(warp_sum(sq) * inv_c1 + eps);   // unbiased variance (modules.py:24)
      T* orow = out + ((long long)(b0 + r) * HW + p) * C;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = i * 128 + lane * 4;
        if (c < C)
          Pack4<T>::store(orow + c, fmaf(v[r][i].x * rs, mu[i].x, bi[i].x), fmaf(v[r][i].y * rs, mu[i].y, bi[i].y),
                          fmaf(v[r][i].z * rs, mu[i].z, bi[i].z), fmaf(v[r][i].w * rs, mu[i].w, bi[i].w));
      }
    }
  }
}

// C = 128 variant: 16 lanes per row (8 channels = two float4 per lane), so one warp normalises TWO pixels x ROWS images
// per pass and every instruction (index math, shuffles, FiLM arithmetic) serves two rows -- the 32-lanes-per-row
// kernel above is issue-bound at this width (143 warp instructions per 128-element row, ncu).
template <typename T, int ROWS>
__global__ void __launch_bounds__(256) norm_film_c128_kernel(const float* __restrict__ x, const float* __restrict__ film,
                                                             const int* __restrict__ t_index, T* __restrict__ out,
                                                             int M, int HW, float eps, const int* __restrict__ skip) {
  pdl_wait();
  if (skip != nullptr && *skip != 0) return;
  constexpr int C = 128;
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const int B = M / HW, groups = (B + ROWS - 1) / ROWS, pairs = (HW + 1) / 2;
  for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < groups * pairs; item += warps_per_grid) {
    const int p = (item % pairs) * 2 + half, b0 = (item / pairs) * ROWS;
    const bool pix_ok = p < HW;
    float4 v[ROWS][2];
    float sum[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      sum[r] = 0.f;
      if (pix_ok && b0 + r < B) {
        const float* xr = x + ((long long)(b0 + r) * HW + p) * C + hl * 8;
        v[r][0] = __ldg(reinterpret_cast<const float4*>(xr));
        v[r][1] = __ldg(reinterpret_cast<const float4*>(xr + 4));
        sum[r] = ((v[r][0].x + v[r][0].y) + (v[r][0].z + v[r][0].w)) + ((v[r][1].x + v[r][1].y) + (v[r][1].z + v[r][1].w));
      } else {
        v[r][0] = v[r][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float4 mu[2], bi[2];
    int t_cur = -1;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (b0 + r >= B) break;                          // warp-uniform
      const int ti = t_index ? t_index[b0 + r] : 0;
      if (ti != t_cur && pix_ok) {
        const float* fr = film + ((long long)ti * HW + p) * 2 * C + hl * 8;
        mu[0] = __ldg(reinterpret_cast<const float4*>(fr)); mu[1] = __ldg(reinterpret_cast<const float4*>(fr + 4));
        bi[0] = __ldg(reinterpret_cast<const float4*>(fr + C)); bi[1] = __ldg(reinterpret_cast<const float4*>(fr + C + 4));
      }
      t_cur = ti;
      float s1 = sum[r];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);     // within the 16-lane half
      const float mean = s1 * (1.f / C);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        v[r][i].x -= mean; v[r][i].y -= mean; v[r][i].z -= mean; v[r][i].w -= mean;
        sq += (v[r][i].x * v[r][i].x + v[r][i].y * v[r][i].y) + (v[r][i].z * v[r][i].z + v[r][i].w * v[r][i].w);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rs = 1.f / sqrtf(sq * (1.f / (C - 1)) + eps);      // unbiased variance (modules.py:24)
      if (pix_ok) {
        T* orow = out + ((long long)(b0 + r) * HW + p) * C + hl * 8;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          Pack4<T>::store(orow + 4 * i, fmaf(v[r][i].x * rs, mu[i].x, bi[i].x), fmaf(v[r][i].y * rs, mu[i].y, bi[i].y),
                          fmaf(v[r][i].z * rs, mu[i].z, bi[i].z), fmaf(v[r][i].w * rs, mu[i].w, bi[i].w));
      }
    }
  }
}

template <typename T>
__global__ void emb_build_kernel(const float* __restrict__ pe, const float* __restrict__ te, T* __restrict__ emb,
                                 int n_t, int HW, int C) {
  pdl_wait();
  const long long total = (long long)n_t * HW * 2 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (2 * C));
    const long long row = i / (2 * C);
    const int p = (int)(row % HW), ti = (int)(row / HW);
    emb[i] = from_f<T>(c < C ? pe[(long long)p * C + c] : te[(long long)ti * C + (c - C)]);
  }
}

template <typename TI, typename TO>
__global__ void pool2_kernel(const TI* __restrict__ x, TO* __restrict__ out, int B, int H, int W, int C) {
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const TI* p = x + (((long long)b * H + 2 * ho) * W + 2 * wo) * C + c;
    const float s = (to_f(p[0]) + to_f(p[C])) + (to_f(p[(long long)W * C]) + to_f(p[(long long)W * C + C]));
    out[i] = from_f<TO>(s * 0.25f);
  }
}

// AvgPool2d(2) fp32 NHWC -> T, four channels per thread (16-byte loads)
template <typename TO>
__global__ void pool2_vec4_kernel(const float* __restrict__ x, TO* __restrict__ out, int B, int H, int W, int C4) {
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * C4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const float4* p = x4 + (((long long)b * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const float4 a = __ldg(p), bq = __ldg(p + C4), cq = __ldg(p + (long long)W * C4), d = __ldg(p + (long long)W * C4 + C4);
    Pack4<TO>::store(out + i * 4, ((a.x + bq.x) + (cq.x + d.x)) * 0.25f, ((a.y + bq.y) + (cq.y + d.y)) * 0.25f,
                     ((a.z + bq.z) + (cq.z + d.z)) * 0.25f, ((a.w + bq.w) + (cq.w + d.w)) * 0.25f);
  }
}

template <typename T>
__global__ void cast_vec4_kernel(const float* __restrict__ x, T* __restrict__ out, long long n4) {
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    Pack4<T>::store(out + i * 4, v.x, v.y, v.z, v.w);
  }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, T* __restrict__ out, long long n) {
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f<T>(x[i]);
}

// =====================================================================================
// Window attention core.  One CTA per (image, window, head); thread i < L owns query i.
// Index arithmetic replaces the reference's pad / roll / split / concat copies (attention.py:19-58):
// slot (si,sj) of window (wi,wj) is rolled position r=(wi*wh+si, wj*ww+sj), i.e. padded position
// p=((r_i-s) mod Hp, (r_j-s) mod Wp); pad positions (p_i>=H or p_j>=W) carry q/k/v = in_proj bias.
// Key bias: shift==0 -> pad keys masked (-inf); shift!=0 -> + xm[channel 0] at ((p_i-s) mod Hp, (p_j-s) mod Wp)
// (attention.py:40 rolls the activation into `mask`), 0 where that position is padding.
// =====================================================================================
template <typename T> struct Vec16;   // 16-byte vector of T
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* out) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const bf16* p, float* out) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      out[2 * i] = __low2float(h2); out[2 * i + 1] = __high2float(h2);
    }
  }
};

// One CTA per (image, window, chunk of HC heads); thread (token, head).  K and V of the window are staged in
// shared memory with 16-byte loads; each thread keeps its query row and output row (D = 32) in registers.
template <typename T, int D, int HC>
__global__ void __launch_bounds__(64 * HC) window_attention_kernel(const T* __restrict__ qkv, const T* __restrict__ xm,
                                                                 const float* __restrict__ b_in, T* __restrict__ att,
                                                                 long long ldo, int H, int W, int C, int wh, int ww,
                                                                 int shift, int Hp, int Wp, const int* __restrict__ skip) {
  pdl_wait();
  if (skip != nullptr && *skip != 0) return;
  constexpr int VN = Vec16<T>::N;
  constexpr int ROW = HC * D;                     // channels of this head chunk
  __shared__ __align__(16) T Ks[64 * ROW];
  __shared__ __align__(16) T Vs[64 * ROW];
  __shared__ float kb[64];
  __shared__ long long rowm[64];
  const int L = wh * ww;
  const int nww = Wp / ww, nwin = (Hp / wh) * nww;
  const int b = blockIdx.x / nwin, win = blockIdx.x % nwin;
  const int wi = win / nww, wj = win % nww;
  const int ch0 = blockIdx.y * ROW;               // first channel of the chunk
  const int tid = threadIdx.x, nthreads = blockDim.x;
  if (tid < L) {
    const int si = tid / ww, sj = tid % ww;
    const int pi = (wi * wh + si - shift + Hp) % Hp, pj = (wj * ww + sj - shift + Wp) % Wp;
    const bool pad = pi >= H || pj >= W;
    rowm[tid] = pad ? -1 : ((long long)b * H + pi) * W + pj;
    float bias = 0.f;
    if (shift == 0) {
      bias = pad ? -INFINITY : 0.f;
    } else {
      const int qi = (pi - shift + Hp) % Hp, qj = (pj - shift + Wp) % Wp;
      if (qi < H && qj < W) bias = to_f(xm[(((long long)b * H + qi) * W + qj) * C]);
    }
    kb[tid] = bias;
  }
  __syncthreads();
  for (int idx = tid; idx < L * (ROW / VN); idx += nthreads) {
    const int tok = idx / (ROW / VN), c = (idx % (ROW / VN)) * VN;
    const long long m = rowm[tok];
    T* kd = Ks + tok * ROW + c;
    T* vd = Vs + tok * ROW + c;
    if (m >= 0) {
      const T* row = qkv + m * 3 * C + ch0 + c;
      *reinterpret_cast<uint4*>(kd) = *reinterpret_cast<const uint4*>(row + C);
      *reinterpret_cast<uint4*>(vd) = *reinterpret_cast<const uint4*>(row + 2 * C);
    } else {
#pragma unroll
      for (int i = 0; i < VN; ++i) { kd[i] = from_f<T>(b_in[C + ch0 + c + i]); vd[i] = from_f<T>(b_in[2 * C + ch0 + c + i]); }
    }
  }
  __syncthreads();
  const int tok = tid % L, hl = tid / L;
  if (hl >= HC) return;
  const long long m = rowm[tok];
  if (m < 0) return;                               // outputs at pad positions are cropped (attention.py:56)
  float q[D], acc[D];
  const float scale = (float)sqrt(1.0 / (double)D);
  {
    const T* row = qkv + m * 3 * C + ch0 + hl * D;
#pragma unroll
    for (int c = 0; c < D; c += VN) Vec16<T>::load(row + c, q + c);
#pragma unroll
    for (int dd = 0; dd < D; ++dd) { q[dd] *= scale; acc[dd] = 0.f; }
  }
  float mx = -INFINITY, den = 0.f;
  for (int j = 0; j < L; ++j) {
    const float bj = kb[j];
    if (bj == -INFINITY) continue;
    const T* kr = Ks + j * ROW + hl * D;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < D; c += VN) {
      float kv[VN];
      Vec16<T>::load(kr + c, kv);
#pragma unroll
      for (int i = 0; i < VN; ++i) s = fmaf(q[c + i], kv[i], s);
    }
    s += bj;
    const float nm = fmaxf(mx, s);
    const float corr = expf(mx - nm), p = expf(s - nm);
    den = den * corr + p;
    const T* vr = Vs + j * ROW + hl * D;
#pragma unroll
    for (int c = 0; c < D; c += VN) {
      float vv[VN];
      Vec16<T>::load(vr + c, vv);
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[c + i] = acc[c + i] * corr + p * vv[i];
    }
    mx = nm;
  }
  T* o = att + m * ldo + ch0 + hl * D;
  const float inv = 1.f / den;
#pragma unroll
  for (int c = 0; c < D; c += 4) Pack4<T>::store(o + c, acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
}

// =====================================================================================
// Few-channel pointwise convolutions at the NCHW fp32 boundary
// =====================================================================================
// in: NCHW fp32 [B, Cin, H*s, W*s], k = stride = s (s = 1: 1x1 conv).  out: [B*H*W, Cout] in TO.
// One CTA per 64 pixels: the J = Cin*s*s inputs of each pixel and the [Cout][J] weights are staged in shared
// memory (coalesced reads along w), then threads sweep (pixel, channel) pairs with coalesced NHWC stores.
constexpr int kPixIn = 64;
template <typename TO>
__global__ void __launch_bounds__(256) pointwise_in_kernel(const float* __restrict__ x, const StepParams* __restrict__ sp,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           TO* __restrict__ out, int B, int Cin, int H, int W, int s, int Cout) {
  pdl_wait();
  extern __shared__ float smem_in[];
  if (sp != nullptr) x = sp->x_in;
  const int J = Cin * s * s;
  float* xs = smem_in;                    // [kPixIn][J + 1]
  float* ws = smem_in + kPixIn * (J + 1); // [J][Cout]  (transposed: consecutive threads read consecutive channels)
  const int M = B * H * W, m0 = blockIdx.x * kPixIn;     // 32-bit pixel index (64-bit divisions are slow)
  for (int i = threadIdx.x; i < Cout * J; i += blockDim.x) ws[(i % J) * Cout + i / J] = w[i];
  for (int i = threadIdx.x; i < kPixIn * J; i += blockDim.x) {
    const int p = i % kPixIn, j = i / kPixIn;           // consecutive threads -> consecutive pixels (contiguous along w)
    const int m = m0 + p;
    float v = 0.f;
    if (m < M) {
      const int ww = m % W, hh = (m / W) % H, b = m / (W * H);
      const int ci = j / (s * s), dy = (j / s) % s, dx = j % s;
      v = x[(((long long)b * Cin + ci) * (H * s) + hh * s + dy) * (W * s) + ww * s + dx];
    }
    xs[p * (J + 1) + j] = v;
  }
  __syncthreads();
  if ((Cout & 3) == 0) {
    const int q = Cout >> 2;                             // channel quads per pixel
    for (int i = threadIdx.x; i < kPixIn * q; i += blockDim.x) {
      const int p = i / q, co = (i % q) * 4;
      const int m = m0 + p;
      if (m >= M) break;
      float4 acc = *reinterpret_cast<const float4*>(bias + co);
      const float* xr = xs + p * (J + 1);
      for (int j = 0; j < J; ++j) {
        const float xv = xr[j];
        const float4 wv = *reinterpret_cast<const float4*>(ws + j * Cout + co);
        acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
      }
      Pack4<TO>::store(out + (long long)m * Cout + co, acc.x, acc.y, acc.z, acc.w);
    }
    return;
  }
  for (int i = threadIdx.x; i < kPixIn * Cout; i += blockDim.x) {
    const int p = i / Cout, co = i % Cout;
    const int m = m0 + p;
    if (m >= M) break;
    float acc = bias[co];
    const float* xr = xs + p * (J + 1);
    for (int j = 0; j < J; ++j) acc = fmaf(xr[j], ws[j * Cout + co], acc);
    out[(long long)m * Cout + co] = from_f<TO>(acc);
  }
}

// Few-output pointwise convolutions out of an NHWC tensor: J (<= 32) outputs per pixel.
// One CTA per 32 pixels: the pixel rows are staged in shared memory in chunks of 128 channels (coalesced), thread
// (pixel p = tid % 32, output j = tid / 32 + 8k) accumulates its dot product, results land in e[j][p].
// wmat is [C][J] (ConvTranspose layout) when w_cj, else [J][C] (Conv2d layout).
template <typename T>
__device__ __forceinline__ void pixel_dots_32(const T* __restrict__ x, long long m0, long long M, int C, int J,
                                              const float* __restrict__ wmat, bool w_cj, float* xs /*[32][129]*/,
                                              float* ws /*[128][J]*/, float (*e)[33]) {
  const int p = threadIdx.x & 31, jg = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < C; c0 += 128) {
    const int cn = min(128, C - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * cn; i += blockDim.x) {
      const int pp = i / cn, c = i % cn;
      const long long m = m0 + pp;
      xs[pp * 129 + c] = m < M ? to_f(x[m * C + c0 + c]) : 0.f;
    }
    for (int i = threadIdx.x; i < cn * J; i += blockDim.x) {
      const int c = i / J, j = i % J;
      ws[i] = w_cj ? wmat[(long long)(c0 + c) * J + j] : wmat[(long long)j * C + c0 + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = jg + 8 * k;
      if (j < J) {
        float a = acc[k];
        for (int c = 0; c < cn; ++c) a = fmaf(xs[p * 129 + c], ws[c * J + j], a);
        acc[k] = a;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = jg + 8 * k;
    if (j < J) e[j][p] = acc[k];
  }
}

// decoder_last (ConvTranspose2d k=s stride=s, unet.py:78,102) + DDIM update (ddpm.py:81-91)
__global__ void __launch_bounds__(256) final_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ bias, const StepParams* __restrict__ sp,
                                                    int B, int Cin, int H, int W, int s, int C0) {
  pdl_wait();
  __shared__ float e[32][33];
  __shared__ float xs[32 * 129];
  __shared__ float ws[128 * 32];
  const StepParams co = *sp;
  const float* __restrict__ xin = co.x_in;
  const float* __restrict__ noise = co.noise;
  float* __restrict__ out = co.out;
  const long long M = (long long)B * H * W, m0 = (long long)blockIdx.x * 32;
  const int J = Cin * s * s;
  pixel_dots_32<float>(x, m0, M, C0, J, w, true, xs, ws, e);
  __syncthreads();
  for (int idx = threadIdx.x; idx < J * 32; idx += blockDim.x) {
    const int j = idx >> 5, p = idx & 31;
    const long long m = m0 + p;
    if (m >= M) continue;
    const int ci = j / (s * s), dy = (j / s) % s, dx = j % s;
    const int ww = (int)(m % W), hh = (int)((m / W) % H), b = (int)(m / ((long long)W * H));
    const long long o = (((long long)b * Cin + ci) * (H * s) + hh * s + dy) * (W * s) + ww * s + dx;
    const float eps = e[j][p] + bias[ci];
    float r = eps;
    if (co.ddim_enabled) {
      const float x0 = (xin[o] - co.c_eps_in * eps) / co.c_div;
      if (co.final_step) r = x0;
      else {
        r = co.c_x0 * x0 + co.c_eps_out * eps;
        r += co.sigma * (noise ? noise[o] : 0.f);
      }
    }
    out[o] = r;
  }
}

// decoder_last + DDIM update, fast path for C0 = 128*V channels and J = Cin*s*s <= 8 outputs per pixel: one warp per
// pixel (the row is one coalesced 16-byte load per lane), the lane's [J][4V] weight slice lives in registers, the 8
// partial dot products are folded across the warp with 9 shuffles, and the 8 lanes that end up owning one output
// each apply the DDIM update (ddpm.py:81-91) and write NCHW.  Bytes: 4*C0 per pixel in, 12*J in/out.
template <int V>
__global__ void __launch_bounds__(256) final_warp_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const StepParams* __restrict__ sp,
                                                         int B, int Cin, int H, int W, int s, int C0) {
  pdl_wait();
  const StepParams co = *sp;
  const float* __restrict__ xin = co.x_in;
  const float* __restrict__ noise = co.noise;
  float* __restrict__ out = co.out;
  const int lane = threadIdx.x & 31;
  const int J = Cin * s * s;
  // weights [C0][J] -> shared [8][C0] (transposed, zero rows for j >= J) with coalesced loads, then 16-byte reads:
  // loading the lane's slice straight from global costs 32 sectors per load instruction (stride J floats per lane)
  __shared__ __align__(16) float wsh[8 * 128 * V];
  for (int i = threadIdx.x; i < 8 * 128 * V; i += blockDim.x) wsh[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * V * J; i += blockDim.x) wsh[(i % J) * (128 * V) + i / J] = w[i];
  __syncthreads();
  float wr[8][4 * V];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(wsh + j * (128 * V) + i * 128 + lane * 4);
      wr[j][4 * i] = t.x; wr[j][4 * i + 1] = t.y; wr[j][4 * i + 2] = t.z; wr[j][4 * i + 3] = t.w;
    }
  // output owned by this lane after the fold: bit 4 -> +4, bit 3 -> +2, bit 2 -> +1
  const int jl = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  const float bj = jl < J ? bias[jl / (s * s)] : 0.f;
  const int M = B * H * W;                                // 32-bit pixel index: 64-bit divisions cost ~100 instructions each
  const int wpg = (gridDim.x * blockDim.x) >> 5;
  constexpr int PX = 4;                                   // pixels per warp iteration: PX independent row loads in flight
  const bool owner = (lane & 3) == 0 && jl < J;
  const int ci = jl / (s * s), dy = (jl / s) % s, dx = jl % s;
  for (int mb = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * PX; mb < M; mb += wpg * PX) {
    float4 v[PX][V];
    long long o[PX];
    float xi[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      const int m = mb + p;
#pragma unroll
      for (int i = 0; i < V; ++i)
        v[p][i] = m < M ? __ldg(reinterpret_cast<const float4*>(x + (long long)m * C0 + i * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int ww = m % W, hh = (m / W) % H, b = m / (W * H);
      o[p] = (((long long)b * Cin + ci) * (H * s) + hh * s + dy) * (W * s) + ww * s + dx;
      xi[p] = (owner && m < M && co.ddim_enabled) ? xin[o[p]] : 0.f;
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < V; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          a[j] = fmaf(v[p][i].x, wr[j][4 * i], fmaf(v[p][i].y, wr[j][4 * i + 1], fmaf(v[p][i].z, wr[j][4 * i + 2], fmaf(v[p][i].w, wr[j][4 * i + 3], a[j]))));
      }
      {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float r = __shfl_xor_sync(0xffffffffu, hi ? a[i] : a[i + 4], 16); a[i] = (hi ? a[i + 4] : a[i]) + r; }
      }
      {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) { const float r = __shfl_xor_sync(0xffffffffu, hi ? a[i] : a[i + 2], 8); a[i] = (hi ? a[i + 2] : a[i]) + r; }
      }
      {
        const bool hi = lane & 4;
        const float r = __shfl_xor_sync(0xffffffffu, hi ? a[0] : a[1], 4);
        a[0] = (hi ? a[1] : a[0]) + r;
      }
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
      if (owner && mb + p < M) {
        const float eps = a[0] + bj;
        float r = eps;
        if (co.ddim_enabled) {
          const float x0 = (xi[p] - co.c_eps_in * eps) / co.c_div;
          if (co.final_step) r = x0;
          else {
            r = co.c_x0 * x0 + co.c_eps_out * eps;
            r += co.sigma * (noise ? noise[o[p]] : 0.f);
          }
        }
        out[o[p]] = r;
      }
    }
  }
}

// -------------------------------------------------------------------------------------
// TF32 tensor-core versions of the two edge layers for the bf16 mode (the fp32 validation mode keeps the exact kernels
// above).  Both are skinny GEMMs -- stem [M x J<=8] . [J x C0], last layer [M x C0] . [C0 x J<=8] -- whose SIMT forms
// are instruction-bound (190 / 225 warp instructions per pixel against 33.5 MB of fp32 traffic at config 2); as
// mma.sync.m16n8k8 they cost 4 / 10 and run at the speed of the x tensor they write / read.
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// offset of input/output element j of pixel m in the NCHW fp32 tensor [B, Cin, H*s, W*s] (k = stride = s)
__device__ __forceinline__ long long nchw_offset(int m, int j, int Cin, int H, int W, int s) {
  const int ww = m % W, hh = (m / W) % H, b = m / (W * H);
  const int ci = j / (s * s), dy = (j / s) % s, dx = j % s;
  return (((long long)b * Cin + ci) * (H * s) + hh * s + dy) * (W * s) + ww * s + dx;
}

// encoder_first (unet.py:90): out fp32 [M, C0 = 128*V] = x_in(NCHW) . W^T + bias.  A warp takes 16 pixels: the A fragment is
// 4 scalar loads per lane (8 consecutive pixels of one input plane per 32-byte sector), the [C0 x 8] weights live in
// registers as B fragments, every MMA yields an 8-channel slice that is stored as one full 32-byte sector per pixel.
template <int V>
__global__ void __launch_bounds__(256) stem_mma_kernel(const StepParams* __restrict__ sp, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out,
                                                       int B, int Cin, int H, int W, int s) {
  pdl_wait();
  constexpr int C0 = 128 * V, NT = C0 / 8;
  const float* __restrict__ x = sp->x_in;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int J = Cin * s * s, M = B * H * W;
  uint32_t wb[NT][2];
  float bz[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    wb[nt][0] = to_tf32(t < J ? w[(nt * 8 + g) * J + t] : 0.f);            // B(k = t,     n = g)
    wb[nt][1] = to_tf32(t + 4 < J ? w[(nt * 8 + g) * J + t + 4] : 0.f);    // B(k = t + 4, n = g)
    bz[nt][0] = bias[nt * 8 + 2 * t]; bz[nt][1] = bias[nt * 8 + 2 * t + 1];
  }
  const int wpg = (gridDim.x * blockDim.x) >> 5;
  for (int mb = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 16; mb < M; mb += wpg * 16) {
    const int m0 = mb + g, m1 = mb + g + 8;
    const uint32_t a0 = to_tf32(m0 < M && t < J ? x[nchw_offset(m0, t, Cin, H, W, s)] : 0.f);
    const uint32_t a1 = to_tf32(m1 < M && t < J ? x[nchw_offset(m1, t, Cin, H, W, s)] : 0.f);
    const uint32_t a2 = to_tf32(m0 < M && t + 4 < J ? x[nchw_offset(m0, t + 4, Cin, H, W, s)] : 0.f);
    const uint32_t a3 = to_tf32(m1 < M && t + 4 < J ? x[nchw_offset(m1, t + 4, Cin, H, W, s)] : 0.f);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float c[4] = {bz[nt][0], bz[nt][1], bz[nt][0], bz[nt][1]};
      mma_tf32(c, a0, a1, a2, a3, wb[nt][0], wb[nt][1]);
      if (m0 < M) *reinterpret_cast<float2*>(out + (long long)m0 * C0 + nt * 8 + 2 * t) = make_float2(c[0], c[1]);
      if (m1 < M) *reinterpret_cast<float2*>(out + (long long)m1 * C0 + nt * 8 + 2 * t) = make_float2(c[2], c[3]);
    }
  }
}

// decoder_last (unet.py:102) + DDIM update (ddpm.py:81-91): eps[m, j] = x[m, :] . W[:, j] + bias.  A warp takes 16 pixels;
// a lane loads 16-byte pieces of rows g and g + 8 and feeds them to the MMAs with the k index permuted (the dot product
// does not care; the register-resident B fragments use the same permutation): piece (x, y, z, w) of 16-channel chunk kc =
// channels 16 kc + 4 t + {0, 1, 2, 3} -> k-step 2 kc uses (x, y) as columns (t, t + 4), k-step 2 kc + 1 uses (z, w).
template <int V>
__global__ void __launch_bounds__(256) final_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const StepParams* __restrict__ sp,
                                                        int B, int Cin, int H, int W, int s) {
  pdl_wait();
  constexpr int C0 = 128 * V, KC = C0 / 16;
  const StepParams co = *sp;
  const float* __restrict__ xin = co.x_in;
  const float* __restrict__ noise = co.noise;
  float* __restrict__ out = co.out;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int J = Cin * s * s, M = B * H * W;
  uint32_t wb[KC][4];                                     // B(k, n = g) for the four channels of this lane's piece
#pragma unroll
  for (int kc = 0; kc < KC; ++kc)
#pragma unroll
    for (int i = 0; i < 4; ++i) wb[kc][i] = to_tf32(g < J ? w[(kc * 16 + 4 * t + i) * J + g] : 0.f);
  const int j0 = 2 * t, j1 = 2 * t + 1;                   // the outputs this lane owns after the MMAs (rows g and g + 8)
  const float b0 = j0 < J ? bias[j0 / (s * s)] : 0.f, b1 = j1 < J ? bias[j1 / (s * s)] : 0.f;
  const int wpg = (gridDim.x * blockDim.x) >> 5;
  for (int mb = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 16; mb < M; mb += wpg * 16) {
    const int m0 = mb + g, m1 = mb + g + 8;
    float4 xa[KC], xb[KC];
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      xa[kc] = m0 < M ? __ldg(reinterpret_cast<const float4*>(x + (long long)m0 * C0 + kc * 16 + 4 * t)) : make_float4(0.f, 0.f, 0.f, 0.f);
      xb[kc] = m1 < M ? __ldg(reinterpret_cast<const float4*>(x + (long long)m1 * C0 + kc * 16 + 4 * t)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float c[4] = {b0, b1, b0, b1};
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      mma_tf32(c, to_tf32(xa[kc].x), to_tf32(xb[kc].x), to_tf32(xa[kc].y), to_tf32(xb[kc].y), wb[kc][0], wb[kc][1]);
      mma_tf32(c, to_tf32(xa[kc].z), to_tf32(xb[kc].z), to_tf32(xa[kc].w), to_tf32(xb[kc].w), wb[kc][2], wb[kc][3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = i < 2 ? m0 : m1, j = (i & 1) ? j1 : j0;
      if (m >= M || j >= J) continue;
      const long long o = nchw_offset(m, j, Cin, H, W, s);
      const float eps = c[i];
      float r = eps;
      if (co.ddim_enabled) {
        const float x0 = (xin[o] - co.c_eps_in * eps) / co.c_div;
        if (co.final_step) r = x0;
        else {
          r = co.c_x0 * x0 + co.c_eps_out * eps;
          r += co.sigma * (noise ? noise[o] : 0.f);
        }
      }
      out[o] = r;
    }
  }
}

// NHWC T [B,H,W,C] -> NCHW fp32 [B,Cout,H,W] 1x1 conv (+ bilinear x2 of prev [B,Cout,H/2,W/2], + uint8 HWC copy)
template <typename T>
__global__ void __launch_bounds__(256) pointwise_out_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, const float* __restrict__ prev,
                                                            float* __restrict__ out, uint8_t* __restrict__ out_u8,
                                                            int B, int H, int W, int C, int Cout) {
  pdl_wait();
  __shared__ float e[32][33];
  __shared__ float xs[32 * 129];
  __shared__ float ws[128 * 32];
  const long long M = (long long)B * H * W, m0 = (long long)blockIdx.x * 32;
  pixel_dots_32<T>(x, m0, M, C, Cout, w, false, xs, ws, e);
  __syncthreads();
  for (int idx = threadIdx.x; idx < Cout * 32; idx += blockDim.x) {
    const int j = idx >> 5, p = idx & 31;
    const long long m = m0 + p;
    if (m >= M) continue;
    const int ww = (int)(m % W), hh = (int)((m / W) % H), b = (int)(m / ((long long)W * H));
    float v = e[j][p] + bias[j];
    if (prev) {
      // F.interpolate(scale_factor=2, 'bilinear', align_corners=False), vae.py:131
      const int Hs = H / 2, Ws = W / 2;
      const int i0 = hh >> 1, j0 = ww >> 1;
      const int ia = (hh & 1) ? i0 : max(i0 - 1, 0), ib = (hh & 1) ? min(i0 + 1, Hs - 1) : i0;
      const int ja = (ww & 1) ? j0 : max(j0 - 1, 0), jb = (ww & 1) ? min(j0 + 1, Ws - 1) : j0;
      const float wa_h = (hh & 1) ? 0.75f : 0.25f, wa_w = (ww & 1) ? 0.75f : 0.25f;
      const float* pp = prev + ((long long)b * Cout + j) * Hs * Ws;
      const float top = wa_w * pp[ia * Ws + ja] + (1.f - wa_w) * pp[ia * Ws + jb];
      const float bot = wa_w * pp[ib * Ws + ja] + (1.f - wa_w) * pp[ib * Ws + jb];
      v += wa_h * top + (1.f - wa_h) * bot;
    }
    if (out) out[(((long long)b * Cout + j) * H + hh) * W + ww] = v;
    if (out_u8) {
      const float c = fminf(fmaxf(v, -1.f), 1.f);
      out_u8[m * Cout + j] = (uint8_t)(c * 127.5f + 127.5f);   // truncation, sample_ldm.py:77
    }
  }
}

// to_rgb-style few-output 1x1 conv out of an NHWC bf16 tensor, fast path (Cout <= 4, C a multiple of 64): G lanes
// share a pixel (8 channels = one 16-byte load per lane per 8*G channels), partial dot products are folded with
// log2(G) shuffle steps, the group's first lane adds the bilinear-upsampled running sum (vae.py:131) and writes
// NCHW fp32 (+ the uint8 HWC image of sample_ldm.py:75-77).  Bytes per pixel: 2C in, 4*Cout (+ Cout) out.
template <int G, int ITERS>
__global__ void __launch_bounds__(256) pointwise_out_warp_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, const float* __restrict__ prev,
                                                                 float* __restrict__ out, uint8_t* __restrict__ out_u8,
                                                                 int B, int H, int W, int C, int Cout) {
  pdl_wait();
  constexpr int PPW = 32 / G;                              // pixels per warp per pass
  constexpr int UNR = 4;                                   // pixel groups per pass: UNR*ITERS independent 16-byte loads in flight per lane
  constexpr int SUBS = 32 / (PPW * UNR);                   // passes per 32-pixel chunk
  static_assert(SUBS >= 1 && SUBS * PPW * UNR == 32, "a warp finishes 32 consecutive pixels per chunk");
  __shared__ float s_dot[8][4][32];                        // [warp][output channel][pixel of the chunk]: dot products, re-read pixel-per-lane
  const int lane = threadIdx.x & 31, gl = lane % G, gp = lane / G, wib = threadIdx.x >> 5;
  const int M = B * H * W;
  float wr[ITERS][4][8];                                   // [pass][j][8 channels]; C = ITERS * 8 * G
#pragma unroll
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) wr[it][j][e] = j < Cout ? w[(long long)j * C + it * 8 * G + gl * 8 + e] : 0.f;
  const int wpg = (gridDim.x * blockDim.x) >> 5;
  for (int mb = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; mb < M; mb += wpg * 32) {
    // ---- phase 1: G lanes per pixel reduce the C channels (16-byte loads), every pixel's Cout dot products go to shared memory
#pragma unroll
    for (int sub = 0; sub < SUBS; ++sub) {
      uint4 u[UNR][ITERS];
#pragma unroll
      for (int r = 0; r < UNR; ++r) {
        const int m = mb + (sub * UNR + r) * PPW + gp;
#pragma unroll
        for (int it = 0; it < ITERS; ++it)
          u[r][it] = m < M ? __ldg(reinterpret_cast<const uint4*>(x + (long long)m * C + it * 8 * G + gl * 8)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int r = 0; r < UNR; ++r) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const uint32_t w4[4] = {u[r][it].x, u[r][it].y, u[r][it].z, u[r][it].w};
          float xv[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[k]);
            xv[2 * k] = __low2float(h2); xv[2 * k + 1] = __high2float(h2);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) a[j] = fmaf(xv[e], wr[it][j][e], a[j]);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
        // the butterfly leaves every lane of the group with all sums: lane gl publishes output channel j = gl
        if (gl < 4) s_dot[wib][gl][(sub * UNR + r) * PPW + gp] = gl == 0 ? a[0] : (gl == 1 ? a[1] : (gl == 2 ? a[2] : a[3]));
      }
    }
    __syncwarp();
    // ---- phase 2: lane = pixel: + bias, + bilinear-upsampled running sum (vae.py:131), coalesced NCHW fp32 rows and HWC uint8 bytes
    const int m = mb + lane;
    if (m < M) {
      const int ww = m % W, hh = (m / W) % H, b = m / (W * H);
      for (int j = 0; j < Cout; ++j) {
        float v = s_dot[wib][j][lane] + bias[j];
        if (prev) {
          const int Hs = H / 2, Ws = W / 2;
          const int i0 = hh >> 1, j0 = ww >> 1;
          const int ia = (hh & 1) ? i0 : max(i0 - 1, 0), ib = (hh & 1) ? min(i0 + 1, Hs - 1) : i0;
          const int ja = (ww & 1) ? j0 : max(j0 - 1, 0), jb = (ww & 1) ? min(j0 + 1, Ws - 1) : j0;
          const float wa_h = (hh & 1) ? 0.75f : 0.25f, wa_w = (ww & 1) ? 0.75f : 0.25f;
          const float* pp = prev + ((long long)b * Cout + j) * Hs * Ws;
          const float top = wa_w * pp[ia * Ws + ja] + (1.f - wa_w) * pp[ia * Ws + jb];
          const float bot = wa_w * pp[ib * Ws + ja] + (1.f - wa_w) * pp[ib * Ws + jb];
          v += wa_h * top + (1.f - wa_h) * bot;
        }
        if (out) out[(((long long)b * Cout + j) * H + hh) * W + ww] = v;
        if (out_u8) {
          const float c = fminf(fmaxf(v, -1.f), 1.f);
          out_u8[(long long)m * Cout + j] = (uint8_t)(c * 127.5f + 127.5f);   // truncation, sample_ldm.py:77
        }
      }
    }
    __syncwarp();                                            // s_dot is rewritten by the next chunk
  }
}

// =====================================================================================
// Per-image plans (the reference's batch-1 loops draw the stochastic-depth / expert decisions per image,
// sample_ldm.py:71-72): every image of a batch carries its own plan word  skip | e1 << 8 | e2 << 16.
// The feed-forward then evaluates all five ReGLU experts and these kernels apply the per-image choice.
// =====================================================================================
// Per-image decisions (a batch of the reference's batch-1 forwards): rows of one SwinBlock entering / leaving it.
// Entering (after the norm has read x): rows of images that skip the block (unet.py:39-40) are saved; the other rows get
// the c-projection biases of their image's experts, b_c[general] + b_c[e1] + b_c[e2] (+ out_proj bias) -- the branch
// kernels that follow only ever ADD to x, so the bias can go first.  Leaving: the saved rows are put back.
__global__ void rows_enter_kernel(float* __restrict__ x, float* __restrict__ backup, const float* __restrict__ b_c,
                                  const int* __restrict__ plan_img, int M, int HW, int C, int attn) {
  pdl_wait();
  const int C4 = C / 4;
  const long long total = (long long)M * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / C4), c = (int)(i % C4) * 4;
    const int w = plan_img[m / HW];
    float4* xp = reinterpret_cast<float4*>(x) + i;
    if (w & 1) { reinterpret_cast<float4*>(backup)[i] = *xp; continue; }
    if (b_c == nullptr) continue;
    const int e1 = (w >> 8) & 0xff, e2 = (w >> 16) & 0xff;
    float4 v = *xp;
    const float4 g = __ldg(reinterpret_cast<const float4*>(b_c + c));
    const float4 a = __ldg(reinterpret_cast<const float4*>(b_c + (long long)(1 + e1) * C + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(b_c + (long long)(1 + e2) * C + c));
    v.x += (g.x + a.x) + b.x; v.y += (g.y + a.y) + b.y; v.z += (g.z + a.z) + b.z; v.w += (g.w + a.w) + b.w;
    if (attn) {
      const float4 o = __ldg(reinterpret_cast<const float4*>(b_c + 5LL * C + c));
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *xp = v;
  }
}
__global__ void rows_leave_kernel(float* __restrict__ x, const float* __restrict__ backup, const int* __restrict__ plan_img, int M,
                                  int HW, int C) {
  pdl_wait();
  const int C4 = C / 4;
  const long long total = (long long)M * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / C4);
    if (plan_img[m / HW] & 1) reinterpret_cast<float4*>(x)[i] = reinterpret_cast<const float4*>(backup)[i];
  }
}

inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

// =====================================================================================
// launchers
// =====================================================================================
cudaError_t launch_gemm_simt(const GemmDesc& d, bool is_bf16, cudaStream_t s) {
  return is_bf16 ? gemm_simt_dispatch<bf16>(d, s) : gemm_simt_dispatch<float>(d, s);
}

cudaError_t launch_repack(const float* src, void* dst, bool dst_bf16, const int dims[4], const long long sstr[4],
                          const long long dstr[4], cudaStream_t s) {
  Repack4 r;
  long long total = 1;
  for (int i = 0; i < 4; ++i) { r.dims[i] = dims[i]; r.ss[i] = sstr[i]; r.ds[i] = dstr[i]; total *= dims[i]; }
  if (total == 0) return cudaSuccess;
  if (dst_bf16) launch_k((repack_kernel<bf16>), grid_for(total, 256), 256, 0, s, src, (bf16*)dst, r, total);
  else launch_k((repack_kernel<float>), grid_for(total, 256), 256, 0, s, src, (float*)dst, r, total);
  return cudaGetLastError();
}

cudaError_t launch_stem(const StepParams* sp, const float* w, const float* bias, float* out, int B, int Cin, int H, int W,
                        int s, int C0, bool tf32, cudaStream_t st) {
  const long long M = (long long)B * H * W;
  const int J = Cin * s * s;
  if (tf32 && J <= 8 && (C0 == 128 || C0 == 256) && M < (1LL << 31)) {
    // a warp per 16 pixels, at most one resident wave of CTAs (122 / 220 registers per thread): the warps loop instead
    const int grid = grid_for(M * 2, 256, C0 == 128 ? 148 * 2 : 148);
    if (C0 == 128) launch_k((stem_mma_kernel<1>), grid, 256, 0, st, sp, w, bias, out, B, Cin, H, W, s);
    else launch_k((stem_mma_kernel<2>), grid, 256, 0, st, sp, w, bias, out, B, Cin, H, W, s);
    return cudaGetLastError();
  }
  const size_t smem = (size_t)(kPixIn * (J + 1) + C0 * J) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorNotSupported;
  launch_k((pointwise_in_kernel<float>), (unsigned)((M + kPixIn - 1) / kPixIn), 256, smem, st, nullptr, sp, w, bias, out, B, Cin, H, W, s, C0);
  return cudaGetLastError();
}

cudaError_t launch_nchw_pointwise_in(const float* x, const float* w, const float* bias, void* out, bool is_bf16,
                                     int B, int Cin, int H, int W, int Cout, cudaStream_t st) {
  const long long M = (long long)B * H * W;
  const size_t smem = (size_t)(kPixIn * (Cin + 1) + Cout * Cin) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorNotSupported;
  const unsigned grid = (unsigned)((M + kPixIn - 1) / kPixIn);
  if (is_bf16) launch_k((pointwise_in_kernel<bf16>), grid, 256, smem, st, x, nullptr, w, bias, (bf16*)out, B, Cin, H, W, 1, Cout);
  else launch_k((pointwise_in_kernel<float>), grid, 256, smem, st, x, nullptr, w, bias, (float*)out, B, Cin, H, W, 1, Cout);
  return cudaGetLastError();
}

template <typename T, int MAXV, int ROWS>
static void norm_film_launch(const float* x, const float* film, const int* t_index, T* out, int M, int C, int HW, float eps,
                             const int* skip, cudaStream_t st) {
  const long long items = (long long)((M / HW + ROWS - 1) / ROWS) * HW;     // warps of work
  launch_k((norm_film_kernel<T, MAXV, ROWS>), grid_for(items * 32, 256, 148 * 8), 256, 0, st, x, film, t_index, out, M, C, HW, eps, skip);
}

template <typename T>
static cudaError_t norm_film_dispatch(const float* x, const float* film, const int* t_index, T* out, int M, int C,
                                      int HW, float eps, const int* skip, cudaStream_t st) {
  if (M % HW != 0) return cudaErrorInvalidValue;
  if (C == 128) {
    const long long items = (long long)((M / HW + 3) / 4) * ((HW + 1) / 2);
    launch_k((norm_film_c128_kernel<T, 4>), grid_for(items * 32, 256, 148 * 8), 256, 0, st, x, film, t_index, out, M, HW, eps, skip);
  } else if (C <= 128) norm_film_launch<T, 1, 4>(x, film, t_index, out, M, C, HW, eps, skip, st);
  else if (C <= 256) norm_film_launch<T, 2, 4>(x, film, t_index, out, M, C, HW, eps, skip, st);
  else if (C <= 512) norm_film_launch<T, 4, 2>(x, film, t_index, out, M, C, HW, eps, skip, st);
  else if (C <= 1024) norm_film_launch<T, 8, 1>(x, film, t_index, out, M, C, HW, eps, skip, st);
  else if (C <= 2048) norm_film_launch<T, 16, 1>(x, film, t_index, out, M, C, HW, eps, skip, st);
  else return cudaErrorNotSupported;
  return cudaGetLastError();
}

cudaError_t launch_norm_film(const float* x, const float* film, const int* t_index, void* out, bool is_bf16, int M,
                             int C, int HW, float eps, const int* skip, cudaStream_t st) {
  if (C % 4 != 0) return cudaErrorNotSupported;
  return is_bf16 ? norm_film_dispatch<bf16>(x, film, t_index, (bf16*)out, M, C, HW, eps, skip, st)
                 : norm_film_dispatch<float>(x, film, t_index, (float*)out, M, C, HW, eps, skip, st);
}

cudaError_t launch_emb_build(const float* pe, const float* te, void* emb, bool is_bf16, int n_t, int HW, int C,
                             cudaStream_t st) {
  const long long total = (long long)n_t * HW * 2 * C;
  if (is_bf16) launch_k((emb_build_kernel<bf16>), grid_for(total, 256), 256, 0, st, pe, te, (bf16*)emb, n_t, HW, C);
  else launch_k((emb_build_kernel<float>), grid_for(total, 256), 256, 0, st, pe, te, (float*)emb, n_t, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_pool_cast(const float* x, void* out, bool is_bf16, int B, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)B * (H / 2) * (W / 2) * C;
  if (C % 4 == 0) {
    if (is_bf16) launch_k((pool2_vec4_kernel<bf16>), grid_for(total / 4, 256), 256, 0, st, x, (bf16*)out, B, H, W, C / 4);
    else launch_k((pool2_vec4_kernel<float>), grid_for(total / 4, 256), 256, 0, st, x, (float*)out, B, H, W, C / 4);
    return cudaGetLastError();
  }
  if (is_bf16) launch_k((pool2_kernel<float, bf16>), grid_for(total, 256), 256, 0, st, x, (bf16*)out, B, H, W, C);
  else launch_k((pool2_kernel<float, float>), grid_for(total, 256), 256, 0, st, x, (float*)out, B, H, W, C);
  return cudaGetLastError();
}

cudaError_t launch_pool_t(const void* x, void* out, bool is_bf16, int B, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)B * (H / 2) * (W / 2) * C;
  if (is_bf16) launch_k((pool2_kernel<bf16, bf16>), grid_for(total, 256), 256, 0, st, (const bf16*)x, (bf16*)out, B, H, W, C);
  else launch_k((pool2_kernel<float, float>), grid_for(total, 256), 256, 0, st, (const float*)x, (float*)out, B, H, W, C);
  return cudaGetLastError();
}

cudaError_t launch_cast(const float* x, void* out, bool is_bf16, long long n, cudaStream_t st) {
  if (n % 4 == 0) {
    if (is_bf16) launch_k((cast_vec4_kernel<bf16>), grid_for(n / 4, 256), 256, 0, st, x, (bf16*)out, n / 4);
    else launch_k((cast_vec4_kernel<float>), grid_for(n / 4, 256), 256, 0, st, x, (float*)out, n / 4);
    return cudaGetLastError();
  }
  if (is_bf16) launch_k((cast_kernel<bf16>), grid_for(n, 256), 256, 0, st, x, (bf16*)out, n);
  else launch_k((cast_kernel<float>), grid_for(n, 256), 256, 0, st, x, (float*)out, n);
  return cudaGetLastError();
}

template <typename T, int HC>
static cudaError_t window_attention_launch(const void* qkv, const void* xm, const float* b_in, void* att, long long ldo,
                                           int B, int H, int W, int C, int win_h, int win_w, int shift, const int* skip,
                                           cudaStream_t st) {
  const int Hp = (H + win_h - 1) / win_h * win_h, Wp = (W + win_w - 1) / win_w * win_w;
  const int L = win_h * win_w, heads = C / 32;
  dim3 grid(B * (Hp / win_h) * (Wp / win_w), heads / HC);
  if (grid.y > 65535u) return cudaErrorNotSupported;
  int threads = ((L * HC + 31) / 32) * 32;
  launch_k((window_attention_kernel<T, 32, HC>), grid, threads, 0, st, (const T*)qkv, (const T*)xm, b_in, (T*)att, ldo, H, W, C,
                                                              win_h, win_w, shift, Hp, Wp, skip);
  return cudaGetLastError();
}

cudaError_t launch_window_attention(const void* qkv, const void* xm, const float* b_in, void* att, long long ldo,
                                    bool is_bf16, int B, int H, int W, int C, int head_dim, int win_h, int win_w,
                                    int shift, const int* skip, cudaStream_t st, bool force_simt) {
  if (head_dim != 32 || win_h * win_w > 64 || C % 32) return cudaErrorNotSupported;
  const int heads = C / 32;
  if (is_bf16 && !force_simt && window_attention_mma_supported(C, head_dim, win_h, win_w, ldo))
    return launch_window_attention_mma(qkv, xm, b_in, att, ldo, B, H, W, C, win_h, win_w, shift, skip, st);
  if (is_bf16) {
    if (heads % 4 == 0) return window_attention_launch<bf16, 4>(qkv, xm, b_in, att, ldo, B, H, W, C, win_h, win_w, shift, skip, st);
    if (heads % 2 == 0) return window_attention_launch<bf16, 2>(qkv, xm, b_in, att, ldo, B, H, W, C, win_h, win_w, shift, skip, st);
    return window_attention_launch<bf16, 1>(qkv, xm, b_in, att, ldo, B, H, W, C, win_h, win_w, shift, skip, st);
  }
  if (heads % 2 == 0) return window_attention_launch<float, 2>(qkv, xm, b_in, att, ldo, B, H, W, C, win_h, win_w, shift, skip, st);
  return window_attention_launch<float, 1>(qkv, xm, b_in, att, ldo, B, H, W, C, win_h, win_w, shift, skip, st);
}

cudaError_t launch_final(const float* x, const float* w, const float* bias, const StepParams* sp, int B, int Cin, int H,
                         int W, int s, int C0, bool tf32, cudaStream_t st) {
  if (Cin * s * s > 32) return cudaErrorNotSupported;
  const long long M = (long long)B * H * W;
  if (tf32 && Cin * s * s <= 8 && (C0 == 128 || C0 == 256) && M < (1LL << 31)) {
    const int grid = grid_for(M * 2, 256, 148 * 2);           // a warp per 16 pixels, one resident wave (measured: 296 CTAs best)
    if (C0 == 128) launch_k((final_mma_kernel<1>), grid, 256, 0, st, x, w, bias, sp, B, Cin, H, W, s);
    else launch_k((final_mma_kernel<2>), grid, 256, 0, st, x, w, bias, sp, B, Cin, H, W, s);
    return cudaGetLastError();
  }
  if (Cin * s * s <= 8 && (C0 == 128 || C0 == 256)) {
    const int grid = grid_for(M * 8, 256, 148 * 8);
    if (C0 == 128) launch_k((final_warp_kernel<1>), grid, 256, 0, st, x, w, bias, sp, B, Cin, H, W, s, C0);
    else launch_k((final_warp_kernel<2>), grid, 256, 0, st, x, w, bias, sp, B, Cin, H, W, s, C0);
    return cudaGetLastError();
  }
  launch_k((final_kernel), (unsigned)((M + 31) / 32), 256, 0, st, x, w, bias, sp, B, Cin, H, W, s, C0);
  return cudaGetLastError();
}

cudaError_t launch_nhwc_pointwise_out(const void* x, bool is_bf16, const float* w, const float* bias,
                                      const float* prev, float* out, uint8_t* out_u8, int B, int H, int W, int C,
                                      int Cout, cudaStream_t st) {
  if (Cout > 32) return cudaErrorNotSupported;
  const long long M = (long long)B * H * W;
  const unsigned grid = (unsigned)((M + 31) / 32);
  if (is_bf16 && Cout <= 4 && (C == 64 || C == 128 || C == 256 || C == 512) && M < (1LL << 31)) {
    const int G = C >= 256 ? 32 : C / 8;                   // lanes per pixel, 8 channels per lane per pass
    const int g2 = grid_for((long long)((M + 31) / 32) * 32, 256, 148 * 8);      // a warp per 32-pixel chunk
    const bf16* xb = (const bf16*)x;
    if (C == 512) launch_k((pointwise_out_warp_kernel<32, 2>), g2, 256, 0, st, xb, w, bias, prev, out, out_u8, B, H, W, C, Cout);
    else if (C == 256) launch_k((pointwise_out_warp_kernel<32, 1>), g2, 256, 0, st, xb, w, bias, prev, out, out_u8, B, H, W, C, Cout);
    else if (C == 128) launch_k((pointwise_out_warp_kernel<16, 1>), g2, 256, 0, st, xb, w, bias, prev, out, out_u8, B, H, W, C, Cout);
    else launch_k((pointwise_out_warp_kernel<8, 1>), g2, 256, 0, st, xb, w, bias, prev, out, out_u8, B, H, W, C, Cout);
    return cudaGetLastError();
  }
  if (is_bf16) launch_k((pointwise_out_kernel<bf16>), grid, 256, 0, st, (const bf16*)x, w, bias, prev, out, out_u8, B, H, W, C, Cout);
  else launch_k((pointwise_out_kernel<float>), grid, 256, 0, st, (const float*)x, w, bias, prev, out, out_u8, B, H, W, C, Cout);
  return cudaGetLastError();
}

cudaError_t launch_rows_enter(float* x, float* backup, const float* b_c, const int* plan_img, int M, int HW, int C, bool attn,
                              cudaStream_t st) {
  if (C % 4) return cudaErrorNotSupported;
  launch_k((rows_enter_kernel), grid_for((long long)M * C / 4, 256), 256, 0, st, x, backup, b_c, plan_img, M, HW, C, attn ? 1 : 0);
  return cudaGetLastError();
}
cudaError_t launch_rows_leave(float* x, const float* backup, const int* plan_img, int M, int HW, int C, cudaStream_t st) {
  if (C % 4) return cudaErrorNotSupported;
  launch_k((rows_leave_kernel), grid_for((long long)M * C / 4, 256), 256, 0, st, x, backup, plan_img, M, HW, C);
  return cudaGetLastError();
}
