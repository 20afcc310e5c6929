// Window attention core on tcgen05 / TMEM (sm_100a): S = q k^T and O = P v as UMMA tiles, fp32 softmax from TMEM.
//
// attention.py:13-85 + torch MHA as called there (attention.py:82): per (image, window, head)
//   S = (q / sqrt(32)) k^T + key bias,  P = softmax_keys(S),  O = P v.
// A window has <= 36 tokens and a head 32 channels, so one (window, head) problem is far below a 128-row UMMA tile.
// Instead of one problem per tile, a tile packs 128 / SLOT windows, each in a SLOT-row slot (SLOT = 64 for the 6x6
// windows, 32 for windows of <= 4 rows), token (y, x) of a window at slot row 8 y + x, and ONE M128 x N128 x K32 UMMA
// computes every query x key product of the tile; the softmax only reads its own window's diagonal block of the
// accumulator, and P is written into a [128 x 128] bf16 operand that is zero outside the diagonal blocks, so
// O = P v (M128 x N32 x K128) never mixes windows.  The wasted off-diagonal MACs are free (the kernel is bound by moving
// q/k/v and by the softmax, not by the tensor pipe).  Slot rows that hold no token (x >= window width, y >= window height)
// carry a -inf key bias: their k / v rows only have to be finite.
// The 8-rows-per-window-row layout is what lets TMA do the gather: one 4-d box (64 channels, 8 pixels, window height, 1 image)
// per (window, q|k|v) lands exactly on the window's slot (the 2 pixels beyond the window are the dead rows; out-of-image
// pixels are zero-filled = the reference's zero padding).  Windows of a shifted block that wrap around the frame or contain
// pad tokens (which carry the in-projection bias there) are gathered with 16-byte cp.async instead.
//
// Work item = (tile, pair of heads) = 64 channels = 128-byte rows.  Warp roles (13 warps):
//   warps 0-7   softmax / epilogue: two groups of 4 warps (TMEM lane quadrant = warp % 4), group g owns head g of the
//               pair: tcgen05.ld of its window's S block -> scale, key bias, row softmax in fp32 -> P (bf16, 128B-swizzled
//               smem) -> later O from TMEM, normalised, 64-byte store per token (pad tokens are cropped, attention.py:56)
//   warps 8-11  loaders: thread = tile row = one token slot for the index arithmetic that replaces the pad / roll /
//               window-split copies of the reference (attention.py:27-50); q | k | v into 128B-swizzled K-major panels
//               (3 stages) by TMA box loads or cp.async; pad tokens: k = v = 0 (shift == 0, masked by the -inf key bias) or
//               the in-projection bias (shift != 0: x = 0 there, and the float "mask" keeps those keys live)
//   warp 12     MMA issuer (whole warp runs the loop, one elected lane issues): S for item j, then P v for item j - 1
// V is the B operand of P v in MN-major form ([key][d] rows as loaded, no transpose): instruction descriptor bit 16.
#include <cuda.h>
#include <math.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tc_context.h"

namespace {

constexpr int kD = 32;                       // head_dim (unet.py:26)
constexpr int kStages = 3;
constexpr int kPanel = 128 * 128;            // [128 rows x 64 bf16] = 16 KB, 128-byte rows
constexpr int kStageBytes = 3 * kPanel;      // q | k | v panels of one item
constexpr int kPBytes = 2 * kPanel;          // P [128 x 128 keys] bf16 = two K-panels
constexpr int kSoftWarps = 8, kLoadWarps = 4;
constexpr int kMmaWarp = kSoftWarps + kLoadWarps;
constexpr int kThreads = 32 * (kMmaWarp + 1);
constexpr int kMetaBytes = kStages * 128 * 12;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kPBytes + kMetaBytes + 256;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnGeom {
  int B, H, W, C, wh, ww, shift, Hp, Wp;
  int wpt;                    // windows per 128-row tile
  int box_bytes;              // bytes one TMA box (64 channels x 8 pixels x wh rows) delivers
  int nww, nwin, n_windows, n_pairs;
  int n_units, ppu, n_groups;     // work units = tiles x n_groups, each ppu = n_pairs / n_groups consecutive head pairs
  long long ldo;
  int dbg;   // debug experiments: 1 no softmax arithmetic, 2 no P v MMAs, 4 no S MMAs, 8 no q/k/v copies, 16 no output stores, 32 no P writes
};

__device__ __noinline__ bool wait_bar_slow(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* fault, int code);
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* fault, int code) {
  if (ptx::mbar_try_wait(bar, parity)) return true;
  return wait_bar_slow(bar, parity, s_abort, fault, code);
}
__device__ __noinline__ bool wait_bar_slow(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* fault, int code) {
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

__device__ __forceinline__ float ex2_approx(float x) {     // 2^x, one MUFU op; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// SLOT = rows of a tile reserved per window: 64 (32 < L <= 36) or 32 (L <= 32)
// WWC: compile-time bound on the window width (8 = any): columns x >= WWC of every window row are dead and skipped
template <int SLOT, int WWC>
__global__ void __launch_bounds__(kThreads, 1)
window_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const bf16* __restrict__ qkv, const bf16* __restrict__ xm, const float* __restrict__ b_in,
                           bf16* __restrict__ att, const AttnGeom g, const int* __restrict__ skip, int* fault, long long* trace) {
  constexpr int NV = SLOT == 64 ? 48 : 32;            // accumulator columns a softmax thread reads: 8 per window row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* pbuf = stages + kStages * kStageBytes;
  float* meta_kb = reinterpret_cast<float*>(pbuf + 2 * kPBytes);     // [stage][128] key bias * log2(e)
  int* meta_m = reinterpret_cast<int*>(meta_kb + kStages * 128);     // [stage][128] token index of the row, -1: not a stored query
  int* meta_c = meta_m + kStages * 128;                              // [stage][128] copy code of the row (loaders)
  uint64_t* full = reinterpret_cast<uint64_t*>(meta_c + kStages * 128);
  uint64_t* empty = full + kStages;
  uint64_t* s_full = empty + kStages;
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* pv_done = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  volatile int* s_abort = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // debug (ldmb_debug_tc_trace): cycles one lane of each role spends per phase, 16 slots per CTA
  // (compiled in with -DLDMB_ATTN_TRACE only: the accounting costs registers and ~10 KB of code)
#ifdef LDMB_ATTN_TRACE
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  auto lap = [&](int slot) { if (trace != nullptr) { const long long now = clock64(); tacc[slot] += now - tlast; tlast = now; } };
#else
  const long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto lap = [](int) {};
  (void)trace;
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&full[i], 2 * 32 * kLoadWarps); ptx::mbar_init(&empty[i], 1 + kSoftWarps); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&s_empty[i], 4);
      ptx::mbar_init(&p_full[i], 4); ptx::mbar_init(&pv_done[i], 1);
    }
    *s_abort = 0;
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmQ);
  }
  if (warp == kMmaWarp) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  // Zero every operand buffer once: rows no loader ever writes (slot rows >= L) and P outside the diagonal blocks must
  // read as 0 -- a NaN bit pattern there would reach every output through 0 * NaN.
  for (int i = threadIdx.x; i < (kStages * kStageBytes + 2 * kPBytes) / 16; i += kThreads)
    ptx::st_shared_v4(ptx::smem_u32(stages) + i * 16, 0u, 0u, 0u, 0u);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const bool skipped = skip != nullptr && *skip != 0;              // stochastic depth (unet.py:39-40)
  // work unit = (tile, group of ppu head pairs): the tile's index arithmetic is shared by the unit's ppu items
  const int n_units = skipped ? 0 : g.n_units;
  const int n_units_my = (int)blockIdx.x < n_units ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n_my = n_units_my * g.ppu;

  if (warp >= kSoftWarps && warp < kMmaWarp) {
    // ===================================================== loaders
    // Two thread mappings: (1) thread = tile row = token for the per-tile index arithmetic (published in meta_m / meta_kb),
    // (2) thread = 16-byte chunk c of rows r0 + 16 i for the copies, so that a warp instruction moves 4 whole 128-byte rows.
    const int t = threadIdx.x - 32 * kSoftWarps;
    const int w = t / SLOT, sl = t % SLOT;
    const int si = sl >> 3, sj = sl & 7;                   // token (y, x) of the window this slot row holds
    const bool in_window = si < g.wh && sj < g.ww;
    const int c = t & 7, r0 = t >> 3;
    const uint32_t dst_off = static_cast<uint32_t>(r0 * 128 + ((c ^ (r0 & 7)) << 4));     // (r0 + 16 i) & 7 == r0 & 7
    // row state of the unit's tile.  mq: token index of the row as a query (-1: no output).  code, for the cp.async path:
    // >= 0 token index (q, k, v copied); -1: k, v zero-filled; -2: k, v = in-projection bias; -3: nothing to write (dead slot
    // row, or the window is loaded by TMA).  tma: this thread issues its window's three box loads (slot row 0 of a TMA window).
    int code = -3, mq = -1, tma_b = 0, tma_pi = 0, tma_pj = 0;
    bool tma = false;
    float kbv = 0.f;
    // The key bias of a shifted window is a gather from xm (one dependent global load per row and tile).  Issued right after a
    // burst of cp.async it would queue behind them in the SM's load/store unit, so it is requested one unit AHEAD (kb_next),
    // before the current unit's copies, and only converted when that unit starts.
    int code_next = -3, mq_next = -1, tma_b_next = 0, tma_pi_next = 0, tma_pj_next = 0;
    bool tma_next = false;
    float kbv_next = 0.f;
    bf16 kb_raw_next = __float2bfloat16_rn(0.f);
    bool kb_from_xm_next = false;
    auto tile_meta_issue = [&](int tile) {
      const int gw = tile * g.wpt + w;                      // window index over the batch
      code_next = -3; mq_next = -1; kbv_next = -INFINITY; kb_from_xm_next = false; tma_next = false;
      if (gw >= g.n_windows) { if (in_window) code_next = -1; return; }      // beyond the batch: keys must still be finite
      const int b = gw / g.nwin, win = gw - b * g.nwin, wi = win / g.nww, wj = win - wi * g.nww;
      // origin of the window in the zero-padded frame BEFORE the roll (attention.py:39): rolled position r holds frame position r - shift
      const int oi = wi * g.wh - g.shift, oj = wj * g.ww - g.shift;
      // TMA box: the window is a rectangle of the frame (no wrap-around) and none of its tokens is a pad token of a shifted
      // block (those need the in-projection bias, TMA zero-fills)
      const bool box_ok = g.shift == 0 || (oi >= 0 && oj >= 0 && oi + g.wh <= g.H && oj + g.ww <= g.W);
      if (box_ok && sl == 0) { tma_next = true; tma_b_next = b; tma_pi_next = oi; tma_pj_next = oj; }
      if (!in_window) return;
      int pi = oi + si, pj = oj + sj;
      if (pi < 0) pi += g.Hp;
      if (pj < 0) pj += g.Wp;
      const bool pad = pi >= g.H || pj >= g.W;
      if (!pad) mq_next = (b * g.H + pi) * g.W + pj;
      if (!box_ok) code_next = pad ? -1 : mq_next;
      if (g.shift == 0) {
        kbv_next = pad ? -INFINITY : 0.f;                   // bool key_padding_mask (attention.py:27-35)
      } else {                                              // float "mask" = rolled activation, channel 0 (attention.py:40)
        int qi = pi - g.shift, qj = pj - g.shift;
        if (qi < 0) qi += g.Hp;
        if (qj < 0) qj += g.Wp;
        kbv_next = 0.f;
        if (qi < g.H && qj < g.W) { kb_raw_next = xm[(((long long)b * g.H + qi) * g.W + qj) * g.C]; kb_from_xm_next = true; }
        if (pad) code_next = -2;                            // x = 0 there: k, v = in-projection bias, and the key stays live
      }
    };
    auto tile_meta_take = [&]() {
      code = code_next; mq = mq_next; tma = tma_next; tma_b = tma_b_next; tma_pi = tma_pi_next; tma_pj = tma_pj_next;
      kbv = kb_from_xm_next ? __bfloat162float(kb_raw_next) * kLog2e : kbv_next;
    };
    int it = 0;
    bool ok = true;
    if (n_units_my > 0) tile_meta_issue((int)blockIdx.x / g.n_groups);
    for (int k = 0; ok && k < n_units_my; ++k) {
      const int u = (int)blockIdx.x + k * (int)gridDim.x;
      const int grp = u % g.n_groups;
      tile_meta_take();
      if (k + 1 < n_units_my) tile_meta_issue((u + (int)gridDim.x) / g.n_groups);      // next unit's gather, ahead of this unit's copies
      for (int pp = 0; pp < g.ppu; ++pp, ++it) {
        const int pair = grp * g.ppu + pp, s = it % kStages;
        lap(0);
        if (!wait_bar(&empty[s], ((it / kStages) & 1) ^ 1, s_abort, fault, 41)) { ok = false; break; }
        lap(1);
        meta_kb[s * 128 + t] = kbv;
        meta_m[s * 128 + t] = mq;
        meta_c[s * 128 + t] = code;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kLoadWarps) : "memory");
        const uint32_t base = ptx::smem_u32(stages + s * kStageBytes) + dst_off;
        const bf16* src0 = qkv + pair * 64 + c * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int cd = meta_c[s * 128 + r0 + 16 * i];
          const uint32_t dst = base + i * (16 * 128);
          if (cd >= 0) {
            if (!(g.dbg & 8)) {
              const bf16* src = src0 + (long long)cd * (3 * g.C);
              ptx::cp_async_16(dst, src);
              ptx::cp_async_16(dst + kPanel, src + g.C);
              ptx::cp_async_16(dst + 2 * kPanel, src + 2 * g.C);
            }
          } else if (cd >= -2) {
#pragma unroll
            for (int which = 1; which < 3; ++which) {
              uint32_t v0 = 0u, v1 = 0u, v2 = 0u, v3 = 0u;
              if (cd == -2) {
                const float* bp = b_in + (long long)which * g.C + pair * 64 + c * 8;
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(bp)), b4 = __ldg(reinterpret_cast<const float4*>(bp + 4));
                v0 = pack_bf16(a4.x, a4.y); v1 = pack_bf16(a4.z, a4.w); v2 = pack_bf16(b4.x, b4.y); v3 = pack_bf16(b4.z, b4.w);
              }
              ptx::st_shared_v4(dst + which * kPanel, v0, v1, v2, v3);
            }
          }
        }
        lap(2);
        // publish: a plain arrive releases this thread's metadata / fill stores, the cp.async arrive fires when its copies
        // have landed -- the loader never waits for its own copies and runs up to kStages items ahead
        ptx::fence_proxy_async();
        if (tma && !(g.dbg & 8)) {                // this window's q | k | v: three box loads straight onto its slot
          ptx::mbar_arrive_expect_tx(&full[s], 3 * g.box_bytes);
          uint8_t* dstw = stages + s * kStageBytes + w * (SLOT * 128);
#pragma unroll
          for (int which = 0; which < 3; ++which)
            ptx::tma_load_4d(dstw + which * kPanel, &tmQ, &full[s], which * g.C + pair * 64, tma_pj, tma_pi, tma_b);
        } else {
          ptx::mbar_arrive(&full[s]);
        }
        ptx::cp_async_mbar_arrive_noinc(&full[s]);
        lap(3);
      }
    }
    if (trace != nullptr && t == 0) for (int i = 0; i < 4; ++i) trace[blockIdx.x * 16 + 12 + i] = tacc[i];
  } else if (warp == kMmaWarp) {
    // ===================================================== MMA issuer
    const bool issuer = ptx::elect_one();
    constexpr uint32_t idesc_s = ptx::idesc_bf16(128, 128);
    constexpr uint32_t idesc_pv = ptx::idesc_bf16(128, kD) | (1u << 16);      // B (= v) MN-major: rows = keys, 32 d contiguous
    bool ok = true;
    for (int j = 0; ok && j <= n_my; ++j) {
      if (j < n_my) {           // S = q k^T of item j, both heads
        const int s = j % kStages;
        lap(0);
        if (!wait_bar(&full[s], (j / kStages) & 1, s_abort, fault, 42)) break;
        lap(1);
        ptx::fence_proxy_async();      // cp.async-written operands (generic proxy) -> the tensor core's async-proxy reads
        ptx::tc_fence_after();
        const uint64_t q_desc = ptx::smem_desc_sw128(ptx::smem_u32(stages + s * kStageBytes));
        const uint64_t k_desc = q_desc + kPanel / 16;
        for (int h = 0; h < 2; ++h) {
          lap(0);
          if (!wait_bar(&s_empty[h], (j & 1) ^ 1, s_abort, fault, 43)) { ok = false; break; }
          lap(2);
          ptx::tc_fence_after();
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 2 && !(g.dbg & 4); ++k)      // head h = bytes [64 h, 64 h + 64) of the 128-byte rows, K = 32 = 2 x 16
              ptx::umma_f16(tmem_base + h * 128, q_desc + 4 * h + 2 * k, k_desc + 4 * h + 2 * k, idesc_s, k);
            ptx::umma_commit(&s_full[h]);
          }
          __syncwarp();
        }
        if (!ok) break;
      }
      if (j >= 1) {             // O = P v of item j - 1
        const int i = j - 1, s = i % kStages;
        const uint64_t v_desc = ptx::smem_desc_sw128(ptx::smem_u32(stages + s * kStageBytes + 2 * kPanel));
        for (int h = 0; h < 2; ++h) {
          lap(0);
          if (!wait_bar(&p_full[h], i & 1, s_abort, fault, 44)) { ok = false; break; }
          lap(3);
          ptx::tc_fence_after();
          if (issuer) {
            const uint64_t p_desc = ptx::smem_desc_sw128(ptx::smem_u32(pbuf + h * kPBytes));
            uint32_t acc = 0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {           // 16 keys per step
              if ((kk * 16) % SLOT >= 8 * g.wh || (g.dbg & 2)) continue;   // both window rows of this step lie beyond the window: P is 0 there
              ptx::umma_f16(tmem_base + 256 + h * kD, p_desc + (kk / 4) * (kPanel / 16) + (kk % 4) * 2,
                            v_desc + kk * (16 * 128 / 16) + 4 * h, idesc_pv, acc);
              acc = 1;
            }
            ptx::umma_commit(&pv_done[h]);
          }
          __syncwarp();
        }
        if (!ok) break;
        if (issuer) ptx::umma_commit(&empty[s]);       // q, k, v of the stage are consumed once these MMAs complete
        __syncwarp();
      }
    }
    lap(0);
    if (trace != nullptr && lane == 0) for (int i = 0; i < 4; ++i) trace[blockIdx.x * 16 + 8 + i] = tacc[i];
  } else {
    // ===================================================== softmax / epilogue: group = head of the pair
    const int grp = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const int w = row / SLOT;                                   // warp-uniform (SLOT is a multiple of 32)
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const float scale_l2 = 0.17677669529663687f * kLog2e;       // 1/sqrt(32) * log2(e)
    // P row of this thread: key columns [w SLOT, w SLOT + 8 wh) -> K-panel (w SLOT) / 64, 16-byte chunks (= window rows) from ((w SLOT) % 64) / 8
    const uint32_t p_row = ptx::smem_u32(pbuf + grp * kPBytes) + ((w * SLOT) / 64) * kPanel + row * 128;
    const int chunk0 = ((w * SLOT) % 64) / 8;
    int m_prev = -1, pair_prev = 0, pair_cur = 0;
    float inv_prev = 0.f;
    bool ok = true;
    auto epilogue = [&](int i) -> bool {                        // O of item i (this group's head) -> att
      lap(0);
      if (!wait_bar(&pv_done[grp], i & 1, s_abort, fault, 45)) return false;
      lap(4);
      ptx::tc_fence_after();
      uint32_t o[32];
      ptx::tmem_ld_32x32(tmem_base + lane_off + 256 + grp * kD, o);
      ptx::tmem_ld_wait();
      if (m_prev >= 0 && !(g.dbg & 16)) {
        uint4* dst = reinterpret_cast<uint4*>(att + (long long)m_prev * g.ldo + pair_prev * 64 + grp * kD);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          dst[u] = make_uint4(pack_bf16(__uint_as_float(o[8 * u]) * inv_prev, __uint_as_float(o[8 * u + 1]) * inv_prev),
                              pack_bf16(__uint_as_float(o[8 * u + 2]) * inv_prev, __uint_as_float(o[8 * u + 3]) * inv_prev),
                              pack_bf16(__uint_as_float(o[8 * u + 4]) * inv_prev, __uint_as_float(o[8 * u + 5]) * inv_prev),
                              pack_bf16(__uint_as_float(o[8 * u + 6]) * inv_prev, __uint_as_float(o[8 * u + 7]) * inv_prev));
      }
      return true;
    };
    for (int it = 0, pp = 0, k = 0; it < n_my; ++it) {
      const int s = it % kStages;
      if (pp == 0) pair_cur = (((int)blockIdx.x + k * (int)gridDim.x) % g.n_groups) * g.ppu;     // first pair of the unit
      else ++pair_cur;
      if (++pp == g.ppu) { pp = 0; ++k; }
      lap(0);
      if (!wait_bar(&full[s], (it / kStages) & 1, s_abort, fault, 46)) { ok = false; break; }
      lap(1);
      const int m_cur = meta_m[s * 128 + row];
      if (!wait_bar(&s_full[grp], it & 1, s_abort, fault, 47)) { ok = false; break; }
      lap(2);
      ptx::tc_fence_after();
      float v[NV];
      {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + lane_off + grp * 128 + w * SLOT, r);
        if (SLOT == 64) {
          uint32_t r16[16];
          ptx::tmem_ld_32x16(tmem_base + lane_off + grp * 128 + w * SLOT + 32, r16);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[(32 + i) % NV] = __uint_as_float(r16[i]);
        } else {
          ptx::tmem_ld_wait();
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[grp]);           // the next item's S may overwrite the accumulator
      lap(3);
      // ---- softmax over the window's keys (fp32, exp2 domain)
      const uint32_t kbp = ptx::smem_u32(meta_kb + s * 128 + w * SLOT);
      float mx = -INFINITY;
      float sum = 0.f;
      if (!(g.dbg & 1)) {
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i4 = 0; i4 < NV / 4; ++i4) {               // key biases: 16-byte broadcast loads
          const float4 kb4 = ptx::ld_shared_v4(kbp + i4 * 16);
          const float kb[4] = {kb4.x, kb4.y, kb4.z, kb4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i4 * 4 + u;
            if ((i & 7) >= WWC) continue;
            v[i] = fmaf(v[i], scale_l2, kb[u]);              // dead slot rows and masked keys: kb = -inf
            mx4[u] = fmaxf(mx4[u], v[i]);
          }
        }
        mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if (mx == -INFINITY) mx = 0.f;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if ((i & 7) >= WWC) { v[i] = 0.f; continue; }
          v[i] = ex2_approx(v[i] - mx);
          s4[i & 3] += v[i];
        }
        sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      const float inv = 1.f / sum;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[s]);               // this stage's metadata has been read
      // ---- previous item of this head: P v has completed (P buffer free), its O leaves through this warp
      lap(5);
      if (it > 0 && !epilogue(it - 1)) { ok = false; break; }
      lap(6);
      if (m_cur >= 0 && !(g.dbg & 32)) {
#pragma unroll
        for (int c = 0; c < NV / 8; ++c) {                  // one 16-byte chunk per window row (dead keys: exp2(-inf) = 0)
          if (c >= g.wh) break;
          ptx::st_shared_v4(p_row + (((chunk0 + c) ^ sw) << 4), pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                            pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
        }
      }
      ptx::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      ptx::tc_fence_before();            // orders the O read above before the MMA that overwrites the accumulator
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[grp]);
      lap(7);
      m_prev = m_cur;
      pair_prev = pair_cur;
      inv_prev = inv;
    }
    if (ok && n_my > 0) epilogue(n_my - 1);
    if (trace != nullptr && warp == 0 && lane == 0) for (int i = 0; i < 8; ++i) trace[blockIdx.x * 16 + i] = tacc[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

static const bool g_attn_tc = getenv("LDMB_ATTN_TC") == nullptr || atoi(getenv("LDMB_ATTN_TC")) != 0;   // debug: 0 = mma.sync kernel
static const int g_attn_tc_min_hw = getenv("LDMB_ATTN_TC_MIN_HW") ? atoi(getenv("LDMB_ATTN_TC_MIN_HW")) : 0;   // experiment: tcgen05 only for feature maps of at least this many pixels

bool window_attention_tc_supported(int B, int H, int W, int C, int head_dim, int win_h, int win_w, long long ldo) {
  const long long tokens = (long long)B * H * W;
  return g_attn_tc && H * W >= g_attn_tc_min_hw && head_dim == kD && C % 64 == 0 && win_h >= 1 && win_h <= 6 && win_w >= 1 && win_w <= 8 && ldo % 8 == 0 &&
         tokens < (1LL << 31);
}

cudaError_t launch_window_attention_tc(TcContext* ctx, const void* qkv, const void* xm, const float* b_in, void* att, long long ldo,
                                       int B, int H, int W, int C, int win_h, int win_w, int shift, const int* skip,
                                       cudaStream_t st) {
  if (!window_attention_tc_supported(B, H, W, C, kD, win_h, win_w, ldo)) return cudaErrorNotSupported;
  AttnGeom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.H = H; g.W = W; g.C = C; g.wh = win_h; g.ww = win_w; g.shift = shift; g.ldo = ldo;
  g.dbg = tc_knobs().attn_dbg;
  g.box_bytes = 64 * 2 * 8 * win_h;
  g.Hp = (H + win_h - 1) / win_h * win_h; g.Wp = (W + win_w - 1) / win_w * win_w;
  const int slot = win_h <= 4 ? 32 : 64;                 // slot rows per window: 8 per window row
  g.wpt = 128 / slot;
  g.nww = g.Wp / win_w; g.nwin = (g.Hp / win_h) * g.nww;
  const long long n_windows = (long long)B * g.nwin;
  const long long n_tiles = (n_windows + g.wpt - 1) / g.wpt;
  g.n_pairs = C / 64;
  if (n_windows >= (1LL << 31) || n_tiles * g.n_pairs >= (1LL << 31)) return cudaErrorNotSupported;
  g.n_windows = (int)n_windows;
  // pairs per work unit: whole waves over the SMs first, then as many pairs per unit as possible (the tile's index
  // arithmetic and key-bias gather are per unit)
  long long best_cost = -1;
  for (int ppu = 1; ppu <= g.n_pairs; ++ppu) {
    if (g.n_pairs % ppu) continue;
    const long long units = n_tiles * (g.n_pairs / ppu);
    const long long waves = (units + ctx->num_sms - 1) / ctx->num_sms;
    const long long cost = waves * (4 * ppu + 1);          // ~ a quarter of an item's time per unit for the index arithmetic
    if (best_cost < 0 || cost <= best_cost) { best_cost = cost; g.ppu = ppu; }
  }
  g.n_groups = g.n_pairs / g.ppu;
  g.n_units = (int)(n_tiles * g.n_groups);
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc_kernel<64, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attention_tc_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attention_tc_kernel<64, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attention_tc_kernel<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr.mark(ctx->device);
  }
  const int grid = g.n_units < ctx->num_sms ? g.n_units : ctx->num_sms;
  // qkv [B, H, W, 3C] bf16 as a 4-d tensor (channel, x, y, image); box = 64 channels x 8 pixels x win_h rows of one image
  CUtensorMap tmQ;
  {
    const cuuint64_t gdim[4] = {(cuuint64_t)(3 * C), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)(3 * C) * 2, (cuuint64_t)(3 * C) * 2 * W, (cuuint64_t)(3 * C) * 2 * W * H};
    const cuuint32_t box[4] = {64, 8, (cuuint32_t)win_h, 1};
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    if (ctx->encode(&tmQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(qkv), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const bf16* q = static_cast<const bf16*>(qkv);
  const bf16* x = static_cast<const bf16*>(xm);
  bf16* o = static_cast<bf16*>(att);
  if (slot == 64) {
    if (win_w <= 6) return launch_k((window_attention_tc_kernel<64, 6>), dim3(grid), dim3(kThreads), kSmemBytes, st, tmQ, q, x, b_in, o, g, skip, ctx->fault_dev, ctx->trace_dev);
    return launch_k((window_attention_tc_kernel<64, 8>), dim3(grid), dim3(kThreads), kSmemBytes, st, tmQ, q, x, b_in, o, g, skip, ctx->fault_dev, ctx->trace_dev);
  }
  if (win_w <= 4) return launch_k((window_attention_tc_kernel<32, 4>), dim3(grid), dim3(kThreads), kSmemBytes, st, tmQ, q, x, b_in, o, g, skip, ctx->fault_dev, ctx->trace_dev);
  return launch_k((window_attention_tc_kernel<32, 8>), dim3(grid), dim3(kThreads), kSmemBytes, st, tmQ, q, x, b_in, o, g, skip, ctx->fault_dev, ctx->trace_dev);
}
