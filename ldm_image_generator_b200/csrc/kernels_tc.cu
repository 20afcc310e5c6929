// tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// One persistent CTA per SM, 6 warps:
//   warp 0     TMA producer  (one lane): A tile 128x64 + B tile BNx64 per stage, 128B-swizzled, mbarrier tx-count
//   warp 1     MMA issuer    (one lane): 4 x tcgen05.mma (M128, N=BN, K16) per stage into one of two TMEM accumulators
//   warps 2-9  epilogue: tcgen05.ld 32 lanes x 32 columns -> bias / activation / ReGLU gate -> 128B-swizzled smem slab
//              (32 rows x 128 B per warp, double buffered) -> TMA store, or TMA fp32 reduce-add into the residual
//              stream (x += ... without reading x).  A direct register->global path remains for the ConvTranspose
//              scatter and for epilogues that add a second tensor.
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty pair (MMA <-> epilogue), so the epilogue of
// tile i overlaps the main loop of tile i+1.
//
// A operand modes: AM_ROWS  -- plain [M,K] row-major matrix (1x1 convolutions on NHWC activations)
//                  AM_CONV3 -- NHWC tensor seen through a 4-d tensor map; the 9 taps are 9 shifted box loads whose
//                              out-of-bounds elements TMA zero-fills (= the convolution's zero padding, vae.py:57-58)
// Every mbarrier wait carries a watchdog: a stuck pipeline sets a fault flag and drains instead of hanging the GPU.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>

#include "kernels.h"
#include "ptx.cuh"

#include "tc_context.h"
constexpr int kTraceSlots = 16;

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8;                       // two per TMEM lane quadrant, each takes half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;

struct TcTiling {
  int m_tiles, n_tiles, num_kb, total;
  int TW, TH, TB;   // AM_CONV3: the 128-pixel tile is TB images x TH rows x TW columns
  int tma_out;      // epilogue goes through smem + TMA store / reduce
  int max_stages;       // debug (LDMB_TC_STAGES): use only this many of the smem pipeline stages
  int dbg;              // debug (LDMB_TC_DBG): 1 = the epilogue drains nothing (accumulators are released unread), 2 = no MMAs,
                        // 4 = staged epilogue without its TMA stores / reduces, 8 = no per-tile bias staging + CTA barrier
  int splits, kb_per;   // split-K (EPI_ACCUM_F32 through TMA reduce-add only): tile t covers k-blocks [sp*kb_per, ...)
  int out_col_b, out_row_b;   // per-batch (grid z) column / row offset of the output tile in the out tensor map
  int ast_ppm, ast_tpp;       // A-stationary mode: CTA pairs per m-tile, consecutive n-tiles per pair
  int ast_persist;            // A-stationary, persistent: pair p walks the m-tiles p, p + P, ... and takes ALL n-tiles of each (A loaded once per m-tile)
  int ast_nbuf;               //   A buffers (an m-tile's k-block tiles each): the next m-tile's A loads while this one's n-tiles run
  int ast_bres;               //   B resident: every (n-tile, k-block) tile of B sits in its own ring slot for the whole kernel
};

// CG = CTAs per tile: 1 (128 x BN per CTA) or 2 (a CTA pair computes 256 x BN with cta_group::2; each CTA stages its
// own 128 rows of A and BN/2 rows of B, so the pair moves 2/3 of the L2->SM bytes per FLOP of two independent CTAs).
// AST (A-stationary; K <= 512, CTA pairs, 256-wide tiles): a pair keeps its 256 x K block of A resident in shared memory for
// several consecutive n-tiles and streams only B through the ring -- the main loop of these GEMMs is bound by the bytes an SM
// ingests from L2 (A 16 KB + B 16 KB per k-block otherwise), not by the tensor pipe.
constexpr int kAstMaxKb = 8;
template <int BN, int CG, bool AST = false> struct TcCfg;
template <int BN, int CG> struct TcCfg<BN, CG, true> {
  static constexpr int B_ROWS = BN / CG;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = B_BYTES;                // the ring holds B tiles only
  static constexpr int A_RES_BYTES = kAstMaxKb * A_BYTES;    // resident A: up to 8 k-block tiles of [128 x 128 B]
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SLAB_BYTES = 32 * 128;
  static constexpr int SLABS = 1;                            // one staging slab per epilogue warp (bf16 epilogues: one slab per tile and warp)
  static constexpr int STAGING_BYTES = kEpiWarps * SLABS * SLAB_BYTES;
  static constexpr int BIAS_FLOATS = BN == 256 ? 3072 : (BN == 128 ? 448 : 2 * BN);   // >= 2 * BN; sized so that no pipeline stage is lost
  static constexpr int FIXED_BYTES = 1024 + A_RES_BYTES + STAGING_BYTES + BAR_BYTES + BIAS_FLOATS * 4;
  static constexpr int STAGES_FIT = (232448 - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int SMEM_BYTES = FIXED_BYTES + STAGES * STAGE_BYTES;
  static_assert(STAGES >= 3, "B ring");
};
template <int BN, int CG> struct TcCfg<BN, CG, false> {
  static constexpr int B_ROWS = BN / CG;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;

  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SLAB_BYTES = 32 * 128;               // one warp's 32 rows x 128 B
  static constexpr int SLABS = 2;
  static constexpr int A_RES_BYTES = 0;
  static constexpr int STAGING_BYTES = kEpiWarps * SLABS * SLAB_BYTES;   // per epilogue warp, double buffered
  // bias staging: the whole launch's bias vector when it fits (N <= BIAS_FLOATS, one K slice, no batch), else two tiles' worth
  static constexpr int BIAS_FLOATS = BN == 256 ? 3072 : (BN == 128 ? 448 : 2 * BN);   // >= 2 * BN; sized so that no pipeline stage is lost
  static constexpr int FIXED_BYTES = 1024 + STAGING_BYTES + BAR_BYTES + BIAS_FLOATS * 4;
  static constexpr int STAGES_FIT = (232448 - FIXED_BYTES) / STAGE_BYTES;       // 227 KB of dynamic smem per CTA
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int SMEM_BYTES = FIXED_BYTES + STAGES * STAGE_BYTES;
};

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* fault, int code) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (true) {
    if (ptx::mbar_try_wait(bar, parity)) return;
    if (*s_abort) return;
    if (clock64() - t0 > 3000000000LL) {   // ~1.5 s at 2 GHz: pipeline is stuck, drain instead of hanging
      *s_abort = 1;
      report_fault(fault, code);
      return;
    }
  }
}

__device__ __forceinline__ void trace_stamp(long long* trace, int slot) {
  if (trace != nullptr) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * kTraceSlots + slot] = t;
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// packed fp32 pairs (sm_100 f32x2 pipe): two adds / multiplies per instruction in the issue-bound gate epilogue
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&r);
}

template <int BN, int AMODE, int CG, bool AST>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ GemmDesc d, const TcTiling tl, int* fault, long long* trace) {
  using Cfg = TcCfg<BN, CG, AST>;
  constexpr int STAGES_MAX = Cfg::STAGES;
  const int STAGES = tl.max_stages > 0 && tl.max_stages < STAGES_MAX ? tl.max_stages : STAGES_MAX;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;   // position in the CTA pair
  const bool leader = rank == 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* a_res = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));   // AST: resident A k-block tiles
  uint8_t* tiles = a_res + Cfg::A_RES_BYTES;
  uint8_t* staging = tiles + STAGES_MAX * Cfg::STAGE_BYTES;      // 1024-aligned (every stage is a multiple of 1024 B)
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* empty = full + STAGES_MAX;
  uint64_t* tfull = empty + STAGES_MAX;
  uint64_t* tempty = tfull + 2;
  uint64_t* a_full = tempty + 2;                                 // [kAstMaxKb] (AST): (buffer, k-block)
  uint64_t* a_empty = a_full + kAstMaxKb;                        // [2] (AST, persistent): the m-tile's last MMA has read the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 2);
  volatile int* s_abort = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-role cycle accounting (compiled in with -DLDMB_TC_TRACE only; tools/trace_gemm_roles.py reads it through ldmb_debug_tc_trace)
#ifdef LDMB_TC_TRACE
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  auto lap = [&](int slot) { if (trace != nullptr) { const long long now = clock64(); tacc[slot] += now - tlast; tlast = now; } };
#else
  auto lap = [](int) {};
#endif

  if (threadIdx.x == 0) trace_stamp(trace, 0);                 // kernel entry
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], kEpiWarps * CG); }
    if (AST) { for (int i = 0; i < kAstMaxKb; ++i) ptx::mbar_init(&a_full[i], 1); for (int i = 0; i < 2; ++i) ptx::mbar_init(&a_empty[i], 1); }
    *s_abort = 0;
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (tl.tma_out) ptx::prefetch_tensormap(&tmO);
  }
  if (warp == 1) {
    if (CG == 2) { ptx::tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS); ptx::tmem_relinquish_2sm(); }
    else { ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();      // the peer's barriers are initialised before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_z = tl.m_tiles * tl.n_tiles;
  const int tiles_per_split = tl.total / tl.splits;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail.
  if (threadIdx.x == 0) trace_stamp(trace, 1);                 // setup done
  // The step's plan entry was uploaded before the first kernel of the step (an ancestor of every kernel in the
  // stream / graph), so it is read BEFORE waiting on the previous kernel, like the weights (see the producer).
  // expert rows of the (up to 4) selection slots, resolved from the device-side plan when there is one
  int srow0 = d.sel_rows[0], srow1 = d.sel_rows[1], srow2 = d.sel_rows[2], srow3 = d.sel_rows[3];
  bool skip_block = false;
  if (d.plan != nullptr) {
    skip_block = d.plan[0] != 0;   // stochastic depth skipped this block for this step (uniform over the grid)
    if (d.sel == 1 || d.sel == 2) { srow0 = 0; srow1 = (1 + d.plan[1]) * d.sel_stride; srow2 = (1 + d.plan[2]) * d.sel_stride; srow3 = 5 * d.sel_stride; }
  }
  const bool sel_k = d.sel == 2 || d.sel == 3;        // weight rows selected per K slot
  auto srow = [&](int q) { return d.sel == 3 ? q * d.sel_stride : (q == 0 ? srow0 : (q == 1 ? srow1 : (q == 2 ? srow2 : srow3))); };
  const int total_tiles = skip_block ? 0 : tl.total;
  // this CTA (pair)'s tiles: t_first + i * t_step, i < t_count.  Default: strided over the persistent grid.  AST: pair p owns ast_tpp
  // consecutive n-tiles of m-tile p / ast_ppm.
  int t_first = blockIdx.x / CG, t_step = gridDim.x / CG, t_count = total_tiles > t_first ? (total_tiles - t_first + t_step - 1) / t_step : 0;
  const bool ast_p = AST && tl.ast_persist != 0;
  if (AST && !ast_p) {
    const int pair = blockIdx.x / CG, mt = pair / tl.ast_ppm, nt0 = (pair % tl.ast_ppm) * tl.ast_tpp;
    t_first = mt * tl.n_tiles + nt0; t_step = 1;
    t_count = (skip_block || mt >= tl.m_tiles || nt0 >= tl.n_tiles) ? 0 : min(tl.ast_tpp, tl.n_tiles - nt0);
  }
  if (ast_p) {
    const int my_mt = t_first < tl.m_tiles ? (tl.m_tiles - t_first + t_step - 1) / t_step : 0;     // t_first = pair, t_step = pairs
    t_count = skip_block ? 0 : my_mt * tl.n_tiles;
  }
  // tile t -> (split, batch z, m-tile, n-tile).  The common launch (one K slice, no batch) needs one division, not five: the epilogue
  // warps spent ~700 clk per tile on this decode (tools/trace_gemm_roles.py) -- a third of a K = 128 tile.
  const bool plain_tiles = tl.splits == 1 && tiles_per_z == tl.total;
  auto decode = [&](int t, int& sp, int& z, int& mt, int& nt) {
    if (plain_tiles) { sp = 0; z = 0; mt = t / tl.n_tiles; nt = t - mt * tl.n_tiles; return; }
    sp = t / tiles_per_split;
    const int ts = t - sp * tiles_per_split;
    z = ts / tiles_per_z;
    const int rem = ts - z * tiles_per_z;
    mt = rem / tl.n_tiles; nt = rem - mt * tl.n_tiles;
  };
  // tile index of this pair's ti-th tile
  auto tile_of = [&](int ti) -> int {
    return ast_p ? (t_first + (ti / tl.n_tiles) * t_step) * tl.n_tiles + ti % tl.n_tiles : t_first + ti * t_step;
  };
  // AST: which A buffer / use count the ti-th tile reads, and whether it is the first / last n-tile of its m-tile
  const int ast_nbuf = AST ? (tl.ast_nbuf > 0 ? tl.ast_nbuf : 1) : 1;
  const bool is_producer = warp == 0;          // the producer warp waits for the previous kernel only after its weight prefetch
  if (!is_producer) {
    pdl_wait();
    if (threadIdx.x == 32) trace_stamp(trace, 2);              // previous kernel complete
  }

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp, uniform control flow; one elected lane issues)
    {
      const bool issuer = ptx::elect_one();
      auto b_coords = [&](int z, int n0, int kk, int& brow, int& bcol) {
        bcol = kk;
        if (sel_k) { brow = srow(kk / d.sel_span) + n0; bcol = kk % d.sel_span; }
        else if (d.sel == 1) brow = srow(n0 / d.sel_span) + n0 % d.sel_span;
        else brow = n0;
        brow += (int)(z * d.w_row_b) + (int)rank * Cfg::B_ROWS;      // this CTA's share of the tile's weight rows
      };
      // Weights are never written by a kernel of the step: the B tiles of the first pipeline fill are requested
      // while the previous kernel is still draining; only the activation (A) loads wait for it.
      int pre = 0;
      if (AST && tl.ast_bres) {
        // every (n-tile, k-block) tile of B into its own slot, once, while the previous kernel drains (weights)
        if (t_count > 0) {
          for (int nt = 0; nt < tl.n_tiles; ++nt)
            for (int kb = 0; kb < tl.num_kb; ++kb) {
              const int slot = nt * tl.num_kb + kb;
              int brow, bcol;
              b_coords(0, nt * BN, kb * BK, brow, bcol);
              if (issuer) {
                if (leader) ptx::mbar_arrive_expect_tx(&full[slot], Cfg::STAGE_BYTES * CG);
                ptx::tma_load_2d_2sm(tiles + slot * Cfg::STAGE_BYTES, &tmB, &full[slot], bcol, brow);
              }
            }
        }
      } else
      {
        const int t = tile_of(0);
        if (t_count > 0) {
          int sp, z, mt_, nt;
          decode(t, sp, z, mt_, nt);
          const int kb0 = sp * tl.kb_per, kb1 = min(tl.num_kb, kb0 + tl.kb_per);
          for (int kb = kb0; kb < kb1 && pre < STAGES; ++kb, ++pre) {
            uint8_t* b_dst = tiles + pre * Cfg::STAGE_BYTES + (AST ? 0 : Cfg::A_BYTES);
            int brow, bcol;
            b_coords(z, nt * BN, kb * BK, brow, bcol);
            if (issuer) {
              if (leader) ptx::mbar_arrive_expect_tx(&full[pre], Cfg::STAGE_BYTES * CG);
              if (CG == 2) ptx::tma_load_2d_2sm(b_dst, &tmB, &full[pre], bcol, brow);
              else ptx::tma_load_2d(b_dst, &tmB, &full[pre], bcol, brow);
            }
          }
        }
      }
      __syncwarp();
      pdl_wait();
      if (issuer) trace_stamp(trace, 2);                         // previous kernel complete (producer's view)
      if (AST && !ast_p && AMODE == AM_ROWS && t_count > 0 && issuer) {    // the pair's 256 x K block of A, once: one barrier per k-block so the MMAs start on the first
        const int m0r = (t_first / tl.n_tiles) * (BM * CG) + (int)rank * BM;
        for (int kb = 0; kb < tl.num_kb; ++kb) {
          if (leader) ptx::mbar_arrive_expect_tx(&a_full[kb], Cfg::A_BYTES * CG);
          if (CG == 2) ptx::tma_load_2d_2sm(a_res + kb * Cfg::A_BYTES, &tmA, &a_full[kb], kb * BK, m0r);
          else ptx::tma_load_2d(a_res + kb * Cfg::A_BYTES, &tmA, &a_full[kb], kb * BK, m0r);
        }
      }
      __syncwarp();
      uint32_t stage = 0, phase = 0;
      for (int ti = 0; ti < t_count; ++ti) {
        const int t = tile_of(ti);
        int sp, z, mt, nt;
        decode(t, sp, z, mt, nt);
        const int m0 = mt * (BM * CG) + (int)rank * BM, n0 = nt * BN;
        const int kb0 = sp * tl.kb_per, kb1 = min(tl.num_kb, kb0 + tl.kb_per);
        if (ast_p && nt == 0) {
          // a new m-tile: its A k-block tiles into the next buffer (the buffer's previous m-tile must have been consumed)
          const int i = ti / tl.n_tiles, ab = i % ast_nbuf;
          if (i >= ast_nbuf) wait_bar(&a_empty[ab], ((i / ast_nbuf) - 1) & 1, s_abort, fault, 6);
          if (issuer) {
            for (int kb = 0; kb < tl.num_kb; ++kb) {
              if (leader) ptx::mbar_arrive_expect_tx(&a_full[ab * tl.num_kb + kb], Cfg::A_BYTES * CG);
              ptx::tma_load_2d_2sm(a_res + (ab * tl.num_kb + kb) * Cfg::A_BYTES, &tmA, &a_full[ab * tl.num_kb + kb], kb * BK, m0);
            }
          }
          __syncwarp();
        }
        if (AST && tl.ast_bres) continue;                        // B is resident: nothing to stream per tile
        int b0 = 0, h0 = 0, w0 = 0;
        if (AMODE == AM_CONV3) {
          const int hw = d.cH * d.cW;
          b0 = m0 / hw;
          const int r = m0 % hw;
          h0 = r / d.cW;
          w0 = r % d.cW;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          const bool b_done = pre > 0;                          // this stage's B tile (and expect_tx) was issued up front
          if (b_done) --pre;
          else wait_bar(&empty[stage], phase ^ 1, s_abort, fault, 1);
          uint8_t* a_dst = tiles + stage * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + (AST ? 0 : Cfg::A_BYTES);
          const int kk = kb * BK;
          int brow = 0, bcol = 0;
          if (!b_done) b_coords(z, n0, kk, brow, bcol);
          const int tap = AMODE == AM_CONV3 ? kk / d.cC : 0, c0 = AMODE == AM_CONV3 ? kk % d.cC : 0;
          if (issuer) {
            if (leader && !b_done) ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES * CG);   // both CTAs' bytes land on the leader's barrier
            if (AST) {
              // A is resident
            } else if (AMODE == AM_ROWS) {
              if (CG == 2) ptx::tma_load_2d_2sm(a_dst, &tmA, &full[stage], kk + (int)(z * d.a_koff_b), m0);
              else ptx::tma_load_2d(a_dst, &tmA, &full[stage], kk + (int)(z * d.a_koff_b), m0);
            } else {
              if (CG == 2) ptx::tma_load_4d_2sm(a_dst, &tmA, &full[stage], c0 + (int)(z * d.a_koff_b), w0 + tap % 3 - 1, h0 + tap / 3 - 1, b0);
              else ptx::tma_load_4d(a_dst, &tmA, &full[stage], c0 + (int)(z * d.a_koff_b), w0 + tap % 3 - 1, h0 + tap / 3 - 1, b0);
            }
            if (!b_done) {
              if (CG == 2) ptx::tma_load_2d_2sm(b_dst, &tmB, &full[stage], bcol, brow);
              else ptx::tma_load_2d(b_dst, &tmB, &full[stage], bcol, brow);
            }
            if (ti == 0 && kb == kb0) trace_stamp(trace, 3);   // first loads issued
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (leader) {
      // The whole warp runs the loop (uniform control flow keeps descriptors, addresses and counters in uniform registers:
      // back-to-back tcgen05.mma instead of ~20 SASS instructions per issue); one elected lane issues.
      const bool issuer = ptx::elect_one();
      constexpr uint32_t idesc = ptx::idesc_bf16(BM * CG, BN);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int ti = 0; ti < t_count; ++ti) {
        const int t = tile_of(ti);
        lap(10);
        wait_bar(&tempty[as], aphase ^ 1, s_abort, fault, 2);
        lap(8);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        const int kb0 = (plain_tiles ? 0 : t / tiles_per_split) * tl.kb_per, kb1 = min(tl.num_kb, kb0 + tl.kb_per);
        // AST: the A buffer of this tile's m-tile; its k-block barriers are waited for by the m-tile's first n-tile
        const int a_i = ast_p ? ti / tl.n_tiles : 0, a_buf = a_i % ast_nbuf, a_nt = ast_p ? ti % tl.n_tiles : ti;
        const uint32_t a_par = static_cast<uint32_t>((a_i / ast_nbuf) & 1);
        const bool bres = AST && tl.ast_bres != 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (AST && a_nt == 0) wait_bar(&a_full[a_buf * tl.num_kb + kb], a_par, s_abort, fault, 5);
          const uint32_t slot = bres ? static_cast<uint32_t>(a_nt * tl.num_kb + kb) : stage;
          lap(10);
          if (!bres) wait_bar(&full[stage], phase, s_abort, fault, 3);
          else if (a_i == 0) wait_bar(&full[slot], 0, s_abort, fault, 3);       // resident B: landed once, stays (a poll of a completed barrier still costs ~90 clk)
          lap(9);
          if (issuer && ti == 0 && kb == kb0) trace_stamp(trace, 4);   // first operands landed
          ptx::tc_fence_after();
          // one descriptor per operand tile, k offsets added in 16-byte units (the MMA thread is issue-bound, see kernels_gconv.cu)
          const uint64_t s_desc = ptx::smem_desc_sw128(ptx::smem_u32(tiles + slot * Cfg::STAGE_BYTES));
          const uint64_t a_desc = AST ? ptx::smem_desc_sw128(ptx::smem_u32(a_res + (a_buf * tl.num_kb + kb) * Cfg::A_BYTES)) : s_desc;
          const uint64_t b_desc = AST ? s_desc : s_desc + Cfg::A_BYTES / 16;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if ((tl.dbg & 2) || !issuer) continue;
            if (CG == 2) ptx::umma_f16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            else ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (!bres) {
            if (issuer) { if (CG == 2) ptx::umma_commit_2sm(&empty[stage], 3); else ptx::umma_commit(&empty[stage]); }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (ast_p && a_nt == tl.n_tiles - 1) { if (issuer) ptx::umma_commit_2sm(&a_empty[a_buf], 3); __syncwarp(); }   // the m-tile's A buffer may be reloaded
        if (issuer) { if (CG == 2) ptx::umma_commit_2sm(&tfull[as], 3); else ptx::umma_commit(&tfull[as]); }   // accumulator complete -> epilogue(s)
        __syncwarp();
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
      if (issuer) trace_stamp(trace, 5);                               // all MMAs issued
#ifdef LDMB_TC_TRACE
      lap(10);
      if (trace != nullptr && lane == 0 && blockIdx.x < 256) for (int i = 8; i < 11; ++i) trace[blockIdx.x * 16 + 16 * 256 + i] = tacc[i];
#endif
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (warps 2..9 -> TMEM lane quadrant warp%4,
    //                                                        column half (warp-2)/4 of the tile)
    const int q = warp & 3;
    const int ew = warp - 2;                 // 0..7
    const int chalf = ew >> 2;               // which half of the tile's accumulator columns this warp drains
    const int et = threadIdx.x - 64;
    // branch-free activation: act(v) = max(v,0) + ns*min(v,0) with ns = 1 (none), 0 (relu), slope (leaky)
    const float ns = d.act == ACT_RELU ? 0.f : (d.act == ACT_LEAKY ? d.slope : 1.f);
    uint32_t as = 0, aphase = 0;
    int slab_sel = 0;
    auto bias_of = [&](int z, int n) -> float {
      const float* bp = d.bias + z * d.bias_off_b;
      if (sel_k) { float bv = 0.f; for (int qq = 0; qq < d.K / d.sel_span; ++qq) bv += bp[srow(qq) + n]; return bv; }
      if (d.sel == 1) return bp[srow(n / d.sel_span) + n % d.sel_span];
      return bp[n];
    };
    // The whole bias vector once per kernel where it fits: per tile the staging cost a global load and a 256-thread barrier on the
    // epilogue's critical path (0.2 us per tile, tools/trace_gemm_roles.py).
    const bool bias_all = plain_tiles && d.N <= Cfg::BIAS_FLOATS && t_count > 0 && !(tl.dbg & 8);
    if (bias_all) {
      for (int n = et; n < d.N; n += 32 * kEpiWarps) s_bias[n] = d.bias != nullptr ? bias_of(0, n) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    }
    for (int ti = 0; ti < t_count; ++ti) {
      const int t = tile_of(ti);
      int sp, z, mt, nt;
      decode(t, sp, z, mt, nt);
      const int m0 = mt * (BM * CG) + (int)rank * BM, n0 = nt * BN;
      float* sb = bias_all ? s_bias + n0 : s_bias + as * BN;
      lap(0);
      if (!bias_all)
      if (!(tl.dbg & 8) || ti == 0)
      for (int c = et; c < BN; c += 32 * kEpiWarps) {
        const int n = n0 + c;
        float bv = 0.f;
        if (n < d.N && d.bias != nullptr && sp == 0) bv = bias_of(z, n);      // split-K: the bias rides with the first K slice
        sb[c] = bv;
      }
      if (!bias_all && (!(tl.dbg & 8) || ti == 0)) asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      lap(1);
      wait_bar(&tfull[as], aphase, s_abort, fault, 4);
      lap(2);
      if (threadIdx.x == 64) trace_stamp(trace, ti == 0 ? 6 : 7);   // first / latest accumulator ready
      ptx::tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < d.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      if (tl.dbg & 1) {
      } else if (tl.tma_out) {
        // ---- smem-staged epilogue: each warp owns rows [q*32, q*32+32) of the tile and its own two 4 KB slabs
        const int orow = m0 + q * 32 + z * tl.out_row_b;
        const int ocol_z = z * tl.out_col_b;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        if (d.epi == EPI_STORE || d.epi == EPI_REGLU) {
          const int accw = d.epi == EPI_REGLU ? 128 : 64;        // accumulator columns consumed per 64-column bf16 slab
#pragma unroll 1
          for (int a0 = chalf * accw; a0 < BN; a0 += 2 * accw) {      // slabs interleaved between the two column warps
            if (n0 + a0 >= d.N) break;
            uint8_t* slab = staging + (ew * Cfg::SLABS + slab_sel % Cfg::SLABS) * Cfg::SLAB_BYTES;
            lap(0);
            if (lane == 0) { if (Cfg::SLABS == 1) ptx::bulk_wait_read<0>(); else ptx::bulk_wait_read<1>(); }   // the store that last used this slab has read it
            __syncwarp();
            lap(3);
            const uint32_t srow = ptx::smem_u32(slab) + lane * 128;
            if (d.epi == EPI_STORE && d.act == ACT_NONE && !(tl.dbg & (16 | 32))) {
              // in-projections: bias add only, both 32-column halves of the slab behind ONE TMEM round trip (the epilogue is issue- and
              // latency-bound: tools/trace_gemm_roles.py)
              uint32_t r0[32], r1[32];
              ptx::tmem_ld_32x32(t_row + a0, r0);
              ptx::tmem_ld_32x32(t_row + a0 + 32, r1);
              ptx::tmem_ld_wait();
              const uint32_t sba = ptx::smem_u32(sb + a0);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 b0 = ptx::ld_shared_v4(sba + u * 32), b1 = ptx::ld_shared_v4(sba + u * 32 + 16);
                ptx::st_shared_v4(srow + ((u ^ sw) << 4),
                                  pack_bf16(__uint_as_float(r0[8 * u]) + b0.x, __uint_as_float(r0[8 * u + 1]) + b0.y),
                                  pack_bf16(__uint_as_float(r0[8 * u + 2]) + b0.z, __uint_as_float(r0[8 * u + 3]) + b0.w),
                                  pack_bf16(__uint_as_float(r0[8 * u + 4]) + b1.x, __uint_as_float(r0[8 * u + 5]) + b1.y),
                                  pack_bf16(__uint_as_float(r0[8 * u + 6]) + b1.z, __uint_as_float(r0[8 * u + 7]) + b1.w));
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 b0 = ptx::ld_shared_v4(sba + 128 + u * 32), b1 = ptx::ld_shared_v4(sba + 128 + u * 32 + 16);
                ptx::st_shared_v4(srow + (((4 + u) ^ sw) << 4),
                                  pack_bf16(__uint_as_float(r1[8 * u]) + b0.x, __uint_as_float(r1[8 * u + 1]) + b0.y),
                                  pack_bf16(__uint_as_float(r1[8 * u + 2]) + b0.z, __uint_as_float(r1[8 * u + 3]) + b0.w),
                                  pack_bf16(__uint_as_float(r1[8 * u + 4]) + b1.x, __uint_as_float(r1[8 * u + 5]) + b1.y),
                                  pack_bf16(__uint_as_float(r1[8 * u + 6]) + b1.z, __uint_as_float(r1[8 * u + 7]) + b1.w));
              }
            } else
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t r[32];
              float v[32];
              if (d.epi == EPI_REGLU) {
                uint32_t rb[32];
                if (tl.dbg & 16) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) { r[i] = 0u; rb[i] = 0u; }
                } else {
                ptx::tmem_ld_32x32(t_row + a0 + half * 32, r);
                ptx::tmem_ld_32x32(t_row + a0 + 64 + half * 32, rb);
                ptx::tmem_ld_wait();
                }
                // per-image expert decisions: the 64-column slab lies inside one expert's block
                const bool keep = d.mask_plan == nullptr || expert_kept(d, m, (n0 + a0) >> 1);
                // biases as 16-byte shared loads (through a generic pointer they were 64 scalar loads per half)
                const uint32_t sba = ptx::smem_u32(sb + a0 + half * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  // (a + bias_a) * relu(b + bias_b): adds and the multiply as packed f32x2 (same roundings as the scalar form)
                  const float4 ba = ptx::ld_shared_v4(sba + i * 16), bg = ptx::ld_shared_v4(sba + 256 + i * 16);
                  float2 a01 = add2(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), make_float2(ba.x, ba.y));
                  float2 a23 = add2(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), make_float2(ba.z, ba.w));
                  float2 g01 = add2(make_float2(__uint_as_float(rb[4 * i]), __uint_as_float(rb[4 * i + 1])), make_float2(bg.x, bg.y));
                  float2 g23 = add2(make_float2(__uint_as_float(rb[4 * i + 2]), __uint_as_float(rb[4 * i + 3])), make_float2(bg.z, bg.w));
                  g01.x = fmaxf(g01.x, 0.f); g01.y = fmaxf(g01.y, 0.f); g23.x = fmaxf(g23.x, 0.f); g23.y = fmaxf(g23.y, 0.f);
                  a01 = mul2(a01, g01); a23 = mul2(a23, g23);
                  v[4 * i] = a01.x; v[4 * i + 1] = a01.y; v[4 * i + 2] = a23.x; v[4 * i + 3] = a23.y;
                }
                if (!keep) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
              } else {
                if (tl.dbg & 16) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) r[i] = 0u;
                } else {
                ptx::tmem_ld_32x32(t_row + a0 + half * 32, r);
                ptx::tmem_ld_wait();
                }
                const uint32_t sba = ptx::smem_u32(sb + a0 + half * 32);
                if (d.act == ACT_NONE) {      // in-projections: bias add only (the epilogue is issue-bound: ~6 instructions per element with the activation)
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 ba = ptx::ld_shared_v4(sba + i * 16);
                    v[4 * i] = __uint_as_float(r[4 * i]) + ba.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + ba.y;
                    v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + ba.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + ba.w;
                  }
                } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 ba = ptx::ld_shared_v4(sba + i * 16);
                  const float t0 = __uint_as_float(r[4 * i]) + ba.x, t1 = __uint_as_float(r[4 * i + 1]) + ba.y;
                  const float t2 = __uint_as_float(r[4 * i + 2]) + ba.z, t3 = __uint_as_float(r[4 * i + 3]) + ba.w;
                  v[4 * i] = fmaxf(t0, 0.f) + ns * fminf(t0, 0.f); v[4 * i + 1] = fmaxf(t1, 0.f) + ns * fminf(t1, 0.f);
                  v[4 * i + 2] = fmaxf(t2, 0.f) + ns * fminf(t2, 0.f); v[4 * i + 3] = fmaxf(t3, 0.f) + ns * fminf(t3, 0.f);
                }
                }
              }
              if (!(tl.dbg & 32))
#pragma unroll
              for (int u = 0; u < 4; ++u)
                ptx::st_shared_v4(srow + (((half * 4 + u) ^ sw) << 4), pack_bf16(v[8 * u], v[8 * u + 1]),
                                  pack_bf16(v[8 * u + 2], v[8 * u + 3]), pack_bf16(v[8 * u + 4], v[8 * u + 5]),
                                  pack_bf16(v[8 * u + 6], v[8 * u + 7]));
            }
            lap(5);
            if (!(tl.dbg & 64)) ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int ocol = (d.epi == EPI_REGLU ? ((n0 + a0) >> 1) : (n0 + a0)) + ocol_z;
              if (!(tl.dbg & 4)) ptx::tma_store_2d(&tmO, slab, ocol, orow);
              ptx::bulk_commit();
            }
            lap(6);
            slab_sel ^= 1;
          }
        } else {   // EPI_STORE_F32 / EPI_ACCUM_F32: 32 fp32 columns per slab
#pragma unroll 1
          for (int a0 = chalf * 32; a0 < BN; a0 += 64) {
            if (n0 + a0 >= d.N) break;
            uint8_t* slab = staging + (ew * Cfg::SLABS + slab_sel % Cfg::SLABS) * Cfg::SLAB_BYTES;
            lap(0);
            if (lane == 0) { if (Cfg::SLABS == 1) ptx::bulk_wait_read<0>(); else ptx::bulk_wait_read<1>(); }
            __syncwarp();
            lap(3);
            const uint32_t srow = ptx::smem_u32(slab) + lane * 128;
            uint32_t r[32];
            ptx::tmem_ld_32x32(t_row + a0, r);
            ptx::tmem_ld_wait();
            float v[32];
            {
              const uint32_t sba = ptx::smem_u32(sb + a0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 ba = ptx::ld_shared_v4(sba + i * 16);
                v[4 * i] = __uint_as_float(r[4 * i]) + ba.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + ba.y;
                v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + ba.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + ba.w;
              }
            }
            if (d.epi == EPI_STORE_F32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f) + ns * fminf(v[i], 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              ptx::st_shared_v4(srow + ((u ^ sw) << 4), __float_as_uint(v[4 * u]), __float_as_uint(v[4 * u + 1]),
                                __float_as_uint(v[4 * u + 2]), __float_as_uint(v[4 * u + 3]));
            lap(5);
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (tl.dbg & 4) {}
              else if (d.epi == EPI_ACCUM_F32) ptx::tma_reduce_add_2d(&tmO, slab, n0 + a0 + ocol_z, orow);
              else if (d.epi == EPI_UPADD) {
                // nearest x2 up-sampling (unet.py:85): the slab's 32 low-resolution pixels (a ww x bh rectangle) are added to
                // their four (dy, dx) positions of x viewed as [bh][dy][ww][dx][C] -- four bulk reduce-adds, one smem source
                const int ms = m0 + q * 32;
                const int ww0 = d.ctW >= 32 ? ms % d.ctW : 0, bh0 = ms / d.ctW;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) ptx::tma_reduce_add_5d(&tmO, slab, n0 + a0, q4 & 1, ww0, q4 >> 1, bh0);
              }
              else ptx::tma_store_2d(&tmO, slab, n0 + a0 + ocol_z, orow);
              ptx::bulk_commit();
            }
            slab_sel ^= 1;
          }
        }
      } else
#pragma unroll 1
      for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
        const int n = n0 + c * 32;
        if (n >= d.N) break;
        uint32_t r[32];
        if (d.epi == EPI_REGLU) {
          if ((c >> 1) & 1) continue;            // b-chunk, consumed together with its a-chunk
          uint32_t rb[32];
          ptx::tmem_ld_32x32(t_row + c * 32, r);
          ptx::tmem_ld_32x32(t_row + c * 32 + 64, rb);
          ptx::tmem_ld_wait();
          if (row_ok) {
            const int j = (n >> 7) * 64 + (n & 127);
            bf16* o = reinterpret_cast<bf16*>(d.out) + (long long)m * d.ldo + z * d.out_off_b + j;
            const bool keep = d.mask_plan == nullptr || expert_kept(d, m, j);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a0 = __uint_as_float(r[2 * i]) + sb[c * 32 + 2 * i];
              const float a1 = __uint_as_float(r[2 * i + 1]) + sb[c * 32 + 2 * i + 1];
              const float g0 = fmaxf(__uint_as_float(rb[2 * i]) + sb[c * 32 + 64 + 2 * i], 0.f);
              const float g1 = fmaxf(__uint_as_float(rb[2 * i + 1]) + sb[c * 32 + 64 + 2 * i + 1], 0.f);
              pk[i] = keep ? pack_bf16(a0 * g0, a1 * g1) : 0u;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<uint4*>(o)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
          continue;
        }
        ptx::tmem_ld_32x32(t_row + c * 32, r);
        ptx::tmem_ld_wait();
        if (!row_ok) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + sb[c * 32 + i];
        const bool full_chunk = n + 32 <= d.N;
        if (d.epi == EPI_ACCUM_F32) {
          float* o = reinterpret_cast<float*>(d.out) + (long long)m * d.ldo + z * d.out_off_b + n;
          if (d.res != nullptr) {     // another branch of the block, kept in bf16 (grouped-conv output)
            const bf16* rp = reinterpret_cast<const bf16*>(d.res) + (long long)m * d.ldr + n;
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + i);
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[k]);
                  v[8 * i + 2 * k] += __low2float(h2);
                  v[8 * i + 2 * k + 1] += __high2float(h2);
                }
              }
            } else {
              for (int i = 0; i < 32; ++i) if (n + i < d.N) v[i] += __bfloat162float(rp[i]);
            }
          }
          if (full_chunk) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 cur = reinterpret_cast<float4*>(o)[i];
              cur.x += v[4 * i]; cur.y += v[4 * i + 1]; cur.z += v[4 * i + 2]; cur.w += v[4 * i + 3];
              reinterpret_cast<float4*>(o)[i] = cur;
            }
          } else {
            for (int i = 0; i < 32; ++i) if (n + i < d.N) o[i] += v[i];
          }
        } else if (d.epi == EPI_UPADD) {
          // the low-resolution row feeds the 2x2 pixels of its nearest-upsampled position; x is only added to here
          const int HW = d.ctH * d.ctW, bb = m / HW, rr = m % HW, hh = rr / d.ctW, ww = rr % d.ctW;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float* o = reinterpret_cast<float*>(d.out) +
                       (((long long)bb * 2 * d.ctH + 2 * hh + (q4 >> 1)) * 2 * d.ctW + 2 * ww + (q4 & 1)) * d.ldo + n;
            if (full_chunk) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                             "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
            } else {
              for (int i = 0; i < 32; ++i) if (n + i < d.N) atomicAdd(o + i, v[i]);
            }
          }
        } else if (d.epi == EPI_STORE_F32) {
          float* o = reinterpret_cast<float*>(d.out) + (long long)m * d.ldo + z * d.out_off_b + n;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f) + ns * fminf(v[i], 0.f);
          if (full_chunk) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            for (int i = 0; i < 32; ++i) if (n + i < d.N) o[i] = v[i];
          }
        } else {   // EPI_STORE / EPI_CONVT: bf16 out
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f) + ns * fminf(v[i], 0.f);
          bf16* o;
          if (d.epi == EPI_CONVT) o = reinterpret_cast<bf16*>(d.out) + convt_offset(d, m, n);
          else o = reinterpret_cast<bf16*>(d.out) + (long long)m * d.ldo + z * d.out_off_b + n;
          if (full_chunk) {
            if (d.res != nullptr) {
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(d.res) + (long long)m * d.ldr + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(rp + i);
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[k]);
                  v[8 * i + 2 * k] += __low2float(h2);
                  v[8 * i + 2 * k + 1] += __high2float(h2);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<uint4*>(o)[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                                          pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
          } else {
            for (int i = 0; i < 32; ++i)
              if (n + i < d.N) {
                float x = v[i];
                if (d.res != nullptr) x += __bfloat162float(reinterpret_cast<const bf16*>(d.res)[(long long)m * d.ldr + n + i]);
                o[i] = __float2bfloat16_rn(x);
              }
          }
        }
      }
      lap(0);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CG == 2) ptx::mbar_arrive_leader(&tempty[as]); else ptx::mbar_arrive(&tempty[as]); }
      lap(7);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (tl.tma_out && lane == 0) ptx::bulk_wait_read<0>();  // the stores have read their smem slabs; the writes themselves complete with the grid
    __syncwarp();
    if (threadIdx.x == 64) trace_stamp(trace, 8);            // epilogue done
#ifdef LDMB_TC_TRACE
    if (trace != nullptr && threadIdx.x == 64 && blockIdx.x < 256) for (int i = 0; i < 8; ++i) trace[blockIdx.x * 16 + 16 * 256 + i] = tacc[i];
#endif
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();      // the peer has stopped reading our smem / signalling our barriers
  if (warp == 1) { if (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
  if (threadIdx.x == 0) trace_stamp(trace, 9);               // exit
}

bool conv_tile(const GemmDesc& d, int& TW, int& TH, int& TB) {
  const int W = d.cW, H = d.cH;
  if (W >= 128) {
    if (W % 128) return false;
    TW = 128; TH = 1; TB = 1;
    return true;
  }
  if (W <= 0 || (128 % W) != 0) return false;
  TW = W;
  const int rows = 128 / W;
  if (H >= rows) {
    if (H % rows) return false;
    TH = rows; TB = 1;
    return true;
  }
  if (rows % H) return false;
  TH = H; TB = rows / H;
  return true;
}

int pick_bn(const GemmDesc& d, int num_sms) {
  const int batch = d.batch > 0 ? d.batch : 1;
  const long long m_tiles = (d.M + BM - 1) / BM;
  auto ok = [&](int bn) {
    if (d.N % bn != 0 && !(d.sel == 0 && d.epi != EPI_REGLU && bn == 64 && d.N % 32 == 0)) return false;
    if (d.sel == 1 && d.sel_span % bn) return false;
    if (d.epi == EPI_REGLU && bn < 128) return false;
    return true;
  };
  if (ok(256)) return 256;       // the widest tile: L2->SM bytes per FLOP decide the main-loop rate
  if (ok(128)) return 128;
  if (ok(64)) return 64;
  return 0;
}

}  // namespace

bool tc_supported(const GemmDesc& d) {
  if (d.K <= 0 || d.K % BK) return false;
  if (d.ldw % 8) return false;
  if (d.amode == AM_ROWS) {
    if (d.lda % 8 || d.a_koff_b % 8) return false;
  } else {
    int a, b, c;
    if (d.cC % BK || d.lda % 8 || d.a_koff_b % 8) return false;
    if (!conv_tile(d, a, b, c)) return false;
  }
  if ((d.sel == 2 || d.sel == 3) && d.sel_span % BK) return false;
  if (d.epi == EPI_REGLU && d.glu_chunk != 64) return false;
  if (d.mask_plan != nullptr && (d.epi != EPI_REGLU || d.mask_span % 64 != 0 || d.mask_rows <= 0)) return false;   // a 64-column slab lies in one expert
  if (d.epi == EPI_CONVT && d.ctC % 32) return false;
  if (d.epi != EPI_CONVT && (d.ldo % 8 || d.out_off_b % 8)) return false;
  if (d.res && d.ldr % 8) return false;
  if (d.N % 32) return false;
  return pick_bn(d, 148) != 0;
}

// Mirrors launch_gemm_tc's choice of the smem-staged TMA epilogue for an accumulate-into-the-residual GEMM.
bool tc_accum_is_reduction(const GemmDesc& d) {
  if (d.epi != EPI_ACCUM_F32 || !tc_supported(d)) return false;
  const int batch = d.batch > 0 ? d.batch : 1;
  return d.res == nullptr && batch == 1 && d.N % 32 == 0 && (d.ldo * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(d.out) % 16) == 0;
}

TcContext* tc_context_create(int device, char* err, int errlen) {
  TcContext* ctx = new TcContext();
  ctx->device = device;
  ctx->encode = nullptr;
  ctx->fault_dev = nullptr;
  ctx->fault_host = nullptr;
  ctx->trace_dev = nullptr;
  ctx->splitk = true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    delete ctx;
    return nullptr;
  }
  ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { snprintf(err, errlen, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); delete ctx; return nullptr; }
  if (prop.major != 10) {
    snprintf(err, errlen, "device %d is sm_%d%d; libldmb200 is built for sm_100a only", device, prop.major, prop.minor);
    delete ctx;
    return nullptr;
  }
  ctx->num_sms = prop.multiProcessorCount;
  e = cudaMalloc(&ctx->fault_dev, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(ctx->fault_dev, 0, 4 * sizeof(int));
  if (e != cudaSuccess) { snprintf(err, errlen, "cudaMalloc(fault flag): %s", cudaGetErrorString(e)); delete ctx; return nullptr; }
  // host-mapped mirror of the fault word: polled by every API call without a synchronisation (LDMB_ERR_KERNEL)
  int* fh = nullptr;
  if (cudaHostAlloc(reinterpret_cast<void**>(&fh), sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    *fh = 0;
    int* fh_dev = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&fh_dev), fh, 0) == cudaSuccess) {
      const unsigned long long addr = reinterpret_cast<unsigned long long>(fh_dev);
      if (cudaMemcpy(ctx->fault_dev + 2, &addr, sizeof(addr), cudaMemcpyHostToDevice) == cudaSuccess) ctx->fault_host = fh;
    }
    if (!ctx->fault_host) cudaFreeHost(fh);
  }
  return ctx;
}

void tc_context_destroy(TcContext* ctx) {
  if (!ctx) return;
  if (ctx->fault_dev) cudaFree(ctx->fault_dev);
  if (ctx->fault_host) cudaFreeHost(const_cast<int*>(ctx->fault_host));
  if (ctx->trace_dev) cudaFree(ctx->trace_dev);
  delete ctx;
}

int tc_trace_enable(TcContext* ctx, int on) {
  if (on && !ctx->trace_dev) {
    if (cudaMalloc(&ctx->trace_dev, sizeof(long long) * kTraceSlots * 512) != cudaSuccess) return -1;     // [256 CTAs][16] stamps + [256][16] per-role accounting
    cudaMemset(ctx->trace_dev, 0, sizeof(long long) * kTraceSlots * 512);
  } else if (!on && ctx->trace_dev) {
    cudaFree(ctx->trace_dev);
    ctx->trace_dev = nullptr;
  }
  return 0;
}
int tc_trace_read(TcContext* ctx, long long* host, int max_ctas) {
  if (!ctx->trace_dev) return -1;
  const int n = max_ctas < 512 ? max_ctas : 512;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(host, ctx->trace_dev, sizeof(long long) * kTraceSlots * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}

void tc_set_splitk(TcContext* ctx, bool on) { ctx->splitk = on; }

int tc_num_sms(const TcContext* ctx) { return ctx->num_sms; }
int tc_poll_fault(const TcContext* ctx) { return ctx->fault_host ? *ctx->fault_host : 0; }

int tc_read_fault(TcContext* ctx, cudaStream_t s) {
  int v = 0;
  if (cudaMemcpyAsync(&v, ctx->fault_dev, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
  return v;
}

template <int BN, int AMODE, int CG, bool AST = false>
static cudaError_t launch_tc_inst(TcContext* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                                  const GemmDesc& d, const TcTiling& tl, cudaStream_t s) {
  using Cfg = TcCfg<BN, CG, AST>;
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, AMODE, CG, AST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr.mark(ctx->device);
  }
  int grid = tl.total * CG < ctx->num_sms ? tl.total * CG : (ctx->num_sms / CG) * CG;
  if (AST) grid = tl.m_tiles * tl.ast_ppm * CG;            // one pair per (m-tile, run of ast_tpp n-tiles)
  if (AST && tl.ast_persist) { const int pairs = ctx->num_sms / CG; grid = (tl.m_tiles < pairs ? tl.m_tiles : pairs) * CG; }   // persistent over the m-tiles
  if (!AST && tc_knobs().grid > 0 && tc_knobs().grid < grid) grid = (tc_knobs().grid / CG) * CG;   // debug
  if (!AST && d.max_ctas >= CG && d.max_ctas < grid) grid = (d.max_ctas / CG) * CG;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (g_ldmb_pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
  if (CG == 2) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, AMODE, CG, AST>, tmA, tmB, tmO, d, tl, ctx->fault_dev, ctx->trace_dev);
}

template <int AMODE, int CG>
static cudaError_t launch_tc_bn(TcContext* ctx, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                                const GemmDesc& d, const TcTiling& tl, cudaStream_t s) {
  if (bn == 256) return launch_tc_inst<256, AMODE, CG>(ctx, tmA, tmB, tmO, d, tl, s);
  if (bn == 128) return launch_tc_inst<128, AMODE, CG>(ctx, tmA, tmB, tmO, d, tl, s);
  return launch_tc_inst<64, AMODE, CG>(ctx, tmA, tmB, tmO, d, tl, s);
}

const TcKnobs& tc_knobs() {
  static const TcKnobs k = [] {
    auto geti = [](const char* n, int dflt) { const char* v = getenv(n); return v ? atoi(v) : dflt; };
    TcKnobs t;
    t.grid = geti("LDMB_TC_GRID", 0);          // debug: cap the persistent grid
    t.stages = geti("LDMB_TC_STAGES", 0);      // debug: use only this many smem pipeline stages
    t.dbg = geti("LDMB_TC_DBG", 0);            // debug bits, see TcTiling::dbg
    t.splits = geti("LDMB_TC_SPLITS", 0);      // debug: force the split-K factor
    t.force_cg = geti("LDMB_TC_CG", 0);        // debug: 1 or 2 forces the CTA-group size
    t.no_splitk = getenv("LDMB_NO_SPLITK") != nullptr;
    t.mlp_dbg = geti("LDMB_MLP_DBG", 0);
    t.gconv_dbg = geti("LDMB_GCONV_DBG", 0);
    t.attn_dbg = geti("LDMB_ATTN_DBG", 0);
    t.astp = geti("LDMB_TC_ASTP", 0);        // 1: persistent A-stationary tiling for the in-projection GEMMs at C <= 256
    t.ast = geti("LDMB_TC_AST", 0);          // 1: A-stationary tiling for bf16-out K <= 512 GEMMs (measured slower: 17.7 vs 15.8 us at M4096 N3072 K512, see profiles/r2_experiment_a_stationary.txt)
    return t;
  }();
  return k;
}

static cudaError_t launch_gemm_tc_impl(TcContext* ctx, const GemmDesc& d, cudaStream_t s, int* ctas_only);
cudaError_t launch_gemm_tc(TcContext* ctx, const GemmDesc& d, cudaStream_t s) { return launch_gemm_tc_impl(ctx, d, s, nullptr); }
int tc_gemm_ctas(TcContext* ctx, const GemmDesc& d) {
  int n = 0;
  return launch_gemm_tc_impl(ctx, d, nullptr, &n) == cudaSuccess ? n : 0;
}

// ctas_only != NULL: dry run -- only reports the grid the launch would use
static cudaError_t launch_gemm_tc_impl(TcContext* ctx, const GemmDesc& d, cudaStream_t s, int* ctas_only) {
  if (!tc_supported(d)) return cudaErrorNotSupported;
  int bn = pick_bn(d, ctx->num_sms);
  const int batch = d.batch > 0 ? d.batch : 1;
  // the up-sampling ch_conv (EPI_UPADD: M = low-resolution pixels, 1024..16384 rows) has few output tiles: narrower
  // tiles, single CTAs and split-K (below) put it on ~128 SMs instead of 16..32
  if (d.epi == EPI_UPADD && bn == 256 && d.N % 128 == 0) bn = 128;
  // CTA pairs (cta_group::2) whenever there are at least two 128-row tiles to pair up
  int cg = d.M > BM ? 2 : 1;
  // the up-sampling GEMM with few output tiles is cut along K below: single-CTA tiles give twice the CTAs per split.
  // (The residual GEMMs used the same rule while the issuing thread was the bound; with the warp-uniform issue loops
  // pairs + split-K measure 1.3 % faster end to end.)
  if (d.epi == EPI_UPADD && (long long)((d.M + BM - 1) / BM) * ((d.N + bn - 1) / bn) * batch <= ctx->num_sms) cg = 1;
  if (tc_knobs().force_cg == 1 || tc_knobs().force_cg == 2) cg = tc_knobs().force_cg;
  TcTiling tl;
  tl.m_tiles = (d.M + BM * cg - 1) / (BM * cg);
  tl.n_tiles = (d.N + bn - 1) / bn;
  tl.num_kb = d.K / BK;
  tl.total = tl.m_tiles * tl.n_tiles * batch;
  tl.splits = 1; tl.kb_per = tl.num_kb;
  tl.max_stages = tc_knobs().stages;
  tl.dbg = tc_knobs().dbg;
  tl.TW = tl.TH = tl.TB = 0;
  tl.ast_ppm = tl.ast_tpp = 0;
  tl.ast_persist = tl.ast_nbuf = tl.ast_bres = 0;

  CUtensorMap tmA, tmB;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  CUresult r;
  if (d.amode == AM_ROWS) {
    // inner extent: every column reachable from this A (covers the batch k-offsets); rows = M
    const cuuint64_t gdim[2] = {(cuuint64_t)d.lda, (cuuint64_t)d.M};
    const cuuint64_t gstr[1] = {(cuuint64_t)d.lda * 2};
    const cuuint32_t box[2] = {BK, BM};
    r = ctx->encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d.A), gdim, gstr, box, ones,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    conv_tile(d, tl.TW, tl.TH, tl.TB);
    const int Bimg = d.M / (d.cH * d.cW);
    const cuuint64_t gdim[4] = {(cuuint64_t)d.lda, (cuuint64_t)d.cW, (cuuint64_t)d.cH, (cuuint64_t)Bimg};
    const cuuint64_t gstr[3] = {(cuuint64_t)d.lda * 2, (cuuint64_t)d.lda * 2 * d.cW, (cuuint64_t)d.lda * 2 * d.cW * d.cH};
    const cuuint32_t box[4] = {BK, (cuuint32_t)tl.TW, (cuuint32_t)tl.TH, (cuuint32_t)tl.TB};
    r = ctx->encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.A), gdim, gstr, box, ones,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  {
    // weight rows: everything reachable (stacked experts / batches); the epilogue guards columns >= N
    long long rows = d.N;
    if (d.sel == 1) { rows = 0; for (int q = 0; q < d.N / d.sel_span; ++q) if (d.sel_rows[q] + d.sel_span > rows) rows = d.sel_rows[q] + d.sel_span; }
    if (d.sel == 2) { rows = 0; for (int q = 0; q < d.K / d.sel_span; ++q) if (d.sel_rows[q] + d.N > rows) rows = d.sel_rows[q] + d.N; }
    if (d.plan != nullptr && d.sel == 1) rows = 5LL * d.sel_stride;      // any expert may be picked at replay time
    if (d.plan != nullptr && d.sel == 2) rows = (d.K / d.sel_span == 4 ? 6LL : 5LL) * d.sel_stride;
    if (d.sel == 3) rows = (long long)(d.K / d.sel_span - 1) * d.sel_stride + d.N;
    rows += (long long)(batch - 1) * d.w_row_b;
    const cuuint64_t gdim[2] = {(cuuint64_t)d.ldw, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)d.ldw * 2};
    const cuuint32_t box[2] = {BK, (cuuint32_t)(bn / cg)};
    r = ctx->encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d.W), gdim, gstr, box, ones,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  // ---- output tensor map (smem-staged epilogue), when the shape allows it
  CUtensorMap tmO;
  memset(&tmO, 0, sizeof(tmO));
  tl.tma_out = 0; tl.out_col_b = 0; tl.out_row_b = 0;
  {
    // EPI_UPADD: staged when a 32-row slab of low-resolution pixels is a rectangle of the image grid (else: red.v4 scatter)
    const bool up = d.epi == EPI_UPADD && batch == 1 && d.ctW > 0 && (d.ctW % 32 == 0 || 32 % d.ctW == 0) && d.M % d.ctW == 0;
    const bool f32 = d.epi == EPI_STORE_F32 || d.epi == EPI_ACCUM_F32 || up;
    const bool mode_ok = (d.epi == EPI_STORE || d.epi == EPI_REGLU || f32) && d.res == nullptr;
    const int esz = f32 ? 4 : 2;
    const int slab_cols = f32 ? 32 : 64;
    const int out_n = d.epi == EPI_REGLU ? d.N / 2 : d.N;
    bool ok = mode_ok && out_n % slab_cols == 0 && (d.ldo * esz) % 16 == 0 && ((uintptr_t)d.out % 16) == 0;
    long long rows = d.M;
    if (ok && batch > 1) {
      if (d.out_off_b + out_n <= d.ldo && d.out_off_b * (batch - 1) + out_n <= d.ldo) tl.out_col_b = (int)d.out_off_b;
      else if (d.out_off_b % d.ldo == 0 && d.M % (BM * cg) == 0) {   /* no tile may spill into the next batch's rows */ tl.out_row_b = (int)(d.out_off_b / d.ldo); rows = (long long)tl.out_row_b * (batch - 1) + d.M; }
      else ok = false;
    }
    if (ok && up) {
      const cuuint64_t W = (cuuint64_t)d.ctW, BH = (cuuint64_t)(d.M / d.ctW), ld = (cuuint64_t)d.ldo * 4;
      const cuuint64_t gdim[5] = {(cuuint64_t)d.ldo, 2, W, 2, BH};
      const cuuint64_t gstr[4] = {ld, 2 * ld, 2 * W * ld, 4 * W * ld};
      const cuuint32_t wb = d.ctW >= 32 ? 32u : (cuuint32_t)d.ctW;
      const cuuint32_t box[5] = {32, 1, wb, 1, 32 / wb};
      const cuuint32_t ones5[5] = {1, 1, 1, 1, 1};
      r = ctx->encode(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d.out, gdim, gstr, box, ones5, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
      tl.tma_out = 1;
    } else if (ok) {
      const cuuint64_t gdim[2] = {(cuuint64_t)d.ldo, (cuuint64_t)rows};
      const cuuint64_t gstr[1] = {(cuuint64_t)d.ldo * esz};
      const cuuint32_t box[2] = {(cuuint32_t)slab_cols, 32};
      r = ctx->encode(&tmO, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d.out, gdim, gstr, box,
                      ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
      tl.tma_out = 1;
    }
  }
  // split-K: an accumulate-into-the-residual GEMM with too few output tiles to fill the machine (deep UNet levels:
  // M = 1024..4096 rows, K = 3C..4C) is cut along K; every slice reduce-adds its partial tile into x with TMA.
  // (EPI_UPADD scatters its partial tiles with red.global.add: just as linear)
  if (((tl.tma_out && d.epi == EPI_ACCUM_F32) || d.epi == EPI_UPADD) && !tc_knobs().no_splitk && ctx->splitk) {
    const int base = tl.total * cg;
    int sp = ctx->num_sms / (base > 0 ? base : 1);
    if (sp > tl.num_kb / 4) sp = tl.num_kb / 4;              // at least 4 k-blocks (K = 256) per slice
    if (tc_knobs().splits > 0) sp = tc_knobs().splits;   // debug
    if (sp > 1) {
      tl.kb_per = (tl.num_kb + sp - 1) / sp;
      tl.splits = (tl.num_kb + tl.kb_per - 1) / tl.kb_per;
      tl.total *= tl.splits;
    }
  }
  // A-stationary: bf16-out GEMMs with K <= 512 whose pairs would otherwise re-load the same A rows for every n-tile (ReGLU a|b at
  // C = 512, in-projections): each pair takes ast_tpp consecutive n-tiles of one m-tile.  Same number of tile rounds as the strided
  // persistent grid (ast_tpp = ceil(tiles / pairs)), fewer bytes per SM.
  bool ast = false;
  if (tc_knobs().ast && d.amode == AM_ROWS && cg == 2 && (bn == 256 || bn == 128) && batch == 1 && tl.splits == 1 && tl.tma_out &&
      (d.epi == EPI_STORE || d.epi == EPI_REGLU) && tl.num_kb <= kAstMaxKb && d.a_koff_b == 0 && tl.n_tiles >= 2) {
    const int pairs = ctx->num_sms / 2;
    int tpp = (tl.total + pairs - 1) / pairs;
    if (tpp > tl.n_tiles) tpp = tl.n_tiles;          // more m-tiles than pairs: one pair per m-tile (all its n-tiles), several waves of CTAs
    const int ppm = (tl.n_tiles + tpp - 1) / tpp;
    if (tpp >= 2 && (tl.m_tiles * ppm <= pairs || tc_knobs().ast == 2)) { ast = true; tl.ast_tpp = tpp; tl.ast_ppm = ppm; }
  }
  // Persistent A-stationary (LDMB_TC_ASTP, in-projection GEMMs at C <= 256: K = C, N = 3C, many more m-tiles than CTA pairs): pair p walks
  // the m-tiles p, p + P, ..., loads each m-tile's A block ONCE (double buffered) and runs all its n-tiles against it; where all of B fits
  // the ring's slots it is loaded once and stays.  The strided grid re-reads A once per n-tile (3x) and B once per tile.
  if (tc_knobs().astp && !ast && d.amode == AM_ROWS && cg == 2 && (bn == 256 || bn == 128) && batch == 1 && tl.splits == 1 && tl.tma_out &&
      d.epi == EPI_STORE && d.sel == 0 && tl.num_kb <= 4 && d.a_koff_b == 0 && tl.n_tiles >= 2 && tl.m_tiles > ctx->num_sms / 2) {
    ast = true; tl.ast_persist = 1; tl.ast_ppm = 1; tl.ast_tpp = tl.n_tiles;
    tl.ast_nbuf = kAstMaxKb / tl.num_kb >= 2 ? 2 : 1;
    const int slots = bn == 256 ? TcCfg<256, 2, true>::STAGES : TcCfg<128, 2, true>::STAGES;
    tl.ast_bres = tl.n_tiles * tl.num_kb <= slots ? 1 : 0;
  }
  if (ctas_only != nullptr) {
    int grid = ast ? tl.m_tiles * tl.ast_ppm * cg : (tl.total * cg < ctx->num_sms ? tl.total * cg : (ctx->num_sms / cg) * cg);
    if (ast && tl.ast_persist) grid = (tl.m_tiles < ctx->num_sms / cg ? tl.m_tiles : ctx->num_sms / cg) * cg;
    if (!ast && d.max_ctas >= cg && d.max_ctas < grid) grid = (d.max_ctas / cg) * cg;
    *ctas_only = grid;
    return cudaSuccess;
  }
  if (ast) return bn == 256 ? launch_tc_inst<256, AM_ROWS, 2, true>(ctx, tmA, tmB, tmO, d, tl, s) : launch_tc_inst<128, AM_ROWS, 2, true>(ctx, tmA, tmB, tmO, d, tl, s);
  if (d.amode == AM_ROWS)
    return cg == 2 ? launch_tc_bn<AM_ROWS, 2>(ctx, bn, tmA, tmB, tmO, d, tl, s) : launch_tc_bn<AM_ROWS, 1>(ctx, bn, tmA, tmB, tmO, d, tl, s);
  return cg == 2 ? launch_tc_bn<AM_CONV3, 2>(ctx, bn, tmA, tmB, tmO, d, tl, s) : launch_tc_bn<AM_CONV3, 1>(ctx, bn, tmA, tmB, tmO, d, tl, s);
}
