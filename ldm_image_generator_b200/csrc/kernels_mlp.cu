// Fused RandomMoE / ReGLU feed-forward of one SwinBlock (modules.py:14-15,34-36; unet.py:44,47) for C = 128 / 256:
//     x += sum_{e in {general, e1, e2}}  c_e( a_e(xm) * relu(b_e(xm)) )          (all 1x1 convs = GEMMs over pixels)
// in ONE tcgen05 kernel: the gated hidden activations h (3C bf16 per pixel, the largest tensor of the block) never
// leave the SM.  Per 128-row tile (256 rows per CTA pair with cta_group::2):
//   GEMM1 unit (e, j):  D1[128 x 128] = xm_tile[128 x C] . Wab_e[rows j*128 .. +128 (64 a | 64 b), C]^T      (TMEM, x2 buffers)
//   epilogue 1:         h_(e,j)[128 x 64] = (D1_a + bias_a) * relu(D1_b + bias_b)  -> bf16, 128B-swizzled smem ring
//   GEMM2 unit (e, j):  D2[128 x C] += h_(e,j) . Wc_e[:, j*64 .. +64]^T                                        (TMEM)
//   epilogue 2:         D2 + sum_e bias_c  -> smem slabs -> TMA fp32 reduce-add into the residual stream x.
// The MMA thread software-pipelines GEMM1(s) with GEMM2(s-LAG) (2-3 D1 accumulators) so the tensor pipe works while the
// epilogue warps gate the units in between.  The xm tile is loaded once per tile (NKB k-block tiles) and re-used by all 3*NKB GEMM1 units; weights
// stream through two TMA rings (Wab tiles, Wc tiles).  Experts are resolved on the device from the block's plan entry
// so the launch is static under CUDA-graph replay.  Warp roles: 0 TMA producer, 1 MMA issuer, 2..9 epilogue.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tc_context.h"

namespace {

constexpr int kEpiWarps = 8;                       // warps of ONE epilogue group (two per TMEM lane quadrant)
// Two epilogue groups take alternate units: with one group the gate epilogue (tcgen05.ld -> gate -> bf16 -> swizzled smem,
// ~1 500 busy clk per unit on 8 warps) was slower than the unit's MMAs (~1 200 clk), and the MMA thread waited for h
// (tools/trace_mlp.py, profiles/r2_trace_mlp_roles.txt).
constexpr int kEpiGroups = 2;
constexpr int kThreads = 64 + 32 * kEpiWarps * kEpiGroups;

// RES (C = 128 pairs, one plan for the batch): the three experts' Wab / Wc tiles stay RESIDENT in shared memory for all
// tiles of the CTA (loaded once, before the previous kernel has finished) instead of streaming through the rings per
// tile -- at 4 tiles per CTA the re-streamed weights were 144 of the 176 KB a CTA ingests per tile.
// ATT (attention blocks): NKB extra GEMM2 units per tile whose A operand is the tile of the attention core's output (TMA-loaded into the
// h ring) and whose weights are the MHA out_proj rows (slot 5 of w_c): x += att . W_out^T rides in the same D2 accumulator
// (attention.py:82 out_proj; unet.py:44,47) -- no separate K = C GEMM and no second pass of reduce-adds over x.
template <int C, int CG, bool RES, bool ATT = false> struct MlpCfg {
  static constexpr int NKB = C / 64;                       // k-blocks of GEMM1 = 64-column h chunks per expert
  static constexpr int UNITS = 3 * NKB;                   // gated units (GEMM1 + gate + GEMM2)
  static constexpr int UNITS2 = UNITS + (ATT ? NKB : 0);   // GEMM2 units
  static constexpr int A1_BYTES = NKB * 128 * 128;         // xm tile: NKB k-block tiles of [128 rows x 128 B]
  static constexpr int A1_BUFS = (C == 128 && CG == 2 && !RES) ? 2 : 1;
  // One W1 stage = all NKB k-block tiles of a unit's Wab rows, one W2 stage = the unit's Wc tile: a unit costs the MMA
  // thread two tcgen05.commit (GEMM1 done, GEMM2 done) -- commits, not MMAs, bounded the first version of this kernel.
  static constexpr int B1_ROWS = 128 / CG, W1_BYTES = NKB * B1_ROWS * 128;
  static constexpr int B2_ROWS = C / CG, W2_BYTES = B2_ROWS * 128;
  static constexpr int ND1 = C == 128 ? 3 : 2;             // D1 accumulators (128 TMEM columns each)
  static constexpr int LAG = ND1 - 1;                      // GEMM2(u) is issued LAG units behind GEMM1: hides the MMA->epilogue->MMA round trip
  static constexpr int W1S = RES ? UNITS : (C == 128 ? (CG == 2 ? 5 : 3) : 2);
  static constexpr int HS = (C == 128 && CG == 2 && !RES) ? 3 : 2;  // h slots; streaming: == W2 stages (one barrier frees both)
  static constexpr int W2S = RES ? UNITS : HS;
  static constexpr int H_BYTES = 128 * 128;
  static constexpr int D2_COL = ND1 * 128;
  static constexpr int SLAB_BYTES = 32 * 128;              // epilogue-2 staging aliases the (then idle) h ring
  static constexpr int TILE_BYTES = A1_BUFS * A1_BYTES + W1S * W1_BYTES + W2S * W2_BYTES + HS * H_BYTES;
  static_assert(!RES || (C == 128 && CG == 2), "resident weights fit for C = 128 pairs only");
  static_assert(!(RES && ATT), "the attention units stream their operands");
  static_assert(W1S <= 8 && W2S <= 8, "barrier slots");
  static constexpr int BAR_BYTES = 512;
  static constexpr int BIAS_FLOATS = 5 * 2 * C + 2 * C;    // a|b biases of all five experts + the tile's summed c biases (x2)
  static constexpr int SMEM_BYTES = 1024 + TILE_BYTES + BAR_BYTES + BIAS_FLOATS * 4;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(W1_BYTES % 1024 == 0 && W2_BYTES % 1024 == 0, "1024-byte aligned tiles");
  static_assert(W1S >= ND1, "a W1 slot must not complete twice before the epilogue has seen it");
  static_assert(kEpiWarps * SLAB_BYTES <= HS * H_BYTES, "staging fits the h ring");
  static_assert(D2_COL + C <= 512, "TMEM columns");
};

struct MlpArgs {
  const float* b_ab; const float* b_c;
  const int* plan; int e1, e2;          // plan entry {skip, e1, e2, -} of the block, or explicit experts when NULL
  const int* plan_img; int rows_per_image;   // per-image decisions (skip | e1 << 8 | e2 << 16 per image; a tile lies inside one image), or NULL
  int M, m_tiles;
  int dbg;                              // debug experiments: 1 = epilogue 1 idle, 2 = no GEMM2 MMAs, 4 = no GEMM1 MMAs, 8 = no x update
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

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// packed fp32 pairs (sm_100 f32x2 pipe): two adds / multiplies per instruction in the gate epilogue
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

// debug (ldmb_debug_tc_trace): slots 0..9 %globaltimer stamps
__device__ __forceinline__ void trace_stamp(long long* trace, int slot) {
  if (trace != nullptr) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 16 + slot] = t;
  }
}

__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(ptx::smem_u32(bar)), "r"(n) : "memory");
}

struct Ring {
  uint32_t i = 0, ph = 0;
  __device__ __forceinline__ void next(uint32_t n) { if (++i == n) { i = 0; ph ^= 1; } }
};

template <int C, int CG, bool RES, bool ATT>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWab,
                 const __grid_constant__ CUtensorMap tmWc, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmAtt,
                 const MlpArgs a, int* fault, long long* trace) {
  using Cfg = MlpCfg<C, CG, RES, ATT>;
  constexpr int NKB = Cfg::NKB, UNITS = Cfg::UNITS, UNITS2 = Cfg::UNITS2, LAG = Cfg::LAG, ND1 = Cfg::ND1, W1S = Cfg::W1S, HS = Cfg::HS, W2S = Cfg::W2S;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* a1 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w1 = a1 + Cfg::A1_BUFS * Cfg::A1_BYTES;
  uint8_t* w2 = w1 + W1S * Cfg::W1_BYTES;
  uint8_t* hs = w2 + W2S * Cfg::W2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(hs + HS * Cfg::H_BYTES);
  uint64_t* a1_full = bars;                       uint64_t* a1_empty = a1_full + 2;
  uint64_t* w1_full = a1_empty + 2;               uint64_t* g1_done = w1_full + 8;     // GEMM1(u) complete: D1 full AND W1 stage free
  uint64_t* w2_full = g1_done + 8;                uint64_t* g2_done = w2_full + 8;     // GEMM2(u) complete: h slot AND W2 stage free
  uint64_t* h_full = g2_done + 4;                 uint64_t* d1_empty = h_full + 4;
  uint64_t* d2_full = d1_empty + 4;               uint64_t* d2_empty = d2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* sb_ab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);   // [5][2C]: every expert
  float* sb_c = sb_ab + 5 * 2 * C;                                                                // [2][C]: this / the next tile's sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-role cycle accounting (compiled in with -DLDMB_MLP_TRACE only; read by tools/trace_mlp.py through ldmb_debug_tc_trace)
#ifdef LDMB_MLP_TRACE
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  auto lap = [&](int slot) { if (trace != nullptr) { const long long now = clock64(); tacc[slot] += now - tlast; tlast = now; } };
#else
  auto lap = [](int) {};
#endif
  if (threadIdx.x == 0) trace_stamp(trace, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::A1_BUFS; ++i) { ptx::mbar_init(&a1_full[i], 1); ptx::mbar_init(&a1_empty[i], 1); }
    for (int i = 0; i < W1S; ++i) { ptx::mbar_init(&w1_full[i], 1); ptx::mbar_init(&g1_done[i], 1); }
    for (int i = 0; i < W2S; ++i) ptx::mbar_init(&w2_full[i], 1);
    for (int i = 0; i < HS; ++i) { ptx::mbar_init(&g2_done[i], 1); ptx::mbar_init(&h_full[i], kEpiWarps * CG); }
    for (int i = 0; i < ND1; ++i) ptx::mbar_init(&d1_empty[i], kEpiWarps * CG);
    ptx::mbar_init(d2_full, 1); ptx::mbar_init(d2_empty, kEpiWarps * CG);
    *s_abort = 0;
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmX); ptx::prefetch_tensormap(&tmWab); ptx::prefetch_tensormap(&tmWc); ptx::prefetch_tensormap(&tmO);
    if (ATT) ptx::prefetch_tensormap(&tmAtt);
  }
  if (warp == 1) {
    if (CG == 2) { ptx::tmem_alloc_2sm(tmem_slot, 512); ptx::tmem_relinquish_2sm(); }
    else { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  }
  // the block's plan entry, biases and weights are older than the previous kernel: read before waiting on it
  // A tile's decisions: word = skip | e1 << 8 | e2 << 16, shared by the batch or looked up per image (per-image plans)
  int shared_word = (a.e1 << 8) | (a.e2 << 16);
  if (a.plan != nullptr) shared_word = (a.plan[0] != 0 ? 1 : 0) | (a.plan[1] << 8) | (a.plan[2] << 16);
  const int t0 = blockIdx.x / CG, t_step = gridDim.x / CG;
  const int n_tiles = (a.plan_img == nullptr && (shared_word & 1)) ? 0 : a.m_tiles;
  const int my_tiles = t0 < n_tiles ? (n_tiles - t0 + t_step - 1) / t_step : 0;
  auto tile_word = [&](int ti) -> int {
    return a.plan_img != nullptr ? a.plan_img[((long long)(t0 + ti * t_step) * (128 * CG)) / a.rows_per_image] : shared_word;
  };
  auto next_active = [&](int ti) -> int {                    // first tile of this CTA at or after ti whose image runs the block
    while (ti < my_tiles && (tile_word(ti) & 1)) ++ti;
    return ti;
  };
  auto c_bias_sum = [&](int word, int i) -> float {          // general + e1 + e2 (modules.py:15)
    return a.b_c[i] + a.b_c[(1 + ((word >> 8) & 0xff)) * C + i] + a.b_c[(1 + ((word >> 16) & 0xff)) * C + i] + (ATT ? a.b_c[5 * C + i] : 0.f);   // + out_proj bias
  };
  for (int i = threadIdx.x; i < 5 * 2 * C; i += kThreads) sb_ab[i] = a.b_ab[i];
  {
    const int first = next_active(0);
    if (first < my_tiles) {
      const int word = tile_word(first);
      for (int i = threadIdx.x; i < C; i += kThreads) sb_c[i] = c_bias_sum(word, i);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool is_producer = warp == 0;          // the producer warp waits for the previous kernel after its weight prefetch
  if (threadIdx.x == 0) trace_stamp(trace, 1);
  if (!is_producer) pdl_wait();

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp, uniform control flow; one elected lane issues)
    {
      const bool issuer = ptx::elect_one();
      Ring ra, r1, r2;
      int n1 = 0, n2 = 0;                                   // next unit (counted over the active tiles of this CTA) of each weight ring
      int w1_ti = next_active(0), w2_ti = w1_ti;            // the tile those units belong to (its image picks the experts)
      bool ok = true;
      auto slot_of = [&](int ti, int e) -> int { return e == 0 ? 0 : 1 + ((tile_word(ti) >> (8 * e)) & 0xff); };
      auto issue_w1 = [&]() {
        const int u = n1 % UNITS, e = u / NKB, j = u % NKB;
        if (!RES && !wait_bar(&g1_done[r1.i], r1.ph ^ 1, s_abort, fault, 22)) { ok = false; return; }
        const int row = slot_of(w1_ti, e) * 2 * C + j * 128 + (int)rank * Cfg::B1_ROWS;
        if (issuer) {
          if (leader) ptx::mbar_arrive_expect_tx(&w1_full[r1.i], Cfg::W1_BYTES * CG);
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            uint8_t* dst = w1 + r1.i * Cfg::W1_BYTES + kb * (Cfg::B1_ROWS * 128);
            if (CG == 2) ptx::tma_load_2d_2sm(dst, &tmWab, &w1_full[r1.i], kb * 64, row);
            else ptx::tma_load_2d(dst, &tmWab, &w1_full[r1.i], kb * 64, row);
          }
        }
        __syncwarp();
        r1.next(W1S);
        if (++n1 % UNITS == 0) w1_ti = next_active(w1_ti + 1);
      };
      auto issue_w2 = [&]() {
        const int u = n2 % UNITS2;
        const bool att_unit = ATT && u >= UNITS;             // out_proj unit: weights = slot 5 of w_c, A operand = the attention output tile
        const int e = u / NKB, j = att_unit ? u - UNITS : u % NKB;
        if (!RES && !wait_bar(&g2_done[r2.i], r2.ph ^ 1, s_abort, fault, 23)) { ok = false; return; }
        const int row = (att_unit ? 5 : slot_of(w2_ti, e)) * C + (int)rank * Cfg::B2_ROWS;
        if (issuer) {
          if (leader) ptx::mbar_arrive_expect_tx(&w2_full[r2.i], Cfg::W2_BYTES * CG);
          if (CG == 2) ptx::tma_load_2d_2sm(w2 + r2.i * Cfg::W2_BYTES, &tmWc, &w2_full[r2.i], j * 64, row);
          else ptx::tma_load_2d(w2 + r2.i * Cfg::W2_BYTES, &tmWc, &w2_full[r2.i], j * 64, row);
          if (att_unit) {
            // the h slot of this unit is filled by TMA instead of the gate epilogue: its barrier still counts the epilogue warps' arrivals
            const int m0w = (t0 + w2_ti * t_step) * (128 * CG) + (int)rank * 128;
            if (leader) { ptx::mbar_arrive_expect_tx(&h_full[r2.i], Cfg::H_BYTES * CG); mbar_arrive_n(&h_full[r2.i], kEpiWarps * CG - 1); }
            if (CG == 2) ptx::tma_load_2d_2sm(hs + r2.i * Cfg::H_BYTES, &tmAtt, &h_full[r2.i], j * 64, m0w);
            else ptx::tma_load_2d(hs + r2.i * Cfg::H_BYTES, &tmAtt, &h_full[r2.i], j * 64, m0w);
          }
        }
        __syncwarp();
        r2.next(W2S);
        if (++n2 % UNITS2 == 0) w2_ti = next_active(w2_ti + 1);
      };
      // fill both weight rings while the previous kernel is still draining; only the xm tiles wait for it
      while (ok && w1_ti < my_tiles && n1 < W1S) issue_w1();
      while (ok && w2_ti < my_tiles && n2 < W2S) issue_w2();      // RES: W1S = W2S = UNITS -- every weight tile, once
      __syncwarp();
      pdl_wait();
      if (issuer) trace_stamp(trace, 2);
      int ta = 0;                                           // ordinal of the tile among this CTA's active tiles
      for (int ti = next_active(0); ok && ti < my_tiles; ti = next_active(ti + 1), ++ta) {
        const int m0 = (t0 + ti * t_step) * (128 * CG) + (int)rank * 128;
        if (!wait_bar(&a1_empty[ra.i], ra.ph ^ 1, s_abort, fault, 21)) break;
        if (issuer) {
          if (leader) ptx::mbar_arrive_expect_tx(&a1_full[ra.i], Cfg::A1_BYTES * CG);
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            uint8_t* dst = a1 + ra.i * Cfg::A1_BYTES + kb * 16384;
            if (CG == 2) ptx::tma_load_2d_2sm(dst, &tmX, &a1_full[ra.i], kb * 64, m0);
            else ptx::tma_load_2d(dst, &tmX, &a1_full[ra.i], kb * 64, m0);
          }
        }
        __syncwarp();
        ra.next(Cfg::A1_BUFS);
        for (int s = 0; !RES && ok && s < UNITS2 + LAG; ++s) {     // same order as the MMA thread consumes
          if (s < UNITS && ta * UNITS + s >= n1) issue_w1();
          if (ok && s >= LAG && ta * UNITS2 + s - LAG >= n2) issue_w2();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA of a pair issues for both)
    if (leader) {
      // the whole warp runs the loop (uniform control flow: descriptors and ring state in uniform registers), one elected lane issues
      const bool issuer = ptx::elect_one();
      constexpr uint32_t idesc1 = ptx::idesc_bf16(128 * CG, 128);
      constexpr uint32_t idesc2 = ptx::idesc_bf16(128 * CG, C);
      auto commit = [&](uint64_t* bar) { if (issuer) { if (CG == 2) ptx::umma_commit_2sm(bar, 3); else ptx::umma_commit(bar); } __syncwarp(); };
      // descriptors: built once per operand tile, (k-block, k) offsets ADDED in 16-byte units (the MMA thread is issue-bound:
      // rebuilding both descriptors per tcgen05.mma cost ~21 SASS instructions each, see kernels_gconv.cu)
      auto mma = [&](uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
        if (!issuer) return;
        if (CG == 2) ptx::umma_f16_2sm(d, adesc, bdesc, idesc, acc);
        else ptx::umma_f16(d, adesc, bdesc, idesc, acc);
      };
      Ring ra, r1, r2, rd1, rd2;
      const uint32_t d2 = tmem_base + Cfg::D2_COL;
      bool ok = true;
      bool first_tile = true;
      for (int ti = next_active(0); ok && ti < my_tiles; ti = next_active(ti + 1)) {
        lap(0);
        if (!wait_bar(&a1_full[ra.i], ra.ph, s_abort, fault, 24)) break;
        lap(5);
        if (first_tile && issuer) trace_stamp(trace, 3);
        ptx::tc_fence_after();
        const uint64_t a1_desc = ptx::smem_desc_sw128(ptx::smem_u32(a1 + ra.i * Cfg::A1_BYTES));
        for (int s = 0; ok && s < UNITS2 + LAG; ++s) {
          if (s < UNITS) {
            lap(0);
            if (!wait_bar(&d1_empty[rd1.i], rd1.ph ^ 1, s_abort, fault, 25)) { ok = false; break; }
            lap(1);
            if ((!RES || first_tile) && !wait_bar(&w1_full[r1.i], RES ? 0u : r1.ph, s_abort, fault, 26)) { ok = false; break; }   // RES: landed once, stays
                                                                                                                                // (a try_wait on a completed barrier still costs ~90 clk)
            lap(2);
            ptx::tc_fence_after();
            const uint32_t d1 = tmem_base + rd1.i * 128;
            const uint64_t b1_desc = ptx::smem_desc_sw128(ptx::smem_u32(w1 + r1.i * Cfg::W1_BYTES));
            if (!(a.dbg & 4)) {
#pragma unroll
              for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma(d1, a1_desc + (kb * 16384 + k * 32) / 16, b1_desc + (kb * (Cfg::B1_ROWS * 128) + k * 32) / 16, idesc1, (kb | k) != 0 ? 1u : 0u);
            }
            commit(&g1_done[r1.i]);                               // -> epilogue (D1 full) and producer (W1 stage free)
            if (s == UNITS - 1) commit(&a1_empty[ra.i]);          // the xm tile has been consumed by all GEMM1 units
            r1.next(W1S);
            rd1.next(ND1);
          }
          if (s >= LAG) {
            const int u = s - LAG;
            if (u == 0) {
              lap(0);
              if (!wait_bar(d2_empty, rd2.ph ^ 1, s_abort, fault, 27)) { ok = false; break; }
              lap(6);
            }
            lap(0);
            if (!wait_bar(&h_full[r2.i], r2.ph, s_abort, fault, 28)) { ok = false; break; }
            lap(3);
            const uint32_t w2i = RES ? (uint32_t)u : r2.i;
            if ((!RES || first_tile) && !wait_bar(&w2_full[w2i], RES ? 0u : r2.ph, s_abort, fault, 29)) { ok = false; break; }
            lap(4);
            ptx::tc_fence_after();
            const uint64_t h_desc = ptx::smem_desc_sw128(ptx::smem_u32(hs + r2.i * Cfg::H_BYTES));
            const uint64_t b2_desc = ptx::smem_desc_sw128(ptx::smem_u32(w2 + w2i * Cfg::W2_BYTES));
            if (!(a.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d2, h_desc + 2 * k, b2_desc + 2 * k, idesc2, (u | k) != 0 ? 1u : 0u);
            }
            commit(&g2_done[r2.i]);                               // -> epilogue (h slot free) and producer (W2 stage free)
            r2.next(HS);
          }
        }
        if (!ok) break;
        commit(d2_full);
        rd2.next(1);
        ra.next(Cfg::A1_BUFS);
        if (first_tile && issuer) trace_stamp(trace, 4);
        first_tile = false;
      }
      if (issuer) trace_stamp(trace, 5);
#ifdef LDMB_MLP_TRACE
      lap(0);
      if (trace != nullptr && lane == 0) for (int i = 0; i < 7; ++i) trace[blockIdx.x * 16 + 16 * 256 + i] = tacc[i];
#endif
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue warps: group = units of that parity; TMEM lane quadrant q, column half chalf
    const int grp = (warp - 2) >> 3, ew = (warp - 2) & 7, q = warp & 3, chalf = ew >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint8_t* slab = hs + ew * Cfg::SLAB_BYTES;              // epilogue-2 staging (group 0): the h ring is idle between tiles
    auto arrive = [&](uint64_t* bar) { if (CG == 2) ptx::mbar_arrive_leader(bar); else ptx::mbar_arrive(bar); };
    Ring rd2;
    bool ok = true;
    int ta = 0;
    for (int ti = next_active(0); ok && ti < my_tiles; ++ta) {
      const int m0 = (t0 + ti * t_step) * (128 * CG) + (int)rank * 128;
      const int word = tile_word(ti);
      const int ti_next = next_active(ti + 1);
      if (ti_next < my_tiles && (int)threadIdx.x - 64 < C)      // the next tile's c biases: visible after this tile's closing barrier
        sb_c[((ta + 1) & 1) * C + threadIdx.x - 64] = c_bias_sum(tile_word(ti_next), threadIdx.x - 64);
      const float* sbc = sb_c + (ta & 1) * C;
      for (int u = grp; u < UNITS; u += kEpiGroups) {
        const int gu = ta * UNITS + u, gu2 = ta * UNITS2 + u;  // gated-unit / GEMM2-unit counters over this CTA's active tiles: ring positions follow from them
        const uint32_t i1 = gu % W1S, p1 = (gu / W1S) & 1, ih = gu2 % HS, ph = (gu2 / HS) & 1, id = gu % ND1;
        const int e = u / NKB, slot = e == 0 ? 0 : 1 + ((word >> (8 * e)) & 0xff);   // row block of the stacked expert biases
        const float* sb = sb_ab + slot * 2 * C + (u % NKB) * 128;                     // [64 a-biases | 64 b-biases] of this unit
        lap(0);
        if (!wait_bar(&g1_done[i1], p1, s_abort, fault, 30)) { ok = false; break; }
        lap(1);
        if (threadIdx.x == 64 && ta == 0 && u == 0) trace_stamp(trace, 6);
        if (!wait_bar(&g2_done[ih], ph ^ 1, s_abort, fault, 31)) { ok = false; break; }   // h slot free
        lap(2);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + lane_off + id * 128;
        if (!(a.dbg & 1)) {
          // row r = q*32 + lane of the [128 x 64] bf16 chunk, 16-byte pieces chalf*4 .. +4, 128B-swizzled (piece ^ (r & 7))
          const uint32_t hrow = ptx::smem_u32(hs + ih * Cfg::H_BYTES) + (q * 32 + lane) * 128;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {                      // 16 of this warp's 32 h columns at a time (register budget of 576 threads)
            const int c0 = chalf * 32 + hh * 16;
            uint32_t ra_[16], rb_[16];
            ptx::tmem_ld_32x16(t_row + c0, ra_);
            ptx::tmem_ld_32x16(t_row + 64 + c0, rb_);
            ptx::tmem_ld_wait();
            // gate: (a + bias_a) * relu(b + bias_b); biases as 16-byte shared loads, adds / multiply as packed f32x2
            float v[16];
            const float4* sa4 = reinterpret_cast<const float4*>(sb + c0);
            const float4* sb4 = reinterpret_cast<const float4*>(sb + 64 + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 ba = sa4[i], bb = sb4[i];
              float2 a01 = add2(make_float2(__uint_as_float(ra_[4 * i]), __uint_as_float(ra_[4 * i + 1])), make_float2(ba.x, ba.y));
              float2 a23 = add2(make_float2(__uint_as_float(ra_[4 * i + 2]), __uint_as_float(ra_[4 * i + 3])), make_float2(ba.z, ba.w));
              float2 g01 = add2(make_float2(__uint_as_float(rb_[4 * i]), __uint_as_float(rb_[4 * i + 1])), make_float2(bb.x, bb.y));
              float2 g23 = add2(make_float2(__uint_as_float(rb_[4 * i + 2]), __uint_as_float(rb_[4 * i + 3])), make_float2(bb.z, bb.w));
              g01.x = fmaxf(g01.x, 0.f); g01.y = fmaxf(g01.y, 0.f); g23.x = fmaxf(g23.x, 0.f); g23.y = fmaxf(g23.y, 0.f);
              a01 = mul2(a01, g01); a23 = mul2(a23, g23);
              v[4 * i] = a01.x; v[4 * i + 1] = a01.y; v[4 * i + 2] = a23.x; v[4 * i + 3] = a23.y;
            }
#pragma unroll
            for (int p = 0; p < 2; ++p)
              ptx::st_shared_v4(hrow + (((chalf * 4 + hh * 2 + p) ^ sw) << 4), pack_bf16(v[8 * p], v[8 * p + 1]), pack_bf16(v[8 * p + 2], v[8 * p + 3]),
                                pack_bf16(v[8 * p + 4], v[8 * p + 5]), pack_bf16(v[8 * p + 6], v[8 * p + 7]));
          }
        }
        lap(3);
        ptx::fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { arrive(&h_full[ih]); arrive(&d1_empty[id]); }
        lap(4);
      }
      if (!ok) break;
      if (grp == 0) {
        // ---- epilogue 2 (group 0): D2 + biases -> 32 fp32 columns per slab (h ring as staging) -> TMA reduce-add into x
        lap(0);
        if (!wait_bar(d2_full, rd2.ph, s_abort, fault, 32)) break;
        lap(5);
        ptx::tc_fence_after();
        const uint32_t t_row2 = tmem_base + lane_off + Cfg::D2_COL;
        const int orow = m0 + q * 32;
#pragma unroll 1
        for (int c0 = chalf * 32; c0 < C; c0 += 64) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_row2 + c0, r);
          ptx::tmem_ld_wait();
          if (lane == 0) ptx::bulk_wait_read<0>();          // the reduce that last used this slab has read it
          __syncwarp();
          const uint32_t srow = ptx::smem_u32(slab) + lane * 128;
#pragma unroll
          for (int p = 0; p < 8; ++p)
            ptx::st_shared_v4(srow + ((p ^ sw) << 4), __float_as_uint(__uint_as_float(r[4 * p]) + sbc[c0 + 4 * p]),
                              __float_as_uint(__uint_as_float(r[4 * p + 1]) + sbc[c0 + 4 * p + 1]),
                              __float_as_uint(__uint_as_float(r[4 * p + 2]) + sbc[c0 + 4 * p + 2]),
                              __float_as_uint(__uint_as_float(r[4 * p + 3]) + sbc[c0 + 4 * p + 3]));
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(a.dbg & 8)) { ptx::tma_reduce_add_2d(&tmO, slab, c0, orow); ptx::bulk_commit(); }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          arrive(d2_empty);
          ptx::bulk_wait_read<0>();                           // staging (h ring) is re-used by the next tile's epilogue 1
        }
        rd2.next(1);
      }
      // every reduce has read its slab before ANY warp of either group writes h rows again (slabs alias the h ring)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps * kEpiGroups) : "memory");
      lap(6);
      if (threadIdx.x == 64 && ta == 0) trace_stamp(trace, 7);
      ti = ti_next;
    }
    if (lane == 0) ptx::bulk_wait_read<0>();                  // smem read by the reduces; the writes complete with the grid
    __syncwarp();
    if (threadIdx.x == 64) trace_stamp(trace, 8);
#ifdef LDMB_MLP_TRACE
    if (trace != nullptr && (threadIdx.x == 64 || threadIdx.x == 64 + 32 * kEpiWarps)) for (int i = 0; i < 7; ++i) trace[blockIdx.x * 16 + 16 * 256 + 8 + i] = tacc[i];
#endif
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) ptx::cluster_sync();
  if (warp == 1) { if (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512); }
  if (threadIdx.x == 0) trace_stamp(trace, 9);
}

template <int C, int CG, bool RES, bool ATT = false>
cudaError_t launch_inst(TcContext* ctx, const CUtensorMap& tmX, const CUtensorMap& tmWab, const CUtensorMap& tmWc, const CUtensorMap& tmO,
                        const CUtensorMap& tmAtt, const MlpArgs& a, cudaStream_t st, int max_ctas) {
  using Cfg = MlpCfg<C, CG, RES, ATT>;
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<C, CG, RES, ATT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr.mark(ctx->device);
  }
  int grid = a.m_tiles * CG < ctx->num_sms ? a.m_tiles * CG : (ctx->num_sms / CG) * CG;
  if (max_ctas >= CG && max_ctas < grid) grid = (max_ctas / CG) * CG;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (g_ldmb_pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
  if (CG == 2) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, mlp_fused_kernel<C, CG, RES, ATT>, tmX, tmWab, tmWc, tmO, tmAtt, a, ctx->fault_dev, ctx->trace_dev);
}

}  // namespace

static int g_mlp_mode = getenv("LDMB_MLP_FUSED") ? atoi(getenv("LDMB_MLP_FUSED")) : 2;   // 0 off, 1 single CTA, 2 CTA pairs

bool mlp_fused_supported(int M, int C) { return g_mlp_mode != 0 && (C == 128 || C == 256) && M >= 1; }

static const bool g_mlp_att = getenv("LDMB_MLP_ATT") == nullptr || atoi(getenv("LDMB_MLP_ATT")) != 0;   // 0: out_proj as a separate K = C GEMM
// the out_proj units need CTA pairs (the single-CTA variant covers a lone 128-row tile of C = 128 only)
bool mlp_fused_att_supported(int M, int C) { return g_mlp_att && mlp_fused_supported(M, C) && g_mlp_mode == 2 && (M > 128 || C == 256); }

// CTAs of a cluster: pairs, except where a lone 128-row tile (C = 128) or -- with per-image decisions -- an image of an odd
// number of 128-row tiles needs the single-CTA variant (C = 128 only).  0 = this shape cannot run fused.
static int mlp_cta_group(int M, int C, int rows_per_image) {
  int cg = (g_mlp_mode == 2 && (M > 128 || C == 256)) ? 2 : 1;   // a lone tile of C = 256 still runs as a pair (rows >= M: zero-filled / clipped)
  if (rows_per_image > 0) {                                     // a tile must lie inside one image
    if (rows_per_image % 128 != 0) return 0;
    if (rows_per_image % (128 * cg) != 0) cg = 1;
  }
  if (cg == 1 && C == 256) return 0;                            // the single-CTA variant exists for C = 128 only
  return cg;
}

bool mlp_fused_per_image_supported(int M, int C, int rows_per_image) {
  return mlp_fused_supported(M, C) && rows_per_image > 0 && M % rows_per_image == 0 && mlp_cta_group(M, C, rows_per_image) != 0;
}

// x fp32 [M,C] += sum_e c_e(a_e(xm) * relu(b_e(xm))) over {general, e1, e2}; xm bf16 [M,C];
// plan_img != NULL: per-image decisions (word per image: skip | e1 << 8 | e2 << 16), rows_per_image = H*W: each tile
// resolves its own image's experts, tiles of images that skip the block do nothing.
// w_ab bf16 [5*2C, C] (per expert: a|b rows interleaved in chunks of 64), b_ab fp32 [5*2C]; w_c bf16 [>=5C, C], b_c fp32 [>=5C].
cudaError_t launch_mlp_fused(TcContext* ctx, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                             float* x, int M, int C, int w_c_rows, const int* plan, int e1, int e2, const int* plan_img, int rows_per_image,
                             cudaStream_t st, int max_ctas, const void* att, long long ld_att) {
  if (!mlp_fused_supported(M, C)) return cudaErrorNotSupported;
  if (plan_img != nullptr && !mlp_fused_per_image_supported(M, C, rows_per_image)) return cudaErrorNotSupported;
  const int cg = mlp_cta_group(M, C, plan_img != nullptr ? rows_per_image : 0);
  if (cg == 0) return cudaErrorNotSupported;
  MlpArgs a;
  a.b_ab = b_ab; a.b_c = b_c; a.plan = plan; a.e1 = e1; a.e2 = e2; a.M = M;
  a.plan_img = plan_img; a.rows_per_image = rows_per_image;
  a.m_tiles = (M + 128 * cg - 1) / (128 * cg);
  a.dbg = tc_knobs().mlp_dbg;
  CUtensorMap tmX, tmWab, tmWc, tmO;
  const cuuint32_t ones[2] = {1, 1};
  auto enc = [&](CUtensorMap* tm, CUtensorMapDataType dt, const void* p, cuuint64_t cols, cuuint64_t rows, int esz, cuuint32_t bc, cuuint32_t br) {
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstr[1] = {cols * esz};
    const cuuint32_t box[2] = {bc, br};
    return ctx->encode(tm, dt, 2, const_cast<void*>(p), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  if (!enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, xm, C, M, 2, 64, 128)) return cudaErrorInvalidValue;
  if (!enc(&tmWab, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_ab, C, 5 * 2 * C, 2, 64, 128 / cg)) return cudaErrorInvalidValue;
  if (!enc(&tmWc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_c, C, w_c_rows, 2, 64, C / cg)) return cudaErrorInvalidValue;
  if (!enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, C, M, 4, 32, 32)) return cudaErrorInvalidValue;
  CUtensorMap tmAtt;
  memset(&tmAtt, 0, sizeof(tmAtt));
  if (att != nullptr) {
    // x += att . W_out^T in the same kernel (attention blocks): att bf16 [M, C] rows of stride ld_att, W_out = rows 5C .. 6C of w_c
    if (!mlp_fused_att_supported(M, C) || plan_img != nullptr || w_c_rows < 6 * C || ld_att % 8 != 0) return cudaErrorNotSupported;
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)M};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld_att * 2};
    const cuuint32_t box[2] = {64, 128};
    if (ctx->encode(&tmAtt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(att), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
    if (C == 128) return launch_inst<128, 2, false, true>(ctx, tmX, tmWab, tmWc, tmO, tmAtt, a, st, max_ctas);
    return launch_inst<256, 2, false, true>(ctx, tmX, tmWab, tmWc, tmO, tmAtt, a, st, max_ctas);
  }
  static const bool resident = getenv("LDMB_MLP_RES") == nullptr || atoi(getenv("LDMB_MLP_RES")) != 0;
  if (C == 128 && cg == 2 && plan_img == nullptr && resident) return launch_inst<128, 2, true>(ctx, tmX, tmWab, tmWc, tmO, tmAtt, a, st, max_ctas);
  if (C == 128) return cg == 2 ? launch_inst<128, 2, false>(ctx, tmX, tmWab, tmWc, tmO, tmAtt, a, st, max_ctas) : launch_inst<128, 1, false>(ctx, tmX, tmWab, tmWc, tmO, tmAtt, a, st, max_ctas);
  return launch_inst<256, 2, false>(ctx, tmX, tmWab, tmWc, tmO, tmAtt, a, st, max_ctas);
}
