// Fused RandomMoE / ReGLU feed-forward of one SwinBlock (modules.py:14-15,34-36; unet.py:44,47) at C = 512 (UNet level 2:
// 18 of the 36 blocks of the default model):
//     x += sum_{e in {general, e1, e2}}  c_e( a_e(xm) * relu(b_e(xm)) )
// as ONE tcgen05 kernel over 8-CTA thread-block clusters.  At C = 512 the accumulator of the c-projection (512 fp32
// columns) fills TMEM on its own, so the single-CTA-pair design of kernels_mlp.cu (C <= 256) does not fit.  Here a cluster
// of FOUR CTA pairs (cta_group::2, 256 pixel rows) shares one row tile:
//   pair p, GEMM1:  its quarter of the 24 hidden chunks (64 hidden units each; chunk u = 4c + p, c = 0..5):
//                   D1[256 x 128 (64 a | 64 b)] = xm_tile[256 x 512] . Wab[chunk rows, 512]^T      (3 TMEM accumulators)
//   gate epilogue:  h_u[256 x 64] = (D1_a + bias_a) * relu(D1_b + bias_b) -> bf16, 128B-swizzled, into the pair's own h ring
//                   AND, by one bulk shared->shared::cluster copy per peer, into the h rings of the other three pairs
//                   (distributed shared memory; the copy completes on the destination CTA's mbarrier)
//   pair p, GEMM2:  its quarter of the OUTPUT columns over ALL 24 chunks:
//                   D2[256 x 128] += h_u . Wc[out cols 128p .. +128, chunk u]^T                    (TMEM)
//   epilogue 2:     D2 + biases -> smem slabs -> TMA fp32 reduce-add into the residual stream x.
// The hidden tensor h (3C bf16 per pixel), the a|b GEMM's output round trip through L2 and the split-K partial sums of the
// two-GEMM path never exist; every CTA streams only its own quarter of Wab and its own column slice of Wc, and the xm tile
// is loaded once and stays resident (128 KB per CTA).  16 row tiles x 8 CTAs = 128 SMs at the config-2 batch (M = 4096).
//
// Flow control of the h ring (slot s = u % 3 in every CTA of the cluster):
//   h_full[s]  (each CTA)    one arrival per chunk, from the gating CTA's epilogue: a local arrive, or a remote arrive.expect_tx
//                            whose byte count the bulk copy into this CTA completes
//   h_peer[s]  (pair leader) the non-leader's relay warp has seen ITS half of the chunk land
//   h_free[s]  (gating pair) 4 arrivals: each pair leader's tcgen05.commit after GEMM2(u), multicast to the two CTAs of the pair that
//                            gates chunk u + 3 -- chunk u has been consumed everywhere, so every copy out of / into slot s is complete.
// Warp roles: 0 TMA producer (weights ring in the MMA warp's static order + the xm tile), 1 MMA issuer (leader) / h relay
// (non-leader), 2..9 epilogue.  Every mbarrier wait carries the watchdog of the other tcgen05 kernels.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tc_context.h"

namespace {

constexpr int C = 512;
constexpr int NKB = C / 64;                 // k-blocks of GEMM1
constexpr int NU = 3 * NKB;                 // hidden chunks of 64 (general + two experts)
constexpr int NP = 4;                       // CTA pairs per cluster
constexpr int CPP = NU / NP;                // chunks gated by one pair
constexpr int N2 = C / NP;                  // output columns of one pair
constexpr int kCluster = 2 * NP;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int WS = 6, HS = 3, ND1 = 3;
constexpr int XM_BYTES = NKB * 16384;       // [128 rows x 128 B] per k-block
constexpr int W_BYTES = 64 * 128;           // one weight stage: this CTA's 64 of the MMA's 128 B rows x one 64-wide k-block
constexpr int H_BYTES = 128 * 128;
constexpr int BAR_BYTES = 512;
constexpr int SMEM_BYTES = 1024 + XM_BYTES + WS * W_BYTES + HS * H_BYTES + BAR_BYTES + N2 * 4;
constexpr int NSTEPS = CPP + NU;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(ND1 * 128 + N2 <= 512, "TMEM columns");

struct FfnArgs {
  const float* b_ab; const float* b_c;
  const int* plan; int e1, e2;          // plan entry {skip, e1, e2, -} of the block, or explicit experts when NULL
  int M;
  int dbg;                              // debug (LDMB_MLP_DBG): 1 = gate epilogue writes nothing, 2 = no GEMM2 MMAs, 4 = no GEMM1 MMAs, 8 = no x update, 16 = no h copies to the peers
  int nsteps;                           // debug: run only the first nsteps of the schedule (NSTEPS = all)
  int* progress;                        // debug: [grid][64] %globaltimer stamps (leader MMA warp: start of step s at [s]; epilogue warp 2: [32 + 4c + k])
  signed char sched[NSTEPS + 2];        // the MMA warp's static order: v >= 0: GEMM1 of this pair's chunk v; v < 0: GEMM2 of chunk -v-1
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
// acquire at cluster scope: the arrival came from another CTA of the cluster
__device__ __forceinline__ bool try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool wait_bar_cluster(uint64_t* bar, uint32_t parity, volatile int* s_abort, int* fault, int code) {
  if (try_wait_cluster(bar, parity)) return true;
  const long long t0 = clock64();
  while (true) {
    if (try_wait_cluster(bar, parity)) return true;
    if (*s_abort) return false;
    if (clock64() - t0 > 3000000000LL) {
      *s_abort = 1;
      report_fault(fault, code);
      return false;
    }
  }
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void arrive_expect_tx_remote(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// bulk copy of this CTA's shared memory into another CTA of the cluster; the bytes complete on an mbarrier of the destination CTA
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(bar_cluster_addr) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void trace_stamp(long long* trace, int slot) {
  if (trace != nullptr && blockIdx.x < 256) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 16 + slot] = t;
  }
}


// debug timeline: low 32 bits of %globaltimer into progress[cta][64] (device memory; LDMB_FFN_PROGRESS = its address)
#define STAMP(idx) do { if (a.progress != nullptr && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.progress[blockIdx.x * 64 + (idx)] = (int)(unsigned)t_; } } while (0)

__global__ void __launch_bounds__(kThreads, 1)
ffn_cluster_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWab,
                   const __grid_constant__ CUtensorMap tmWc, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ FfnArgs a,
                   int* fault, long long* trace) {
  const uint32_t crank = ptx::cluster_ctarank();
  const uint32_t pr = crank >> 1, rk = crank & 1u;          // pair within the cluster, CTA within the pair
  const bool leader = rk == 0;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pr));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* xm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wr = xm + XM_BYTES;
  uint8_t* hs = wr + WS * W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(hs + HS * H_BYTES);
  uint64_t* xm_full = bars;                       // [NKB]  leader
  uint64_t* w_full = xm_full + NKB;               // [WS]   leader
  uint64_t* w_empty = w_full + WS;                // [WS]   both
  uint64_t* d1_full = w_empty + WS;               // [ND1]  both
  uint64_t* d1_empty = d1_full + ND1;             // [ND1]  leader
  uint64_t* h_full = d1_empty + ND1;              // [HS]   each CTA
  uint64_t* h_peer = h_full + HS;                 // [HS]   leader
  uint64_t* h_free = h_peer + HS;                 // [HS]   each CTA
  uint64_t* d2_full = h_free + HS;                // [1]    both
  uint64_t* h_written = d2_full + 1;              // [ND1]  each CTA: the eight epilogue warps have written their rows of the chunk
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_written + ND1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* sb_c = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // [N2]: summed c biases of this pair's columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_stamp(trace, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NKB; ++i) ptx::mbar_init(&xm_full[i], 1);
    for (int i = 0; i < WS; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < ND1; ++i) { ptx::mbar_init(&d1_full[i], 1); ptx::mbar_init(&d1_empty[i], kEpiWarps * 2); }
    for (int i = 0; i < HS; ++i) { ptx::mbar_init(&h_full[i], 1); ptx::mbar_init(&h_peer[i], 1); ptx::mbar_init(&h_free[i], NP); }
    ptx::mbar_init(d2_full, 1);
    for (int i = 0; i < ND1; ++i) ptx::mbar_init(&h_written[i], kEpiWarps);
    *s_abort = 0;
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmX); ptx::prefetch_tensormap(&tmWab); ptx::prefetch_tensormap(&tmWc); ptx::prefetch_tensormap(&tmO);
  }
  if (warp == 1) { ptx::tmem_alloc_2sm(tmem_slot, 512); ptx::tmem_relinquish_2sm(); }
  // the block's plan entry, biases and weights are older than the previous kernel: read before waiting on it
  int e1 = a.e1, e2 = a.e2;
  bool skip = false;
  if (a.plan != nullptr) { skip = a.plan[0] != 0; e1 = a.plan[1]; e2 = a.plan[2]; }
  auto slot_of = [&](int e) -> int { return e == 0 ? 0 : 1 + (e == 1 ? e1 : e2); };     // row block of the stacked expert weights / biases
  if (threadIdx.x < N2) {
    const int n = (int)pr * N2 + threadIdx.x;
    sb_c[threadIdx.x] = a.b_c[n] + a.b_c[(1 + e1) * C + n] + a.b_c[(1 + e2) * C + n];     // general + e1 + e2 (modules.py:15)
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                 // every CTA's barriers are initialised before any peer signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int m0 = (int)(blockIdx.x / kCluster) * 256 + (int)rk * 128;
  const int nsteps = skip ? 0 : a.nsteps;
  int n_gate = 0;                      // chunks this pair gates within the (possibly truncated) schedule
  for (int s = 0; s < nsteps; ++s) n_gate += a.sched[s] >= 0 ? 1 : 0;
  if (threadIdx.x == 0) trace_stamp(trace, 1);
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp, uniform control flow; one elected lane issues)
    const bool issuer = ptx::elect_one();
    uint32_t stage = 0, phase = 0;
    int n = 0;
    bool xm_loaded = false, ok = true;
    auto load_xm = [&]() {
      __syncwarp();
      pdl_wait();                      // xm is the previous kernel's output; everything requested before this line is weights
      if (issuer) {
        trace_stamp(trace, 2);
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          if (leader) ptx::mbar_arrive_expect_tx(&xm_full[kb], 2 * 16384);
          ptx::tma_load_2d_2sm(xm + kb * 16384, &tmX, &xm_full[kb], kb * 64, m0);
        }
      }
      __syncwarp();
      xm_loaded = true;
    };
    auto issue = [&](const CUtensorMap* tm, int col, int row) {
      if (n == WS && !xm_loaded) load_xm();
      if (n >= WS && !wait_bar(&w_empty[stage], phase ^ 1, s_abort, fault, 41)) { ok = false; return; }
      if (issuer) {
        if (leader) ptx::mbar_arrive_expect_tx(&w_full[stage], 2 * W_BYTES);
        ptx::tma_load_2d_2sm(wr + stage * W_BYTES, tm, &w_full[stage], col, row);
      }
      __syncwarp();
      ++n;
      if (++stage == WS) { stage = 0; phase ^= 1; }
    };
    for (int s = 0; ok && s < nsteps; ++s) {
      const int v = a.sched[s];
      if (v >= 0) {                    // GEMM1 of this pair's chunk v: Wab rows of hidden chunk u, all k-blocks
        const int u = NP * v + (int)pr, e = u / NKB, jj = u % NKB;
        const int row = slot_of(e) * 2 * C + jj * 128 + (int)rk * 64;
        for (int kb = 0; ok && kb < NKB; ++kb) issue(&tmWab, kb * 64, row);
      } else {                         // GEMM2 of chunk u: Wc rows = this pair's output columns, k = the chunk's 64 hidden units
        const int u = -v - 1, e = u / NKB, jj = u % NKB;
        issue(&tmWc, jj * 64, slot_of(e) * C + (int)pr * N2 + (int)rk * 64);
      }
    }
    if (nsteps > 0 && !xm_loaded) load_xm();
    if (nsteps == 0) pdl_wait();
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (pair leader) / h relay (non-leader)
    const bool issuer = ptx::elect_one();
    bool ok = true;
    if (leader) {
      constexpr uint32_t idesc = ptx::idesc_bf16(256, 128);
      auto commit = [&](uint64_t* bar, uint16_t mask) { if (issuer) ptx::umma_commit_2sm(bar, mask); __syncwarp(); };
      auto mma = [&](uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t acc) { if (issuer) ptx::umma_f16_2sm(d, adesc, bdesc, idesc, acc); };
      uint32_t stage = 0, phase = 0;
      const uint32_t d2 = tmem_base + ND1 * 128;
      const uint64_t xm_desc = ptx::smem_desc_sw128(ptx::smem_u32(xm));
      bool first_g1 = true, first_g2 = true;
      for (int s = 0; ok && s < nsteps; ++s) {
        const int v = a.sched[s];
        STAMP(s);
        if (v >= 0) {
          const int c = v, id = c % ND1;
          if (c >= ND1 && !wait_bar(&d1_empty[id], ((c / ND1) - 1) & 1, s_abort, fault, 42)) { ok = false; break; }
          const uint32_t d1 = tmem_base + id * 128;
          for (int kb = 0; kb < NKB; ++kb) {
            if (first_g1 && !wait_bar(&xm_full[kb], 0, s_abort, fault, 43)) { ok = false; break; }
            if (!wait_bar(&w_full[stage], phase, s_abort, fault, 44)) { ok = false; break; }
            if (first_g1 && kb == 0 && issuer) trace_stamp(trace, 3);
            ptx::tc_fence_after();
            const uint64_t b_desc = ptx::smem_desc_sw128(ptx::smem_u32(wr + stage * W_BYTES));
            if (!(a.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d1, xm_desc + (kb * 16384 + k * 32) / 16, b_desc + 2 * k, (kb | k) != 0 ? 1u : 0u);
            }
            commit(&w_empty[stage], pair_mask);
            if (++stage == WS) { stage = 0; phase ^= 1; }
          }
          if (!ok) break;
          commit(&d1_full[id], pair_mask);                       // -> the gate epilogues of both CTAs
          first_g1 = false;
        } else {
          const int u = -v - 1, ih = u % HS;
          const uint32_t par = (u / HS) & 1;
          if (!wait_bar(&h_full[ih], par, s_abort, fault, 45)) { ok = false; break; }
          if (!wait_bar_cluster(&h_peer[ih], par, s_abort, fault, 46)) { ok = false; break; }
          if (!wait_bar(&w_full[stage], phase, s_abort, fault, 47)) { ok = false; break; }
          ptx::tc_fence_after();
          const uint64_t h_desc = ptx::smem_desc_sw128(ptx::smem_u32(hs + ih * H_BYTES));
          const uint64_t b_desc = ptx::smem_desc_sw128(ptx::smem_u32(wr + stage * W_BYTES));
          if (!(a.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(d2, h_desc + 2 * k, b_desc + 2 * k, (!first_g2 || k != 0) ? 1u : 0u);
          }
          first_g2 = false;
          commit(&w_empty[stage], pair_mask);
          // consumed here: one of the four arrivals the pair that gates the slot's NEXT chunk (u + HS) waits for.  (Signalling only that
          // pair keeps every h_free phase observed by its waiter -- a parity wait that skips phases passes vacuously.)
          if (u + HS < NU) commit(&h_free[ih], static_cast<uint16_t>(3u << (2 * ((u + HS) % NP))));
          if (++stage == WS) { stage = 0; phase ^= 1; }
        }
      }
      STAMP(NSTEPS);
      if (ok && nsteps == NSTEPS) commit(d2_full, pair_mask);
      if (issuer) trace_stamp(trace, 5);
    } else {
      // the leader's MMAs read this CTA's half of every h chunk: tell it when that half has landed
      for (int s = 0; ok && s < nsteps; ++s) {
        const int v = a.sched[s];
        if (v >= 0) continue;
        STAMP(s);
        const int u = -v - 1, ih = u % HS;
        if (!wait_bar(&h_full[ih], (u / HS) & 1, s_abort, fault, 48)) { ok = false; break; }
        if (issuer) arrive_remote(mapa(ptx::smem_u32(&h_peer[ih]), crank & ~1u));
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue warps: TMEM lane quadrant q, column half chalf
    const int ew = warp - 2, q = warp & 3, chalf = ew >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    bool ok = true;
    for (int c = 0; ok && c < n_gate; ++c) {
      const int u = NP * c + (int)pr, e = u / NKB, jj = u % NKB, id = c % ND1, ih = u % HS;
      const float* bp = a.b_ab + slot_of(e) * 2 * C + jj * 128;        // [64 a-biases | 64 b-biases] of this chunk
      if (!wait_bar(&d1_full[id], (c / ND1) & 1, s_abort, fault, 50)) { ok = false; break; }
      if (threadIdx.x == 64 && c == 0) trace_stamp(trace, 6);
      if (ew == 0) STAMP(32 + 4 * c);
      // slot free: the chunk HS before this one has been consumed by all four pairs (this pair's n-th wait on the slot: c and c - 3 share it)
      if (u >= HS && !wait_bar_cluster(&h_free[ih], (c >= HS && NP * (c - HS) + (int)pr >= HS) ? 1u : 0u, s_abort, fault, 51)) { ok = false; break; }
      if (ew == 0) STAMP(32 + 4 * c + 1);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + lane_off + id * 128;
      if (!(a.dbg & 1)) {
        // row r = q*32 + lane of the [128 x 64] bf16 chunk, 16-byte pieces chalf*4 .. +4, 128B-swizzled (piece ^ (r & 7))
        const uint32_t hrow = ptx::smem_u32(hs + ih * H_BYTES) + (q * 32 + lane) * 128;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int c0 = chalf * 32 + hh * 16;
          uint32_t ra_[16], rb_[16];
          ptx::tmem_ld_32x16(t_row + c0, ra_);
          ptx::tmem_ld_32x16(t_row + 64 + c0, rb_);
          float4 ba[4], bb[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { ba[i] = __ldg(reinterpret_cast<const float4*>(bp + c0) + i); bb[i] = __ldg(reinterpret_cast<const float4*>(bp + 64 + c0) + i); }
          ptx::tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[4 * i] = (__uint_as_float(ra_[4 * i]) + ba[i].x) * fmaxf(__uint_as_float(rb_[4 * i]) + bb[i].x, 0.f);
            v[4 * i + 1] = (__uint_as_float(ra_[4 * i + 1]) + ba[i].y) * fmaxf(__uint_as_float(rb_[4 * i + 1]) + bb[i].y, 0.f);
            v[4 * i + 2] = (__uint_as_float(ra_[4 * i + 2]) + ba[i].z) * fmaxf(__uint_as_float(rb_[4 * i + 2]) + bb[i].z, 0.f);
            v[4 * i + 3] = (__uint_as_float(ra_[4 * i + 3]) + ba[i].w) * fmaxf(__uint_as_float(rb_[4 * i + 3]) + bb[i].w, 0.f);
          }
#pragma unroll
          for (int p = 0; p < 2; ++p)
            ptx::st_shared_v4(hrow + (((chalf * 4 + hh * 2 + p) ^ sw) << 4), pack_bf16(v[8 * p], v[8 * p + 1]), pack_bf16(v[8 * p + 2], v[8 * p + 3]),
                              pack_bf16(v[8 * p + 4], v[8 * p + 5]), pack_bf16(v[8 * p + 6], v[8 * p + 7]));
        }
      }
      ptx::fence_proxy_async();           // generic-proxy smem writes -> visible to the async proxy (tensor core reads, bulk copies)
      ptx::tc_fence_before();
      __syncwarp();
      // (h_written before d1_empty: a warp can only be ND1 chunks ahead of another one's d1_empty arrival, so the ND1 barriers never alias)
      if (lane == 0) { ptx::mbar_arrive(&h_written[id]); ptx::mbar_arrive_leader(&d1_empty[id]); }
      if (ew != 0) continue;
      if (!wait_bar(&h_written[id], (c / ND1) & 1, s_abort, fault, 53)) { ok = false; break; }     // the whole chunk is written
      STAMP(32 + 4 * c + 2);
      if (threadIdx.x == 64) {
        ptx::mbar_arrive(&h_full[ih]);                                          // local consumer (MMA / relay warp of this CTA)
        const uint32_t src = ptx::smem_u32(hs + ih * H_BYTES), bar = ptx::smem_u32(&h_full[ih]);
#pragma unroll
        for (uint32_t p2 = 0; p2 < NP; ++p2) {
          if (p2 == pr || (a.dbg & 16)) continue;
          const uint32_t dst_rank = 2 * p2 + rk;                               // same half of the row tile in the other pairs
          const uint32_t rbar = mapa(bar, dst_rank);
          arrive_expect_tx_remote(rbar, H_BYTES);                              // the chunk's one arrival on the peer's h_full + its byte count
          bulk_copy_to_cluster(mapa(src, dst_rank), src, H_BYTES, rbar);
        }
      }
      __syncwarp();
      STAMP(32 + 4 * c + 3);
    }
    if (ok && nsteps == NSTEPS) {
      // ---- epilogue 2: D2 + biases -> 32 fp32 columns per slab (the xm region is idle once every MMA has completed) -> TMA reduce-add
      if (wait_bar(d2_full, 0, s_abort, fault, 52)) {
        if (threadIdx.x == 64) trace_stamp(trace, 7);
        ptx::tc_fence_after();
        const uint32_t t_row2 = tmem_base + lane_off + ND1 * 128;
        const int orow = m0 + q * 32;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c0 = chalf * 64 + i * 32;
          uint8_t* slab = xm + (ew * 2 + i) * 4096;
          uint32_t r[32];
          ptx::tmem_ld_32x32(t_row2 + c0, r);
          ptx::tmem_ld_wait();
          const uint32_t srow = ptx::smem_u32(slab) + lane * 128;
#pragma unroll
          for (int p = 0; p < 8; ++p)
            ptx::st_shared_v4(srow + ((p ^ sw) << 4), __float_as_uint(__uint_as_float(r[4 * p]) + sb_c[c0 + 4 * p]),
                              __float_as_uint(__uint_as_float(r[4 * p + 1]) + sb_c[c0 + 4 * p + 1]),
                              __float_as_uint(__uint_as_float(r[4 * p + 2]) + sb_c[c0 + 4 * p + 2]),
                              __float_as_uint(__uint_as_float(r[4 * p + 3]) + sb_c[c0 + 4 * p + 3]));
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(a.dbg & 8)) { ptx::tma_reduce_add_2d(&tmO, slab, (int)pr * N2 + c0, orow); ptx::bulk_commit(); }
        }
        if (lane == 0) ptx::bulk_wait_read<0>();              // smem read by the reduces; the writes complete with the grid
        __syncwarp();
      }
    }
    if (threadIdx.x == 64) trace_stamp(trace, 8);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                 // no peer still copies into / signals this CTA
  if (warp == 1) ptx::tmem_dealloc_2sm(tmem_base, 512);
  if (threadIdx.x == 0) trace_stamp(trace, 9);
}

}  // namespace

static const int g_ffn_cluster = getenv("LDMB_FFN_CLUSTER") ? atoi(getenv("LDMB_FFN_CLUSTER")) : 1;     // 0: the two-GEMM path
static const int g_ffn_sched = getenv("LDMB_FFN_SCHED") ? atoi(getenv("LDMB_FFN_SCHED")) : 0;           // experiment: MMA order variants

bool ffn_cluster_supported(int M, int Cc) { return g_ffn_cluster != 0 && Cc == C && M >= 1; }

// x fp32 [M,512] += sum_e c_e(a_e(xm) * relu(b_e(xm))) over {general, e1, e2}; operand layouts as launch_mlp_fused.
cudaError_t launch_ffn_cluster(TcContext* ctx, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                               float* x, int M, int Cc, int w_c_rows, const int* plan, int e1, int e2, cudaStream_t st) {
  if (!ffn_cluster_supported(M, Cc)) return cudaErrorNotSupported;
  FfnArgs a;
  memset(&a, 0, sizeof(a));
  a.b_ab = b_ab; a.b_c = b_c; a.plan = plan; a.e1 = e1; a.e2 = e2; a.M = M;
  a.dbg = tc_knobs().mlp_dbg;
  static const int steps_knob = getenv("LDMB_FFN_STEPS") ? atoi(getenv("LDMB_FFN_STEPS")) : NSTEPS;
  a.nsteps = steps_knob < NSTEPS ? steps_knob : NSTEPS;
  static int* const progress = getenv("LDMB_FFN_PROGRESS") ? reinterpret_cast<int*>(strtoull(getenv("LDMB_FFN_PROGRESS"), nullptr, 0)) : nullptr;
  a.progress = progress;
  {
    // GEMM1 leads by ND1 - 1 chunks; then per round the four chunks of GEMM2 that the cluster gated one round earlier
    int n = 0;
    auto g1 = [&](int c) { a.sched[n++] = (signed char)c; };
    auto g2 = [&](int u) { a.sched[n++] = (signed char)(-u - 1); };
    if (g_ffn_sched == 1) {            // the round's fourth chunk waits for a ring slot: consume it after the next GEMM1
      g1(0); g1(1); g1(2);
      for (int r = 0; r < CPP - 3; ++r) { g2(4 * r); g2(4 * r + 1); g2(4 * r + 2); g1(3 + r); g2(4 * r + 3); }
      for (int u = 4 * (CPP - 3); u < NU; ++u) g2(u);
    } else if (g_ffn_sched == 2) {     // lead of one chunk
      g1(0); g1(1);
      for (int r = 0; r < CPP - 2; ++r) { for (int k = 0; k < 4; ++k) g2(4 * r + k); g1(2 + r); }
      for (int u = 4 * (CPP - 2); u < NU; ++u) g2(u);
    } else {
      g1(0); g1(1); g1(2);
      for (int r = 0; r < CPP - 3; ++r) { for (int k = 0; k < 4; ++k) g2(4 * r + k); g1(3 + r); }
      for (int u = 4 * (CPP - 3); u < NU; ++u) g2(u);
    }
  }
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
  if (!enc(&tmWab, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_ab, C, 5 * 2 * C, 2, 64, 64)) return cudaErrorInvalidValue;
  if (!enc(&tmWc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_c, C, w_c_rows, 2, 64, 64)) return cudaErrorInvalidValue;
  if (!enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, C, M, 4, 32, 32)) return cudaErrorInvalidValue;
  static PerDeviceOnce attr;
  if (attr.need(ctx->device)) {
    cudaError_t e = cudaFuncSetAttribute(ffn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr.mark(ctx->device);
  }
  const int tiles = (M + 255) / 256;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(tiles * kCluster); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (g_ldmb_pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
  at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = kCluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na;
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, ffn_cluster_kernel, tmX, tmWab, tmWc, tmO, a, ctx->fault_dev, ctx->trace_dev);
}
