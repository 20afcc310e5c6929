// libldmb200: C ABI (include/ldmb.h), weight arena, workspaces and the per-call kernel schedules of
// the UNet step (unet.py:89-103 + ddpm.py:76-91) and the VAE decode/encode (vae.py:91-96,122-132).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ldmb.h"
#include "kernels.h"

namespace {

constexpr int kHeadDim = 32;        // unet.py:26
constexpr int kWindow = 6;          // unet.py:26
constexpr int kExperts = 4;         // modules.py:29
constexpr float kNormEps = 1e-4f;   // modules.py:19
constexpr float kLeaky = 0.01f;     // F.leaky_relu default, vae.py:63
constexpr int kStagingSlots = 32;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// Bump arena: slots are registered first, one cudaMalloc backs them all.
struct Arena {
  struct Slot { void** dst; size_t off; };
  std::vector<Slot> slots;
  size_t total = 0;
  void* base = nullptr;
  void add(void** dst, size_t bytes) {
    slots.push_back({dst, total});
    total += (bytes + 255) & ~size_t(255);
  }
  cudaError_t commit() {
    if (base) { cudaFree(base); base = nullptr; }
    cudaError_t e = cudaMalloc(&base, total ? total : 256);
    if (e != cudaSuccess) return e;
    e = cudaMemset(base, 0, total ? total : 256);   // block-diagonal packings rely on zero fill
    if (e != cudaSuccess) return e;
    for (auto& s : slots) *s.dst = static_cast<char*>(base) + s.off;
    return cudaSuccess;
  }
  void release() {
    if (base) cudaFree(base);
    base = nullptr; slots.clear(); total = 0;
  }
};

struct BlockW {
  int level = 0, C = 0, lb = 0, shift = 0;
  bool attn = false;
  // w_c: [5C, C] stacked c-projections (general, experts 0..3) + for attention blocks the MHA out_proj as rows 5C..6C,
  //      so that ffn-c and out_proj run as ONE K-concatenated GEMM with a single residual update
  // w_g: grouped 3x3 conv; C % 64 == 0: block-diagonal pairs of groups [C/64][64][9*64], else per group [C][9*32]
  void *w_ab = nullptr, *w_c = nullptr, *w_g = nullptr, *w_in = nullptr;
  float *b_ab = nullptr, *b_c = nullptr, *b_g = nullptr, *b_in = nullptr;
};

struct LevelW {
  int C = 0, nb = 0;
  void *w1 = nullptr, *w2 = nullptr, *w_down = nullptr, *w_up = nullptr;
  float *b1 = nullptr, *b2 = nullptr, *b_down = nullptr, *b_up = nullptr;
  DevBuf pe;            // [HW, C] fp32
  int peH = 0, peW = 0;
  // workspaces
  DevBuf xs, emb, h1, film, te;
};

struct UNetState {
  bool configured = false;
  ldmb_unet_config cfg{};
  std::vector<BlockW> blocks;       // execution order
  std::vector<LevelW> levels;
  std::map<std::string, int> block_of;   // "e.<lvl>.<b>" / "d.<i>.<b>" -> block index
  float *w_first = nullptr, *b_first = nullptr, *w_last = nullptr, *b_last = nullptr;
  std::set<std::string> missing;
  Arena arena;
  DevBuf xm, hbuf, qkv, pooled, stepbuf, te_pre, backup, pe_tmp;
  const int* plan_img_dev = nullptr;   // per-image plans [n_blocks][B] (skip | e1 << 8 | e2 << 16), NULL in the shared-plan mode
  bool per_image_ws = false;           // workspaces sized for the per-image path (hbuf 6C wide, skip backup)
  // FiLM tables precomputed for a whole schedule (ldmb_unet_precompute_film): film holds n_t = film_nt timesteps
  int film_nt = 0, film_Hs = 0, film_Ws = 0;
  // device views into stepbuf (layout fixed per (B, n_t)): StepParams | plan[n_blocks][4] | t_index[B] | te tables
  const StepParams* sp_dev = nullptr; const int* plan_dev = nullptr; const int* tindex_dev = nullptr;
  struct GraphEntry { int B, Hs, Ws, n_t, pre, per_image; unsigned long long epoch; cudaGraphExec_t exec; long long launches; int seen; };
  std::vector<GraphEntry> graphs;
  unsigned long long ws_epoch = 0;     // bumped whenever a workspace is reallocated (cached graphs hold raw pointers)
  bool use_graphs = true;
  cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the caller's may be the legacy NULL stream)
  // The grouped conv of a block only needs xm and only adds into x: it runs on a side stream (a forked branch of the
  // captured graph) next to the block's GEMM chain, so its CTAs fill the SMs the GEMMs' partial last waves leave idle.
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  const bool no_fork_env = getenv("LDMB_NO_FORK") != nullptr;      // read once (handle creation), like every knob
  bool fork_conv = !no_fork_env;
  bool deterministic = false;          // ldmb_set_deterministic: no split-K, no two concurrent updaters of x
  // where the forked conv runs in the two-GEMM blocks (C >= 512): 1 (default) = from the norm on, beside the a|b GEMM; 2 = beside
  // the c-projection only, capped to the SMs that GEMM's grid leaves idle -- measured 22 % SLOWER end to end (16 CTAs walk 32 tiles
  // each, the block waits for them), kept as an experiment knob (profiles/r2_experiment_conv_fork_placement.txt);
  // 3 = fork only in the fused feed-forward blocks (C <= 256), 4 = only in the two-GEMM blocks
  int fork_mode = getenv("LDMB_FORK_MODE") ? atoi(getenv("LDMB_FORK_MODE")) : 5;     // 5: SM-partitioned (run_block); 1: plain early fork; 2: late fork; 3 / 4: fork only the fused / two-GEMM blocks
  int attn_part = getenv("LDMB_ATTN_PART") ? atoi(getenv("LDMB_ATTN_PART")) : 0;   // attention blocks at C >= 512, partitioned mode: 1 = conv beside a capped in-projection GEMM (a|b GEMM then uncapped)
  int attn_fused_fork = getenv("LDMB_ATTN_FUSED_FORK") ? atoi(getenv("LDMB_ATTN_FUSED_FORK")) : 2;   // attention blocks at C <= 256: 0 = conv serial, 1 = forked at the start of the block, 2 = forked after the attention core
  int fork_fused = getenv("LDMB_FORK_FUSED") ? atoi(getenv("LDMB_FORK_FUSED")) : 3;   // fork the conv beside the fused feed-forward: bit 0 at C = 128, bit 1 at C = 256
  // fork mode 5, fused feed-forward blocks (C <= 256): the conv is split over two launches -- the first conv_split / 1000 of its tiles
  // run on the split_free SMs the (capped) feed-forward kernel leaves, the rest on the whole machine afterwards.  0 = no split.
  int conv_split128 = getenv("LDMB_CONV_SPLIT128") ? atoi(getenv("LDMB_CONV_SPLIT128")) : 0;     // measured slower (99.6 vs 98.9 ms per 50 steps at 220 / 260): off
  int conv_split256 = getenv("LDMB_CONV_SPLIT256") ? atoi(getenv("LDMB_CONV_SPLIT256")) : 0;
  int split_free = getenv("LDMB_SPLIT_FREE") ? atoi(getenv("LDMB_SPLIT_FREE")) : 20;
  int part_min_free = getenv("LDMB_PART_MIN_FREE") ? atoi(getenv("LDMB_PART_MIN_FREE")) : 48;   // fork mode 5: fewest SMs worth giving the conv
  // pinned staging ring for the per-call host tables
  char* staging = nullptr;
  size_t staging_slot_bytes = 0;
  cudaEvent_t staging_ev[kStagingSlots]{};
  bool staging_used[kStagingSlots]{};
  int staging_next = 0;
};

struct ResW { void *w1 = nullptr, *w2 = nullptr; float *b1 = nullptr, *b2 = nullptr; };
struct VaeLevelW {
  int C = 0;
  std::vector<ResW> res;
  float *w_rgb = nullptr, *b_rgb = nullptr;      // decoder to_rgb [img_ch][C]
  void* w_resample = nullptr; float* b_resample = nullptr;   // decoder: ConvT into this level; encoder: 1x1 out of this level
};
struct VaeState {
  bool configured = false;
  ldmb_vae_config cfg{};
  std::vector<VaeLevelW> levels;
  float *w_in = nullptr, *b_in = nullptr, *w_out = nullptr, *b_out = nullptr;
  std::set<std::string> missing;
  Arena arena;
  DevBuf act[3], rgb[2];
};

}  // namespace

enum { PK_FFN_AB = 0, PK_FFN_C = 1, PK_QKV = 2, PK_ENC = 3, PK_LEVEL = 4, PK_GCONV = 5, PK_VAE_CONV = 6, PK_VAE_GEMM = 7,
       PK_GEMM_SIMT = 8, PK_NORM = 9, PK_ATTN = 10, PK_EDGE = 11, PK_OTHER = 12, PK_COUNT = 13 };
static_assert(PK_COUNT == LDMB_PROFILE_CLASSES, "profile classes");
struct ProfRec { int kind; double work; cudaEvent_t a, b; };

struct ldmb_handle {
  bool prof_on = false;
  uint32_t skip_mask = 0;      // debug: classes whose launches are dropped (ldmb_debug_skip_classes)
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  cudaEvent_t get_event() {
    if (ev_next == ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
    return ev_pool[ev_next++];
  }
  int device = 0;
  int precision = LDMB_BF16;
  bool force_simt = false;
  long long launches = 0;
  TcContext* tc = nullptr;
  char err[512] = {0};
  UNetState unet;
  VaeState vae[2];
  bool bf16() const { return precision == LDMB_BF16; }
  size_t tsize() const { return precision == LDMB_BF16 ? 2 : 4; }
};

namespace {

int fail(ldmb_handle* h, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h->err, sizeof(h->err), fmt, ap);
  va_end(ap);
  return code;
}

// A pipeline watchdog fired in an earlier kernel of this handle: every later compute call refuses with LDMB_ERR_KERNEL
// (the results since then are garbage).  Read from the host-mapped mirror of the fault word: no synchronisation.
#define CHECK_FAULT()                                                                              \
  do {                                                                                             \
    const int f__ = tc_poll_fault(h->tc);                                                          \
    if (f__ != 0) return fail(h, LDMB_ERR_KERNEL, "a tcgen05 pipeline watchdog fired earlier on this handle (code %d)", f__); \
  } while (0)

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(h, LDMB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// A kernel launch on stream `st`: counted, and bracketed by events when profiling is on.
#define CKLP(KIND_, WORK_, call)                                                                     \
  do {                                                                                             \
    if (h->skip_mask & (1u << (KIND_))) break;                                                     \
    ProfRec pr__;                                                                                  \
    const bool p__ = h->prof_on;                                                                   \
    if (p__) { pr__.kind = (KIND_); pr__.work = (double)(WORK_); pr__.a = h->get_event(); pr__.b = h->get_event(); cudaEventRecord(pr__.a, st); } \
    CK(call);                                                                                      \
    h->launches++;                                                                                 \
    if (p__) { cudaEventRecord(pr__.b, st); h->prof.push_back(pr__); }                             \
  } while (0)
#define CKL(call) CKLP(PK_OTHER, 0, call)

int ensure(ldmb_handle* h, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return LDMB_OK;
  h->unet.ws_epoch++;
  if (b.p) { CK(cudaDeviceSynchronize()); CK(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
  CK(cudaMalloc(&b.p, bytes ? bytes : 256));
  b.bytes = bytes;
  return LDMB_OK;
}

void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.bytes = 0;
}

int gemm(ldmb_handle* h, const GemmDesc& d, cudaStream_t st, int kind, bool force_simt = false) {
  const double flops = 2.0 * d.M * (double)d.N * d.K * (d.batch > 0 ? d.batch : 1);
  if (h->bf16() && !h->force_simt && !force_simt && tc_supported(d))
    CKLP(kind, flops, launch_gemm_tc(h->tc, d, st));
  else CKLP(PK_GEMM_SIMT, flops, launch_gemm_simt(d, h->bf16(), st));
  return LDMB_OK;
}

GemmDesc gd() {
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.amode = AM_ROWS;
  d.batch = 1;
  d.glu_chunk = 64;
  d.slope = kLeaky;
  return d;
}

// Window attention core (attention.py:13-85 + torch MHA): tcgen05 kernel where the shape allows, else the mma.sync / CUDA-core ones.
// mode: 0 default, 1 CUDA-core kernel, 2 mma.sync kernel (tests compare the three)
int window_attention(ldmb_handle* h, const void* qkv, const void* xm, const float* b_in, void* att, long long ldo, int B, int Hl, int Wl,
                     int C, int wh, int ww, int shift, const int* pl, cudaStream_t st, int mode) {
  const double bytes = (double)B * Hl * Wl * C * 4 * h->tsize();
  if (h->bf16() && !h->force_simt && mode == 0 && window_attention_tc_supported(B, Hl, Wl, C, kHeadDim, wh, ww, ldo)) {
    CKLP(PK_ATTN, bytes, launch_window_attention_tc(h->tc, qkv, xm, b_in, att, ldo, B, Hl, Wl, C, wh, ww, shift, pl, st));
    return LDMB_OK;
  }
  CKLP(PK_ATTN, bytes, launch_window_attention(qkv, xm, b_in, att, ldo, h->bf16(), B, Hl, Wl, C, kHeadDim, wh, ww, shift, pl, st,
                                               mode == 1 || h->force_simt));
  return LDMB_OK;
}

// x fp32 [B,H,W,C] += grouped conv3x3(xm) + bias (unet.py:30: groups of 32 channels).
// tcgen05, halo-patch kernel (every activation read once) when C % 64 == 0, else the generic implicit-GEMM path.
int grouped_conv(ldmb_handle* h, const void* xm, const void* w_g, const float* b_g, float* x, int B, int Hl, int Wl, int C,
                 const int* pl, cudaStream_t st, bool force_generic, int max_ctas = 0, int part = 0, int split_permille = 0) {
  const int M = B * Hl * Wl;
  if (h->bf16() && !h->force_simt && !force_generic && gconv_halo_supported(B, Hl, Wl, C)) {
    const double share = part == 1 ? split_permille / 1000.0 : (part == 2 ? 1.0 - split_permille / 1000.0 : 1.0);
    CKLP(PK_GCONV, share * 2.0 * M * (double)C * 9 * kHeadDim, launch_gconv_halo(h->tc, xm, w_g, b_g, x, B, Hl, Wl, C, pl, st, max_ctas, part, split_permille));
    return LDMB_OK;
  }
  GemmDesc d = gd();
  const int gw = (C % 64 == 0) ? 64 : kHeadDim;          // channels per launch-batch: a pair of groups, or one group
  d.A = xm; d.lda = C; d.amode = AM_CONV3; d.cH = Hl; d.cW = Wl; d.cC = gw;
  d.W = w_g; d.ldw = 9 * gw; d.bias = b_g; d.out = x; d.ldo = C;
  d.M = M; d.N = gw; d.K = 9 * gw; d.epi = EPI_ACCUM_F32; d.plan = pl;
  d.batch = C / gw; d.a_koff_b = gw; d.w_row_b = gw; d.out_off_b = gw; d.bias_off_b = gw;
  return gemm(h, d, st, PK_GCONV);
}

// ---------------------------------------------------------------- repack helpers
int repack(ldmb_handle* h, const float* src, void* dst, bool to_t, int d0, int d1, int d2, long long s0, long long s1,
           long long s2, long long t0, long long t1, long long t2, cudaStream_t st) {
  const int dims[4] = {d0, d1, d2, 1};
  const long long ss[4] = {s0, s1, s2, 0}, ds[4] = {t0, t1, t2, 0};
  CKL(launch_repack(src, dst, to_t && h->bf16(), dims, ss, ds, st));
  return LDMB_OK;
}
int copy_t(ldmb_handle* h, const float* src, void* dst, long long n, cudaStream_t st) {   // cast-copy into T
  return repack(h, src, dst, true, 1, 1, (int)n, 0, 0, 1, 0, 0, 1, st);
}
int copy_f(ldmb_handle* h, const float* src, float* dst, long long n, cudaStream_t st) {
  return repack(h, src, dst, false, 1, 1, (int)n, 0, 0, 1, 0, 0, 1, st);
}
char* toff(ldmb_handle* h, void* base, long long elems) { return static_cast<char*>(base) + elems * (long long)h->tsize(); }

bool shape_is(const int64_t* shape, int ndim, std::initializer_list<int64_t> want) {
  // trailing 1s are optional (conv weights are [O,I,1,1])
  size_t i = 0;
  for (int64_t w : want) {
    if ((int)i >= ndim) { if (w != 1) return false; }
    else if (shape[i] != w) return false;
    ++i;
  }
  for (; (int)i < ndim; ++i) if (shape[i] != 1) return false;
  return true;
}

int glu_chunk_for(int C) { return C % 64 == 0 ? 64 : 32; }

}  // namespace

// =====================================================================================
// handle
// =====================================================================================
extern "C" int ldmb_abi_version(void) { return LDMB_ABI_VERSION; }

extern "C" int ldmb_create(int device, int precision, ldmb_handle** out) {
  if (!out) return LDMB_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return LDMB_ERR_CUDA;
  if (precision != LDMB_BF16 && precision != LDMB_FP32_VALIDATE) return LDMB_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return LDMB_ERR_CUDA;
  ldmb_handle* h = new ldmb_handle();
  h->device = device;
  h->precision = precision;
  h->tc = tc_context_create(device, h->err, sizeof(h->err));
  if (!h->tc) {
    fprintf(stderr, "ldmb_create: %s\n", h->err);
    delete h;
    return LDMB_ERR_CUDA;
  }
  if (getenv("LDMB_DETERMINISTIC")) ldmb_set_deterministic(h, 1);
  *out = h;
  return LDMB_OK;
}

extern "C" void ldmb_destroy(ldmb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  UNetState& u = h->unet;
  u.arena.release();
  for (auto& l : u.levels) { release(l.pe); release(l.xs); release(l.emb); release(l.h1); release(l.film); release(l.te); }
  release(u.xm); release(u.hbuf); release(u.qkv); release(u.pooled); release(u.stepbuf); release(u.te_pre); release(u.backup); release(u.pe_tmp);
  for (auto& g : u.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (u.cap_stream) cudaStreamDestroy(u.cap_stream);
  if (u.side_stream) cudaStreamDestroy(u.side_stream);
  if (u.ev_fork) cudaEventDestroy(u.ev_fork);
  if (u.ev_join) cudaEventDestroy(u.ev_join);
  if (u.staging) {
    cudaFreeHost(u.staging);
    for (int i = 0; i < kStagingSlots; ++i) if (u.staging_ev[i]) cudaEventDestroy(u.staging_ev[i]);
  }
  for (auto& v : h->vae) { v.arena.release(); for (auto& b : v.act) release(b); for (auto& b : v.rgb) release(b); }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  tc_context_destroy(h->tc);
  delete h;
}

extern "C" const char* ldmb_last_error(const ldmb_handle* h) { return h ? h->err : "null handle"; }
extern "C" int ldmb_precision_of(const ldmb_handle* h) { return h ? h->precision : -1; }
extern "C" int ldmb_set_force_simt(ldmb_handle* h, int on) {
  if (!h) return LDMB_ERR_INVALID;
  h->force_simt = on != 0;
  h->unet.ws_epoch++;          // captured graphs hold the old launch set
  return LDMB_OK;
}
extern "C" int64_t ldmb_launch_count(const ldmb_handle* h) { return h ? h->launches : 0; }
extern "C" int ldmb_debug_tc_trace(ldmb_handle* h, int enable, int64_t* stamps_host, int max_ctas) {
  if (!h) return -1;
  if (stamps_host == nullptr) return tc_trace_enable(h->tc, enable);
  return tc_trace_read(h->tc, reinterpret_cast<long long*>(stamps_host), max_ctas);
}
extern "C" int ldmb_profile_begin(ldmb_handle* h) {
  if (!h) return LDMB_ERR_INVALID;
  h->prof.clear();
  h->ev_next = 0;
  h->prof_on = true;
  return LDMB_OK;
}
extern "C" int ldmb_profile_end(ldmb_handle* h, double* ms, double* work, int64_t* launches) {
  if (!h || !ms || !work || !launches) return LDMB_ERR_INVALID;
  h->prof_on = false;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  for (int k = 0; k < PK_COUNT; ++k) { ms[k] = 0; work[k] = 0; launches[k] = 0; }
  for (const ProfRec& r : h->prof) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.kind] += t; work[r.kind] += r.work; launches[r.kind] += 1;
  }
  h->prof.clear();
  h->ev_next = 0;
  return LDMB_OK;
}
extern "C" int ldmb_debug_skip_classes(ldmb_handle* h, uint32_t mask) {
  if (!h) return LDMB_ERR_INVALID;
  h->skip_mask = mask;
  h->unet.ws_epoch++;          // captured graphs hold the old launch set
  return LDMB_OK;
}
extern "C" int ldmb_check_device_fault(ldmb_handle* h, void* stream) {
  if (!h) return -1;
  return tc_read_fault(h->tc, static_cast<cudaStream_t>(stream));
}
extern "C" int ldmb_poll_device_fault(const ldmb_handle* h) { return h ? tc_poll_fault(h->tc) : -1; }

// =====================================================================================
// UNet: configure + parameters
// =====================================================================================
extern "C" int ldmb_unet_configure(ldmb_handle* h, const ldmb_unet_config* cfg) {
  if (!h || !cfg) return LDMB_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const int S = cfg->num_levels;
  if (S < 1 || S > LDMB_MAX_LEVELS || cfg->input_channels < 1 || cfg->stem_size < 1)
    return fail(h, LDMB_ERR_INVALID, "unet config: bad level count / channels / stem");
  for (int l = 0; l < S; ++l) {
    if (cfg->channels[l] % kHeadDim || cfg->channels[l] <= 0 || cfg->channels[l] > 2048)
      return fail(h, LDMB_ERR_INVALID, "unet config: channels[%d]=%d must be a multiple of %d (head_dim/group size, unet.py:26,30) and <= 2048",
                  l, cfg->channels[l], kHeadDim);
    if (cfg->blocks[l] < 1) return fail(h, LDMB_ERR_INVALID, "unet config: stages[%d] must be >= 1", l);
  }
  if (cfg->input_channels * cfg->stem_size * cfg->stem_size > 32)
    return fail(h, LDMB_ERR_UNSUPPORTED, "input_channels*stem_size^2 > 32 is not supported");
  UNetState& u = h->unet;
  if (u.configured) {
    CK(cudaDeviceSynchronize());
    u.arena.release();
    for (auto& l : u.levels) { release(l.pe); release(l.xs); release(l.emb); release(l.h1); release(l.film); release(l.te); }
  }
  u.cfg = *cfg;
  u.blocks.clear(); u.levels.clear(); u.block_of.clear(); u.missing.clear();
  u.levels.resize(S);
  for (int l = 0; l < S; ++l) { u.levels[l].C = cfg->channels[l]; u.levels[l].nb = 2 * cfg->blocks[l]; }
  char key[64];
  // execution order: encoder levels 0..S-1, decoder levels S-1..0 (unet.py:92-101); decoder_stages index i = S-1-level
  for (int l = 0; l < S; ++l)
    for (int b = 0; b < cfg->blocks[l]; ++b) {
      BlockW w; w.level = l; w.C = cfg->channels[l]; w.lb = b; w.attn = false; w.shift = (b % 2 == 0) ? kWindow / 2 : 0;
      snprintf(key, sizeof(key), "e.%d.%d", l, b);
      u.block_of[key] = (int)u.blocks.size();
      u.blocks.push_back(w);
    }
  for (int i = 0; i < S; ++i) {
    const int l = S - 1 - i, n = cfg->blocks[l];
    for (int b = 0; b < n; ++b) {
      BlockW w; w.level = l; w.C = cfg->channels[l]; w.lb = n + b; w.attn = b >= n - 2; w.shift = (b % 2 == 0) ? kWindow / 2 : 0;
      snprintf(key, sizeof(key), "d.%d.%d", i, b);
      u.block_of[key] = (int)u.blocks.size();
      u.blocks.push_back(w);
    }
  }
  // arena layout + the list of state_dict entries we need
  const size_t ts = h->tsize();
  Arena& a = u.arena;
  const int J = cfg->input_channels * cfg->stem_size * cfg->stem_size;
  a.add((void**)&u.w_first, (size_t)cfg->channels[0] * J * 4); a.add((void**)&u.b_first, (size_t)cfg->channels[0] * 4);
  a.add((void**)&u.w_last, (size_t)cfg->channels[0] * J * 4);  a.add((void**)&u.b_last, (size_t)cfg->input_channels * 4);
  for (const char* n : {"encoder_first.weight", "encoder_first.bias", "decoder_last.weight", "decoder_last.bias"}) u.missing.insert(n);
  for (int l = 0; l < S; ++l) {
    LevelW& L = u.levels[l];
    const size_t C = L.C;
    a.add(&L.w1, (size_t)L.nb * 4 * C * 2 * C * ts); a.add((void**)&L.b1, (size_t)L.nb * 4 * C * 4);
    a.add(&L.w2, (size_t)L.nb * 2 * C * 4 * C * ts); a.add((void**)&L.b2, (size_t)L.nb * 2 * C * 4);
    if (l < S - 1) {
      const size_t Cn = cfg->channels[l + 1];
      a.add(&L.w_down, Cn * C * ts); a.add((void**)&L.b_down, Cn * 4);
      a.add(&L.w_up, C * Cn * ts);   a.add((void**)&L.b_up, C * 4);
      char nm[96];
      for (const char* sfx : {"weight", "bias"}) {
        snprintf(nm, sizeof(nm), "encoder_stages.%d.ch_conv.0.%s", l, sfx); u.missing.insert(nm);
        snprintf(nm, sizeof(nm), "decoder_stages.%d.ch_conv.1.%s", S - 1 - l, sfx); u.missing.insert(nm);
      }
    }
  }
  for (auto& kv : u.block_of) {
    BlockW& w = u.blocks[kv.second];
    const size_t C = w.C;
    a.add(&w.w_ab, 5 * 2 * C * C * ts); a.add((void**)&w.b_ab, 5 * 2 * C * 4);
    const size_t nc = w.attn ? 6 : 5;
    a.add(&w.w_c, nc * C * C * ts);     a.add((void**)&w.b_c, nc * C * 4);
    a.add(&w.w_g, C * (C % 64 == 0 ? 576 : 288) * ts); a.add((void**)&w.b_g, C * 4);
    int si, bi; char ed;
    sscanf(kv.first.c_str(), "%c.%d.%d", &ed, &si, &bi);
    char pre[96];
    snprintf(pre, sizeof(pre), "%s_stages.%d.stage.blocks.%d.", ed == 'e' ? "encoder" : "decoder", si, bi);
    auto need = [&](const std::string& s) { u.missing.insert(std::string(pre) + s); };
    for (const char* m : {"a", "b", "c"})
      for (const char* sfx : {"weight", "bias"}) {
        need(std::string("ffn.general.") + m + "." + sfx);
        for (int e = 0; e < kExperts; ++e) need("ffn.experts." + std::to_string(e) + "." + m + "." + sfx);
      }
    for (const char* sfx : {"weight", "bias"}) {
      need(std::string("conv.") + sfx);
      need(std::string("encodings.proj1.") + sfx);
      need(std::string("encodings.proj2.") + sfx);
    }
    if (w.attn) {
      a.add(&w.w_in, 3 * C * C * ts); a.add((void**)&w.b_in, 3 * C * 4);
      for (const char* n : {"in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias"})
        need(std::string("self_attention.attention.") + n);
    }
  }
  CK(a.commit());
  u.configured = true;
  return LDMB_OK;
}

extern "C" int ldmb_unet_params_missing(const ldmb_handle* h) {
  if (!h || !h->unet.configured) return -1;
  return (int)h->unet.missing.size();
}

extern "C" int ldmb_unet_load_param(ldmb_handle* h, const char* name, const float* src, const int64_t* shape, int ndim,
                                    void* stream) {
  if (!h || !name || !src || !shape) return LDMB_ERR_INVALID;
  UNetState& u = h->unet;
  if (!u.configured) return fail(h, LDMB_ERR_STATE, "ldmb_unet_configure has not been called");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ldmb_unet_config& cfg = u.cfg;
  const int S = cfg.num_levels, s = cfg.stem_size, Cin = cfg.input_channels, C0 = cfg.channels[0];
  u.film_nt = 0;                       // FiLM tables depend on the Encodings weights
  std::string nm(name);
  if (nm.rfind("model.", 0) == 0) nm = nm.substr(6);
  auto bad_shape = [&]() { return fail(h, LDMB_ERR_INVALID, "size mismatch for %s", name); };
  auto done = [&]() { u.missing.erase(nm); return (int)LDMB_OK; };
  int rc;

  if (nm == "encoder_first.weight") {
    if (!shape_is(shape, ndim, {C0, Cin, s, s})) return bad_shape();
    if ((rc = copy_f(h, src, u.w_first, (long long)C0 * Cin * s * s, st))) return rc;
    return done();
  }
  if (nm == "encoder_first.bias") { if (!shape_is(shape, ndim, {C0})) return bad_shape(); if ((rc = copy_f(h, src, u.b_first, C0, st))) return rc; return done(); }
  if (nm == "decoder_last.weight") {   // ConvTranspose2d: [in=C0, out=Cin, s, s]
    if (!shape_is(shape, ndim, {C0, Cin, s, s})) return bad_shape();
    if ((rc = copy_f(h, src, u.w_last, (long long)C0 * Cin * s * s, st))) return rc;
    return done();
  }
  if (nm == "decoder_last.bias") { if (!shape_is(shape, ndim, {Cin})) return bad_shape(); if ((rc = copy_f(h, src, u.b_last, Cin, st))) return rc; return done(); }

  int si = -1, bi = -1, consumed = 0;
  bool dec = false;
  if (sscanf(nm.c_str(), "encoder_stages.%d.%n", &si, &consumed) == 1 && consumed > 0) dec = false;
  else if (sscanf(nm.c_str(), "decoder_stages.%d.%n", &si, &consumed) == 1 && consumed > 0) dec = true;
  else return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
  if (si < 0 || si >= S) return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
  std::string rest = nm.substr(consumed);
  const int lvl = dec ? S - 1 - si : si;

  if (rest.rfind("ch_conv.", 0) == 0) {
    if (lvl >= S - 1) return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
    LevelW& L = u.levels[lvl];
    const int C = L.C, Cn = cfg.channels[lvl + 1];
    if (!dec && rest == "ch_conv.0.weight") { if (!shape_is(shape, ndim, {Cn, C})) return bad_shape(); if ((rc = copy_t(h, src, L.w_down, (long long)Cn * C, st))) return rc; return done(); }
    if (!dec && rest == "ch_conv.0.bias") { if (!shape_is(shape, ndim, {Cn})) return bad_shape(); if ((rc = copy_f(h, src, L.b_down, Cn, st))) return rc; return done(); }
    if (dec && rest == "ch_conv.1.weight") { if (!shape_is(shape, ndim, {C, Cn})) return bad_shape(); if ((rc = copy_t(h, src, L.w_up, (long long)C * Cn, st))) return rc; return done(); }
    if (dec && rest == "ch_conv.1.bias") { if (!shape_is(shape, ndim, {C})) return bad_shape(); if ((rc = copy_f(h, src, L.b_up, C, st))) return rc; return done(); }
    return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
  }
  consumed = 0;
  if (sscanf(rest.c_str(), "stage.blocks.%d.%n", &bi, &consumed) != 1 || consumed == 0)
    return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
  char key[64];
  snprintf(key, sizeof(key), "%c.%d.%d", dec ? 'd' : 'e', si, bi);
  auto it = u.block_of.find(key);
  if (it == u.block_of.end()) return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
  BlockW& w = u.blocks[it->second];
  LevelW& L = u.levels[w.level];
  const int C = w.C, G = glu_chunk_for(C);
  rest = rest.substr(consumed);

  if (rest.rfind("cross_attention.", 0) == 0) return LDMB_OK;   // dead code in the reference (attention.py:92-98)

  if (rest.rfind("ffn.", 0) == 0) {
    int e = -1; char m = 0; char kind[16] = {0};
    if (sscanf(rest.c_str(), "ffn.general.%c.%15s", &m, kind) == 2) e = 0;
    else if (sscanf(rest.c_str(), "ffn.experts.%d.%c.%15s", &e, &m, kind) == 3 && e >= 0 && e < kExperts) e += 1;
    else return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
    const bool is_w = strcmp(kind, "weight") == 0;
    if (!is_w && strcmp(kind, "bias") != 0) return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
    if (is_w ? !shape_is(shape, ndim, {C, C}) : !shape_is(shape, ndim, {C})) return bad_shape();
    if (m == 'a' || m == 'b') {
      // rows of a and b interleaved in chunks of G so one accumulator tile holds matching a|b columns
      const long long base = (long long)e * 2 * C + (m == 'b' ? G : 0);
      if (is_w) rc = repack(h, src, toff(h, w.w_ab, base * C), true, C / G, G, C, (long long)G * C, C, 1, 2LL * G * C, C, 1, st);
      else rc = repack(h, src, w.b_ab + base, false, C / G, G, 1, G, 1, 0, 2 * G, 1, 0, st);
    } else if (m == 'c') {
      if (is_w) rc = copy_t(h, src, toff(h, w.w_c, (long long)e * C * C), (long long)C * C, st);
      else rc = copy_f(h, src, w.b_c + (long long)e * C, C, st);
    } else return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
    if (rc) return rc;
    return done();
  }
  if (rest == "conv.weight") {   // [C, 32, 3, 3]
    if (!shape_is(shape, ndim, {C, kHeadDim, 3, 3})) return bad_shape();
    if (C % 64 == 0) {
      // pairs of groups as one dense 64->64 convolution with block-diagonal weights (zeros from the arena fill):
      // row (pair p, gl*32+col) holds, per tap, the 32 inputs of group 2p+gl at columns tap*64 + gl*32 + ci
      for (int gl = 0; gl < 2; ++gl) {
        const int dims[4] = {C / 64, kHeadDim, kHeadDim, 9};
        const long long ss[4] = {64 * 288, 288, 9, 1}, ds[4] = {64 * 576, 576, 1, 64};
        CKL(launch_repack(src + (long long)gl * kHeadDim * 288, toff(h, w.w_g, (long long)gl * (kHeadDim * 576 + kHeadDim)),
                          h->bf16(), dims, ss, ds, st));
      }
    } else if ((rc = repack(h, src, w.w_g, true, C, kHeadDim, 9, 288, 9, 1, 288, 1, kHeadDim, st))) return rc;   // [C][tap*32 + ci]
    return done();
  }
  if (rest == "conv.bias") { if (!shape_is(shape, ndim, {C})) return bad_shape(); if ((rc = copy_f(h, src, w.b_g, C, st))) return rc; return done(); }
  if (rest == "encodings.proj1.weight") { if (!shape_is(shape, ndim, {4 * C, 2 * C})) return bad_shape(); if ((rc = copy_t(h, src, toff(h, L.w1, (long long)w.lb * 8 * C * C), 8LL * C * C, st))) return rc; return done(); }
  if (rest == "encodings.proj1.bias") { if (!shape_is(shape, ndim, {4 * C})) return bad_shape(); if ((rc = copy_f(h, src, L.b1 + (long long)w.lb * 4 * C, 4 * C, st))) return rc; return done(); }
  if (rest == "encodings.proj2.weight") { if (!shape_is(shape, ndim, {2 * C, 4 * C})) return bad_shape(); if ((rc = copy_t(h, src, toff(h, L.w2, (long long)w.lb * 8 * C * C), 8LL * C * C, st))) return rc; return done(); }
  if (rest == "encodings.proj2.bias") { if (!shape_is(shape, ndim, {2 * C})) return bad_shape(); if ((rc = copy_f(h, src, L.b2 + (long long)w.lb * 2 * C, 2 * C, st))) return rc; return done(); }
  if (w.attn && rest.rfind("self_attention.attention.", 0) == 0) {
    const std::string t = rest.substr(strlen("self_attention.attention."));
    if (t == "in_proj_weight") { if (!shape_is(shape, ndim, {3 * C, C})) return bad_shape(); if ((rc = copy_t(h, src, w.w_in, 3LL * C * C, st))) return rc; return done(); }
    if (t == "in_proj_bias") { if (!shape_is(shape, ndim, {3 * C})) return bad_shape(); if ((rc = copy_f(h, src, w.b_in, 3 * C, st))) return rc; return done(); }
    if (t == "out_proj.weight") { if (!shape_is(shape, ndim, {C, C})) return bad_shape(); if ((rc = copy_t(h, src, toff(h, w.w_c, 5LL * C * C), (long long)C * C, st))) return rc; return done(); }
    if (t == "out_proj.bias") { if (!shape_is(shape, ndim, {C})) return bad_shape(); if ((rc = copy_f(h, src, w.b_c + 5LL * C, C, st))) return rc; return done(); }
  }
  return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
}

// =====================================================================================
// UNet: workspaces, position tables, forward
// =====================================================================================
namespace {

int unet_reserve(ldmb_handle* h, int B, int Hs, int Ws, int n_t, bool per_image = false) {   // Hs, Ws: post-stem resolution
  UNetState& u = h->unet;
  const int S = u.cfg.num_levels;
  const size_t ts = h->tsize();
  size_t mx_mc = 0, mx_low = 0;
  int rc;
  for (int l = 0; l < S; ++l) {
    LevelW& L = u.levels[l];
    const size_t HW = (size_t)(Hs >> l) * (Ws >> l), M = HW * B, C = L.C;
    if ((rc = ensure(h, L.xs, M * C * 4))) return rc;
    if ((rc = ensure(h, L.emb, (size_t)n_t * HW * 2 * C * ts))) return rc;
    if ((rc = ensure(h, L.h1, (size_t)n_t * HW * L.nb * 4 * C * ts))) return rc;
    if ((rc = ensure(h, L.film, (size_t)L.nb * n_t * HW * 2 * C * 4))) return rc;
    if (M * C > mx_mc) mx_mc = M * C;
    if (l < S - 1) {
      const size_t Ml = M / 4, Cn = u.cfg.channels[l + 1];
      const size_t need = Ml * (C > Cn ? C : Cn);
      if (need > mx_low) mx_low = need;
    }
  }
  if ((rc = ensure(h, u.xm, mx_mc * ts))) return rc;
  if (per_image) u.per_image_ws = true;
  if ((rc = ensure(h, u.hbuf, mx_mc * (u.per_image_ws ? 6 : 4) * ts))) return rc;   // per-image plans: all five experts + attention
  if (u.per_image_ws && (rc = ensure(h, u.backup, mx_mc * 4))) return rc;
  if ((rc = ensure(h, u.qkv, mx_mc * 3 * ts))) return rc;
  if ((rc = ensure(h, u.pooled, (mx_low ? mx_low : 64) * ts))) return rc;
  return LDMB_OK;
}

}  // namespace

extern "C" int ldmb_unet_reserve(ldmb_handle* h, int max_batch, int H, int W, int max_t) {
  if (!h) return LDMB_ERR_INVALID;
  if (!h->unet.configured) return fail(h, LDMB_ERR_STATE, "ldmb_unet_configure has not been called");
  CK(cudaSetDevice(h->device));
  const int s = h->unet.cfg.stem_size;
  if (max_batch < 1 || H < s || W < s || max_t < 1) return fail(h, LDMB_ERR_INVALID, "reserve: bad sizes");
  return unet_reserve(h, max_batch, H / s, W / s, max_t);
}

extern "C" int ldmb_unet_set_position_table(ldmb_handle* h, int level, const float* pe_host, int C, int Hl, int Wl,
                                            void* stream) {
  if (!h || !pe_host) return LDMB_ERR_INVALID;
  UNetState& u = h->unet;
  if (!u.configured) return fail(h, LDMB_ERR_STATE, "ldmb_unet_configure has not been called");
  if (level < 0 || level >= u.cfg.num_levels || C != u.levels[level].C || Hl < 1 || Wl < 1)
    return fail(h, LDMB_ERR_INVALID, "position table: bad level/shape");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LevelW& L = u.levels[level];
  const size_t n = (size_t)C * Hl * Wl;
  int rc;
  if ((rc = ensure(h, L.pe, n * 4))) return rc;
  // [C,H,W] -> [HW,C] through a library-owned scratch (grows like any workspace; no allocation / synchronisation per call).
  // pe_host is pageable host memory: cudaMemcpyAsync has staged it when it returns, so the caller may free it.
  if ((rc = ensure(h, u.pe_tmp, n * 4))) return rc;
  CK(cudaMemcpyAsync(u.pe_tmp.p, pe_host, n * 4, cudaMemcpyHostToDevice, st));
  if ((rc = repack(h, static_cast<const float*>(u.pe_tmp.p), L.pe.p, false, Hl * Wl, C, 1, 1, (long long)Hl * Wl, 0, C, 1, 0, st))) return rc;
  L.peH = Hl; L.peW = Wl;
  return LDMB_OK;
}

namespace {

int run_block_per_image(ldmb_handle* h, const BlockW& w, int block_index, int B, int Hl, int Wl, int n_t, cudaStream_t st);

int run_block(ldmb_handle* h, const BlockW& w, int block_index, int B, int Hl, int Wl, int n_t, cudaStream_t st) {
  UNetState& u = h->unet;
  if (u.plan_img_dev != nullptr) return run_block_per_image(h, w, block_index, B, Hl, Wl, n_t, st);
  LevelW& L = u.levels[w.level];
  const int C = w.C, HW = Hl * Wl, M = B * HW;
  const int* pl = u.plan_dev + 4 * block_index;     // {skip, e1, e2, -}: every kernel of the block reads it on the device
  float* x = static_cast<float*>(L.xs.p);
  const float* film = static_cast<const float*>(L.film.p) + (size_t)w.lb * n_t * HW * 2 * C;
  int rc;
  // Deep levels (an image fits one 128-row tile): ChannelNorm + FiLM + grouped conv in ONE kernel, x updated in place
  // before the block's GEMMs start (no side stream, no reduction traffic).
  const bool fused_nc = h->bf16() && !h->force_simt && !(h->skip_mask & ((1u << PK_NORM) | (1u << PK_GCONV))) && normconv_in_step() && normconv_supported(B, Hl, Wl, C);
  if (fused_nc) {
    CKLP(PK_GCONV, 2.0 * M * (double)C * 9 * kHeadDim,
         launch_normconv(h->tc, x, film, u.tindex_dev, u.xm.p, w.w_g, w.b_g, B, Hl, Wl, C, kNormEps, pl, st));
  } else {
  // ChannelNorm + FiLM (modules.py:23-25, unet.py:22)
  CKLP(PK_NORM, (double)M * C * (4 + h->tsize()),
       launch_norm_film(x, film, u.tindex_dev, u.xm.p, h->bf16(), M, C, HW, kNormEps, pl, st));
  }
  // grouped 3x3 (unet.py:30): x += conv(xm); the residual stream is only ever added to (x itself is not read),
  // so the conv is forked onto the side stream and joined at the end of the block
  // (only where every concurrent update of x is an L2 reduction: halo conv reds, TMA reduce-add GEMM epilogues)
  const int ldh = 4 * C;                                   // hbuf row: [h_general | h_e1 | h_e2 | attention]
  const bool fused_ffn = h->bf16() && !h->force_simt && mlp_fused_supported(M, C);
  // the block's GEMM that updates x beside the forked conv: out_proj behind the fused feed-forward (attention blocks only),
  // else the K-concatenated c-projection
  GemmDesc c = gd();
  if (fused_ffn) {   // x += att . W_out^T + b_out   (attention.py:82 out_proj; unet.py:44,47)
    c.A = toff(h, u.hbuf.p, 3LL * C); c.lda = ldh; c.W = toff(h, w.w_c, 5LL * C * C); c.ldw = C; c.bias = w.b_c + 5LL * C;
    c.out = x; c.ldo = C; c.M = M; c.N = C; c.K = C; c.epi = EPI_ACCUM_F32; c.plan = pl;
  } else {           // x += [h_g|h_e1|h_e2|att] . [Wc_g|Wc_e1|Wc_e2|W_out]^T + biases      (unet.py:44,47: ffn and attention in one update)
    c.A = u.hbuf.p; c.lda = ldh; c.W = w.w_c; c.ldw = C; c.bias = w.b_c; c.out = x; c.ldo = C;
    c.M = M; c.N = C; c.K = (w.attn ? 4 : 3) * C; c.epi = EPI_ACCUM_F32;
    c.sel = 2; c.sel_span = C; c.sel_stride = C; c.plan = pl;
  }
  // concurrent updaters of x must ALL be L2 reductions (halo conv reds, fused feed-forward / GEMM TMA reduce-adds): a shape
  // that would send the GEMM down a read-modify-write epilogue (CUDA-core fallback, unaligned x) is not forked
  // Deterministic mode keeps ONE concurrency: in the two-GEMM blocks the conv may run beside the a|b GEMM (which does not touch x) as long
  // as it is joined BEFORE the c-projection -- every element of x still receives its two updates in a fixed order (conv, then c).
  const bool det_fork = u.deterministic && !fused_ffn && u.fork_mode == 5 && !u.no_fork_env;
  bool fork = !fused_nc && (u.fork_conv || det_fork) && h->bf16() && !h->force_simt && !h->prof_on && C % 128 == 0 && gconv_halo_supported(B, Hl, Wl, C) &&
                    ((fused_ffn && !w.attn) || tc_accum_is_reduction(c)) && !(u.fork_mode == 3 && !fused_ffn) && !(u.fork_mode == 4 && fused_ffn) && !(fused_ffn && !((u.fork_fused >> (C == 128 ? 0 : 1)) & 1)) &&
                    !(fused_ffn && w.attn && u.attn_fused_fork == 0);
  if (fork && !u.side_stream) {
    CK(cudaStreamCreateWithFlags(&u.side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&u.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&u.ev_join, cudaEventDisableTiming));
  }
  // Late fork (two-GEMM blocks): the conv starts with the c-projection, whose grid (output tiles x split-K) leaves some SMs idle,
  // and is capped to those -- beside the a|b GEMM it would take SMs that GEMM's persistent CTAs then wait for.
  int conv_cap = 0;
  if (fork && !fused_ffn && u.fork_mode == 2) {
    conv_cap = tc_num_sms(h->tc) - tc_gemm_ctas(h->tc, c);
    if (conv_cap < C / 64) conv_cap = 0;                     // fewer idle SMs than channel slices: fork early instead
  }
  // SM partitioning (fork mode 5, two-GEMM blocks): the a|b GEMM runs on just enough CTA pairs for its number of tile rounds (one
  // round more where that would free fewer than ~1/3 of the SMs), the conv takes the SMs it leaves from the start of the block --
  // forked beside a GEMM that fills the machine, the conv's CTAs only got SMs as the GEMM's persistent CTAs retired, i.e. the
  // two ran back to back at the conv's own (latency-bound, 8 us) pace.
  int ab_cap = 0, part_conv_cap = 0;
  if (fork && !fused_ffn && u.fork_mode == 5) {
    const int sms = tc_num_sms(h->tc), pairs = sms / 2;
    const long long tiles = (long long)((M + 255) / 256) * ((6 * C + 255) / 256);
    const int rounds0 = (int)((tiles + pairs - 1) / pairs);
    // the fewest rounds (the GEMM's own count, or one more if that is at most a third longer) that leave the conv part_min_free SMs;
    // large batches have many rounds and nothing to gain: no partition, the plain early fork
    for (int rounds = rounds0; rounds <= rounds0 + 1 && 3 * rounds <= 4 * rounds0; ++rounds) {
      const int need = (int)((tiles + rounds - 1) / rounds);
      if (sms - 2 * need >= u.part_min_free && sms - 2 * need >= C / 64) { ab_cap = 2 * need; part_conv_cap = sms - 2 * need; break; }
    }
  }
  if (u.deterministic && fork && ab_cap == 0) fork = false;      // no SMs to give the conv: serial
  const bool late = conv_cap > 0;
  // fused feed-forward blocks in partitioned mode: conv part 1 beside the capped feed-forward kernel, part 2 after the join
  const int conv_split = (fork && fused_ffn && !w.attn && u.fork_mode == 5 && u.split_free >= C / 64) ? (C == 128 ? u.conv_split128 : u.conv_split256) : 0;
  auto fork_conv_now = [&](int cap) -> int {
    CK(cudaEventRecord(u.ev_fork, st));
    CK(cudaStreamWaitEvent(u.side_stream, u.ev_fork, 0));
    int rc2 = conv_split > 0 ? grouped_conv(h, u.xm.p, w.w_g, w.b_g, x, B, Hl, Wl, C, pl, u.side_stream, false, u.split_free, 1, conv_split)
                             : grouped_conv(h, u.xm.p, w.w_g, w.b_g, x, B, Hl, Wl, C, pl, u.side_stream, false, cap);
    if (rc2) return rc2;
    CK(cudaEventRecord(u.ev_join, u.side_stream));
    return LDMB_OK;
  };
  // attention blocks in partitioned mode: the in-projection GEMM and the attention core want the whole machine (a conv beside them
  // stretched the core from 13 to 30 us), so the conv is forked after them, beside the a|b GEMM that leaves it its SMs
  // (LDMB_ATTN_PART=1: instead beside the in-projection GEMM, capped to the pairs its tile rounds need -- 96 / 48 tiles take 2 / 1 rounds
  // on 48 pairs as on 74 -- and the a|b GEMM then gets the whole machine)
  int qkv_cap = 0;
  if (fork && !late && w.attn && ab_cap > 0 && u.attn_part == 1) {
    const int sms = tc_num_sms(h->tc), pairs = sms / 2;
    const long long tiles = (long long)((M + 255) / 256) * ((3 * C + 255) / 256);
    const int rounds = (int)((tiles + pairs - 1) / pairs), need = (int)((tiles + rounds - 1) / rounds);
    if (sms - 2 * need >= u.part_min_free) { qkv_cap = 2 * need; part_conv_cap = sms - 2 * need; ab_cap = 0; }
  }
  const bool fork_after_attn = fork && !late && w.attn && qkv_cap == 0 && (ab_cap > 0 || (fused_ffn && u.attn_fused_fork == 2));
  if (fork && !late && !fork_after_attn) {
    if ((rc = fork_conv_now(part_conv_cap))) return rc;
  } else if (!fork && !fused_nc && (rc = grouped_conv(h, u.xm.p, w.w_g, w.b_g, x, B, Hl, Wl, C, pl, st, false))) return rc;
  if (w.attn) {   // WindowAttention (attention.py:13-85): in_proj GEMM, per-window core; out_proj rides in the last GEMM
    GemmDesc d = gd();
    d.A = u.xm.p; d.lda = C; d.W = w.w_in; d.ldw = C; d.bias = w.b_in; d.out = u.qkv.p; d.ldo = 3 * C;
    d.M = M; d.N = 3 * C; d.K = C; d.epi = EPI_STORE; d.plan = pl;
    d.max_ctas = qkv_cap;
    if ((rc = gemm(h, d, st, PK_QKV))) return rc;
    const bool global = Hl <= kWindow && Wl <= kWindow;      // attention.py:15-16
    if ((rc = window_attention(h, u.qkv.p, u.xm.p, w.b_in, toff(h, u.hbuf.p, 3LL * C), ldh, B, Hl, Wl, C, global ? Hl : kWindow,
                               global ? Wl : kWindow, global ? 0 : w.shift, pl, st, 0))) return rc;
    if (fork_after_attn && (rc = fork_conv_now(part_conv_cap))) return rc;
  }
  // RandomMoE of ReGLU experts (modules.py:14-15,34-36): general + e1 + e2, experts resolved on the device from the plan
  if (fused_ffn) {
    // C = 128 / 256: a|b GEMM, gate and c GEMM in one kernel, h stays on the SM; in attention blocks the out_proj rides along
    const bool att_in_mlp = w.attn && mlp_fused_att_supported(M, C);
    CKLP(PK_FFN_AB, 2.0 * M * (double)C * (att_in_mlp ? 10 : 9) * C,
         launch_mlp_fused(h->tc, u.xm.p, w.w_ab, w.b_ab, w.w_c, w.b_c, x, M, C, (w.attn ? 6 : 5) * C, pl, 0, 0, nullptr, 0, st,
                          conv_split > 0 ? tc_num_sms(h->tc) - u.split_free : 0, att_in_mlp ? toff(h, u.hbuf.p, 3LL * C) : nullptr, ldh));
    if (w.attn && !att_in_mlp && (rc = gemm(h, c, st, PK_FFN_C))) return rc;
    if (fork) CK(cudaStreamWaitEvent(st, u.ev_join, 0));
    if (conv_split > 0 && (rc = grouped_conv(h, u.xm.p, w.w_g, w.b_g, x, B, Hl, Wl, C, pl, st, false, 0, 2, conv_split))) return rc;
    return LDMB_OK;
  }
  {
    GemmDesc d = gd();
    d.A = u.xm.p; d.lda = C; d.W = w.w_ab; d.ldw = C; d.bias = w.b_ab; d.out = u.hbuf.p; d.ldo = ldh;
    d.M = M; d.N = 6 * C; d.K = C; d.epi = EPI_REGLU; d.glu_chunk = glu_chunk_for(C);
    d.sel = 1; d.sel_span = 2 * C; d.sel_stride = 2 * C; d.plan = pl;
    d.max_ctas = ab_cap;
    if ((rc = gemm(h, d, st, PK_FFN_AB))) return rc;
    if (late && (rc = fork_conv_now(conv_cap))) return rc;
    if (fork && u.deterministic) CK(cudaStreamWaitEvent(st, u.ev_join, 0));     // the conv's update of x is complete before the c-projection's
    if ((rc = gemm(h, c, st, PK_FFN_C))) return rc;
  }
  if (fork && !u.deterministic) CK(cudaStreamWaitEvent(st, u.ev_join, 0));
  return LDMB_OK;
}

// One SwinBlock with PER-IMAGE stochastic-depth / expert decisions (the reference's batch-1 loops, sample_ldm.py:71-72,
// draw them per image).  Exact.  Where an image is a whole number of 128-row tiles (C = 128 / 256 levels) the fused
// feed-forward kernel resolves the experts per tile and does nothing for tiles of skipped images: same work as one
// shared plan.  Elsewhere (deep levels, 16..64 pixels per image) dense kernels: all five ReGLU experts are evaluated,
// the two an image did not draw are written as zeros by the gate epilogue, the c-projection runs over all slots and
// its biases are added per image.  Rows of images that skip the block are saved on entry and put back on exit.
int run_block_per_image(ldmb_handle* h, const BlockW& w, int block_index, int B, int Hl, int Wl, int n_t, cudaStream_t st) {
  UNetState& u = h->unet;
  LevelW& L = u.levels[w.level];
  const int C = w.C, HW = Hl * Wl, M = B * HW;
  const int* pimg = u.plan_img_dev + (size_t)block_index * B;
  float* x = static_cast<float*>(L.xs.p);
  const float* film = static_cast<const float*>(L.film.p) + (size_t)w.lb * n_t * HW * 2 * C;
  const bool fused = h->bf16() && !h->force_simt && mlp_fused_per_image_supported(M, C, HW);
  int rc;
  CKLP(PK_NORM, (double)M * C * (4 + h->tsize()),
       launch_norm_film(x, film, u.tindex_dev, u.xm.p, h->bf16(), M, C, HW, kNormEps, nullptr, st));
  CKL(launch_rows_enter(x, static_cast<float*>(u.backup.p), fused ? nullptr : w.b_c, pimg, M, HW, C, w.attn, st));
  // from here on x is only added to (L2 reductions): the grouped conv runs on the side stream, as in run_block
  const bool fork = u.fork_conv && h->bf16() && !h->force_simt && !h->prof_on && C % 128 == 0 && gconv_halo_supported(B, Hl, Wl, C);
  if (fork) {
    if (!u.side_stream) {
      CK(cudaStreamCreateWithFlags(&u.side_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&u.ev_fork, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&u.ev_join, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(u.ev_fork, st));
    CK(cudaStreamWaitEvent(u.side_stream, u.ev_fork, 0));
    if ((rc = grouped_conv(h, u.xm.p, w.w_g, w.b_g, x, B, Hl, Wl, C, nullptr, u.side_stream, false))) return rc;
    CK(cudaEventRecord(u.ev_join, u.side_stream));
  } else if ((rc = grouped_conv(h, u.xm.p, w.w_g, w.b_g, x, B, Hl, Wl, C, nullptr, st, false))) return rc;
  const int ldh = 6 * C;                                   // hbuf row: [h_general | h_e0 .. h_e3 | attention]
  if (w.attn) {
    GemmDesc d = gd();
    d.A = u.xm.p; d.lda = C; d.W = w.w_in; d.ldw = C; d.bias = w.b_in; d.out = u.qkv.p; d.ldo = 3 * C;
    d.M = M; d.N = 3 * C; d.K = C; d.epi = EPI_STORE;
    if ((rc = gemm(h, d, st, PK_QKV))) return rc;
    const bool global = Hl <= kWindow && Wl <= kWindow;
    if ((rc = window_attention(h, u.qkv.p, u.xm.p, w.b_in, toff(h, u.hbuf.p, 5LL * C), ldh, B, Hl, Wl, C, global ? Hl : kWindow,
                               global ? Wl : kWindow, global ? 0 : w.shift, nullptr, st, 0))) return rc;
  }
  if (fused) {
    CKLP(PK_FFN_AB, 2.0 * M * (double)C * 9 * C,
         launch_mlp_fused(h->tc, u.xm.p, w.w_ab, w.b_ab, w.w_c, w.b_c, x, M, C, (w.attn ? 6 : 5) * C, nullptr, 0, 0, pimg, HW, st));
    if (w.attn) {   // x += att . W_out^T + b_out for every image (skipped ones are restored below)
      GemmDesc c = gd();
      c.A = toff(h, u.hbuf.p, 5LL * C); c.lda = ldh; c.W = toff(h, w.w_c, 5LL * C * C); c.ldw = C; c.bias = w.b_c + 5LL * C;
      c.out = x; c.ldo = C; c.M = M; c.N = C; c.K = C; c.epi = EPI_ACCUM_F32;
      if ((rc = gemm(h, c, st, PK_FFN_C))) return rc;
    }
  } else {
    GemmDesc d = gd();                                     // all five experts: [M,C] . [C, 10C], ReGLU gate (masked per image) -> h [M, 5C]
    d.A = u.xm.p; d.lda = C; d.W = w.w_ab; d.ldw = C; d.bias = w.b_ab; d.out = u.hbuf.p; d.ldo = ldh;
    d.M = M; d.N = 10 * C; d.K = C; d.epi = EPI_REGLU; d.glu_chunk = glu_chunk_for(C);
    d.mask_plan = pimg; d.mask_rows = HW; d.mask_span = C;
    if ((rc = gemm(h, d, st, PK_FFN_AB))) return rc;
    GemmDesc c = gd();                                     // x += [h | att] . [Wc_g | Wc_0..3 | W_out]^T (biases: rows_enter)
    c.A = u.hbuf.p; c.lda = ldh; c.W = w.w_c; c.ldw = C; c.bias = nullptr; c.out = x; c.ldo = C;
    c.M = M; c.N = C; c.K = (w.attn ? 6 : 5) * C; c.epi = EPI_ACCUM_F32;
    c.sel = 3; c.sel_span = C; c.sel_stride = C;
    if ((rc = gemm(h, c, st, PK_FFN_C))) return rc;
  }
  if (fork) CK(cudaStreamWaitEvent(st, u.ev_join, 0));
  CKL(launch_rows_leave(x, static_cast<const float*>(u.backup.p), pimg, M, HW, C, st));
  return LDMB_OK;
}

// Every launch of one UNet step, in order, on `st`.  Static given (B, Hs, Ws, n_t): all per-step values are read from
// the device-side step buffer, so the sequence can be captured once into a CUDA graph and replayed.
// Encodings (unet.py:18-21) of one level, hoisted: the FiLM rows depend on (t,h,w) only -> one evaluation for all
// blocks of the level and all n_t timesteps: film[block][t][pixel][mul | bias].
int issue_encodings(ldmb_handle* h, int l, int HW, int n_t, const float* te_dev, cudaStream_t st) {
  LevelW& L = h->unet.levels[l];
  const int C = L.C, Mt = n_t * HW;
  int rc;
  CKL(launch_emb_build(static_cast<const float*>(L.pe.p), te_dev, L.emb.p, h->bf16(), n_t, HW, C, st));
  GemmDesc a = gd();
  a.A = L.emb.p; a.lda = 2 * C; a.W = L.w1; a.ldw = 2 * C; a.bias = L.b1; a.out = L.h1.p; a.ldo = (long long)L.nb * 4 * C;
  a.M = Mt; a.N = L.nb * 4 * C; a.K = 2 * C; a.epi = EPI_STORE; a.act = ACT_RELU;
  if ((rc = gemm(h, a, st, PK_ENC))) return rc;
  GemmDesc b = gd();
  b.A = L.h1.p; b.lda = (long long)L.nb * 4 * C; b.W = L.w2; b.ldw = 4 * C; b.bias = L.b2; b.out = L.film.p; b.ldo = 2 * C;
  b.M = Mt; b.N = 2 * C; b.K = 4 * C; b.epi = EPI_STORE_F32;
  b.batch = L.nb; b.a_koff_b = 4 * C; b.w_row_b = 2 * C; b.out_off_b = (long long)Mt * 2 * C; b.bias_off_b = 2 * C;
  return gemm(h, b, st, PK_ENC);
}

int issue_forward(ldmb_handle* h, int B, int Hs, int Ws, int n_t, const float* const* te_dev, cudaStream_t st, bool pre) {
  UNetState& u = h->unet;
  const ldmb_unet_config& cfg = u.cfg;
  const int S = cfg.num_levels, s = cfg.stem_size;
  int rc;
  if (!pre)
    for (int l = 0; l < S; ++l)
      if ((rc = issue_encodings(h, l, (Hs >> l) * (Ws >> l), n_t, te_dev[l], st))) return rc;
  // ---- encoder_first (unet.py:90)
  CKLP(PK_EDGE, 0, launch_stem(u.sp_dev, u.w_first, u.b_first, static_cast<float*>(u.levels[0].xs.p), B, cfg.input_channels, Hs, Ws, s,
                  cfg.channels[0], h->bf16() && !h->force_simt, st));
  // ---- encoder (unet.py:92-98)
  int bi = 0;
  for (int l = 0; l < S; ++l) {
    const int Hl = Hs >> l, Wl = Ws >> l;
    for (int b = 0; b < cfg.blocks[l]; ++b, ++bi)
      if ((rc = run_block(h, u.blocks[bi], bi, B, Hl, Wl, n_t, st))) return rc;
    if (l < S - 1) {
      // ch_conv = Conv1x1 then AvgPool2 (unet.py:83); the two commute, pool first = 4x fewer FLOPs
      LevelW& L = u.levels[l];
      const int C = L.C, Cn = cfg.channels[l + 1];
      CKL(launch_pool_cast(static_cast<const float*>(L.xs.p), u.pooled.p, h->bf16(), B, Hl, Wl, C, st));
      GemmDesc d = gd();
      d.A = u.pooled.p; d.lda = C; d.W = L.w_down; d.ldw = C; d.bias = L.b_down; d.out = u.levels[l + 1].xs.p; d.ldo = Cn;
      d.M = B * (Hl / 2) * (Wl / 2); d.N = Cn; d.K = C; d.epi = EPI_STORE_F32;
      if ((rc = gemm(h, d, st, PK_LEVEL))) return rc;
    }
  }
  // ---- decoder (unet.py:99-101); the deepest level continues in place (skip = 0)
  for (int l = S - 1; l >= 0; --l) {
    const int Hl = Hs >> l, Wl = Ws >> l;
    if (l < S - 1) {
      // ch_conv = Upsample(2, nearest) then Conv1x1 (unet.py:85): conv at low resolution, replicated into the skip-connected
      // residual stream by the GEMM's epilogue (EPI_UPADD)
      LevelW& L = u.levels[l];
      const int C = L.C, Cn = cfg.channels[l + 1];
      const long long Mlow = (long long)B * (Hl / 2) * (Wl / 2);
      CKL(launch_cast(static_cast<const float*>(u.levels[l + 1].xs.p), u.pooled.p, h->bf16(), Mlow * Cn, st));
      GemmDesc d = gd();
      d.A = u.pooled.p; d.lda = Cn; d.W = L.w_up; d.ldw = Cn; d.bias = L.b_up; d.out = L.xs.p; d.ldo = C;
      d.M = (int)Mlow; d.N = C; d.K = Cn; d.epi = EPI_UPADD; d.ctH = Hl / 2; d.ctW = Wl / 2;   // x_l += up2(conv(x_{l+1})) in the epilogue
      if ((rc = gemm(h, d, st, PK_LEVEL))) return rc;
    }
    for (int b = 0; b < cfg.blocks[l]; ++b, ++bi)
      if ((rc = run_block(h, u.blocks[bi], bi, B, Hl, Wl, n_t, st))) return rc;
  }
  // ---- decoder_last (unet.py:102) + DDIM update (ddpm.py:81-91)
  CKLP(PK_EDGE, 0, launch_final(static_cast<const float*>(u.levels[0].xs.p), u.w_last, u.b_last, u.sp_dev, B, cfg.input_channels, Hs, Ws, s,
                   cfg.channels[0], h->bf16() && !h->force_simt, st));
  return LDMB_OK;
}

}  // namespace

extern "C" int ldmb_host_draw_plans(const uint32_t* raw, int64_t n_raw, int n_plans, int n_blocks, const uint8_t* training,
                                    const double* depth_p, int n_experts, int32_t* plan_out, int64_t* used) {
  if (!raw || !training || !depth_p || !plan_out || !used || n_plans < 0 || n_blocks <= 0 || n_experts < 2 || n_experts > 21)
    return LDMB_ERR_INVALID;                      // random.sample switches to its set-based path above 21 items
  int bits[2];
  for (int i = 0; i < 2; ++i) {
    int n = n_experts - i, b = 0;
    while (n >> b) ++b;                            // int.bit_length()
    bits[i] = b;
  }
  int64_t pos = 0;
  for (int p = 0; p < n_plans; ++p)
    for (int k = 0; k < n_blocks; ++k) {
      int32_t* o = plan_out + ((size_t)p * n_blocks + k) * 3;
      o[0] = o[1] = o[2] = 0;
      if (training[k]) {
        if (pos + 2 > n_raw) return LDMB_ERR_INVALID;
        const double a = (double)(raw[pos] >> 5), b = (double)(raw[pos + 1] >> 6);
        pos += 2;
        if ((a * 67108864.0 + b) * (1.0 / 9007199254740992.0) <= depth_p[k]) { o[0] = 1; continue; }
      }
      int j[2];
      for (int i = 0; i < 2; ++i) {
        uint32_t r;
        do {
          if (pos >= n_raw) return LDMB_ERR_INVALID;
          r = raw[pos++] >> (32 - bits[i]);
        } while (r >= (uint32_t)(n_experts - i));
        j[i] = (int)r;
      }
      // pool = [0..n); first = pool[j0]; pool[j0] = pool[n-1]; second = pool[j1]
      o[1] = j[0];
      o[2] = j[1] == j[0] ? n_experts - 1 : j[1];
    }
  *used = pos;
  return LDMB_OK;
}

extern "C" int ldmb_set_deterministic(ldmb_handle* h, int on) {
  if (!h) return LDMB_ERR_INVALID;
  tc_set_splitk(h->tc, on == 0);
  h->unet.fork_conv = on == 0 && !h->unet.no_fork_env;
  h->unet.deterministic = on != 0;
  h->unet.ws_epoch++;          // captured graphs hold the old launch set
  return LDMB_OK;
}

extern "C" int ldmb_set_use_graphs(ldmb_handle* h, int on) {
  if (!h) return LDMB_ERR_INVALID;
  h->unet.use_graphs = on != 0;
  return LDMB_OK;
}

static int unet_forward_impl(ldmb_handle* h, const float* x_dev, float* out_dev, int B, int H, int W,
                             const int32_t* t_index, int n_t, const float* const* te_host, const int32_t* plan,
                             const int32_t* plan_img, const ldmb_ddim_coef* coef, const float* noise_dev, void* stream) {
  if (!h || !x_dev || !out_dev || !t_index || (!plan && !plan_img)) return LDMB_ERR_INVALID;
  CHECK_FAULT();
  const bool per_image = plan_img != nullptr;
  const bool pre = te_host == nullptr;      // FiLM tables of the n_t timesteps were precomputed (ldmb_unet_precompute_film)
  UNetState& u = h->unet;
  if (!u.configured) return fail(h, LDMB_ERR_STATE, "ldmb_unet_configure has not been called");
  if (!u.missing.empty())
    return fail(h, LDMB_ERR_STATE, "%d UNet parameters not loaded (first: %s)", (int)u.missing.size(), u.missing.begin()->c_str());
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ldmb_unet_config& cfg = u.cfg;
  const int S = cfg.num_levels, s = cfg.stem_size;
  const int nblk = (int)u.blocks.size();
  if (B < 1 || n_t < 1 || H < s || W < s || H % s || W % s) return fail(h, LDMB_ERR_INVALID, "forward: bad batch/resolution");
  const int Hs = H / s, Ws = W / s;
  if ((Hs % (1 << (S - 1))) || (Ws % (1 << (S - 1))))
    return fail(h, LDMB_ERR_INVALID, "resolution %dx%d is not divisible by 2^%d: skip shapes would not match (unet.py:101)", Hs, Ws, S - 1);
  for (int b = 0; b < B; ++b) if (t_index[b] < 0 || t_index[b] >= n_t) return fail(h, LDMB_ERR_INVALID, "t_index out of range");
  if (pre && (u.film_nt != n_t || u.film_Hs != Hs || u.film_Ws != Ws))
    return fail(h, LDMB_ERR_STATE, "te_host == NULL needs ldmb_unet_precompute_film for %d timesteps at this resolution", n_t);
  if (!pre) u.film_nt = 0;                   // this call overwrites the FiLM workspace
  for (int b = 0; b < nblk && !per_image; ++b)
    if (!plan[3 * b] && (plan[3 * b + 1] < 0 || plan[3 * b + 1] >= kExperts || plan[3 * b + 2] < 0 || plan[3 * b + 2] >= kExperts))
      return fail(h, LDMB_ERR_INVALID, "plan: expert index out of range");
  for (long long i = 0; per_image && i < (long long)nblk * B; ++i)
    if (!plan_img[3 * i] && (plan_img[3 * i + 1] < 0 || plan_img[3 * i + 1] >= kExperts || plan_img[3 * i + 2] < 0 ||
                             plan_img[3 * i + 2] >= kExperts || plan_img[3 * i + 1] == plan_img[3 * i + 2]))
      return fail(h, LDMB_ERR_INVALID, "per-image plan: expert indices out of range or equal");
  if (coef && coef->sigma != 0.f && !noise_dev) return fail(h, LDMB_ERR_INVALID, "sigma != 0 needs a noise tensor");
  int rc;
  if ((rc = unet_reserve(h, B, Hs, Ws, n_t, per_image))) return rc;
  for (int l = 0; l < S; ++l)
    if (u.levels[l].peH != (Hs >> l) || u.levels[l].peW != (Ws >> l) || !u.levels[l].pe.p)
      return fail(h, LDMB_ERR_STATE, "position table of level %d not set for %dx%d", l, Hs >> l, Ws >> l);

  // ---- step buffer layout: StepParams | plan[nblk][4] | t_index[B] | te tables (16-byte aligned pieces)
  auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
  const size_t off_plan = al(sizeof(StepParams));
  const size_t off_tidx = off_plan + al((size_t)nblk * 16);
  size_t off_te[LDMB_MAX_LEVELS];
  const size_t off_pimg = off_tidx + al((size_t)B * 4);
  size_t total = off_pimg + (per_image ? al((size_t)nblk * B * 4) : 0);
  for (int l = 0; l < S; ++l) { off_te[l] = total; total += pre ? 0 : al((size_t)n_t * u.levels[l].C * 4); }
  if ((rc = ensure(h, u.stepbuf, total))) return rc;
  char* sb = static_cast<char*>(u.stepbuf.p);
  u.sp_dev = reinterpret_cast<const StepParams*>(sb);
  u.plan_dev = reinterpret_cast<const int*>(sb + off_plan);
  u.tindex_dev = reinterpret_cast<const int*>(sb + off_tidx);
  u.plan_img_dev = per_image ? reinterpret_cast<const int*>(sb + off_pimg) : nullptr;
  const float* te_dev[LDMB_MAX_LEVELS];
  for (int l = 0; l < S; ++l) te_dev[l] = reinterpret_cast<const float*>(sb + off_te[l]);

  // ---- fill one slot of the pinned staging ring and upload it with a single stream-ordered copy
  if (total > u.staging_slot_bytes) {
    CK(cudaDeviceSynchronize());
    if (u.staging) CK(cudaFreeHost(u.staging));
    u.staging_slot_bytes = total * 2 + 4096;
    CK(cudaMallocHost((void**)&u.staging, u.staging_slot_bytes * kStagingSlots));
    for (int i = 0; i < kStagingSlots; ++i) {
      if (!u.staging_ev[i]) CK(cudaEventCreateWithFlags(&u.staging_ev[i], cudaEventDisableTiming));
      u.staging_used[i] = false;
    }
  }
  const int slot = u.staging_next;
  u.staging_next = (slot + 1) % kStagingSlots;
  if (u.staging_used[slot]) CK(cudaEventSynchronize(u.staging_ev[slot]));
  char* hp = u.staging + (size_t)slot * u.staging_slot_bytes;
  StepParams spv;
  memset(&spv, 0, sizeof(spv));
  spv.x_in = x_dev; spv.out = out_dev; spv.noise = noise_dev;
  if (coef) {
    spv.c_eps_in = coef->c_eps_in; spv.c_div = coef->c_div; spv.c_x0 = coef->c_x0; spv.c_eps_out = coef->c_eps_out;
    spv.sigma = coef->sigma; spv.final_step = coef->final_step; spv.ddim_enabled = 1;
  }
  memcpy(hp, &spv, sizeof(spv));
  int* hplan = reinterpret_cast<int*>(hp + off_plan);
  for (int b = 0; b < nblk; ++b) {
    hplan[4 * b] = per_image ? 0 : plan[3 * b]; hplan[4 * b + 1] = per_image ? 0 : plan[3 * b + 1];
    hplan[4 * b + 2] = per_image ? 1 : plan[3 * b + 2]; hplan[4 * b + 3] = 0;
  }
  if (per_image) {
    int* hp_img = reinterpret_cast<int*>(hp + off_pimg);
    for (long long i = 0; i < (long long)nblk * B; ++i)
      hp_img[i] = (plan_img[3 * i] ? 1 : 0) | (plan_img[3 * i + 1] << 8) | (plan_img[3 * i + 2] << 16);
  }
  memcpy(hp + off_tidx, t_index, (size_t)B * 4);
  for (int l = 0; l < S && !pre; ++l) {
    if (!te_host[l]) return fail(h, LDMB_ERR_INVALID, "te_host[%d] is NULL", l);
    memcpy(hp + off_te[l], te_host[l], (size_t)n_t * u.levels[l].C * 4);
  }
  CK(cudaMemcpyAsync(sb, hp, total, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(u.staging_ev[slot], st));
  u.staging_used[slot] = true;

  // ---- launch: eagerly the first time a shape is seen (and when profiling), then capture once and replay the graph
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CK(cudaStreamIsCapturing(st, &cap));
  UNetState::GraphEntry* ge = nullptr;
  if (u.use_graphs && !h->prof_on && cap == cudaStreamCaptureStatusNone) {
    for (auto& g : u.graphs)
      if (g.B == B && g.Hs == Hs && g.Ws == Ws && g.n_t == n_t && g.pre == (int)pre && g.per_image == (int)per_image) { ge = &g; break; }
    if (!ge) {
      if (u.graphs.size() >= 16) { for (auto& g : u.graphs) if (g.exec) cudaGraphExecDestroy(g.exec); u.graphs.clear(); }
      u.graphs.push_back({B, Hs, Ws, n_t, (int)pre, (int)per_image, u.ws_epoch, nullptr, 0, 0});
      ge = &u.graphs.back();
    }
    if (ge->epoch != u.ws_epoch) {     // a workspace moved since capture: the graph holds stale pointers
      if (ge->exec) { cudaGraphExecDestroy(ge->exec); ge->exec = nullptr; }
      ge->epoch = u.ws_epoch; ge->seen = 0;
    }
  }
  if (ge && ge->exec) {
    CK(cudaGraphLaunch(ge->exec, st));
    h->launches += ge->launches;
    return LDMB_OK;
  }
  if (ge && ge->seen >= 1) {
    const long long l0 = h->launches;
    if (!u.cap_stream) CK(cudaStreamCreateWithFlags(&u.cap_stream, cudaStreamNonBlocking));
    CK(cudaStreamBeginCapture(u.cap_stream, cudaStreamCaptureModeThreadLocal));
    rc = issue_forward(h, B, Hs, Ws, n_t, te_dev, u.cap_stream, pre);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(u.cap_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || !graph) return fail(h, LDMB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ge->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { ge->exec = nullptr; return fail(h, LDMB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
    ge->launches = h->launches - l0;
    CK(cudaGraphLaunch(ge->exec, st));
    return LDMB_OK;
  }
  if (ge) ge->seen++;
  return issue_forward(h, B, Hs, Ws, n_t, te_dev, st, pre);
}

extern "C" int ldmb_unet_forward(ldmb_handle* h, const float* x_dev, float* out_dev, int B, int H, int W,
                                 const int32_t* t_index, int n_t, const float* const* te_host, const int32_t* plan,
                                 const ldmb_ddim_coef* coef, const float* noise_dev, void* stream) {
  if (!plan) return LDMB_ERR_INVALID;
  return unet_forward_impl(h, x_dev, out_dev, B, H, W, t_index, n_t, te_host, plan, nullptr, coef, noise_dev, stream);
}

extern "C" int ldmb_unet_forward_per_image(ldmb_handle* h, const float* x_dev, float* out_dev, int B, int H, int W,
                                           const int32_t* t_index, int n_t, const float* const* te_host,
                                           const int32_t* plan_img, const ldmb_ddim_coef* coef, const float* noise_dev,
                                           void* stream) {
  if (!plan_img) return LDMB_ERR_INVALID;
  return unet_forward_impl(h, x_dev, out_dev, B, H, W, t_index, n_t, te_host, nullptr, plan_img, coef, noise_dev, stream);
}

extern "C" int ldmb_unet_precompute_film(ldmb_handle* h, int H, int W, int n_t, const float* const* te_host, void* stream) {
  if (!h || !te_host || n_t < 1) return LDMB_ERR_INVALID;
  CHECK_FAULT();
  UNetState& u = h->unet;
  if (!u.configured) return fail(h, LDMB_ERR_STATE, "ldmb_unet_configure has not been called");
  if (!u.missing.empty())
    return fail(h, LDMB_ERR_STATE, "%d UNet parameters not loaded (first: %s)", (int)u.missing.size(), u.missing.begin()->c_str());
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = u.cfg.num_levels, s = u.cfg.stem_size;
  if (H < s || W < s || H % s || W % s) return fail(h, LDMB_ERR_INVALID, "precompute: bad resolution");
  const int Hs = H / s, Ws = W / s;
  if ((Hs % (1 << (S - 1))) || (Ws % (1 << (S - 1)))) return fail(h, LDMB_ERR_INVALID, "resolution %dx%d is not divisible by 2^%d", Hs, Ws, S - 1);
  for (int l = 0; l < S; ++l)
    if (u.levels[l].peH != (Hs >> l) || u.levels[l].peW != (Ws >> l) || !u.levels[l].pe.p)
      return fail(h, LDMB_ERR_STATE, "position table of level %d not set for %dx%d", l, Hs >> l, Ws >> l);
  int rc;
  u.film_nt = 0;
  if ((rc = unet_reserve(h, 1, Hs, Ws, n_t))) return rc;
  size_t off[LDMB_MAX_LEVELS], total = 0;
  for (int l = 0; l < S; ++l) { off[l] = total; total += ((size_t)n_t * u.levels[l].C * 4 + 255) & ~size_t(255); }
  if ((rc = ensure(h, u.te_pre, total))) return rc;
  for (int l = 0; l < S; ++l) {
    if (!te_host[l]) return fail(h, LDMB_ERR_INVALID, "te_host[%d] is NULL", l);
    CK(cudaMemcpyAsync(static_cast<char*>(u.te_pre.p) + off[l], te_host[l], (size_t)n_t * u.levels[l].C * 4, cudaMemcpyHostToDevice, st));
  }
  CK(cudaStreamSynchronize(st));            // the host tables may go away when we return (one-off per schedule)
  for (int l = 0; l < S; ++l)
    if ((rc = issue_encodings(h, l, (Hs >> l) * (Ws >> l), n_t, reinterpret_cast<const float*>(static_cast<char*>(u.te_pre.p) + off[l]), st)))
      return rc;
  u.film_nt = n_t; u.film_Hs = Hs; u.film_Ws = Ws;
  return LDMB_OK;
}

// =====================================================================================
// VAE
// =====================================================================================
extern "C" int ldmb_vae_configure(ldmb_handle* h, int which, const ldmb_vae_config* cfg) {
  if (!h || !cfg || (which != LDMB_VAE_DECODER && which != LDMB_VAE_ENCODER)) return LDMB_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  const int S = cfg->num_levels;
  if (S < 1 || S > LDMB_MAX_LEVELS || cfg->image_channels < 1 || cfg->image_channels > 32 || cfg->latent_channels < 1 ||
      cfg->latent_channels > 32)
    return fail(h, LDMB_ERR_INVALID, "vae config: bad level count / channels");
  for (int l = 0; l < S; ++l)
    if (cfg->channels[l] < 1 || cfg->channels[l] % 8 || cfg->blocks[l] < 0)
      return fail(h, LDMB_ERR_INVALID, "vae config: channels[%d]=%d must be a positive multiple of 8", l, cfg->channels[l]);
  VaeState& v = h->vae[which];
  if (v.configured) { CK(cudaDeviceSynchronize()); v.arena.release(); }
  v.cfg = *cfg;
  v.levels.clear(); v.levels.resize(S); v.missing.clear();
  const size_t ts = h->tsize();
  Arena& a = v.arena;
  const bool dec = which == LDMB_VAE_DECODER;
  const int cin = dec ? cfg->latent_channels : cfg->image_channels;     // input_layer
  const int cout = dec ? cfg->image_channels : cfg->latent_channels;    // to_rgb / output_layer
  a.add((void**)&v.w_in, (size_t)cfg->channels[0] * cin * 4); a.add((void**)&v.b_in, (size_t)cfg->channels[0] * 4);
  v.missing.insert("input_layer.weight"); v.missing.insert("input_layer.bias");
  if (!dec) {
    a.add((void**)&v.w_out, (size_t)cout * cfg->channels[S - 1] * 4); a.add((void**)&v.b_out, (size_t)cout * 4);
    v.missing.insert("output_layer.weight"); v.missing.insert("output_layer.bias");
  }
  char nm[96];
  for (int l = 0; l < S; ++l) {
    VaeLevelW& L = v.levels[l];
    const size_t C = cfg->channels[l];
    L.C = (int)C;
    L.res.resize(cfg->blocks[l]);
    for (int r = 0; r < cfg->blocks[l]; ++r) {
      a.add(&L.res[r].w1, C * 9 * C * ts); a.add((void**)&L.res[r].b1, C * 4);
      a.add(&L.res[r].w2, C * 9 * C * ts); a.add((void**)&L.res[r].b2, C * 4);
      for (const char* c : {"c1", "c2"})
        for (const char* sfx : {"weight", "bias"}) {
          snprintf(nm, sizeof(nm), dec ? "stages.%d.layers.%d.%s.%s" : "stages.%d.seq.%d.%s.%s", l, r, c, sfx);
          v.missing.insert(nm);
        }
    }
    if (dec) {
      a.add((void**)&L.w_rgb, (size_t)cout * C * 4); a.add((void**)&L.b_rgb, (size_t)cout * 4);
      snprintf(nm, sizeof(nm), "stages.%d.to_rgb.weight", l); v.missing.insert(nm);
      snprintf(nm, sizeof(nm), "stages.%d.to_rgb.bias", l); v.missing.insert(nm);
      if (l > 0) {   // ConvTranspose2d(channels[l-1] -> channels[l], 2, 2): packed [4*C][Cprev], bias replicated x4
        const size_t Cp = cfg->channels[l - 1];
        a.add(&L.w_resample, 4 * C * Cp * ts); a.add((void**)&L.b_resample, 4 * C * 4);
        snprintf(nm, sizeof(nm), "upsamples.%d.weight", l); v.missing.insert(nm);
        snprintf(nm, sizeof(nm), "upsamples.%d.bias", l); v.missing.insert(nm);
      }
    } else if (l < S - 1) {   // AvgPool2 + Conv1x1(channels[l] -> channels[l+1])
      const size_t Cn = cfg->channels[l + 1];
      a.add(&L.w_resample, Cn * C * ts); a.add((void**)&L.b_resample, Cn * 4);
      snprintf(nm, sizeof(nm), "downsamples.%d.1.weight", l); v.missing.insert(nm);
      snprintf(nm, sizeof(nm), "downsamples.%d.1.bias", l); v.missing.insert(nm);
    }
  }
  CK(a.commit());
  v.configured = true;
  return LDMB_OK;
}

extern "C" int ldmb_vae_params_missing(const ldmb_handle* h, int which) {
  if (!h || which < 0 || which > 1 || !h->vae[which].configured) return -1;
  return (int)h->vae[which].missing.size();
}

extern "C" int ldmb_vae_load_param(ldmb_handle* h, int which, const char* name, const float* src, const int64_t* shape,
                                   int ndim, void* stream) {
  if (!h || !name || !src || !shape || which < 0 || which > 1) return LDMB_ERR_INVALID;
  VaeState& v = h->vae[which];
  if (!v.configured) return fail(h, LDMB_ERR_STATE, "ldmb_vae_configure has not been called");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool dec = which == LDMB_VAE_DECODER;
  const ldmb_vae_config& cfg = v.cfg;
  const int S = cfg.num_levels;
  const int cin = dec ? cfg.latent_channels : cfg.image_channels;
  const int cout = dec ? cfg.image_channels : cfg.latent_channels;
  std::string nm(name);
  auto bad_shape = [&]() { return fail(h, LDMB_ERR_INVALID, "size mismatch for %s", name); };
  auto done = [&]() { v.missing.erase(nm); return (int)LDMB_OK; };
  int rc, l = -1, r = -1, c = 0;
  char kind[16] = {0};
  if (nm == "input_layer.weight") { if (!shape_is(shape, ndim, {cfg.channels[0], cin})) return bad_shape(); if ((rc = copy_f(h, src, v.w_in, (long long)cfg.channels[0] * cin, st))) return rc; return done(); }
  if (nm == "input_layer.bias") { if (!shape_is(shape, ndim, {cfg.channels[0]})) return bad_shape(); if ((rc = copy_f(h, src, v.b_in, cfg.channels[0], st))) return rc; return done(); }
  if (nm == "output_layer.weight" || nm == "output_layer.bias") {
    if (dec) return LDMB_OK;   // defined but never applied by Decoder.forward (vae.py:113,122-132)
    const int C = cfg.channels[S - 1];
    if (nm == "output_layer.weight") { if (!shape_is(shape, ndim, {cout, C})) return bad_shape(); if ((rc = copy_f(h, src, v.w_out, (long long)cout * C, st))) return rc; return done(); }
    if (!shape_is(shape, ndim, {cout})) return bad_shape();
    if ((rc = copy_f(h, src, v.b_out, cout, st))) return rc;
    return done();
  }
  if (sscanf(nm.c_str(), dec ? "stages.%d.layers.%d.c%d.%15s" : "stages.%d.seq.%d.c%d.%15s", &l, &r, &c, kind) == 4) {
    if (l < 0 || l >= S || r < 0 || r >= cfg.blocks[l] || (c != 1 && c != 2)) return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
    const int C = cfg.channels[l];
    ResW& R = v.levels[l].res[r];
    if (!strcmp(kind, "weight")) {   // [C, C, 3, 3] -> [co][tap*C + ci]
      if (!shape_is(shape, ndim, {C, C, 3, 3})) return bad_shape();
      if ((rc = repack(h, src, c == 1 ? R.w1 : R.w2, true, C, C, 9, 9LL * C, 9, 1, 9LL * C, 1, C, st))) return rc;
      return done();
    }
    if (!strcmp(kind, "bias")) { if (!shape_is(shape, ndim, {C})) return bad_shape(); if ((rc = copy_f(h, src, c == 1 ? R.b1 : R.b2, C, st))) return rc; return done(); }
  }
  if (dec && sscanf(nm.c_str(), "stages.%d.to_rgb.%15s", &l, kind) == 2 && l >= 0 && l < S) {
    const int C = cfg.channels[l];
    if (!strcmp(kind, "weight")) { if (!shape_is(shape, ndim, {cout, C})) return bad_shape(); if ((rc = copy_f(h, src, v.levels[l].w_rgb, (long long)cout * C, st))) return rc; return done(); }
    if (!strcmp(kind, "bias")) { if (!shape_is(shape, ndim, {cout})) return bad_shape(); if ((rc = copy_f(h, src, v.levels[l].b_rgb, cout, st))) return rc; return done(); }
  }
  if (dec && sscanf(nm.c_str(), "upsamples.%d.%15s", &l, kind) == 2 && l >= 1 && l < S) {
    const int C = cfg.channels[l], Cp = cfg.channels[l - 1];
    if (!strcmp(kind, "weight")) {   // ConvTranspose2d weight [in=Cp, out=C, 2, 2] -> rows n = (dy*2+dx)*C + co, cols ci
      if (!shape_is(shape, ndim, {Cp, C, 2, 2})) return bad_shape();
      if ((rc = repack(h, src, v.levels[l].w_resample, true, Cp, C, 4, 4LL * C, 4, 1, 1, Cp, (long long)C * Cp, st))) return rc;
      return done();
    }
    if (!strcmp(kind, "bias")) {
      if (!shape_is(shape, ndim, {C})) return bad_shape();
      if ((rc = repack(h, src, v.levels[l].b_resample, false, 4, C, 1, 0, 1, 0, C, 1, 0, st))) return rc;
      return done();
    }
  }
  if (!dec && sscanf(nm.c_str(), "downsamples.%d.1.%15s", &l, kind) == 2 && l >= 0 && l < S - 1) {
    const int C = cfg.channels[l], Cn = cfg.channels[l + 1];
    if (!strcmp(kind, "weight")) { if (!shape_is(shape, ndim, {Cn, C})) return bad_shape(); if ((rc = copy_t(h, src, v.levels[l].w_resample, (long long)Cn * C, st))) return rc; return done(); }
    if (!strcmp(kind, "bias")) { if (!shape_is(shape, ndim, {Cn})) return bad_shape(); if ((rc = copy_f(h, src, v.levels[l].b_resample, Cn, st))) return rc; return done(); }
  }
  return fail(h, LDMB_ERR_INVALID, "unexpected key %s", name);
}

namespace {

int vae_reserve(ldmb_handle* h, int which, int B, int H0, int W0) {   // H0,W0: resolution of level 0
  VaeState& v = h->vae[which];
  const bool dec = which == LDMB_VAE_DECODER;
  const int S = v.cfg.num_levels;
  size_t mx = 0, mx_rgb = 0;
  for (int l = 0; l < S; ++l) {
    const size_t Hl = dec ? ((size_t)H0 << l) : ((size_t)H0 >> l), Wl = dec ? ((size_t)W0 << l) : ((size_t)W0 >> l);
    const size_t n = (size_t)B * Hl * Wl * v.cfg.channels[l];
    if (n > mx) mx = n;
    if (dec && l < S - 1) { const size_t r = (size_t)B * Hl * Wl * v.cfg.image_channels; if (r > mx_rgb) mx_rgb = r; }
  }
  int rc;
  for (auto& b : v.act) if ((rc = ensure(h, b, mx * h->tsize()))) return rc;
  if (dec) for (auto& b : v.rgb) if ((rc = ensure(h, b, (mx_rgb ? mx_rgb : 64) * 4))) return rc;
  return LDMB_OK;
}

// x -> ResBlock (vae.py:60-66): leaky(c2(leaky(c1(x)))) + x.  bufs: x in cur, scratch tmp, result in out.
int res_block(ldmb_handle* h, const ResW& R, const void* x, void* tmp, void* out, int B, int Hl, int Wl, int C, cudaStream_t st) {
  int rc;
  if (h->bf16() && !h->force_simt && conv64_halo_supported(C, C)) {   // C = 64: halo-patch kernel, every activation read once
    const double fl = 2.0 * B * Hl * Wl * (double)C * 9 * C;
    CKLP(PK_VAE_CONV, fl, launch_conv64_halo(h->tc, x, R.w1, R.b1, tmp, nullptr, B, Hl, Wl, kLeaky, st));
    CKLP(PK_VAE_CONV, fl, launch_conv64_halo(h->tc, tmp, R.w2, R.b2, out, x, B, Hl, Wl, kLeaky, st));
    return LDMB_OK;
  }
  GemmDesc d = gd();
  d.A = x; d.lda = C; d.amode = AM_CONV3; d.cH = Hl; d.cW = Wl; d.cC = C; d.W = R.w1; d.ldw = 9LL * C; d.bias = R.b1;
  d.out = tmp; d.ldo = C; d.M = B * Hl * Wl; d.N = C; d.K = 9 * C; d.epi = EPI_STORE; d.act = ACT_LEAKY;
  if ((rc = gemm(h, d, st, PK_VAE_CONV))) return rc;
  d.A = tmp; d.W = R.w2; d.bias = R.b2; d.out = out; d.res = x; d.ldr = C;
  return gemm(h, d, st, PK_VAE_CONV);
}

}  // namespace

extern "C" int ldmb_vae_reserve(ldmb_handle* h, int which, int max_batch, int H, int W) {
  if (!h || which < 0 || which > 1) return LDMB_ERR_INVALID;
  if (!h->vae[which].configured) return fail(h, LDMB_ERR_STATE, "ldmb_vae_configure has not been called");
  CK(cudaSetDevice(h->device));
  return vae_reserve(h, which, max_batch, H, W);
}

extern "C" int ldmb_vae_decode(ldmb_handle* h, const float* z_dev, float* img_dev, uint8_t* img_u8_dev, int B, int hl,
                               int wl, void* stream) {
  if (!h || !z_dev || (!img_dev && !img_u8_dev)) return LDMB_ERR_INVALID;
  CHECK_FAULT();
  VaeState& v = h->vae[LDMB_VAE_DECODER];
  if (!v.configured) return fail(h, LDMB_ERR_STATE, "decoder not configured");
  if (!v.missing.empty())
    return fail(h, LDMB_ERR_STATE, "%d decoder parameters not loaded (first: %s)", (int)v.missing.size(), v.missing.begin()->c_str());
  if (B < 1 || hl < 1 || wl < 1) return fail(h, LDMB_ERR_INVALID, "decode: bad sizes");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ldmb_vae_config& cfg = v.cfg;
  const int S = cfg.num_levels;
  int rc;
  if ((rc = vae_reserve(h, LDMB_VAE_DECODER, B, hl, wl))) return rc;
  void *cur = v.act[0].p, *t1 = v.act[1].p, *t2 = v.act[2].p;
  CKL(launch_nchw_pointwise_in(z_dev, v.w_in, v.b_in, cur, h->bf16(), B, cfg.latent_channels, hl, wl, cfg.channels[0], st));
  const float* rgb_prev = nullptr;
  for (int l = 0; l < S; ++l) {
    const int Hl = hl << l, Wl = wl << l, C = cfg.channels[l];
    VaeLevelW& L = v.levels[l];
    if (l > 0) {   // ConvTranspose2d(k=2,s=2) = GEMM [M, Cp] x [Cp, 4C] with a 2x2 scatter (vae.py:120)
      const int Cp = cfg.channels[l - 1];
      GemmDesc d = gd();
      d.A = cur; d.lda = Cp; d.W = L.w_resample; d.ldw = Cp; d.bias = L.b_resample; d.out = t1; d.ldo = C;
      d.M = B * (Hl / 2) * (Wl / 2); d.N = 4 * C; d.K = Cp; d.epi = EPI_CONVT; d.ctH = Hl / 2; d.ctW = Wl / 2; d.ctC = C;
      if ((rc = gemm(h, d, st, PK_VAE_GEMM))) return rc;
      std::swap(cur, t1);
    }
    for (const ResW& R : L.res) {
      if ((rc = res_block(h, R, cur, t1, t2, B, Hl, Wl, C, st))) return rc;
      std::swap(cur, t2);
    }
    // to_rgb + running bilinear-upsampled sum (vae.py:103,129-131)
    const bool last = l == S - 1;
    float* rgb_out = last ? img_dev : static_cast<float*>(v.rgb[l & 1].p);
    CKL(launch_nhwc_pointwise_out(cur, h->bf16(), L.w_rgb, L.b_rgb, rgb_prev, rgb_out, last ? img_u8_dev : nullptr, B, Hl, Wl, C,
                                  cfg.image_channels, st));
    rgb_prev = rgb_out;
    if (last && !img_dev && S > 1) { /* u8-only output: nothing further reads rgb_out */ }
  }
  return LDMB_OK;
}

extern "C" int ldmb_vae_encode(ldmb_handle* h, const float* img_dev, float* z_dev, int B, int H, int W, void* stream) {
  if (!h || !img_dev || !z_dev) return LDMB_ERR_INVALID;
  CHECK_FAULT();
  VaeState& v = h->vae[LDMB_VAE_ENCODER];
  if (!v.configured) return fail(h, LDMB_ERR_STATE, "encoder not configured");
  if (!v.missing.empty())
    return fail(h, LDMB_ERR_STATE, "%d encoder parameters not loaded (first: %s)", (int)v.missing.size(), v.missing.begin()->c_str());
  const ldmb_vae_config& cfg = v.cfg;
  const int S = cfg.num_levels;
  if (B < 1 || H < (1 << (S - 1)) || W < (1 << (S - 1))) return fail(h, LDMB_ERR_INVALID, "encode: bad sizes");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = vae_reserve(h, LDMB_VAE_ENCODER, B, H, W))) return rc;
  void *cur = v.act[0].p, *t1 = v.act[1].p, *t2 = v.act[2].p;
  CKL(launch_nchw_pointwise_in(img_dev, v.w_in, v.b_in, cur, h->bf16(), B, cfg.image_channels, H, W, cfg.channels[0], st));
  int Hl = H, Wl = W;
  for (int l = 0; l < S; ++l) {
    const int C = cfg.channels[l];
    VaeLevelW& L = v.levels[l];
    for (const ResW& R : L.res) {
      if ((rc = res_block(h, R, cur, t1, t2, B, Hl, Wl, C, st))) return rc;
      std::swap(cur, t2);
    }
    if (l < S - 1) {   // AvgPool2d(2) then Conv1x1 (vae.py:87-89)
      const int Cn = cfg.channels[l + 1];
      CKL(launch_pool_t(cur, t1, h->bf16(), B, Hl, Wl, C, st));
      Hl /= 2; Wl /= 2;
      GemmDesc d = gd();
      d.A = t1; d.lda = C; d.W = L.w_resample; d.ldw = C; d.bias = L.b_resample; d.out = t2; d.ldo = Cn;
      d.M = B * Hl * Wl; d.N = Cn; d.K = C; d.epi = EPI_STORE;
      if ((rc = gemm(h, d, st, PK_VAE_GEMM))) return rc;
      std::swap(cur, t2);
    }
  }
  CKL(launch_nhwc_pointwise_out(cur, h->bf16(), v.w_out, v.b_out, nullptr, z_dev, nullptr, B, Hl, Wl, cfg.channels[S - 1],
                                cfg.latent_channels, st));
  return LDMB_OK;
}

// =====================================================================================
// kernel-level entry points (tests / bench)
// =====================================================================================
extern "C" int ldmb_gemm(ldmb_handle* h, const void* A, const void* W, const float* bias, void* out, int M, int N, int K,
                         int out_f32, int act, int force_simt, void* stream) {
  if (!h || !A || !W || !out || M < 1 || N < 1 || K < 1) return LDMB_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  GemmDesc d = gd();
  d.A = A; d.lda = K; d.W = W; d.ldw = K; d.bias = bias; d.out = out; d.ldo = N; d.M = M; d.N = N; d.K = K;
  d.epi = out_f32 == 0 ? EPI_STORE : (out_f32 == 1 ? EPI_STORE_F32 : EPI_ACCUM_F32);
  d.act = act;
  return gemm(h, d, static_cast<cudaStream_t>(stream), PK_FFN_AB, force_simt != 0);
}

extern "C" int ldmb_conv3x3(ldmb_handle* h, const void* in, const void* W, const float* bias, void* out, int B, int H,
                            int Wd, int C, int N, int act, int force_simt, void* stream) {
  if (!h || !in || !W || !out || B < 1 || H < 1 || Wd < 1 || C < 1 || N < 1) return LDMB_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  if (h->bf16() && !h->force_simt && !force_simt && bias && conv64_halo_supported(C, N)) {
    const float slope = act == ACT_RELU ? 0.f : (act == ACT_LEAKY ? kLeaky : 1.f);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CKLP(PK_VAE_CONV, 2.0 * B * H * Wd * (double)C * 9 * N, launch_conv64_halo(h->tc, in, W, bias, out, nullptr, B, H, Wd, slope, st));
    return LDMB_OK;
  }
  GemmDesc d = gd();
  d.A = in; d.lda = C; d.amode = AM_CONV3; d.cH = H; d.cW = Wd; d.cC = C; d.W = W; d.ldw = 9LL * C; d.bias = bias;
  d.out = out; d.ldo = N; d.M = B * H * Wd; d.N = N; d.K = 9 * C; d.epi = EPI_STORE; d.act = act;
  return gemm(h, d, static_cast<cudaStream_t>(stream), PK_VAE_CONV, force_simt != 0);
}

extern "C" int ldmb_channelnorm_film(ldmb_handle* h, const float* x, const float* film, void* out, int M, int C, int HW,
                                     void* stream) {
  if (!h || !x || !film || !out) return LDMB_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CKLP(PK_NORM, (double)M * C * (4 + h->tsize()), launch_norm_film(x, film, nullptr, out, h->bf16(), M, C, HW, kNormEps, nullptr, st));
  return LDMB_OK;
}

extern "C" int ldmb_window_attention(ldmb_handle* h, const void* qkv, const void* xm, const float* b_in, void* att, int64_t ldo,
                                     int B, int H, int W, int C, int win_h, int win_w, int shift, int force_simt, void* stream) {
  if (!h || !qkv || !xm || !b_in || !att || B < 1 || H < 1 || W < 1 || C < kHeadDim || win_h < 1 || win_w < 1 || ldo < C)
    return LDMB_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return window_attention(h, qkv, xm, b_in, att, ldo, B, H, W, C, win_h, win_w, shift, nullptr, st, force_simt);
}

extern "C" int ldmb_grouped_conv3x3(ldmb_handle* h, const void* xm, const void* w_packed, const float* bias, float* x, int B,
                                    int H, int W, int C, int force_generic, void* stream) {
  if (!h || !xm || !w_packed || !bias || !x || B < 1 || H < 1 || W < 1 || C < kHeadDim || C % kHeadDim) return LDMB_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  return grouped_conv(h, xm, w_packed, bias, x, B, H, W, C, nullptr, static_cast<cudaStream_t>(stream), force_generic != 0);
}

extern "C" int ldmb_normconv(ldmb_handle* h, float* x, const float* film, void* xm, const void* w_packed, const float* bias, int B,
                             int H, int W, int C, void* stream) {
  if (!h || !x || !film || !xm || !w_packed || !bias || B < 1 || H < 1 || W < 1) return LDMB_ERR_INVALID;
  if (!h->bf16() || !normconv_supported(B, H, W, C)) return fail(h, LDMB_ERR_UNSUPPORTED, "fused norm + conv: bf16 mode, power-of-two feature maps that fit one 128-row tile, C/64 in {1,2,4,8,16}");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CKLP(PK_GCONV, 2.0 * B * H * W * (double)C * 9 * kHeadDim, launch_normconv(h->tc, x, film, nullptr, xm, w_packed, bias, B, H, W, C, kNormEps, nullptr, st));
  return LDMB_OK;
}

extern "C" int ldmb_mlp_fused(ldmb_handle* h, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                              float* x, int M, int C, int e1, int e2, void* stream) {
  if (!h || !xm || !w_ab || !b_ab || !w_c || !b_c || !x || M < 1 || e1 < 0 || e1 >= kExperts || e2 < 0 || e2 >= kExperts)
    return LDMB_ERR_INVALID;
  if (h->bf16() && ffn_cluster_supported(M, C)) {        // C = 512: the 8-CTA-cluster kernel
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CKLP(PK_FFN_AB, 2.0 * M * (double)C * 9 * C, launch_ffn_cluster(h->tc, xm, w_ab, b_ab, w_c, b_c, x, M, C, 5 * C, nullptr, e1, e2, st));
    return LDMB_OK;
  }
  if (!h->bf16() || !mlp_fused_supported(M, C)) return fail(h, LDMB_ERR_UNSUPPORTED, "fused feed-forward: bf16 mode, C = 128, 256 or 512");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CKLP(PK_FFN_AB, 2.0 * M * (double)C * 9 * C, launch_mlp_fused(h->tc, xm, w_ab, b_ab, w_c, b_c, x, M, C, 5 * C, nullptr, e1, e2, nullptr, 0, st));
  return LDMB_OK;
}

extern "C" int ldmb_mlp_fused_attn(ldmb_handle* h, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                                   const void* att, int64_t ld_att, float* x, int M, int C, int e1, int e2, void* stream) {
  if (!h || !xm || !w_ab || !b_ab || !w_c || !b_c || !att || !x || M < 1 || e1 < 0 || e1 >= kExperts || e2 < 0 || e2 >= kExperts || ld_att < C)
    return LDMB_ERR_INVALID;
  if (!h->bf16() || !mlp_fused_att_supported(M, C))
    return fail(h, LDMB_ERR_UNSUPPORTED, "fused feed-forward + out_proj: bf16 mode, C = 128 (more than one 128-row tile) or 256");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CKLP(PK_FFN_AB, 2.0 * M * (double)C * 10 * C,
       launch_mlp_fused(h->tc, xm, w_ab, b_ab, w_c, b_c, x, M, C, 6 * C, nullptr, e1, e2, nullptr, 0, st, 0, att, ld_att));
  return LDMB_OK;
}
