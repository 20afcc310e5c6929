/*
 * ldmb.h -- C ABI of libldmb200.so: the B200 (sm_100a) implementation of the
 * latent-diffusion sampling hot path of uthree/ldm-image-generator.
 *
 * The reference has no FFI of its own (it is pure PyTorch); the drop-in boundary is
 * its Python module API.  This header is what the drop-in Python
 * modules in ldm_image_generator_b200/ bind with ctypes, one entry point per
 * reference call site it replaces:
 *
 *   ldmb_unet_forward   <- UNet.forward            /root/reference/unet.py:89-103
 *                          (+ SwinBlock unet.py:38-48, Encodings unet.py:18-23,
 *                           RandomMoE/ReGLU/ChannelNorm modules.py:14-36,
 *                           WindowAttention attention.py:13-85)
 *   ldmb_unet_forward with a ldmb_ddim_coef
 *                       <- one iteration of DDPM.sample   ddpm.py:76-91
 *   ldmb_vae_decode     <- Decoder.forward         vae.py:122-132 (+ sample_ldm.py:75-77 for the u8 output)
 *   ldmb_vae_encode     <- Encoder.forward         vae.py:91-96
 *   ldmb_*_load_param   <- nn.Module.load_state_dict with the reference key layout
 *
 * Rules of the ABI
 *   - extern "C", plain pointers and sizes; no C++/torch types cross it.
 *   - every call returns an int status (LDMB_OK == 0); ldmb_last_error(h) explains a failure.
 *   - the CALLER owns every tensor passed in; "device" pointers are CUDA device memory on
 *     the handle's device, "host" pointers are ordinary host memory read before the call returns.
 *   - calls are stream-ordered on the cudaStream_t passed as `void* stream` and never
 *     synchronise the device, except where a workspace has to grow (ldmb_*_reserve avoids that).
 *   - one handle per GPU per process; a handle is not thread-safe.
 *   - there is no CPU fallback: without a CUDA device ldmb_create fails.
 *
 * Tensor layouts at the boundary are the reference's: NCHW fp32, contiguous.
 */
#ifndef LDMB_H_
#define LDMB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDMB_MAX_LEVELS 8
#define LDMB_ABI_VERSION 1

typedef struct ldmb_handle ldmb_handle;

enum ldmb_status {
  LDMB_OK = 0,
  LDMB_ERR_INVALID = 1,      /* bad argument / shape the reference would also reject */
  LDMB_ERR_CUDA = 2,         /* a CUDA runtime/driver call failed */
  LDMB_ERR_STATE = 3,        /* call order: not configured, parameters missing, ... */
  LDMB_ERR_UNSUPPORTED = 4,
  LDMB_ERR_KERNEL = 5        /* a kernel reported an internal fault (pipeline watchdog) */
};

enum ldmb_precision {
  LDMB_BF16 = 0,             /* bf16 operands, fp32 accumulate, fp32 residual stream; tcgen05 GEMMs */
  LDMB_FP32_VALIDATE = 1     /* everything fp32 on CUDA cores; the <=1e-5 validation mode */
};

/* unet.py:75  UNet(input_channels, stages, channels, stem_size) */
typedef struct ldmb_unet_config {
  int32_t input_channels;
  int32_t num_levels;
  int32_t stem_size;
  int32_t blocks[LDMB_MAX_LEVELS];     /* `stages`  */
  int32_t channels[LDMB_MAX_LEVELS];   /* `channels`, each a multiple of 32 (head_dim, unet.py:26) */
} ldmb_unet_config;

/* vae.py:110 Decoder(output_channels, latent_channels, channels, stages)
 * vae.py:77  Encoder(input_channels,  latent_channels, channels, stages) */
typedef struct ldmb_vae_config {
  int32_t image_channels;
  int32_t latent_channels;
  int32_t num_levels;
  int32_t channels[LDMB_MAX_LEVELS];
  int32_t blocks[LDMB_MAX_LEVELS];     /* `stages` */
} ldmb_vae_config;

/* The five fp32 scalars of one DDIM iteration, computed by the caller on the CPU exactly
 * as ddpm.py:81-85 does:   x0 = (x - c_eps_in*eps) / c_div
 *                          x' = final ? x0 : c_x0*x0 + c_eps_out*eps + sigma*noise        */
typedef struct ldmb_ddim_coef {
  float c_eps_in, c_div, c_x0, c_eps_out, sigma;
  int32_t final_step;                  /* t == 0 (ddpm.py:88-89) */
} ldmb_ddim_coef;

enum ldmb_vae_which { LDMB_VAE_DECODER = 0, LDMB_VAE_ENCODER = 1 };

int ldmb_abi_version(void);

/* Create a handle on CUDA device `device`.  Fails (LDMB_ERR_CUDA) when there is no device. */
int ldmb_create(int device, int precision, ldmb_handle** out);
void ldmb_destroy(ldmb_handle* h);
const char* ldmb_last_error(const ldmb_handle* h);
int ldmb_precision_of(const ldmb_handle* h);

/* Debug switch: route every bf16 GEMM/convolution through the CUDA-core kernels instead of
 * tcgen05 (to bisect a parity failure).  Never set on a measured run. */
int ldmb_set_force_simt(ldmb_handle* h, int on);
/* CUDA-graph replay of the UNet step (default on): the launch sequence of a (batch, resolution, n_t) shape is
 * captured the second time it is seen and replayed afterwards; everything that changes per step lives in a
 * device-side step buffer.  Off = every kernel launched individually. */
int ldmb_set_use_graphs(ldmb_handle* h, int on);
/* Number of kernels this library has launched on the handle since creation (graph replays count their kernels). */
/* bf16 mode adds the branch outputs of a block into the fp32 residual stream with L2 reductions; with split-K GEMM
 * slices and the grouped conv running as a concurrent graph branch their arrival order varies, so two runs agree only
 * to fp32 rounding of x (then re-quantised by later bf16 roundings: ~1e-3 rel-L2 at the default UNet, inside the 1e-2
 * parity budget).  on != 0: no split-K, no concurrent branch -> bit-reproducible results (a few % slower).
 * Also enabled by the environment variable LDMB_DETERMINISTIC. */
int ldmb_set_deterministic(ldmb_handle* h, int on);
int64_t ldmb_launch_count(const ldmb_handle* h);
/* 0, or the watchdog code a tcgen05 pipeline wrote when an mbarrier wait timed out (device read; synchronises). */
int ldmb_check_device_fault(ldmb_handle* h, void* stream);
/* The same word through its host-mapped mirror: no synchronisation, so it only reflects kernels that have already run.
 * Once it is non-zero every later ldmb_unet_* / ldmb_vae_* compute call on the handle returns LDMB_ERR_KERNEL. */
int ldmb_poll_device_fault(const ldmb_handle* h);

/* Debug: %globaltimer stamps inside the tcgen05 kernel.  stamps_host == NULL: enable/disable recording.
 * Otherwise copies 16 int64 stamps per CTA of the most recent tcgen05 launch (synchronises) and returns the CTA count:
 * 0 entry, 1 setup done, 2 previous kernel complete, 3 first TMA issued, 4 first operands landed, 5 all MMAs issued,
 * 6 first accumulator ready, 7 latest accumulator ready, 8 epilogue done, 9 exit. */
int ldmb_debug_tc_trace(ldmb_handle* h, int enable, int64_t* stamps_host, int max_ctas);

/* Per-kernel-class device timing for roofline reports (bench.py).  Between begin and end every launch is
 * bracketed by CUDA events on its stream; end synchronises and returns, per class, the summed event time (ms),
 * the summed algorithmic work (FLOPs for the GEMM/conv classes, bytes for the HBM-bound ones) and the launch count.
 * Classes (tcgen05 unless noted): 0 ReGLU a|b GEMM (C = 128/256: the fused a|b -> gate -> c kernel), 1 ReGLU c (+ attention out_proj) GEMM into the residual stream,
 * 2 attention in_proj GEMM, 3 hoisted Encodings MLP GEMMs, 4 level-change 1x1 convs, 5 grouped 3x3 conv,
 * 6 VAE dense 3x3 conv, 7 VAE ConvTranspose / 1x1 GEMMs, 8 CUDA-core GEMM/conv (fp32 validation mode, toy widths),
 * 9 ChannelNorm+FiLM, 10 window attention core, 11 stem + final (decoder_last + DDIM update), 12 other.
 * Arrays must hold LDMB_PROFILE_CLASSES entries. */
#define LDMB_PROFILE_CLASSES 13
int ldmb_profile_begin(ldmb_handle* h);
int ldmb_profile_end(ldmb_handle* h, double* ms, double* work, int64_t* launches);
/* Debug (tools/ablate.py): launches of the classes in `mask` (bit k = class k) are skipped -- results are garbage,
 * but timing a step with and without a class gives that class's cost inside the CUDA-graph replay. */
int ldmb_debug_skip_classes(ldmb_handle* h, uint32_t mask);

/* ---------------------------------------------------------------- UNet (unet.py) */
int ldmb_unet_configure(ldmb_handle* h, const ldmb_unet_config* cfg);
/* One state_dict entry, named as in UNet.state_dict() (no "model." prefix), device fp32,
 * contiguous, shape as in the reference.  Repacked (cast, expert-concatenated, K-major) into the
 * library's arena on `stream`.  cross_attention.* entries are accepted and ignored (dead code in
 * the reference, attention.py:92-98). */
int ldmb_unet_load_param(ldmb_handle* h, const char* name, const float* data_dev,
                         const int64_t* shape, int ndim, void* stream);
/* How many parameters ldmb_unet_forward still needs (0 = ready). */
int ldmb_unet_params_missing(const ldmb_handle* h);
/* Pre-size workspaces for batches up to max_batch at input resolution H x W and up to
 * max_t distinct timesteps per call (optional; forward grows them on demand). */
int ldmb_unet_reserve(ldmb_handle* h, int max_batch, int H, int W, int max_t);
/* Position-encoding table of one level, host fp32 [C, Hl, Wl], computed by the caller exactly as
 * sinusoidal.py:12-19 (kept on the host so the sin/cos bits are the reference's). */
int ldmb_unet_set_position_table(ldmb_handle* h, int level, const float* pe_host, int C, int Hl, int Wl,
                                 void* stream);
/*
 * eps = UNet(x, time) and, when `coef` is non-NULL, the DDIM update fused behind it.
 *   x_dev      [B, Cin, H, W] fp32 device
 *   out_dev    [B, Cin, H, W] fp32 device: eps if coef == NULL, else the updated x (may alias x_dev)
 *   t_index    host int32 [B]: row of the time tables each image uses (all 0 in DDPM.sample, ddpm.py:77)
 *   te_host    host: num_levels pointers to fp32 [n_t, C_level] tables = sinusoidal.py:31-39 per distinct t
 *   plan       host int32 [n_blocks][3] = (skip, e1, e2) per SwinBlock in execution order: the Python
 *              `random` decisions of unet.py:39 and modules.py:35, drawn by the caller in lock-step
 *   noise_dev  [B, Cin, H, W] or NULL (required only when coef->sigma != 0)
 * te_host == NULL: the FiLM tables of the n_t timesteps registered by ldmb_unet_precompute_film are used (t_index
 * indexes that table) and the Encodings MLP is not evaluated by this call.
 */
int ldmb_unet_forward(ldmb_handle* h, const float* x_dev, float* out_dev, int B, int H, int W,
                      const int32_t* t_index, int n_t, const float* const* te_host,
                      const int32_t* plan, const ldmb_ddim_coef* coef, const float* noise_dev,
                      void* stream);

/* The same step with PER-IMAGE stochastic-depth / expert decisions: plan_img is host int32 [n_blocks][B][3] =
 * (skip, e1, e2) of block k for image b.  This is what a batch of the reference's batch-1 calls computes (its sample
 * scripts loop `sample((1, ...))`, sample_ldm.py:71-72 / sample_ddpm.py:35-36, so every image draws its own decisions
 * from Python's `random`).  Exact: all five ReGLU experts are evaluated and the image's choice is applied by masking. */
int ldmb_unet_forward_per_image(ldmb_handle* h, const float* x_dev, float* out_dev, int B, int H, int W,
                                const int32_t* t_index, int n_t, const float* const* te_host,
                                const int32_t* plan_img, const ldmb_ddim_coef* coef, const float* noise_dev,
                                void* stream);

/* Host-only helper of the per-image path (no handle, no device work): the block decisions of n_plans consecutive
 * reference forwards, replayed from a stream of raw MT19937 outputs exactly as CPython's `random` consumes them --
 * per block, `random.random() <= depth_p[k]` iff training[k] (unet.py:39: two 32-bit words -> a 53-bit double), then
 * `random.sample(experts, 2)` unless skipped (modules.py:35: _randbelow(n) then _randbelow(n-1), each
 * getrandbits(bit_length) = word >> (32 - bits) with rejection).  raw: the next n_raw 32-bit outputs of the caller's
 * generator; plan_out: int32 [n_plans][n_blocks][3] = (skip, e1, e2); *used = words consumed (the caller advances its
 * generator by that many).  Returns LDMB_ERR_INVALID when raw runs out (call again with a longer stream). */
int ldmb_host_draw_plans(const uint32_t* raw, int64_t n_raw, int n_plans, int n_blocks, const uint8_t* training,
                         const double* depth_p, int n_experts, int32_t* plan_out, int64_t* used);

/* The Encodings MLP (unet.py:18-21) of every block for ALL n_t timesteps of a sampling schedule in one batched pass
 * (its weights, 58 % of the parameters, are then streamed once per schedule instead of once per step).
 * te_host as in ldmb_unet_forward.  Valid until the next ldmb_unet_load_param / ldmb_unet_forward with te_host != NULL. */
int ldmb_unet_precompute_film(ldmb_handle* h, int H, int W, int n_t, const float* const* te_host, void* stream);

/* ---------------------------------------------------------------- VAE (vae.py) */
int ldmb_vae_configure(ldmb_handle* h, int which, const ldmb_vae_config* cfg);
int ldmb_vae_load_param(ldmb_handle* h, int which, const char* name, const float* data_dev,
                        const int64_t* shape, int ndim, void* stream);
int ldmb_vae_params_missing(const ldmb_handle* h, int which);
int ldmb_vae_reserve(ldmb_handle* h, int which, int max_batch, int H, int W);
/* z_dev [B, L, h, w] -> img_dev [B, 3, h*2^(levels-1), w*2^(levels-1)] fp32 (vae.py:122-132).
 * img_u8_dev, if non-NULL, also receives clamp(-1,1)*127.5+127.5 truncated to uint8 in HWC order
 * ([B, Hout, Wout, 3]), i.e. sample_ldm.py:75-77 fused into the last kernel. img_dev may be NULL then. */
int ldmb_vae_decode(ldmb_handle* h, const float* z_dev, float* img_dev, uint8_t* img_u8_dev,
                    int B, int h_lat, int w_lat, void* stream);
/* img_dev [B, 3, H, W] -> z_dev [B, L, H/2^(levels-1), W/2^(levels-1)] (vae.py:91-96). */
int ldmb_vae_encode(ldmb_handle* h, const float* img_dev, float* z_dev, int B, int H, int W, void* stream);

/* ---------------------------------------------------------------- kernel-level entry points
 * Used by tests/ and bench.py to check and time one kernel in isolation against the oracle.
 * A, W, out are device pointers in the handle's precision (bf16 or fp32) unless noted. */

/* out[M,N] (+)= A[M,K] * W[N,K]^T + bias[N].  out_f32: 0 -> handle precision, 1 -> fp32 store, 2 -> fp32 accumulate.
 * act: 0 none, 1 relu, 2 leaky(0.01).  force_simt: use the CUDA-core kernel even in bf16 mode. */
int ldmb_gemm(ldmb_handle* h, const void* A, const void* W, const float* bias, void* out,
              int M, int N, int K, int out_f32, int act, int force_simt, void* stream);
/* 3x3, pad 1, dense convolution on NHWC: in [B,H,W,C] -> out [B,H,W,N]; W packed [N][9*C] tap-major. */
int ldmb_conv3x3(ldmb_handle* h, const void* in, const void* W, const float* bias, void* out,
                 int B, int H, int Wd, int C, int N, int act, int force_simt, void* stream);
/* ChannelNorm + FiLM (modules.py:23-25, unet.py:22): x fp32 [M,C], film fp32 [HW, 2C] -> out [M,C] in handle precision. */
int ldmb_channelnorm_film(ldmb_handle* h, const float* x, const float* film, void* out, int M, int C, int HW,
                          void* stream);

/* Window attention core (attention.py:13-85 + torch MHA, after in_proj and before out_proj): qkv [B*H*W, 3C] and
 * xm [B*H*W, C] in the handle's precision, b_in fp32 [3C] (in_proj bias: the q/k/v of zero-padded tokens),
 * att [B*H*W, ldo>=C].  window win_h x win_w, shift 0 (bool pad mask) or != 0 (rolled window, float key bias =
 * channel 0 of the rolled xm).  force_simt: 0 = the tcgen05 kernel (bf16 mode; falls back to mma.sync / CUDA cores for shapes it does not
 * cover), 1 = the CUDA-core kernel, 2 = the mma.sync kernel. */
int ldmb_window_attention(ldmb_handle* h, const void* qkv, const void* xm, const float* b_in, void* att, int64_t ldo,
                          int B, int H, int W, int C, int win_h, int win_w, int shift, int force_simt, void* stream);

/* SwinBlock.conv (unet.py:30): x fp32 [B,H,W,C] += conv3x3(xm, groups of 32 channels, pad 1) + bias.
 * xm [B,H,W,C] in the handle's precision.  w_packed: C % 64 == 0 -> block-diagonal pairs of groups [C/64][64][9*64]
 * (row = output channel of the pair, column = tap*64 + input channel of the pair, zeros off the diagonal blocks);
 * otherwise per group [C][9*32] (column = tap*32 + input channel of the group).  force_generic: the 9-tap-load
 * implicit-GEMM kernel instead of the halo-patch kernel. */
int ldmb_grouped_conv3x3(ldmb_handle* h, const void* xm, const void* w_packed, const float* bias, float* x, int B,
                         int H, int W, int C, int force_generic, void* stream);

/* ChannelNorm + FiLM + grouped 3x3 conv of a SwinBlock fused (unet.py:22,30,42-44; modules.py:23-25), for feature maps that fit
 * one 128-row tile (8x8, 4x4): xm bf16 [B,H,W,C] = norm(x) * film[:, :C] + film[:, C:]; x fp32 [B,H,W,C] += conv(xm) + bias in place.
 * film fp32 [H*W][2C]; w_packed as for ldmb_grouped_conv3x3.  LDMB_ERR_UNSUPPORTED for other shapes (the UNet then runs the
 * separate kernels). */
int ldmb_normconv(ldmb_handle* h, float* x, const float* film, void* xm, const void* w_packed, const float* bias, int B, int H,
                  int W, int C, void* stream);
/* RandomMoE of ReGLU experts (modules.py:14-15,34-36) as one fused kernel, bf16 mode, C = 128 or 256:
 * x fp32 [M,C] += sum over {general, experts e1, e2} of c_e(a_e(xm) * relu(b_e(xm))).
 * w_ab [5*2C, C]: per expert block (general first) the rows of a and b interleaved in chunks of 64 (64 a rows, 64 b rows, ...),
 * b_ab fp32 [5*2C] alike; w_c [5*C, C] (row = expert*C + output channel), b_c fp32 [5*C]. */
int ldmb_mlp_fused(ldmb_handle* h, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                   float* x, int M, int C, int e1, int e2, void* stream);

/* The same kernel for an attention block (unet.py:44,47 + attention.py:82): additionally x += att . W_out^T + b_out, the MHA out_proj
 * of the attention core's output att bf16 [M, C] (row stride ld_att elements), as extra K-chunks of the c-projection accumulator.
 * w_c [6*C, C] / b_c [6*C]: rows 5C .. 6C are out_proj.weight / out_proj.bias.  C = 128 (M > 128) or 256. */
int ldmb_mlp_fused_attn(ldmb_handle* h, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                        const void* att, int64_t ld_att, float* x, int M, int C, int e1, int e2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDMB_H_ */
