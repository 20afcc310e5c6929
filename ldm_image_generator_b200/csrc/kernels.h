// Host-side launchers of the libldmb200 kernels.  `is_bf16` selects the activation/weight type T
// (bf16 in LDMB_BF16, float in LDMB_FP32_VALIDATE); the residual stream is always fp32.
#pragma once
#include "common.cuh"

struct TcContext;   // tcgen05 GEMM state (tensor-map encoder entry point, SM count, fault flag)

// ---- GEMM-shaped work
cudaError_t launch_gemm_simt(const GemmDesc& d, bool is_bf16, cudaStream_t s);
// Returns cudaErrorNotSupported when the shape cannot be tiled for tcgen05 (caller falls back to
// the CUDA-core kernel, which only happens for toy channel counts).
TcContext* tc_context_create(int device, char* err, int errlen);
void tc_context_destroy(TcContext*);
cudaError_t launch_gemm_tc(TcContext* ctx, const GemmDesc& d, cudaStream_t s);
bool tc_supported(const GemmDesc& d);
int tc_read_fault(TcContext* ctx, cudaStream_t s);   // synchronises `s`
int tc_poll_fault(const TcContext* ctx);              // host-mapped mirror: no synchronisation (faults of kernels that have run)
// true when launch_gemm_tc would update d.out (EPI_ACCUM_F32) through TMA reduce-adds only -- the precondition for running
// another updater of the same tensor concurrently (the forked grouped conv)
bool tc_accum_is_reduction(const GemmDesc& d);
void tc_set_splitk(TcContext* ctx, bool on);
// debug: per-CTA %globaltimer stamps (16 slots per CTA) of the most recent tcgen05 launch
int tc_trace_enable(TcContext* ctx, int on);
int tc_trace_read(TcContext* ctx, long long* host, int max_ctas);

// Grouped 3x3 convolution with one halo-patch load per tile (kernels_gconv.cu): x fp32 [B,H,W,C] += conv(xm) + bias;
// w packed as block-diagonal pairs of 32-channel groups [C/64][64][9*64].  plan: {skip,...} of the block or NULL.
bool gconv_halo_supported(int B, int H, int W, int C);
// max_ctas > 0 caps the persistent grid (rounded down to a whole number of channel slices): the conv then fits beside a GEMM
// that leaves that many SMs idle
// part / split_permille: 0 = the whole conv; 1 = only the first split_permille / 1000 of the spatial tiles, 2 = only the rest
cudaError_t launch_gconv_halo(TcContext* ctx, const void* xm, const void* w, const float* bias, float* x, int B, int H, int W,
                              int C, const int* plan, cudaStream_t st, int max_ctas = 0, int part = 0, int split_permille = 0);
// CTAs launch_gemm_tc puts on the machine for this GEMM (0: not a tcgen05 shape)
int tc_gemm_ctas(TcContext* ctx, const GemmDesc& d);
int tc_num_sms(const TcContext* ctx);

// ChannelNorm + FiLM + grouped 3x3 conv in one kernel (kernels_normconv.cu) for feature maps that fit one 128-row tile:
// xm(bf16) = FiLM(norm(x)); x += conv(xm) + bias in place.  plan: {skip,...} of the block or NULL.
bool normconv_supported(int B, int H, int W, int C);
bool normconv_in_step();      // whether the UNet step uses it (environment knob; off: it is slower than the separate kernels today)
cudaError_t launch_normconv(TcContext* ctx, float* x, const float* film, const int* t_index, void* xm, const void* w, const float* bias,
                            int B, int H, int W, int C, float eps, const int* plan, cudaStream_t st);

// Fused ReGLU feed-forward (kernels_mlp.cu), C = 128 / 256: x fp32 [M,C] += sum over {general, e1, e2} of
// c_e(a_e(xm) * relu(b_e(xm))).  Weight layouts as in the two-GEMM path (w_ab [5*2C, C] a|b interleaved in chunks of 64,
// w_c [w_c_rows >= 5C, C]).  plan: {skip, e1, e2, -} on the device, or NULL to use e1 / e2.
bool mlp_fused_supported(int M, int C);
cudaError_t launch_mlp_fused(TcContext* ctx, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                             float* x, int M, int C, int w_c_rows, const int* plan, int e1, int e2, const int* plan_img, int rows_per_image,
                             cudaStream_t st, int max_ctas = 0,       // max_ctas > 0: cap of the persistent grid (SM partitioning)
                             const void* att = nullptr, long long ld_att = 0);   // attention blocks: x += att . W_out^T (rows 5C .. 6C of w_c, bias b_c[5C ..]) in the same kernel
bool mlp_fused_att_supported(int M, int C);
bool mlp_fused_per_image_supported(int M, int C, int rows_per_image);
// The same feed-forward at C = 512 (kernels_ffn_cluster.cu): clusters of four CTA pairs share a 256-row tile, every pair gates a
// quarter of the hidden chunks and exchanges them with its peers through distributed shared memory, then accumulates its quarter
// of the output columns.  One shared plan for the batch (no per-image decisions).
bool ffn_cluster_supported(int M, int C);
cudaError_t launch_ffn_cluster(TcContext* ctx, const void* xm, const void* w_ab, const float* b_ab, const void* w_c, const float* b_c,
                               float* x, int M, int C, int w_c_rows, const int* plan, int e1, int e2, cudaStream_t st);

// Dense 3x3 conv with 64 input and 64 output channels through the same halo-patch kernel (VAE level at full resolution):
// out bf16 [B,H,W,64] = act(conv(in) + bias) (+ res), act(v) = max(v,0) + slope*min(v,0); w [64][9*64] tap-major.
bool conv64_halo_supported(int C, int N);
cudaError_t launch_conv64_halo(TcContext* ctx, const void* in, const void* w, const float* bias, void* out, const void* res, int B,
                               int H, int W, float slope, cudaStream_t st);

// ---- weight repack: dst[T] (4-d, dst strides) = src[fp32] (4-d, src strides)
cudaError_t launch_repack(const float* src, void* dst, bool dst_bf16, const int dims[4],
                          const long long sstr[4], const long long dstr[4], cudaStream_t s);

// ---- UNet pieces
// encoder_first (unet.py:77,90): x NCHW fp32 [B,Cin,H*s,W*s] -> out fp32 [B*H*W, C0]; w fp32 [C0][Cin*s*s]
// x is read through sp->x_in (device-side step parameters)
cudaError_t launch_stem(const StepParams* sp, const float* w, const float* bias, float* out,
                        int B, int Cin, int H, int W, int s, int C0, bool tf32, cudaStream_t st);
// ChannelNorm + FiLM (modules.py:23-25, unet.py:22): out(T)[m,c] = norm(x[m,:])[c]*film[row,c] + film[row,C+c]
// film row = t_index[m / HW] * HW + m % HW  (t_index may be NULL -> 0)
// skip (device, may be NULL): non-zero => the block is skipped this step and the kernel exits
cudaError_t launch_norm_film(const float* x, const float* film, const int* t_index, void* out, bool is_bf16,
                             int M, int C, int HW, float eps, const int* skip, cudaStream_t st);
// emb(T)[ti*HW + p][0:C] = pe[p][:],  [C:2C] = te[ti][:]      (unet.py:19)
cudaError_t launch_emb_build(const float* pe, const float* te, void* emb, bool is_bf16, int n_t, int HW, int C,
                             cudaStream_t st);
// AvgPool2d(2) on NHWC fp32 -> T (feeds the encoder ch_conv GEMM; pool and 1x1 conv commute, unet.py:83)
cudaError_t launch_pool_cast(const float* x, void* out, bool is_bf16, int B, int H, int W, int C, cudaStream_t st);
cudaError_t launch_cast(const float* x, void* out, bool is_bf16, long long n, cudaStream_t st);
// Window attention core (attention.py:13-85 + torch MHA): qkv(T) [M,3C] (+ in_proj bias for pad tokens),
// key bias from xm(T) channel 0 when shift != 0, writes att(T) rows of stride ldo
cudaError_t launch_window_attention(const void* qkv, const void* xm, const float* b_in, void* att, long long ldo,
                                    bool is_bf16, int B, int H, int W, int C, int head_dim, int win_h, int win_w,
                                    int shift, const int* skip, cudaStream_t st, bool force_simt = false);
// bf16 tensor-path (mma.sync) implementation of the same contract, 4 heads per CTA (kernels_attn.cu)
bool window_attention_mma_supported(int C, int head_dim, int win_h, int win_w, long long ldo);
cudaError_t launch_window_attention_mma(const void* qkv, const void* xm, const float* b_in, void* att, long long ldo, int B,
                                        int H, int W, int C, int win_h, int win_w, int shift, const int* skip,
                                        cudaStream_t st);
// tcgen05 implementation of the same contract (kernels_attn_tc.cu): S = q k^T and O = P v as UMMA tiles over 128-row tiles
// that pack several windows, softmax from TMEM; two heads (64 channels) per work item
bool window_attention_tc_supported(int B, int H, int W, int C, int head_dim, int win_h, int win_w, long long ldo);
cudaError_t launch_window_attention_tc(TcContext* ctx, const void* qkv, const void* xm, const float* b_in, void* att, long long ldo,
                                       int B, int H, int W, int C, int win_h, int win_w, int shift, const int* skip,
                                       cudaStream_t st);
// decoder_last ConvTranspose (unet.py:78,102) fused with the DDIM update (ddpm.py:81-91).
// x fp32 [B*H*W, C0]; w fp32 [C0][Cin*s*s]; x_in/out/noise NCHW fp32 [B,Cin,H*s,W*s]; ddim_enabled 0 = eps only.
// sp (device): x_in / out / noise pointers and the DDIM scalars of this step.
cudaError_t launch_final(const float* x, const float* w, const float* bias, const StepParams* sp,
                         int B, int Cin, int H, int W, int s, int C0, bool tf32, cudaStream_t st);

// ---- per-image plans (plan word per image: skip | e1 << 8 | e2 << 16; see kernels_simt.cu)
cudaError_t launch_rows_enter(float* x, float* backup, const float* b_c, const int* plan_img, int M, int HW, int C, bool attn,
                              cudaStream_t st);
cudaError_t launch_rows_leave(float* x, const float* backup, const int* plan_img, int M, int HW, int C, cudaStream_t st);

// ---- VAE pieces
// 1x1 conv from an NCHW fp32 tensor with few channels (latent 8 / RGB 3) into NHWC T
cudaError_t launch_nchw_pointwise_in(const float* x, const float* w, const float* bias, void* out, bool is_bf16,
                                     int B, int Cin, int H, int W, int Cout, cudaStream_t st);
// 1x1 conv from NHWC T down to few channels, NCHW fp32 out; optionally + bilinear_up2(prev) (vae.py:131) and
// the uint8 HWC image of sample_ldm.py:75-77
cudaError_t launch_nhwc_pointwise_out(const void* x, bool is_bf16, const float* w, const float* bias,
                                      const float* prev, float* out, uint8_t* out_u8,
                                      int B, int H, int W, int C, int Cout, cudaStream_t st);
// AvgPool2d(2) NHWC T -> T (vae.py:88)
cudaError_t launch_pool_t(const void* x, void* out, bool is_bf16, int B, int H, int W, int C, cudaStream_t st);
