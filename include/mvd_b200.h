/*
 * mvd_b200.h — C ABI of libmvd_b200.so: the hand-written sm_100a kernels behind the MVD-Fusion
 * multi-view denoising hot path.
 *
 * The reference (zhizdev/mvdfusion) is pure PyTorch and has no FFI layer; each entry point below
 * names the reference call site(s) (file:line under the reference root) whose library dispatch
 * (cuDNN / cuBLAS / ATen) it replaces.  Conventions:
 *   - plain device pointers + explicit sizes; no torch types; the caller owns every buffer
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     never allocates device memory, and is CUDA-graph capturable
 *   - returns 0 on success, a negative MVD_E* code otherwise; mvd_last_error() gives the text
 *   - activations are "rows x channels" (NHWC flattened): row = (image*H + y)*W + x
 *   - fp16 tensors are IEEE binary16 (`__half`), "f32" is float
 */
#ifndef MVD_B200_H_
#define MVD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVD_OK 0
#define MVD_EINVAL (-1)  /* bad argument / unsupported shape */
#define MVD_ECUDA (-2)   /* CUDA runtime or driver error */
#define MVD_EALIGN (-3)  /* pointer or leading dimension not aligned as required */

const char* mvd_last_error(void);
int mvd_abi_version(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
long long mvd_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM / implicit-GEMM convolution (tcgen05.mma, TMA, TMEM accumulators).
 *
 *   acc[m, n] = sum_k A[m, k] * W[n, k]          (fp16 operands, fp32 accumulate)
 *
 * a_mode = MVD_A_ROWMAJOR : A is fp16 [M, lda] row-major, k < K.
 *   replaces nn.Linear (external/sd1/ldm/modules/attention.py:40,60,161-168; mvdfusion/attention.py:100,114;
 *   mvdfusion/view_attn_efficient2.py:158,162; timm Attention/Mlp) and the 1x1 convs
 *   (external/sd1/ldm/modules/attention.py:245,259; openaimodel.py:241).
 * a_mode = MVD_A_CONV3X3 : A is an fp16 NHWC image batch [n_img, H, W, C]; stride 1, zero pad 1;
 *   K = 9*C with k = (ky*3 + kx)*C + c; M = n_img*H*W.  W is [N, 9*C] in that k order.
 *   replaces conv_nd(…,3,padding=1) (openaimodel.py:107,204,230; mvdfusion/unet.py:323,499).
 *
 * Epilogue (per element, in this order):
 *   v = acc + bias[n] + rowbias[(m / rows_per_group), n] ; v = act(v) ; v *= colscale[n] ;
 *   v += residual[m, n] ; store        (colscale = the adaLN gate of a DiT block, view_attn_efficient2.py:65-66)
 *   act: NONE, GELU (exact erf), SILU, GEGLU (weights pre-interleaved per tile so that column
 *        j and column j + BN/2 of a tile are value/gate; output has N/2 columns:
 *        out = value * gelu(gate); external/sd1/ldm/modules/attention.py:42-44)
 *   out_mode: F32 / F16 row-major [M, ldc]; QKV_HEADS scatters q,k into [img*heads + h, seq, dpad]
 *        and v transposed into [img*heads + h, dpad, seq] for mvd_attn_self_f16.
 *   split_k: K is cut into slices that run as separate work units; each slice parks its fp32 partial tile in
 *        `splitk_ws`, the slice finishing last sums them and runs the epilogue (any act / out_mode except GEGLU).
 *   The kernel is persistent (one CTA per SM walks the tile list) and keeps two accumulators in TMEM so that the
 *   epilogue of one tile overlaps the MMAs of the next; outputs leave through a swizzled smem staging tile as coalesced
 *   16-byte stores.  Deep-K problems run as CTA pairs (tcgen05 cta_group::2): two m-tiles share each W tile.
 *   tile_n / split_k / cta_pair = 0 let the library choose; mvdfusion_b200/gemm_tuning.json holds measured choices.
 * ---------------------------------------------------------------------------------------------- */
enum { MVD_A_ROWMAJOR = 0, MVD_A_CONV3X3 = 1 };
enum { MVD_ACT_NONE = 0, MVD_ACT_GELU = 1, MVD_ACT_SILU = 2, MVD_ACT_GEGLU = 3 };
enum { MVD_OUT_F32 = 0, MVD_OUT_F16 = 1, MVD_OUT_QKV_HEADS = 2 };

typedef struct mvd_gemm_args {
  int32_t M, N, K;
  int32_t a_mode;
  const void* A;       /* fp16 */
  int32_t lda;         /* elements (multiple of 8); ROWMAJOR: row pitch; CONV3X3: pixel pitch, 0 = C */
  int32_t n_img, H, W, C; /* CONV3X3 only; C multiple of 8, W a power of two (rows wider than 128 pixels are tiled in 128-pixel segments) */
  const void* Wt;      /* fp16 [N, ldw] */
  int32_t ldw;         /* elements, multiple of 8, >= K */
  const float* bias;   /* [N] or NULL */
  const float* rowbias;/* [ceil(M/rows_per_group), N] or NULL */
  int32_t rows_per_group;
  const float* colscale; /* [N] or NULL */
  const float* residual; /* fp32 [M, ldr] or NULL */
  int32_t ldr;
  int32_t act;
  int32_t out_mode;
  void* out;           /* F32/F16: [M, ldc]; QKV_HEADS: q base (fp16) */
  int32_t ldc;
  /* QKV_HEADS only: N = 3*heads*dhead; seq rows per image */
  void* out_k;
  void* out_vt;
  int32_t heads, dhead, dpad, seq;
  int32_t split_k;     /* 0 = auto, 1 = off, > 1 = that many K slices (needs splitk_ws) */
  int32_t tile_n;      /* 0 = auto; else a multiple of 32 in [32, 256], or a multiple of 16 >= N (GEGLU: a multiple of 64 dividing N) */
  int32_t cta_pair;    /* 0 = auto, 1 = one CTA per 128-row tile, 2 = CTA pairs (cta_group::2, 256-row tiles; needs >= 2 m-tiles) */
  void* splitk_ws;     /* caller-owned split-K workspace or NULL; MUST be zero-filled once before its first use
                          (the first 16 KB are self-resetting tile semaphores, the rest holds fp32 partial tiles) */
  long long splitk_ws_bytes;
  /* ABI 7 */
  void* out16;         /* optional (out_mode F32 only): the stored values once more as fp16 [M, ld16] — the operand of the
                          GEMM that consumes this output, written here instead of by a separate cast / concat pass */
  int32_t ld16;
  /* ABI 9: split-precision ("hi/lo") operands for the few GEMMs whose fp16 operand rounding dominates the end-to-end error
   * (the stem / head convolutions and the ResBlock 1x1 skip convolutions: they sit on the residual trunk, not on a branch).
   * hilo = 1: A holds [A_hi | A_lo] (ROWMAJOR: columns [0,K) and [K,2K), lda >= 2K; CONV3X3: 2C channels per pixel) and Wt holds
   *   [W_hi | W_lo] ([N, 2K], ldw >= 2K) with x_hi = fp16(x), x_lo = fp16(x - x_hi); the kernel accumulates
   *   A_hi W_hi + A_lo W_hi + A_hi W_lo (three passes over K in one launch; K, C multiples of 64).
   * out16_lo > 0: next to the fp16 copy `out16` of the output, its rounding residual fp16(v - fp16(v)) is stored out16_lo
   *   columns to the right (the [hi | lo] operand of a consuming hilo GEMM). */
  int32_t hilo;
  int32_t out16_lo;
  int32_t a_lo_off;    /* ROWMAJOR + hilo: column of A_lo (0 = K); > K when A is a column window of a wider [hi | lo] buffer */
  /* ABI 10: strided implicit-GEMM convolution — Downsample.op, conv3x3 stride 2 padding 1 (openaimodel.py:151), reads its taps
   * straight from the full-resolution image through a TMA box with element strides 2 (no im2col).  conv_stride = 2: H, W are the
   * OUTPUT extent, A is the fp16 NHWC image [n_img, 2H, 2W, C].  conv_no_pad_lo = 1: pad the high side only (the VAE encoder's
   * F.pad(x, (0,1,0,1)) + stride-2 conv, external/sd1/ldm/modules/diffusionmodules/model.py:65-76).  In CONV3X3 mode `lda`, when
   * non-zero, is the pixel pitch of A in elements (the image may be a column window of a wider buffer). */
  int32_t conv_stride;
  int32_t conv_no_pad_lo;
  /* ABI 13: nn.LayerNorm between two GEMMs without a pass of its own (external/sd1/ldm/modules/attention.py:211-213,220-222 norm1 / norm3
   * in front of to_q|to_k|to_v and GEGLU.proj; mvdfusion/attention.py:35-37,52,64).
   *   producer side — ln_stats_out: fp32 pairs [N/32][M][2]; for every output row and every 32-column chunk the epilogue stores
   *     (sum, sum of squares) of the values it writes (F32 output through the TMA epilogue: unsplit, K <= 1536, N % 32 == 0; the call
   *     fails otherwise).  Together with out16 (the raw rows as fp16) this is everything the consumer needs.
   *   consumer side — ln_stats (what a producer wrote: [K/32][M][2], K % 32 == 0), ln_colsum, ln_eps: A holds the RAW rows x (fp16),
   *     Wt the gamma-scaled weights W' = W diag(gamma), ln_colsum[n] = sum_k W'[n, k] (of the fp16-rounded W'), bias already contains
   *     W beta.  Per row: mean = sum / K, rstd = rsqrt(sumsq / K - mean^2 + ln_eps), and the accumulator becomes
   *         acc' = rstd[m] * (acc - mean[m] * ln_colsum[n])  ==  (LayerNorm_noaffine(x) W'^T)[m, n]
   *     before bias / activation.  ROWMAJOR A, QKV_HEADS or GEGLU output, unsplit. */
  float* ln_stats_out;
  const float* ln_stats;
  const float* ln_colsum;
  float ln_eps;
  /* ABI 14: nearest x2 upsample folded into the convolution — Upsample.forward, F.interpolate(scale_factor=2, mode="nearest") followed by
   * conv3x3 padding 1 (external/sd1/ldm/modules/diffusionmodules/openaimodel.py:107-119).  An output pixel (2y + py, 2x + px) sees only
   * the 2 x 2 source pixels (y + py - 1 + a, x + px - 1 + b), a, b in {0, 1}, each with the SUM of the 3 x 3 taps that land on it:
   * four 2 x 2 convolutions of the source image, 16 C instead of 36 C multiply-adds per output value, no upsampled tensor.
   * conv_up2 = 1 (CONV3X3): A is the SOURCE image fp16 [n_img, H, W, C] (H, W powers of two), M = n_img * H * W, K = 4 * C,
   *   N = 4 * Cout; Wt is [N, K]: row phase * Cout + n (phase = 2 py + px), column (2 a + b) * C + c, holding
   *   sum over ky in Sa(py), kx in Sb(px) of w[n, c, ky, kx] with S0(0) = {0}, S1(0) = {1, 2}, S0(1) = {0, 1}, S1(1) = {2};
   *   bias is [N] (the convolution's bias once per phase); out / out16 are the UPSAMPLED-resolution matrices [n_img * 2H * 2W, ldc]
   *   with ldc >= Cout.  Bias only (no residual / row bias / activation / split precision); tile_n must divide Cout. */
  int32_t conv_up2;
} mvd_gemm_args;

int mvd_gemm_f16(const mvd_gemm_args* args, void* stream);

/* GEGLU weight interleave used by MVD_ACT_GEGLU for a given tile width: row index of the
 * packed weight -> row index of the original nn.Linear(dim, 2*inner) weight. */
int mvd_geglu_row_permutation(int32_t inner_dim, int32_t tile_n, int32_t* perm_out /* [2*inner_dim] */);


/* ------------------------------------------------------------------------------------------------
 * Fused multi-head self-attention (tcgen05 QK^T and PV, softmax in registers, scores stay in TMEM).
 * Replaces CrossAttention.forward with context=None: external/sd1/ldm/modules/attention.py:170-193
 * (called from :220 and mvdfusion/attention.py:52).
 *   q, k : fp16 [n_img*heads, seq, dpad]   vt : fp16 [n_img*heads, dpad, seq]
 *   out  : fp16 [n_img*seq, ldo], columns h*dhead + j  ('b n (h d)')
 * ---------------------------------------------------------------------------------------------- */
int mvd_attn_self_f16(const void* q, const void* k, const void* vt, void* out, int32_t n_img, int32_t heads,
                      int32_t seq, int32_t dhead, int32_t dpad, int32_t ldo, void* stream);
/* the same with keys [seq_valid, seq) masked out: a sequence whose length is not a multiple of 16 lives in a padded layout
 * (CLIP ViT-L/14: 257 tokens in 272 rows; nn.MultiheadAttention of clip.model.ResidualAttentionBlock, called from
 * external/sd1/ldm/modules/encoders/modules.py:431-436 through model.encode_image) */
int mvd_attn_self_masked_f16(const void* q, const void* k, const void* vt, void* out, int32_t n_img, int32_t heads, int32_t seq,
                             int32_t seq_valid, int32_t dhead, int32_t dpad, int32_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation (fp32 residual stream in, fp16 GEMM operand out).
 *   groupnorm : 32 groups over [n_img, hw, C]; optional SiLU.  util.py:200-217 (eps 1e-5, ResBlock/out),
 *               external/sd1/ldm/modules/attention.py:76-77 (eps 1e-6).  stats_ws: unused since the single-pass kernel (may be NULL).
 *   layernorm : nn.LayerNorm(C) (external/sd1/ldm/modules/attention.py:211-213, mvdfusion/attention.py:35-37)
 *   ln_modulate : LayerNorm(no affine) then x*(1+scale)+shift (mvdfusion/view_attn_efficient2.py:15-16,65-66)
 * ---------------------------------------------------------------------------------------------- */
int mvd_groupnorm_f32_f16(const float* x, const float* gamma, const float* beta, void* y, void* stats_ws,
                          int32_t n_img, int32_t hw, int32_t C, float eps, int32_t apply_silu, void* stream);
/* the same, written as a split-precision operand y fp16 [n_img*hw, 2C] = [hi | lo], hi = fp16(v), lo = fp16(v - hi): the A operand of a
 * `hilo` GEMM (the UNet head, GroupNorm32-SiLU-conv 320->5, mvdfusion/unet.py:496-500, whose operand rounding would otherwise land on
 * the output unattenuated) */
int mvd_groupnorm_hilo_f32_f16(const float* x, const float* gamma, const float* beta, void* y, int32_t n_img, int32_t hw, int32_t C,
                               float eps, int32_t apply_silu, void* stream);
/* GroupNorm of the channel concatenation [x1 | x2] (x1: [n_img, hw, C1], x2: [n_img, hw, C2]) without materialising it:
 * `h = th.cat([h, hs.pop()], dim=1)` followed by ResBlock.in_layers[0] in the UNet's output blocks (mvdfusion/unet.py:550). */
int mvd_groupnorm2_f32_f16(const float* x1, int32_t C1, const float* x2, int32_t C2, const float* gamma, const float* beta,
                           void* y, int32_t n_img, int32_t hw, float eps, int32_t apply_silu, void* stream);
int mvd_layernorm_f32_f16(const float* x, const float* gamma, const float* beta, void* y, int32_t rows, int32_t C,
                          float eps, void* stream);
/* nn.LayerNorm with fp32 output and explicit row pitches (x: [rows, ldx], y: [rows, ldy]; in place allowed): CLIP's ln_pre (the
 * normalised tokens ARE the residual stream) and ln_post on the class-token rows (clip/model.py VisionTransformer.forward) */
int mvd_layernorm_f32_f32(const float* x, long long ldx, const float* gamma, const float* beta, float* y, long long ldy, int32_t rows,
                          int32_t C, float eps, void* stream);
int mvd_ln_modulate_f32_f16(const float* x, const float* shift, const float* scale, void* y, int32_t rows, int32_t C,
                            float eps, void* stream);
/* p[r, :cols] = softmax(scale * s[r, :cols]) as fp16: the VAE decoder's single-head attention weights
 * (external/sd1/ldm/modules/diffusionmodules/model.py:186-190, w_ = softmax(q k^T c^-1/2)); s fp32 [rows, ld_in], p fp16 [rows, ld_out] */
int mvd_softmax_rows_f32_f16(const float* s, void* p, int32_t rows, int32_t cols, int32_t ld_in, int32_t ld_out, float scale,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data movement / elementwise.
 * ---------------------------------------------------------------------------------------------- */
int mvd_cast_f32_f16(const float* x, void* y, long long n, void* stream);
/* torch.cat([h, hs.pop()], dim=1) in rows x channels form (mvdfusion/unet.py:550) */
int mvd_concat_f32(const float* a, const float* b, float* out, long long rows, int32_t C1, int32_t C2, void* stream);
/* the same concatenation as the fp16 operand of the ResBlock's 1x1 skip convolution (openaimodel.py:241,273) */
int mvd_concat_f32_f16(const float* a, const float* b, void* out, long long rows, int32_t C1, int32_t C2, void* stream);
/* F.interpolate(scale_factor=2, mode="nearest") (openaimodel.py:116); fp32 NHWC -> fp16 NHWC */
int mvd_upsample2x_f32_f16(const float* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C, void* stream);
/* im2col of the stride-2 Downsample conv (openaimodel.py:151): fp32 NHWC -> fp16 [n*(H/2)*(W/2), 9*C] */
int mvd_im2col_s2_f32_f16(const float* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C, void* stream);
/* the same with the low-side padding selectable: pad_lo = 0 is the VAE encoder's Downsample, F.pad(x, (0,1,0,1)) + conv3x3 stride 2
 * padding 0 (external/sd1/ldm/modules/diffusionmodules/model.py:65-76); pad_lo = 1 is the call above */
int mvd_im2col_s2_pad_f32_f16(const float* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t pad_lo, void* stream);
/* small-M linear: y = act_out(act_in(x) W^T + b); x fp32 [M, ldx], W fp16 [N, ldw], y fp32 [M, ldy].
 * t-only MLPs: ResBlock.emb_layers (openaimodel.py:218-224), UNetModel.time_embed (unet.py:310-314),
 * ViewFusion.time_embed / cc_projection (viewfusion_zero_depth_rgb.py:107-132), DiT adaLN (view_attn_efficient2.py:58-61) */
int mvd_gemv_f16(const float* x, int32_t ldx, const void* W, int32_t ldw, const float* bias, float* y, int32_t ldy,
                 int32_t M, int32_t N, int32_t K, int32_t silu_in, int32_t silu_out, void* stream);
/* the same for a table of jobs that share ONE input row (M = 1), one launch: the 22 ResBlock.emb_layers of a UNet pass all
 * read SiLU(emb) (openaimodel.py:218-224,266-270).  jobs_dev: int64 [n_jobs, 5] in device memory =
 * {W (fp16 [N, ldw]), bias (fp32 [N] or 0), y (fp32 [N]), N | (int64)ldw << 32, first global column}; columns are
 * numbered consecutively over the jobs, total_cols = sum of N; W rows must be 16-byte aligned (ldw % 8 == 0). */
int mvd_gemv_grouped_f16(const float* x, int32_t K, int32_t silu_in, const void* jobs_dev, int32_t n_jobs,
                         int32_t total_cols, void* stream);
/* timestep_embedding (util.py:152-172): out[dim] = [cos(t f) | sin(t f)], t and f tables in device memory */
int mvd_timestep_embedding(const float* t_dev, const float* freqs_dev, float* out, int32_t dim, void* stream);
/* UNet input assembly incl. the unconditional CFG branch (mvdfusion/unet.py:153-161,173-186) -> fp16 NHWC */
/* cond_scale: optional [n_views] multiplier of the concat channels (condition drop, mvdfusion/unet.py:140-151) */
/* hilo = 1 (Cpad >= 32): channels [0,10) hold fp16(v), [10,20) fp16(v - fp16(v)), [20,30) fp16(v) again — the three K segments of a
 * split-precision stem convolution laid out as plain input channels (weights [W_hi | W_hi | W_lo]) */
int mvd_unet_input_f16(const float* noisy, const float* cond, int32_t cond_batched, const float* cond_scale, void* out,
                       int32_t n_views, int32_t n_img, int32_t hw, int32_t Cpad, int32_t hilo, void* stream);
/* CFG combine (mvdfusion/unet.py:195) + optional DDIM update (mvdfusion/sampler.py:55-65).
 * coef_dev = {a_t, a_prev, sqrt(1-a_t), sigma_t, add_noise, cfg_scale} in device memory. */
int mvd_cfg_ddim(const float* head, int32_t ld, int32_t two_branch, const float* coef_dev, const float* xt,
                 const float* noise, float* eps_out, float* x_prev, float* x0_out, int32_t n_views, int32_t hw,
                 void* stream);
/* out[0:row_len] = table[*idx_dev * row_len + ...]: per-step schedule constants (mvdfusion/sampler.py:55-58 indexes
 * them on the host every step) and pre-drawn noise rows, selected by a device-resident step counter so that one
 * captured CUDA graph replays the whole DDIM loop (mvdfusion/sampler.py:119-142); mvd_increment_i32 advances it. */
int mvd_gather_rows_f32(const float* table, long long row_len, const int32_t* idx_dev, float* out, void* stream);
int mvd_increment_i32(int32_t* counter_dev, int32_t delta, void* stream);
int mvd_nchw_to_rows_f32(const float* x, float* y, int32_t n_img, int32_t C, int32_t hw, void* stream);
int mvd_rows_to_nchw_f32(const float* x, float* y, int32_t n_img, int32_t C, int32_t ld, int32_t hw, void* stream);
/* NCHW fp32 -> NHWC fp16 with the channel dim zero-padded to Cpad (UNetModel.forward input, mvdfusion/unet.py:524,544) */
int mvd_nchw_to_nhwc_f16(const float* x, void* y, int32_t n_img, int32_t C, int32_t hw, int32_t Cpad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GridAttn: depth-guided cross-view aggregation (mvdfusion/view_attn_efficient2.py:269-442).
 *   prep    : z-depth samples (:418-432) and z_embedder maps (:434-437).  scal_dev = {sqrt_alphas_cumprod[t], depth_std}
 *             feat_out fp16 [n_views+1, S*S, 256] (last = input view); zdepth_out fp32 [n_views, D, S*S]
 *   tokens  : unproject -> reproject -> bilinear gather -> Plucker/depth harmonics -> fp16 [q_count*S*S*D*n_views, 736]
 *             cams fp32 [n_views+1, 16] = R(9, row-major) T(3) f(2) pp(2); last = input camera (pytorch3d conventions)
 *   view_attention : timm Attention over the view axis; qkv fp16 [P*V, 768] -> out fp16 [P*V, 256]
 *   view_pool      : weight_layer + softmax over V + weighted sum (:83,92,396-397) -> fp16 [P, 256]
 *   frustum_pool   : area pyramid of the frustum features (mvdfusion/unet.py:198-209)
 *   pixel_cross_attn : DualAttnetionBlock.attn2 with D > 1 keys per pixel (mvdfusion/attention.py:56-62)
 * ---------------------------------------------------------------------------------------------- */
int mvd_gridattn_prep(const float* noisy, const float* input_latent, const float* depth_override, const float* depth_eps,
                      const float* scal_dev, const float* Wz, const float* bz, void* feat_out, float* zdepth_out,
                      int32_t n_views, int32_t S, int32_t D, float depth_scale, float depth_shift, void* stream);
int mvd_gridattn_tokens(const void* feat, const float* zdepth, const float* cams, const float* mask, const float* freqs,
                        const float* ndc_grid, void* tokens, int32_t n_views, int32_t S, int32_t D, int32_t q_first,
                        int32_t q_count, void* stream);
int mvd_view_attention_f16(const void* qkv, void* out, int32_t P, int32_t V, int32_t heads, int32_t hd, void* stream);
int mvd_view_pool_f16(const float* x, const float* w, const float* b, void* out, int32_t P, int32_t V, int32_t C,
                      void* stream);
int mvd_frustum_pool_f16(const void* in, void* out, int32_t n_img, int32_t S, int32_t D, int32_t C, int32_t factor,
                         void* stream);
int mvd_pixel_cross_attn_f16(const void* q, const void* kv, void* out, int32_t M, int32_t D, int32_t heads,
                             int32_t dhead, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GridAttn aggregation transformer as ONE kernel (SURVEY.md kernel K10): pre_layer_b (Linear 723 -> 256 + GELU), the adaLN-Zero DiT
 * blocks over the view axis and the view-softmax pooling (mvdfusion/view_attn_efficient2.py:365-410, :45-70 DiTBlock,
 * :83-97 weight_layer / pooling).  A persistent CTA per SM takes 128 token rows (128 / V points x V views, rows = point * V + view),
 * keeps their fp32 residual stream in TMEM through every block — LayerNorm-modulate, the per-head q | k | v projections, the V x V
 * attention, proj, fc1-GELU-fc2 all run out of shared memory / TMEM — and writes only the pooled fp16 [P, 256] features.
 * Replaces 1 + 7 * layers + 1 launches and the fp32 [P*V, 256] stream round trips of the unfused program.
 *
 * Weights are fp16 [out, in] row-major (nn.Linear), biases / vectors fp32:
 *   w_qkv rows in HEAD order: head h -> rows [96h, 96h + 96) = q_h (32) | k_h (32) | v_h (32); b_qkv likewise
 *   w_proj, w_fc2 (and b_proj, b_fc2) carry the adaLN gate of this step: W' = diag(gate) W, b' = gate * b  (mvd_dit_fold_gates),
 *   so that x += gate * (a W^T + b) is a plain accumulation into the resident stream
 *   shift / scale: the adaLN modulate vectors of norm1 (msa) and norm2 (mlp)
 * 128 % V == 0, V <= 32; hidden 256, 8 heads x 32, mlp 512 (the reference's only configuration).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mvd_dit_layer {
  const void* w_qkv;   const float* b_qkv;    /* [768, 256], [768]  (head order) */
  const void* w_proj;  const float* b_proj;   /* [256, 256], [256]  (gate folded) */
  const void* w_fc1;   const float* b_fc1;    /* [512, 256], [512] */
  const void* w_fc2;   const float* b_fc2;    /* [256, 512], [256]  (gate folded) */
  const float* shift_msa; const float* scale_msa; const float* shift_mlp; const float* scale_mlp;  /* [256] each */
} mvd_dit_layer;
typedef struct mvd_dit_args {
  int32_t R, V, layers;          /* token rows (= P * V), views per point, DiT blocks (1..4) */
  int32_t token_k, token_ld;     /* token width (K of pre_layer_b) and row pitch, both multiples of 8 */
  const void* tokens;            /* fp16 [R, token_ld] */
  const void* w_pre; int32_t w_pre_ld; const float* b_pre;   /* fp16 [256, w_pre_ld >= token_k], fp32 [256] */
  mvd_dit_layer layer[4];
  const float* pool_w; const float* pool_b;   /* weight_layer: fp32 [256], [1] */
  void* pooled;                  /* fp16 [R / V, 256] */
  float* x_out;                  /* optional fp32 [R, 256]: the stream after the last block (parity tests); NULL in the product */
  float eps;                     /* LayerNorm eps (1e-6) */
} mvd_dit_args;
int mvd_gridattn_dit_f16(const mvd_dit_args* args, void* stream);
/* W_out[n, :] = fp16(gate[n] * W[n, :]), b_out[n] = gate[n] * b[n] for up to 8 (W, b) pairs in one launch */
typedef struct mvd_fold_job { const void* w; const float* gate; const float* bias; void* w_out; float* b_out; int32_t N, K; } mvd_fold_job;
int mvd_dit_fold_gates(const mvd_fold_job* jobs, int32_t n_jobs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training (ABI 15): fp32 forward passes that keep what their backward needs, and the backward passes, of
 * the normalisation / activation layers between the GEMMs.  In the reference they are ATen kernels recorded by autograd and
 * replayed by loss.backward() (train.py:90-94; mvdfusion/viewfusion_zero_depth_rgb.py:362-392).  Activations are fp32 [rows, C]
 * channels-last rows; every call zeroes and then accumulates the reductions it owns (dgamma, dbeta, ws) on `stream`.
 * ---------------------------------------------------------------------------------------------- */
/* nn.LayerNorm (external/sd1/ldm/modules/attention.py:210-212; timm LayerNorm + adaLN modulate with gamma = 1 + scale, beta = shift,
 * mvdfusion/view_attn_efficient2.py:61-66).  y = (x - mean) * rstd * gamma + beta; stats fp32 [rows, 2] = (mean, rstd).
 * gamma = beta = NULL: no affine part.  C % 4 == 0. */
int mvd_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* stats, int32_t rows, int32_t C,
                          float eps, void* stream);
/* dx [rows, C]; dgamma / dbeta fp32 [C] (NULL together when the layer has no affine part).  gamma may be NULL (= 1). */
int mvd_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* stats, float* dx, float* dgamma,
                          float* dbeta, int32_t rows, int32_t C, void* stream);
/* GroupNorm32 (+ SiLU) on x fp32 [n_img, hw, C], 32 groups (external/sd1/ldm/modules/diffusionmodules/util.py:204-216,
 * openaimodel.py:199-203,224-228; attention.py:242).  stats fp32 [n_img, 32, 2] = (mean, rstd) per (image, group), written by the
 * forward and read by the backward; ws: scratch of n_img * 16384 bytes, 16-byte aligned (ABI 17: fp64 per-chunk group partials, <= 32
 * chunks x 32 groups x 2 per image — two launches each way, no memset / atomics on the statistics).  dgamma / dbeta are zeroed by the
 * backward call (one memset when dbeta == dgamma + C). */
int mvd_groupnorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* stats, void* ws, int32_t n_img,
                          int32_t hw, int32_t C, float eps, int32_t apply_silu, void* stream);
int mvd_groupnorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* beta, const float* stats, float* dx,
                          float* dgamma, float* dbeta, void* ws, int32_t n_img, int32_t hw, int32_t C, int32_t apply_silu, void* stream);
/* mode 1 GELU (exact erf), 2 SiLU: y[i] = act(x[i]) over rows * cols elements; mode 3 GEGLU: x [rows, 2 cols] = (a | gate),
 * y [rows, cols] = a * gelu(gate) (external/sd1/ldm/modules/attention.py:42-44; cols % 4 == 0).
 * Backward: dx has x's shape; GEGLU: dx = (dy * gelu(gate) | dy * a * gelu'(gate)). */
int mvd_act_fwd_f32(const float* x, float* y, long long rows, int32_t cols, int32_t mode, void* stream);
int mvd_act_bwd_f32(const float* dy, const float* x, float* dx, long long rows, int32_t cols, int32_t mode, void* stream);
/* ABI 16: F.grid_sample(fmap, grid, mode="bilinear", padding_mode="border", align_corners=True) of GridAttn's aggregate_features
 * (mvdfusion/view_attn_efficient2.py:303-318) on channels-last maps: fmap fp32 [V, H, W, C], xy fp32 [V, P, 2] = the grid's (x, y) in
 * [-1, 1], out fp32 [V, P, C].  Backward: dfmap [V, H, W, C] is zeroed, then dfmap[v, tap, :] += w_tap * dout[v, p, :] (16-byte
 * reductions).  The grid carries no gradient on this path (the sampled depth is detached, :291-296).  C % 4 == 0. */
int mvd_bilinear_gather_fwd_f32(const float* fmap, const float* xy, float* out, int32_t V, int32_t H, int32_t W, int32_t C, long long P,
                                void* stream);
int mvd_bilinear_gather_bwd_f32(const float* dout, const float* xy, float* dfmap, int32_t V, int32_t H, int32_t W, int32_t C, long long P,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVD_B200_H_ */
