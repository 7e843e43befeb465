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
 *   v = acc + bias[n] + rowbias[(m / rows_per_group), n] ; v = act(v) ; v += residual[m, n] ; store
 *   act: NONE, GELU (exact erf), SILU, GEGLU (weights pre-interleaved per tile so that column
 *        j and column j + BN/2 of a tile are value/gate; output has N/2 columns:
 *        out = value * gelu(gate); external/sd1/ldm/modules/attention.py:42-44)
 *   out_mode: F32 / F16 row-major [M, ldc]; QKV_HEADS scatters q,k into [img*heads + h, seq, dpad]
 *        and v transposed into [img*heads + h, dpad, seq] for mvd_attn_self_f16.
 *   split_k > 1: partial sums are red.add'ed into an fp32 output the library zeroes first
 *        (F32 out, act NONE only).
 * ---------------------------------------------------------------------------------------------- */
enum { MVD_A_ROWMAJOR = 0, MVD_A_CONV3X3 = 1 };
enum { MVD_ACT_NONE = 0, MVD_ACT_GELU = 1, MVD_ACT_SILU = 2, MVD_ACT_GEGLU = 3 };
enum { MVD_OUT_F32 = 0, MVD_OUT_F16 = 1, MVD_OUT_QKV_HEADS = 2 };

typedef struct mvd_gemm_args {
  int32_t M, N, K;
  int32_t a_mode;
  const void* A;       /* fp16 */
  int32_t lda;         /* elements; ROWMAJOR only (multiple of 8) */
  int32_t n_img, H, W, C; /* CONV3X3 only; C multiple of 8, W a power of two <= 128 */
  const void* Wt;      /* fp16 [N, ldw] */
  int32_t ldw;         /* elements, multiple of 8, >= K */
  const float* bias;   /* [N] or NULL */
  const float* rowbias;/* [ceil(M/rows_per_group), N] or NULL */
  int32_t rows_per_group;
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
  int32_t split_k;     /* >= 1 */
  int32_t tile_n;      /* 0 = auto; 64, 128 or 256 */
} mvd_gemm_args;

int mvd_gemm_f16(const mvd_gemm_args* args, void* stream);

/* GEGLU weight interleave used by MVD_ACT_GEGLU for a given tile width: row index of the
 * packed weight -> row index of the original nn.Linear(dim, 2*inner) weight. */
int mvd_geglu_row_permutation(int32_t inner_dim, int32_t tile_n, int32_t* perm_out /* [2*inner_dim] */);

#ifdef __cplusplus
}
#endif
#endif /* MVD_B200_H_ */
