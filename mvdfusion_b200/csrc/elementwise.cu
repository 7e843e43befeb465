// Data-movement / elementwise kernels of the UNet path (HBM/L2-bound; coalesced, vectorised):
//   cast, channel concat, nearest x2 upsample, stride-2 im2col, small-M GEMV, sinusoidal timestep
//   embedding, UNet input assembly, CFG combine + DDIM update.
#ifdef MVD_CPU_EMULATION
// test infrastructure: this file compiled as plain C++ and run on host threads (tests/native/cpu_emul/cuda_on_cpu.h)
#include "cuda_on_cpu.h"
#else
#include "common.h"
#include "ptx.cuh"
#endif

namespace mvd {

__device__ __forceinline__ uint2 pack4(float a, float b, float c, float d) {
  __half2 h0 = __floats2half2_rn(a, b);
  __half2 h1 = __floats2half2_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  return u;
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, size_t n4) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = pack4(v.x, v.y, v.z, v.w);
  }
}

// out[r, 0:C1] = a[r, :], out[r, C1:C1+C2] = b[r, :]   (torch.cat([h, hs.pop()], dim=1), mvdfusion/unet.py:550)
__global__ void concat_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                  int C1, int C2, size_t total4) {
  pdl_trigger();
  pdl_wait();
  const int C = C1 + C2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 4;
    const size_t r = e / C;
    const int c = static_cast<int>(e - r * C);
    float4 v;
    if (c < C1)
      v = *reinterpret_cast<const float4*>(a + r * C1 + c);
    else
      v = *reinterpret_cast<const float4*>(b + r * C2 + (c - C1));
    *reinterpret_cast<float4*>(out + e) = v;
  }
}

// the same concatenation written as the fp16 operand of the ResBlock's 1x1 skip convolution (openaimodel.py:241,273)
__global__ void concat_f16_kernel(const float* __restrict__ a, const float* __restrict__ b, __half* __restrict__ out,
                                  int C1, int C2, size_t total4) {
  pdl_trigger();
  pdl_wait();
  const int C = C1 + C2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 4;
    const size_t r = e / C;
    const int c = static_cast<int>(e - r * C);
    const float4 v = c < C1 ? *reinterpret_cast<const float4*>(a + r * C1 + c) : *reinterpret_cast<const float4*>(b + r * C2 + (c - C1));
    __half2 h0 = __floats2half2_rn(v.x, v.y);
    __half2 h1 = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out + e) = u;
  }
}

// F.interpolate(scale_factor=2, mode='nearest') (openaimodel.py:116): fp32 [n,H,W,C] -> fp16 [n,2H,2W,C]
__global__ void upsample2x_kernel(const float* __restrict__ x, __half* __restrict__ y, int H, int W, int C,
                                  size_t total4) {
  pdl_trigger();
  pdl_wait();
  const int W2 = 2 * W, H2 = 2 * H;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 4;
    const int c = static_cast<int>(e % C);
    size_t pix = e / C;
    const int ox = static_cast<int>(pix % W2);
    pix /= W2;
    const int oy = static_cast<int>(pix % H2);
    const size_t img = pix / H2;
    const float4 v = *reinterpret_cast<const float4*>(x + ((img * H + (oy >> 1)) * W + (ox >> 1)) * C + c);
    *reinterpret_cast<uint2*>(y + e) = pack4(v.x, v.y, v.z, v.w);
  }
}

// im2col for conv3x3 stride 2: fp32 [n,H,W,C] -> fp16 [n*(H/2)*(W/2), 9*C], k = (ky*3+kx)*C + c.
// pad_lo = 1: padding 1 on every side (Downsample.op of the UNet, openaimodel.py:151);
// pad_lo = 0: zero padding on the right / bottom only — the VAE encoder pads (0,1,0,1) by hand in front of a padding-0 conv
//             (external/sd1/ldm/modules/diffusionmodules/model.py:65-76).
__global__ void im2col_s2_kernel(const float* __restrict__ x, __half* __restrict__ y, int H, int W, int C,
                                 size_t total4, int pad_lo) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, K = 9 * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 4;
    const int k = static_cast<int>(e % K);
    size_t row = e / K;
    const int tap = k / C, c = k - tap * C;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int ox = static_cast<int>(row % Wo);
    row /= Wo;
    const int oy = static_cast<int>(row % Ho);
    const size_t img = row / Ho;
    const int iy = 2 * oy - pad_lo + ky, ix = 2 * ox - pad_lo + kx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const float4*>(x + ((img * H + iy) * W + ix) * C + c);
    *reinterpret_cast<uint2*>(y + e) = pack4(v.x, v.y, v.z, v.w);
  }
}

// y[m, n] = act_out( sum_k act_in(x[m, k]) * W[n, k] + b[n] ),  tiny M (t-only MLPs, per-view context vectors).
// One warp per output column n; W is fp16 [N, ldw]; x fp32 [M, ldx]; y fp32 [M, ldy].
__global__ void gemv_kernel(const float* __restrict__ x, int ldx, const __half* __restrict__ W, int ldw,
                            const float* __restrict__ bias, float* __restrict__ y, int ldy, int M, int N, int K,
                            int silu_in, int silu_out) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const __half* w = W + static_cast<size_t>(n) * ldw;
  for (int m0 = 0; m0 < M; m0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 u = *reinterpret_cast<const uint4*>(w + k);
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
      float wf[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hp[j]);
        wf[2 * j] = f.x;
        wf[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (m0 + r < M) {
          const float* xr = x + static_cast<size_t>(m0 + r) * ldx + k;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float xv = (k + j < K) ? xr[j] : 0.f;
            if (silu_in) xv = xv / (1.f + expf(-xv));
            acc[r] = fmaf(xv, wf[j], acc[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && m0 + r < M) {
        if (bias != nullptr) v += bias[n];
        if (silu_out) v = v / (1.f + expf(-v));
        y[static_cast<size_t>(m0 + r) * ldy + n] = v;
      }
    }
  }
}

// Grouped GEMV on one shared input row: y_j[n] = sum_k act_in(x[k]) * W_j[n, k] + b_j[n] for a table of jobs, one launch.
// Used for the 22 ResBlock.emb_layers of a UNet pass (openaimodel.py:218-224,266-270): all of them see the same SiLU(emb).
// jobs: int64 [n_jobs, 5] = {W ptr (fp16 [N, ldw]), bias ptr (fp32 or 0), y ptr (fp32), N | ldw << 32, first global column}.
__global__ void gemv_grouped_kernel(const float* __restrict__ x, int K, int silu_in, const long long* __restrict__ jobs,
                                    int n_jobs, int total_cols) {
  pdl_trigger();
  pdl_wait();
  MVD_DYNAMIC_SHARED(float, gx);  // act_in(x), K floats (zero-padded to a multiple of 8)
  const int K8 = (K + 7) & ~7;
  for (int k = threadIdx.x; k < K8; k += blockDim.x) {
    float v = k < K ? x[k] : 0.f;
    if (silu_in) v = v / (1.f + expf(-v));
    gx[k] = v;
  }
  __syncthreads();
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (col >= total_cols) return;
  int j = 0;
  while (j + 1 < n_jobs && col >= static_cast<int>(jobs[(j + 1) * 5 + 4])) ++j;
  const long long* job = jobs + j * 5;
  const int n = col - static_cast<int>(job[4]);
  const int N = static_cast<int>(job[3] & 0xffffffffll);
  const int ldw = static_cast<int>(job[3] >> 32);
  if (n >= N) return;
  const __half* w = reinterpret_cast<const __half*>(job[0]) + static_cast<size_t>(n) * ldw;
  const float* bias = reinterpret_cast<const float*>(job[1]);
  float* y = reinterpret_cast<float*>(job[2]);
  float acc = 0.f;
  for (int k0 = 0; k0 < K8; k0 += 1024) {  // four independent 16-byte weight loads in flight per lane
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + i * 256 + lane * 8;
      u[i] = k < K8 ? __ldg(reinterpret_cast<const uint4*>(w + k)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + i * 256 + lane * 8;
      if (k < K8) {
        const __half2* hp = reinterpret_cast<const __half2*>(&u[i]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float2 f = __half22float2(hp[jj]);
          acc = fmaf(gx[k + 2 * jj], f.x, acc);
          acc = fmaf(gx[k + 2 * jj + 1], f.y, acc);
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[n] = acc + (bias != nullptr ? bias[n] : 0.f);
}

// timestep_embedding (util.py:152-172 / mvdfusion/embedder.py:114-134): [cos(t*f_i) | sin(t*f_i)],
// f_i = exp(-ln(max_period) * i / half) supplied as a host-built fp32 table (bit-identical to the
// reference's torch.exp).  t is read from device memory (graph-replay friendly).
__global__ void timestep_embed_kernel(const float* __restrict__ t_ptr, const float* __restrict__ freqs,
                                      float* __restrict__ out, int dim) {
  pdl_trigger();
  pdl_wait();
  const int half = dim / 2;
  const float t = *t_ptr;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = freqs[i];
    const float a = t * f;
    out[i] = cosf(a);
    out[half + i] = sinf(a);
  }
  if ((dim & 1) && threadIdx.x == 0) out[dim - 1] = 0.f;
}

// UNet input (mvdfusion/unet.py:153-161,173-186): per image, NHWC fp16 with Cpad channels:
//   ch 0..4  = noisy latent, ch 5..8 = input latent / 0.18215, ch 9 = input depth channel, rest 0.
//   images [n_views, 2*n_views) are the unconditional branch: concat channels zeroed.
//   cond_scale (optional, [n_views]) multiplies the concat channels of a view (condition drop, unet.py:140-151).
__global__ void unet_input_kernel(const float* __restrict__ noisy /*[n,5,hw]*/, const float* __restrict__ cond /*[1 or n,5,hw]*/,
                                  int cond_batched, const float* __restrict__ cond_scale, __half* __restrict__ out,
                                  int n_views, int n_img, int hw, int Cpad, int hilo) {
  pdl_trigger();
  pdl_wait();
  const size_t total = static_cast<size_t>(n_img) * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int img = static_cast<int>(i / hw);
    const int pix = static_cast<int>(i - static_cast<size_t>(img) * hw);
    const int view = img % n_views;
    const bool uncond = img >= n_views;
    __half* o = out + i * Cpad;
    float val[10];
    for (int c = 0; c < 5; ++c) val[c] = noisy[(static_cast<size_t>(view) * 5 + c) * hw + pix];
    const float* cb = cond + (cond_batched ? static_cast<size_t>(view) * 5 * hw : 0);
    for (int c = 0; c < 5; ++c) {
      float v = uncond ? 0.f : cb[static_cast<size_t>(c) * hw + pix];
      if (cond_scale != nullptr) v *= cond_scale[view];
      if (c < 4) v = v / 0.18215f;
      val[5 + c] = v;
    }
    for (int c = 0; c < 10; ++c) {
      const __half h = __float2half_rn(val[c]);
      o[c] = h;
      if (hilo) {
        o[10 + c] = __float2half_rn(val[c] - __half2float(h));
        o[20 + c] = h;
      }
    }
    for (int c = hilo ? 30 : 10; c < Cpad; ++c) o[c] = __float2half_rn(0.f);
  }
}

// eps = s_uc + w (s - s_uc) (mvdfusion/unet.py:195) from the UNet head output [n_img*hw, ld] (rows of
// images [0,n) = conditional, [n,2n) = unconditional), written NCHW [n,5,hw]; when `xt` is given also the
// DDIM update (mvdfusion/sampler.py:55-65):
//   x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t); x_prev = sqrt(a_prev) x0 + sqrt(max(1-a_prev-sigma^2,1e-7)) eps + sigma*noise
// coef (device) = {a_t, a_prev, sqrt_one_minus_at, sigma_t, add_noise(0/1), cfg_scale}
__global__ void cfg_ddim_kernel(const float* __restrict__ head, int ld, int two_branch, const float* __restrict__ coef,
                                const float* __restrict__ xt, const float* __restrict__ noise, float* __restrict__ eps_out,
                                float* __restrict__ x_prev, float* __restrict__ x0_out, int n, int hw) {
  pdl_trigger();
  pdl_wait();
  const size_t total = static_cast<size_t>(n) * 5 * hw;
  const float w = coef[5];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int pix = static_cast<int>(i % hw);
    const int c = static_cast<int>((i / hw) % 5);
    const int view = static_cast<int>(i / (static_cast<size_t>(hw) * 5));
    const float s = head[(static_cast<size_t>(view) * hw + pix) * ld + c];
    float e = s;
    if (two_branch) {
      const float su = head[(static_cast<size_t>(view + n) * hw + pix) * ld + c];
      e = su + w * (s - su);
    }
    if (eps_out != nullptr) eps_out[i] = e;
    if (xt != nullptr) {
      const float a_t = coef[0], a_prev = coef[1], somat = coef[2], sigma = coef[3];
      const float x = xt[i];
      const float x0 = (x - somat * e) / sqrtf(a_t);
      const float dir = sqrtf(fmaxf(1.f - a_prev - sigma * sigma, 1e-7f)) * e;
      float xp = sqrtf(a_prev) * x0 + dir;
      if (coef[4] != 0.f) xp += sigma * noise[i];
      x_prev[i] = xp;
      if (x0_out != nullptr) x0_out[i] = x0;
    }
  }
}

// NCHW fp32 <-> rows x channels fp32 (public per-module entry points only; the fused path never transposes)
__global__ void nchw_to_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int hw, size_t total) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t r = i / C;
    const size_t img = r / hw, pix = r % hw;
    y[i] = x[(img * C + c) * hw + pix];
  }
}
__global__ void rows_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int ld, int hw,
                                    size_t total) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = i % hw;
    const int c = static_cast<int>((i / hw) % C);
    const size_t img = i / (static_cast<size_t>(hw) * C);
    y[i] = x[(img * hw + pix) * ld + c];
  }
}

// out[0:row_len] = table[idx[0] * row_len + ...]: per-step constants / pre-drawn noise selected by a device-resident
// step counter, so that one captured CUDA graph replays every DDIM iteration (mvdfusion/sampler.py:119-142).
__global__ void gather_rows_kernel(const float* __restrict__ table, long long row_len, const int* __restrict__ idx,
                                   float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const float* src = table + static_cast<long long>(*idx) * row_len;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < row_len;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = src[i];
}
__global__ void increment_kernel(int* p, int delta) {
  pdl_trigger();
  pdl_wait(); *p += delta; }

// NCHW fp32 [n, C, hw] -> NHWC fp16 [n, hw, Cpad] (channels >= C zero-filled): input of the stem conv when a caller
// hands UNetModel.forward an already-assembled tensor (mvdfusion/unet.py:524).
__global__ void nchw_to_nhwc_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, int C, int hw, int Cpad,
                                        size_t total) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const size_t r = i / Cpad;
    const size_t img = r / hw, pix = r % hw;
    y[i] = __float2half_rn(c < C ? x[(img * C + c) * hw + pix] : 0.f);
  }
}

static inline int grid_for(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  if (b > 148u * 16u) b = 148u * 16u;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_cast_f32_f16(const float* x, void* y, long long n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || n <= 0 || (n & 3)) return set_error(MVD_EINVAL, "mvd_cast_f32_f16: n must be a positive multiple of 4");
  MVD_LAUNCH((cast_f32_f16_kernel), grid_for(n / 4), 256, 0, stream, x, static_cast<__half*>(y), static_cast<size_t>(n / 4));
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_concat_f32(const float* a, const float* b, float* out, long long rows, int32_t C1, int32_t C2,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !b || !out || rows <= 0 || C1 <= 0 || C2 <= 0 || (C1 & 3) || (C2 & 3))
    return set_error(MVD_EINVAL, "mvd_concat_f32: channel counts must be multiples of 4");
  const size_t total4 = static_cast<size_t>(rows) * (C1 + C2) / 4;
  MVD_LAUNCH((concat_f32_kernel), grid_for(total4), 256, 0, stream, a, b, out, C1, C2, total4);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_concat_f32_f16(const float* a, const float* b, void* out, long long rows, int32_t C1, int32_t C2,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !b || !out || rows <= 0 || C1 <= 0 || C2 <= 0 || (C1 & 3) || (C2 & 3))
    return set_error(MVD_EINVAL, "mvd_concat_f32_f16: channel counts must be multiples of 4");
  const size_t total4 = static_cast<size_t>(rows) * (C1 + C2) / 4;
  MVD_LAUNCH((concat_f16_kernel), grid_for(total4), 256, 0, stream, a, b, static_cast<__half*>(out), C1, C2, total4);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_upsample2x_f32_f16(const float* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return set_error(MVD_EINVAL, "mvd_upsample2x_f32_f16: bad arguments");
  const size_t total4 = static_cast<size_t>(n_img) * H * W * 4 * C / 4;
  MVD_LAUNCH((upsample2x_kernel), grid_for(total4), 256, 0, stream, x, static_cast<__half*>(y), H, W, C, total4);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_im2col_s2_pad_f32_f16(const float* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t pad_lo,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || n_img <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || (C & 3) || pad_lo < 0 || pad_lo > 1)
    return set_error(MVD_EINVAL, "mvd_im2col_s2_pad_f32_f16: bad arguments");
  const size_t total4 = static_cast<size_t>(n_img) * (H / 2) * (W / 2) * 9 * C / 4;
  MVD_LAUNCH((im2col_s2_kernel), grid_for(total4), 256, 0, stream, x, static_cast<__half*>(y), H, W, C, total4, pad_lo);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_im2col_s2_f32_f16(const float* x, void* y, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                     void* stream_) {
  return mvd_im2col_s2_pad_f32_f16(x, y, n_img, H, W, C, 1, stream_);
}

extern "C" int mvd_gemv_f16(const float* x, int32_t ldx, const void* W, int32_t ldw, const float* bias, float* y,
                            int32_t ldy, int32_t M, int32_t N, int32_t K, int32_t silu_in, int32_t silu_out,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !W || !y || M <= 0 || N <= 0 || K <= 0) return set_error(MVD_EINVAL, "mvd_gemv_f16: bad arguments");
  if ((ldw & 7) || ldw < ((K + 7) & ~7)) return set_error(MVD_EALIGN, "mvd_gemv_f16: ldw must be a multiple of 8 and >= K rounded up to 8");
  MVD_LAUNCH((gemv_kernel), (N + 7) / 8, 256, 0, stream, x, ldx, static_cast<const __half*>(W), ldw, bias, y, ldy, M, N, K, silu_in,
                                              silu_out);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_gemv_grouped_f16(const float* x, int32_t K, int32_t silu_in, const void* jobs_dev, int32_t n_jobs,
                                    int32_t total_cols, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !jobs_dev || K <= 0 || K > 8192 || n_jobs <= 0 || total_cols <= 0)
    return set_error(MVD_EINVAL, "mvd_gemv_grouped_f16: bad arguments");
  const int K8 = (K + 7) & ~7;
  MVD_LAUNCH((gemv_grouped_kernel), (total_cols + 7) / 8, 256, K8 * sizeof(float), stream, x, K, silu_in, static_cast<const long long*>(jobs_dev),
                                                                             n_jobs, total_cols);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_timestep_embedding(const float* t_dev, const float* freqs_dev, float* out, int32_t dim,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!t_dev || !freqs_dev || !out || dim < 2) return set_error(MVD_EINVAL, "mvd_timestep_embedding: bad arguments");
  MVD_LAUNCH((timestep_embed_kernel), 1, 256, 0, stream, t_dev, freqs_dev, out, dim);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_unet_input_f16(const float* noisy, const float* cond, int32_t cond_batched, const float* cond_scale,
                                  void* out, int32_t n_views, int32_t n_img, int32_t hw, int32_t Cpad, int32_t hilo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!noisy || !cond || !out || n_views <= 0 || (n_img != n_views && n_img != 2 * n_views) || hw <= 0 || Cpad < (hilo ? 32 : 10) ||
      (Cpad & 7))
    return set_error(MVD_EINVAL, "mvd_unet_input_f16: bad arguments");
  MVD_LAUNCH((unet_input_kernel), grid_for(static_cast<size_t>(n_img) * hw), 256, 0, stream, 
      noisy, cond, cond_batched, cond_scale, static_cast<__half*>(out), n_views, n_img, hw, Cpad, hilo ? 1 : 0);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_cfg_ddim(const float* head, int32_t ld, int32_t two_branch, const float* coef_dev, const float* xt,
                            const float* noise, float* eps_out, float* x_prev, float* x0_out, int32_t n_views,
                            int32_t hw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!head || !coef_dev || n_views <= 0 || hw <= 0 || ld < 5) return set_error(MVD_EINVAL, "mvd_cfg_ddim: bad arguments");
  if (xt != nullptr && (x_prev == nullptr || noise == nullptr)) return set_error(MVD_EINVAL, "mvd_cfg_ddim: x_prev and noise required with xt");
  if (xt == nullptr && eps_out == nullptr) return set_error(MVD_EINVAL, "mvd_cfg_ddim: nothing to write");
  MVD_LAUNCH((cfg_ddim_kernel), grid_for(static_cast<size_t>(n_views) * 5 * hw), 256, 0, stream, 
      head, ld, two_branch, coef_dev, xt, noise, eps_out, x_prev, x0_out, n_views, hw);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_nchw_to_rows_f32(const float* x, float* y, int32_t n_img, int32_t C, int32_t hw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || n_img <= 0 || C <= 0 || hw <= 0) return set_error(MVD_EINVAL, "mvd_nchw_to_rows_f32: bad arguments");
  const size_t total = static_cast<size_t>(n_img) * C * hw;
  MVD_LAUNCH((nchw_to_rows_kernel), grid_for(total), 256, 0, stream, x, y, C, hw, total);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_rows_to_nchw_f32(const float* x, float* y, int32_t n_img, int32_t C, int32_t ld, int32_t hw,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || n_img <= 0 || C <= 0 || hw <= 0 || ld < C) return set_error(MVD_EINVAL, "mvd_rows_to_nchw_f32: bad arguments");
  const size_t total = static_cast<size_t>(n_img) * C * hw;
  MVD_LAUNCH((rows_to_nchw_kernel), grid_for(total), 256, 0, stream, x, y, C, ld, hw, total);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_gather_rows_f32(const float* table, long long row_len, const int32_t* idx_dev, float* out,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!table || !idx_dev || !out || row_len <= 0) return set_error(MVD_EINVAL, "mvd_gather_rows_f32: bad arguments");
  MVD_LAUNCH((gather_rows_kernel), grid_for(static_cast<size_t>(row_len)), 256, 0, stream, table, row_len, idx_dev, out);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_increment_i32(int32_t* p, int32_t delta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p) return set_error(MVD_EINVAL, "mvd_increment_i32: null pointer");
  MVD_LAUNCH((increment_kernel), 1, 1, 0, stream, p, delta);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_nchw_to_nhwc_f16(const float* x, void* y, int32_t n_img, int32_t C, int32_t hw, int32_t Cpad,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || n_img <= 0 || C <= 0 || hw <= 0 || Cpad < C) return set_error(MVD_EINVAL, "mvd_nchw_to_nhwc_f16: bad arguments");
  const size_t total = static_cast<size_t>(n_img) * hw * Cpad;
  MVD_LAUNCH((nchw_to_nhwc_f16_kernel), grid_for(total), 256, 0, stream, x, static_cast<__half*>(y), C, hw, Cpad, total);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}
