// Normalisation kernels (HBM/L2-bound, CUDA cores): GroupNorm(+SiLU), LayerNorm, adaLN modulate.
// Inputs are the fp32 residual stream [rows, C]; outputs are fp16 operands for the tensor-core GEMMs.
//   GroupNorm32 / Normalize : external/sd1/ldm/modules/diffusionmodules/util.py:200-217,
//                             external/sd1/ldm/modules/attention.py:76-77
//   nn.LayerNorm            : external/sd1/ldm/modules/attention.py:211-213, mvdfusion/attention.py:35-37
//   DiT LayerNorm+modulate  : mvdfusion/view_attn_efficient2.py:15-16,51,53,65-66
#include "common.h"
#include "ptx.cuh"

namespace mvd {

// ---------------------------------------------------------------------------- GroupNorm
// stats[img][group] = {sum, sumsq} in double (zeroed by the host wrapper first)
__global__ void gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int hw, int C, int cpg,
                                int pix_per_block) {
  __shared__ double s_sum[32], s_sq[32];
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(hw, p0 + pix_per_block);
  if (threadIdx.x < 32) {
    s_sum[threadIdx.x] = 0.0;
    s_sq[threadIdx.x] = 0.0;
  }
  __syncthreads();
  const float* base = x + (static_cast<size_t>(img) * hw) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float v = __ldg(base + static_cast<size_t>(p) * C + c);
      s += v;
      q = fmaf(v, v, q);
    }
    const int g = c / cpg;
    atomicAdd(&s_sum[g], static_cast<double>(s));
    atomicAdd(&s_sq[g], static_cast<double>(q));
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double* dst = stats + (static_cast<size_t>(img) * 32 + threadIdx.x) * 2;
    atomicAdd(dst, s_sum[threadIdx.x]);
    atomicAdd(dst + 1, s_sq[threadIdx.x]);
  }
}

__global__ void gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, __half* __restrict__ y,
                                int hw, int C, int cpg, float eps, int apply_silu, size_t total4) {
  const double inv_cnt = 1.0 / (static_cast<double>(hw) * cpg);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 4;
    const int c = static_cast<int>(e % C);
    const int img = static_cast<int>(e / (static_cast<size_t>(hw) * C));
    const float4 v = *reinterpret_cast<const float4*>(x + e);
    float in[4] = {v.x, v.y, v.z, v.w};
    float out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = (c + k) / cpg;
      const double* st = stats + (static_cast<size_t>(img) * 32 + g) * 2;
      const double mean = st[0] * inv_cnt;
      const double var = fmax(st[1] * inv_cnt - mean * mean, 0.0);
      const float rstd = rsqrtf(static_cast<float>(var) + eps);
      float t = (in[k] - static_cast<float>(mean)) * rstd * __ldg(gamma + c + k) + __ldg(beta + c + k);
      if (apply_silu) t = t / (1.f + expf(-t));
      out[k] = t;
    }
    __half2 h0 = __floats2half2_rn(out[0], out[1]);
    __half2 h1 = __floats2half2_rn(out[2], out[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(y + e) = u;
  }
}

// ---------------------------------------------------------------------------- LayerNorm family
// One warp per row, C <= 1280 and a multiple of 4.  mode 0: affine (gamma, beta); mode 1: adaLN
// modulate y = n * (1 + scale[c]) + shift[c] (no affine).
template <int MODE>
__global__ void ln_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                          __half* __restrict__ y, int rows, int C, float eps) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * C;
  float4 v[10];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
      q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  __half* yr = y + static_cast<size_t>(row) * C;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      const float4 ga = *reinterpret_cast<const float4*>(a + c);
      const float4 be = *reinterpret_cast<const float4*>(b + c);
      float o0, o1, o2, o3;
      if (MODE == 0) {
        o0 = (v[i].x - mean) * rstd * ga.x + be.x;
        o1 = (v[i].y - mean) * rstd * ga.y + be.y;
        o2 = (v[i].z - mean) * rstd * ga.z + be.z;
        o3 = (v[i].w - mean) * rstd * ga.w + be.w;
      } else {  // a = scale, b = shift
        o0 = (v[i].x - mean) * rstd * (1.f + ga.x) + be.x;
        o1 = (v[i].y - mean) * rstd * (1.f + ga.y) + be.y;
        o2 = (v[i].z - mean) * rstd * (1.f + ga.z) + be.z;
        o3 = (v[i].w - mean) * rstd * (1.f + ga.w) + be.w;
      }
      __half2 h0 = __floats2half2_rn(o0, o1);
      __half2 h1 = __floats2half2_rn(o2, o3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(yr + c) = u;
    }
  }
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_groupnorm_f32_f16(const float* x, const float* gamma, const float* beta, void* y, void* stats_ws,
                                     int32_t n_img, int32_t hw, int32_t C, float eps, int32_t apply_silu,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y || !stats_ws) return set_error(MVD_EINVAL, "mvd_groupnorm_f32_f16: null pointer");
  if (n_img <= 0 || hw <= 0 || C <= 0 || (C % 32) != 0 || (C & 3) != 0)
    return set_error(MVD_EINVAL, "mvd_groupnorm_f32_f16: C must be a multiple of 32");
  const int cpg = C / 32;
  MVD_CUDA_CHECK(cudaMemsetAsync(stats_ws, 0, static_cast<size_t>(n_img) * 32 * 2 * sizeof(double), stream));
  // enough blocks to cover the machine: ~ 148*4 blocks in total
  int chunks = (592 + n_img - 1) / n_img;
  if (chunks > hw) chunks = hw;
  if (chunks < 1) chunks = 1;
  const int ppb = (hw + chunks - 1) / chunks;
  chunks = (hw + ppb - 1) / ppb;
  gn_stats_kernel<<<dim3(chunks, n_img), 256, 0, stream>>>(x, static_cast<double*>(stats_ws), hw, C, cpg, ppb);
  count_launch();
  const size_t total4 = static_cast<size_t>(n_img) * hw * C / 4;
  int blocks = static_cast<int>((total4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_apply_kernel<<<blocks, 256, 0, stream>>>(x, static_cast<const double*>(stats_ws), gamma, beta,
                                              static_cast<__half*>(y), hw, C, cpg, eps, apply_silu, total4);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_layernorm_f32_f16(const float* x, const float* gamma, const float* beta, void* y, int32_t rows,
                                     int32_t C, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y) return set_error(MVD_EINVAL, "mvd_layernorm_f32_f16: null pointer");
  if (rows <= 0 || C <= 0 || (C & 3) != 0 || C > 1280) return set_error(MVD_EINVAL, "mvd_layernorm_f32_f16: C must be a multiple of 4, <= 1280");
  ln_kernel<0><<<(rows + 7) / 8, 256, 0, stream>>>(x, gamma, beta, static_cast<__half*>(y), rows, C, eps);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_ln_modulate_f32_f16(const float* x, const float* shift, const float* scale, void* y, int32_t rows,
                                       int32_t C, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !shift || !scale || !y) return set_error(MVD_EINVAL, "mvd_ln_modulate_f32_f16: null pointer");
  if (rows <= 0 || C <= 0 || (C & 3) != 0 || C > 1280) return set_error(MVD_EINVAL, "mvd_ln_modulate_f32_f16: C must be a multiple of 4, <= 1280");
  ln_kernel<1><<<(rows + 7) / 8, 256, 0, stream>>>(x, scale, shift, static_cast<__half*>(y), rows, C, eps);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}
