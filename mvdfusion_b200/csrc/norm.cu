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
// Thread mapping shared by both passes: a block is `rows` x (C/4) threads; thread (row, cq) owns channels 4cq..4cq+3 and
// walks pixels row, row+rows, ... of its block's pixel range, so every load is a float4 and a warp reads contiguous
// memory.  stats[img][group] = {sum, sumsq} in double (zeroed by the host wrapper first).
__global__ void gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int hw, int C, int cpg,
                                int pix_per_block, int rows) {
  extern __shared__ float gn_sm[];  // [rows][C] sums, then [rows][C] sums of squares
  const int cq4 = C >> 2;
  const int cq = threadIdx.x % cq4;
  const int row = threadIdx.x / cq4;
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(hw, p0 + pix_per_block);
  const float4* base = reinterpret_cast<const float4*>(x + static_cast<size_t>(img) * hw * C) + cq;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  int p = p0 + row;
  for (; p + 3 * rows < p1; p += 4 * rows) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(base + static_cast<size_t>(p + u * rows) * cq4);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
      q[0] = fmaf(v[u].x, v[u].x, q[0]); q[1] = fmaf(v[u].y, v[u].y, q[1]);
      q[2] = fmaf(v[u].z, v[u].z, q[2]); q[3] = fmaf(v[u].w, v[u].w, q[3]);
    }
  }
  for (; p < p1; p += rows) {
    const float4 v = __ldg(base + static_cast<size_t>(p) * cq4);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
  }
  float* ssum = gn_sm;
  float* ssq = gn_sm + rows * C;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ssum[row * C + cq * 4 + k] = s[k];
    ssq[row * C + cq * 4 + k] = q[k];
  }
  __syncthreads();
  // 32 groups x 8 lanes: lane l of group g sums channels g*cpg + l, l+8, ... over all rows
  const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
  if (g < 32) {
    double a = 0.0, b = 0.0;
    for (int c = l; c < cpg; c += 8)
      for (int rr = 0; rr < rows; ++rr) {
        a += static_cast<double>(ssum[rr * C + g * cpg + c]);
        b += static_cast<double>(ssq[rr * C + g * cpg + c]);
      }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o, 8);
      b += __shfl_xor_sync(0xffffffffu, b, o, 8);
    }
    if (l == 0) {
      double* dst = stats + (static_cast<size_t>(img) * 32 + g) * 2;
      atomicAdd(dst, a);
      atomicAdd(dst + 1, b);
    }
  }
}

__global__ void gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, __half* __restrict__ y,
                                int hw, int C, int cpg, float eps, int apply_silu, int pix_per_block, int rows) {
  __shared__ float s_mean[32], s_rstd[32];
  const int cq4 = C >> 2;
  const int cq = threadIdx.x % cq4;
  const int row = threadIdx.x / cq4;
  const int img = blockIdx.y;
  if (threadIdx.x < 32) {
    const double inv_cnt = 1.0 / (static_cast<double>(hw) * cpg);
    const double* st = stats + (static_cast<size_t>(img) * 32 + threadIdx.x) * 2;
    const double mean = st[0] * inv_cnt;
    const double var = fmax(st[1] * inv_cnt - mean * mean, 0.0);
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = rsqrtf(static_cast<float>(var) + eps);
  }
  __syncthreads();
  float sc[4], sh[4];  // y = x * sc + sh
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = cq * 4 + k;
    const int g = c / cpg;
    const float ga = __ldg(gamma + c);
    sc[k] = s_rstd[g] * ga;
    sh[k] = __ldg(beta + c) - s_mean[g] * s_rstd[g] * ga;
  }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(hw, p0 + pix_per_block);
  const size_t img_off = static_cast<size_t>(img) * hw * cq4;
  const float4* src = reinterpret_cast<const float4*>(x) + img_off + cq;
  uint2* dst = reinterpret_cast<uint2*>(y) + img_off + cq;
  for (int p = p0 + row; p < p1; p += rows) {
    const float4 v = __ldg(src + static_cast<size_t>(p) * cq4);
    float o[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
    if (apply_silu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = o[k] / (1.f + __expf(-o[k]));
    }
    __half2 h0 = __floats2half2_rn(o[0], o[1]);
    __half2 h1 = __floats2half2_rn(o[2], o[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    dst[static_cast<size_t>(p) * cq4] = u;
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
  if (C > 4096) return set_error(MVD_EINVAL, "mvd_groupnorm_f32_f16: C must be <= 4096");
  MVD_CUDA_CHECK(cudaMemsetAsync(stats_ws, 0, static_cast<size_t>(n_img) * 32 * 2 * sizeof(double), stream));
  const int cq4 = C / 4;
  int rows = 512 / cq4;
  if (rows < 1) rows = 1;
  if (rows > hw) rows = hw;
  if (rows * cq4 < 256) rows = (256 + cq4 - 1) / cq4;  // the group reduction needs 256 threads
  const int threads = rows * cq4;
  // enough blocks to cover the machine (~4 per SM), at least 4 pixels per thread
  int chunks = (592 + n_img - 1) / n_img;
  const int max_chunks = (hw + 4 * rows - 1) / (4 * rows);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int ppb = (hw + chunks - 1) / chunks;
  chunks = (hw + ppb - 1) / ppb;
  const size_t sm = static_cast<size_t>(2) * rows * C * sizeof(float);
  gn_stats_kernel<<<dim3(chunks, n_img), threads, sm, stream>>>(x, static_cast<double*>(stats_ws), hw, C, cpg, ppb, rows);
  count_launch();
  gn_apply_kernel<<<dim3(chunks, n_img), threads, 0, stream>>>(x, static_cast<const double*>(stats_ws), gamma, beta,
                                                             static_cast<__half*>(y), hw, C, cpg, eps, apply_silu, ppb, rows);
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
