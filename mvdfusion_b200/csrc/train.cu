// Training-side normalisation / activation / gather kernels (second slice): the fp32 forward passes that save what
// their backward needs, and the backward passes, of the non-contraction layers between the GEMMs of mvdfusion_b200/training.py —
// nn.LayerNorm (external/sd1/ldm/modules/attention.py:210-212; timm LayerNorm + adaLN modulate, mvdfusion/view_attn_efficient2.py:61-66),
// GroupNorm32 (+ SiLU) on channels-last rows (external/sd1/ldm/modules/diffusionmodules/util.py:204-216, openaimodel.py:199-203,224-228),
// GELU / SiLU / GEGLU (external/sd1/ldm/modules/attention.py:42-44), and GridAttn's bilinear gather with its scatter-add backward
// (F.grid_sample, mvdfusion/view_attn_efficient2.py:303-318).  In the reference these are ATen kernels recorded by autograd
// (train.py:90-94, loss.backward()).
//
// All of them are HBM-bound streaming passes: activations are fp32 [rows, C] (row = (image*H + y)*W + x), consecutive threads read
// consecutive channels, every per-channel / per-group constant is loaded once per thread and kept in registers over the pixel loop.
// Parameter gradients (dgamma, dbeta) end in fp32 atomics on small zero-initialised vectors; the (sum, sum of squares) of a group,
// whose difference is the variance, is carried in fp64 through per-chunk partials with one writer each.
// The launches are plain stream launches (their neighbours in the training graph are ATen kernels); only the second kernel of a
// GroupNorm pair carries the programmatic-dependent-launch attribute.
#ifdef MVD_CPU_EMULATION
// test infrastructure: the same source compiled as plain C++ and run on host threads (tests/native/cpu_emul/cuda_on_cpu.h), so that
// the kernels' indexing and arithmetic are checked in a container without a GPU; the product is the nvcc build below
#include "cuda_on_cpu.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"
#define MVD_KLAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
// second kernel of a pair: launched with the programmatic-dependent-launch attribute (common.h) so that its launch overlaps the first
// one's tail; it calls pdl_wait() before touching memory
#define MVD_KLAUNCH_PDL(kernel, grid, block, stream, ...) \
  MVD_CUDA_CHECK(::mvd::launch_kernel(kernel, dim3(grid), dim3(block), 0, stream, 1, __VA_ARGS__))
#endif

namespace mvd {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float silu_d(float x) {
  const float s = sigmoid_f(x);
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_d(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * expf(-0.5f * x * x);
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// one warp per row; the row is read three times (mean, centred variance, output) out of L1 / L2
__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ y,
                              float* __restrict__ stats, int rows, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C);
  const int n4 = C >> 2;
  float s = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  if (lane == 0) {
    stats[2 * static_cast<size_t>(row)] = mean;
    stats[2 * static_cast<size_t>(row) + 1] = rstd;
  }
  float4* yr = reinterpret_cast<float4*>(y + static_cast<size_t>(row) * C);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g != nullptr) {
      ga = __ldg(reinterpret_cast<const float4*>(g) + i);
      be = __ldg(reinterpret_cast<const float4*>(b) + i);
    }
    yr[i] = make_float4((v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y, (v.z - mean) * rstd * ga.z + be.z,
                        (v.w - mean) * rstd * ga.w + be.w);
  }
}

// dx = rstd * (g - mean_C(g) - xhat * mean_C(g * xhat)),  g = dy * gamma,  xhat = (x - mean) * rstd
__global__ void ln_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ g,
                                 const float* __restrict__ stats, float* __restrict__ dx, int rows, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float mean = stats[2 * static_cast<size_t>(row)], rstd = stats[2 * static_cast<size_t>(row) + 1];
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C);
  const float4* dr = reinterpret_cast<const float4*>(dy + static_cast<size_t>(row) * C);
  const int n4 = C >> 2;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i], d = dr[i];
    const float4 ga = g != nullptr ? __ldg(reinterpret_cast<const float4*>(g) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float g0 = d.x * ga.x, g1 = d.y * ga.y, g2 = d.z * ga.z, g3 = d.w * ga.w;
    s1 += (g0 + g1) + (g2 + g3);
    s2 += g0 * ((v.x - mean) * rstd) + g1 * ((v.y - mean) * rstd) + g2 * ((v.z - mean) * rstd) + g3 * ((v.w - mean) * rstd);
  }
  s1 = warp_sum(s1) / C;
  s2 = warp_sum(s2) / C;
  float4* or_ = reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * C);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i], d = dr[i];
    const float4 ga = g != nullptr ? __ldg(reinterpret_cast<const float4*>(g) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    or_[i] = make_float4(rstd * (d.x * ga.x - s1 - (v.x - mean) * rstd * s2), rstd * (d.y * ga.y - s1 - (v.y - mean) * rstd * s2),
                         rstd * (d.z * ga.z - s1 - (v.z - mean) * rstd * s2), rstd * (d.w * ga.w - s1 - (v.w - mean) * rstd * s2));
  }
}

// dgamma[c] += sum_r dy[r, c] * xhat[r, c], dbeta[c] += sum_r dy[r, c] over one chunk of rows; block (32 columns, 8 rows in flight)
__global__ void ln_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int C, int rows_per_chunk) {
  __shared__ float sa[8][33], sb[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int r = r0 + ty; r < r1; r += 8) {
      const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + r);
      const size_t o = static_cast<size_t>(r) * C + c;
      const float d = dy[o];
      a += d;
      b += d * ((x[o] - st.x) * st.y);
    }
  }
  sa[ty][tx] = a;
  sb[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      a += sa[j][tx];
      b += sb[j][tx];
    }
    atomicAdd(dbeta + c, a);
    atomicAdd(dgamma + c, b);
  }
}

// ---------------------------------------------------------------------------------------------- GroupNorm (32 groups, channels-last)
// Two launches each way, no memset, no global atomics on the statistics.  grid = (pixel chunks, channel blocks, images); a channel
// block covers `gpb` WHOLE groups (gpb * cpg <= 256 channels), thread = one channel, loop over the chunk's pixels with independent
// loads in flight.  Pass 1 leaves the chunk's per-group partial sums as fp64 pairs in ws[image][chunk][group] (exactly one writer per
// slot); pass 2 sums the <= 32 chunk partials of its group (broadcast loads), forms the statistics, and streams the pixels again.
struct GnDims {
  int hw, C, cpg, gpb, px, chunks;
};

struct GnThread {
  int c, g, lg, p0, p1;
  bool valid;
};
__device__ __forceinline__ GnThread gn_thread(const GnDims& d) {
  GnThread t;
  const int tid = threadIdx.x;
  t.lg = tid / d.cpg;
  t.g = blockIdx.y * d.gpb + t.lg;
  t.c = blockIdx.y * d.gpb * d.cpg + tid;
  t.valid = tid < d.gpb * d.cpg && t.g < 32;
  t.p0 = blockIdx.x * d.px;
  t.p1 = min(d.hw, t.p0 + d.px);
  return t;
}

// the block's per-group (u, v) pairs -> ws[image][chunk][group]: every thread parks its pair in shared memory, one warp per group
// sums the group's cpg entries (shuffle reduction; no atomics), lane 0 stores the fp64 pair
__device__ __forceinline__ void gn_store_partials(float (*sh)[256], const GnThread& t, const GnDims& d, float u, float v, double* ws) {
  const int tid = threadIdx.x;
  sh[0][tid] = t.valid ? u : 0.f;
  sh[1][tid] = t.valid ? v : 0.f;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int lg = warp; lg < d.gpb; lg += nwarps) {  // uniform per warp
    float a = 0.f, b = 0.f;
    for (int j = lane; j < d.cpg; j += 32) {
      a += sh[0][lg * d.cpg + j];
      b += sh[1][lg * d.cpg + j];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    const int g = blockIdx.y * d.gpb + lg;
    if (lane == 0 && g < 32) {
      double* w = ws + ((static_cast<size_t>(blockIdx.z) * d.chunks + blockIdx.x) * 32 + g) * 2;
      w[0] = static_cast<double>(a);
      w[1] = static_cast<double>(b);
    }
  }
}

// sum of the group's chunk partials
__device__ __forceinline__ void gn_sum_partials(const double* __restrict__ ws, const GnDims& d, int g, double& u, double& v) {
  const double* w = ws + (static_cast<size_t>(blockIdx.z) * d.chunks * 32 + g) * 2;
  u = 0.0;
  v = 0.0;
  for (int ch = 0; ch < d.chunks; ++ch, w += 64) {
    u += w[0];
    v += w[1];
  }
}

__global__ void gn_partials_kernel(const float* __restrict__ x, double* __restrict__ ws, GnDims d) {
  __shared__ float sh[2][256];
  pdl_trigger();  // the apply pass may be launched under this one's tail
  const GnThread t = gn_thread(d);
  float s = 0.f, q = 0.f;
  if (t.valid) {
    const float* xp = x + (static_cast<size_t>(blockIdx.z) * d.hw + t.p0) * d.C + t.c;
#pragma unroll 8
    for (int p = t.p0; p < t.p1; ++p, xp += d.C) {
      const float v = *xp;
      s += v;
      q += v * v;
    }
  }
  gn_store_partials(sh, t, d, s, q, ws);
}

__global__ void gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const double* __restrict__ ws, float* __restrict__ stats, float* __restrict__ y, GnDims d, double eps,
                                int silu) {
  pdl_wait();  // launched with the PDL attribute behind gn_partials_kernel
  const GnThread t = gn_thread(d);
  if (!t.valid) return;
  double su, sq;
  gn_sum_partials(ws, d, t.g, su, sq);
  const double m = static_cast<double>(d.hw) * d.cpg;
  const double mean_d = su / m;
  double var = sq / m - mean_d * mean_d;
  var = var > 0.0 ? var : 0.0;
  const float mean = static_cast<float>(mean_d), rstd = static_cast<float>(1.0 / sqrt(var + eps));
  if (blockIdx.x == 0 && threadIdx.x == t.lg * d.cpg) {  // the group's first channel thread of the first chunk: (mean, rstd) for the backward
    stats[(static_cast<size_t>(blockIdx.z) * 32 + t.g) * 2] = mean;
    stats[(static_cast<size_t>(blockIdx.z) * 32 + t.g) * 2 + 1] = rstd;
  }
  const float sc = rstd * __ldg(gamma + t.c);
  const float be = __ldg(beta + t.c);
  size_t o = (static_cast<size_t>(blockIdx.z) * d.hw + t.p0) * d.C + t.c;
#pragma unroll 8
  for (int p = t.p0; p < t.p1; ++p, o += d.C) {
    float v = (x[o] - mean) * sc + be;
    if (silu) v = silu_f(v);
    y[o] = v;
  }
}

// per channel: a = sum dz, b = sum dz * xhat over the chunk -> dbeta, dgamma (fp32 atomics on the zeroed vectors) and the group's
// (sum dz * gamma, sum dz * gamma * xhat) partials;  dz = dy * silu'(xhat * gamma + beta) or dy
__global__ void gn_bwd_partials_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ stats, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, double* __restrict__ ws, GnDims d, int silu) {
  __shared__ float sh[2][256];
  pdl_trigger();
  const GnThread t = gn_thread(d);
  float a = 0.f, b = 0.f, ga = 0.f;
  if (t.valid) {
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + blockIdx.z * 32 + t.g);
    ga = __ldg(gamma + t.c);
    const float be = __ldg(beta + t.c);
    size_t o = (static_cast<size_t>(blockIdx.z) * d.hw + t.p0) * d.C + t.c;
#pragma unroll 8
    for (int p = t.p0; p < t.p1; ++p, o += d.C) {
      const float xh = (x[o] - st.x) * st.y;
      float dz = dy[o];
      if (silu) dz *= silu_d(xh * ga + be);
      a += dz;
      b += dz * xh;
    }
    atomicAdd(dbeta + t.c, a);
    atomicAdd(dgamma + t.c, b);
  }
  gn_store_partials(sh, t, d, a * ga, b * ga, ws);
}

// dx = rstd * (dz * gamma - s1 / m - xhat * s2 / m)
__global__ void gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ stats, const double* __restrict__ ws,
                                    float* __restrict__ dx, GnDims d, int silu) {
  pdl_wait();
  const GnThread t = gn_thread(d);
  if (!t.valid) return;
  double s1d, s2d;
  gn_sum_partials(ws, d, t.g, s1d, s2d);
  const double m = static_cast<double>(d.hw) * d.cpg;
  const float s1 = static_cast<float>(s1d / m), s2 = static_cast<float>(s2d / m);
  const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + blockIdx.z * 32 + t.g);
  const float ga = __ldg(gamma + t.c), be = __ldg(beta + t.c);
  size_t o = (static_cast<size_t>(blockIdx.z) * d.hw + t.p0) * d.C + t.c;
#pragma unroll 8
  for (int p = t.p0; p < t.p1; ++p, o += d.C) {
    const float xh = (x[o] - st.x) * st.y;
    float dz = dy[o];
    if (silu) dz *= silu_d(xh * ga + be);
    dx[o] = st.y * (dz * ga - s1 - xh * s2);
  }
}

// ---------------------------------------------------------------------------------------------- activations
enum { ACT_GELU = 1, ACT_SILU = 2, ACT_GEGLU = 3 };

template <int MODE, bool BWD>
__device__ __forceinline__ float act1(float x, float dy) {
  if (MODE == ACT_GELU) return BWD ? dy * gelu_d(x) : gelu_f(x);
  return BWD ? dy * silu_d(x) : silu_f(x);
}

// y[i] = act(x[i]) (BWD: dx[i] = dy[i] * act'(x[i])); n4 float4 elements + a scalar tail
template <int MODE, bool BWD>
__global__ void act_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, long long n, int vec) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n4 = vec ? (n >> 2) : 0;
  for (long long i = t; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BWD) d = reinterpret_cast<const float4*>(dy)[i];
    reinterpret_cast<float4*>(out)[i] =
        make_float4(act1<MODE, BWD>(v.x, d.x), act1<MODE, BWD>(v.y, d.y), act1<MODE, BWD>(v.z, d.z), act1<MODE, BWD>(v.w, d.w));
  }
  for (long long i = (n4 << 2) + t; i < n; i += stride) out[i] = act1<MODE, BWD>(x[i], BWD ? dy[i] : 0.f);
}

// GEGLU: x [rows, 2 cols] = (a | gate); forward y [rows, cols] = a * gelu(gate);
// backward dx [rows, 2 cols] = (dy * gelu(gate) | dy * a * gelu'(gate)).  One CTA per row, float4 columns.
template <bool BWD>
__global__ void geglu_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, int cols) {
  const size_t row = blockIdx.x;
  const float4* a4 = reinterpret_cast<const float4*>(x + row * 2 * cols);
  const float4* g4 = reinterpret_cast<const float4*>(x + row * 2 * cols + cols);
  const int n4 = cols >> 2;
  if (!BWD) {
    float4* y4 = reinterpret_cast<float4*>(out + row * cols);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 a = a4[i], g = g4[i];
      y4[i] = make_float4(a.x * gelu_f(g.x), a.y * gelu_f(g.y), a.z * gelu_f(g.z), a.w * gelu_f(g.w));
    }
  } else {
    const float4* d4 = reinterpret_cast<const float4*>(dy + row * cols);
    float4* da4 = reinterpret_cast<float4*>(out + row * 2 * cols);
    float4* dg4 = reinterpret_cast<float4*>(out + row * 2 * cols + cols);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 a = a4[i], g = g4[i], d = d4[i];
      da4[i] = make_float4(d.x * gelu_f(g.x), d.y * gelu_f(g.y), d.z * gelu_f(g.z), d.w * gelu_f(g.w));
      dg4[i] = make_float4(d.x * a.x * gelu_d(g.x), d.y * a.y * gelu_d(g.y), d.z * a.z * gelu_d(g.z), d.w * a.w * gelu_d(g.w));
    }
  }
}

// ---------------------------------------------------------------------------------------------- bilinear gather (GridAttn)
// F.grid_sample(fmap, grid, mode="bilinear", padding_mode="border", align_corners=True) on a channels-last map: one warp per
// (view, point), lanes over float4 channel groups.  Taps past the border carry zero weight (the clamped coordinate makes tx or ty 0).
struct Taps {
  int idx[4];
  float w[4];
};
__device__ __forceinline__ Taps bilinear_taps_f32(float gx, float gy, int H, int W) {
  float ix = (gx + 1.f) * 0.5f * (W - 1);
  float iy = (gy + 1.f) * 0.5f * (H - 1);
  ix = fminf(fmaxf(ix, 0.f), static_cast<float>(W - 1));
  iy = fminf(fmaxf(iy, 0.f), static_cast<float>(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f);
  const float tx = ix - x0f, ty = iy - y0f;
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  Taps t;
  t.idx[0] = y0 * W + x0; t.w[0] = (1.f - tx) * (1.f - ty);
  t.idx[1] = y0 * W + x1; t.w[1] = tx * (1.f - ty);
  t.idx[2] = y1 * W + x0; t.w[2] = (1.f - tx) * ty;
  t.idx[3] = y1 * W + x1; t.w[3] = tx * ty;
  return t;
}

__device__ __forceinline__ void atomic_add4(float* p, float4 v) {
#ifdef MVD_CPU_EMULATION
  atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
#else
  atomicAdd(reinterpret_cast<float4*>(p), v);  // one 16-byte reduction (sm_90+)
#endif
}

// BWD = false: out[v, p, :] = sum_t w_t fmap[v, idx_t, :];  BWD = true: dfmap[v, idx_t, :] += w_t dout[v, p, :]
template <bool BWD>
__global__ void gather_kernel(const float* __restrict__ src, const float* __restrict__ xy, float* __restrict__ dst, int H, int W, int C,
                              long long P, long long total) {
  const long long pt = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pt >= total) return;
  const long long v = pt / P;
  const float2 g = __ldg(reinterpret_cast<const float2*>(xy) + pt);
  const Taps t = bilinear_taps_f32(g.x, g.y, H, W);
  const size_t map0 = static_cast<size_t>(v) * H * W * C;
  const int n4 = C >> 2;
  if (!BWD) {
    float4* o = reinterpret_cast<float4*>(dst + static_cast<size_t>(pt) * C);
    for (int i = lane; i < n4; i += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(src + map0 + static_cast<size_t>(t.idx[k]) * C) + i);
        acc.x = fmaf(t.w[k], f.x, acc.x);
        acc.y = fmaf(t.w[k], f.y, acc.y);
        acc.z = fmaf(t.w[k], f.z, acc.z);
        acc.w = fmaf(t.w[k], f.w, acc.w);
      }
      o[i] = acc;
    }
  } else {
    const float4* d = reinterpret_cast<const float4*>(src + static_cast<size_t>(pt) * C);
    for (int i = lane; i < n4; i += 32) {
      const float4 f = d[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t.w[k] == 0.f) continue;
        atomic_add4(dst + map0 + static_cast<size_t>(t.idx[k]) * C + 4 * i, make_float4(t.w[k] * f.x, t.w[k] * f.y, t.w[k] * f.z, t.w[k] * f.w));
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct GnGeometry {
  dim3 grid, block;
  GnDims d;
};
inline GnGeometry gn_geometry(int n_img, int hw, int C) {
  GnGeometry g;
  GnDims& d = g.d;
  d.hw = hw;
  d.C = C;
  d.cpg = C / 32;
  d.gpb = 256 / d.cpg < 1 ? 1 : (256 / d.cpg > 32 ? 32 : 256 / d.cpg);  // whole groups per block, <= 256 channels
  d.px = (hw + 31) / 32 < 8 ? 8 : (hw + 31) / 32;                           // <= 32 chunks per image, >= 8 pixels per thread
  d.chunks = (hw + d.px - 1) / d.px;
  g.block = dim3(((d.gpb * d.cpg + 31) / 32) * 32);
  g.grid = dim3(d.chunks, (32 + d.gpb - 1) / d.gpb, n_img);
  return g;
}

template <bool BWD>
int act_launch(const char* name, const float* x, const float* dy, float* out, long long rows, int32_t cols, int32_t mode, cudaStream_t stream) {
  if (!x || !out || (BWD && !dy)) return set_error(MVD_EINVAL, "%s: null pointer", name);
  if (rows <= 0 || cols <= 0) return set_error(MVD_EINVAL, "%s: empty input", name);
  if (mode == ACT_GEGLU) {
    if ((cols & 3) != 0 || !aligned16(x) || !aligned16(out) || (BWD && !aligned16(dy)))
      return set_error(MVD_EALIGN, "%s: GEGLU needs cols %% 4 == 0 and 16-byte aligned pointers", name);
    if (rows > 2147483647LL) return set_error(MVD_EINVAL, "%s: too many rows", name);
    MVD_KLAUNCH(geglu_kernel<BWD>, static_cast<unsigned>(rows), 256, stream, x, dy, out, cols);
  } else if (mode == ACT_GELU || mode == ACT_SILU) {
    const long long n = rows * cols;
    const int vec = aligned16(x) && aligned16(out) && (!BWD || aligned16(dy));
    const long long work = vec ? (n + 3) / 4 : n;
    long long blocks = (work + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    auto kernel = mode == ACT_GELU ? act_kernel<ACT_GELU, BWD> : act_kernel<ACT_SILU, BWD>;  // (a template-id with a comma cannot be a macro argument)
    MVD_KLAUNCH(kernel, static_cast<unsigned>(blocks), 256, stream, x, dy, out, n, vec);
  } else {
    return set_error(MVD_EINVAL, "%s: mode must be 1 (GELU), 2 (SiLU) or 3 (GEGLU)", name);
  }
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" int mvd_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* stats, int32_t rows,
                                     int32_t C, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || !stats) return set_error(MVD_EINVAL, "mvd_layernorm_fwd_f32: null pointer");
  if ((gamma == nullptr) != (beta == nullptr)) return set_error(MVD_EINVAL, "mvd_layernorm_fwd_f32: gamma and beta go together");
  if (rows <= 0 || C <= 0 || (C & 3) != 0) return set_error(MVD_EINVAL, "mvd_layernorm_fwd_f32: C must be a multiple of 4");
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta) || (reinterpret_cast<uintptr_t>(stats) & 7))
    return set_error(MVD_EALIGN, "mvd_layernorm_fwd_f32: x / y / gamma / beta must be 16-byte, stats 8-byte aligned");
  MVD_KLAUNCH(ln_fwd_kernel, (rows + 7) / 8, 256, stream, x, gamma, beta, y, stats, rows, C, eps);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* stats, float* dx, float* dgamma,
                                     float* dbeta, int32_t rows, int32_t C, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dy || !x || !stats || !dx) return set_error(MVD_EINVAL, "mvd_layernorm_bwd_f32: null pointer");
  if ((dgamma == nullptr) != (dbeta == nullptr)) return set_error(MVD_EINVAL, "mvd_layernorm_bwd_f32: dgamma and dbeta go together");
  if (rows <= 0 || C <= 0 || (C & 3) != 0) return set_error(MVD_EINVAL, "mvd_layernorm_bwd_f32: C must be a multiple of 4");
  if (!aligned16(dy) || !aligned16(x) || !aligned16(dx) || !aligned16(gamma) || (reinterpret_cast<uintptr_t>(stats) & 7))
    return set_error(MVD_EALIGN, "mvd_layernorm_bwd_f32: dy / x / dx / gamma must be 16-byte, stats 8-byte aligned");
  MVD_KLAUNCH(ln_bwd_dx_kernel, (rows + 7) / 8, 256, stream, dy, x, gamma, stats, dx, rows, C);
  count_launch();
  if (dgamma != nullptr) {
    MVD_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, stream));
    MVD_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, sizeof(float) * C, stream));
    int rpc = (rows + 127) / 128;
    if (rpc < 64) rpc = 64;
    MVD_KLAUNCH(ln_bwd_param_kernel, dim3((C + 31) / 32, (rows + rpc - 1) / rpc), dim3(32, 8), stream, dy, x, stats, dgamma, dbeta, rows, C, rpc);
    count_launch();
  }
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_groupnorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* stats, void* ws,
                                     int32_t n_img, int32_t hw, int32_t C, float eps, int32_t apply_silu, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y || !stats || !ws) return set_error(MVD_EINVAL, "mvd_groupnorm_fwd_f32: null pointer");
  if (n_img <= 0 || hw <= 0 || C <= 0 || (C % 32) != 0 || C > 8192 || n_img > 65535)
    return set_error(MVD_EINVAL, "mvd_groupnorm_fwd_f32: C must be a multiple of 32 (32 groups) and <= 8192, n_img <= 65535");
  if ((reinterpret_cast<uintptr_t>(ws) & 15) || (reinterpret_cast<uintptr_t>(stats) & 7))
    return set_error(MVD_EALIGN, "mvd_groupnorm_fwd_f32: ws must be 16-byte, stats 8-byte aligned");
  const GnGeometry g = gn_geometry(n_img, hw, C);
  double* part = static_cast<double*>(ws);
  MVD_KLAUNCH(gn_partials_kernel, g.grid, g.block, stream, x, part, g.d);
  MVD_KLAUNCH_PDL(gn_apply_kernel, g.grid, g.block, stream, x, gamma, beta, part, stats, y, g.d, static_cast<double>(eps), apply_silu);
  count_launch(2);
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_groupnorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* beta, const float* stats,
                                     float* dx, float* dgamma, float* dbeta, void* ws, int32_t n_img, int32_t hw, int32_t C,
                                     int32_t apply_silu, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dy || !x || !gamma || !beta || !stats || !dx || !dgamma || !dbeta || !ws)
    return set_error(MVD_EINVAL, "mvd_groupnorm_bwd_f32: null pointer");
  if (n_img <= 0 || hw <= 0 || C <= 0 || (C % 32) != 0 || C > 8192 || n_img > 65535)
    return set_error(MVD_EINVAL, "mvd_groupnorm_bwd_f32: C must be a multiple of 32 (32 groups) and <= 8192, n_img <= 65535");
  if ((reinterpret_cast<uintptr_t>(ws) & 15) || (reinterpret_cast<uintptr_t>(stats) & 7))
    return set_error(MVD_EALIGN, "mvd_groupnorm_bwd_f32: ws must be 16-byte, stats 8-byte aligned");
  const GnGeometry g = gn_geometry(n_img, hw, C);
  double* part = static_cast<double*>(ws);
  if (dbeta == dgamma + C) {  // one vector [2, C]: one memset
    MVD_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * 2 * C, stream));
  } else {
    MVD_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, stream));
    MVD_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, sizeof(float) * C, stream));
  }
  MVD_KLAUNCH(gn_bwd_partials_kernel, g.grid, g.block, stream, dy, x, gamma, beta, stats, dgamma, dbeta, part, g.d, apply_silu);
  MVD_KLAUNCH_PDL(gn_bwd_apply_kernel, g.grid, g.block, stream, dy, x, gamma, beta, stats, part, dx, g.d, apply_silu);
  count_launch(2);
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_act_fwd_f32(const float* x, float* y, long long rows, int32_t cols, int32_t mode, void* stream_) {
  return act_launch<false>("mvd_act_fwd_f32", x, nullptr, y, rows, cols, mode, static_cast<cudaStream_t>(stream_));
}

extern "C" int mvd_act_bwd_f32(const float* dy, const float* x, float* dx, long long rows, int32_t cols, int32_t mode, void* stream_) {
  return act_launch<true>("mvd_act_bwd_f32", x, dy, dx, rows, cols, mode, static_cast<cudaStream_t>(stream_));
}

template <bool BWD>
static int gather_launch(const char* name, const float* src, const float* xy, float* dst, int32_t V, int32_t H, int32_t W, int32_t C, long long P,
                         cudaStream_t stream) {
  if (!src || !xy || !dst) return set_error(MVD_EINVAL, "%s: null pointer", name);
  if (V <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) != 0 || P <= 0 || static_cast<long long>(V) * P > (1LL << 33))
    return set_error(MVD_EINVAL, "%s: C must be a multiple of 4, V * P <= 2^33", name);
  if (!aligned16(src) || !aligned16(dst) || (reinterpret_cast<uintptr_t>(xy) & 7))
    return set_error(MVD_EALIGN, "%s: the maps / rows must be 16-byte, xy 8-byte aligned", name);
  const long long total = static_cast<long long>(V) * P;
  if (BWD) MVD_CUDA_CHECK(cudaMemsetAsync(dst, 0, sizeof(float) * static_cast<size_t>(V) * H * W * C, stream));
  MVD_KLAUNCH(gather_kernel<BWD>, static_cast<unsigned>((total + 7) / 8), 256, stream, src, xy, dst, H, W, C, P, total);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_bilinear_gather_fwd_f32(const float* fmap, const float* xy, float* out, int32_t V, int32_t H, int32_t W, int32_t C,
                                           long long P, void* stream_) {
  return gather_launch<false>("mvd_bilinear_gather_fwd_f32", fmap, xy, out, V, H, W, C, P, static_cast<cudaStream_t>(stream_));
}

extern "C" int mvd_bilinear_gather_bwd_f32(const float* dout, const float* xy, float* dfmap, int32_t V, int32_t H, int32_t W, int32_t C,
                                           long long P, void* stream_) {
  return gather_launch<true>("mvd_bilinear_gather_bwd_f32", dout, xy, dfmap, V, H, W, C, P, static_cast<cudaStream_t>(stream_));
}
