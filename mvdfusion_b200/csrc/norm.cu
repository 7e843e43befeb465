// Normalisation kernels (HBM/L2-bound, CUDA cores): GroupNorm(+SiLU), LayerNorm, adaLN modulate.
// Inputs are the fp32 residual stream [rows, C]; outputs are fp16 operands for the tensor-core GEMMs.
//   GroupNorm32 / Normalize : external/sd1/ldm/modules/diffusionmodules/util.py:200-217,
//                             external/sd1/ldm/modules/attention.py:76-77
//   nn.LayerNorm            : external/sd1/ldm/modules/attention.py:211-213, mvdfusion/attention.py:35-37
//   DiT LayerNorm+modulate  : mvdfusion/view_attn_efficient2.py:15-16,51,53,65-66
#include <cstdio>
#include <cstdlib>

#ifdef MVD_CPU_EMULATION
// test infrastructure: this file compiled as plain C++ and run on host threads (tests/native/cpu_emul/cuda_on_cpu.h); the blocks of a
// cluster run together there, the distributed-shared-memory reads become plain loads from the peer block's buffer
#include "cuda_on_cpu.h"
#else
#include "common.h"
#include "ptx.cuh"
#endif

namespace mvd {

// ---------------------------------------------------------------------------- GroupNorm
// One kernel, one pass over HBM/L2, statistics exchanged inside a thread-block cluster.
//   slice   = `gpc` consecutive groups of one image (span = gpc * C/32 channels; a multiple of 8 channels when possible,
//             so that every pixel's part of the slice is whole 32-byte sectors)
//   cluster = the CTAs that share one slice; CTA r of `csplit` owns pixels r, r + csplit, ... (interleaved)
//   thread (row, cq) owns channels 4cq .. 4cq+3 of the CTA's pixels row, row + rows, ...: at most MAXP pixels, which
//             stay in REGISTERS between the statistics and the normalisation (all loads of a thread are issued at once:
//             one memory latency per CTA; several small CTAs per SM overlap each other's phases).
//   per-thread fp32 partial sums -> smem [row][span] -> one warp per group sums them in fp64 -> cluster barrier ->
//   every CTA adds up its peers' group sums through distributed shared memory -> normalise, gamma / beta (+SiLU), fp16.
// NT = threads per CTA: 256, or 320 for the one geometry where 256 leaves a second partial wave (16 x 32^2 x 320: see the host code)
template <int MAXP, int NT = 256>
__global__ void __launch_bounds__(NT, MAXP == 16 ? 2 : (MAXP == 4 ? 4 : 3))  // register budget sized to the pixels a thread keeps: occupancy is what hides the latency here
    gn_cluster_kernel(const float* __restrict__ x, const float* __restrict__ x2, int C1, const float* __restrict__ gamma,
                      const float* __restrict__ beta, __half* __restrict__ y, int hw, int C, int cpg, int gpc, int rows, int csplit,
                      float eps, int apply_silu, int ldy, int lo_off) {
  pdl_trigger();
  MVD_DYNAMIC_SHARED_ALIGNED16(uint8_t, gn_smem);
  const int span = gpc * cpg;
  const int span4 = span >> 2;
  double* gsum = reinterpret_cast<double*>(gn_smem);      // [32][2] group sums of this CTA (read by the cluster peers)
  float* s_mean = reinterpret_cast<float*>(gsum + 64);    // [32]
  float* s_rstd = s_mean + 32;                            // [32]
  float* part_s = s_rstd + 32;                            // [rows][span]
  float* part_q = part_s + rows * span;                   // [rows][span]

  const int img = blockIdx.y;
  const int slice = blockIdx.x / csplit;
  const int rank = blockIdx.x - slice * csplit;  // == %cluster_ctarank (cluster dims (csplit, 1, 1))
  const int c0 = slice * span;
  const int cq = threadIdx.x % span4;
  const int row = threadIdx.x / span4;
  const bool active = row < rows;
  // two-source form (x2 != nullptr): channels [0, C1) live in x [.., C1], channels [C1, C) in x2 [.., C - C1] — the
  // torch.cat([h, skip], dim=1) of the UNet's output blocks (mvdfusion/unet.py:550) is never materialised.
  // C1 % 4 == 0, so a thread's four channels always come from one source.
  const int cg = c0 + cq * 4;
  const bool second = x2 != nullptr && cg >= C1;
  const int Csrc = x2 == nullptr ? C : (second ? C - C1 : C1);
  const size_t pix_stride4 = static_cast<size_t>(Csrc) >> 2;
  const float4* src = reinterpret_cast<const float4*>((second ? x2 : x) + static_cast<size_t>(img) * hw * Csrc + (second ? cg - C1 : cg));
  const int pstep = rows * csplit;         // pixel stride of one thread
  const int p_first = row * csplit + rank;  // its first pixel

  // gamma / beta do not depend on the predecessor kernel: they are in flight before the dependency wait and long before the
  // statistics are done (fetched after the cluster barrier they cost one more memory round trip per launch)
  float ga[4], be[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ga[k] = active ? __ldg(gamma + c0 + cq * 4 + k) : 0.f;
    be[k] = active ? __ldg(beta + c0 + cq * 4 + k) : 0.f;
  }
  pdl_wait();
  constexpr int NV = MAXP > 0 ? MAXP : 8;  // MAXP == 0: any number of pixels per thread, eight at a time, re-read for the output
  float4 v[NV];
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    for (int pb = p_first; pb < (MAXP > 0 ? p_first + 1 : hw); pb += NV * pstep) {
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int p = pb + u * pstep;
        v[u] = p < hw ? __ldg(src + static_cast<size_t>(p) * pix_stride4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
        q[0] = fmaf(v[u].x, v[u].x, q[0]); q[1] = fmaf(v[u].y, v[u].y, q[1]);
        q[2] = fmaf(v[u].z, v[u].z, q[2]); q[3] = fmaf(v[u].w, v[u].w, q[3]);
      }
    }
    *reinterpret_cast<float4*>(part_s + row * span + cq * 4) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(part_q + row * span + cq * 4) = make_float4(q[0], q[1], q[2], q[3]);
  }
  __syncthreads();
  {
    // warp w sums group w, w + 8, ...: lane = row, cpg contiguous fp32 partials each, accumulated in double
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < gpc; g += NT / 32) {
      double a = 0.0, b = 0.0;
      for (int rr = lane; rr < rows; rr += 32) {
        const float* ps = part_s + rr * span + g * cpg;
        const float* pq = part_q + rr * span + g * cpg;
        float fa = 0.f, fb = 0.f;  // <= 128 values of one row: fp32 is ample, the cross-row / cross-CTA sums are double
        for (int c = 0; c < cpg; ++c) {
          fa += ps[c];
          fb += pq[c];
        }
        a += static_cast<double>(fa);
        b += static_cast<double>(fb);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) {
        gsum[2 * g] = a;
        gsum[2 * g + 1] = b;
      }
    }
  }
  // publish the group sums to the cluster, then add up every CTA's (own included) in rank order: all CTAs get identical bits
  if (csplit > 1) cluster_sync_all();
  else __syncthreads();
  if (threadIdx.x < gpc) {
    const int g = threadIdx.x;
    double a = 0.0, b = 0.0;
    if (csplit > 1) {
      double ra[8], rb[8];
#ifdef MVD_CPU_EMULATION
      for (int r = 0; r < 8; ++r) {
        ra[r] = rb[r] = 0.0;
        if (r < csplit) {
          const double* remote = cpu_emul::cluster_peer(gsum + 2 * g, r);
          ra[r] = remote[0];
          rb[r] = remote[1];
        }
      }
#else
      const uint32_t local = smem_u32(gsum + 2 * g);
#pragma unroll
      for (int r = 0; r < 8; ++r) {  // all remote loads in flight together
        ra[r] = rb[r] = 0.0;
        if (r < csplit) {
          const uint32_t remote = mapa_u32(local, static_cast<uint32_t>(r));
          asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(ra[r]), "=d"(rb[r]) : "r"(remote) : "memory");
        }
      }
#endif
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        a += ra[r];
        b += rb[r];
      }
    } else {
      a = gsum[2 * g];
      b = gsum[2 * g + 1];
    }
    const double inv_cnt = 1.0 / (static_cast<double>(hw) * cpg);
    const double mean = a * inv_cnt;
    const double var = fmax(b * inv_cnt - mean * mean, 0.0);
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = rsqrtf(static_cast<float>(var) + eps);
  }
  // peers may still be reading this CTA's gsum: arrive now, wait just before exit (the normalisation runs in between)
#ifndef MVD_CPU_EMULATION
  if (csplit > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#endif
  __syncthreads();
  if (active) {
    float sc[4], sh[4];  // y = x * sc + sh
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cl = cq * 4 + k;
      const int g = cl / cpg;
      sc[k] = s_rstd[g] * ga[k];
      sh[k] = be[k] - s_mean[g] * s_rstd[g] * ga[k];
    }
    // ldy = row pitch of y (C, or 2C for the [hi | lo] split-precision form: lo_off > 0 stores fp16(o - fp16(o)) lo_off columns right)
    uint2* dst = reinterpret_cast<uint2*>(y + static_cast<size_t>(img) * hw * ldy + c0) + cq;
    const size_t out_stride4 = static_cast<size_t>(ldy) >> 2;
    for (int pb = p_first; pb < (MAXP > 0 ? p_first + 1 : hw); pb += NV * pstep) {
      if (MAXP == 0) {  // the registers hold the last batch only: read this one again (L2)
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          const int p = pb + u * pstep;
          if (p < hw) v[u] = __ldg(src + static_cast<size_t>(p) * pix_stride4);
        }
      }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int p = pb + u * pstep;
      if (p >= hw) continue;
      float o[4] = {fmaf(v[u].x, sc[0], sh[0]), fmaf(v[u].y, sc[1], sh[1]), fmaf(v[u].z, sc[2], sh[2]), fmaf(v[u].w, sc[3], sh[3])};
      if (apply_silu) {
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = __fdividef(o[k], 1.f + __expf(-o[k]));
      }
      __half2 h0 = __floats2half2_rn(o[0], o[1]);
      __half2 h1 = __floats2half2_rn(o[2], o[3]);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&h0);
      w.y = *reinterpret_cast<uint32_t*>(&h1);
      dst[static_cast<size_t>(p) * out_stride4] = w;
      if (lo_off > 0) {
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        __half2 l0 = __floats2half2_rn(o[0] - f0.x, o[1] - f0.y);
        __half2 l1 = __floats2half2_rn(o[2] - f1.x, o[3] - f1.y);
        uint2 wl;
        wl.x = *reinterpret_cast<uint32_t*>(&l0);
        wl.y = *reinterpret_cast<uint32_t*>(&l1);
        dst[static_cast<size_t>(p) * out_stride4 + (lo_off >> 2)] = wl;
      }
    }
    }
  }
#ifdef MVD_CPU_EMULATION
  if (csplit > 1) cpu_emul::cluster_barrier();  // arrive (above, a no-op here) + wait as one barrier
#else
  if (csplit > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}

// ---------------------------------------------------------------------------- LayerNorm family
// One warp per row, C <= 1280 and a multiple of 4.  mode 0: affine (gamma, beta); mode 1: adaLN
// modulate y = n * (1 + scale[c]) + shift[c] (no affine).
// NV = float4 slots per lane (covers C <= 128 * NV): sized to the row so that narrow rows do not pay for 10 slots of registers
template <int MODE, int RPW, int NV>
__global__ void ln_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                          __half* __restrict__ y, int rows, int C, float eps) {
  pdl_trigger();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5;
  const int row0 = (blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * RPW;  // RPW consecutive rows per warp: all their loads in flight together
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;
  float4 v[RPW][NV];
  float s[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    s[r] = 0.f;
    const float* xr = x + static_cast<size_t>(min(row0 + r, rows - 1)) * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) v[r][i] = *reinterpret_cast<const float4*>(xr + c);
    }
  }
  // wide rows run as a few latency-bound blocks: their gamma / beta loads go out together with the row instead of after the
  // two reductions (narrow rows keep the registers for occupancy: those launches are bandwidth-bound)
  constexpr bool EARLY_AFFINE = NV >= 10;
  float4 ga_[EARLY_AFFINE ? NV : 1], be_[EARLY_AFFINE ? NV : 1];
  if (EARLY_AFFINE) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        ga_[i] = __ldg(reinterpret_cast<const float4*>(a + c));
        be_[i] = __ldg(reinterpret_cast<const float4*>(b + c));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) s[r] += v[r][i].x + v[r][i].y + v[r][i].z + v[r][i].w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  }
  float mean[RPW], q[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    mean[r] = s[r] / C;
    q[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        const float d0 = v[r][i].x - mean[r], d1 = v[r][i].y - mean[r], d2 = v[r][i].z - mean[r], d3 = v[r][i].w - mean[r];
        q[r] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
  }
  float rstd_[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) rstd_[r] = rsqrtf(q[r] / C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      const float4 ga = EARLY_AFFINE ? ga_[i] : *reinterpret_cast<const float4*>(a + c);
      const float4 be = EARLY_AFFINE ? be_[i] : *reinterpret_cast<const float4*>(b + c);
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        if (row0 + r >= rows) continue;
        const float rstd = rstd_[r];
        float o0, o1, o2, o3;
        if (MODE == 0) {
          o0 = (v[r][i].x - mean[r]) * rstd * ga.x + be.x;
          o1 = (v[r][i].y - mean[r]) * rstd * ga.y + be.y;
          o2 = (v[r][i].z - mean[r]) * rstd * ga.z + be.z;
          o3 = (v[r][i].w - mean[r]) * rstd * ga.w + be.w;
        } else {  // a = scale, b = shift
          o0 = (v[r][i].x - mean[r]) * rstd * (1.f + ga.x) + be.x;
          o1 = (v[r][i].y - mean[r]) * rstd * (1.f + ga.y) + be.y;
          o2 = (v[r][i].z - mean[r]) * rstd * (1.f + ga.z) + be.z;
          o3 = (v[r][i].w - mean[r]) * rstd * (1.f + ga.w) + be.w;
        }
        __half2 h0 = __floats2half2_rn(o0, o1);
        __half2 h1 = __floats2half2_rn(o2, o3);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(y + static_cast<size_t>(row0 + r) * C + c) = u;
      }
    }
  }
}

// one row per warp; register slots sized to the row width (C = 320: 6.0 us instead of 10.0 us at 16384 rows, occupancy)
template <int MODE>
static cudaError_t ln_launch(const float* x, const float* a, const float* b, __half* y, int rows, int C, float eps, cudaStream_t stream) {
  static const int force = getenv("MVD_LN_RPW") ? atoi(getenv("MVD_LN_RPW")) : 0;  // measurement override
  const int rpw = force ? force : 1;  // measured on B200: one row per warp wins at every shape of the step (tests/native/norm_bench)
  const dim3 grid2((rows + 15) / 16), grid1((rows + 7) / 8), block(256);
  if (C <= 384) {
    if (rpw == 2) return launch_kernel(ln_kernel<MODE, 2, 3>, grid2, block, 0, stream, 1, x, a, b, y, rows, C, eps);
    return launch_kernel(ln_kernel<MODE, 1, 3>, grid1, block, 0, stream, 1, x, a, b, y, rows, C, eps);
  }
  if (C <= 640) {
    if (rpw == 2) return launch_kernel(ln_kernel<MODE, 2, 5>, grid2, block, 0, stream, 1, x, a, b, y, rows, C, eps);
    return launch_kernel(ln_kernel<MODE, 1, 5>, grid1, block, 0, stream, 1, x, a, b, y, rows, C, eps);
  }
  if (rpw == 2) return launch_kernel(ln_kernel<MODE, 2, 10>, grid2, block, 0, stream, 1, x, a, b, y, rows, C, eps);
  return launch_kernel(ln_kernel<MODE, 1, 10>, grid1, block, 0, stream, 1, x, a, b, y, rows, C, eps);
}

// LayerNorm fp32 -> fp32 with row pitches: one warp per row, any C % 4 == 0 (two passes over the row, L1 / L2 resident).  For the few
// places where the normalised values stay on the fp32 residual stream or feed a GEMV (CLIP's ln_pre / ln_post on strided rows).
__global__ void ln_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ g, const float* __restrict__ b,
                              float* __restrict__ y, long long ldy, int rows, int C, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
  const int n4 = C >> 2;
  float s = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  float4* yr = reinterpret_cast<float4*>(y + static_cast<size_t>(row) * ldy);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    const float4 ga = __ldg(reinterpret_cast<const float4*>(g) + i), be = __ldg(reinterpret_cast<const float4*>(b) + i);
    yr[i] = make_float4((v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y, (v.z - mean) * rstd * ga.z + be.z,
                        (v.w - mean) * rstd * ga.w + be.w);
  }
}

// ---------------------------------------------------------------------------- row softmax
// p[r, :] = softmax(scale * s[r, :]) as fp16 (the P operand of the P V product).  One warp per row; the row is read three
// times (maximum, sum, output) — it is L2-resident, and this runs once per scene, in the VAE decoder's single-head 512-wide
// attention block (external/sd1/ldm/modules/diffusionmodules/model.py:186-190), not in the denoising loop.
__global__ void softmax_rows_kernel(const float* __restrict__ s, __half* __restrict__ p, int rows, int cols, int ld_in, int ld_out,
                                    float scale_log2) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(s + static_cast<size_t>(row) * ld_in);
  const int n4 = cols >> 2;
  float m = -INFINITY;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = src[i];
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float mc = m * scale_log2;
  float l = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = src[i];
    l += exp2f(fmaf(v.x, scale_log2, -mc)) + exp2f(fmaf(v.y, scale_log2, -mc)) + exp2f(fmaf(v.z, scale_log2, -mc)) +
         exp2f(fmaf(v.w, scale_log2, -mc));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float inv = 1.f / l;
  uint2* dst = reinterpret_cast<uint2*>(p + static_cast<size_t>(row) * ld_out);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = src[i];
    __half2 h0 = __floats2half2_rn(exp2f(fmaf(v.x, scale_log2, -mc)) * inv, exp2f(fmaf(v.y, scale_log2, -mc)) * inv);
    __half2 h1 = __floats2half2_rn(exp2f(fmaf(v.z, scale_log2, -mc)) * inv, exp2f(fmaf(v.w, scale_log2, -mc)) * inv);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    dst[i] = u;
  }
}

}  // namespace mvd

using namespace mvd;

static int groupnorm_launch(const float* x, const float* x2, int C1, const float* gamma, const float* beta, void* y, int32_t n_img,
                            int32_t hw, int32_t C, float eps, int32_t apply_silu, cudaStream_t stream, int hilo = 0) {
  const int ldy = hilo ? 2 * C : C, lo_off = hilo ? C : 0;
  if (!x || !gamma || !beta || !y) return set_error(MVD_EINVAL, "mvd_groupnorm_f32_f16: null pointer");
  if (x2 != nullptr && (C1 <= 0 || C1 >= C || (C1 & 3) != 0 || ((C - C1) & 3) != 0))
    return set_error(MVD_EINVAL, "mvd_groupnorm2_f32_f16: C1 and C2 must be positive multiples of 4");
  if (n_img <= 0 || hw <= 0 || C <= 0 || (C % 32) != 0 || (C & 3) != 0)
    return set_error(MVD_EINVAL, "mvd_groupnorm_f32_f16: C must be a multiple of 32");
  const int cpg = C / 32;
  if (C > 4096) return set_error(MVD_EINVAL, "mvd_groupnorm_f32_f16: C must be <= 4096");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(x2) & 15) || (reinterpret_cast<uintptr_t>(y) & 7))
    return set_error(MVD_EALIGN, "mvd_groupnorm_f32_f16: x must be 16-byte and y 8-byte aligned");
  // Geometry.  Candidates: groups per CTA = 1, 2, 4, ... with a span of whole float4s (span % 4 == 0) and at most 256
  // channel quads; pixels of a slice are split over a cluster of 1, 2, 4 or 8 CTAs until a thread keeps <= 16 pixels
  // (the cluster exchange costs more than a longer thread: measured on B200 with MVD_GN_GEOMETRY, tests/native/norm_bench).
  // Preferred: whole 32-byte sectors per pixel (span % 8 == 0) and <= 16 pixels per thread; the first candidate that meets
  // both wins, else the one with the fewest pixels per thread.
  int best_gpc = 0, best_split = 1, best_pp = 1 << 30, best_rows = 0;
  bool best_pref = false;
  for (int gpc = 1; gpc <= 32; gpc <<= 1) {
    const int span = gpc * cpg;
    if ((span & 3) != 0 || span / 4 > 256) continue;
    int rows = 256 / (span / 4);
    if (rows > hw) rows = hw;
    int split = 1;
    while (split < 8 && (hw + rows * split - 1) / (rows * split) > 16) split <<= 1;
    const int pp = (hw + rows * split - 1) / (rows * split);
    const bool pref = (span & 7) == 0 && pp <= 16;
    if (best_gpc == 0 || (pref && !best_pref) || (pref == best_pref && pp < best_pp)) {
      best_gpc = gpc; best_split = split; best_pp = pp; best_rows = rows; best_pref = pref;
    }
    if (pref) break;
  }
  if (best_gpc == 0) return set_error(MVD_EINVAL, "mvd_groupnorm_f32_f16: unsupported shape (hw %d, C %d)", hw, C);
  if (!best_pref && best_pp > 16) {
    // large images (the VAE decoder's 128^2 / 256^2 maps) stream through whatever the slice: take the WIDEST slice of whole
    // sectors that still gives a couple of CTAs per SM, so that a pixel contributes a long contiguous run instead of 16 bytes
    for (int gpc = 32; gpc >= 1; gpc >>= 1) {
      const int span = gpc * cpg;
      if ((span & 7) != 0 || span / 4 > 256) continue;
      if ((32 / gpc) * 8 * n_img >= 256 || gpc == 1) {
        best_gpc = gpc;
        best_split = 8;
        best_rows = 256 / (span / 4);
        if (best_rows > hw) best_rows = hw;
        best_pp = (hw + best_rows * 8 - 1) / (best_rows * 8);
        break;
      }
    }
  }
  if (const char* ov = getenv("MVD_GN_GEOMETRY")) {  // "gpc,split": measurement override (tests/native/norm_bench)
    int og = 0, os = 0;
    if (sscanf(ov, "%d,%d", &og, &os) == 2 && og >= 1 && og <= 32 && 32 % og == 0 && ((og * cpg) & 3) == 0 && og * cpg / 4 <= 256 &&
        (os == 1 || os == 2 || os == 4 || os == 8)) {
      best_gpc = og;
      best_split = os;
      best_rows = 256 / (og * cpg / 4);
      if (best_rows > hw) best_rows = hw;
      best_pp = (hw + best_rows * os - 1) / (best_rows * os);
    }
  }
  // One more candidate: when the chosen geometry does not fit the machine in one wave (16 x 32^2 x 320: 512 CTAs of 256 threads
  // on 3 x 148 slots = a full wave and a 15 % one, 14.8 us), 320-thread CTAs that keep 16 pixels per thread halve the cluster split:
  // 256 CTAs on 2 x 148 slots, one wave.  MVD_GN_NT320=0 keeps the 256-thread form (A/B measurements).
  bool nt320 = false;
  {
    static const bool off = getenv("MVD_GN_NT320") != nullptr && atoi(getenv("MVD_GN_NT320")) == 0;
    const long long ctas = static_cast<long long>(32 / best_gpc) * best_split * n_img;
    const int span4 = best_gpc * cpg / 4;
    if (!off && getenv("MVD_GN_GEOMETRY") == nullptr && best_pp <= 16 && best_split >= 2 && ctas > 3 * 148 && (320 % span4) == 0) {
      const int rows320 = 320 / span4, split2 = best_split / 2;
      const int pp2 = (hw + rows320 * split2 - 1) / (rows320 * split2);
      if (rows320 <= hw && pp2 <= 16 && static_cast<long long>(32 / best_gpc) * split2 * n_img <= 2 * 148) {
        nt320 = true;
        best_split = split2;
        best_rows = rows320;
        best_pp = pp2;
      }
    }
  }
  const int gpc = best_gpc, csplit = best_split, rows = best_rows, span = gpc * cpg;
  const size_t sm = 64 * sizeof(double) + 64 * sizeof(float) + static_cast<size_t>(2) * rows * span * sizeof(float);
  const dim3 grid((32 / gpc) * csplit, n_img);
  __half* yh = static_cast<__half*>(y);
  if (nt320)
    MVD_CUDA_CHECK(launch_kernel(gn_cluster_kernel<16, 320>, grid, dim3(320), sm, stream, csplit, x, x2, C1, gamma, beta, yh, hw, C, cpg, gpc, rows, csplit, eps, apply_silu, ldy, lo_off));
  else if (best_pp <= 4)
    MVD_CUDA_CHECK(launch_kernel(gn_cluster_kernel<4>, grid, dim3(256), sm, stream, csplit, x, x2, C1, gamma, beta, yh, hw, C, cpg, gpc, rows, csplit, eps, apply_silu, ldy, lo_off));
  else if (best_pp <= 8)
    MVD_CUDA_CHECK(launch_kernel(gn_cluster_kernel<8>, grid, dim3(256), sm, stream, csplit, x, x2, C1, gamma, beta, yh, hw, C, cpg, gpc, rows, csplit, eps, apply_silu, ldy, lo_off));
  else if (best_pp <= 12)
    MVD_CUDA_CHECK(launch_kernel(gn_cluster_kernel<12>, grid, dim3(256), sm, stream, csplit, x, x2, C1, gamma, beta, yh, hw, C, cpg, gpc, rows, csplit, eps, apply_silu, ldy, lo_off));
  else if (best_pp > 16)  // large images (64x64 latents and up): pixels stream through in batches and are read twice
    MVD_CUDA_CHECK(launch_kernel(gn_cluster_kernel<0>, grid, dim3(256), sm, stream, csplit, x, x2, C1, gamma, beta, yh, hw, C, cpg, gpc, rows, csplit, eps, apply_silu, ldy, lo_off));
  else
    MVD_CUDA_CHECK(launch_kernel(gn_cluster_kernel<16>, grid, dim3(256), sm, stream, csplit, x, x2, C1, gamma, beta, yh, hw, C, cpg, gpc, rows, csplit, eps, apply_silu, ldy, lo_off));
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_groupnorm_f32_f16(const float* x, const float* gamma, const float* beta, void* y, void* stats_ws,
                                     int32_t n_img, int32_t hw, int32_t C, float eps, int32_t apply_silu,
                                     void* stream_) {
  (void)stats_ws;  // the single-pass kernel keeps its statistics on chip; the argument stays for ABI stability
  return groupnorm_launch(x, nullptr, 0, gamma, beta, y, n_img, hw, C, eps, apply_silu, static_cast<cudaStream_t>(stream_));
}

extern "C" int mvd_groupnorm_hilo_f32_f16(const float* x, const float* gamma, const float* beta, void* y, int32_t n_img, int32_t hw,
                                          int32_t C, float eps, int32_t apply_silu, void* stream_) {
  return groupnorm_launch(x, nullptr, 0, gamma, beta, y, n_img, hw, C, eps, apply_silu, static_cast<cudaStream_t>(stream_), 1);
}

extern "C" int mvd_groupnorm2_f32_f16(const float* x1, int32_t C1, const float* x2, int32_t C2, const float* gamma,
                                      const float* beta, void* y, int32_t n_img, int32_t hw, float eps, int32_t apply_silu,
                                      void* stream_) {
  if (!x2) return set_error(MVD_EINVAL, "mvd_groupnorm2_f32_f16: null pointer");
  return groupnorm_launch(x1, x2, C1, gamma, beta, y, n_img, hw, C1 + C2, eps, apply_silu, static_cast<cudaStream_t>(stream_));
}

extern "C" int mvd_layernorm_f32_f16(const float* x, const float* gamma, const float* beta, void* y, int32_t rows,
                                     int32_t C, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y) return set_error(MVD_EINVAL, "mvd_layernorm_f32_f16: null pointer");
  if (rows <= 0 || C <= 0 || (C & 3) != 0 || C > 1280) return set_error(MVD_EINVAL, "mvd_layernorm_f32_f16: C must be a multiple of 4, <= 1280");
  MVD_CUDA_CHECK(ln_launch<0>(x, gamma, beta, static_cast<__half*>(y), rows, C, eps, stream));
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_layernorm_f32_f32(const float* x, long long ldx, const float* gamma, const float* beta, float* y, long long ldy,
                                     int32_t rows, int32_t C, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !beta || !y) return set_error(MVD_EINVAL, "mvd_layernorm_f32_f32: null pointer");
  if (rows <= 0 || C <= 0 || (C & 3) != 0 || ldx < C || ldy < C || (ldx & 3) != 0 || (ldy & 3) != 0)
    return set_error(MVD_EINVAL, "mvd_layernorm_f32_f32: C and the row pitches must be multiples of 4, pitches >= C");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(gamma) & 15) ||
      (reinterpret_cast<uintptr_t>(beta) & 15))
    return set_error(MVD_EALIGN, "mvd_layernorm_f32_f32: pointers must be 16-byte aligned");
  MVD_LAUNCH(ln_f32_kernel, (rows + 7) / 8, 256, 0, stream, x, ldx, gamma, beta, y, ldy, rows, C, eps);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_ln_modulate_f32_f16(const float* x, const float* shift, const float* scale, void* y, int32_t rows,
                                       int32_t C, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !shift || !scale || !y) return set_error(MVD_EINVAL, "mvd_ln_modulate_f32_f16: null pointer");
  if (rows <= 0 || C <= 0 || (C & 3) != 0 || C > 1280) return set_error(MVD_EINVAL, "mvd_ln_modulate_f32_f16: C must be a multiple of 4, <= 1280");
  MVD_CUDA_CHECK(ln_launch<1>(x, scale, shift, static_cast<__half*>(y), rows, C, eps, stream));
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_softmax_rows_f32_f16(const float* s, void* p, int32_t rows, int32_t cols, int32_t ld_in, int32_t ld_out, float scale,
                                        void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!s || !p) return set_error(MVD_EINVAL, "mvd_softmax_rows_f32_f16: null pointer");
  if (rows <= 0 || cols <= 0 || (cols & 3) != 0 || ld_in < cols || ld_out < cols || (ld_in & 3) != 0 || (ld_out & 3) != 0)
    return set_error(MVD_EINVAL, "mvd_softmax_rows_f32_f16: cols and the leading dimensions must be multiples of 4, ld >= cols");
  if ((reinterpret_cast<uintptr_t>(s) & 15) || (reinterpret_cast<uintptr_t>(p) & 7))
    return set_error(MVD_EALIGN, "mvd_softmax_rows_f32_f16: s must be 16-byte and p 8-byte aligned");
  MVD_LAUNCH(softmax_rows_kernel, (rows + 7) / 8, 256, 0, stream, s, static_cast<__half*>(p), rows, cols, ld_in, ld_out,
             scale * 1.4426950408889634f);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}
