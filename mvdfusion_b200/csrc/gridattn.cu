// GridAttn (depth-guided cross-view attention) kernels, mvdfusion/view_attn_efficient2.py:269-442.
//
//   gridattn_prep   : depth de-bias + jitter + unnormalise -> z-depth (:418-432); z_embedder Linear(5,256)+GELU
//                     on the noisy latents of every view and on the input latent (:434-437) -> fp16 NHWC maps
//   gridattn_tokens : per (query point, view): unproject (utils/ray_utils.py:174-202, pytorch3d conventions
//                     restated in SURVEY.md §8c), reproject into the view and into the input view
//                     (:303,321), bilinear gather with border padding on the negated xy (:310-329),
//                     Plücker / depth harmonic encodings (:334-360, utils/common_utils.py:229-244),
//                     concatenated into the 723-d token (:365-370), zero-padded to 736 fp16
//   view_attention  : timm Attention over the V axis (seq = V, 8 heads x 32) for every point
//   view_pool       : weight_layer + softmax over V + weighted sum (:83,92,396-397)
//   frustum_pool    : 2^l x 2^l area pooling of the frustum features (mvdfusion/unet.py:198-209)
//
// These are gather / transcendental / tiny-attention kernels: HBM/L2-bound, no tensor cores.
#ifdef MVD_CPU_EMULATION
// test infrastructure: this file compiled as plain C++ and run on host threads (tests/native/cpu_emul/cuda_on_cpu.h)
#include "cuda_on_cpu.h"
#else
#include "common.h"
#include "ptx.cuh"
#endif

namespace mvd {

constexpr int ZC = 256;         // z_embedder width
constexpr int TOKEN_LD = 736;   // 723 rounded up to a multiple of 16
constexpr int N_HARM = 7;

struct Cam {  // pytorch3d PerspectiveCameras, row-vector convention X_view = X_world R + T
  float R[9];
  float T[3];
  float f[2];
  float pp[2];
};

__device__ __forceinline__ void cam_load(const float* __restrict__ cams, int i, Cam& c) {
  const float* p = cams + i * 16;
#pragma unroll
  for (int k = 0; k < 9; ++k) c.R[k] = p[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) c.T[k] = p[9 + k];
  c.f[0] = p[12];
  c.f[1] = p[13];
  c.pp[0] = p[14];
  c.pp[1] = p[15];
}
// camera centre C = -T R^T
__device__ __forceinline__ void cam_center(const Cam& c, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i) C[i] = -(c.T[0] * c.R[i * 3 + 0] + c.T[1] * c.R[i * 3 + 1] + c.T[2] * c.R[i * 3 + 2]);
}
// NDC projection: X_view = X R + T; x = fx X/Z + px, y = fy Y/Z + py
__device__ __forceinline__ void cam_project(const Cam& c, const float* X, float& x, float& y) {
  float v[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) v[j] = X[0] * c.R[0 * 3 + j] + X[1] * c.R[1 * 3 + j] + X[2] * c.R[2 * 3 + j] + c.T[j];
  x = c.f[0] * v[0] / v[2] + c.pp[0];
  y = c.f[1] * v[1] / v[2] + c.pp[1];
}

// ---------------------------------------------------------------------------------------------- prep
// lat: [n_views, 5, hw] noisy latents; inp: [1, 5, hw] input latent.  feat: fp16 [(n_views+1), hw, 256]
// (index n_views = input view).  zdepth: fp32 [n_views, D, hw].
// scal (device): {sqrt_alphas_cumprod[t], depth_std}
__global__ void gridattn_prep_kernel(const float* __restrict__ lat, const float* __restrict__ inp,
                                     const float* __restrict__ depth_override, const float* __restrict__ eps,
                                     const float* __restrict__ scal, const float* __restrict__ Wz,
                                     const float* __restrict__ bz, __half* __restrict__ feat,
                                     float* __restrict__ zdepth, int n_views, int hw, int D, float depth_scale,
                                     float depth_shift) {
  pdl_trigger();
  pdl_wait();
  const int pix = blockIdx.x;
  const int view = blockIdx.y;  // n_views == input view
  const float* src = (view < n_views) ? lat + static_cast<size_t>(view) * 5 * hw : inp;
  float x[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) x[c] = src[static_cast<size_t>(c) * hw + pix];
  for (int o = threadIdx.x; o < ZC; o += blockDim.x) {
    float a = bz[o];
#pragma unroll
    for (int c = 0; c < 5; ++c) a = fmaf(x[c], Wz[o * 5 + c], a);
    feat[(static_cast<size_t>(view) * hw + pix) * ZC + o] = __float2half_rn(gelu_erf(a));
  }
  if (view < n_views && threadIdx.x < D) {
    const int d = threadIdx.x;
    const float mean = (depth_override != nullptr) ? depth_override[static_cast<size_t>(view) * hw + pix] : x[4] / scal[0];
    const float s = mean + scal[1] * eps[(static_cast<size_t>(view) * D + d) * hw + pix];
    const float u = fminf(fmaxf((s + 1.0f) / 2.0f, 0.f), 1.f);
    zdepth[(static_cast<size_t>(view) * D + d) * hw + pix] = u * depth_scale + depth_shift;
  }
}

// ---------------------------------------------------------------------------------------------- tokens
__device__ __forceinline__ void bilinear_taps(float gx, float gy, int S, int* idx, float* w) {
  // grid_sample(align_corners=True, padding_mode='border'), x -> width, y -> height
  float ix = (gx + 1.f) * 0.5f * (S - 1);
  float iy = (gy + 1.f) * 0.5f * (S - 1);
  ix = fminf(fmaxf(ix, 0.f), static_cast<float>(S - 1));
  iy = fminf(fmaxf(iy, 0.f), static_cast<float>(S - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f);
  const float tx = ix - x0f, ty = iy - y0f;
  const int x1 = min(x0 + 1, S - 1), y1 = min(y0 + 1, S - 1);
  // taps past the border carry zero weight in the reference (tx or ty == 0 there)
  idx[0] = y0 * S + x0; w[0] = (1.f - tx) * (1.f - ty);
  idx[1] = y0 * S + x1; w[1] = tx * (1.f - ty);
  idx[2] = y1 * S + x0; w[2] = (1.f - tx) * ty;
  idx[3] = y1 * S + x1; w[3] = tx * ty;
}

__device__ __forceinline__ void gather8(const __half* __restrict__ map, const int* idx, const float* w, int ch,
                                        __half* __restrict__ dst) {
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint4 u = *reinterpret_cast<const uint4*>(map + static_cast<size_t>(idx[t]) * ZC + ch);
    const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(hp[j]);
      acc[2 * j] = fmaf(w[t], f.x, acc[2 * j]);
      acc[2 * j + 1] = fmaf(w[t], f.y, acc[2 * j + 1]);
    }
  }
  __half2 h0 = __floats2half2_rn(acc[0], acc[1]);
  __half2 h1 = __floats2half2_rn(acc[2], acc[3]);
  __half2 h2 = __floats2half2_rn(acc[4], acc[5]);
  __half2 h3 = __floats2half2_rn(acc[6], acc[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&h0);
  o.y = *reinterpret_cast<uint32_t*>(&h1);
  o.z = *reinterpret_cast<uint32_t*>(&h2);
  o.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst + ch) = o;
}

// harmonic embedding of `dim` values: [sin(x_i f_k) (i-major, k minor) | cos(...) | x]
// Arguments stay below ~30 (|x| <= a few scene units, f <= 6.4): the SFU sine / cosine are accurate to ~1e-5 absolute there,
// two orders below the fp16 rounding of the token they are stored into.
__device__ __forceinline__ float harmonic_value(const float* x, int dim, int j, const float* freqs) {
  const int nh = dim * N_HARM;
  if (j < nh) return __sinf(x[j / N_HARM] * freqs[j % N_HARM]);
  if (j < 2 * nh) {
    const int jj = j - nh;
    return __cosf(x[jj / N_HARM] * freqs[jj % N_HARM]);
  }
  return x[j - 2 * nh];
}

// One warp per token (query point p = ((b*hw + pix)*D + d), view v).  tokens row = p*V + v.
// cams: [n_views+1, 16] (R row-major 9, T 3, f 2, pp 2), index n_views = input camera.
// q0: first query view handled by this rank (view sharding), nq: number of local query views.
__global__ void gridattn_tokens_kernel(const __half* __restrict__ feat, const float* __restrict__ zdepth,
                                       const float* __restrict__ cams, const float* __restrict__ mask,
                                       const float* __restrict__ freqs, const float* __restrict__ ndc_grid,
                                       __half* __restrict__ tokens, int n_views, int S, int D, int q0, int nq) {
  pdl_trigger();
  pdl_wait();
  const int hw = S * S;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int total = nq * hw * D * n_views;
  if (warp_global >= total) return;
  const int v = warp_global % n_views;
  const int p = warp_global / n_views;  // local point index
  const int d = p % D;
  const int pix = (p / D) % hw;
  const int b = q0 + p / (D * hw);  // global query view
  const int py = pix / S, px = pix % S;

  float fr[N_HARM];
#pragma unroll
  for (int k = 0; k < N_HARM; ++k) fr[k] = freqs[k];

  Cam cb, cv, ci;
  cam_load(cams, b, cb);
  cam_load(cams, v, cv);
  cam_load(cams, n_views, ci);

  // --- query ray: NDC grid x,y = linspace(1-1/S, -1+1/S, S) (host-built table, bit-identical to torch.linspace);
  //     direction = ((x-px)/fx, (y-py)/fy, 1) R^T
  const float gx = ndc_grid[px];
  const float gy = ndc_grid[py];
  float dv[3] = {(gx - cb.pp[0]) / cb.f[0], (gy - cb.pp[1]) / cb.f[1], 1.f};
  float dir[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) dir[i] = dv[0] * cb.R[i * 3 + 0] + dv[1] * cb.R[i * 3 + 1] + dv[2] * cb.R[i * 3 + 2];
  float Cb[3];
  cam_center(cb, Cb);
  const float z = zdepth[(static_cast<size_t>(b) * D + d) * hw + pix];
  float X[3] = {Cb[0] + z * dir[0], Cb[1] + z * dir[1], Cb[2] + z * dir[2]};

  __half* tok = tokens + (static_cast<size_t>(p) * n_views + v) * TOKEN_LD;

  // --- features: reference view v (256) and input view (256); lane handles 8 channels of each
  {
    float x, y;
    int idx[4];
    float w[4];
    cam_project(cv, X, x, y);
    bilinear_taps(-x, -y, S, idx, w);
    gather8(feat + static_cast<size_t>(v) * hw * ZC, idx, w, lane * 8, tok);
    cam_project(ci, X, x, y);
    bilinear_taps(-x, -y, S, idx, w);
    gather8(feat + static_cast<size_t>(n_views) * hw * ZC, idx, w, lane * 8, tok + ZC);
  }

  // --- encodings
  float Cv[3];
  cam_center(cv, Cv);
  float rd[3] = {X[0] - Cv[0], X[1] - Cv[1], X[2] - Cv[2]};
  const float rlen = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
  const float rinv = 1.f / fmaxf(rlen, 1e-12f);  // F.normalize eps
  float ref_pl[6] = {rd[0] * rinv, rd[1] * rinv, rd[2] * rinv, 0.f, 0.f, 0.f};
  ref_pl[3] = Cv[1] * ref_pl[2] - Cv[2] * ref_pl[1];
  ref_pl[4] = Cv[2] * ref_pl[0] - Cv[0] * ref_pl[2];
  ref_pl[5] = Cv[0] * ref_pl[1] - Cv[1] * ref_pl[0];
  const float qlen = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  const float qinv = 1.f / fmaxf(qlen, 1e-12f);
  float q_pl[6] = {dir[0] * qinv, dir[1] * qinv, dir[2] * qinv, 0.f, 0.f, 0.f};
  q_pl[3] = Cb[1] * q_pl[2] - Cb[2] * q_pl[1];
  q_pl[4] = Cb[2] * q_pl[0] - Cb[0] * q_pl[2];
  q_pl[5] = Cb[0] * q_pl[1] - Cb[1] * q_pl[0];

  constexpr int PL = 6 * (2 * N_HARM + 1);  // 90
  constexpr int DP = 2 * N_HARM + 1;        // 15
  // layout after the 512 feature channels: ref_plucker 90 | ref_depth 15 | q_plucker 90 | q_depth 15 | mask 1 | pad
  for (int j = lane; j < TOKEN_LD - 2 * ZC; j += 32) {
    float val;
    if (j < PL) val = harmonic_value(ref_pl, 6, j, fr);
    else if (j < PL + DP) val = harmonic_value(&rlen, 1, j - PL, fr);
    else if (j < 2 * PL + DP) val = harmonic_value(q_pl, 6, j - PL - DP, fr);
    else if (j < 2 * PL + 2 * DP) val = harmonic_value(&z, 1, j - 2 * PL - DP, fr);
    else if (j == 2 * PL + 2 * DP) val = mask[v];
    else val = 0.f;
    tok[2 * ZC + j] = __float2half_rn(val);
  }
}

// ---------------------------------------------------------------------------------------------- view attention
// qkv: fp16 [P*V, 3*heads*hd] (timm layout [3][heads][hd]); out fp16 [P*V, heads*hd].
// One thread per (point, head, query view).
template <int HD>
__global__ void view_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int P, int V, int heads) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(P) * heads * V;
  if (idx >= total) return;
  const int qi = static_cast<int>(idx % V);
  const int h = static_cast<int>((idx / V) % heads);
  const size_t p = idx / (static_cast<size_t>(V) * heads);
  const int C = heads * HD;
  const int ld = 3 * C;
  const __half* base = qkv + p * V * ld;
  float q[HD];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(qi) * ld + h * HD);
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const uint4 u = qp[i];
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hp[j]);
        q[i * 8 + 2 * j] = f.x;
        q[i * 8 + 2 * j + 1] = f.y;
      }
    }
  }
  const float scale = rsqrtf(static_cast<float>(HD));
  float m = -INFINITY, l = 0.f;
  float acc[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) acc[i] = 0.f;
  for (int kj = 0; kj < V; ++kj) {
    const uint4* kp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(kj) * ld + C + h * HD);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const uint4 u = kp[i];
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hp[j]);
        s = fmaf(q[i * 8 + 2 * j], f.x, s);
        s = fmaf(q[i * 8 + 2 * j + 1], f.y, s);
      }
    }
    s *= scale;
    const float m_new = fmaxf(m, s);
    const float corr = expf(m - m_new);
    const float pexp = expf(s - m_new);
    l = l * corr + pexp;
    const uint4* vp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(kj) * ld + 2 * C + h * HD);
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const uint4 u = vp[i];
      const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hp[j]);
        acc[i * 8 + 2 * j] = acc[i * 8 + 2 * j] * corr + pexp * f.x;
        acc[i * 8 + 2 * j + 1] = acc[i * 8 + 2 * j + 1] * corr + pexp * f.y;
      }
    }
    m = m_new;
  }
  const float inv = 1.f / l;
  __half* o = out + (p * V + qi) * C + h * HD;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    __half2 h0 = __floats2half2_rn(acc[i * 8 + 0] * inv, acc[i * 8 + 1] * inv);
    __half2 h1 = __floats2half2_rn(acc[i * 8 + 2] * inv, acc[i * 8 + 3] * inv);
    __half2 h2 = __floats2half2_rn(acc[i * 8 + 4] * inv, acc[i * 8 + 5] * inv);
    __half2 h3 = __floats2half2_rn(acc[i * 8 + 6] * inv, acc[i * 8 + 7] * inv);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    reinterpret_cast<uint4*>(o)[i] = u;
  }
}

// acc += a.lo * b.lo + a.hi * b.hi   (half2 operands, fp32 accumulator; two FHFMA)
__device__ __forceinline__ void dot2_f16(float& acc, uint32_t a2, uint32_t b2) {
#ifdef MVD_CPU_EMULATION
  acc = cpu_emul::fma_f16(cpu_emul::fma_f16(acc, a2 & 0xffffu, b2 & 0xffffu), a2 >> 16, b2 >> 16);
  return;
#endif
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.f16 %0, al, bl, %0;\n\tfma.rn.f32.f16 %0, ah, bh, %0;\n\t}\n" : "+f"(acc) : "r"(a2), "r"(b2));
}
// acc0 += p * v.lo ; acc1 += p * v.hi   (p duplicated in both halves of p2)
__device__ __forceinline__ void axpy2_f16(float& acc0, float& acc1, uint32_t p2, uint32_t v2) {
#ifdef MVD_CPU_EMULATION
  acc0 = cpu_emul::fma_f16(acc0, p2 & 0xffffu, v2 & 0xffffu);
  acc1 = cpu_emul::fma_f16(acc1, p2 >> 16, v2 >> 16);
  return;
#endif
  asm("{\n\t.reg .b16 pl, ph, vl, vh;\n\tmov.b32 {pl, ph}, %2;\n\tmov.b32 {vl, vh}, %3;\n\t"
      "fma.rn.f32.f16 %0, pl, vl, %0;\n\tfma.rn.f32.f16 %1, ph, vh, %1;\n\t}\n" : "+f"(acc0), "+f"(acc1) : "r"(p2), "r"(v2));
}

// Staged variant (heads * V divides 256, V <= 16): a CTA of 256 threads owns 256 / (heads * V) consecutive points.  The q | k | v rows
// of a point are one contiguous block of V * 3C halves, so the CTA pulls its 32 rows (48 KB) into shared memory with
// fully coalesced 16-byte loads, every thread = (point, head, query view) then works out of shared memory (k / v reads are
// broadcasts across the query views), and writes its 64-byte output segment.
template <int HD>
__global__ void __launch_bounds__(256)
    view_attention_staged_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int P, int V, int heads) {
  MVD_DYNAMIC_SHARED_ALIGNED16(uint8_t, va_smem);
  pdl_trigger();
  pdl_wait();
  const int C = heads * HD;
  const int ld = 3 * C;                       // halves per row
  const int pts = 256 / (heads * V);          // points per CTA
  const size_t p0 = static_cast<size_t>(blockIdx.x) * pts;
  const int npts = static_cast<int>(min(static_cast<size_t>(pts), static_cast<size_t>(P) - p0));
  const int rows = npts * V;
  const int vec_per_row = ld / 8;             // 16-byte vectors per row
  const uint4* src = reinterpret_cast<const uint4*>(qkv + p0 * V * ld);
  uint4* sm = reinterpret_cast<uint4*>(va_smem);
  const int nvec = rows * vec_per_row;
  for (int i = threadIdx.x; i < nvec; i += 256) sm[i] = __ldg(src + i);
  __syncthreads();
  const int t = threadIdx.x;
  const int qi = t % V;
  const int h = (t / V) % heads;
  const int pl = t / (V * heads);
  if (pl >= npts) return;
  const __half* base = reinterpret_cast<const __half*>(va_smem) + static_cast<size_t>(pl) * V * ld;
  // q stays packed (half2); products run on sm_100's mixed-precision FMA (fp16 x fp16 -> exact, accumulated in fp32: FHFMA),
  // so no operand is ever converted.  Two passes over the <= 16 keys: scores -> maximum -> P = exp2(s - m) rounded to fp16
  // (as in the tensor-core attention; l is the sum of the rounded values) -> P V.
  uint32_t qh[HD / 2];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(base + qi * ld + h * HD);
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      const uint4 u = qp[i];
      qh[i * 4] = u.x; qh[i * 4 + 1] = u.y; qh[i * 4 + 2] = u.z; qh[i * 4 + 3] = u.w;
    }
  }
  const float scale_log2 = rsqrtf(static_cast<float>(HD)) * 1.4426950408889634f;
  constexpr int MAXV = 16;
  float sc[MAXV];
  float m = -INFINITY;
#pragma unroll
  for (int kj = 0; kj < MAXV; ++kj) {
    if (kj < V) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + kj * ld + C + h * HD);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        const uint4 u = kp[i];
        dot2_f16(s0, qh[i * 4], u.x);
        dot2_f16(s1, qh[i * 4 + 1], u.y);
        dot2_f16(s0, qh[i * 4 + 2], u.z);
        dot2_f16(s1, qh[i * 4 + 3], u.w);
      }
      sc[kj] = (s0 + s1) * scale_log2;
      m = fmaxf(m, sc[kj]);
    } else {
      sc[kj] = -INFINITY;
    }
  }
  float l = 0.f;
  float acc[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) acc[i] = 0.f;
#pragma unroll
  for (int kj = 0; kj < MAXV; ++kj) {
    if (kj < V) {
      float pe;
#ifdef MVD_CPU_EMULATION
      pe = exp2f(sc[kj] - m);
#else
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(sc[kj] - m));
#endif
      const __half ph = __float2half_rn(pe);
      l += __half2float(ph);
      const uint32_t p2 = static_cast<uint32_t>(__half_as_ushort(ph)) * 0x10001u;
      const uint4* vp = reinterpret_cast<const uint4*>(base + kj * ld + 2 * C + h * HD);
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        const uint4 u = vp[i];
        axpy2_f16(acc[i * 8 + 0], acc[i * 8 + 1], p2, u.x);
        axpy2_f16(acc[i * 8 + 2], acc[i * 8 + 3], p2, u.y);
        axpy2_f16(acc[i * 8 + 4], acc[i * 8 + 5], p2, u.z);
        axpy2_f16(acc[i * 8 + 6], acc[i * 8 + 7], p2, u.w);
      }
    }
  }
  const float inv = 1.f / l;
  __half* o = out + ((p0 + pl) * V + qi) * C + h * HD;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    __half2 h0 = __floats2half2_rn(acc[i * 8 + 0] * inv, acc[i * 8 + 1] * inv);
    __half2 h1 = __floats2half2_rn(acc[i * 8 + 2] * inv, acc[i * 8 + 3] * inv);
    __half2 h2 = __floats2half2_rn(acc[i * 8 + 4] * inv, acc[i * 8 + 5] * inv);
    __half2 h3 = __floats2half2_rn(acc[i * 8 + 6] * inv, acc[i * 8 + 7] * inv);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    reinterpret_cast<uint4*>(o)[i] = u;
  }
}

// ---------------------------------------------------------------------------------------------- view pool
// x fp32 [P*V, 256]; w = x . ww + wb; softmax over V; out[p] = sum_v softmax_v x_v  -> fp16 [P, 256]
__global__ void view_pool_kernel(const float* __restrict__ x, const float* __restrict__ ww, const float* __restrict__ wb,
                                 __half* __restrict__ out, int P, int V) {
  pdl_trigger();
  pdl_wait();
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= P) return;
  const float* xp = x + static_cast<size_t>(p) * V * ZC;
  float wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) wv[j] = ww[lane * 8 + j];
  const float bias = wb[0];
  float m = -INFINITY, l = 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int v = 0; v < V; ++v) {
    const float4 a = *reinterpret_cast<const float4*>(xp + static_cast<size_t>(v) * ZC + lane * 8);
    const float4 b = *reinterpret_cast<const float4*>(xp + static_cast<size_t>(v) * ZC + lane * 8 + 4);
    const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(xv[j], wv[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += bias;
    const float m_new = fmaxf(m, s);
    const float corr = expf(m - m_new);
    const float pe = expf(s - m_new);
    l = l * corr + pe;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = acc[j] * corr + pe * xv[j];
    m = m_new;
  }
  const float inv = 1.f / l;
  __half2 h0 = __floats2half2_rn(acc[0] * inv, acc[1] * inv);
  __half2 h1 = __floats2half2_rn(acc[2] * inv, acc[3] * inv);
  __half2 h2 = __floats2half2_rn(acc[4] * inv, acc[5] * inv);
  __half2 h3 = __floats2half2_rn(acc[6] * inv, acc[7] * inv);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(out + static_cast<size_t>(p) * ZC + lane * 8) = u;
}

// ---------------------------------------------------------------------------------------------- frustum pyramid
// in fp16 [n, S, S, D, C] -> out fp16 [n, S/f, S/f, D, C], mean over f x f pixel blocks (interpolate mode='area')
__global__ void frustum_pool_kernel(const __half* __restrict__ in, __half* __restrict__ out, int S, int D, int C, int f,
                                    size_t total8) {
  pdl_trigger();
  pdl_wait();
  const int So = S / f;
  const float inv = 1.f / static_cast<float>(f * f);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 8;
    const int c = static_cast<int>(e % C);
    size_t r = e / C;
    const int d = static_cast<int>(r % D);
    r /= D;
    const int ox = static_cast<int>(r % So);
    r /= So;
    const int oy = static_cast<int>(r % So);
    const size_t img = r / So;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) {
        const size_t src = (((img * S + (oy * f + dy)) * S + (ox * f + dx)) * D + d) * C + c;
        const uint4 u = *reinterpret_cast<const uint4*>(in + src);
        const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(hp[j]);
          acc[2 * j] += t.x;
          acc[2 * j + 1] += t.y;
        }
      }
    __half2 h0 = __floats2half2_rn(acc[0] * inv, acc[1] * inv);
    __half2 h1 = __floats2half2_rn(acc[2] * inv, acc[3] * inv);
    __half2 h2 = __floats2half2_rn(acc[4] * inv, acc[5] * inv);
    __half2 h3 = __floats2half2_rn(acc[6] * inv, acc[7] * inv);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(out + e) = u;
  }
}

// ---------------------------------------------------------------------------------------------- pixel cross-attention
// DualAttnetionBlock.attn2 for D > 1 (mvdfusion/attention.py:56-62): each pixel is one query against its D frustum
// keys.  q fp16 [M, C]; kv fp16 [M*D, 2C] (k | v); out fp16 [M, C].  One thread per (pixel, head, 8-channel group).
__global__ void pixel_cross_attn_kernel(const __half* __restrict__ q, const __half* __restrict__ kv,
                                        __half* __restrict__ out, int M, int D, int heads, int dhead) {
  pdl_trigger();
  pdl_wait();
  // one warp per (pixel, head): lanes stride the head dim
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global >= M * heads) return;
  const int h = warp_global % heads;
  const size_t m = warp_global / heads;
  const int C = heads * dhead;
  const float scale = rsqrtf(static_cast<float>(dhead));
  float s[8];  // D <= 8
  for (int d = 0; d < D; ++d) {
    float a = 0.f;
    for (int j = lane; j < dhead; j += 32)
      a = fmaf(__half2float(q[m * C + h * dhead + j]), __half2float(kv[(m * D + d) * 2 * C + h * dhead + j]), a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    s[d] = a * scale;
  }
  float mx = -INFINITY;
  for (int d = 0; d < D; ++d) mx = fmaxf(mx, s[d]);
  float l = 0.f;
  for (int d = 0; d < D; ++d) {
    s[d] = expf(s[d] - mx);
    l += s[d];
  }
  const float inv = 1.f / l;
  for (int j = lane; j < dhead; j += 32) {
    float a = 0.f;
    for (int d = 0; d < D; ++d) a = fmaf(s[d] * inv, __half2float(kv[(m * D + d) * 2 * C + C + h * dhead + j]), a);
    out[m * C + h * dhead + j] = __float2half_rn(a);
  }
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_gridattn_prep(const float* noisy, const float* input_latent, const float* depth_override,
                                 const float* depth_eps, const float* scal_dev, const float* Wz, const float* bz,
                                 void* feat_out, float* zdepth_out, int32_t n_views, int32_t S, int32_t D,
                                 float depth_scale, float depth_shift, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!noisy || !input_latent || !depth_eps || !scal_dev || !Wz || !bz || !feat_out || !zdepth_out)
    return set_error(MVD_EINVAL, "mvd_gridattn_prep: null pointer");
  if (n_views <= 0 || S <= 0 || D <= 0 || D > 32) return set_error(MVD_EINVAL, "mvd_gridattn_prep: bad sizes");
  MVD_LAUNCH((gridattn_prep_kernel), dim3(S * S, n_views + 1), 64, 0, stream, noisy, input_latent, depth_override, depth_eps,
                                                                   scal_dev, Wz, bz, static_cast<__half*>(feat_out),
                                                                   zdepth_out, n_views, S * S, D, depth_scale,
                                                                   depth_shift);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_gridattn_tokens(const void* feat, const float* zdepth, const float* cams, const float* mask,
                                   const float* freqs, const float* ndc_grid, void* tokens, int32_t n_views,
                                   int32_t S, int32_t D, int32_t q_first, int32_t q_count, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!feat || !zdepth || !cams || !mask || !freqs || !ndc_grid || !tokens) return set_error(MVD_EINVAL, "mvd_gridattn_tokens: null pointer");
  if (n_views <= 0 || S <= 1 || D <= 0 || q_first < 0 || q_count <= 0 || q_first + q_count > n_views)
    return set_error(MVD_EINVAL, "mvd_gridattn_tokens: bad sizes");
  const long long total = static_cast<long long>(q_count) * S * S * D * n_views;
  const int wpb = 8;
  MVD_LAUNCH((gridattn_tokens_kernel), static_cast<unsigned>((total + wpb - 1) / wpb), wpb * 32, 0, stream, 
      static_cast<const __half*>(feat), zdepth, cams, mask, freqs, ndc_grid, static_cast<__half*>(tokens), n_views, S, D,
      q_first, q_count);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_view_attention_f16(const void* qkv, void* out, int32_t P, int32_t V, int32_t heads, int32_t hd,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !out || P <= 0 || V <= 0 || heads <= 0) return set_error(MVD_EINVAL, "mvd_view_attention_f16: bad arguments");
  if (hd != 32) return set_error(MVD_EINVAL, "mvd_view_attention_f16: head dim must be 32");
  const size_t total = static_cast<size_t>(P) * heads * V;
  if (256 % (heads * V) == 0 && V <= 16 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int pts = 256 / (heads * V);
    const size_t smem = static_cast<size_t>(pts) * V * 3 * heads * 32 * sizeof(__half);
    static bool configured = false;
    if (!configured) {
      MVD_CUDA_CHECK(cudaFuncSetAttribute(view_attention_staged_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      configured = true;
    }
    if (smem <= 64 * 1024) {
      MVD_LAUNCH((view_attention_staged_kernel<32>), static_cast<unsigned>((P + pts - 1) / pts), 256, smem, stream,
                 static_cast<const __half*>(qkv), static_cast<__half*>(out), P, V, heads);
      count_launch();
      MVD_CUDA_CHECK(cudaGetLastError());
      return MVD_OK;
    }
  }
  MVD_LAUNCH((view_attention_kernel<32>), static_cast<unsigned>((total + 127) / 128), 128, 0, stream, 
      static_cast<const __half*>(qkv), static_cast<__half*>(out), P, V, heads);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_view_pool_f16(const float* x, const float* w, const float* b, void* out, int32_t P, int32_t V,
                                 int32_t C, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w || !b || !out || P <= 0 || V <= 0) return set_error(MVD_EINVAL, "mvd_view_pool_f16: bad arguments");
  if (C != ZC) return set_error(MVD_EINVAL, "mvd_view_pool_f16: hidden size must be 256");
  MVD_LAUNCH((view_pool_kernel), (P + 7) / 8, 256, 0, stream, x, w, b, static_cast<__half*>(out), P, V);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_frustum_pool_f16(const void* in, void* out, int32_t n_img, int32_t S, int32_t D, int32_t C,
                                    int32_t factor, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || n_img <= 0 || S <= 0 || D <= 0 || C <= 0 || (C & 7) || factor <= 0 || (S % factor) != 0)
    return set_error(MVD_EINVAL, "mvd_frustum_pool_f16: bad arguments");
  const int So = S / factor;
  const size_t total8 = static_cast<size_t>(n_img) * So * So * D * C / 8;
  size_t blocks = (total8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  MVD_LAUNCH((frustum_pool_kernel), static_cast<unsigned>(blocks), 256, 0, stream, 
      static_cast<const __half*>(in), static_cast<__half*>(out), S, D, C, factor, total8);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_pixel_cross_attn_f16(const void* q, const void* kv, void* out, int32_t M, int32_t D, int32_t heads,
                                        int32_t dhead, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!q || !kv || !out || M <= 0 || D <= 0 || D > 8 || heads <= 0 || dhead <= 0)
    return set_error(MVD_EINVAL, "mvd_pixel_cross_attn_f16: bad arguments (D <= 8)");
  const long long warps = static_cast<long long>(M) * heads;
  MVD_LAUNCH((pixel_cross_attn_kernel), static_cast<unsigned>((warps + 7) / 8), 256, 0, stream, 
      static_cast<const __half*>(q), static_cast<const __half*>(kv), static_cast<__half*>(out), M, D, heads, dhead);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}
