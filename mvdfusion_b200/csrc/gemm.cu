// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   acc[m, n] = sum_k A[m, k] * W[n, k]     fp16 operands (K-major, 128B-swizzled smem tiles via TMA),
//                                           fp32 accumulators in TMEM, fused epilogues.
//
// One CTA computes a 128 x BN output tile over a contiguous range of 64-wide k-blocks:
//   warp 0      : TMA producer (one elected lane) — A tile + W tile per k-block into a STAGES-deep ring
//   warp 1      : TMEM allocator + UMMA issuer (one elected lane), tcgen05.commit releases ring slots
//   warps 2..5  : epilogue — tcgen05.ld (each warp owns the TMEM lane quarter warp%4), bias / row-bias /
//                 activation / GEGLU / residual, then direct global stores (or red.add for split-K)
// For MVD_A_CONV3X3 the A tile of k-block (tap, c-block) is a 4-D TMA box (64 ch, tw, th, tn) of the
// NHWC image shifted by (kx-1, ky-1); out-of-bounds pixels are zero-filled by TMA == zero padding,
// so im2col never exists in memory.
//
// Replaces the cuBLAS/cuDNN dispatch behind nn.Linear / nn.Conv2d on the reference hot path
// (see include/mvd_b200.h for the file:line list).
#include "common.h"
#include "ptx.cuh"

namespace mvd {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmKParams {
  int M, N;
  int num_kb;        // total 64-wide k-blocks
  int kb_per_split;  // k-blocks per blockIdx.z
  int a_mode;
  int kb_per_tap;    // conv: ceil(C/64)
  int C;             // conv: channels (W column offset of a tap = tap*C)
  int n_img, H, W;
  int tw, th, tn, tiles_x, tiles_y;
  const float* bias;
  const float* rowbias;
  int rows_per_group;
  const float* colscale;
  const float* residual;
  int ldr;
  int act, out_mode;
  void* out;
  int ldc;
  void* out_k;
  void* out_vt;
  int heads, dhead, dpad, seq;
  int split_k;
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
};

__device__ __forceinline__ void store8_f16(__half* dst, const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]);
  __half2 h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]);
  __half2 h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = u;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmKParams p) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;

  // conv tile origin
  int x0 = 0, y0 = 0, img0 = 0;
  if (p.a_mode == MVD_A_CONV3X3) {
    const int tx = m_tile % p.tiles_x;
    const int ty = (m_tile / p.tiles_x) % p.tiles_y;
    const int tz = m_tile / (p.tiles_x * p.tiles_y);
    x0 = tx * p.tw;
    y0 = ty * p.th;
    img0 = tz * p.tn;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb0 + i;
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        uint8_t* sa = smem + s * S::STAGE_BYTES;
        uint8_t* sb = sa + S::A_BYTES;
        int kcol;
        if (p.a_mode == MVD_A_CONV3X3) {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          const int ky = tap / 3, kx = tap - ky * 3;
          tma_load_4d(sa, &tmA, &full_bar[s], cb * BK, x0 + kx - 1, y0 + ky - 1, img0);
          kcol = tap * p.C + cb * BK;
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m_tile * BM);
          kcol = kb * BK;
        }
        tma_load_2d(sb, &tmB, &full_bar[s], kcol, n_tile * BN);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t sb = sa + S::A_BYTES;
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 fp16 = 32 B inside the 128-B swizzle atom: +2 in the (addr >> 4) field
          umma_f16(tmem_acc, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);
      }
      tc_commit(accum_bar);
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    int grow;
    bool valid;
    if (p.a_mode == MVD_A_CONV3X3) {
      const int wi = r % p.tw;
      const int hi = (r / p.tw) % p.th;
      const int ni = r / (p.tw * p.th);
      const int img = img0 + ni;
      valid = img < p.n_img;
      grow = (img * p.H + y0 + hi) * p.W + x0 + wi;
    } else {
      grow = m_tile * BM + r;
      valid = grow < p.M;
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const int n_base = n_tile * BN;
    const bool lead = (blockIdx.z == 0);
    const float* rb = nullptr;
    if (p.rowbias != nullptr && valid) rb = p.rowbias + static_cast<size_t>(grow / p.rows_per_group) * p.N;

    if (p.act == MVD_ACT_GEGLU) {
      const int n_out = p.N / 2;
      const int o_base = n_tile * (BN / 2);
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        if (o_base + c * 32 >= n_out) break;
        float v[32], g[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld32(taddr + BN / 2 + c * 32, g);
        tmem_ld_wait();
        if (!valid) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int nv = n_base + c * 32 + i;
          const int ng = nv + BN / 2;
          float a = v[i], b = g[i];
          if (p.bias != nullptr) {
            a += (nv < p.N) ? __ldg(p.bias + nv) : 0.f;
            b += (ng < p.N) ? __ldg(p.bias + ng) : 0.f;
          }
          v[i] = a * gelu_erf(b);
        }
        const int oc = o_base + c * 32;
        if (p.out_mode == MVD_OUT_F16) {
          __half* dst = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(grow) * p.ldc + oc;
          if (oc + 32 <= n_out) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) store8_f16(dst + i, v + i);
          } else {
            for (int i = 0; i < 32 && oc + i < n_out; ++i) dst[i] = __float2half_rn(v[i]);
          }
        } else {
          float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(grow) * p.ldc + oc;
          for (int i = 0; i < 32 && oc + i < n_out; ++i) dst[i] = v[i];
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nc = n_base + c * 32;
        if (nc >= p.N) break;
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (!valid) continue;
        const bool full = (nc + 32 <= p.N);
        if (lead) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += (full || nc + i < p.N) ? __ldg(p.bias + nc + i) : 0.f;
          }
          if (rb != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += (full || nc + i < p.N) ? __ldg(rb + nc + i) : 0.f;
          }
        }
        if (p.act == MVD_ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        } else if (p.act == MVD_ACT_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = silu(v[i]);
        }
        if (p.colscale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= (full || nc + i < p.N) ? __ldg(p.colscale + nc + i) : 0.f;
        }
        if (lead && p.residual != nullptr) {
          const float* res = p.residual + static_cast<size_t>(grow) * p.ldr + nc;
          if (full && (p.ldr & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 t = *reinterpret_cast<const float4*>(res + i);
              v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
            }
          } else {
            for (int i = 0; i < 32 && nc + i < p.N; ++i) v[i] += res[i];
          }
        }
        if (p.out_mode == MVD_OUT_F32) {
          float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(grow) * p.ldc + nc;
          if (p.split_k > 1) {
            for (int i = 0; i < 32 && nc + i < p.N; ++i) atomicAdd(dst + i, v[i]);
          } else if (full && (p.ldc & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            for (int i = 0; i < 32 && nc + i < p.N; ++i) dst[i] = v[i];
          }
        } else if (p.out_mode == MVD_OUT_F16) {
          __half* dst = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(grow) * p.ldc + nc;
          if (full && (p.ldc & 7) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) store8_f16(dst + i, v + i);
          } else {
            for (int i = 0; i < 32 && nc + i < p.N; ++i) dst[i] = __float2half_rn(v[i]);
          }
        } else {  // MVD_OUT_QKV_HEADS
          const int inner = p.heads * p.dhead;
          const int img = grow / p.seq;
          const int pos = grow - img * p.seq;
#pragma unroll 1
          for (int i = 0; i < 32; i += 8) {
            const int n = nc + i;
            if (n >= p.N) break;
            const int which = n / inner;
            const int rem = n - which * inner;
            const int h = rem / p.dhead;
            const int j = rem - h * p.dhead;
            const size_t bh = static_cast<size_t>(img) * p.heads + h;
            if (which < 2) {
              __half* base = reinterpret_cast<__half*>(which == 0 ? p.out : p.out_k);
              store8_f16(base + (bh * p.seq + pos) * p.dpad + j, v + i);
            } else {
              __half* base = reinterpret_cast<__half*>(p.out_vt) + (bh * p.dpad + j) * p.seq + pos;
#pragma unroll
              for (int e = 0; e < 8; ++e) base[static_cast<size_t>(e) * p.seq] = __float2half_rn(v[i + e]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, BN);
  }
}

// ------------------------------------------------------------------------------------------------ host
template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKParams& p, int m_tiles,
                       cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, STAGES>;
  if (!configured) {
    MVD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES));
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, m_tiles, p.split_k);
  kern<<<grid, GEMM_THREADS, S::DYN_BYTES, stream>>>(tmA, tmB, p);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_gemm_f16(const mvd_gemm_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return set_error(MVD_EINVAL, "mvd_gemm_f16: null args");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(MVD_EINVAL, "mvd_gemm_f16: M, N, K must be positive");
  if (a->A == nullptr || a->Wt == nullptr || a->out == nullptr)
    return set_error(MVD_EINVAL, "mvd_gemm_f16: null A / Wt / out");
  if ((a->ldw & 7) != 0 || a->ldw < a->K) return set_error(MVD_EALIGN, "mvd_gemm_f16: ldw must be >= K and a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->Wt) & 15))
    return set_error(MVD_EALIGN, "mvd_gemm_f16: A and Wt must be 16-byte aligned");
  if (a->act < MVD_ACT_NONE || a->act > MVD_ACT_GEGLU) return set_error(MVD_EINVAL, "mvd_gemm_f16: bad act");
  if (a->out_mode < MVD_OUT_F32 || a->out_mode > MVD_OUT_QKV_HEADS)
    return set_error(MVD_EINVAL, "mvd_gemm_f16: bad out_mode");

  GemmKParams p{};
  p.M = a->M;
  p.N = a->N;
  p.a_mode = a->a_mode;
  p.bias = a->bias;
  p.rowbias = a->rowbias;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.colscale = a->colscale;
  p.residual = a->residual;
  p.ldr = a->ldr;
  p.act = a->act;
  p.out_mode = a->out_mode;
  p.out = a->out;
  p.ldc = a->ldc;
  p.out_k = a->out_k;
  p.out_vt = a->out_vt;
  p.heads = a->heads;
  p.dhead = a->dhead;
  p.dpad = a->dpad;
  p.seq = a->seq;

  int bn = a->tile_n;
  if (bn == 0) bn = (a->N <= 64) ? 64 : 128;
  if (bn != 64 && bn != 128 && bn != 256) return set_error(MVD_EINVAL, "mvd_gemm_f16: tile_n must be 0, 64, 128 or 256");
  if (a->act == MVD_ACT_GEGLU && (a->N % bn) != 0)
    return set_error(MVD_EINVAL, "mvd_gemm_f16: GEGLU needs N to be a multiple of tile_n");
  if (a->out_mode == MVD_OUT_QKV_HEADS) {
    if (a->out_k == nullptr || a->out_vt == nullptr || a->heads <= 0 || a->dhead <= 0 || (a->dhead & 7) != 0 ||
        (a->dpad & 7) != 0 || a->dpad < a->dhead || a->seq <= 0 || a->N != 3 * a->heads * a->dhead ||
        (a->M % a->seq) != 0 || a->act != MVD_ACT_NONE)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: bad QKV_HEADS arguments");
  }

  CUtensorMap tmA, tmB;
  int m_tiles;
  if (a->a_mode == MVD_A_ROWMAJOR) {
    if ((a->lda & 7) != 0 || a->lda < a->K) return set_error(MVD_EALIGN, "mvd_gemm_f16: lda must be >= K and a multiple of 8");
    p.num_kb = (a->K + BK - 1) / BK;
    m_tiles = (a->M + BM - 1) / BM;
    int rc = make_tmap_2d(&tmA, a->A, /*cols=*/a->K, /*rows=*/a->M, /*ld=*/a->lda, BK, BM);
    if (rc != MVD_OK) return rc;
  } else if (a->a_mode == MVD_A_CONV3X3) {
    if (a->n_img <= 0 || a->H <= 0 || a->W <= 0 || a->C <= 0 || (a->C & 7) != 0 || !is_pow2(a->W) || a->W > 128 ||
        a->K != 9 * a->C || a->M != a->n_img * a->H * a->W)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: bad CONV3X3 geometry");
    p.n_img = a->n_img;
    p.H = a->H;
    p.W = a->W;
    p.C = a->C;
    p.tw = a->W;
    p.th = BM / p.tw;
    if (p.th > a->H) p.th = a->H;
    if (!is_pow2(p.th) || (a->H % p.th) != 0) return set_error(MVD_EINVAL, "mvd_gemm_f16: CONV3X3 needs H a multiple of the tile height");
    p.tn = BM / (p.tw * p.th);
    p.tiles_x = 1;
    p.tiles_y = a->H / p.th;
    const int tiles_z = (a->n_img + p.tn - 1) / p.tn;
    m_tiles = p.tiles_x * p.tiles_y * tiles_z;
    p.kb_per_tap = (a->C + BK - 1) / BK;
    p.num_kb = 9 * p.kb_per_tap;
    int rc = make_tmap_nhwc(&tmA, a->A, a->n_img, a->H, a->W, a->C, BK, p.tw, p.th, p.tn);
    if (rc != MVD_OK) return rc;
  } else {
    return set_error(MVD_EINVAL, "mvd_gemm_f16: bad a_mode");
  }
  {
    int rc = make_tmap_2d(&tmB, a->Wt, /*cols=*/a->K, /*rows=*/a->N, /*ld=*/a->ldw, BK, bn);
    if (rc != MVD_OK) return rc;
  }

  int split = a->split_k > 0 ? a->split_k : 1;
  if (split > p.num_kb) split = p.num_kb;
  p.kb_per_split = (p.num_kb + split - 1) / split;
  split = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.split_k = split;
  if (split > 1) {
    if (a->out_mode != MVD_OUT_F32 || a->act != MVD_ACT_NONE || a->colscale != nullptr)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: split_k needs F32 output and no activation");
    if (a->ldc == a->N) {
      MVD_CUDA_CHECK(cudaMemsetAsync(a->out, 0, static_cast<size_t>(a->M) * a->N * sizeof(float), stream));
    } else {
      MVD_CUDA_CHECK(cudaMemset2DAsync(a->out, static_cast<size_t>(a->ldc) * sizeof(float), 0,
                                       static_cast<size_t>(a->N) * sizeof(float), a->M, stream));
    }
  }

  switch (bn) {
    case 64:
      return launch_gemm<64, 4>(tmA, tmB, p, m_tiles, stream);
    case 128:
      return launch_gemm<128, 3>(tmA, tmB, p, m_tiles, stream);
    default:
      return launch_gemm<256, 4>(tmA, tmB, p, m_tiles, stream);
  }
}

extern "C" int mvd_geglu_row_permutation(int32_t inner, int32_t tile_n, int32_t* perm) {
  if (inner <= 0 || perm == nullptr || (tile_n != 64 && tile_n != 128 && tile_n != 256) || (2 * inner) % tile_n != 0)
    return set_error(MVD_EINVAL, "mvd_geglu_row_permutation: bad arguments");
  const int half = tile_n / 2;
  for (int r = 0; r < 2 * inner; ++r) {
    const int tile = r / tile_n, w = r % tile_n;
    // nn.Linear(dim, 2*inner): rows [0, inner) are the value half, [inner, 2*inner) the gate half (chunk(2, dim=-1))
    perm[r] = (w < half) ? tile * half + w : inner + tile * half + (w - half);
  }
  return MVD_OK;
}
