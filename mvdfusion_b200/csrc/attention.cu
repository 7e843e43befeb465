// Fused multi-head self-attention for sm_100a: softmax(Q K^T * d^-1/2) V without materialising the
// (seq x seq) score matrix.  Replaces the einsum / softmax / einsum of CrossAttention.forward
// (external/sd1/ldm/modules/attention.py:177-192) for context=None (BasicTransformerBlock.attn1,
// DualAttnetionBlock.attn1).
//
// Layouts (written by mvd_gemm_f16 with MVD_OUT_QKV_HEADS):
//   Q, K : fp16 [BH, seq, dpad]  (head dim zero-padded to dpad = multiple of 64)
//   Vt   : fp16 [BH, dpad, seq]  (transposed so the PV product is K-major on both operands)
//   out  : fp16 [n_img*seq, heads*dhead]  ('b n (h d)'), the A operand of to_out.
//
// One CTA = 128 query rows of one (image, head).  Warps 0-3: one thread per query row (TMEM lane),
// warp 4: control (TMA loads + tcgen05.mma issue).  S = Q K^T lives in TMEM (BKV fp32 columns),
// O accumulates in TMEM next to it.  Exact two-pass softmax: pass 1 computes the row maxima from
// S tiles only, pass 2 recomputes S, writes P = exp2(S*c - m*c) as fp16 into a 128B-swizzled smem
// tile and accumulates O += P V_j and l += rowsum(P).  When the whole key range fits one tile
// (seq <= BKV) S is computed once.
#include "common.h"
#include "ptx.cuh"

namespace mvd {

constexpr int ATT_BM = 128;
constexpr int ATT_THREADS = 160;

struct AttnParams {
  int seq, heads, dhead, dpad, bkv;
  int n_tiles;      // key tiles
  int kd_steps;     // ceil(dhead/16): k-steps of the QK^T product
  int n_o;          // kd_steps*16: columns of O
  int tmem_cols;    // power of two >= bkv + n_o
  float scale_log2; // dhead^-0.5 * log2(e)
  __half* out;
  int ldo;
  // smem byte offsets (from the 1024-aligned base)
  int off_k, off_v, off_p, off_bar;
};

__device__ __forceinline__ void st_shared_16(uint32_t addr, const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]);
  __half2 h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]);
  __half2 h3 = __floats2half2_rn(v[6], v[7]);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&h0)),
               "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
               "r"(*reinterpret_cast<uint32_t*>(&h3))
               : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS)
    attn_self_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + p.off_k;
  uint8_t* sV = smem + p.off_v;
  uint8_t* sP = smem + p.off_p;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* bar_q = bars + 0;       // TMA Q landed
  uint64_t* bar_kv = bars + 1;      // TMA K (+V) tile landed
  uint64_t* bar_s_full = bars + 2;  // QK^T MMA retired
  uint64_t* bar_s_free = bars + 3;  // 128 row threads finished reading S
  uint64_t* bar_p_ready = bars + 4; // 128 row threads wrote P
  uint64_t* bar_pv_done = bars + 5; // PV MMA retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int bh = blockIdx.y;
  const bool two_pass = p.n_tiles > 1;
  const int atoms_d = p.dpad / 64;   // 64-wide k atoms of Q / K tiles
  const int atoms_kv = (p.bkv + 63) / 64;  // 64-key atoms of the P / Vt tiles

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_free, ATT_BM);
    mbar_init(bar_p_ready, ATT_BM);
    mbar_init(bar_pv_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;
  const uint32_t tmem_o = tmem_s + p.bkv;

  const int total_iters = (two_pass ? 2 : 1) * p.n_tiles;

  if (warp == 4) {
    if (lane == 0) {
      // ---------------------------------------------------------------- control thread
      const uint32_t q_bytes = ATT_BM * p.dpad * 2;
      const uint32_t k_bytes = p.bkv * p.dpad * 2;
      const uint32_t v_bytes = atoms_kv * p.dpad * 128;
      mbar_expect_tx(bar_q, q_bytes);
      for (int a = 0; a < atoms_d; ++a) tma_load_3d(sQ + a * (ATT_BM * 128), &tmQ, bar_q, a * 64, q0, bh);
      mbar_wait(bar_q, 0);
      const uint32_t idesc_s = umma_idesc_f16(ATT_BM, p.bkv);
      const uint32_t idesc_o = umma_idesc_f16(ATT_BM, p.n_o);
      int pv_count = 0;
      for (int it = 0; it < total_iters; ++it) {
        const bool pv_pass = !two_pass || it >= p.n_tiles;
        const int j = two_pass ? (it % p.n_tiles) : it;
        mbar_expect_tx(bar_kv, k_bytes + (pv_pass ? v_bytes : 0));
        for (int a = 0; a < atoms_d; ++a) tma_load_3d(sK + a * (p.bkv * 128), &tmK, bar_kv, a * 64, j * p.bkv, bh);
        if (pv_pass)
          for (int a = 0; a < atoms_kv; ++a) tma_load_3d(sV + a * (p.dpad * 128), &tmV, bar_kv, j * p.bkv + a * 64, 0, bh);
        mbar_wait(bar_kv, it & 1);
        if (it > 0) mbar_wait(bar_s_free, (it - 1) & 1);
        tc_fence_after();
        for (int ks = 0; ks < p.kd_steps; ++ks) {
          const uint32_t aq = smem_u32(sQ) + (ks >> 2) * (ATT_BM * 128) + (ks & 3) * 32;
          const uint32_t ak = smem_u32(sK) + (ks >> 2) * (p.bkv * 128) + (ks & 3) * 32;
          umma_f16(tmem_s, umma_desc_sw128(aq), umma_desc_sw128(ak), idesc_s, ks > 0 ? 1u : 0u);
        }
        tc_commit(bar_s_full);
        if (pv_pass) {
          mbar_wait(bar_p_ready, pv_count & 1);
          tc_fence_after();
          const int ksteps = (min(p.bkv, p.seq - j * p.bkv) + 15) / 16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t ap = smem_u32(sP) + (ks >> 2) * (ATT_BM * 128) + (ks & 3) * 32;
            const uint32_t av = smem_u32(sV) + (ks >> 2) * (p.dpad * 128) + (ks & 3) * 32;
            umma_f16(tmem_o, umma_desc_sw128(ap), umma_desc_sw128(av), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(bar_pv_done);
          mbar_wait(bar_pv_done, pv_count & 1);
          ++pv_count;
        } else {
          mbar_wait(bar_s_full, it & 1);  // K tile may be overwritten once the MMA has read it
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ one thread per query row
    const int r = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    float m = -INFINITY;  // running max of raw scores (unscaled)
    float l = 0.f;
    float mc = 0.f;       // m * scale_log2, fixed during the P pass
    for (int it = 0; it < total_iters; ++it) {
      const bool pv_pass = !two_pass || it >= p.n_tiles;
      const int j = two_pass ? (it % p.n_tiles) : it;
      const int kv_valid = min(p.bkv, p.seq - j * p.bkv);
      mbar_wait(bar_s_full, it & 1);
      tc_fence_after();
      if (!pv_pass || !two_pass) {
        // row max over this tile
        for (int c = 0; c < p.bkv; c += 16) {
          float s[16];
          tmem_ld16(tmem_s + lane_base + c, s);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c + i < kv_valid) m = fmaxf(m, s[i]);
        }
      }
      if (pv_pass) {
        if (!two_pass || it == p.n_tiles) mc = m * p.scale_log2;
        for (int c = 0; c < p.bkv; c += 16) {
          float s[16];
          tmem_ld16(tmem_s + lane_base + c, s);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float e = (c + i < kv_valid) ? exp2f(fmaf(s[i], p.scale_log2, -mc)) : 0.f;
            s[i] = e;
            l += e;
          }
          // P[r, c..c+15] -> 128B-swizzled K-major tile (atom = 64 keys x 128 rows)
          const uint32_t atom = smem_u32(sP) + (c >> 6) * (ATT_BM * 128) + r * 128;
          const int ch = (c & 63) >> 3;
          st_shared_16(atom + (((ch) ^ (r & 7)) << 4), s);
          st_shared_16(atom + (((ch + 1) ^ (r & 7)) << 4), s + 8);
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(bar_p_ready);
      } else {
        tc_fence_before();
      }
      mbar_arrive(bar_s_free);
    }
    // ---- epilogue: O / l
    const int n_pv = p.n_tiles;
    mbar_wait(bar_pv_done, (n_pv - 1) & 1);
    tc_fence_after();
    const int q = q0 + r;
    const bool valid = q < p.seq;
    const float inv_l = 1.f / l;
    const int img = bh / p.heads, h = bh - img * p.heads;
    __half* dst = p.out + (static_cast<size_t>(img) * p.seq + q) * p.ldo + h * p.dhead;
    for (int c = 0; c < p.n_o; c += 16) {
      float o[16];
      tmem_ld16(tmem_o + lane_base + c, o);
      tmem_ld_wait();
      if (!valid) continue;
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] *= inv_l;
      for (int i = 0; i < 16 && c + i < p.dhead; i += 8) {
        __half2 h0 = __floats2half2_rn(o[i], o[i + 1]);
        __half2 h1 = __floats2half2_rn(o[i + 2], o[i + 3]);
        __half2 h2 = __floats2half2_rn(o[i + 4], o[i + 5]);
        __half2 h3 = __floats2half2_rn(o[i + 6], o[i + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + c + i) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_s, p.tmem_cols);
  }
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_attn_self_f16(const void* q, const void* k, const void* vt, void* out, int32_t n_img, int32_t heads,
                                 int32_t seq, int32_t dhead, int32_t dpad, int32_t ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (q == nullptr || k == nullptr || vt == nullptr || out == nullptr) return set_error(MVD_EINVAL, "mvd_attn_self_f16: null pointer");
  if (n_img <= 0 || heads <= 0 || seq <= 0 || dhead <= 0) return set_error(MVD_EINVAL, "mvd_attn_self_f16: bad sizes");
  if ((dhead & 7) || (dpad & 63) || dpad < dhead || dpad > 192)
    return set_error(MVD_EINVAL, "mvd_attn_self_f16: dhead must be a multiple of 8, dpad a multiple of 64 in [dhead, 192]");
  if ((seq & 15) != 0) return set_error(MVD_EINVAL, "mvd_attn_self_f16: seq must be a multiple of 16");
  if ((ldo & 7) != 0 || ldo < heads * dhead) return set_error(MVD_EALIGN, "mvd_attn_self_f16: ldo must be a multiple of 8 and >= heads*dhead");

  AttnParams p{};
  p.seq = seq;
  p.heads = heads;
  p.dhead = dhead;
  p.dpad = dpad;
  p.bkv = seq >= 128 ? 128 : seq;  // 16 <= bkv <= 128, multiple of 16
  p.n_tiles = (seq + p.bkv - 1) / p.bkv;
  p.kd_steps = (dhead + 15) / 16;
  p.n_o = p.kd_steps * 16;
  int cols = p.bkv + p.n_o;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dhead));
  p.out = static_cast<__half*>(out);
  p.ldo = ldo;
  const int atoms_kv = (p.bkv + 63) / 64;
  auto up1k = [](int x) { return (x + 1023) & ~1023; };
  p.off_k = up1k(ATT_BM * dpad * 2);
  p.off_v = p.off_k + up1k(p.bkv * dpad * 2);
  p.off_p = p.off_v + up1k(atoms_kv * dpad * 128);
  p.off_bar = p.off_p + atoms_kv * ATT_BM * 128;
  const int smem_bytes = p.off_bar + 64 + 1024;

  const int BH = n_img * heads;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_3d(&tmQ, q, dpad, seq, BH, dpad, static_cast<long long>(seq) * dpad, 64, ATT_BM, 1);
  if (rc != MVD_OK) return rc;
  rc = make_tmap_3d(&tmK, k, dpad, seq, BH, dpad, static_cast<long long>(seq) * dpad, 64, p.bkv, 1);
  if (rc != MVD_OK) return rc;
  rc = make_tmap_3d(&tmV, vt, seq, dpad, BH, seq, static_cast<long long>(seq) * dpad, 64, dpad, 1);
  if (rc != MVD_OK) return rc;

  static int configured_smem = 0;
  if (smem_bytes > configured_smem) {
    MVD_CUDA_CHECK(cudaFuncSetAttribute(attn_self_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured_smem = smem_bytes;
  }
  dim3 grid((seq + ATT_BM - 1) / ATT_BM, BH);
  attn_self_kernel<<<grid, ATT_THREADS, smem_bytes, stream>>>(tmQ, tmK, tmV, p);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}
