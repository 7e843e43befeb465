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
// One CTA = 128 query rows of one (image, head); two CTAs share an SM so that one CTA's softmax overlaps the other's
// tensor work.  Warps 0-3: one thread per query row (TMEM lane); warp 4: tcgen05.mma issue; warp 5: TMA producer
// (Q once, K / V^T tiles through a 2-stage ring).  S = Q K^T lives in TMEM (BKV fp32 columns), O accumulates next to it.
// Single pass over the keys with an exact online softmax: the running row maximum is only advanced (and O / l rescaled,
// tcgen05.ld -> scale -> tcgen05.st) when it grows by more than 2^8 — P = exp2(S c - m c) stays far inside fp16 range
// and the final O / l is exact for any reference maximum.  P goes to the tensor core as an fp16 128B-swizzled smem tile.
#include "common.h"
#include "ptx.cuh"

namespace mvd {

constexpr int ATT_BM = 128;
constexpr int ATT_THREADS = 192;
constexpr int ATT_MAX_STAGES = 2;

struct AttnParams {
  int seq, heads, dhead, dpad, bkv;
  int seq_valid;    // keys >= seq_valid are masked out (rows / keys [seq_valid, seq) are layout padding: CLIP's 257 tokens in 272 rows)
  int n_tiles;      // key tiles
  int stages;       // K / V ring depth
  int kd_steps;     // ceil(dhead/16): k-steps of the QK^T product
  int n_o;          // kd_steps*16: columns of O
  int tmem_cols;    // power of two >= bkv + n_o
  float scale_log2; // dhead^-0.5 * log2(e)
  __half* out;
  int ldo;
  // smem byte offsets (from the 1024-aligned base)
  int off_k, off_v, off_p, off_bar;
  int k_stage, v_stage;  // bytes per ring slot
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

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// BKV = key-tile width (16 / 32 / 64 / 128): a whole score row of the tile lives in registers between the maximum
// and the exponentiation, so S is read from TMEM once.
template <int BKV>
__global__ void __launch_bounds__(ATT_THREADS, 2)
    attn_self_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + p.off_k;
  uint8_t* sV = smem + p.off_v;
  uint8_t* sP = smem + p.off_p;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* bar_q = bars + 0;          // TMA Q landed
  uint64_t* bar_k_full = bars + 1;     // [2] K tile landed
  uint64_t* bar_k_empty = bars + 3;    // [2] QK^T MMA that read the slot retired (early: K(j+2) loads under softmax(j))
  uint64_t* bar_v_full = bars + 5;     // [2] V^T tile landed
  uint64_t* bar_v_empty = bars + 7;    // [2] PV MMA that read the slot retired
  uint64_t* bar_s_full = bars + 9;     // QK^T MMA retired
  uint64_t* bar_s_free = bars + 10;    // 128 row threads pulled S into registers
  uint64_t* bar_p_ready = bars + 11;   // 128 row threads wrote P (and rescaled O)
  uint64_t* bar_pv_done = bars + 12;   // PV MMA retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int bh = blockIdx.y;
  const int atoms_d = p.dpad / 64;         // 64-wide k atoms of Q / K tiles
  const int atoms_kv = (BKV + 63) / 64;  // 64-key atoms of the P / Vt tiles

  pdl_trigger();
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023) != 0) __trap();  // swizzled tiles need the 1024-byte base
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    for (int s = 0; s < ATT_MAX_STAGES; ++s) {
      mbar_init(&bar_k_full[s], 1);
      mbar_init(&bar_k_empty[s], 1);
      mbar_init(&bar_v_full[s], 1);
      mbar_init(&bar_v_empty[s], 1);
    }
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
  const uint32_t tmem_o = tmem_s + BKV;

  if (warp == 5) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      pdl_wait();  // q / k / v^T come from the QKV GEMM just before us
      const uint32_t q_bytes = ATT_BM * p.dpad * 2;
      const uint32_t k_bytes = BKV * p.dpad * 2;
      const uint32_t v_bytes = atoms_kv * p.dpad * 128;
      mbar_expect_tx(bar_q, q_bytes);
      for (int a = 0; a < atoms_d; ++a) tma_load_3d(sQ + a * (ATT_BM * 128), &tmQ, bar_q, a * 64, q0, bh);
      for (int j = 0; j < p.n_tiles; ++j) {
        const int st = j % p.stages;
        const uint32_t ph = ((j / p.stages) & 1) ^ 1;
        uint8_t* k = sK + st * p.k_stage;
        uint8_t* v = sV + st * p.v_stage;
        mbar_wait(&bar_k_empty[st], ph);  // released as soon as QK^T(j - stages) retired
        mbar_expect_tx(&bar_k_full[st], k_bytes);
        for (int a = 0; a < atoms_d; ++a) tma_load_3d(k + a * (BKV * 128), &tmK, &bar_k_full[st], a * 64, j * BKV, bh);
        mbar_wait(&bar_v_empty[st], ph);  // released when PV(j - stages) retired
        mbar_expect_tx(&bar_v_full[st], v_bytes);
        for (int a = 0; a < atoms_kv; ++a) tma_load_3d(v + a * (p.dpad * 128), &tmV, &bar_v_full[st], j * BKV + a * 64, 0, bh);
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      // ---------------------------------------------------------------- UMMA issuer
      const uint32_t idesc_s = umma_idesc_f16(ATT_BM, BKV);
      const uint32_t idesc_o = umma_idesc_f16(ATT_BM, p.n_o);
      mbar_wait(bar_q, 0);
      // QK^T of tile j+1 is issued as soon as every row thread has pulled S(j) into registers, i.e. it runs under the
      // softmax of tile j; PV(j) follows when P(j) is in shared memory.
      auto issue_qk = [&](int j) {
        const int st = j % p.stages;
        mbar_wait(&bar_k_full[st], (j / p.stages) & 1);
        if (j > 0) mbar_wait(bar_s_free, (j - 1) & 1);
        tc_fence_after();
        const uint32_t k = smem_u32(sK + st * p.k_stage);
        for (int ks = 0; ks < p.kd_steps; ++ks) {
          const uint32_t aq = smem_u32(sQ) + (ks >> 2) * (ATT_BM * 128) + (ks & 3) * 32;
          const uint32_t ak = k + (ks >> 2) * (BKV * 128) + (ks & 3) * 32;
          umma_f16(tmem_s, umma_desc_sw128(aq), umma_desc_sw128(ak), idesc_s, ks > 0 ? 1u : 0u);
        }
        tc_commit(&bar_k_empty[st]);
        tc_commit(bar_s_full);
      };
      issue_qk(0);
      for (int j = 0; j < p.n_tiles; ++j) {
        const int st = j % p.stages;
        if (j + 1 < p.n_tiles) issue_qk(j + 1);
        const uint32_t v = smem_u32(sV + st * p.v_stage);
        mbar_wait(&bar_v_full[st], (j / p.stages) & 1);
        mbar_wait(bar_p_ready, j & 1);
        tc_fence_after();
        const int ksteps = (min(BKV, p.seq - j * BKV) + 15) / 16;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t ap = smem_u32(sP) + (ks >> 2) * (ATT_BM * 128) + (ks & 3) * 32;
          const uint32_t av = v + (ks >> 2) * (p.dpad * 128) + (ks & 3) * 32;
          umma_f16(tmem_o, umma_desc_sw128(ap), umma_desc_sw128(av), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&bar_v_empty[st]);
        tc_commit(bar_pv_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ one thread per query row
    pdl_wait();  // `out` may alias a buffer the predecessor is still reading
    const int r = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    float m = -INFINITY;  // reference maximum of the raw scores (may lag the true running maximum by <= 8 / scale_log2)
    float l = 0.f;
    constexpr int LDW = BKV >= 32 ? 32 : 16;  // columns per tcgen05.ld
    for (int j = 0; j < p.n_tiles; ++j) {
      const int kv_valid = min(BKV, p.seq_valid - j * BKV);
      mbar_wait(bar_s_full, j & 1);
      tc_fence_after();
      float s[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += LDW) {
        if (LDW == 32) tmem_ld32(tmem_s + lane_base + c, s + c);
        else tmem_ld16(tmem_s + lane_base + c, s + c);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_s_free);  // S is in registers: the next QK^T may overwrite it
      if (kv_valid < BKV) {     // ragged last tile (seq is a multiple of 16)
#pragma unroll
        for (int i = 0; i < BKV; ++i)
          if (i >= kv_valid) s[i] = -INFINITY;
      }
      // row maximum as four independent chains (one chain of BKV / 2 dependent maxima is ~300 cycles of latency per tile that the
      // two warps per scheduler cannot cover)
      float mt;
      if (BKV >= 16) {
        float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
        for (int i = 4; i < BKV; i += 4) {
          m0 = fmaxf(m0, s[i]);
          m1 = fmaxf(m1, s[i + 1]);
          m2 = fmaxf(m2, s[i + 2]);
          m3 = fmaxf(m3, s[i + 3]);
        }
        mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
        mt = s[0];
#pragma unroll
        for (int i = 1; i < BKV; ++i) mt = fmaxf(mt, s[i]);
      }
      float f = 1.f;
      bool rescale = false;
      if (j == 0) {
        m = mt;
      } else {
        const bool grow = (mt - m) * p.scale_log2 > 8.f;
        rescale = __any_sync(0xffffffffu, grow);
        if (rescale) {
          const float m_new = fmaxf(m, mt);
          f = exp2f((m - m_new) * p.scale_log2);
          m = m_new;
          l *= f;
        }
      }
      const float mc = m * p.scale_log2;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < BKV; c += 8) {  // exponentials in place, under the PV MMA of the previous tile
#pragma unroll
        for (int i = 0; i < 8; ++i) s[c + i] = ex2_fast(fmaf(s[c + i], p.scale_log2, -mc));
        l0 += (s[c] + s[c + 1]) + (s[c + 2] + s[c + 3]);
        l1 += (s[c + 4] + s[c + 5]) + (s[c + 6] + s[c + 7]);
      }
      l += l0 + l1;
      if (j > 0) {
        mbar_wait(bar_pv_done, (j - 1) & 1);  // P tile free again, O quiescent
        tc_fence_after();
        if (rescale) {
          for (int c = 0; c < p.n_o; c += 16) {
            float o[16];
            tmem_ld16(tmem_o + lane_base + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] *= f;
            tmem_st16(tmem_o + lane_base + c, o);
          }
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int c = 0; c < BKV; c += 8) {
        // P[r, c..c+7] -> 128B-swizzled K-major tile (atom = 64 keys x 128 rows)
        const uint32_t atom = smem_u32(sP) + (c >> 6) * (ATT_BM * 128) + r * 128;
        st_shared_16(atom + ((((c & 63) >> 3) ^ (r & 7)) << 4), s + c);
      }
      tc_fence_before();
      fence_async_smem();
      mbar_arrive(bar_p_ready);
    }
    // ---- epilogue: O / l
    mbar_wait(bar_pv_done, (p.n_tiles - 1) & 1);
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
#pragma unroll
      for (int i = 0; i < 16; i += 8) {
        if (c + i >= p.dhead) continue;
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
  return mvd_attn_self_masked_f16(q, k, vt, out, n_img, heads, seq, seq, dhead, dpad, ldo, stream_);
}

extern "C" int mvd_attn_self_masked_f16(const void* q, const void* k, const void* vt, void* out, int32_t n_img, int32_t heads,
                                        int32_t seq, int32_t seq_valid, int32_t dhead, int32_t dpad, int32_t ldo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (seq_valid <= 0 || seq_valid > seq) return set_error(MVD_EINVAL, "mvd_attn_self_masked_f16: seq_valid must be in [1, seq]");
  if (q == nullptr || k == nullptr || vt == nullptr || out == nullptr) return set_error(MVD_EINVAL, "mvd_attn_self_f16: null pointer");
  if (n_img <= 0 || heads <= 0 || seq <= 0 || dhead <= 0) return set_error(MVD_EINVAL, "mvd_attn_self_f16: bad sizes");
  if ((dhead & 7) || (dpad & 63) || dpad < dhead || dpad > 192)
    return set_error(MVD_EINVAL, "mvd_attn_self_f16: dhead must be a multiple of 8, dpad a multiple of 64 in [dhead, 192]");
  if ((seq & 15) != 0) return set_error(MVD_EINVAL, "mvd_attn_self_f16: seq must be a multiple of 16");
  if ((ldo & 7) != 0 || ldo < heads * dhead) return set_error(MVD_EALIGN, "mvd_attn_self_f16: ldo must be a multiple of 8 and >= heads*dhead");

  AttnParams p{};
  p.seq = seq;
  p.seq_valid = seq_valid;
  p.heads = heads;
  p.dhead = dhead;
  p.dpad = dpad;
  p.kd_steps = (dhead + 15) / 16;
  p.n_o = p.kd_steps * 16;
  // key tile: the widest of {128, 64} (or the whole sequence) whose footprint lets two CTAs share an SM
  // (<= 113 KB of shared memory and <= 256 TMEM columns each); otherwise the widest that fits at all.
  auto up1k = [](int x) { return (x + 1023) & ~1023; };
  auto layout = [&](int bkv) {
    p.bkv = bkv;
    p.n_tiles = (seq_valid + bkv - 1) / bkv;  // fully masked key tiles are never visited
    p.stages = p.n_tiles > 1 ? ATT_MAX_STAGES : 1;
    const int atoms_kv = (bkv + 63) / 64;
    p.k_stage = up1k(bkv * dpad * 2);
    p.v_stage = up1k(atoms_kv * dpad * 128);
    p.off_k = up1k(ATT_BM * dpad * 2);
    p.off_v = p.off_k + p.stages * p.k_stage;
    p.off_p = p.off_v + p.stages * p.v_stage;
    p.off_bar = p.off_p + atoms_kv * ATT_BM * 128;
    return p.off_bar + 128;  // 13 barriers + the TMEM slot
  };
  int smem_bytes = 0;
  {
    const int cands[4] = {128, 64, 32, 16};
    int chosen = -1;
    // a sequence of one or two 32-key tiles runs as two tiles, so that softmax(0) overlaps QK^T(1) (seq 64: 6.5 -> 5.3 us on B200)
    if (seq == 64 && layout(32) <= 115712 && 32 + p.n_o <= 256) chosen = 32;
    for (int i = 0; i < 4 && chosen < 0; ++i)
      if (cands[i] <= seq && layout(cands[i]) <= 115712 && cands[i] + p.n_o <= 256) chosen = cands[i];
    for (int i = 0; i < 4 && chosen < 0; ++i)
      if (cands[i] <= seq && layout(cands[i]) <= 232448 && cands[i] + p.n_o <= 512) chosen = cands[i];
    if (chosen < 0) return set_error(MVD_EINVAL, "mvd_attn_self_f16: tile does not fit in shared memory");
    smem_bytes = layout(chosen);
  }
  int cols = p.bkv + p.n_o;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dhead));
  p.out = static_cast<__half*>(out);
  p.ldo = ldo;

  const int BH = n_img * heads;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_3d(&tmQ, q, dpad, seq, BH, dpad, static_cast<long long>(seq) * dpad, 64, ATT_BM, 1);
  if (rc != MVD_OK) return rc;
  rc = make_tmap_3d(&tmK, k, dpad, seq, BH, dpad, static_cast<long long>(seq) * dpad, 64, p.bkv, 1);
  if (rc != MVD_OK) return rc;
  rc = make_tmap_3d(&tmV, vt, seq, dpad, BH, seq, static_cast<long long>(seq) * dpad, 64, dpad, 1);
  if (rc != MVD_OK) return rc;

  typedef void (*AttnFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);
  AttnFn fn = p.bkv == 128 ? attn_self_kernel<128> : p.bkv == 64 ? attn_self_kernel<64> : p.bkv == 32 ? attn_self_kernel<32> : attn_self_kernel<16>;
  static bool configured = false;
  if (!configured) {
    AttnFn all[] = {attn_self_kernel<128>, attn_self_kernel<64>, attn_self_kernel<32>, attn_self_kernel<16>};
    for (AttnFn f : all) MVD_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    configured = true;
  }
  dim3 grid((seq + ATT_BM - 1) / ATT_BM, BH);
  MVD_CUDA_CHECK(launch_kernel(fn, grid, dim3(ATT_THREADS), static_cast<size_t>(smem_bytes), stream, 1, tmQ, tmK, tmV, p));
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}
