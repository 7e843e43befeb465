// GridAttn aggregation transformer as one persistent kernel for sm_100a (SURVEY.md kernel K10).
//
// Reference: mvdfusion/view_attn_efficient2.py:365-410 (aggregate_features tail): pre_layer_b (Linear + GELU) on the 723-d
// tokens, `num_layers` adaLN-Zero DiT blocks (:45-70) whose attention runs over the V views of one 3-D point, and the
// weight_layer view-softmax pooling (:83-97).  The unfused program (engine.emit_gridattn) runs this as 1 + 7 * layers + 1
// launches around an fp32 [P*V, 256] residual stream (67 MB at N = 8 views, 256^2): every GEMM / LayerNorm of it is bound by
// reading and writing that stream.  The blocks only ever mix the V rows of one point, so a tile of 128 rows (128 / V points) is
// independent of all others: here one CTA per SM walks over such tiles and keeps everything on chip.
//
//   TMEM   columns [0, 256)   X    the fp32 residual stream of the tile (row = TMEM lane).  pre_layer_b, attn.proj and mlp.fc2
//                                  ACCUMULATE straight into it (the adaLN gates are folded into proj / fc2: mvd_dit_fold_gates)
//          columns [256, 512) ACC  two 128-column accumulators: per-head q | k | v (96 columns) or an fc1 quarter (128 columns)
//   smem   A tile  64 KB  LayerNorm-modulate output, fp16, K-major 128B-swizzled (4 k-blocks) — the A operand of qkv / fc1;
//                         during pre_layer_b it is the 4-slot ring of token k-blocks (TMA)
//          B tile  64 KB  attention output (A operand of proj), then the two fc1-GELU quarter buffers (A operand of fc2)
//          KV      32 KB  k | v rows of the head a warpgroup is working on (fp16)
//          W ring  64 KB  4 x 16 KB slots of weight k-blocks (TMA; 128 rows x 64 k, or 96 rows for a head's q | k | v)
//   warps  0: TMA producer   1: tcgen05.mma issuer (+ TMEM allocation)   2-9: two warpgroups, thread = (warpgroup, tile row)
//          LayerNorm / GELU passes split the columns between the warpgroups; attention alternates heads (even / odd).
//
// Per tile: 12 + layers * (32 + 4 + 16 + 8) weight k-block loads (3.1 MB at 3 layers, L2-resident) against 128 x 256 resident
// activations; the only global traffic besides the weights is the token tile in and 128 / V pooled rows out.
#include <cstring>

#include "common.h"
#include "ptx.cuh"

namespace mvd {

constexpr int DT_C = 256, DT_HID = 512, DT_HEADS = 8, DT_HD = 32, DT_BM = 128;
constexpr int DT_THREADS = 320;  // 10 warps: 65536 / 320 = 204 registers per thread (the attention phase wants ~180)
constexpr int DT_MAXL = 4;
constexpr int DT_MAXV = 16;
constexpr int DT_SLOT = 16384;
constexpr int DT_OFF_A = 0, DT_OFF_B = 65536, DT_OFF_KV = 131072, DT_OFF_W = 163840, DT_OFF_MISC = 229376;
constexpr int DT_SMEM = DT_OFF_MISC + 3072;

struct DitMaps {
  CUtensorMap tok, wpre;
  CUtensorMap qkv[DT_MAXL], proj[DT_MAXL], fc1[DT_MAXL], fc2[DT_MAXL];
};
struct DitLayerP {
  const float *b_qkv, *b_proj, *b_fc1, *b_fc2, *shift_msa, *scale_msa, *shift_mlp, *scale_mlp;
};
struct DitParams {
  int R, V, layers, nkb_pre, n_tiles;
  const float* b_pre;
  DitLayerP L[DT_MAXL];
  const float *pool_w, *pool_b;
  __half* pooled;
  float* x_out;
  float eps;
};

// barrier indices inside the misc block
enum {
  BAR_TOK_FULL = 0, BAR_TOK_EMPTY = 4, BAR_W_FULL = 8, BAR_W_EMPTY = 12, BAR_X_READY = 16, BAR_A_READY = 17, BAR_A_FREE = 18,
  BAR_QACC_FULL = 19, BAR_QACC_FREE = 21, BAR_B_READY = 23, BAR_X_DONE = 24, BAR_FACC_FULL = 25, BAR_FACC_FREE = 27,
  BAR_F_READY = 29, BAR_F_FREE = 31, BAR_TILE_DONE = 33, BAR_COUNT = 34
};

// Use counters of ring slots / double buffers as bit fields (bit i = parity of the number of uses of slot i, `used` bit i = it has
// been used at all): the roles index them with run-time slot numbers, and arrays would live in local memory.
struct Uses {
  uint32_t par, used;
  __device__ __forceinline__ void init() { par = 0; used = 0; }
  // consumer side: parity to wait for on the slot's "full" barrier, then count the use
  __device__ __forceinline__ uint32_t full_parity(int s) {
    const uint32_t ph = (par >> s) & 1u;
    par ^= 1u << s;
    return ph;
  }
  // producer side: true (and the parity of the "empty" barrier to wait for) when the slot has been used before; counts the use
  __device__ __forceinline__ bool empty_parity(int s, uint32_t& ph) {
    const bool again = (used >> s) & 1u;
    ph = ((par >> s) & 1u) ^ 1u;
    par ^= 1u << s;
    used |= 1u << s;
    return again;
  }
};
// weight ring cursor: producer and MMA issuer walk the same sequence of single-slot / slot-pair uses
struct WRing {
  int pos;
  Uses u;
  __device__ __forceinline__ void init() { pos = 0; u.init(); }
  __device__ __forceinline__ int take() { const int s = pos; pos = (pos + 1) & 3; return s; }
  __device__ __forceinline__ void align_pair() { pos = (pos + 1) & 2; }  // next even slot (0 or 2)
};

__device__ __forceinline__ void dt_mma_kblock(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool acc_first) {
  const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(b_addr);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (acc_first || k > 0) ? 1u : 0u);
}
__device__ __forceinline__ uint32_t dt_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void dt_sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 dt_lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void dt_dot2(float& acc, uint32_t a2, uint32_t b2) {  // acc += a.lo b.lo + a.hi b.hi (fp16 products exact, fp32 sum)
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.f16 %0, al, bl, %0;\n\tfma.rn.f32.f16 %0, ah, bh, %0;\n\t}\n" : "+f"(acc) : "r"(a2), "r"(b2));
}
__device__ __forceinline__ void dt_axpy2(float& acc0, float& acc1, uint32_t p2, uint32_t v2) {  // acc0 += p v.lo ; acc1 += p v.hi
  asm("{\n\t.reg .b16 pl, ph, vl, vh;\n\tmov.b32 {pl, ph}, %2;\n\tmov.b32 {vl, vh}, %3;\n\t"
      "fma.rn.f32.f16 %0, pl, vl, %0;\n\tfma.rn.f32.f16 %1, ph, vh, %1;\n\t}\n" : "+f"(acc0), "+f"(acc1) : "r"(p2), "r"(v2));
}

// optional timeline (-DMVD_GEMM_TRACE, `make trace`): SM clock at the phase boundaries of the first tile, per CTA, as seen by row 0 of
// warpgroup 0 (tools/dit_trace.py)
#ifdef MVD_GEMM_TRACE
__device__ unsigned int g_dit_trace[160 * 64];
#define DT_TR(i) do { if (it == 0 && wg == 0 && r == 0 && (i) < 64) g_dit_trace[blockIdx.x * 64 + (i)] = static_cast<unsigned int>(clock64()); } while (0)
#else
#define DT_TR(i) do { } while (0)
#endif

// VT = views per point (compile time: the attention loops unroll exactly, without predicated tails)
template <int VT>
__global__ void __launch_bounds__(DT_THREADS, 1) dit_kernel(const __grid_constant__ DitMaps maps, const DitParams p) {
  extern __shared__ __align__(1024) uint8_t dt_smem[];
  uint8_t* smem = dt_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DT_OFF_MISC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT + 2);
  float* xch = reinterpret_cast<float*>(smem + DT_OFF_MISC + 512);  // [2][128] warpgroup exchange (LayerNorm sums, pooling scores)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.layers;
  pdl_trigger();
  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023) != 0) __trap();  // swizzled operand tiles need the 1024-byte base
    tma_prefetch_desc(&maps.tok);
    tma_prefetch_desc(&maps.wpre);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars[BAR_TOK_FULL + i], 1);
      mbar_init(&bars[BAR_TOK_EMPTY + i], 1);
      mbar_init(&bars[BAR_W_FULL + i], 1);
      mbar_init(&bars[BAR_W_EMPTY + i], 1);
    }
    mbar_init(&bars[BAR_X_READY], 1);
    mbar_init(&bars[BAR_A_READY], 256);
    mbar_init(&bars[BAR_A_FREE], 1);
    mbar_init(&bars[BAR_B_READY], 256);
    mbar_init(&bars[BAR_X_DONE], 1);
    mbar_init(&bars[BAR_TILE_DONE], 256);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[BAR_QACC_FULL + i], 1);
      mbar_init(&bars[BAR_QACC_FREE + i], 128);
      mbar_init(&bars[BAR_FACC_FULL + i], 1);
      mbar_init(&bars[BAR_FACC_FREE + i], 256);
      mbar_init(&bars[BAR_F_READY + i], 256);
      mbar_init(&bars[BAR_F_FREE + i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_x = *tmem_slot;
  const uint32_t tmem_acc = tmem_x + 256;
  const uint32_t sA = smem_u32(smem + DT_OFF_A), sB = smem_u32(smem + DT_OFF_B), sKV = smem_u32(smem + DT_OFF_KV), sW = smem_u32(smem + DT_OFF_W);
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = first < p.n_tiles ? (p.n_tiles - first + stride - 1) / stride : 0;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      WRing ring;
      ring.init();
      Uses tok;
      tok.init();
      auto w_load = [&](const CUtensorMap* m, int kcol, int row, uint32_t bytes) {
        const int s = ring.take();
        uint32_t ph;
        if (ring.u.empty_parity(s, ph)) mbar_wait(&bars[BAR_W_EMPTY + s], ph);
        mbar_expect_tx(&bars[BAR_W_FULL + s], bytes);
        tma_load_2d(smem + DT_OFF_W + s * DT_SLOT, m, &bars[BAR_W_FULL + s], kcol, row);
      };
      auto w_pair = [&](const CUtensorMap* m, int kcol) {
        ring.align_pair();
        w_load(m, kcol, 0, DT_SLOT);
        w_load(m, kcol, 128, DT_SLOT);
      };
      uint32_t n_afree = 0;  // a_free completions this role has stepped over
      for (int it = 0; it < n_local; ++it) {
        const int tile = first + it * stride;
        if (it == 0) pdl_wait();  // the tokens come from the kernel before us (the weights / folded gates too)
        // the token ring lives in the A tile: the previous tile's last LayerNorm output must have been consumed
        if (it > 0) {
          // a_free completes twice per block (after the q | k | v products, after the last fc1 quarter): wait for the last one of the
          // previous tile.  (A parity wait cannot tell completion m from m - 2; the 4-slot weight ring keeps this role within a few
          // k-blocks of the issuer, which by then is past the last block's q | k | v products.)
          n_afree = static_cast<uint32_t>(it) * 2u * L;
          mbar_wait(&bars[BAR_A_FREE], (n_afree - 1) & 1);
        }
        for (int kb = 0; kb < p.nkb_pre; ++kb) {
          const int s = kb & 3;
          uint32_t ph;
          if (tok.empty_parity(s, ph)) mbar_wait(&bars[BAR_TOK_EMPTY + s], ph);
          mbar_expect_tx(&bars[BAR_TOK_FULL + s], DT_SLOT);
          tma_load_2d(smem + DT_OFF_A + s * DT_SLOT, &maps.tok, &bars[BAR_TOK_FULL + s], kb * 64, tile * DT_BM);
          w_pair(&maps.wpre, kb * 64);
        }
        for (int l = 0; l < L; ++l) {
          for (int h = 0; h < DT_HEADS; ++h)
            for (int kb = 0; kb < 4; ++kb) w_load(&maps.qkv[l], kb * 64, h * 96, 96 * 128);
          for (int kb = 0; kb < 4; ++kb) w_pair(&maps.proj[l], kb * 64);
          // fc1 / fc2 quarters in the issuer's order: fc1(0) fc1(1) fc2(0) fc1(2) fc2(1) fc1(3) fc2(2) fc2(3)
          auto ld_fc1 = [&](int q) {
            for (int kb = 0; kb < 4; ++kb) w_load(&maps.fc1[l], kb * 64, q * 128, DT_SLOT);
          };
          auto ld_fc2 = [&](int q) {
            for (int kk = 0; kk < 2; ++kk) w_pair(&maps.fc2[l], (q * 2 + kk) * 64);
          };
          ld_fc1(0); ld_fc1(1); ld_fc2(0); ld_fc1(2); ld_fc2(1); ld_fc1(3); ld_fc2(2); ld_fc2(3);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ tcgen05.mma issuer
    if (lane == 0) {
      WRing ring;
      ring.init();
      Uses tok, qacc, facc, fbuf;
      tok.init(); qacc.init(); facc.init(); fbuf.init();
      uint32_t n_aready = 0, n_bready = 0, n_tile = 0;
      const uint32_t idesc256 = umma_idesc_f16(DT_BM, 256), idesc128 = umma_idesc_f16(DT_BM, 128), idesc96 = umma_idesc_f16(DT_BM, 96);
      auto w_wait = [&](int& slot) {
        slot = ring.take();
        mbar_wait(&bars[BAR_W_FULL + slot], ring.u.full_parity(slot));
      };
      // one k-block of an N = 256 product: the weight rows [0, 128) and [128, 256) sit in an even / odd slot pair (contiguous)
      auto mma_pair = [&](uint32_t d, uint32_t a_addr, bool acc_first) {
        ring.align_pair();
        int s0, s1;
        w_wait(s0);
        w_wait(s1);
        tc_fence_after();
        dt_mma_kblock(d, a_addr, sW + s0 * DT_SLOT, idesc256, acc_first);
        tc_commit(&bars[BAR_W_EMPTY + s0]);
        tc_commit(&bars[BAR_W_EMPTY + s1]);
      };
      auto mma_single = [&](uint32_t d, uint32_t a_addr, uint32_t idesc, bool acc_first) {
        int s;
        w_wait(s);
        tc_fence_after();
        dt_mma_kblock(d, a_addr, sW + s * DT_SLOT, idesc, acc_first);
        tc_commit(&bars[BAR_W_EMPTY + s]);
      };
      for (int it = 0; it < n_local; ++it) {
        if (it > 0) {  // the warpgroups have read the previous tile out of TMEM
          mbar_wait(&bars[BAR_TILE_DONE], (n_tile - 1) & 1);
          tc_fence_after();
        }
        ++n_tile;
        // ---- pre_layer_b: X = tokens W_pre^T
        for (int kb = 0; kb < p.nkb_pre; ++kb) {
          const int s = kb & 3;
          mbar_wait(&bars[BAR_TOK_FULL + s], tok.full_parity(s));
          mma_pair(tmem_x, sA + s * DT_SLOT, kb > 0);
          tc_commit(&bars[BAR_TOK_EMPTY + s]);
        }
        tc_commit(&bars[BAR_X_READY]);
        for (int l = 0; l < L; ++l) {
          // ---- per-head q | k | v = a W_qkv[h]^T  (two accumulators: head h + 1 is issued under the attention of head h)
          mbar_wait(&bars[BAR_A_READY], n_aready & 1);
          ++n_aready;
          tc_fence_after();
          for (int h = 0; h < DT_HEADS; ++h) {
            const int b = h & 1;
            uint32_t ph;
            if (qacc.empty_parity(b, ph)) {
              mbar_wait(&bars[BAR_QACC_FREE + b], ph);
              tc_fence_after();
            }
            for (int kb = 0; kb < 4; ++kb) mma_single(tmem_acc + b * 128, sA + kb * DT_SLOT, idesc96, kb > 0);
            tc_commit(&bars[BAR_QACC_FULL + b]);
          }
          tc_commit(&bars[BAR_A_FREE]);
          // ---- X += attn_out W_proj'^T
          mbar_wait(&bars[BAR_B_READY], n_bready & 1);
          ++n_bready;
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) mma_pair(tmem_x, sB + kb * DT_SLOT, true);
          tc_commit(&bars[BAR_X_DONE]);
          // ---- MLP in hidden quarters: fc1(q) -> ACC[q & 1]; fc2(q): X += gelu(fc1 quarter) W_fc2'[:, quarter]^T
          mbar_wait(&bars[BAR_A_READY], n_aready & 1);
          ++n_aready;
          tc_fence_after();
          auto fc1 = [&](int q) {
            const int b = q & 1;
            uint32_t ph;
            if (facc.empty_parity(b, ph)) {
              mbar_wait(&bars[BAR_FACC_FREE + b], ph);
              tc_fence_after();
            }
            for (int kb = 0; kb < 4; ++kb) mma_single(tmem_acc + b * 128, sA + kb * DT_SLOT, idesc128, kb > 0);
            tc_commit(&bars[BAR_FACC_FULL + b]);
          };
          auto fc2 = [&](int q) {
            const int b = q & 1;
            mbar_wait(&bars[BAR_F_READY + b], fbuf.full_parity(b));
            tc_fence_after();
            for (int kk = 0; kk < 2; ++kk) mma_pair(tmem_x, sB + b * 32768 + kk * DT_SLOT, true);
            tc_commit(&bars[BAR_F_FREE + b]);
          };
          fc1(0);
          fc1(1);
          fc2(0);
          fc1(2);
          fc2(1);
          fc1(3);
          tc_commit(&bars[BAR_A_FREE]);
          fc2(2);
          fc2(3);
          tc_commit(&bars[BAR_X_DONE]);
        }
      }
    }
  } else if (warp >= 2) {
    // ================================================================ two warpgroups: thread = (warpgroup, tile row)
    const int wg = (warp - 2) >> 2;  // warps 2-5 and 6-9: any four consecutive warps cover the four TMEM lane quadrants (warp % 4)
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t xaddr = tmem_x + lane_base + wg * 128;   // this thread's half of its X row
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    constexpr int V = VT;
    const int j0 = r & ~(V - 1);                             // first row of this row's point
    uint32_t n_xready = 0, n_xdone = 0, n_awrite = 0;
    uint32_t qacc_uses = 0;
    Uses facc, fbuf;
    facc.init(); fbuf.init();
    pdl_wait();

    auto exchange2 = [&](float a, float b2, float& ra, float& rb) {  // sums of the two warpgroups' partials of row r (identical bits in both)
      float2* x2 = reinterpret_cast<float2*>(xch);
      x2[wg * 128 + r] = make_float2(a, b2);
      named_bar_sync(3, 256);
      const float2 o = x2[(wg ^ 1) * 128 + r];
      named_bar_sync(3, 256);
      ra = a + o.x;
      rb = b2 + o.y;
    };
    // LayerNorm(no affine) * (1 + scale) + shift of the resident row -> fp16 A tile.  `pend`: a bias still owed to X (the folded
    // proj / fc2 bias of the product that was just accumulated) is added and written back first.  Two passes over this thread's
    // 128 columns, 64 at a time (a TMEM read costs a few hundred cycles of latency that two warps per scheduler cannot hide):
    // sum and sum of squares (fp32 over 256 values of O(1) magnitude), then the normalisation.  have_stats: the caller already
    // holds this thread's partial sums (the pre_layer GELU pass).
    auto ln_modulate = [&](const float* pend, const float* shift, const float* scale, bool have_stats, float s, float q2) {
      if (!have_stats) {
        s = 0.f;
        q2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[64];
          tmem_ld32(xaddr + c * 64, v);
          tmem_ld32(xaddr + c * 64 + 32, v + 32);
          tmem_ld_wait();
          if (pend != nullptr) {
            const float4* pb = reinterpret_cast<const float4*>(pend + wg * 128 + c * 64);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 b4 = __ldg(pb + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
            tmem_st32(xaddr + c * 64, v);
            tmem_st32(xaddr + c * 64 + 32, v + 32);
          }
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            s0 += v[i]; s1 += v[i + 1];
            q0 = fmaf(v[i], v[i], q0); q1 = fmaf(v[i + 1], v[i + 1], q1);
          }
          s += s0 + s1;
          q2 += q0 + q1;
        }
        if (pend != nullptr) tmem_st_wait();
      }
      float ts, tq;
      exchange2(s, q2, ts, tq);
      const float mean = ts * (1.f / DT_C);
      const float rstd = rsqrtf(fmaxf(tq * (1.f / DT_C) - mean * mean, 0.f) + p.eps);
      // the A tile may be rewritten once the products that read the previous one have retired
      if (n_awrite > 0) mbar_wait(&bars[BAR_A_FREE], (n_awrite - 1) & 1);
      ++n_awrite;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[64];
        tmem_ld32(xaddr + c * 64, v);
        tmem_ld32(xaddr + c * 64 + 32, v + 32);
        tmem_ld_wait();
        const float4* sc4 = reinterpret_cast<const float4*>(scale + wg * 128 + c * 64);
        const float4* sh4 = reinterpret_cast<const float4*>(shift + wg * 128 + c * 64);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 a = __ldg(sc4 + i), b = __ldg(sh4 + i);
          v[4 * i] = fmaf((v[4 * i] - mean) * rstd, 1.f + a.x, b.x);
          v[4 * i + 1] = fmaf((v[4 * i + 1] - mean) * rstd, 1.f + a.y, b.y);
          v[4 * i + 2] = fmaf((v[4 * i + 2] - mean) * rstd, 1.f + a.z, b.z);
          v[4 * i + 3] = fmaf((v[4 * i + 3] - mean) * rstd, 1.f + a.w, b.w);
        }
        // 64 columns = k-block (2 wg + c) of the A tile
        const uint32_t rowaddr = sA + (wg * 2 + c) * DT_SLOT + r * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dt_sts16(rowaddr + ((static_cast<uint32_t>(i) ^ sw) << 4), dt_pack(v[8 * i], v[8 * i + 1]), dt_pack(v[8 * i + 2], v[8 * i + 3]),
                   dt_pack(v[8 * i + 4], v[8 * i + 5]), dt_pack(v[8 * i + 6], v[8 * i + 7]));
      }
      tc_fence_before();
      fence_async_smem();
      mbar_arrive(&bars[BAR_A_READY]);
    };

    for (int it = 0; it < n_local; ++it) {
      const int tile = first + it * stride;
      const long long row_g = static_cast<long long>(tile) * DT_BM + r;
      // ---- pre_layer_b epilogue: X = GELU(X + b_pre)
      DT_TR(0);
      mbar_wait(&bars[BAR_X_READY], n_xready & 1);
      ++n_xready;
      tc_fence_after();
      DT_TR(1);
      float pre_s = 0.f, pre_q = 0.f;  // norm1 statistics of block 0 come out of this pass
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[64];
        tmem_ld32(xaddr + c * 64, v);
        tmem_ld32(xaddr + c * 64 + 32, v + 32);
        tmem_ld_wait();
        const float4* pb = reinterpret_cast<const float4*>(p.b_pre + wg * 128 + c * 64);
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b4 = __ldg(pb + i);
          v[4 * i] = gelu_erf(v[4 * i] + b4.x); v[4 * i + 1] = gelu_erf(v[4 * i + 1] + b4.y);
          v[4 * i + 2] = gelu_erf(v[4 * i + 2] + b4.z); v[4 * i + 3] = gelu_erf(v[4 * i + 3] + b4.w);
          s0 += v[4 * i] + v[4 * i + 2]; s1 += v[4 * i + 1] + v[4 * i + 3];
          q0 = fmaf(v[4 * i], v[4 * i], q0); q1 = fmaf(v[4 * i + 1], v[4 * i + 1], q1);
          q0 = fmaf(v[4 * i + 2], v[4 * i + 2], q0); q1 = fmaf(v[4 * i + 3], v[4 * i + 3], q1);
        }
        pre_s += s0 + s1;
        pre_q += q0 + q1;
        tmem_st32(xaddr + c * 64, v);
        tmem_st32(xaddr + c * 64 + 32, v + 32);
      }
      tmem_st_wait();
      DT_TR(2);

      for (int l = 0; l < L; ++l) {
        const DitLayerP& P = p.L[l];
        DT_TR(3 + 8 * l);
        // ---- norm1 + modulate -> A tile  (layers > 0: the previous block's fc2 bias is still owed to X)
        if (l > 0) {
          mbar_wait(&bars[BAR_X_DONE], n_xdone & 1);
          ++n_xdone;
          tc_fence_after();
        }
        DT_TR(4 + 8 * l);
        ln_modulate(l > 0 ? p.L[l - 1].b_fc2 : nullptr, P.shift_msa, P.scale_msa, l == 0, pre_s, pre_q);
        DT_TR(5 + 8 * l);

        // ---- attention over the V views of each point, heads wg, wg + 2, wg + 4, wg + 6
        const uint32_t kv = sKV + wg * 16384;
        const uint32_t kvsw = static_cast<uint32_t>((r >> 3) & 7);
#pragma unroll 1
        for (int hh = 0; hh < 4; ++hh) {
          const int h = 2 * hh + wg;
          mbar_wait(&bars[BAR_QACC_FULL + wg], qacc_uses & 1);
          ++qacc_uses;
          tc_fence_after();
          const uint32_t acc = tmem_acc + lane_base + wg * 128;
          const float4* bq = reinterpret_cast<const float4*>(P.b_qkv + h * 96);
          const uint32_t myrow = kv + r * 128;
          uint32_t qh[16];
          // q, then k, then v pass through the same 32 registers (one accumulator read each)
          {
            float t[32];
            tmem_ld32(acc, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a = __ldg(bq + i);
              qh[2 * i] = dt_pack(t[4 * i] + a.x, t[4 * i + 1] + a.y);
              qh[2 * i + 1] = dt_pack(t[4 * i + 2] + a.z, t[4 * i + 3] + a.w);
            }
          }
          // k | v row -> staging (16-byte chunk index XOR (row / 8): the V rows a warp reads together sit 8 or 16 rows apart)
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            float t[32];
            tmem_ld32(acc + 32 + part * 32, t);
            tmem_ld_wait();
            if (part == 1) {
              tc_fence_before();
              mbar_arrive(&bars[BAR_QACC_FREE + wg]);  // last accumulator read of this head
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a = __ldg(bq + 8 + part * 8 + i);
              t[4 * i] += a.x; t[4 * i + 1] += a.y; t[4 * i + 2] += a.z; t[4 * i + 3] += a.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dt_sts16(myrow + ((static_cast<uint32_t>(part * 4 + i) ^ kvsw) << 4), dt_pack(t[8 * i], t[8 * i + 1]), dt_pack(t[8 * i + 2], t[8 * i + 3]),
                       dt_pack(t[8 * i + 4], t[8 * i + 5]), dt_pack(t[8 * i + 6], t[8 * i + 7]));
          }
          named_bar_sync(1 + wg, 128);
          const float scale_log2 = rsqrtf(static_cast<float>(DT_HD)) * 1.4426950408889634f;
          // scores of this row against the V rows of its point (k rows are broadcast reads); four independent accumulators
          float sc[VT];
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < VT; ++j) {
            {
              const int jr = j0 + j;
              const uint32_t ja = kv + jr * 128;
              const uint32_t jsw = static_cast<uint32_t>((jr >> 3) & 7);
              float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
              const uint4 u0 = dt_lds16(ja + ((0u ^ jsw) << 4)), u1 = dt_lds16(ja + ((1u ^ jsw) << 4));
              const uint4 u2 = dt_lds16(ja + ((2u ^ jsw) << 4)), u3 = dt_lds16(ja + ((3u ^ jsw) << 4));
              dt_dot2(a0, qh[0], u0.x); dt_dot2(a1, qh[1], u0.y); dt_dot2(a2, qh[2], u0.z); dt_dot2(a3, qh[3], u0.w);
              dt_dot2(a0, qh[4], u1.x); dt_dot2(a1, qh[5], u1.y); dt_dot2(a2, qh[6], u1.z); dt_dot2(a3, qh[7], u1.w);
              dt_dot2(a0, qh[8], u2.x); dt_dot2(a1, qh[9], u2.y); dt_dot2(a2, qh[10], u2.z); dt_dot2(a3, qh[11], u2.w);
              dt_dot2(a0, qh[12], u3.x); dt_dot2(a1, qh[13], u3.y); dt_dot2(a2, qh[14], u3.z); dt_dot2(a3, qh[15], u3.w);
              sc[j] = ((a0 + a1) + (a2 + a3)) * scale_log2;
              m = fmaxf(m, sc[j]);
            }
          }
          // P = exp2(s - m) rounded to fp16 (as the tensor-core attention does; l sums the rounded values), kept packed in sc[]
          float lsum = 0.f;
#pragma unroll
          for (int j = 0; j < VT; ++j) {
            {
              float pe;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pe) : "f"(sc[j] - m));
              const __half ph = __float2half_rn(pe);
              lsum += __half2float(ph);
              sc[j] = __uint_as_float(static_cast<uint32_t>(__half_as_ushort(ph)) * 0x10001u);
            }
          }
          const float inv = 1.f / lsum;
          // P V in two halves of 16 dims (16 accumulators live at a time); head h -> columns [32h, 32h + 32) of the B tile
          // (k-block h / 2, chunks 4 (h & 1) .. + 3)
          const uint32_t brow = sB + (h >> 1) * DT_SLOT + r * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = 0.f;
#pragma unroll
            for (int j = 0; j < VT; ++j) {
              {
                const int jr = j0 + j;
                const uint32_t ja = kv + jr * 128;
                const uint32_t jsw = static_cast<uint32_t>((jr >> 3) & 7);
                const uint32_t p2 = __float_as_uint(sc[j]);
                const uint4 u0 = dt_lds16(ja + ((static_cast<uint32_t>(4 + 2 * half) ^ jsw) << 4));
                const uint4 u1 = dt_lds16(ja + ((static_cast<uint32_t>(5 + 2 * half) ^ jsw) << 4));
                dt_axpy2(o[0], o[1], p2, u0.x); dt_axpy2(o[2], o[3], p2, u0.y); dt_axpy2(o[4], o[5], p2, u0.z); dt_axpy2(o[6], o[7], p2, u0.w);
                dt_axpy2(o[8], o[9], p2, u1.x); dt_axpy2(o[10], o[11], p2, u1.y); dt_axpy2(o[12], o[13], p2, u1.z); dt_axpy2(o[14], o[15], p2, u1.w);
              }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
              dt_sts16(brow + ((static_cast<uint32_t>((h & 1) * 4 + 2 * half + i) ^ sw) << 4), dt_pack(o[8 * i] * inv, o[8 * i + 1] * inv),
                       dt_pack(o[8 * i + 2] * inv, o[8 * i + 3] * inv), dt_pack(o[8 * i + 4] * inv, o[8 * i + 5] * inv),
                       dt_pack(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
          }
          named_bar_sync(1 + wg, 128);  // every row of the warpgroup is done with this head's k | v
        }
        fence_async_smem();
        mbar_arrive(&bars[BAR_B_READY]);
        DT_TR(6 + 8 * l);

        // ---- norm2 + modulate (the folded proj bias is owed to X)
        mbar_wait(&bars[BAR_X_DONE], n_xdone & 1);
        ++n_xdone;
        tc_fence_after();
        DT_TR(7 + 8 * l);
        ln_modulate(P.b_proj, P.shift_mlp, P.scale_mlp, false, 0.f, 0.f);
        DT_TR(8 + 8 * l);

        // ---- fc1 quarter -> GELU -> fp16 F buffer (this warpgroup writes k-block `wg` of the quarter)
#pragma unroll 1
        for (int qd = 0; qd < 4; ++qd) {
          const int b = qd & 1;
          mbar_wait(&bars[BAR_FACC_FULL + b], facc.full_parity(b));
          tc_fence_after();
          float v[64];
          const uint32_t acc = tmem_acc + lane_base + b * 128 + wg * 64;
          tmem_ld32(acc, v);
          tmem_ld32(acc + 32, v + 32);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&bars[BAR_FACC_FREE + b]);
          const float4* b1 = reinterpret_cast<const float4*>(P.b_fc1 + qd * 128 + wg * 64);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 b4 = __ldg(b1 + i);
            v[4 * i] = gelu_erf(v[4 * i] + b4.x); v[4 * i + 1] = gelu_erf(v[4 * i + 1] + b4.y);
            v[4 * i + 2] = gelu_erf(v[4 * i + 2] + b4.z); v[4 * i + 3] = gelu_erf(v[4 * i + 3] + b4.w);
          }
          {
            uint32_t ph;
            if (fbuf.empty_parity(b, ph)) mbar_wait(&bars[BAR_F_FREE + b], ph);  // fc2 of two quarters ago has read this buffer
          }
          const uint32_t frow = sB + b * 32768 + wg * DT_SLOT + r * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dt_sts16(frow + ((static_cast<uint32_t>(i) ^ sw) << 4), dt_pack(v[8 * i], v[8 * i + 1]), dt_pack(v[8 * i + 2], v[8 * i + 3]),
                     dt_pack(v[8 * i + 4], v[8 * i + 5]), dt_pack(v[8 * i + 6], v[8 * i + 7]));
          fence_async_smem();
          mbar_arrive(&bars[BAR_F_READY + b]);
        }
        DT_TR(9 + 8 * l);
      }

      // ---- after the last block: X (+ owed fc2 bias) -> weight_layer score -> softmax over the V views -> pooled row
      mbar_wait(&bars[BAR_X_DONE], n_xdone & 1);
      ++n_xdone;
      tc_fence_after();
      DT_TR(40);
      const float* pend = p.L[L - 1].b_fc2;
      float dot = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[64];
        tmem_ld32(xaddr + c * 64, v);
        tmem_ld32(xaddr + c * 64 + 32, v + 32);
        tmem_ld_wait();
        const float4* pb = reinterpret_cast<const float4*>(pend + wg * 128 + c * 64);
        const float4* pw = reinterpret_cast<const float4*>(p.pool_w + wg * 128 + c * 64);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b4 = __ldg(pb + i), w4 = __ldg(pw + i);
          v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
          d0 = fmaf(v[4 * i], w4.x, d0); d1 = fmaf(v[4 * i + 1], w4.y, d1);
          d0 = fmaf(v[4 * i + 2], w4.z, d0); d1 = fmaf(v[4 * i + 3], w4.w, d1);
        }
        dot += d0 + d1;
        if (p.x_out != nullptr && row_g < p.R) {
          float4* dst = reinterpret_cast<float4*>(p.x_out + row_g * DT_C + wg * 128 + c * 64);
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
      float score, unused;
      exchange2(dot, 0.f, score, unused);
      score += __ldg(p.pool_b);
      // softmax over the V rows of the point: the V lanes are consecutive lanes of this warp
      float mx = score;
      for (int o = 1; o < V; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float pe = expf(score - mx);
      float den = pe;
      for (int o = 1; o < V; o <<= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
      const float wt = pe / den;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(xaddr + c * 32, v);
        tmem_ld_wait();
        const float4* pb = reinterpret_cast<const float4*>(pend + wg * 128 + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = __ldg(pb + i);
          v[4 * i] = (v[4 * i] + b4.x) * wt; v[4 * i + 1] = (v[4 * i + 1] + b4.y) * wt;
          v[4 * i + 2] = (v[4 * i + 2] + b4.z) * wt; v[4 * i + 3] = (v[4 * i + 3] + b4.w) * wt;
        }
        for (int o = 1; o < V; o <<= 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        }
        if ((r & (V - 1)) == 0 && row_g < p.R) {
          uint4* dst = reinterpret_cast<uint4*>(p.pooled + (row_g / V) * DT_C + wg * 128 + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_uint4(dt_pack(v[8 * i], v[8 * i + 1]), dt_pack(v[8 * i + 2], v[8 * i + 3]), dt_pack(v[8 * i + 4], v[8 * i + 5]),
                                dt_pack(v[8 * i + 6], v[8 * i + 7]));
        }
      }
      DT_TR(41);
      tc_fence_before();
      mbar_arrive(&bars[BAR_TILE_DONE]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_x, 512);
  }
}

// W'[n, :] = fp16(gate[n] * W[n, :]), b'[n] = gate[n] * b[n]
struct FoldJobs {
  mvd_fold_job j[8];
  int n;
};
__global__ void dit_fold_kernel(const FoldJobs jobs) {
  pdl_trigger();
  pdl_wait();
  const mvd_fold_job& jb = jobs.j[blockIdx.y];
  const int K8 = jb.K >> 3;
  const long long total = static_cast<long long>(jb.N) * K8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / K8);
    const float g = jb.gate[n];
    const uint4 u = reinterpret_cast<const uint4*>(jb.w)[i];
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h[k]);
      ow[k] = dt_pack(f.x * g, f.y * g);
    }
    reinterpret_cast<uint4*>(jb.w_out)[i] = o;
    if (i % K8 == 0) jb.b_out[n] = g * jb.bias[n];
  }
}

}  // namespace mvd

using namespace mvd;

#ifdef MVD_GEMM_TRACE
extern "C" int mvd_debug_dit_trace(void* dst) {  // 160 x 64 clock stamps
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(dst, g_dit_trace, sizeof(unsigned int) * 160 * 64) != cudaSuccess) return -1;
  return 0;
}
#endif

extern "C" int mvd_dit_fold_gates(const mvd_fold_job* jobs, int32_t n_jobs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (jobs == nullptr || n_jobs <= 0 || n_jobs > 8) return set_error(MVD_EINVAL, "mvd_dit_fold_gates: 1..8 jobs");
  FoldJobs fj;
  memset(&fj, 0, sizeof(fj));
  fj.n = n_jobs;
  for (int i = 0; i < n_jobs; ++i) {
    const mvd_fold_job& j = jobs[i];
    if (!j.w || !j.gate || !j.bias || !j.w_out || !j.b_out || j.N <= 0 || j.K <= 0 || (j.K & 7) != 0)
      return set_error(MVD_EINVAL, "mvd_dit_fold_gates: job %d: null pointer or K not a multiple of 8", i);
    if ((reinterpret_cast<uintptr_t>(j.w) & 15) || (reinterpret_cast<uintptr_t>(j.w_out) & 15))
      return set_error(MVD_EALIGN, "mvd_dit_fold_gates: weights must be 16-byte aligned");
    fj.j[i] = j;
  }
  MVD_CUDA_CHECK(launch_kernel(dit_fold_kernel, dim3(32, n_jobs), dim3(256), 0, stream, 1, fj));
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_gridattn_dit_f16(const mvd_dit_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return set_error(MVD_EINVAL, "mvd_gridattn_dit_f16: null args");
  if (a->R <= 0 || a->V <= 0 || a->V > DT_MAXV || (a->V & (a->V - 1)) != 0 || (a->R % a->V) != 0)
    return set_error(MVD_EINVAL, "mvd_gridattn_dit_f16: V must be a power of two <= %d that divides R", DT_MAXV);
  if (a->layers < 1 || a->layers > DT_MAXL) return set_error(MVD_EINVAL, "mvd_gridattn_dit_f16: 1..%d layers", DT_MAXL);
  if (a->token_k <= 0 || (a->token_k & 7) != 0 || a->token_ld < a->token_k || (a->token_ld & 7) != 0 || a->w_pre_ld < a->token_k || (a->w_pre_ld & 7) != 0)
    return set_error(MVD_EINVAL, "mvd_gridattn_dit_f16: token_k / token_ld / w_pre_ld must be multiples of 8 with ld >= k");
  if (!a->tokens || !a->w_pre || !a->b_pre || !a->pool_w || !a->pool_b || !a->pooled) return set_error(MVD_EINVAL, "mvd_gridattn_dit_f16: null pointer");
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  if (!al16(a->tokens) || !al16(a->w_pre) || !al16(a->b_pre) || !al16(a->pool_w) || !al16(a->pooled) || !al16(a->x_out))
    return set_error(MVD_EALIGN, "mvd_gridattn_dit_f16: buffers must be 16-byte aligned");
  DitMaps maps;
  memset(&maps, 0, sizeof(maps));
  DitParams p;
  memset(&p, 0, sizeof(p));
  p.R = a->R;
  p.V = a->V;
  p.layers = a->layers;
  p.nkb_pre = (a->token_k + 63) / 64;
  p.n_tiles = (a->R + DT_BM - 1) / DT_BM;
  p.b_pre = a->b_pre;
  p.pool_w = a->pool_w;
  p.pool_b = a->pool_b;
  p.pooled = static_cast<__half*>(a->pooled);
  p.x_out = a->x_out;
  p.eps = a->eps;
  int rc = make_tmap_2d(&maps.tok, a->tokens, a->token_k, a->R, a->token_ld, 64, DT_BM);
  if (rc == MVD_OK) rc = make_tmap_2d(&maps.wpre, a->w_pre, a->token_k, DT_C, a->w_pre_ld, 64, 128);
  for (int l = 0; l < a->layers && rc == MVD_OK; ++l) {
    const mvd_dit_layer& s = a->layer[l];
    const void* ptrs[] = {s.w_qkv, s.b_qkv, s.w_proj, s.b_proj, s.w_fc1, s.b_fc1, s.w_fc2, s.b_fc2, s.shift_msa, s.scale_msa, s.shift_mlp, s.scale_mlp};
    for (const void* q : ptrs)
      if (q == nullptr || !al16(q)) return set_error(MVD_EINVAL, "mvd_gridattn_dit_f16: layer %d: null or misaligned pointer", l);
    p.L[l] = DitLayerP{s.b_qkv, s.b_proj, s.b_fc1, s.b_fc2, s.shift_msa, s.scale_msa, s.shift_mlp, s.scale_mlp};
    rc = make_tmap_2d(&maps.qkv[l], s.w_qkv, DT_C, 3 * DT_C, DT_C, 64, 96);
    if (rc == MVD_OK) rc = make_tmap_2d(&maps.proj[l], s.w_proj, DT_C, DT_C, DT_C, 64, 128);
    if (rc == MVD_OK) rc = make_tmap_2d(&maps.fc1[l], s.w_fc1, DT_C, DT_HID, DT_C, 64, 128);
    if (rc == MVD_OK) rc = make_tmap_2d(&maps.fc2[l], s.w_fc2, DT_HID, DT_C, DT_HID, 64, 128);
  }
  if (rc != MVD_OK) return rc;
  typedef void (*DitFn)(const DitMaps, const DitParams);
  static const DitFn fns[5] = {dit_kernel<1>, dit_kernel<2>, dit_kernel<4>, dit_kernel<8>, dit_kernel<16>};
  static bool configured = false;
  if (!configured) {
    for (DitFn f : fns) MVD_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM));
    configured = true;
  }
  int vi = 0;
  while ((1 << vi) < a->V) ++vi;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  MVD_CUDA_CHECK(launch_kernel(fns[vi], dim3(grid), dim3(DT_THREADS), DT_SMEM, stream, 1, maps, p));
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}
