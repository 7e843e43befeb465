// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a — persistent, warp-specialised.
//
//   acc[m, n] = sum_k A[m, k] * W[n, k]     fp16 operands (K-major, 128B-swizzled smem tiles via TMA),
//                                           fp32 accumulators in TMEM, fused epilogues.
//
// One CTA per SM loops over work units (output tile 128 x BN, optionally one K-slice of it):
//   warp 0      : TMA producer (one elected lane) — A tile + W tile per 64-wide k-block into a `stages`-deep ring
//   warp 1      : TMEM allocator + UMMA issuer (one elected lane); tcgen05.commit releases ring slots and
//                 publishes the finished accumulator.  TWO accumulators live in TMEM, so the MMA of unit j+1
//                 overlaps the epilogue of unit j.
//   warps 4..7  : epilogue — tcgen05.ld (warp w owns TMEM lanes 32*(w%4)..), bias / per-image row bias /
//                 GELU / SiLU / GEGLU / adaLN gate / residual, then a 128B-swizzled smem staging tile and a
//                 TMA bulk store (coalesced; TMA clips the M / N tails).  The residual tile is TMA-loaded
//                 into smem two chunks ahead.  QKV mode scatters q, k, v^T per head directly.
// Split-K (small-M, weight-bound layers): every K-slice writes its fp32 partial tile to a workspace in a
//   lane-coalesced layout and bumps a per-tile semaphore; the slice that arrives last adds the others'
//   partials to its own TMEM accumulator and runs the normal epilogue (no atomics on the output, no memset).
// For MVD_A_CONV3X3 the A tile of k-block (tap, c-block) is a 4-D TMA box (64 ch, tw, th, tn) of the NHWC image
//   shifted by (kx-1, ky-1); out-of-bounds pixels are zero-filled by TMA == zero padding; im2col never exists.
//
// Replaces the cuBLAS/cuDNN dispatch behind nn.Linear / nn.Conv2d on the reference hot path
// (see include/mvd_b200.h for the file:line list).
#include "common.h"
#include "ptx.cuh"

namespace mvd {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int A_BYTES = BM * BK * 2;
constexpr int STG_BYTES = 128 * 128;  // one staging chunk: 128 rows x 32 fp32 (or 32 fp16 in the first 64 B of... see below)
constexpr int MAX_STAGES = 8;
constexpr int EPI_THREADS = 128;
constexpr int WS_COUNTER_BYTES = 16384;  // 4096 tile semaphores at the head of the split-K workspace

struct GemmKParams {
  int M, N;
  int BN, stages;
  int num_kb, split, kb_per_split;
  int tiles_m, tiles_n, num_units;
  int a_mode;
  int kb_per_tap, C;
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
  int use_out_tma, use_res_tma;
  float4* ws;
  int* counters;
  int acc_stride, tmem_cols;
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
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts_v4_u32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct Unit {
  int tile, s, m_tile, n_tile, kb0, kb1;
  int x0, y0, img0;  // conv tile origin
  int grow0;         // first output row of the tile
};

__device__ __forceinline__ Unit decode_unit(const GemmKParams& p, int u) {
  Unit t;
  t.s = u % p.split;
  t.tile = u / p.split;
  t.n_tile = t.tile / p.tiles_m;
  t.m_tile = t.tile - t.n_tile * p.tiles_m;
  t.kb0 = t.s * p.kb_per_split;
  t.kb1 = min(p.num_kb, t.kb0 + p.kb_per_split);
  t.x0 = t.y0 = t.img0 = 0;
  if (p.a_mode == MVD_A_CONV3X3) {
    const int tx = t.m_tile % p.tiles_x;
    const int ty = (t.m_tile / p.tiles_x) % p.tiles_y;
    const int tz = t.m_tile / (p.tiles_x * p.tiles_y);
    t.x0 = tx * p.tw;
    t.y0 = ty * p.th;
    t.img0 = tz * p.tn;
    t.grow0 = (t.img0 * p.H + t.y0) * p.W + t.x0;  // tiles are full-width rows of whole images: 128 contiguous output rows
  } else {
    t.grow0 = t.m_tile * BM;
  }
  return t;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = A_BYTES + p.BN * 128;
  uint8_t* out_stg = smem + p.stages * stage_bytes;  // 2 x STG_BYTES
  uint8_t* res_stg = out_stg + 2 * STG_BYTES;        // 2 x STG_BYTES (only when use_res_tma)
  uint64_t* bars = reinterpret_cast<uint64_t*>(res_stg + (p.use_res_tma ? 2 * STG_BYTES : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* acc_full = bars + 2 * MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.use_out_tma) tma_prefetch_desc(&tmO);
    if (p.use_res_tma) tma_prefetch_desc(&tmR);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], EPI_THREADS);
      mbar_init(&res_full[b], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first = blockIdx.x;
  const int n_local = (first < p.num_units) ? (p.num_units - first + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const Unit t = decode_unit(p, first + j * gridDim.x);
        for (int kb = t.kb0; kb < t.kb1; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], stage_bytes);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + A_BYTES;
          int kcol;
          if (p.a_mode == MVD_A_CONV3X3) {
            const int tap = kb / p.kb_per_tap;
            const int cb = kb - tap * p.kb_per_tap;
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_4d(sa, &tmA, &full_bar[s], cb * BK, t.x0 + kx - 1, t.y0 + ky - 1, t.img0);
            kcol = tap * p.C + cb * BK;
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, t.m_tile * BM);
            kcol = kb * BK;
          }
          tma_load_2d(sb, &tmB, &full_bar[s], kcol, t.n_tile * p.BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BM, p.BN);
      int it = 0;
      for (int j = 0; j < n_local; ++j) {
        const Unit t = decode_unit(p, first + j * gridDim.x);
        const int buf = j & 1;
        mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * p.acc_stride;
        for (int kb = t.kb0; kb < t.kb1; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * stage_bytes);
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 fp16 = 32 B inside the 128-B swizzle atom: +2 in the (addr >> 4) field
            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (128 threads, thread = output row of the tile)
    const int et = threadIdx.x - 4 * 32;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool geglu = (p.act == MVD_ACT_GEGLU);
    const int n_out = geglu ? p.N / 2 : p.N;         // output columns
    const int out_bn = geglu ? p.BN / 2 : p.BN;      // output columns per tile
    const int nchunks = (out_bn + 31) / 32;  // a single narrow tile (BN < 32, N <= BN) still takes one chunk; TMA / guards clip it
    const bool f16_out = (p.out_mode == MVD_OUT_F16);
    const uint32_t stg_row_bytes = f16_out ? 64 : 128;
    int res_items = 0;  // residual staging items consumed so far (buffer = item & 1, parity = (item >> 1) & 1)
    int out_items = 0;  // out staging items issued so far

    for (int j = 0; j < n_local; ++j) {
      const int u = first + j * gridDim.x;
      const Unit t = decode_unit(p, u);
      const int buf = j & 1;
      const int col_base = t.n_tile * out_bn;
      const int valid_chunks = min(nchunks, (n_out - col_base + 31) / 32);
      const int grow = t.grow0 + r;
      const bool valid = grow < p.M;

      // residual tiles for the first two chunks: requested before the accumulator is even finished
      if (p.use_res_tma && et == 0) {
        for (int c = 0; c < min(2, valid_chunks); ++c) {
          const int b = (res_items + c) & 1;
          mbar_expect_tx(&res_full[b], STG_BYTES);
          tma_load_2d(res_stg + b * STG_BYTES, &tmR, &res_full[b], col_base + c * 32, t.grow0);
        }
      }

      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);

      bool do_final = true;
      if (p.split > 1) {
        // ---- publish this K-slice's partial tile: ws[u][col4][row] (lanes write consecutive float4 -> coalesced)
        const int ws_cols4 = ((p.BN + 31) / 32) * 8;  // float4 columns of one partial tile
        float4* wsp = p.ws + static_cast<size_t>(u) * ws_cols4 * BM;
        for (int c = 0; c < nchunks; ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) wsp[(c * 8 + i) * BM + r] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        __threadfence();
        named_bar_sync(1, EPI_THREADS);
        if (et == 0) {
          const int old = atomicAdd(&p.counters[t.tile], 1);
          const int last = (old == p.split - 1) ? 1 : 0;
          if (last) p.counters[t.tile] = 0;  // self-resetting semaphore: ready for the next launch
          *last_flag = last;
        }
        named_bar_sync(1, EPI_THREADS);
        do_final = (*last_flag != 0);
        if (do_final) __threadfence();
      }

      if (do_final) {
        const float* rb = nullptr;
        if (p.rowbias != nullptr && valid) rb = p.rowbias + static_cast<size_t>(grow / p.rows_per_group) * p.N;
        for (int c = 0; c < valid_chunks; ++c) {
          const int oc = col_base + c * 32;  // first output column of this chunk
          float v[32];
          if (geglu) {
            float g[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_ld32(taddr + p.BN / 2 + c * 32, g);
            tmem_ld_wait();
            const int nv = t.n_tile * p.BN + c * 32;  // packed column of the value half
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float a = v[i], b = g[i];
              if (p.bias != nullptr) {
                a += __ldg(p.bias + nv + i);
                b += __ldg(p.bias + nv + p.BN / 2 + i);
              }
              v[i] = a * gelu_erf(b);
            }
          } else {
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            if (p.split > 1) {
              for (int s2 = 0; s2 < p.split; ++s2) {
                if (s2 == t.s) continue;
                const float4* wo = p.ws + static_cast<size_t>(t.tile * p.split + s2) * (((p.BN + 31) / 32) * 8) * BM;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 w4 = __ldcg(wo + (c * 8 + i) * BM + r);
                  v[4 * i] += w4.x; v[4 * i + 1] += w4.y; v[4 * i + 2] += w4.z; v[4 * i + 3] += w4.w;
                }
              }
            }
            const bool full = (oc + 32 <= p.N);
            if (p.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += (full || oc + i < p.N) ? __ldg(p.bias + oc + i) : 0.f;
            }
            if (rb != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += (full || oc + i < p.N) ? __ldg(rb + oc + i) : 0.f;
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
              for (int i = 0; i < 32; ++i) v[i] *= (full || oc + i < p.N) ? __ldg(p.colscale + oc + i) : 0.f;
            }
            if (p.use_res_tma) {
              const int b = res_items & 1;
              mbar_wait(&res_full[b], (res_items >> 1) & 1);
              const uint32_t rrow = smem_u32(res_stg + b * STG_BYTES) + r * 128;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 t4 = lds_v4(rrow + ((i ^ (r & 7)) << 4));
                v[4 * i] += t4.x; v[4 * i + 1] += t4.y; v[4 * i + 2] += t4.z; v[4 * i + 3] += t4.w;
              }
            } else if (p.residual != nullptr && valid) {
              const float* res = p.residual + static_cast<size_t>(grow) * p.ldr + oc;
              if (full && (p.ldr & 3) == 0) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(res + i);
                  v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
                }
              } else {
                for (int i = 0; i < 32 && oc + i < p.N; ++i) v[i] += res[i];
              }
            }
          }

          if (p.use_out_tma) {
            // ---- staged, coalesced store: registers -> swizzled smem tile -> TMA bulk store (clips the tails)
            const int ob = out_items & 1;
            if (et == 0) tma_store_wait_read<1>();  // the store issued two chunks ago has finished reading buffer `ob`
            named_bar_sync(1, EPI_THREADS);
            const uint32_t srow = smem_u32(out_stg + ob * STG_BYTES) + r * stg_row_bytes;
            if (f16_out) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                sts_v4_u32(srow + (i << 4), pack_h2(v[8 * i], v[8 * i + 1]), pack_h2(v[8 * i + 2], v[8 * i + 3]),
                           pack_h2(v[8 * i + 4], v[8 * i + 5]), pack_h2(v[8 * i + 6], v[8 * i + 7]));
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) sts_v4(srow + ((i ^ (r & 7)) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            fence_async_smem();
            named_bar_sync(1, EPI_THREADS);  // staging tile complete; every thread is also done with residual buffer res_items & 1
            if (et == 0) {
              tma_store_2d(&tmO, out_stg + ob * STG_BYTES, oc, t.grow0);
              tma_store_commit();
              if (p.use_res_tma && c + 2 < valid_chunks) {
                const int b = res_items & 1;
                mbar_expect_tx(&res_full[b], STG_BYTES);
                tma_load_2d(res_stg + b * STG_BYTES, &tmR, &res_full[b], col_base + (c + 2) * 32, t.grow0);
              }
            }
            ++out_items;
            if (p.use_res_tma) ++res_items;
          } else if (valid) {
            // ---- direct stores (QKV head scatter, or an output whose leading dimension TMA cannot address)
            const bool full = (oc + 32 <= n_out);
            if (p.out_mode == MVD_OUT_F32) {
              float* dst = reinterpret_cast<float*>(p.out) + static_cast<size_t>(grow) * p.ldc + oc;
              if (full && (p.ldc & 3) == 0) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              } else {
                for (int i = 0; i < 32 && oc + i < n_out; ++i) dst[i] = v[i];
              }
            } else if (p.out_mode == MVD_OUT_F16) {
              __half* dst = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(grow) * p.ldc + oc;
              if (full && (p.ldc & 7) == 0) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) store8_f16(dst + i, v + i);
              } else {
                for (int i = 0; i < 32 && oc + i < n_out; ++i) dst[i] = __float2half_rn(v[i]);
              }
            } else {  // MVD_OUT_QKV_HEADS
              const int inner = p.heads * p.dhead;
              const int img = grow / p.seq;
              const int pos = grow - img * p.seq;
#pragma unroll 1
              for (int i = 0; i < 32; i += 8) {
                const int n = oc + i;
                if (n >= p.N) break;
                const int which = n / inner;
                const int rem = n - which * inner;
                const int h = rem / p.dhead;
                const int jj = rem - h * p.dhead;
                const size_t bh = static_cast<size_t>(img) * p.heads + h;
                if (which < 2) {
                  __half* base = reinterpret_cast<__half*>(which == 0 ? p.out : p.out_k);
                  store8_f16(base + (bh * p.seq + pos) * p.dpad + jj, v + i);
                } else {
                  __half* base = reinterpret_cast<__half*>(p.out_vt) + (bh * p.dpad + jj) * p.seq + pos;
#pragma unroll
                  for (int e = 0; e < 8; ++e) base[static_cast<size_t>(e) * p.seq] = __float2half_rn(v[i + e]);
                }
              }
            }
          }
        }
      }
      // accumulator `buf` is drained: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
    if (et == 0 && p.use_out_tma) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host
static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// tile width: N <= 64 -> one tile of N rounded up to 16; else the multiple of 32 in [64, 256] that wastes the fewest
// padded columns (ties -> wider).  (The epilogue walks 32-column chunks.)
static int pick_bn(int N) {
  if (N <= 64) return (N + 15) / 16 * 16;
  int best = 64, best_waste = 1 << 30;
  for (int bn = 64; bn <= 256; bn += 32) {
    const int waste = (N + bn - 1) / bn * bn - N;
    if (waste <= best_waste) {
      best = bn;
      best_waste = waste;
    }
  }
  return best;
}

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

  const bool geglu = a->act == MVD_ACT_GEGLU;
  int bn = a->tile_n;
  if (bn == 0) bn = geglu ? 256 : pick_bn(a->N);
  if (bn < 16 || bn > 256 || (bn & 15) != 0 || ((bn & 31) != 0 && bn < a->N))
    return set_error(MVD_EINVAL, "mvd_gemm_f16: tile_n must be 0, a multiple of 32 in [32, 256], or a multiple of 16 that covers N");
  if (geglu) {
    if ((bn & 63) != 0 || (a->N % bn) != 0) return set_error(MVD_EINVAL, "mvd_gemm_f16: GEGLU needs tile_n a multiple of 64 that divides N");
    if (a->colscale != nullptr || a->residual != nullptr || a->rowbias != nullptr || a->out_mode == MVD_OUT_QKV_HEADS)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: GEGLU supports bias only");
  }
  if (a->out_mode == MVD_OUT_QKV_HEADS) {
    if (a->out_k == nullptr || a->out_vt == nullptr || a->heads <= 0 || a->dhead <= 0 || (a->dhead & 7) != 0 ||
        (a->dpad & 7) != 0 || a->dpad < a->dhead || a->seq <= 0 || a->N != 3 * a->heads * a->dhead ||
        (a->M % a->seq) != 0 || a->act != MVD_ACT_NONE)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: bad QKV_HEADS arguments");
  }
  p.BN = bn;

  CUtensorMap tmA, tmB, tmO, tmR;
  if (a->a_mode == MVD_A_ROWMAJOR) {
    if ((a->lda & 7) != 0 || a->lda < a->K) return set_error(MVD_EALIGN, "mvd_gemm_f16: lda must be >= K and a multiple of 8");
    p.num_kb = (a->K + BK - 1) / BK;
    p.tiles_m = (a->M + BM - 1) / BM;
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
    p.tiles_m = p.tiles_x * p.tiles_y * tiles_z;
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
  p.tiles_n = (a->N + bn - 1) / bn;
  const int tiles = p.tiles_m * p.tiles_n;
  const int sms = num_sms();

  // ---- split-K
  int split = a->split_k;
  const size_t ws_avail = (a->splitk_ws != nullptr && a->splitk_ws_bytes > WS_COUNTER_BYTES) ? static_cast<size_t>(a->splitk_ws_bytes) - WS_COUNTER_BYTES : 0;
  const size_t tile_ws = static_cast<size_t>((bn + 31) / 32 * 32) * BM * sizeof(float);
  if (split <= 0) {  // auto: only when the tiles cannot fill half the machine and K is deep
    split = 1;
    if (!geglu && tiles * 2 <= sms && p.num_kb >= 8) {
      split = sms / tiles;
      if (split > p.num_kb / 4) split = p.num_kb / 4;
      if (split > 16) split = 16;
      if (split < 1) split = 1;
    }
    while (split > 1 && (tiles > WS_COUNTER_BYTES / 4 || static_cast<size_t>(tiles) * split * tile_ws > ws_avail)) --split;
  } else if (split > 1) {
    if (geglu) return set_error(MVD_EINVAL, "mvd_gemm_f16: split_k is not supported with GEGLU");
    if (tiles > WS_COUNTER_BYTES / 4 || static_cast<size_t>(tiles) * split * tile_ws > ws_avail)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: split_k=%d needs a split-K workspace of %zu bytes (args.splitk_ws)", split,
                       static_cast<size_t>(tiles) * split * tile_ws + WS_COUNTER_BYTES);
  }
  if (split > p.num_kb) split = p.num_kb;
  p.kb_per_split = (p.num_kb + split - 1) / split;
  split = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.split = split;
  p.num_units = tiles * split;
  if (split > 1) {
    if ((reinterpret_cast<uintptr_t>(a->splitk_ws) & 15) != 0) return set_error(MVD_EALIGN, "mvd_gemm_f16: splitk_ws must be 16-byte aligned");
    p.counters = reinterpret_cast<int*>(a->splitk_ws);
    p.ws = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a->splitk_ws) + WS_COUNTER_BYTES);
  }

  // ---- epilogue paths
  const int out_elem = (a->out_mode == MVD_OUT_F32) ? 4 : 2;
  const int n_out = geglu ? a->N / 2 : a->N;
  p.use_out_tma = (a->out_mode != MVD_OUT_QKV_HEADS) && ((static_cast<long long>(a->ldc) * out_elem) % 16 == 0) &&
                  ((reinterpret_cast<uintptr_t>(a->out) & 15) == 0) && a->ldc >= n_out;
  if (a->out_mode != MVD_OUT_QKV_HEADS && a->ldc < n_out) return set_error(MVD_EINVAL, "mvd_gemm_f16: ldc is smaller than the output width");
  if (p.use_out_tma) {
    int rc = make_tmap_2d_ex(&tmO, a->out, out_elem, n_out, a->M, a->ldc, 32, BM, out_elem == 4 ? 128 : 0);
    if (rc != MVD_OK) return rc;
  } else {
    tmO = tmB;
  }
  p.use_res_tma = p.use_out_tma && a->residual != nullptr && split == 1 && ((static_cast<long long>(a->ldr) * 4) % 16 == 0) &&
                  ((reinterpret_cast<uintptr_t>(a->residual) & 15) == 0) && a->ldr >= a->N;
  if (a->residual != nullptr && a->ldr < a->N) return set_error(MVD_EINVAL, "mvd_gemm_f16: ldr is smaller than N");
  if (p.use_res_tma) {
    int rc = make_tmap_2d_ex(&tmR, a->residual, 4, a->N, a->M, a->ldr, 32, BM, 128);
    if (rc != MVD_OK) return rc;
  } else {
    tmR = tmB;
  }

  // ---- shared memory / TMEM budget
  const int stage_bytes = A_BYTES + bn * 128;
  const int fixed = 2 * STG_BYTES + (p.use_res_tma ? 2 * STG_BYTES : 0) + 512;
  int stages = (232448 - 1024 - fixed) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return set_error(MVD_EINVAL, "mvd_gemm_f16: tile does not fit in shared memory");
  p.stages = stages;
  const int dyn = stages * stage_bytes + fixed + 1024;
  p.acc_stride = bn <= 128 ? 128 : 256;
  p.tmem_cols = 2 * p.acc_stride;

  static bool configured = false;
  if (!configured) {
    MVD_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    configured = true;
  }
  const int grid = p.num_units < sms ? p.num_units : sms;
  gemm_tc_kernel<<<grid, GEMM_THREADS, dyn, stream>>>(tmA, tmB, tmO, tmR, p);
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

extern "C" int mvd_geglu_row_permutation(int32_t inner, int32_t tile_n, int32_t* perm) {
  if (inner <= 0 || perm == nullptr || tile_n < 64 || tile_n > 256 || (tile_n & 63) != 0 || (2 * inner) % tile_n != 0)
    return set_error(MVD_EINVAL, "mvd_geglu_row_permutation: bad arguments");
  const int half = tile_n / 2;
  for (int r = 0; r < 2 * inner; ++r) {
    const int tile = r / tile_n, w = r % tile_n;
    // nn.Linear(dim, 2*inner): rows [0, inner) are the value half, [inner, 2*inner) the gate half (chunk(2, dim=-1))
    perm[r] = (w < half) ? tile * half + w : inner + tile * half + (w - half);
  }
  return MVD_OK;
}
