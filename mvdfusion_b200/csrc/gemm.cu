// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a — persistent, warp-specialised.
//
//   acc[m, n] = sum_k A[m, k] * W[n, k]     fp16 operands (K-major, 128B-swizzled smem tiles via TMA),
//                                           fp32 accumulators in TMEM, fused epilogues.
//
// One CTA (384 or 512 threads) per SM loops over work units (output tile 128 x BN, optionally one K-slice of it):
//   warp 0      : TMA producer (one elected lane) — A tile + W tile per 64-wide k-block into a `stages`-deep ring
//   warp 1      : TMEM allocator + UMMA issuer (one elected lane); tcgen05.commit releases ring slots and
//                 publishes the finished accumulator.  TWO accumulators live in TMEM, so the MMA of unit j+1
//                 overlaps the epilogue of unit j.
//   warps 4..   : two or three epilogue warpgroups (template NWG) that take turns over the 32-column chunks of a tile.
//                 Three (512 threads, 128 registers per thread) for every non-split specialisation: the short-K GEMMs of the
//                 step are bound by the epilogue's latency chains, and a third chunk in flight is worth 13-22 % there; the
//                 split-K bodies need more registers and keep two.  Per chunk:
//                 phase A  tcgen05.ld (thread = tile row) -> 128B-swizzled fp32 staging tile in smem
//                          (GEGLU multiplies value * gelu(gate) here; the v^T part of a QKV scatter leaves from here),
//                 phase B  thread = (row, 16-byte segment): bias / per-image row bias / GELU / SiLU / adaLN gate /
//                          residual / split-K partials, all read and written with coalesced 16-byte accesses; an fp32
//                          output can be stored a second time as fp16 (out16: the next GEMM's operand).
//                 With two warpgroups each owns two staging tiles and the residual of the NEXT chunk is already in flight
//                 (registers) while this one is processed; with three, one staging tile each and no prefetch.
// Split-K (small-M, weight-bound layers; all slices of a tile are co-resident): slice s parks the chunks it does not
//   own in a workspace, bumps the tile semaphore and waits for its siblings; then every slice reduces and finishes the
//   chunks it owns (chunk c belongs to slice c % split) — the reduction is spread over the slices, no atomics on data.
// For MVD_A_CONV3X3 the A tile of k-block (tap, c-block) is a 4-D TMA box (64 ch, tw, th, tn) of the NHWC image
//   shifted by (kx-1, ky-1); out-of-bounds pixels are zero-filled by TMA == zero padding; im2col never exists.
//
// Replaces the cuBLAS/cuDNN dispatch behind nn.Linear / nn.Conv2d on the reference hot path
// (see include/mvd_b200.h for the file:line list).
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "ptx.cuh"

namespace mvd {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int STG_BYTES = 128 * 128;  // one staging chunk: 128 rows x 32 fp32, 128B-swizzled
constexpr int MAX_STAGES = 8;
constexpr int WG_THREADS = 128;
constexpr int WS_COUNTER_BYTES = 16384;  // 2048 x {arrive, done} tile semaphores at the head of the split-K workspace

// ---- optional in-kernel timeline (-DMVD_GEMM_TRACE; tools/gemm_trace.py): every CTA leaves one record with the SM clock at the
//      hand-over points of its warp roles plus the global timer at entry / exit.  Not compiled into the product library.
#ifdef MVD_GEMM_TRACE
struct TraceRec {
  unsigned long long gt_entry, gt_exit;
  unsigned int st[12];
  int bid, grid, M, N, num_kb, BN, split, n_local, flags, pad;
};
constexpr unsigned TRACE_CAP = 1u << 17;
__device__ TraceRec g_trace[TRACE_CAP];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ unsigned long long trace_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define MVD_TR(i) do { tr_s[i] = static_cast<unsigned int>(clock64()); } while (0)
#else
#define MVD_TR(i) do { } while (0)
#endif

// Division by a run-time constant as multiply-high + shift (the persistent loops decode a work unit per tile in every warp
// role, and the QKV / row-bias epilogues divide per chunk: a hardware-emulated integer division is ~20 dependent instructions).
// Exact for 0 <= n < 2^31, 1 <= d < 2^31.
struct FastDiv {
  uint32_t mul, shr, d;
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul) >> shr); }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = div(n);
    r = n - q * static_cast<int>(d);
  }
};
static FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u, static_cast<uint32_t>(d < 1 ? 1 : d)};
  if (f.d > 1) {
    uint32_t lg = 0;
    while ((1ull << lg) < f.d) ++lg;                       // ceil(log2 d)
    const unsigned long long pw = 1ull << (31 + lg);       // n < 2^31: one extra bit of the quotient estimate is enough
    f.mul = static_cast<uint32_t>((pw + f.d - 1) / f.d);
    f.shr = lg - 1;                                        // (n * mul) >> (31 + lg) == umulhi(n, mul) >> (lg - 1)
  }
  return f;
}

struct GemmKParams {
  int M, N;
  int BN, stages;
  int num_kb, split, kb_per_split;
  int tiles_m, tiles_n, num_units;
  int tiles_m_real;  // m-tiles of the problem (tiles_m counts CTA-pair rows in pair mode)
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
  int vec_bias, vec_rowbias, vec_colscale, vec_res, vec_out;  // 16-byte (8-byte for fp16 out) accesses are legal
  int qkv_direct;                                             // QKV scatter of every chunk from phase A (generic geometry)
  float* ws;
  int* counters;
  int acc_stride, tmem_cols;
  __half* out16;            // optional fp16 copy of an F32 output (the next GEMM's operand), [M, ld16]
  int ld16, vec_out16;
  int out16_lo;             // > 0: the fp16 rounding residual v - fp16(v) is stored too, out16_lo columns to the right
  // hi/lo split operands (args.hilo): A = [A_hi | A_lo], W = [W_hi | W_lo]; the k-blocks run over three segments
  // A_hi W_hi, A_lo W_hi, A_hi W_lo (kb_seg k-blocks each).  a_lo_off / w_lo_off: column (channel) offsets of the lo halves.
  int hilo, kb_seg, a_lo_off, w_lo_off;
  int cstride, cpad;        // CONV3X3: stride (1 / 2) and low-side padding (1, or 0 for the VAE encoder's pad-high-only downsample)
  // CONV3X3 + conv_up2 (nearest x2 upsample folded into the convolution): the N columns are four phase blocks of n_real = N / 4 output
  // channels; phase (py, px) reads the 2 x 2 source taps at offsets (py - 1 + a, px - 1 + b) and owns the output pixels (2y + py, 2x + px)
  int taps_w, taps;         // taps per row / per pixel: 3 / 9, or 2 / 4 with up2
  int up2, n_real, up_lw, up_lh;
  FastDiv fd_tpp;           // n-tiles per phase
  // TMA epilogue (template TMAE): per-warpgroup chunk slots [32 fp32 columns x 128 rows | fp16 copy | fp16 lo] that the residual
  // is loaded into and the finished chunk is stored from, both by TMA
  int spw, slot_bytes, slot_h16, slot_lo;  // slots per warpgroup (1 / 2), bytes per slot, offsets of the fp16 tiles inside a slot
  // wide pair tiles (256 < BN <= 320, TMAE only): the tile is two N = BN / 2 MMAs into ONE accumulator of BN columns (single-buffered);
  // each CTA of the pair stages rows [rank q, rank q + q) and [BN / 2 + rank q, ...) of the W tile, q = BN / 4
  int wide;
  // LayerNorm between two GEMMs (mvd_b200.h, ABI 13).  Producer: (sum, sum of squares) per row and 32-column chunk -> ln_stats_out
  // [N/32][M] float2.  Consumer: ln_stats [ln_parts][M] float2 of its A rows, ln_colsum[n] = column sums of the gamma-scaled weights.
  float2* ln_stats_out;
  const float2* ln_stats;
  const float* ln_colsum;
  float ln_eps, ln_inv_k;
  int ln_parts;
  FastDiv fd_split, fd_tiles_m, fd_tiles_x, fd_tiles_y, fd_seq, fd_inner, fd_dhead, fd_rpg;
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
__device__ __forceinline__ void sts_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// TMA tile load / store with the shared-memory side given as a shared-window address
__device__ __forceinline__ void tma_load_2d_raw(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_raw(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
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
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// 4 consecutive floats starting at p; `nvalid` of them exist (>= 4 -> all); vector access only when `vec`
__device__ __forceinline__ float4 ldg4(const float* p, bool vec, int nvalid) {
  if (vec && nvalid >= 4) return __ldg(reinterpret_cast<const float4*>(p));
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid > 0) r.x = __ldg(p);
  if (nvalid > 1) r.y = __ldg(p + 1);
  if (nvalid > 2) r.z = __ldg(p + 2);
  if (nvalid > 3) r.w = __ldg(p + 3);
  return r;
}
// same, coherent loads (the residual stream is written by earlier kernels of the same graph)
__device__ __forceinline__ float4 ld4(const float* p, bool vec, int nvalid) {
  if (vec && nvalid >= 4) return *reinterpret_cast<const float4*>(p);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid > 0) r.x = p[0];
  if (nvalid > 1) r.y = p[1];
  if (nvalid > 2) r.z = p[2];
  if (nvalid > 3) r.w = p[3];
  return r;
}

// k-block iterator of the TMA producer: the coordinates of the next k-block are kept in registers and advanced with adds and
// compares only.  (A division per k-block sits on the producer's critical path: the deep-K convolutions lost 12 % when the
// coordinate code grew by one more runtime division — profiles/r02_gemm_ab_producer.md.)
// Plain: A column kb*BK (conv: tap (ky, kx), channel block) and the same W column.  hilo: segments 0 / 1 / 2 read
// (A_hi, W_hi) / (A_lo, W_hi) / (A_hi, W_lo), kb_seg k-blocks each.
struct KbIter {
  int a_col, w_col, kx, ky;   // what the TMA loads take
  int cb, tap, seg, left;     // position: channel block in the tap (plain: k-block in the segment), tap, segment
  __device__ __forceinline__ void seek(const GemmKParams& p, int kb) {  // the only divisions: once per work unit
    seg = 0;
    if (p.hilo) {
      seg = kb / p.kb_seg;
      kb -= seg * p.kb_seg;
    }
    if (p.a_mode == MVD_A_CONV3X3) {
      tap = kb / p.kb_per_tap;
      cb = kb - tap * p.kb_per_tap;
      ky = tap / p.taps_w;
      kx = tap - ky * p.taps_w;
    } else {
      tap = 0; ky = 0; kx = 0;
      cb = kb;
    }
    place(p);
  }
  __device__ __forceinline__ void place(const GemmKParams& p) {
    const int a_off = seg == 1 ? p.a_lo_off : 0, w_off = seg == 2 ? p.w_lo_off : 0;
    a_col = cb * BK + a_off;
    w_col = (p.a_mode == MVD_A_CONV3X3 ? tap * p.C : 0) + cb * BK + w_off;
  }
  __device__ __forceinline__ void next(const GemmKParams& p) {
    ++cb;
    a_col += BK;
    w_col += BK;
    if (p.a_mode == MVD_A_CONV3X3) {
      if (cb == p.kb_per_tap) {
        cb = 0;
        ++tap;
        if (++kx == p.taps_w) { kx = 0; ++ky; }
        if (tap == p.taps) { tap = 0; kx = 0; ky = 0; ++seg; }
        place(p);
      }
    } else if (cb == p.kb_seg) {  // plain GEMM: only the hilo form ever gets here before the unit ends
      cb = 0;
      ++seg;
      place(p);
    }
  }
};

struct Unit {
  int tile, s, m_tile, n_tile, kb0, kb1;
  int x0, y0, img0;  // conv tile origin
  int grow0;         // first output row of the tile
  int ox, oy;        // conv: low-side offset of the taps (the padding; with up2 it depends on the unit's phase)
  int ph;            // up2: phase 2 py + px of the unit's columns
};

// pair_rank < 0: one CTA per tile.  Otherwise the unit list enumerates 256-row tile PAIRS and CTA `pair_rank` of the
// cluster owns m-tile 2*mt + rank (which may lie beyond the problem when the m-tile count is odd: every access of such
// a tile is clipped, the CTA still takes part in the pair's loads and barriers).
__device__ __forceinline__ Unit decode_unit(const GemmKParams& p, int u, int pair_rank) {
  Unit t;
  p.fd_split.divmod(u, t.tile, t.s);
  p.fd_tiles_m.divmod(t.tile, t.n_tile, t.m_tile);
  if (pair_rank >= 0) {
    t.m_tile = 2 * t.m_tile + pair_rank;
    t.tile = t.n_tile * p.tiles_m_real + t.m_tile;  // split-K semaphores / partials are per real tile
  }
  t.kb0 = t.s * p.kb_per_split;
  t.kb1 = min(p.num_kb, t.kb0 + p.kb_per_split);
  t.x0 = t.y0 = t.img0 = 0;
  t.ox = t.oy = p.cpad;
  t.ph = 0;
  if (p.up2) {
    t.ph = p.fd_tpp.div(t.n_tile);
    t.oy = 1 - (t.ph >> 1);
    t.ox = 1 - (t.ph & 1);
  }
  if (p.a_mode == MVD_A_CONV3X3) {
    int tx, ty, tz, rest;
    p.fd_tiles_x.divmod(t.m_tile, rest, tx);
    p.fd_tiles_y.divmod(rest, tz, ty);
    t.x0 = tx * p.tw;
    t.y0 = ty * p.th;
    t.img0 = tz * p.tn;
    t.grow0 = (t.img0 * p.H + t.y0) * p.W + t.x0;  // full-width rows of whole images, or a 128-pixel row segment: 128 contiguous output rows
  } else {
    t.grow0 = t.m_tile * BM;
  }
  return t;
}

// Epilogue specialisation: ACT / OUT / RES / SPLIT >= 0 fix the activation, output mode, "has residual" and "split-K"
// at compile time (-1 = decided at run time); VEC = every epilogue access is a full, aligned 16-byte (8-byte fp16)
// vector, so no tails exist.  The specialised bodies are several times smaller than the generic one, which matters:
// a warp walks its epilogue code once per chunk, and the generic body does not fit the instruction cache.
// NWG = epilogue warpgroups: 2 (384 threads, two staging tiles each) or 3 (512 threads -> 128 registers per thread, one
// staging tile each).  Short-K GEMMs are bound by the epilogue's latency chains (TMEM -> staging -> global with only two
// warps per scheduler); a third warpgroup is a third chunk in flight.  Used by the specialisations whose epilogue fits
// 128 registers.
//
// TMAE ("TMA epilogue", unsplit F32 / F16 outputs): the in-kernel timeline (tools/gemm_trace.py, profiles/r02_gemm_trace_v2.txt)
// shows the short-K GEMMs of the step waiting on their epilogue — 3-6 us per 128 x 160..256 tile of per-thread global loads /
// stores chained behind their latencies, against 1.3 us of MMAs.  With TMAE no epilogue thread touches global memory for data:
// warp 2 TMA-loads the fp32 residual chunk (128 rows x 32 columns, 128B-swizzled) into the slot of the warpgroup that will
// finish that chunk, one or two chunks ahead; thread = row reads its accumulator row from TMEM, adds bias / row bias / residual
// (from the slot), applies the activation, writes the result back into the slot (and optional fp16 copies next to it), and one
// thread per warpgroup issues the TMA store of the chunk.  TMA clips the M and N tails.
// WIDE: 320-column pair tile (two N = BN / 2 MMAs, one accumulator).  A template parameter rather than a run-time flag: the producer and
// issuer loops are single threads whose instruction stream is latency-critical — a run-time branch in them (and the code it adds
// to the instruction cache the epilogue warps share) cost the ordinary deep-K convolutions 6 %.
template <int ACT, int OUT, int RES, int SPLIT, bool VEC, bool PAIR, int NWG, bool TMAE, bool WIDE = false>
__global__ void __launch_bounds__(128 + 128 * NWG, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmO16, const __grid_constant__ CUtensorMap tmO16lo,
                   const GemmKParams p) {
  const int act = ACT >= 0 ? ACT : p.act;
  const int out_mode = OUT >= 0 ? OUT : p.out_mode;
  const bool has_res = RES >= 0 ? (RES != 0) : (p.residual != nullptr);
  const bool is_split = SPLIT >= 0 ? (SPLIT != 0) : (p.split > 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_rows = PAIR ? p.BN / 2 : p.BN;  // W rows this CTA stages per k-block (a pair splits the tile's columns)
  const int stage_bytes = A_BYTES + b_rows * 128;
  constexpr int EPI_THREADS = NWG * WG_THREADS;
  constexpr int STG_PER_WG = NWG == 2 ? 2 : 1;
  uint8_t* out_stg = smem + p.stages * stage_bytes;                                      // NWG x STG_PER_WG x STG_BYTES, or the TMAE slots
  float* bias_smem = reinterpret_cast<float*>(out_stg + (TMAE ? NWG * p.spw * p.slot_bytes : NWG * STG_PER_WG * STG_BYTES));   // NWG x 256 floats (GEGLU tile bias)
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_smem + NWG * (TMAE ? 640 : 256));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* acc_full = bars + 2 * MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_full = acc_empty + 2;     // TMAE, one per slot (<= 6): the residual chunk has landed in the slot
  uint64_t* res_empty = res_full + 6;     //       the slot's last store has read it: free for the next residual / result
  uint64_t* out_ready = res_empty + 6;    //       the warpgroup has written the finished chunk into the slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_ready + 6);
  // LayerNorm fold (p.ln_stats): warp 2 reduces the chunk statistics of the NEXT unit's rows to {rstd, -mean * rstd} one unit ahead
  uint64_t* st_full = bars + 48;          // [2] the helper has written rowstat[b] (32 lanes arrive)
  uint64_t* st_empty = bars + 50;         // [2] every epilogue thread has read rowstat[b]
  float2* rowstat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [2][128], behind the 512-byte barrier block
  float* lnvec = reinterpret_cast<float*>(rowstat + 256);  // [2][column sums of the unit's BN (<= 256) weight rows | their bias] (QKV form only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair_rank = PAIR ? static_cast<int>(cluster_ctarank()) : -1;
  const bool leader = !PAIR || pair_rank == 0;
#ifdef MVD_GEMM_TRACE
  unsigned int* tr_s = tmem_slot + 4;  // inside the 512-byte barrier block of the dynamic shared memory
  unsigned long long tr_gt0 = 0;
  if (threadIdx.x == 0) {
    tr_gt0 = trace_gtimer();
    for (int i = 0; i < 12; ++i) tr_s[i] = 0;
    MVD_TR(0);
  }
#endif
  pdl_trigger();  // the next kernel's CTAs may take SMs as ours retire; they block in pdl_wait() until this grid is done

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], PAIR ? 2 * EPI_THREADS : EPI_THREADS);  // pair: the leader collects both CTAs' epilogues
    }
    if (p.ln_stats != nullptr) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(&st_full[b], 32);
        mbar_init(&st_empty[b], EPI_THREADS / 32);
      }
    }
    if (TMAE) {
      tma_prefetch_desc(&tmOut);
      if (has_res) tma_prefetch_desc(&tmRes);
      for (int b = 0; b < 6; ++b) {
        mbar_init(&res_full[b], 1);
        mbar_init(&res_empty[b], 1);
        mbar_init(&out_ready[b], WG_THREADS);
      }
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, p.tmem_cols);
    else tmem_alloc(tmem_slot, p.tmem_cols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before either signals the other's
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) MVD_TR(1);

  const int first = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int ustride = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int n_local = (first < p.num_units) ? (p.num_units - first + ustride - 1) / ustride : 0;

  // ---- LayerNorm fold, statistics helper (one warp; lane owns tile rows lane, lane + 32, lane + 64, lane + 96): sums the per-chunk
  //      (sum, sum of squares) the producing GEMM left for the unit's rows — `ln_parts` coalesced 8-byte loads per row, 16 in flight per
  //      lane — and leaves {rstd, -mean * rstd} in rowstat[b] one unit ahead of the epilogue, so that no epilogue thread waits on L2.
  auto ln_helper = [&](const bool stage_vec) {
    pdl_wait();  // the statistics come from the kernel before us
    int k = 0;
    for (int j = 0; j < n_local; ++j) {
      const Unit t = decode_unit(p, first + j * ustride, pair_rank);
      if (PAIR && t.m_tile >= p.tiles_m_real) continue;
      const int b = k & 1;
      if (k >= 2) mbar_wait(&st_empty[b], ((k >> 1) - 1) & 1);
      if (stage_vec) {  // the unit's column sums and bias: loaded here, one unit ahead, so that no epilogue thread waits on a global load
        const int nb = t.n_tile * p.BN;
        float cvv[8], bvv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kk = lane + 32 * i;
          const bool in = kk < p.BN && nb + kk < p.N;
          cvv[i] = in ? __ldg(p.ln_colsum + nb + kk) : 0.f;
          bvv[i] = (in && p.bias != nullptr) ? __ldg(p.bias + nb + kk) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          lnvec[b * 512 + lane + 32 * i] = cvv[i];
          lnvec[b * 512 + 256 + lane + 32 * i] = bvv[i];
        }
      }
      float sm[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
      // 32 loads in flight per lane (64 registers; this warp has nothing else to hold): two L2 round trips for a 320-wide row, three for
      // 640, five for 1280, all of them one unit ahead of the epilogue
      constexpr int LNB = 8;
      for (int c = 0; c < p.ln_parts; c += LNB) {
        float2 v[LNB][4];
#pragma unroll
        for (int cc = 0; cc < LNB; ++cc) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = t.grow0 + lane + 32 * r;
            v[cc][r] = (c + cc < p.ln_parts && row < p.M) ? __ldcg(p.ln_stats + static_cast<size_t>(c + cc) * p.M + row) : make_float2(0.f, 0.f);
          }
        }
#pragma unroll
        for (int cc = 0; cc < LNB; ++cc) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            sm[r] += v[cc][r].x;
            sq[r] += v[cc][r].y;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float mu = sm[r] * p.ln_inv_k;
        const float rs = rsqrtf(fmaxf(sq[r] * p.ln_inv_k - mu * mu, 0.f) + p.ln_eps);
        rowstat[b * 128 + lane + 32 * r] = make_float2(rs, -mu * rs);
      }
      mbar_arrive(&st_full[b]);
      ++k;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // The weights do not depend on the previous kernel: the W tiles of the first ring slots are pulled into L2 BEFORE
      // the grid-dependency wait, so their HBM latency overlaps the predecessor's tail.
      if (n_local > 0) {
        const Unit t0 = decode_unit(p, first, pair_rank);
        const int npre = min(p.stages, t0.kb1 - t0.kb0);
        KbIter ki;
        ki.seek(p, t0.kb0);
        for (int i = 0; i < npre; ++i) {
          if (PAIR && WIDE) {
            tma_prefetch_l2_2d(&tmB, ki.w_col, t0.n_tile * p.BN + pair_rank * (p.BN / 4));
            tma_prefetch_l2_2d(&tmB, ki.w_col, t0.n_tile * p.BN + p.BN / 2 + pair_rank * (p.BN / 4));
          } else {
            tma_prefetch_l2_2d(&tmB, ki.w_col, t0.n_tile * p.BN + (PAIR ? pair_rank * b_rows : 0));
          }
          ki.next(p);
        }
      }
      pdl_wait();
      MVD_TR(2);
      int s = 0;          // ring slot and its phase, advanced without divisions
      uint32_t ph = 0;
      // W rows this CTA of a pair stages: [rank b_rows, +b_rows), or for a wide tile two boxes of BN / 4 rows (loop constants: the
      // per-k-block path of this thread is latency-critical — deep-K convolutions lost 7 % to one more branch with divisions here)
      constexpr bool w_wide = PAIR && WIDE;
      const int w_row0 = PAIR ? (w_wide ? pair_rank * (p.BN >> 2) : pair_rank * b_rows) : 0;
      const int w_row1 = (p.BN >> 1) + pair_rank * (p.BN >> 2);
      const int w_half_bytes = (p.BN >> 2) * 128;
      for (int j = 0; j < n_local; ++j) {
        const Unit t = decode_unit(p, first + j * ustride, pair_rank);
        KbIter ki;
        ki.seek(p, t.kb0);
        if (j == 1) MVD_TR(3);  // every load of the first unit has been issued
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_expect_tx(&full_bar[s], PAIR ? 2 * stage_bytes : stage_bytes);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + A_BYTES;
          if (PAIR) {
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (p.a_mode == MVD_A_CONV3X3) tma_load_4d_pair(sa, &tmA, lbar, ki.a_col, t.x0 * p.cstride + ki.kx - t.ox, t.y0 * p.cstride + ki.ky - t.oy, t.img0);
            else tma_load_2d_pair(sa, &tmA, lbar, ki.a_col, t.m_tile * BM);
            tma_load_2d_pair(sb, &tmB, lbar, ki.w_col, t.n_tile * p.BN + w_row0);
            if (w_wide) tma_load_2d_pair(sb + w_half_bytes, &tmB, lbar, ki.w_col, t.n_tile * p.BN + w_row1);
          } else {
            if (p.a_mode == MVD_A_CONV3X3) tma_load_4d(sa, &tmA, &full_bar[s], ki.a_col, t.x0 * p.cstride + ki.kx - t.ox, t.y0 * p.cstride + ki.ky - t.oy, t.img0);
            else tma_load_2d(sa, &tmA, &full_bar[s], ki.a_col, t.m_tile * BM);
            tma_load_2d(sb, &tmB, &full_bar[s], ki.w_col, t.n_tile * p.BN);
          }
          ki.next(p);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    if (lane == 0 && leader) {
      constexpr bool wide = PAIR && WIDE;
      const uint32_t idesc = umma_idesc_f16(PAIR ? 2 * BM : BM, wide ? p.BN / 2 : p.BN);
      const uint32_t wide_b = static_cast<uint32_t>((p.BN / 4) * 128) >> 4;  // descriptor offset of the W rows of the second MMA
      int s = 0;
      uint32_t ph = 0;
      const uint64_t da0 = umma_desc_sw128(smem_u32(smem));
      const uint64_t stage16 = static_cast<uint64_t>(stage_bytes >> 4);
      uint64_t da_cur = da0;
      for (int j = 0; j < n_local; ++j) {
        const Unit t = decode_unit(p, first + j * ustride, pair_rank);
        const int buf = j & 1;
        mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);
        // one accumulator only: the epilogue of the previous unit must have read it (it alternates the two barrier pairs all the same)
        if (wide && j >= 1) mbar_wait(&acc_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * p.acc_stride;
        // The issue loop is one thread whose instruction latencies add up against ~450 clk of MMA per k-block: operand descriptors advance
        // by adds (a stage is stage_bytes >> 4 in the descriptor's address field), the accumulate flag is a register
        uint32_t accum = 0u;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (j == 0 && kb == t.kb0) MVD_TR(4);
          const uint64_t da = da_cur, db = da_cur + (A_BYTES >> 4);
          if (PAIR && wide) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, k == 0 ? accum : 1u);
              umma_f16_pair(d_tmem + p.BN / 2, da + 2 * k, db + 2 * k + wide_b, idesc, k == 0 ? accum : 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 fp16 = 32 B inside the 128-B swizzle atom: +2 in the (addr >> 4) field
              if (PAIR) umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, k == 0 ? accum : 1u);
              else umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, k == 0 ? accum : 1u);
            }
          }
          accum = 1u;
          if (PAIR) tc_commit_pair(&empty_bar[s], 3);  // frees the slot in BOTH CTAs
          else tc_commit(&empty_bar[s]);
          da_cur += stage16;
          if (++s == p.stages) { s = 0; ph ^= 1; da_cur = da0; }
        }
        if (PAIR) tc_commit_pair(&acc_full[buf], 3);
        else tc_commit(&acc_full[buf]);
        if (j == 0) MVD_TR(5);
        if (j == n_local - 1) MVD_TR(6);
      }
    }
  } else if (TMAE && warp == 2) {
    // ------------------------------------------------------------ TMAE: residual chunks -> slots, in the order the warpgroups take them
    if (lane == 0 && has_res) {
      const int n_out = p.N, out_bn = p.BN;
      const int nchunks = (out_bn + 31) / 32;
      const uint32_t slots = smem_u32(out_stg);
      pdl_wait();
      int gbase = 0;  // chunks of this CTA before unit j; chunk g goes to warpgroup g % NWG as its (g / NWG)-th
      for (int j = 0; j < n_local; ++j) {
        const Unit t = decode_unit(p, first + j * ustride, pair_rank);
        if (PAIR && t.m_tile >= p.tiles_m_real) continue;
        const int valid = min(nchunks, (n_out - t.n_tile * out_bn + 31) / 32);
        for (int c = 0; c < valid; ++c) {
          const int g = gbase + c;
          const int w = g % NWG, k = g / NWG;
          const int sl = w * p.spw + (p.spw == 2 ? (k & 1) : 0);
          const int nth = p.spw == 2 ? (k >> 1) : k;  // how many times this slot has been filled before
          if (nth >= 1) mbar_wait(&res_empty[sl], (nth - 1) & 1);
          mbar_expect_tx(&res_full[sl], 128 * 128);
          tma_load_2d_raw(slots + sl * p.slot_bytes, &tmRes, &res_full[sl], t.n_tile * out_bn + c * 32, t.grow0);
        }
        gbase += valid;
      }
    } else if (!has_res && act == MVD_ACT_GEGLU && p.ln_stats != nullptr) {
      ln_helper(false);  // (a GEGLU GEMM has no residual to load: the warp is free; its vectors ride in the bias staging)
    }
  } else if (TMAE && warp == 3) {
    // ------------------------------------------------------------ TMAE: finished chunks -> global memory (TMA stores), slots handed back
    if (lane == 0) {
      const bool geglu = (act == MVD_ACT_GEGLU);
      const bool f16out = (out_mode == MVD_OUT_F16);
      const int n_out = geglu ? p.N / 2 : p.N;
      const int out_bn = geglu ? p.BN / 2 : p.BN;
      const int nchunks = (out_bn + 31) / 32;
      const uint32_t slots = smem_u32(out_stg);
      const uint32_t f16_off = (f16out && !has_res) ? 0u : static_cast<uint32_t>(p.slot_h16);
      pdl_wait();
      int gbase = 0, prev_sl = -1;  // prev_sl: slot of the store committed last, not yet handed back
      for (int j = 0; j < n_local; ++j) {
        const Unit t = decode_unit(p, first + j * ustride, pair_rank);
        if (PAIR && t.m_tile >= p.tiles_m_real) continue;
        const int valid = min(nchunks, (n_out - t.n_tile * out_bn + 31) / 32);
        for (int c = 0; c < valid; ++c) {
          const int g = gbase + c;
          const int w = g % NWG, k = g / NWG;
          const int sl = w * p.spw + (p.spw == 2 ? (k & 1) : 0);
          const uint32_t par = static_cast<uint32_t>((p.spw == 2 ? (k >> 1) : k) & 1);
          if (prev_sl >= 0 && !mbar_try_wait(&out_ready[sl], par)) {  // nothing to store yet: hand the last slot back right away
            tma_store_wait_read<0>();
            mbar_arrive(&res_empty[prev_sl]);
            prev_sl = -1;
          }
          mbar_wait(&out_ready[sl], par);
          const uint32_t slot = slots + sl * p.slot_bytes;
          const int oc = t.n_tile * out_bn + c * 32;
          if (f16out) {
            tma_store_2d_raw(&tmOut, slot + f16_off, oc, t.grow0);
          } else {
            tma_store_2d_raw(&tmOut, slot, oc, t.grow0);
            if (p.out16 != nullptr) tma_store_2d_raw(&tmO16, slot + f16_off, oc, t.grow0);
            if (p.out16_lo > 0) tma_store_2d_raw(&tmO16lo, slot + p.slot_lo, oc, t.grow0);
          }
          tma_store_commit();
          if (prev_sl >= 0) {  // the store before this one has read its slot
            tma_store_wait_read<1>();
            mbar_arrive(&res_empty[prev_sl]);
          }
          prev_sl = sl;
        }
        gbase += valid;
      }
      tma_store_wait_read<0>();  // shared memory stays valid until the stores have read it; the writes complete with the grid
    }
  } else if (TMAE && warp >= 4) {
    // ------------------------------------------------------------ TMAE epilogue: thread = tile row, chunks of 32 columns
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int et = q * 32 + lane;
    const int bar_id = 1 + wg;
    const bool geglu = (act == MVD_ACT_GEGLU);
    const bool f16out = (out_mode == MVD_OUT_F16);
    const int n_out = geglu ? p.N / 2 : p.N;
    const int out_bn = geglu ? p.BN / 2 : p.BN;
    const int nchunks = (out_bn + 31) / 32;
    const uint32_t slot_wg = smem_u32(out_stg) + wg * p.spw * p.slot_bytes;
    const uint32_t sw128 = static_cast<uint32_t>(et & 7), sw64 = static_cast<uint32_t>((et >> 1) & 3);
    const uint32_t row128 = et * 128, row64 = et * 64;
    const uint32_t f16_off = (f16out && !has_res) ? 0u : static_cast<uint32_t>(p.slot_h16);  // where the fp16 tile of a slot lives
    float* sbias = bias_smem + wg * 640;   // [bias of the tile's BN (<= 320) columns | column scale]
    const bool ln_in = geglu && p.ln_stats != nullptr;  // LayerNorm folded into this GEMM: its column sums ride in the column-scale half
    const bool use_sb = (p.bias != nullptr || p.colscale != nullptr || ln_in) && out_mode != MVD_OUT_QKV_HEADS;
    int sb_tile = -1;
    // the tile's bias / gate (or LayerNorm column-sum) columns -> shared vector; every load is issued before the first store, so a
    // restaged tile costs ONE L2 round trip (load - store - load - store chains cost the folded GEGLU GEMM ~1 us per unit)
    auto stage_vectors = [&](int n_tile) {
      const int nb = n_tile * p.BN;
      const float* second = ln_in ? p.ln_colsum : p.colscale;
      float bv[3], cv[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int kk = et + i * WG_THREADS;
        const bool in = kk < p.BN && nb + kk < p.N;
        bv[i] = (p.bias != nullptr && in) ? __ldg(p.bias + nb + kk) : 0.f;
        cv[i] = (second != nullptr && in) ? __ldg(second + nb + kk) : (ln_in ? 0.f : 1.f);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int kk = et + i * WG_THREADS;
        if (kk < p.BN) {
          if (p.bias != nullptr) sbias[kk] = bv[i];
          if (second != nullptr) sbias[320 + kk] = cv[i];
        }
      }
    };
    if (use_sb && n_local > 0) {            // bias and gate vectors are parameters: they may be fetched before the grid dependency resolves
      const Unit t = decode_unit(p, first, pair_rank);
      stage_vectors(t.n_tile);
      sb_tile = t.n_tile;
      named_bar_sync(bar_id, WG_THREADS);
    }
    pdl_wait();
    float ln_rs = 1.f, ln_nm = 0.f;  // rstd and -mean * rstd of this thread's row (ln_in)
    int kln = 0;                     // units whose statistics have been consumed
    if (warp == 4 && lane == 0) MVD_TR(7);
    int gbase = 0;
    for (int j = 0; j < n_local; ++j) {
      const Unit t = decode_unit(p, first + j * ustride, pair_rank);
      const int buf = j & 1;
      if (PAIR && t.m_tile >= p.tiles_m_real) {  // padding half of the last pair: only the accumulator hand-shake
        mbar_wait(&acc_full[buf], (j >> 1) & 1);
        tc_fence_before();
        mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
        continue;
      }
      const int valid = min(nchunks, (n_out - t.n_tile * out_bn + 31) / 32);
      // the unit's bias (and adaLN gate) columns -> this warpgroup's shared vector while the MMAs are still running: a first-touch
      // global load in front of the first chunk costs ~0.8 us on the critical path of every short GEMM
      if (use_sb && t.n_tile != sb_tile) {
        named_bar_sync(bar_id, WG_THREADS);  // everybody is done with the previous unit's vector
        stage_vectors(t.n_tile);
        sb_tile = t.n_tile;
        named_bar_sync(bar_id, WG_THREADS);
      }
      const int grow = t.grow0 + et;
      if (ln_in) {  // {rstd, -mean * rstd} of this thread's row, reduced by warp 2 while the previous unit was being finished
        const int b = kln & 1;
        mbar_wait(&st_full[b], (kln >> 1) & 1);
        const float2 rsn = rowstat[b * 128 + et];
        ln_rs = rsn.x;
        ln_nm = rsn.y;
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_empty[b]);  // one arrival per warp: its 32 reads were issued before it (shared-memory accesses of a warp stay in order)
        ++kln;
      }
      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      if (j == 0 && warp == 4 && lane == 0) MVD_TR(8);
      if (j == n_local - 1 && warp == 4 && lane == 0) MVD_TR(9);
      const uint32_t taddr = tmem_base + buf * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
      int c0 = (wg - gbase) % NWG;
      if (c0 < 0) c0 += NWG;
      for (int c = c0; c < valid; c += NWG) {
        const int k = (gbase + c) / NWG;  // this warpgroup's chunk counter
        const int sl = p.spw == 2 ? (k & 1) : 0;
        const int nth = p.spw == 2 ? (k >> 1) : k;  // how often this slot has been used before
        const uint32_t slot = slot_wg + sl * p.slot_bytes;
        const int oc = t.n_tile * out_bn + c * 32;
        float v[32];
        if (geglu) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float g[16];
            tmem_ld16(taddr + c * 32 + hh * 16, v + hh * 16);
            tmem_ld16(taddr + p.BN / 2 + c * 32 + hh * 16, g);
            tmem_ld_wait();
            if (ln_in) {  // LayerNorm(x) W'^T + b' from the raw-x product: rstd * acc + (-mean * rstd * colsum + b'), two FMAs per element
              const uint32_t cv = smem_u32(sbias + 320 + c * 32 + hh * 16), cg = smem_u32(sbias + 320 + p.BN / 2 + c * 32 + hh * 16);
              const uint32_t sv = smem_u32(sbias + c * 32 + hh * 16), sg = smem_u32(sbias + p.BN / 2 + c * 32 + hh * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a4 = lds_v4(cv + i * 16), g4 = lds_v4(cg + i * 16);
                float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
                if (p.bias != nullptr) {
                  ba = lds_v4(sv + i * 16);
                  bg = lds_v4(sg + i * 16);
                }
                v[hh * 16 + 4 * i] = fmaf(ln_rs, v[hh * 16 + 4 * i], fmaf(ln_nm, a4.x, ba.x)); v[hh * 16 + 4 * i + 1] = fmaf(ln_rs, v[hh * 16 + 4 * i + 1], fmaf(ln_nm, a4.y, ba.y));
                v[hh * 16 + 4 * i + 2] = fmaf(ln_rs, v[hh * 16 + 4 * i + 2], fmaf(ln_nm, a4.z, ba.z)); v[hh * 16 + 4 * i + 3] = fmaf(ln_rs, v[hh * 16 + 4 * i + 3], fmaf(ln_nm, a4.w, ba.w));
                g[4 * i] = fmaf(ln_rs, g[4 * i], fmaf(ln_nm, g4.x, bg.x)); g[4 * i + 1] = fmaf(ln_rs, g[4 * i + 1], fmaf(ln_nm, g4.y, bg.y));
                g[4 * i + 2] = fmaf(ln_rs, g[4 * i + 2], fmaf(ln_nm, g4.z, bg.z)); g[4 * i + 3] = fmaf(ln_rs, g[4 * i + 3], fmaf(ln_nm, g4.w, bg.w));
              }
            } else if (p.bias != nullptr) {
              const uint32_t sv = smem_u32(sbias + c * 32 + hh * 16), sg = smem_u32(sbias + p.BN / 2 + c * 32 + hh * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 a4 = lds_v4(sv + i * 16), g4 = lds_v4(sg + i * 16);
                v[hh * 16 + 4 * i] += a4.x; v[hh * 16 + 4 * i + 1] += a4.y; v[hh * 16 + 4 * i + 2] += a4.z; v[hh * 16 + 4 * i + 3] += a4.w;
                g[4 * i] += g4.x; g[4 * i + 1] += g4.y; g[4 * i + 2] += g4.z; g[4 * i + 3] += g4.w;
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[hh * 16 + i] *= gelu_erf(g[i]);
          }
        } else {
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
        }
        if (c + NWG >= valid) {  // last TMEM read of this unit by this thread: hand the accumulator back
          tc_fence_before();
          if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
          else mbar_arrive(&acc_empty[buf]);
        }
        if (!geglu) {
          const bool full = oc + 32 <= n_out;
          if (p.bias != nullptr) {
            const uint32_t sv = smem_u32(sbias + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = lds_v4(sv + i * 16);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          }
          if (p.rowbias != nullptr && grow < p.M) {
            const float* rbp = p.rowbias + static_cast<size_t>(p.fd_rpg.div(grow)) * p.N + oc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (full || oc + 4 * i + 4 <= n_out) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(rbp) + i);
                v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
              }
            }
          }
          if (act == MVD_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
          } else if (act == MVD_ACT_SILU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = silu(v[i]);
          }
          if (p.colscale != nullptr) {
            const uint32_t sv = smem_u32(sbias + 320 + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 c4 = lds_v4(sv + i * 16);
              v[4 * i] *= c4.x; v[4 * i + 1] *= c4.y; v[4 * i + 2] *= c4.z; v[4 * i + 3] *= c4.w;
            }
          }
        }
        if (has_res) {
          mbar_wait(&res_full[wg * p.spw + sl], nth & 1);  // the residual chunk has landed (and the slot was free)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 r4 = lds_v4(slot + row128 + ((static_cast<uint32_t>(i) ^ sw128) << 4));
            v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
          }
        } else if (nth >= 1) {
          mbar_wait(&res_empty[wg * p.spw + sl], (nth - 1) & 1);  // the store that last used this slot has read it
        }
        if (!f16out && p.ln_stats_out != nullptr) {  // what the LayerNorm behind this output needs: (sum, sum of squares) of the chunk
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s4[i & 3] += v[i];
            q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]);
          }
          if (grow < p.M) p.ln_stats_out[static_cast<size_t>(oc >> 5) * p.M + grow] = make_float2((s4[0] + s4[1]) + (s4[2] + s4[3]), (q4[0] + q4[1]) + (q4[2] + q4[3]));
        }
        if (!f16out) {
#pragma unroll
          for (int i = 0; i < 8; ++i) sts_v4(slot + row128 + ((static_cast<uint32_t>(i) ^ sw128) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (f16out || p.out16 != nullptr) {
          const uint32_t hrow = slot + f16_off + row64;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts_v4u(hrow + ((static_cast<uint32_t>(i) ^ sw64) << 4), pack_h2(v[8 * i], v[8 * i + 1]), pack_h2(v[8 * i + 2], v[8 * i + 3]),
                    pack_h2(v[8 * i + 4], v[8 * i + 5]), pack_h2(v[8 * i + 6], v[8 * i + 7]));
          if (p.out16_lo > 0) {  // what the fp16 rounding dropped, as a second fp16 tile
            const uint32_t lrow = slot + p.slot_lo + row64;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] -= __half2float(__float2half_rn(v[i]));
#pragma unroll
            for (int i = 0; i < 4; ++i)
              sts_v4u(lrow + ((static_cast<uint32_t>(i) ^ sw64) << 4), pack_h2(v[8 * i], v[8 * i + 1]), pack_h2(v[8 * i + 2], v[8 * i + 3]),
                      pack_h2(v[8 * i + 4], v[8 * i + 5]), pack_h2(v[8 * i + 6], v[8 * i + 7]));
          }
        }
        fence_async_smem();                          // generic-proxy writes -> visible to the TMA store
        mbar_arrive(&out_ready[wg * p.spw + sl]);    // warp 3 stores the chunk and hands the slot back; no warpgroup barrier per chunk
      }
      if (c0 >= valid) {  // nothing to finish in this unit: still part of the accumulator hand-shake
        tc_fence_before();
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
        else mbar_arrive(&acc_empty[buf]);
      }
      gbase += valid;
    }
    if (lane == 0 && (warp & 3) == 0) MVD_TR(10);
  } else if (!TMAE && warp == 2) {
    if (out_mode == MVD_OUT_QKV_HEADS && p.ln_stats != nullptr) ln_helper(true);
  } else if (!TMAE && warp >= 4) {
    // ------------------------------------------------------------ epilogue: 2 warpgroups x 128 threads
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;              // TMEM lane quadrant this warp may read
    const int et = q * 32 + lane;        // thread within the warpgroup == tile row in phase A
    const int seg = et & 7;              // phase B: 16-byte segment of the 32-column chunk
    const int rb = et >> 3;              // phase B: rows rb, rb + 16, ..., rb + 112
    const int bar_id = 1 + wg;
    const bool geglu = (act == MVD_ACT_GEGLU);
    const int n_out = geglu ? p.N / 2 : p.N;       // output columns
    const int out_bn = geglu ? p.BN / 2 : p.BN;    // output columns per tile
    const int nchunks = (out_bn + 31) / 32;        // a narrow tile (BN < 32) still takes one chunk; guards clip it
    const int ws_ld = nchunks * 32;                // leading dimension of one split-K partial tile
    const uint32_t stg_base = smem_u32(out_stg) + wg * STG_PER_WG * STG_BYTES;
    float* sbias = bias_smem + wg * 256;
    const int inner = p.heads * p.dhead;
    int n_staged = 0;  // staging buffer toggle
    pdl_wait();        // residual / split-K workspace reads and every output write come after the predecessor grid
    if (warp == 4 && lane == 0) MVD_TR(7);
    // LayerNorm folded into this GEMM (QKV head scatter): rstd and -mean * rstd of this thread's tile row, set per unit below
    const bool ln_in = out_mode == MVD_OUT_QKV_HEADS && p.ln_stats != nullptr;
    float ln_rs = 1.f, ln_nm = 0.f;
    int kln = 0, ln_b = 0;

    // ---- phase A of one chunk: TMEM -> registers -> (GEGLU) -> swizzled staging tile, or the direct QKV scatter
    auto phase_a = [&](const Unit& t, uint32_t taddr, int c, uint32_t stg) -> bool {
      float v[32];
      const int oc = t.n_tile * out_bn + c * 32;
      if (geglu && NWG == 3) {
        // three warpgroups run at 128 registers: the value / gate columns pass through in two 16-column halves
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float g[16];
          tmem_ld16(taddr + c * 32 + hh * 16, v + hh * 16);
          tmem_ld16(taddr + p.BN / 2 + c * 32 + hh * 16, g);
          tmem_ld_wait();
          if (p.bias != nullptr) {
            const uint32_t sv = smem_u32(sbias + c * 32 + hh * 16), sg = smem_u32(sbias + p.BN / 2 + c * 32 + hh * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bv = lds_v4(sv + i * 16), bg = lds_v4(sg + i * 16);
              v[hh * 16 + 4 * i] += bv.x; v[hh * 16 + 4 * i + 1] += bv.y; v[hh * 16 + 4 * i + 2] += bv.z; v[hh * 16 + 4 * i + 3] += bv.w;
              g[4 * i] += bg.x; g[4 * i + 1] += bg.y; g[4 * i + 2] += bg.z; g[4 * i + 3] += bg.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[hh * 16 + i] *= gelu_erf(g[i]);
        }
      } else if (geglu) {
        float g[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld32(taddr + p.BN / 2 + c * 32, g);
        tmem_ld_wait();
        if (p.bias != nullptr) {
          const uint32_t sv = smem_u32(sbias + c * 32), sg = smem_u32(sbias + p.BN / 2 + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bv = lds_v4(sv + i * 16), bg = lds_v4(sg + i * 16);
            v[4 * i] += bv.x; v[4 * i + 1] += bv.y; v[4 * i + 2] += bv.z; v[4 * i + 3] += bv.w;
            g[4 * i] += bg.x; g[4 * i + 1] += bg.y; g[4 * i + 2] += bg.z; g[4 * i + 3] += bg.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= gelu_erf(g[i]);
      } else {
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (ln_in) {  // LayerNorm(x) W'^T + b' from the raw-x product: rstd * acc + (-mean * rstd * colsum + b') (N % 32 == 0: whole chunks)
          const uint32_t cv = smem_u32(lnvec + ln_b * 512 + c * 32), bv = cv + 256 * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 c4 = lds_v4(cv + i * 16), b4 = lds_v4(bv + i * 16);
            v[4 * i] = fmaf(ln_rs, v[4 * i], fmaf(ln_nm, c4.x, b4.x)); v[4 * i + 1] = fmaf(ln_rs, v[4 * i + 1], fmaf(ln_nm, c4.y, b4.y));
            v[4 * i + 2] = fmaf(ln_rs, v[4 * i + 2], fmaf(ln_nm, c4.z, b4.z)); v[4 * i + 3] = fmaf(ln_rs, v[4 * i + 3], fmaf(ln_nm, c4.w, b4.w));
          }
        }
      }
      if (out_mode == MVD_OUT_QKV_HEADS && (p.qkv_direct || oc >= 2 * inner)) {
        // v^T (keys contiguous) wants thread = row: leave straight from the registers
        const int grow = t.grow0 + et;
        if (grow < p.M) {
          int img, pos;
          p.fd_seq.divmod(grow, img, pos);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const int n = oc + i;
            if (n >= p.N) continue;
            if (p.bias != nullptr && !ln_in) {  // dhead % 8 == 0 and N = 3 * heads * dhead: eight whole columns exist (ln_in: added above)
              if (p.vec_bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + 1);
                v[i] += b0.x; v[i + 1] += b0.y; v[i + 2] += b0.z; v[i + 3] += b0.w;
                v[i + 4] += b1.x; v[i + 5] += b1.y; v[i + 6] += b1.z; v[i + 7] += b1.w;
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[i + e] += __ldg(p.bias + n + e);
              }
            }
            int which, rem, h, jj;
            p.fd_inner.divmod(n, which, rem);
            p.fd_dhead.divmod(rem, h, jj);
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
        return false;
      }
      const uint32_t srow = stg + et * 128;
#pragma unroll
      for (int i = 0; i < 8; ++i) sts_v4(srow + ((i ^ (et & 7)) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      return true;
    };

    // ---- residual of one chunk, phase-B mapping (issued one chunk ahead; consumed in phase B)
    float4 rn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto res_load = [&](int grow0, int oc, int i) -> float4 {
      const int row = grow0 + rb + 16 * i;
      const int col = oc + seg * 4;
      if (row < p.M && col < p.N) return ld4(p.residual + static_cast<size_t>(row) * p.ldr + col, VEC || p.vec_res != 0, VEC ? 4 : p.N - col);
      return make_float4(0.f, 0.f, 0.f, 0.f);
    };

    // ---- phase B of one chunk.  part_tile/part_s: split-K partial tiles to add (nullptr = none); nxt_*: next chunk of
    //      this warpgroup (residual prefetch), nxt_valid = false when there is none.
    constexpr bool PREFETCH_RES = (NWG == 2);  // three warpgroups: a third chunk in flight hides the latency instead (and 32 registers less)
    auto phase_b = [&](const Unit& t, int c, uint32_t stg, bool nxt_valid, int nxt_grow0, int nxt_oc) {
      const int oc = t.n_tile * out_bn + c * 32;
      const int col = oc + seg * 4;
      if (!PREFETCH_RES && has_res) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rn[i] = res_load(t.grow0, oc, i);
      }
      const int nvalid = VEC ? ((n_out - col) > 0 ? 4 : 0) : (n_out - col);  // <= 0: this thread's columns do not exist
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
      if (!geglu && p.bias != nullptr && nvalid > 0 && !ln_in) b4 = ldg4(p.bias + col, VEC || p.vec_bias != 0, nvalid);
      if (p.colscale != nullptr && nvalid > 0) cs4 = ldg4(p.colscale + col, VEC || p.vec_colscale != 0, nvalid);
      // QKV (q / k part): per-thread head coordinates are fixed for the chunk
      int qk_which = 0, qk_h = 0, qk_jj = 0;
      if (out_mode == MVD_OUT_QKV_HEADS && nvalid > 0) {
        int rem;
        p.fd_inner.divmod(col, qk_which, rem);
        p.fd_dhead.divmod(rem, qk_h, qk_jj);
      }
      float4 acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = rb + 16 * i;
        acc[i] = lds_v4(stg + row * 128 + ((seg ^ (row & 7)) << 4));
      }
      if (is_split) {
        // the other slices' partials of this chunk: 16 independent 16-byte loads in flight per step
        const size_t tile_elems = static_cast<size_t>(BM) * ws_ld;
        const float* wbase = p.ws + static_cast<size_t>(t.tile) * p.split * tile_elems + rb * ws_ld + c * 32 + seg * 4;
        const int nother = p.split - 1;
        for (int o = 0; o < nother; o += 2) {
          const int sa = o < t.s ? o : o + 1;
          const bool two = o + 1 < nother;
          const int sb2 = two ? ((o + 1) < t.s ? o + 1 : o + 2) : sa;
          const float* pa = wbase + sa * tile_elems;
          const float* pb = wbase + sb2 * tile_elems;
          float4 wa[8], wb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            wa[i] = __ldcg(reinterpret_cast<const float4*>(pa + 16 * i * ws_ld));
            wb[i] = __ldcg(reinterpret_cast<const float4*>(pb + 16 * i * ws_ld));
          }
          const float fb = two ? 1.f : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i].x += wa[i].x + fb * wb[i].x; acc[i].y += wa[i].y + fb * wb[i].y;
            acc[i].z += wa[i].z + fb * wb[i].z; acc[i].w += wa[i].w + fb * wb[i].w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = rb + 16 * i;
        const int grow = t.grow0 + row;
        float4 v = acc[i];
        const bool live = (grow < p.M) && nvalid > 0;
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        if (p.rowbias != nullptr && live) {
          const float4 r4 = ldg4(p.rowbias + static_cast<size_t>(p.fd_rpg.div(grow)) * p.N + col, VEC || p.vec_rowbias != 0, nvalid);
          v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
        }
        if (act == MVD_ACT_GELU) {
          v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
        } else if (act == MVD_ACT_SILU) {
          v.x = silu(v.x); v.y = silu(v.y); v.z = silu(v.z); v.w = silu(v.w);
        }
        v.x *= cs4.x; v.y *= cs4.y; v.z *= cs4.z; v.w *= cs4.w;
        if (has_res) {
          v.x += rn[i].x; v.y += rn[i].y; v.z += rn[i].z; v.w += rn[i].w;
          if (PREFETCH_RES && nxt_valid) rn[i] = res_load(nxt_grow0, nxt_oc, i);
        }
        if (!live) continue;
        size_t orow = static_cast<size_t>(grow);
        int ocol = col;
        if (p.up2) {  // source pixel (img, y, x), phase (py, px) -> output pixel (img, 2y + py, 2x + px); columns of the phase block
          const int x = grow & ((1 << p.up_lw) - 1), y = (grow >> p.up_lw) & ((1 << p.up_lh) - 1), img = grow >> (p.up_lw + p.up_lh);
          orow = (((static_cast<size_t>(img) << (p.up_lh + 1)) + 2 * y + (t.ph >> 1)) << (p.up_lw + 1)) + 2 * x + (t.ph & 1);
          ocol = col - t.ph * p.n_real;
        }
        if (out_mode == MVD_OUT_F32) {
          float* dst = reinterpret_cast<float*>(p.out) + orow * p.ldc + ocol;
          if (VEC || (p.vec_out && nvalid >= 4)) {
            *reinterpret_cast<float4*>(dst) = v;
          } else {
            dst[0] = v.x;
            if (nvalid > 1) dst[1] = v.y;
            if (nvalid > 2) dst[2] = v.z;
            if (nvalid > 3) dst[3] = v.w;
          }
          if (p.out16 != nullptr) {  // the same values once more as the fp16 operand of the next GEMM
            __half* d16 = p.out16 + orow * p.ld16 + ocol;
            if (p.vec_out16 && nvalid >= 4) {
              *reinterpret_cast<uint2*>(d16) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
            } else {
              d16[0] = __float2half_rn(v.x);
              if (nvalid > 1) d16[1] = __float2half_rn(v.y);
              if (nvalid > 2) d16[2] = __float2half_rn(v.z);
              if (nvalid > 3) d16[3] = __float2half_rn(v.w);
            }
            if (p.out16_lo > 0) {  // what the fp16 rounding dropped, as a second fp16 (hi/lo operand of a split-precision GEMM)
              const float lx = v.x - __half2float(__float2half_rn(v.x)), ly = v.y - __half2float(__float2half_rn(v.y));
              const float lz = v.z - __half2float(__float2half_rn(v.z)), lw = v.w - __half2float(__float2half_rn(v.w));
              __half* dl = d16 + p.out16_lo;
              if (p.vec_out16 && nvalid >= 4) {
                *reinterpret_cast<uint2*>(dl) = make_uint2(pack_h2(lx, ly), pack_h2(lz, lw));
              } else {
                dl[0] = __float2half_rn(lx);
                if (nvalid > 1) dl[1] = __float2half_rn(ly);
                if (nvalid > 2) dl[2] = __float2half_rn(lz);
                if (nvalid > 3) dl[3] = __float2half_rn(lw);
              }
            }
          }
        } else if (out_mode == MVD_OUT_F16) {
          __half* dst = reinterpret_cast<__half*>(p.out) + orow * p.ldc + ocol;
          if (VEC || (p.vec_out && nvalid >= 4)) {
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
          } else {
            dst[0] = __float2half_rn(v.x);
            if (nvalid > 1) dst[1] = __float2half_rn(v.y);
            if (nvalid > 2) dst[2] = __float2half_rn(v.z);
            if (nvalid > 3) dst[3] = __float2half_rn(v.w);
          }
        } else {  // q / k part of the head scatter (dhead % 8 == 0: four columns never straddle a head)
          int img, pos;
          p.fd_seq.divmod(grow, img, pos);
          __half* base = reinterpret_cast<__half*>(qk_which == 0 ? p.out : p.out_k);
          __half* dst = base + ((static_cast<size_t>(img) * p.heads + qk_h) * p.seq + pos) * p.dpad + qk_jj;
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
        }
      }
    };

    // chunks of unit (j, t) this warpgroup finishes: c = cfirst, cfirst + cstride, ... < valid
    auto chunk_walk = [&](const Unit& t, int j, int& cfirst, int& cstride, int& valid) {
      valid = min(nchunks, (n_out - t.n_tile * out_bn + 31) / 32);
      if (is_split) {
        cfirst = t.s + p.split * wg;
        cstride = NWG * p.split;
      } else {
        cfirst = (wg + j) % NWG;
        cstride = NWG;
      }
    };
    // first chunk of this warpgroup at or after unit j (search forward); false when no work is left
    auto find_task = [&](int j, int& jt, int& ct, int& grow0, int& oc) -> bool {
      for (; j < n_local; ++j) {
        const Unit t = decode_unit(p, first + j * ustride, pair_rank);
        int cf, cs, vc;
        chunk_walk(t, j, cf, cs, vc);
        if (PAIR && t.m_tile >= p.tiles_m_real) continue;
        if (cf < vc) {
          jt = j; ct = cf; grow0 = t.grow0; oc = t.n_tile * out_bn + cf * 32;
          return true;
        }
      }
      return false;
    };

    if (PREFETCH_RES && has_res) {  // residual of the very first chunk
      int jt, ct, g0, oc0;
      if (find_task(0, jt, ct, g0, oc0)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rn[i] = res_load(g0, oc0, i);
      }
    }

    for (int j = 0; j < n_local; ++j) {
      const Unit t = decode_unit(p, first + j * ustride, pair_rank);
      const int u = t.tile * p.split + t.s;  // index of this slice's split-K partial tile
      const int buf = j & 1;
      int cfirst, cstride, valid_chunks;
      chunk_walk(t, j, cfirst, cstride, valid_chunks);
      if (PAIR && t.m_tile >= p.tiles_m_real) {  // padding half of the last pair: only the accumulator hand-shake
        mbar_wait(&acc_full[buf], (j >> 1) & 1);
        tc_fence_before();
        mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
        continue;
      }

      if (geglu && p.bias != nullptr) {  // tile bias -> smem (phase A reads it as broadcast vectors)
        for (int k = et; k < p.BN; k += WG_THREADS) sbias[k] = __ldg(p.bias + t.n_tile * p.BN + k);
        named_bar_sync(bar_id, WG_THREADS);
      }

      if (ln_in) {  // {rstd, -mean * rstd} of this thread's row, reduced by warp 2 while the previous unit was being finished
        const int b = kln & 1;
        mbar_wait(&st_full[b], (kln >> 1) & 1);
        const float2 rsn = rowstat[b * 128 + et];
        ln_b = b;
        ln_rs = rsn.x;
        ln_nm = rsn.y;
        ++kln;  // (rowstat[b] / lnvec[b] are handed back after this unit's last chunk: phase A reads the vectors)
      }
      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      if (j == 0 && warp == 4 && lane == 0) MVD_TR(8);
      if (j == n_local - 1 && warp == 4 && lane == 0) MVD_TR(9);
      const uint32_t taddr = tmem_base + buf * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);

      if (is_split) {
        // ---- park the chunks other slices own: ws[u][row][col] fp32, written with the coalesced phase-B mapping
        float* wsu = p.ws + static_cast<size_t>(u) * (BM * ws_ld);
        for (int c = wg; c < valid_chunks; c += NWG) {
          if (c % p.split == t.s) continue;
          const uint32_t stg = stg_base + (STG_PER_WG == 2 ? (n_staged & 1) * STG_BYTES : 0);
          ++n_staged;
          if (STG_PER_WG == 1) named_bar_sync(bar_id, WG_THREADS);  // the previous chunk's readers are done with the tile
          phase_a(t, taddr, c, stg);
          named_bar_sync(bar_id, WG_THREADS);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = rb + 16 * i;
            const float4 v = lds_v4(stg + row * 128 + ((seg ^ (row & 7)) << 4));
            __stcg(reinterpret_cast<float4*>(wsu + row * ws_ld + c * 32 + seg * 4), v);
          }
        }
        __threadfence();
        named_bar_sync(NWG + 1, EPI_THREADS);
        if (wg == 0 && et == 0) {
          atomicAdd(&p.counters[2 * t.tile], 1);
          while (ld_acquire(&p.counters[2 * t.tile]) < p.split) {
          }
        }
        named_bar_sync(NWG + 1, EPI_THREADS);
      }

      for (int c = cfirst; c < valid_chunks; c += cstride) {
        const uint32_t stg = stg_base + (STG_PER_WG == 2 ? (n_staged & 1) * STG_BYTES : 0);
        // next chunk of this warpgroup (this unit or a later one) for the residual prefetch
        bool nxt = false;
        int ng0 = 0, noc = 0;
        if (PREFETCH_RES && has_res) {
          if (c + cstride < valid_chunks) {
            nxt = true; ng0 = t.grow0; noc = t.n_tile * out_bn + (c + cstride) * 32;
          } else {
            int jt, ct;
            nxt = find_task(j + 1, jt, ct, ng0, noc);
          }
        }
        if (STG_PER_WG == 1) named_bar_sync(bar_id, WG_THREADS);  // the previous chunk's readers are done with the tile
        const bool staged = phase_a(t, taddr, c, stg);
        if (c + cstride >= valid_chunks) {  // last TMEM read of this unit by this thread: hand the accumulator back
          tc_fence_before();
          if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
          else mbar_arrive(&acc_empty[buf]);
        }
        if (staged) {
          ++n_staged;
          named_bar_sync(bar_id, WG_THREADS);
          phase_b(t, c, stg, nxt, ng0, noc);
        }
      }
      if (cfirst >= valid_chunks) {  // nothing to finish in this unit: still part of the accumulator hand-shake
        tc_fence_before();
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
        else mbar_arrive(&acc_empty[buf]);
      }
      if (ln_in) {  // one arrival per warp: its reads of rowstat[b] / lnvec[b] were issued before it (a warp's shared-memory accesses stay in order)
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_empty[ln_b]);
      }

      if (is_split) {
        named_bar_sync(NWG + 1, EPI_THREADS);  // every partial of this tile has been consumed by this CTA
        if (wg == 0 && et == 0) {
          const int old = atomicAdd(&p.counters[2 * t.tile + 1], 1);
          if (old == p.split - 1) {  // last slice out resets the semaphores for the next launch
            p.counters[2 * t.tile] = 0;
            p.counters[2 * t.tile + 1] = 0;
          }
        }
      }
    }
    if (lane == 0 && (warp & 3) == 0) MVD_TR(10);  // last warpgroup to get here wins: end of this CTA's epilogue work
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory and signal its barriers until the very end
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
#ifdef MVD_GEMM_TRACE
  if (threadIdx.x == 0) {
    MVD_TR(11);
    const unsigned idx = atomicAdd(&g_trace_n, 1u);
    if (idx < TRACE_CAP) {
      TraceRec& r = g_trace[idx];
      r.gt_entry = tr_gt0;
      r.gt_exit = trace_gtimer();
      for (int i = 0; i < 12; ++i) r.st[i] = tr_s[i];
      r.bid = blockIdx.x; r.grid = gridDim.x; r.M = p.M; r.N = p.N; r.num_kb = p.num_kb; r.BN = p.BN; r.split = p.split; r.n_local = n_local;
      r.flags = (PAIR ? 1 : 0) | (has_res ? 2 : 0) | (act << 4) | (out_mode << 8) | (p.a_mode << 12) | (NWG << 16);
      r.pad = 0;
    }
  }
#endif
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

// tile width.  N <= 64: one tile of N rounded up to 16.  Otherwise the multiple of 32 in [64, 256] with the lowest
// modelled time: rounds of the persistent loop x shared-memory rows filled per k-block (128 of A + bn of W; the
// mainloop is bound by the L2 -> smem fill rate on this part, ~56 B/clk/SM measured) — which accounts for both the
// wave quantisation over the SMs and the columns wasted by padding N up to a multiple of bn.
// tiles_m / slots count CTA pairs when `pair` (each SM of a pair stages 128 + bn/2 rows per k-block).
static int pick_bn(int N, int tiles_m, int slots, bool pair) {
  if (N <= 64) return pair ? (N + 31) / 32 * 32 : (N + 15) / 16 * 16;
  int best = 64;
  double best_cost = 1e300;
  for (int bn = 64; bn <= 256; bn += 32) {
    const int tiles = tiles_m * ((N + bn - 1) / bn);
    const int rounds = (tiles + slots - 1) / slots;
    const double rows = pair ? 128.0 + bn / 2 : 128.0 + bn;
    // a machine that is less than half full will be split along K afterwards: compare per-tile work then
    const double cost = (tiles * 2 <= slots) ? rows * tiles / slots * 1.15 : rounds * rows;
    if (cost < best_cost * 0.999 || (cost <= best_cost * 1.001 && bn > best)) {
      best = bn;
      best_cost = cost;
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
  const int hilo = a->hilo ? 1 : 0;
  if (hilo && ((a->K & 63) != 0 || (a->a_mode == MVD_A_CONV3X3 && (a->C & 63) != 0)))
    return set_error(MVD_EINVAL, "mvd_gemm_f16: hilo needs K (CONV3X3: C) to be a multiple of 64");
  if ((a->ldw & 7) != 0 || a->ldw < (hilo ? 2 * a->K : a->K)) return set_error(MVD_EALIGN, "mvd_gemm_f16: ldw must be >= K (2K with hilo) and a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->Wt) & 15))
    return set_error(MVD_EALIGN, "mvd_gemm_f16: A and Wt must be 16-byte aligned");
  if (a->act < MVD_ACT_NONE || a->act > MVD_ACT_GEGLU) return set_error(MVD_EINVAL, "mvd_gemm_f16: bad act");
  if (a->out_mode < MVD_OUT_F32 || a->out_mode > MVD_OUT_QKV_HEADS)
    return set_error(MVD_EINVAL, "mvd_gemm_f16: bad out_mode");

  // conv_up2: nearest x2 upsample folded into the convolution (mvd_b200.h, ABI 14).  N counts the four phase blocks; the output has N / 4 columns
  const bool up2 = a->conv_up2 != 0;
  if (up2) {
    if (a->a_mode != MVD_A_CONV3X3 || hilo || a->conv_stride == 2 || a->conv_no_pad_lo || (a->N & 3) != 0 || ((a->N / 4) & 3) != 0 ||
        a->residual != nullptr || a->rowbias != nullptr || a->colscale != nullptr || a->act != MVD_ACT_NONE || a->out_mode == MVD_OUT_QKV_HEADS ||
        a->ln_stats_out != nullptr || a->ln_stats != nullptr)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: conv_up2 needs a plain CONV3X3 (bias only, F32 / F16 output, N = 4 x output channels)");
    if (a->H <= 0 || (a->H & (a->H - 1)) != 0) return set_error(MVD_EINVAL, "mvd_gemm_f16: conv_up2 needs H and W to be powers of two");
  }
  const int n_cols = up2 ? a->N / 4 : a->N;  // columns of the output matrix

  GemmKParams p{};
  p.M = a->M;
  p.N = a->N;
  p.up2 = up2 ? 1 : 0;
  p.n_real = n_cols;
  p.taps_w = up2 ? 2 : 3;
  p.taps = up2 ? 4 : 9;
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
  p.out16 = static_cast<__half*>(a->out16);
  p.ld16 = a->ld16;
  p.out16_lo = a->out16 != nullptr ? a->out16_lo : 0;
  if (p.out16_lo < 0 || (p.out16_lo > 0 && (p.out16_lo < n_cols || (p.out16_lo & 3) != 0)))
    return set_error(MVD_EINVAL, "mvd_gemm_f16: out16_lo must be 0 or a multiple of 4 that is >= N");
  if (a->out16 != nullptr) {
    if (a->out_mode != MVD_OUT_F32 || a->act == MVD_ACT_GEGLU) return set_error(MVD_EINVAL, "mvd_gemm_f16: out16 accompanies an F32 output only");
    if (a->ld16 < n_cols) return set_error(MVD_EINVAL, "mvd_gemm_f16: ld16 is smaller than N");
    p.vec_out16 = (reinterpret_cast<uintptr_t>(a->out16) & 7) == 0 && (a->ld16 & 3) == 0;
  }

  p.ln_stats_out = reinterpret_cast<float2*>(a->ln_stats_out);
  p.ln_stats = reinterpret_cast<const float2*>(a->ln_stats);
  p.ln_colsum = a->ln_colsum;
  p.ln_eps = a->ln_eps;
  p.ln_inv_k = 1.0f / static_cast<float>(a->K);
  p.ln_parts = a->K / 32;
  if (a->ln_stats_out != nullptr) {
    if (a->out_mode != MVD_OUT_F32 || a->act == MVD_ACT_GEGLU || (a->N & 31) != 0 || (reinterpret_cast<uintptr_t>(a->ln_stats_out) & 7) != 0)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats_out accompanies an F32 output with N a multiple of 32 (8-byte aligned buffer)");
    if (a->split_k > 1) return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats_out excludes split_k");
  }
  if (a->ln_stats != nullptr) {
    if (a->ln_colsum == nullptr || !(a->ln_eps > 0.f)) return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats needs ln_colsum and a positive ln_eps");
    if (a->a_mode != MVD_A_ROWMAJOR || hilo || (a->K & 31) != 0 || (a->N & 31) != 0 ||
        !(a->out_mode == MVD_OUT_QKV_HEADS || (a->act == MVD_ACT_GEGLU && a->out_mode == MVD_OUT_F16)))
      return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats needs a row-major A, K and N multiples of 32 and a QKV_HEADS or GEGLU output");
    if ((reinterpret_cast<uintptr_t>(a->ln_stats) & 7) != 0 || (reinterpret_cast<uintptr_t>(a->ln_colsum) & 15) != 0)
      return set_error(MVD_EALIGN, "mvd_gemm_f16: ln_stats / ln_colsum must be 8 / 16-byte aligned");
    if (a->out_mode == MVD_OUT_QKV_HEADS && ((a->heads * a->dhead) & 31) != 0)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats with QKV_HEADS needs heads * dhead to be a multiple of 32");
  }

  const bool geglu = a->act == MVD_ACT_GEGLU;
  const int sms = num_sms();
  // m-tiles of the problem (conv tiles are whole image rows: see the geometry block below)
  int tiles_m_real;
  if (a->a_mode == MVD_A_CONV3X3) {
    if (a->n_img <= 0 || a->H <= 0 || a->W <= 0 || !is_pow2(a->W) || a->W > 4096) return set_error(MVD_EINVAL, "mvd_gemm_f16: bad CONV3X3 geometry");
    const int tw = a->W < BM ? a->W : BM;  // rows wider than a tile are cut into 128-pixel segments
    int th = BM / tw;
    if (th > a->H) th = a->H;
    const int tn = BM / (tw * th);
    tiles_m_real = (a->W / tw) * (a->H / th) * ((a->n_img + tn - 1) / tn);
  } else {
    tiles_m_real = (a->M + BM - 1) / BM;
  }
  // CTA pairs (cta_group::2): two m-tiles share one pass over the W tile — each SM stages 128 + bn/2 rows per k-block
  // instead of 128 + bn, which is what bounds the mainloop (L2 -> smem fill).  Used whenever the m-tiles pair up
  // with little padding; MVD_GEMM_NO_PAIR=1 in the environment turns it off (A/B measurements).
  static const bool no_pair = getenv("MVD_GEMM_NO_PAIR") != nullptr;
  // Measured on B200 (profiles/r01_gemm_native_bench_*.log): the pair wins when the mainloop dominates (deep K, or K >= 1280
  // with a wide N); short-K GEMMs are epilogue-bound and run better as independent CTAs.
  static const bool force_pair = getenv("MVD_GEMM_FORCE_PAIR") != nullptr;
  const bool deep = a->K >= 2048 || (a->K >= 1280 && a->N >= 2560);
  if (a->cta_pair < 0 || a->cta_pair > 2) return set_error(MVD_EINVAL, "mvd_gemm_f16: cta_pair must be 0, 1 or 2");
  if (a->cta_pair == 2 && (tiles_m_real < 2 || (sms & 1) != 0)) return set_error(MVD_EINVAL, "mvd_gemm_f16: cta_pair = 2 needs at least two m-tiles");
  const bool pair_auto = (deep || force_pair) && tiles_m_real >= 2 && ((tiles_m_real & 1) == 0 || tiles_m_real >= 9) && (sms & 1) == 0;
  const bool pair = !no_pair && (a->cta_pair == 2 || (a->cta_pair == 0 && pair_auto));
  const int tiles_mp = pair ? (tiles_m_real + 1) / 2 : tiles_m_real;
  const int slots = pair ? sms / 2 : sms;
  int bn = a->tile_n;
  if (bn == 0) bn = geglu ? 256 : pick_bn(a->N, tiles_mp, slots, pair);
  if (up2) {  // a tile stays inside one phase block: the widest multiple of 32 up to the chosen width that divides the output channels
    if (a->tile_n == 0) {
      while (bn > 32 && (n_cols % bn) != 0) bn -= 32;
    }
    if (bn > 256 || (bn & 31) != 0 || (n_cols % bn) != 0)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: conv_up2 needs tile_n to be a multiple of 32 (<= 256) that divides N / 4");
  }
  if (pair && (bn & 31) != 0) return set_error(MVD_EINVAL, "mvd_gemm_f16: tile_n must be a multiple of 32 here");
  const bool wide = bn > 256;
  if (wide && (!pair || bn > 320 || (bn & 63) != 0 || geglu || a->out_mode == MVD_OUT_QKV_HEADS))
    return set_error(MVD_EINVAL, "mvd_gemm_f16: tile_n > 256 needs cta_pair = 2, tile_n in {320}, a plain F32 / F16 output");
  p.wide = wide ? 1 : 0;
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

  CUtensorMap tmA, tmB;
  if (a->a_mode == MVD_A_ROWMAJOR) {
    if ((a->lda & 7) != 0 || a->lda < (hilo ? 2 * a->K : a->K)) return set_error(MVD_EALIGN, "mvd_gemm_f16: lda must be >= K (2K with hilo) and a multiple of 8");
    p.num_kb = (a->K + BK - 1) / BK;
    p.tiles_m_real = (a->M + BM - 1) / BM;
    p.a_lo_off = a->a_lo_off > 0 ? a->a_lo_off : a->K;  // A_lo may sit further right (a column window of a wider [hi | lo] buffer)
    p.w_lo_off = a->K;
    if (hilo && (p.a_lo_off < a->K || (p.a_lo_off & 7) != 0 || a->lda < p.a_lo_off + a->K))
      return set_error(MVD_EINVAL, "mvd_gemm_f16: a_lo_off must be a multiple of 8 with K <= a_lo_off <= lda - K");
    int rc = make_tmap_2d(&tmA, a->A, /*cols=*/hilo ? p.a_lo_off + a->K : a->K, /*rows=*/a->M, /*ld=*/a->lda, BK, BM);
    if (rc != MVD_OK) return rc;
  } else if (a->a_mode == MVD_A_CONV3X3) {
    if (a->n_img <= 0 || a->H <= 0 || a->W <= 0 || a->C <= 0 || (a->C & 7) != 0 || !is_pow2(a->W) || a->W > 4096 ||
        a->K != (up2 ? 4 : 9) * a->C || static_cast<long long>(a->M) != static_cast<long long>(a->n_img) * a->H * a->W)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: bad CONV3X3 geometry");
    p.n_img = a->n_img;
    p.H = a->H;
    p.W = a->W;
    p.C = a->C;
    p.tw = a->W < BM ? a->W : BM;  // W > 128 (the VAE decoder's 256-wide maps): a tile is a 128-pixel segment of one image row
    p.th = BM / p.tw;
    if (p.th > a->H) p.th = a->H;
    if (!is_pow2(p.th) || (a->H % p.th) != 0) return set_error(MVD_EINVAL, "mvd_gemm_f16: CONV3X3 needs H a multiple of the tile height");
    p.tn = BM / (p.tw * p.th);
    p.tiles_x = a->W / p.tw;
    p.tiles_y = a->H / p.th;
    const int tiles_z = (a->n_img + p.tn - 1) / p.tn;
    p.tiles_m_real = p.tiles_x * p.tiles_y * tiles_z;
    p.kb_per_tap = (a->C + BK - 1) / BK;
    p.num_kb = p.taps * p.kb_per_tap;
    p.a_lo_off = a->C;       // the image batch holds 2C channels: [hi | lo]
    p.w_lo_off = p.taps * a->C;
    if (a->conv_stride != 0 && a->conv_stride != 1 && a->conv_stride != 2) return set_error(MVD_EINVAL, "mvd_gemm_f16: conv_stride must be 1 or 2");
    p.cstride = a->conv_stride == 2 ? 2 : 1;
    p.cpad = a->conv_no_pad_lo ? 0 : 1;
    const int c_img = hilo ? 2 * a->C : a->C;
    if (a->lda != 0 && (a->lda < c_img || (a->lda & 7) != 0)) return set_error(MVD_EALIGN, "mvd_gemm_f16: CONV3X3 pixel pitch (lda) must be 0 or a multiple of 8 >= C");
    // H, W are the OUTPUT extent; the image holds (stride H) x (stride W) pixels
    int rc = make_tmap_nhwc(&tmA, a->A, a->n_img, a->H * p.cstride, a->W * p.cstride, c_img, BK, p.tw, p.th, p.tn, a->lda, p.cstride);
    if (rc != MVD_OK) return rc;
  } else {
    return set_error(MVD_EINVAL, "mvd_gemm_f16: bad a_mode");
  }
  {
    int rc = make_tmap_2d(&tmB, a->Wt, /*cols=*/hilo ? 2 * a->K : a->K, /*rows=*/a->N, /*ld=*/a->ldw, BK, pair ? (wide ? bn / 4 : bn / 2) : bn);
    if (rc != MVD_OK) return rc;
  }
  p.hilo = hilo;
  p.kb_seg = p.num_kb;
  if (hilo) p.num_kb *= 3;  // A_hi W_hi + A_lo W_hi + A_hi W_lo
  if (p.tiles_m_real != tiles_m_real) return set_error(MVD_EINVAL, "mvd_gemm_f16: internal tile count mismatch");
  p.tiles_m = tiles_mp;
  p.tiles_n = (a->N + bn - 1) / bn;
  const int tiles = p.tiles_m * p.tiles_n;  // work items per K slice: tiles, or 256-row tile pairs

  // ---- split-K (every slice of a tile must be co-resident: units <= SMs)
  int split = a->split_k;
  const size_t ws_avail = (a->splitk_ws != nullptr && a->splitk_ws_bytes > WS_COUNTER_BYTES) ? static_cast<size_t>(a->splitk_ws_bytes) - WS_COUNTER_BYTES : 0;
  const size_t tile_ws = static_cast<size_t>((bn + 31) / 32 * 32) * BM * sizeof(float);
  const bool can_split = !geglu && a->out_mode != MVD_OUT_QKV_HEADS;
  if (split <= 0) {  // auto: only when the tiles cannot fill half the machine and K is deep
    split = 1;
    if (can_split && tiles * 2 <= slots && p.num_kb >= 8) {
      split = slots / tiles;
      if (split > p.num_kb / 4) split = p.num_kb / 4;
      if (split > 16) split = 16;
      if (split < 1) split = 1;
    }
    while (split > 1 && (p.tiles_m_real * p.tiles_n > WS_COUNTER_BYTES / 8 || static_cast<size_t>(p.tiles_m_real) * p.tiles_n * split * tile_ws > ws_avail)) --split;
  } else if (split > 1) {
    if (!can_split) return set_error(MVD_EINVAL, "mvd_gemm_f16: split_k is not supported with GEGLU / QKV_HEADS");
    if (split > p.num_kb) split = p.num_kb;
    if (tiles * split > slots)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: split_k=%d x %d tiles exceeds the %d SMs / SM pairs (slices of a tile must be co-resident)", split, tiles, slots);
    if (p.tiles_m_real * p.tiles_n > WS_COUNTER_BYTES / 8 || static_cast<size_t>(p.tiles_m_real) * p.tiles_n * split * tile_ws > ws_avail)
      return set_error(MVD_EINVAL, "mvd_gemm_f16: split_k=%d needs a split-K workspace of %zu bytes (args.splitk_ws)", split,
                       static_cast<size_t>(tiles) * split * tile_ws + WS_COUNTER_BYTES);
  }
  if (split > p.num_kb) split = p.num_kb;
  p.kb_per_split = (p.num_kb + split - 1) / split;
  split = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.split = split;
  p.num_units = tiles * split;
  p.fd_split = make_fastdiv(split);
  p.fd_tiles_m = make_fastdiv(p.tiles_m);
  p.fd_tiles_x = make_fastdiv(p.tiles_x > 0 ? p.tiles_x : 1);
  p.fd_tiles_y = make_fastdiv(p.tiles_y > 0 ? p.tiles_y : 1);
  p.fd_seq = make_fastdiv(p.seq > 0 ? p.seq : 1);
  p.fd_inner = make_fastdiv(p.heads * p.dhead > 0 ? p.heads * p.dhead : 1);
  p.fd_dhead = make_fastdiv(p.dhead > 0 ? p.dhead : 1);
  p.fd_rpg = make_fastdiv(p.rows_per_group);
  p.fd_tpp = make_fastdiv(up2 ? n_cols / bn : 1);
  if (up2) {
    p.up_lw = 0;
    while ((1 << p.up_lw) < a->W) ++p.up_lw;
    p.up_lh = 0;
    while ((1 << p.up_lh) < a->H) ++p.up_lh;
  }
  if (split > 1) {
    if ((reinterpret_cast<uintptr_t>(a->splitk_ws) & 15) != 0) return set_error(MVD_EALIGN, "mvd_gemm_f16: splitk_ws must be 16-byte aligned");
    p.counters = reinterpret_cast<int*>(a->splitk_ws);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->splitk_ws) + WS_COUNTER_BYTES);
  }

  // ---- epilogue access widths
  const int n_out = geglu ? a->N / 2 : a->N;
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  if (a->out_mode != MVD_OUT_QKV_HEADS && a->ldc < (up2 ? n_cols : n_out)) return set_error(MVD_EINVAL, "mvd_gemm_f16: ldc is smaller than the output width");
  if (a->residual != nullptr && a->ldr < a->N) return set_error(MVD_EINVAL, "mvd_gemm_f16: ldr is smaller than N");
  p.vec_bias = al16(a->bias);
  p.vec_rowbias = al16(a->rowbias) && (a->N & 3) == 0;
  p.vec_colscale = al16(a->colscale);
  p.vec_res = al16(a->residual) && (a->ldr & 3) == 0;
  if (a->out_mode == MVD_OUT_F32) p.vec_out = al16(a->out) && (a->ldc & 3) == 0;
  else if (a->out_mode == MVD_OUT_F16) p.vec_out = (reinterpret_cast<uintptr_t>(a->out) & 7) == 0 && (a->ldc & 3) == 0;
  else {
    p.vec_out = 1;
    if (!al16(a->out) || !al16(a->out_k) || !al16(a->out_vt)) return set_error(MVD_EALIGN, "mvd_gemm_f16: q / k / v^T must be 16-byte aligned");
    p.qkv_direct = ((a->heads * a->dhead) & 31) != 0;
  }

  // ---- pick the epilogue specialisation
  const bool vec = (n_out % 4 == 0) && (a->bias == nullptr || p.vec_bias) && (a->rowbias == nullptr || p.vec_rowbias) &&
                   (a->colscale == nullptr || p.vec_colscale) && (a->residual == nullptr || p.vec_res) && p.vec_out &&
                   !(a->out_mode == MVD_OUT_QKV_HEADS && p.qkv_direct);
  const int has_res = a->residual != nullptr ? 1 : 0;
  const int is_split = split > 1 ? 1 : 0;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmKParams);
  struct Spec { int key; int nwg; KernelFn one, two, wide; };
#define MVD_SPEC(ACT, OUT, RES, SPL, NWG) \
  { (ACT) * 1000 + (OUT) * 100 + (RES) * 10 + (SPL), NWG, gemm_tc_kernel<ACT, OUT, RES, SPL, true, false, NWG, false>, gemm_tc_kernel<ACT, OUT, RES, SPL, true, true, NWG, false>, nullptr }
#define MVD_SPEC_TMAE(ACT, OUT, RES) \
  { (ACT) * 1000 + (OUT) * 100 + (RES) * 10, 3, gemm_tc_kernel<ACT, OUT, RES, 0, true, false, 3, true>, gemm_tc_kernel<ACT, OUT, RES, 0, true, true, 3, true>, \
    ((ACT) == MVD_ACT_NONE ? static_cast<KernelFn>(gemm_tc_kernel<MVD_ACT_NONE, OUT, RES, 0, true, true, 3, true, true>) : nullptr) }
  // three epilogue warpgroups where the epilogue fits 128 registers (ptxas -v: <= 110 with two warpgroups), two elsewhere
  static const Spec specs[] = {
      MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F32, 0, 0, 3),  MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F32, 0, 1, 2),  MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F32, 1, 0, 3),
      MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F32, 1, 1, 2),  MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F16, 0, 0, 3),  MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F16, 1, 0, 3),
      MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F16, 1, 1, 2),  MVD_SPEC(MVD_ACT_GELU, MVD_OUT_F16, 0, 0, 3),  MVD_SPEC(MVD_ACT_GELU, MVD_OUT_F32, 0, 0, 3),
      MVD_SPEC(MVD_ACT_GEGLU, MVD_OUT_F16, 0, 0, 3), MVD_SPEC(MVD_ACT_NONE, MVD_OUT_QKV_HEADS, 0, 0, 3),
  };
  // the unsplit F32 / F16 forms with the TMA epilogue (MVD_GEMM_NO_TMAE=1 in the environment keeps the thread-store one: A/B measurements)
  static const Spec specs_tmae[] = {
      MVD_SPEC_TMAE(MVD_ACT_NONE, MVD_OUT_F32, 0), MVD_SPEC_TMAE(MVD_ACT_NONE, MVD_OUT_F32, 1), MVD_SPEC_TMAE(MVD_ACT_NONE, MVD_OUT_F16, 0),
      MVD_SPEC_TMAE(MVD_ACT_NONE, MVD_OUT_F16, 1), MVD_SPEC_TMAE(MVD_ACT_GELU, MVD_OUT_F16, 0), MVD_SPEC_TMAE(MVD_ACT_GELU, MVD_OUT_F32, 0),
      MVD_SPEC_TMAE(MVD_ACT_GEGLU, MVD_OUT_F16, 0),
  };
  // the same keys with two warpgroups (MVD_GEMM_WG2=1 in the environment: A/B measurements)
  static const Spec specs_wg2[] = {
      MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F32, 0, 0, 2),  MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F16, 0, 0, 2),  MVD_SPEC(MVD_ACT_GELU, MVD_OUT_F16, 0, 0, 2),
      MVD_SPEC(MVD_ACT_GELU, MVD_OUT_F32, 0, 0, 2),  MVD_SPEC(MVD_ACT_NONE, MVD_OUT_QKV_HEADS, 0, 0, 2),
      MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F32, 1, 0, 2),  MVD_SPEC(MVD_ACT_NONE, MVD_OUT_F16, 1, 0, 2),  MVD_SPEC(MVD_ACT_GEGLU, MVD_OUT_F16, 0, 0, 2),
  };
#undef MVD_SPEC
#undef MVD_SPEC_TMAE
  static const Spec generic = {-1, 2, gemm_tc_kernel<-1, -1, -1, -1, false, false, 2, false>, gemm_tc_kernel<-1, -1, -1, -1, false, true, 2, false>, nullptr};
  static const bool force_wg2 = getenv("MVD_GEMM_WG2") != nullptr;
  static const bool no_tmae = getenv("MVD_GEMM_NO_TMAE") != nullptr;
  const Spec* spec = &generic;
  bool tmae = false;
  if (vec) {
    const int key = a->act * 1000 + a->out_mode * 100 + has_res * 10 + is_split;
    for (const Spec& sp : specs)
      if (sp.key == key) spec = &sp;
    if (force_wg2 && spec->nwg != 2)
      for (const Spec& sp : specs_wg2)
        if (sp.key == key) spec = &sp;
    // TMA epilogue: every tile the TMA engine touches must be describable — 16-byte aligned bases and row pitches
    // (A TMA form of the QKV head scatter — 8-column boxes per (q | k | v, head) — was measured and dropped: 16-byte-wide boxes
    // make the TMA engine write 128 separate 16-byte rows per box; 16384x960x320 went from 22.0 to 23.2 us.)
    // Deep-K GEMMs hide their epilogue behind the mainloop and lose 5-6 % with the TMA stores in flight (conv 16384x640x5760
    // 75.5 -> 80.7 us): they keep the thread-store epilogue.
    // (a wide pair tile runs one tile per CTA pair with nothing to overlap: it takes the TMA epilogue whatever its K)
    const bool tma_ok = !up2 && !is_split && a->out_mode != MVD_OUT_QKV_HEADS && (p.kb_per_split <= 24 || wide) && al16(a->out) &&
                        ((static_cast<long long>(a->ldc) * (a->out_mode == MVD_OUT_F32 ? 4 : 2)) & 15) == 0 &&
                        (a->residual == nullptr || (al16(a->residual) && (a->ldr & 3) == 0)) &&
                        (a->out16 == nullptr || (al16(a->out16) && (a->ld16 & 7) == 0 && (p.out16_lo & 7) == 0));
    if (!no_tmae && !force_wg2 && tma_ok)
      for (const Spec& sp : specs_tmae)
        if (sp.key == key) {
          spec = &sp;
          tmae = true;
        }
  }
  if (a->ln_stats_out != nullptr && !tmae)
    return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats_out is only available with the TMA epilogue (unsplit, K <= 1536, aligned F32 output)");
  if (a->ln_stats != nullptr && geglu && !tmae)
    return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats with GEGLU is only available with the TMA epilogue (K <= 1536, aligned F16 output)");
  if (a->ln_stats != nullptr && !geglu && (spec == &generic || split > 1))
    return set_error(MVD_EINVAL, "mvd_gemm_f16: ln_stats with QKV_HEADS needs aligned, unsplit operands");
  if (wide && (!tmae || spec->wide == nullptr))
    return set_error(MVD_EINVAL, "mvd_gemm_f16: tile_n > 256 is only available with the TMA epilogue (unsplit, aligned F32 / F16 output, no activation)");
  static bool configured = false;
  if (!configured) {
    for (const Spec& sp : specs) {
      MVD_CUDA_CHECK(cudaFuncSetAttribute(sp.one, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      MVD_CUDA_CHECK(cudaFuncSetAttribute(sp.two, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    for (const Spec& sp : specs_tmae) {
      MVD_CUDA_CHECK(cudaFuncSetAttribute(sp.one, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      MVD_CUDA_CHECK(cudaFuncSetAttribute(sp.two, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      if (sp.wide != nullptr) MVD_CUDA_CHECK(cudaFuncSetAttribute(sp.wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    for (const Spec& sp : specs_wg2) {
      MVD_CUDA_CHECK(cudaFuncSetAttribute(sp.one, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      MVD_CUDA_CHECK(cudaFuncSetAttribute(sp.two, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    MVD_CUDA_CHECK(cudaFuncSetAttribute(generic.one, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    MVD_CUDA_CHECK(cudaFuncSetAttribute(generic.two, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    configured = true;
  }
  const int nwg = spec->nwg;
  const int threads = 128 + 128 * nwg;

  // ---- shared memory / TMEM budget: ring stages + staging tiles (2 per warpgroup with two warpgroups, 1 with three; TMAE: the chunk
  //      slots) + GEGLU tile bias (256 floats per warpgroup) + barriers
  const int stage_bytes = A_BYTES + (pair ? bn / 2 : bn) * 128;
  CUtensorMap tmOut, tmRes, tmO16, tmO16lo;
  memset(&tmOut, 0, sizeof(tmOut));
  memset(&tmRes, 0, sizeof(tmRes));
  memset(&tmO16, 0, sizeof(tmO16));
  memset(&tmO16lo, 0, sizeof(tmO16lo));
  int stg_bytes_total = (nwg == 2 ? 4 : 3) * STG_BYTES;
  if (tmae) {
    const bool f16out = a->out_mode == MVD_OUT_F16;
    // slot: [fp32 chunk 128 x 32 (residual in, F32 result out) | fp16 chunk | fp16 lo chunk]; an F16 output without residual needs the fp16 chunk only
    p.slot_h16 = (f16out && !has_res) ? 0 : 16384;
    p.slot_lo = p.slot_h16 + 8192;
    p.slot_bytes = (f16out && !has_res) ? 8192 : 16384 + ((f16out || a->out16 != nullptr) ? 8192 : 0) + (p.out16_lo > 0 ? 8192 : 0);
    // two slots per warpgroup (a store draining / the next residual chunk arriving while this chunk is worked on) where the epilogue is
    // what the tile waits for — short K — or where they are cheap (fp16 chunks); deeper K wants the shared memory as ring stages
    // (measured: 65536x256x736 GELU at 2 ring stages 49 us, at 3 stages 41 us)
    p.spw = (p.kb_per_split <= 10 || p.slot_bytes <= 8192) ? 2 : 1;
    if (p.spw == 2 && (232448 - 1024 - (6 * p.slot_bytes + nwg * 2560 + 512)) / stage_bytes < 3) p.spw = 1;
    stg_bytes_total = 3 * p.spw * p.slot_bytes;
    int rc = f16out ? make_tmap_2d_ex(&tmOut, a->out, 2, n_out, a->M, a->ldc, 32, BM, 64)
                    : make_tmap_2d_ex(&tmOut, a->out, 4, n_out, a->M, a->ldc, 32, BM, 128);
    if (rc != MVD_OK) return rc;
    if (has_res) {
      rc = make_tmap_2d_ex(&tmRes, a->residual, 4, a->N, a->M, a->ldr, 32, BM, 128);
      if (rc != MVD_OK) return rc;
    }
    if (a->out16 != nullptr) {
      rc = make_tmap_2d_ex(&tmO16, a->out16, 2, a->N, a->M, a->ld16, 32, BM, 64);
      if (rc != MVD_OK) return rc;
      if (p.out16_lo > 0) {
        rc = make_tmap_2d_ex(&tmO16lo, static_cast<const __half*>(a->out16) + p.out16_lo, 2, a->N, a->M, a->ld16, 32, BM, 64);
        if (rc != MVD_OK) return rc;
      }
    }
  }
  const int fixed = stg_bytes_total + nwg * (tmae ? 2560 : 1024) + 512 + (a->ln_stats != nullptr ? 2048 + (tmae ? 0 : 4096) : 0);  // + rowstat[2][128] (+ lnvec[2][512])
  int stages = (232448 - 1024 - fixed) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return set_error(MVD_EINVAL, "mvd_gemm_f16: tile does not fit in shared memory");
  p.stages = stages;
  const int dyn = stages * stage_bytes + fixed + 1024;
  p.acc_stride = wide ? 0 : (bn <= 128 ? 128 : 256);
  p.tmem_cols = wide ? 512 : 2 * p.acc_stride;

  if (pair) {
    const int grid = 2 * (p.num_units < slots ? p.num_units : slots);
    MVD_CUDA_CHECK(launch_kernel(wide ? spec->wide : spec->two, dim3(grid), dim3(threads), dyn, stream, 2, tmA, tmB, tmOut, tmRes, tmO16, tmO16lo, p));
  } else {
    const int grid = p.num_units < sms ? p.num_units : sms;
    MVD_CUDA_CHECK(launch_kernel(spec->one, dim3(grid), dim3(threads), dyn, stream, 1, tmA, tmB, tmOut, tmRes, tmO16, tmO16lo, p));
  }
  count_launch();
  MVD_CUDA_CHECK(cudaGetLastError());
  return MVD_OK;
}

#ifdef MVD_GEMM_TRACE
// debug build only: copies the CTA records gathered so far to `dst` (<= max_rec records of 96 bytes) and restarts the log
extern "C" int mvd_debug_gemm_trace(void* dst, int max_rec) {
  unsigned n = 0;
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n)) != cudaSuccess) return -1;
  if (n > TRACE_CAP) n = TRACE_CAP;
  if (static_cast<int>(n) > max_rec) n = static_cast<unsigned>(max_rec);
  if (n > 0 && cudaMemcpyFromSymbol(dst, g_trace, static_cast<size_t>(n) * sizeof(TraceRec)) != cudaSuccess) return -1;
  const unsigned zero = 0;
  if (cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero)) != cudaSuccess) return -1;
  return static_cast<int>(n);
}
#endif

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
