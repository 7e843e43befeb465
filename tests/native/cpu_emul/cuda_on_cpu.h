// TEST INFRASTRUCTURE — a minimal CUDA-on-CPU execution shim, never part of the product.
//
// Purpose: the build container has no GPU.  Kernel files that only use the classic SIMT subset (thread / block indices, shared memory,
// __syncthreads, warp shuffles, atomics, __ldg, vector types, fp16 conversions, thread-block clusters with distributed shared memory)
// — csrc/elementwise.cu, gridattn.cu, norm.cu, train.cu — are compiled a second time as plain C++ (-DMVD_CPU_EMULATION) and executed
// here with one fiber per CUDA thread, blocks (or clusters) one after the other, so that their indexing, reductions, rounding and closed
// forms are checked before the code reaches a B200 (tests/test_kernel_sources_cpu_shim.py, tests/test_training.py).  It says nothing
// about performance, memory coalescing, the hardware memory model or races between blocks; the -m gpu tests remain the proof.
//
// Semantics kept: the threads of a block (and the blocks of a cluster) are all live at once and interleave at barriers,
// __syncthreads() is a block-wide barrier, __shfl_xor_sync exchanges through a per-warp buffer between two warp barriers (every lane
// of the warp must take part, as on hardware with a full mask), `static` stands in for __shared__ (valid because one block is resident
// at a time; cluster kernels use dynamic shared memory only), fp16 is IEEE binary16 (_Float16, round to nearest even).
#pragma once
#include <stdint.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <random>
#include <vector>

#include "../../../include/mvd_b200.h"

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
// IEEE binary16 through the compiler's _Float16 (conversions from float round to nearest even, like __float2half_rn)
typedef _Float16 __half;
struct alignas(4) __half2 { __half x, y; };
inline __half __float2half_rn(float f) { return static_cast<__half>(f); }
inline float __half2float(__half h) { return static_cast<float>(h); }
inline __half2 __floats2half2_rn(float a, float b) { return __half2{static_cast<__half>(a), static_cast<__half>(b)}; }
inline float2 __half22float2(__half2 h) { return float2{static_cast<float>(h.x), static_cast<float>(h.y)}; }
inline unsigned short __half_as_ushort(__half h) { unsigned short u; memcpy(&u, &h, 2); return u; }
inline float __sinf(float x) { return sinf(x); }
inline float __expf(float x) { return expf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __cosf(float x) { return cosf(x); }
#define __launch_bounds__(...)
namespace cpu_emul {
// fma.rn.f32.f16: fp16 x fp16 (exact in fp32) + fp32, one rounding; operands as raw binary16 bit patterns
inline float fma_f16(float acc, uint32_t a_bits, uint32_t b_bits) {
  unsigned short ua = static_cast<unsigned short>(a_bits), ub = static_cast<unsigned short>(b_bits);
  __half a, b;
  memcpy(&a, &ua, 2);
  memcpy(&b, &ub, 2);
  return fmaf(static_cast<float>(a), static_cast<float>(b), acc);
}
}  // namespace cpu_emul

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static

namespace cpu_emul {

// Execution model: every CUDA thread of the resident block(s) is a FIBER (ucontext) on the calling host thread.  A fiber runs until it
// reaches a barrier (__syncthreads, the two barriers inside a warp shuffle, a cluster barrier, the end of its block) and yields to the
// scheduler, which resumes the next runnable fiber in thread order; the last arrival releases the barrier and keeps running.  No OS
// threads, no locks: execution is deterministic, atomics are plain read-modify-writes, and a barrier costs one context switch per
// waiting thread instead of a futex round trip.  A barrier that can never complete (a kernel bug: some threads skipped it) is detected
// as "no runnable fiber" and aborts with a message.
struct Barrier {  // reusable, generation-counted
  explicit Barrier(int n_) : n(n_) {}
  int n, count = 0;
  unsigned gen = 0;
  inline void wait();
};

struct Warp {
  explicit Warp(int lanes) : bar(lanes) {}
  Barrier bar;      // as many lanes as the warp really has (a 1-thread block is a 1-lane warp)
  uint64_t slot[32];
};

struct Ctx {
  dim3 threadIdx, blockIdx, blockDim, gridDim;
  Barrier* block_bar = nullptr;
  Warp* warp = nullptr;
  int lane = 0;
  void* dyn_smem = nullptr;  // the block's dynamic shared memory
  // thread-block cluster (launches with cluster_x > 1: the cluster's blocks are resident together, each with its own dynamic shared memory)
  int cluster_rank = 0, cluster_size = 1;
  Barrier* cluster_bar = nullptr;
  char** peer_dyn = nullptr;
};

struct Fiber {
  ucontext_t uc;
  std::unique_ptr<char[]> stack;
  Ctx ctx;
  bool done = false;
  Barrier* wait_bar = nullptr;  // the barrier this fiber sleeps on, released once its generation moves past wait_gen
  unsigned wait_gen = 0;
};

struct Sched {
  ucontext_t main_uc;
  Fiber* cur = nullptr;
  const std::function<void()>* body = nullptr;
  dim3 grid;
  int cluster = 1;
};
inline Sched& sched() {
  static thread_local Sched s;
  return s;
}
inline Ctx& ctx() { return sched().cur->ctx; }

inline void Barrier::wait() {
  const unsigned g = gen;
  if (++count == n) {  // last arrival: release everybody, keep running
    count = 0;
    ++gen;
    return;
  }
  Fiber* f = sched().cur;
  f->wait_bar = this;
  f->wait_gen = g;
  swapcontext(&f->uc, &sched().main_uc);
}

inline void fiber_main() {
  Sched& s = sched();
  Fiber* f = s.cur;
  Ctx& c = f->ctx;
  for (unsigned bz = 0; bz < s.grid.z; ++bz)
    for (unsigned by = 0; by < s.grid.y; ++by)
      for (unsigned bx = 0; bx < s.grid.x; bx += s.cluster) {
        c.blockIdx = dim3(bx + c.cluster_rank, by, bz);
        (*s.body)();
        c.cluster_bar->wait();  // the next blocks may reuse the shared arrays only after every thread has left these
      }
  f->done = true;  // returning switches to uc_link = the scheduler
}

// run `body` once per (block, thread).  blockDim x cluster_x fibers live for the whole launch; the cluster_x blocks of a cluster are
// resident together (their threads share a cluster barrier and can read each other's dynamic shared memory), clusters one after the other.
// `static` shared arrays are only valid for cluster_x == 1 (one resident block); cluster kernels here use dynamic shared memory only.
inline void launch(dim3 grid, dim3 block, const std::function<void()>& body, size_t dyn_smem_bytes = 0, int cluster_x = 1) {
  const int nt = static_cast<int>(block.x * block.y * block.z);
  const int nw = (nt + 31) / 32;
  const int cs = cluster_x < 1 ? 1 : cluster_x;
  if (grid.x % cs != 0) {
    fprintf(stderr, "cpu_emul::launch: grid.x %u is not a multiple of the cluster size %d\n", grid.x, cs);
    abort();
  }
  Sched& s = sched();
  if (s.cur != nullptr) {
    fprintf(stderr, "cpu_emul::launch: nested launch\n");
    abort();
  }
  Barrier cluster_bar(nt * cs);
  std::vector<std::unique_ptr<Barrier>> block_bars;
  std::vector<std::unique_ptr<Warp>> warps;
  std::vector<std::vector<uint4>> dyn(cs, std::vector<uint4>(dyn_smem_bytes / 16 + 1));  // 16-byte aligned
  std::vector<char*> peer(cs);
  for (int r = 0; r < cs; ++r) {
    block_bars.emplace_back(new Barrier(nt));
    for (int w = 0; w < nw; ++w) warps.emplace_back(new Warp(std::min(32, nt - 32 * w)));
    peer[r] = reinterpret_cast<char*>(dyn[r].data());
  }
  constexpr size_t kStack = 256 * 1024;  // untouched pages are never committed
  std::vector<Fiber> fibers(static_cast<size_t>(nt) * cs);
  s.body = &body;
  s.grid = grid;
  s.cluster = cs;
  for (int r = 0; r < cs; ++r)
    for (int t = 0; t < nt; ++t) {
      Fiber& f = fibers[static_cast<size_t>(r) * nt + t];
      Ctx& c = f.ctx;
      c.blockDim = block;
      c.gridDim = grid;
      c.threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      c.block_bar = block_bars[r].get();
      c.warp = warps[r * nw + t / 32].get();
      c.lane = t % 32;
      c.dyn_smem = peer[r];
      c.cluster_rank = r;
      c.cluster_size = cs;
      c.cluster_bar = &cluster_bar;
      c.peer_dyn = peer.data();
      f.stack.reset(new char[kStack]);
      getcontext(&f.uc);
      f.uc.uc_stack.ss_sp = f.stack.get();
      f.uc.uc_stack.ss_size = kStack;
      f.uc.uc_link = &s.main_uc;
      makecontext(&f.uc, reinterpret_cast<void (*)()>(fiber_main), 0);
    }
  // MVD_SHIM_ORDER = "reverse" | "shuffle:<seed>": the order in which runnable threads are resumed between barriers.  Results must not
  // depend on it — a kernel that misses a barrier (a shared-memory race) gives different answers under a different order.
  std::vector<size_t> order(fibers.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  const char* ord = getenv("MVD_SHIM_ORDER");
  const bool reverse = ord != nullptr && strcmp(ord, "reverse") == 0;
  const bool shuffle = ord != nullptr && strncmp(ord, "shuffle:", 8) == 0;
  std::mt19937 rng(shuffle ? static_cast<unsigned>(atoi(ord + 8)) : 0u);
  if (reverse) std::reverse(order.begin(), order.end());
  size_t remaining = fibers.size();
  while (remaining > 0) {
    bool progressed = false;
    if (shuffle) std::shuffle(order.begin(), order.end(), rng);
    for (size_t idx : order) {
      Fiber& f = fibers[idx];
      if (f.done) continue;
      if (f.wait_bar != nullptr) {
        if (f.wait_bar->gen == f.wait_gen) continue;  // still waiting
        f.wait_bar = nullptr;
      }
      s.cur = &f;
      swapcontext(&s.main_uc, &f.uc);
      progressed = true;
      if (f.done) --remaining;
    }
    if (!progressed) {
      fprintf(stderr, "cpu_emul::launch: dead-lock — %zu thread(s) wait on a barrier that the others never reach\n", remaining);
      abort();
    }
  }
  s.cur = nullptr;
  s.body = nullptr;
}

// barrier.cluster arrive + wait over every thread of the cluster
inline void cluster_barrier() { ctx().cluster_bar->wait(); }
// the same dynamic-shared-memory location in block `rank` of the cluster (mapa + ld.shared::cluster of the real build)
template <typename T>
inline const T* cluster_peer(const T* p, int rank) {
  Ctx& c = ctx();
  return reinterpret_cast<const T*>(c.peer_dyn[rank] + (reinterpret_cast<const char*>(p) - static_cast<const char*>(c.dyn_smem)));
}

}  // namespace cpu_emul

#define threadIdx (cpu_emul::ctx().threadIdx)
#define blockIdx (cpu_emul::ctx().blockIdx)
#define blockDim (cpu_emul::ctx().blockDim)
#define gridDim (cpu_emul::ctx().gridDim)

inline void __syncthreads() { cpu_emul::ctx().block_bar->wait(); }

template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "32- and 64-bit shuffles only");
  cpu_emul::Ctx& c = cpu_emul::ctx();
  memcpy(&c.warp->slot[c.lane], &v, sizeof(T));
  c.warp->bar.wait();
  T r;
  memcpy(&r, &c.warp->slot[c.lane ^ lane_mask], sizeof(T));
  c.warp->bar.wait();
  return r;
}

template <typename T>
inline T atomicAdd(T* p, T v) {  // fibers run one at a time: a plain read-modify-write is atomic
  const T old = *p;
  *p = old + v;
  return old;
}

template <typename T>
inline T __ldg(const T* p) { return *p; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
using std::max;
using std::min;

namespace mvd {
inline int set_error(int code, const char* fmt, ...) {
  static thread_local char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  fprintf(stderr, "mvd (cpu shim): %s\n", buf);
  return code;
}
inline void count_launch(int = 1) {}
inline void pdl_trigger() {}  // programmatic dependent launch: blocks and kernels run strictly in order here
inline void cluster_sync_all() { cpu_emul::cluster_barrier(); }
// common.h's launch_kernel(kernel, grid, block, dynamic shared bytes, stream, cluster size along x, args...)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, int cluster_x, Args&&... args) {
  cpu_emul::launch(grid, block, [&] { kernel(static_cast<KArgs>(args)...); }, smem, cluster_x);
  return cudaSuccess;
}
inline void pdl_wait() {}
// ptx.cuh's gelu_erf restated with exact division / exp2f in place of the two approximate SFU instructions (same A&S 7.1.26 polynomial)
inline float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = 1.0f / fmaf(0.3275911f, z, 1.0f);
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = exp2f(z * z * -1.4426950408889634f);
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);
}
inline float silu(float x) { return x / (1.0f + expf(-x)); }
}  // namespace mvd

#define MVD_CUDA_CHECK(expr)                                                      \
  do {                                                                            \
    if ((expr) != cudaSuccess) return ::mvd::set_error(MVD_ECUDA, "%s", #expr);   \
  } while (0)

// kernel<<<grid, block, 0, stream>>>(args...) of the real build
#define MVD_KLAUNCH(kernel, grid, block, stream, ...) \
  cpu_emul::launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
#define MVD_KLAUNCH_PDL(kernel, grid, block, stream, ...) MVD_KLAUNCH(kernel, grid, block, stream, __VA_ARGS__)
// common.h's MVD_LAUNCH(kernel, grid, block, dynamic shared bytes, stream, args...) and the dynamic shared array declaration
#define MVD_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cpu_emul::launch(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); }, smem)
#define MVD_DYNAMIC_SHARED(type, name) type* name = static_cast<type*>(cpu_emul::ctx().dyn_smem)
#define MVD_DYNAMIC_SHARED_ALIGNED16(type, name) MVD_DYNAMIC_SHARED(type, name)
