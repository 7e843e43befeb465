// Native timing of the normalisation kernels (mvd_groupnorm_f32_f16, mvd_layernorm_f32_f16) at the shapes of one
// denoising step (N=8 views x 2 CFG branches), replayed from a CUDA graph as in the real step.  Inputs rotate over a few
// buffers that stay L2-resident (in the step the producer GEMM has just written them).
//   tests/native/norm_bench            all shapes
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/mvd_b200.h"

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

struct Case { const char* name; int kind; int n_img, hw, C; int silu; int calls; };  // kind 0 = GroupNorm, 1 = LayerNorm (rows = n_img*hw)

static double bench(const Case& c, int iters) {
  const size_t n = static_cast<size_t>(c.n_img) * c.hw * c.C;
  const int ncopy = 3;
  float* x; __half* y; float *g, *b;
  CK(cudaMalloc(&x, n * 4 * ncopy)); CK(cudaMalloc(&y, n * 2 * ncopy)); CK(cudaMalloc(&g, c.C * 4)); CK(cudaMalloc(&b, c.C * 4));
  std::vector<float> h(n * ncopy);
  for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<float>((i * 2654435761u >> 8) & 0xffff) / 65536.f - 0.5f;
  CK(cudaMemcpy(x, h.data(), n * 4 * ncopy, cudaMemcpyHostToDevice));
  std::vector<float> ones(c.C, 1.f);
  CK(cudaMemcpy(g, ones.data(), c.C * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(b, 0, c.C * 4));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  auto launch = [&](int i) {
    const float* xi = x + static_cast<size_t>(i % ncopy) * n;
    __half* yi = y + static_cast<size_t>(i % ncopy) * n;
    return c.kind == 0 ? mvd_groupnorm_f32_f16(xi, g, b, yi, nullptr, c.n_img, c.hw, c.C, 1e-5f, c.silu, st)
                       : mvd_layernorm_f32_f16(xi, g, b, yi, c.n_img * c.hw, c.C, 1e-5f, st);
  };
  int rc = 0;
  for (int i = 0; i < 3; ++i) rc |= launch(i);
  CK(cudaStreamSynchronize(st));
  cudaGraph_t graph; cudaGraphExec_t gexec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < iters; ++i) rc |= launch(i);
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&gexec, graph, 0));
  CK(cudaGraphLaunch(gexec, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0, st);
    CK(cudaGraphLaunch(gexec, st));
    cudaEventRecord(e1, st);
    CK(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double us = best * 1e3 / iters;
  printf("%-34s rc=%d %7.2f us  %7.0f GB/s  x%2d = %6.1f us/step\n", c.name, rc, us, n * 6.0 / (us * 1e-6) / 1e9, c.calls, us * c.calls);
  if (rc) printf("   error: %s\n", mvd_last_error());
  cudaGraphExecDestroy(gexec); cudaGraphDestroy(graph); cudaStreamDestroy(st);
  cudaFree(x); cudaFree(y); cudaFree(g); cudaFree(b);
  return us * c.calls;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 30;  // `norm_bench 1` under compute-sanitizer
  const Case cs[] = {
      {"gn img16 hw1024 C320 silu", 0, 16, 1024, 320, 1, 16}, {"gn img16 hw256 C640 silu", 0, 16, 256, 640, 1, 14},
      {"gn img16 hw64 C1280 silu", 0, 16, 64, 1280, 1, 14},   {"gn img16 hw16 C1280 silu", 0, 16, 16, 1280, 1, 13},
      {"gn img16 hw1024 C640 silu", 0, 16, 1024, 640, 1, 2},  {"gn img16 hw1024 C960 silu", 0, 16, 1024, 960, 1, 2},
      {"gn img16 hw256 C1280 silu", 0, 16, 256, 1280, 1, 1},  {"gn img16 hw256 C1920 silu", 0, 16, 256, 1920, 1, 1},
      {"gn img16 hw64 C2560 silu", 0, 16, 64, 2560, 1, 2},    {"gn img16 hw16 C2560 silu", 0, 16, 16, 2560, 1, 3},
      {"gn img16 hw1024 C320", 0, 16, 1024, 320, 0, 0},
      {"ln rows16384 C320", 1, 16, 1024, 320, 0, 16},         {"ln rows4096 C640", 1, 16, 256, 640, 0, 16},
      {"ln rows1024 C1280", 1, 16, 64, 1280, 0, 16},          {"ln rows256 C1280", 1, 16, 16, 1280, 0, 4},
      {"ln rows65536 C256", 1, 64, 1024, 256, 0, 0},
  };
  double total = 0;
  for (const Case& c : cs) total += bench(c, iters);
  printf("sum over one step: %.1f us\n", total);
  return 0;
}
