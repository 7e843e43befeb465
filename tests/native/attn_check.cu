// Native parity check of mvd_attn_self_f16 against a double-precision CPU softmax attention.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <random>
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

static std::mt19937 rng(7);
static int g_fail = 0;

static void run(int n_img, int heads, int seq, int dhead, int dpad, float amp) {
  const int BH = n_img * heads;
  const size_t nq = static_cast<size_t>(BH) * seq * dpad;
  std::vector<__half> q(nq, __float2half(0.f)), k(nq, __float2half(0.f)), vt(nq, __float2half(0.f));
  std::vector<float> qf(static_cast<size_t>(BH) * seq * dhead), kf(qf.size()), vf(qf.size());
  std::normal_distribution<float> nd(0.f, amp);
  for (int bh = 0; bh < BH; ++bh)
    for (int s = 0; s < seq; ++s)
      for (int j = 0; j < dhead; ++j) {
        const size_t i = (static_cast<size_t>(bh) * seq + s) * dhead + j;
        __half a = __float2half(nd(rng)), b = __float2half(nd(rng)), c = __float2half(nd(rng));
        qf[i] = __half2float(a); kf[i] = __half2float(b); vf[i] = __half2float(c);
        q[(static_cast<size_t>(bh) * seq + s) * dpad + j] = a;
        k[(static_cast<size_t>(bh) * seq + s) * dpad + j] = b;
        vt[(static_cast<size_t>(bh) * dpad + j) * seq + s] = c;
      }
  __half *dq, *dk, *dv, *dout;
  const int C = heads * dhead;
  const size_t nout = static_cast<size_t>(n_img) * seq * C;
  CK(cudaMalloc(&dq, nq * 2)); CK(cudaMalloc(&dk, nq * 2)); CK(cudaMalloc(&dv, nq * 2)); CK(cudaMalloc(&dout, nout * 2));
  CK(cudaMemcpy(dq, q.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, k.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, vt.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, nout * 2));
  int rc = mvd_attn_self_f16(dq, dk, dv, dout, n_img, heads, seq, dhead, dpad, C, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<__half> out(nout);
  CK(cudaMemcpy(out.data(), dout, nout * 2, cudaMemcpyDeviceToHost));

  const double scale = 1.0 / sqrt(static_cast<double>(dhead));
  double max_err = 0, max_ref = 0, se = 0, sr = 0;
  std::vector<double> sc(seq), o(dhead);
  for (int bh = 0; bh < BH; ++bh) {
    const int img = bh / heads, h = bh % heads;
    for (int i = 0; i < seq; ++i) {
      double mx = -1e300;
      for (int j = 0; j < seq; ++j) {
        double s = 0;
        for (int d = 0; d < dhead; ++d)
          s += static_cast<double>(qf[(static_cast<size_t>(bh) * seq + i) * dhead + d]) * kf[(static_cast<size_t>(bh) * seq + j) * dhead + d];
        sc[j] = s * scale;
        if (sc[j] > mx) mx = sc[j];
      }
      double l = 0;
      for (int d = 0; d < dhead; ++d) o[d] = 0;
      for (int j = 0; j < seq; ++j) {
        const double p = exp(sc[j] - mx);
        l += p;
        for (int d = 0; d < dhead; ++d) o[d] += p * vf[(static_cast<size_t>(bh) * seq + j) * dhead + d];
      }
      for (int d = 0; d < dhead; ++d) {
        const double want = o[d] / l;
        const double got = __half2float(out[(static_cast<size_t>(img) * seq + i) * C + h * dhead + d]);
        const double err = fabs(got - want);
        if (err > max_err || std::isnan(err)) max_err = err;
        if (fabs(want) > max_ref) max_ref = fabs(want);
        se += err * err; sr += want * want;
      }
    }
  }
  const double rel = sqrt(se / (sr + 1e-300));
  const bool ok = rc == 0 && rel < 5e-3 && !std::isnan(rel);
  printf("attn n_img=%d heads=%d seq=%4d d=%3d dpad=%3d amp=%.1f  rc=%d max_err=%.3e max_ref=%.3e rel_l2=%.3e %s\n", n_img, heads, seq,
         dhead, dpad, amp, rc, max_err, max_ref, rel, ok ? "OK" : "FAIL");
  if (rc) printf("   error: %s\n", mvd_last_error());
  if (!ok) ++g_fail;
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dout);
}

static void bench(int n_img, int heads, int seq, int dhead, int dpad) {
  const int BH = n_img * heads;
  const size_t nq = static_cast<size_t>(BH) * seq * dpad;
  __half *dq, *dk, *dv, *dout;
  CK(cudaMalloc(&dq, nq * 2)); CK(cudaMalloc(&dk, nq * 2)); CK(cudaMalloc(&dv, nq * 2));
  CK(cudaMalloc(&dout, static_cast<size_t>(n_img) * seq * heads * dhead * 2));
  CK(cudaMemset(dq, 0, nq * 2)); CK(cudaMemset(dk, 0, nq * 2)); CK(cudaMemset(dv, 0, nq * 2));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  int rc = mvd_attn_self_f16(dq, dk, dv, dout, n_img, heads, seq, dhead, dpad, heads * dhead, st);
  CK(cudaStreamSynchronize(st));
  const int iters = 20;
  cudaGraph_t graph; cudaGraphExec_t gexec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < iters; ++i) rc |= mvd_attn_self_f16(dq, dk, dv, dout, n_img, heads, seq, dhead, dpad, heads * dhead, st);
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&gexec, graph, 0));
  CK(cudaGraphLaunch(gexec, st)); CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st); CK(cudaGraphLaunch(gexec, st)); cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / iters;
  printf("attn bench img=%d heads=%d seq=%d d=%d: rc=%d %.1f us  %.1f TF/s (4*seq^2*d)\n", n_img, heads, seq, dhead, rc, us,
         4.0 * BH * seq * static_cast<double>(seq) * dhead / (us * 1e-6) / 1e12);
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dout);
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "bench") == 0) {
    bench(16, 8, 1024, 40, 64);
    bench(16, 8, 256, 80, 128);
    bench(16, 8, 64, 160, 192);
    bench(16, 8, 16, 160, 192);
    bench(16, 8, 4096, 40, 64);
    return 0;
  }
  run(1, 2, 1024, 40, 64, 5.0f);
  run(3, 8, 1024, 40, 64, 2.0f);
  run(1, 1, 16, 160, 192, 1.0f);
  run(2, 3, 64, 160, 192, 1.0f);
  run(1, 2, 128, 40, 64, 1.0f);
  run(2, 2, 256, 80, 128, 1.0f);
  run(1, 2, 1024, 40, 64, 1.0f);
  run(1, 2, 1024, 40, 64, 3.0f);
  run(1, 1, 48, 40, 64, 1.0f);
  run(1, 1, 320, 80, 128, 2.0f);
  run(1, 1, 256, 160, 192, 1.0f);
  printf("%s (%d failing)\n", g_fail ? "ATTN CHECK FAILED" : "ATTN CHECK PASSED", g_fail);
  return g_fail ? 1 : 0;
}
