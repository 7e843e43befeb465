// Native parity check of mvd_gemm_f16 (through the C ABI) against a double-precision CPU loop.
// Build: make -C tests/native ; run on a B200: tests/native/gemm_check
// Exit code 0 = all cases within tolerance.  Prints one line per case.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/mvd_b200.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);       \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

static std::mt19937 rng(1234);
static float frand() { return std::uniform_real_distribution<float>(-1.f, 1.f)(rng); }

template <class T>
static T* dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T) + 16));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
static double gelu(double x) { return 0.5 * x * (1.0 + erf(x * 0.7071067811865476)); }
static double silu(double x) { return x / (1.0 + exp(-x)); }

struct Case {
  std::string name;
  int M, N, K;
  int a_mode = MVD_A_ROWMAJOR;
  int n_img = 0, H = 0, W = 0, C = 0;
  bool bias = false, rowbias = false, residual = false;
  int rows_per_group = 1;
  int act = MVD_ACT_NONE;
  int out_mode = MVD_OUT_F32;
  int split_k = 1;
  int tile_n = 0;
  int heads = 0, dhead = 0, dpad = 0, seq = 0;
  int lda_pad = 0;
  int cta_pair = 0;
};

static int g_fail = 0;

static void run(const Case& c) {
  const int M = c.M, N = c.N, K = c.K;
  const int lda = K + c.lda_pad, ldw = (K + 7) / 8 * 8;
  std::vector<__half> hA, hW(static_cast<size_t>(N) * ldw);
  std::vector<float> fA, fW(static_cast<size_t>(N) * ldw);
  if (c.a_mode == MVD_A_ROWMAJOR) {
    hA.resize(static_cast<size_t>(M) * lda);
  } else {
    hA.resize(static_cast<size_t>(c.n_img) * c.H * c.W * c.C);
  }
  fA.resize(hA.size());
  for (size_t i = 0; i < hA.size(); ++i) {
    hA[i] = __float2half(frand());
    fA[i] = __half2float(hA[i]);
  }
  for (size_t i = 0; i < hW.size(); ++i) {
    hW[i] = __float2half(frand() * 0.25f);
    fW[i] = __half2float(hW[i]);
  }
  std::vector<float> hb(N), hrb, hres;
  for (auto& v : hb) v = frand();
  const int groups = (M + c.rows_per_group - 1) / c.rows_per_group;
  if (c.rowbias) {
    hrb.resize(static_cast<size_t>(groups) * N);
    for (auto& v : hrb) v = frand();
  }
  if (c.residual) {
    hres.resize(static_cast<size_t>(M) * N);
    for (auto& v : hres) v = frand();
  }

  // ---- CPU reference (pre-activation accumulators in double)
  std::vector<double> acc(static_cast<size_t>(M) * N, 0.0);
  if (c.a_mode == MVD_A_ROWMAJOR) {
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        const float* a = &fA[static_cast<size_t>(m) * lda];
        const float* w = &fW[static_cast<size_t>(n) * ldw];
        for (int k = 0; k < K; ++k) s += static_cast<double>(a[k]) * w[k];
        acc[static_cast<size_t>(m) * N + n] = s;
      }
  } else {
    for (int im = 0; im < c.n_img; ++im)
      for (int y = 0; y < c.H; ++y)
        for (int x = 0; x < c.W; ++x) {
          const size_t m = (static_cast<size_t>(im) * c.H + y) * c.W + x;
          for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx) {
                const int yy = y + ky - 1, xx = x + kx - 1;
                if (yy < 0 || yy >= c.H || xx < 0 || xx >= c.W) continue;
                const float* a = &fA[((static_cast<size_t>(im) * c.H + yy) * c.W + xx) * c.C];
                const float* w = &fW[static_cast<size_t>(n) * ldw + (ky * 3 + kx) * c.C];
                for (int ch = 0; ch < c.C; ++ch) s += static_cast<double>(a[ch]) * w[ch];
              }
            acc[m * N + n] = s;
          }
        }
  }
  auto pre = [&](int m, int n) {
    double v = acc[static_cast<size_t>(m) * N + n];
    if (c.bias) v += hb[n];
    if (c.rowbias) v += hrb[static_cast<size_t>(m / c.rows_per_group) * N + n];
    return v;
  };

  // ---- device
  __half* dA = dev(hA);
  __half* dW = dev(hW);
  float* db = dev(hb);
  float* drb = c.rowbias ? dev(hrb) : nullptr;
  float* dres = c.residual ? dev(hres) : nullptr;

  mvd_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.a_mode = c.a_mode;
  g.A = dA; g.lda = c.a_mode == MVD_A_CONV3X3 ? 0 : lda;  // CONV3X3: lda is the pixel pitch (0 = dense)
  g.n_img = c.n_img; g.H = c.H; g.W = c.W; g.C = c.C;
  g.Wt = dW; g.ldw = ldw;
  g.bias = c.bias ? db : nullptr;
  g.rowbias = drb; g.rows_per_group = c.rows_per_group;
  g.residual = dres; g.ldr = N;
  g.act = c.act; g.out_mode = c.out_mode;
  g.split_k = c.split_k; g.tile_n = c.tile_n; g.cta_pair = c.cta_pair;
  static void* ws = nullptr;            // split-K workspace: zero-filled once, semaphores reset themselves
  const size_t ws_bytes = 32u << 20;
  if (ws == nullptr) {
    CK(cudaMalloc(&ws, ws_bytes));
    CK(cudaMemset(ws, 0, ws_bytes));
  }
  g.splitk_ws = ws; g.splitk_ws_bytes = static_cast<long long>(ws_bytes);

  double max_err = 0, max_ref = 0;
  size_t bad = 0, total = 0;
  int first_bad_m = -1, first_bad_n = -1;
  double first_got = 0, first_want = 0;
  auto cmp = [&](double got, double want, int m, int n, double tol_abs) {
    const double e = fabs(got - want);
    if (!(e <= tol_abs) ) {
      if (bad == 0) { first_bad_m = m; first_bad_n = n; first_got = got; first_want = want; }
      ++bad;
    }
    if (e > max_err || std::isnan(e)) max_err = e;
    if (fabs(want) > max_ref) max_ref = fabs(want);
    ++total;
  };
  const double tol32 = 2e-3, tol16 = 2e-2;
  int rc = 0;

  if (c.out_mode == MVD_OUT_QKV_HEADS) {
    const int BH = (M / c.seq) * c.heads;
    const size_t nqk = static_cast<size_t>(BH) * c.seq * c.dpad;
    __half *dq, *dk, *dv;
    CK(cudaMalloc(&dq, nqk * 2)); CK(cudaMalloc(&dk, nqk * 2)); CK(cudaMalloc(&dv, nqk * 2));
    CK(cudaMemset(dq, 0, nqk * 2)); CK(cudaMemset(dk, 0, nqk * 2)); CK(cudaMemset(dv, 0, nqk * 2));
    g.out = dq; g.out_k = dk; g.out_vt = dv;
    g.heads = c.heads; g.dhead = c.dhead; g.dpad = c.dpad; g.seq = c.seq;
    rc = mvd_gemm_f16(&g, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<__half> q(nqk), k(nqk), v(nqk);
    CK(cudaMemcpy(q.data(), dq, nqk * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(k.data(), dk, nqk * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v.data(), dv, nqk * 2, cudaMemcpyDeviceToHost));
    const int inner = c.heads * c.dhead;
    for (int m = 0; m < M; ++m) {
      const int img = m / c.seq, pos = m % c.seq;
      for (int n = 0; n < N; ++n) {
        const int which = n / inner, h = (n % inner) / c.dhead, j = n % c.dhead;
        const size_t bh = static_cast<size_t>(img) * c.heads + h;
        double got;
        if (which == 0) got = __half2float(q[(bh * c.seq + pos) * c.dpad + j]);
        else if (which == 1) got = __half2float(k[(bh * c.seq + pos) * c.dpad + j]);
        else got = __half2float(v[(bh * c.dpad + j) * c.seq + pos]);
        cmp(got, pre(m, n), m, n, tol16);
      }
    }
    // padding must stay zero
    for (size_t bh = 0; bh < static_cast<size_t>(BH); ++bh)
      for (int pos = 0; pos < c.seq; ++pos)
        for (int j = c.dhead; j < c.dpad; ++j)
          if (__half2float(q[(bh * c.seq + pos) * c.dpad + j]) != 0.f) ++bad;
    cudaFree(dq); cudaFree(dk); cudaFree(dv);
  } else {
    const bool geglu = c.act == MVD_ACT_GEGLU;
    const int No = geglu ? N / 2 : N;
    const size_t nout = static_cast<size_t>(M) * No;
    void* dout;
    CK(cudaMalloc(&dout, nout * 4));
    CK(cudaMemset(dout, 0xff, nout * 4));
    g.out = dout; g.ldc = No;
    rc = mvd_gemm_f16(&g, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(nout);
    if (c.out_mode == MVD_OUT_F32) {
      CK(cudaMemcpy(out.data(), dout, nout * 4, cudaMemcpyDeviceToHost));
    } else {
      std::vector<__half> oh(nout);
      CK(cudaMemcpy(oh.data(), dout, nout * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < nout; ++i) out[i] = __half2float(oh[i]);
    }
    const double tol = c.out_mode == MVD_OUT_F32 ? tol32 : tol16;
    const int bn = c.tile_n ? c.tile_n : 256;  // GEGLU default tile
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < No; ++n) {
        double want;
        if (geglu) {
          // packed layout: inside each bn-wide tile, first half = value, second half = gate
          const int tile = n / (bn / 2), w = n % (bn / 2);
          const double val = pre(m, tile * bn + w), gate = pre(m, tile * bn + bn / 2 + w);
          want = val * gelu(gate);
        } else {
          want = pre(m, n);
          if (c.act == MVD_ACT_GELU) want = gelu(want);
          if (c.act == MVD_ACT_SILU) want = silu(want);
          if (c.residual) want += hres[static_cast<size_t>(m) * N + n];
        }
        cmp(out[static_cast<size_t>(m) * No + n], want, m, n, tol);
      }
    cudaFree(dout);
  }
  const bool ok = (rc == 0) && bad == 0;
  printf("%-34s rc=%d max_err=%.3e max_ref=%.3e bad=%zu/%zu %s\n", c.name.c_str(), rc, max_err, max_ref, bad, total,
         ok ? "OK" : "FAIL");
  if (rc != 0) printf("    error: %s\n", mvd_last_error());
  if (bad) printf("    first mismatch at (m=%d, n=%d): got %.6f want %.6f\n", first_bad_m, first_bad_n, first_got, first_want);
  if (!ok) ++g_fail;
  cudaFree(dA); cudaFree(dW); cudaFree(db);
  if (drb) cudaFree(drb);
  if (dres) cudaFree(dres);
}

// ---- timing mode: `gemm_check bench` — representative shapes of one denoising step (N=8 views x 2 CFG branches)
struct BCase { const char* name; int M, N, K; int conv_img, conv_hw, conv_c; bool res; int act; int out_mode; int split; int tile_n; int pair = 0; };
static double bench(const BCase& c, int iters, bool quiet = false) {
  const int ldw = (c.K + 7) / 8 * 8;
  const size_t wbytes = static_cast<size_t>(c.N) * ldw * 2;
  int ncopy = static_cast<int>((300ull << 20) / wbytes) + 1;  // rotate weight copies so they stream from HBM like in the real step
  if (ncopy > 64) ncopy = 64;
  __half* dW; CK(cudaMalloc(&dW, wbytes * ncopy)); CK(cudaMemset(dW, 0, wbytes * ncopy));
  const size_t abytes = c.conv_img ? static_cast<size_t>(c.conv_img) * c.conv_hw * c.conv_hw * c.conv_c * 2 : static_cast<size_t>(c.M) * c.K * 2;
  __half* dA; CK(cudaMalloc(&dA, abytes)); CK(cudaMemset(dA, 0, abytes));
  const bool geglu = c.act == MVD_ACT_GEGLU;
  const int No = geglu ? c.N / 2 : c.N;
  void* dout; CK(cudaMalloc(&dout, static_cast<size_t>(c.M) * No * 4 * 3 + 1024)); 
  float* dres = nullptr; if (c.res) { CK(cudaMalloc(&dres, static_cast<size_t>(c.M) * c.N * 4)); CK(cudaMemset(dres, 0, static_cast<size_t>(c.M) * c.N * 4)); }
  float* db; CK(cudaMalloc(&db, c.N * 4)); CK(cudaMemset(db, 0, c.N * 4));
  static void* ws = nullptr; const size_t ws_bytes = 64u << 20;
  if (!ws) { CK(cudaMalloc(&ws, ws_bytes)); CK(cudaMemset(ws, 0, ws_bytes)); }
  mvd_gemm_args g; memset(&g, 0, sizeof(g));
  g.M = c.M; g.N = c.N; g.K = c.K; g.A = dA; g.lda = c.conv_img ? 0 : c.K; g.ldw = ldw; g.bias = db; g.residual = dres; g.ldr = c.N;
  g.act = c.act; g.out_mode = c.out_mode; g.out = dout; g.ldc = No; g.split_k = c.split; g.tile_n = c.tile_n; g.cta_pair = c.pair;
  g.splitk_ws = ws; g.splitk_ws_bytes = ws_bytes; g.rows_per_group = 1;
  if (c.conv_img) { g.a_mode = MVD_A_CONV3X3; g.n_img = c.conv_img; g.H = g.W = c.conv_hw; g.C = c.conv_c; }
  if (c.out_mode == MVD_OUT_QKV_HEADS) {
    g.heads = 8; g.dhead = c.N / 24; g.dpad = (g.dhead + 63) / 64 * 64; g.seq = c.conv_hw;  // conv_hw doubles as seq here
    const size_t nqk = static_cast<size_t>(c.M) * 8 * g.dpad * 2;
    CK(cudaFree(dout)); CK(cudaMalloc(&dout, nqk * 3));
    g.out = dout; g.out_k = static_cast<char*>(dout) + nqk; g.out_vt = static_cast<char*>(dout) + 2 * nqk; g.n_img = 0;
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rc = 0;
  cudaStream_t st; CK(cudaStreamCreate(&st));
  for (int i = 0; i < 3; ++i) { g.Wt = dW + (static_cast<size_t>(i % ncopy) * wbytes) / 2; rc |= mvd_gemm_f16(&g, st); }
  CK(cudaStreamSynchronize(st));
  // the launches are replayed from a CUDA graph, as in the real step: host-side launch cost is not part of the number
  cudaGraph_t graph; cudaGraphExec_t gexec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < iters; ++i) { g.Wt = dW + (static_cast<size_t>(i % ncopy) * wbytes) / 2; rc |= mvd_gemm_f16(&g, st); }
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&gexec, graph, 0));
  CK(cudaGraphLaunch(gexec, st));
  CK(cudaStreamSynchronize(st));
  cudaEventRecord(e0, st);
  CK(cudaGraphLaunch(gexec, st));
  cudaEventRecord(e1, st);
  CK(cudaStreamSynchronize(st));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(gexec); cudaGraphDestroy(graph); cudaStreamDestroy(st);
  const double us = ms * 1e3 / iters;
  if (!quiet) {
    printf("%-44s rc=%d %8.1f us  %7.1f TF/s\n", c.name, rc, us, 2.0 * c.M * c.N * c.K / (us * 1e-6) / 1e12);
    if (rc) printf("   error: %s\n", mvd_last_error());
  }
  cudaFree(dW); cudaFree(dA); cudaFree(dout); cudaFree(db); if (dres) cudaFree(dres);
  return rc ? 1e30 : us;
}
static int bench_main(int only) {
  const BCase cs[] = {
      {"lin 16384x320x320 f32 res", 16384, 320, 320, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"lin 16384x320x320 f32", 16384, 320, 320, 0, 0, 0, false, 0, MVD_OUT_F32, 0, 0},
      {"lin 4096x640x640 f32 res", 4096, 640, 640, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"lin 1024x1280x1280 f32 res", 1024, 1280, 1280, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"lin 256x1280x1280 f32 res", 256, 1280, 1280, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"lin 16384x320x1280 f32 res", 16384, 320, 1280, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"lin 4096x640x2560 f32 res", 4096, 640, 2560, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"lin 1024x1280x5120 f32 res", 1024, 1280, 5120, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"lin 256x1280x5120 f32 res", 256, 1280, 5120, 0, 0, 0, true, 0, MVD_OUT_F32, 0, 0},
      {"geglu 16384x2560x320", 16384, 2560, 320, 0, 0, 0, false, MVD_ACT_GEGLU, MVD_OUT_F16, 1, 256},
      {"geglu 4096x5120x640", 4096, 5120, 640, 0, 0, 0, false, MVD_ACT_GEGLU, MVD_OUT_F16, 1, 256},
      {"geglu 1024x10240x1280", 1024, 10240, 1280, 0, 0, 0, false, MVD_ACT_GEGLU, MVD_OUT_F16, 1, 256},
      {"geglu 256x10240x1280", 256, 10240, 1280, 0, 0, 0, false, MVD_ACT_GEGLU, MVD_OUT_F16, 1, 256},
      {"qkv 16384x960x320", 16384, 960, 320, 0, 1024, 0, false, 0, MVD_OUT_QKV_HEADS, 1, 0},
      {"qkv 4096x1920x640", 4096, 1920, 640, 0, 256, 0, false, 0, MVD_OUT_QKV_HEADS, 1, 0},
      {"qkv 1024x3840x1280", 1024, 3840, 1280, 0, 64, 0, false, 0, MVD_OUT_QKV_HEADS, 1, 0},
      {"grid 65536x512x256 gelu f16", 65536, 512, 256, 0, 0, 0, false, MVD_ACT_GELU, MVD_OUT_F16, 1, 0},
      {"grid 65536x768x256 f16", 65536, 768, 256, 0, 0, 0, false, 0, MVD_OUT_F16, 1, 0},
      {"grid 65536x256x512 f32 res", 65536, 256, 512, 0, 0, 0, true, 0, MVD_OUT_F32, 1, 0},
      {"grid 65536x256x256 f32 res", 65536, 256, 256, 0, 0, 0, true, 0, MVD_OUT_F32, 1, 0},
      {"grid 65536x256x736 gelu f32", 65536, 256, 736, 0, 0, 0, false, MVD_ACT_GELU, MVD_OUT_F32, 1, 0},
      {"conv 16x32x32x320->320 res", 16384, 320, 2880, 16, 32, 320, true, 0, MVD_OUT_F32, 0, 0},
      {"conv 16x16x16x640->640 res", 4096, 640, 5760, 16, 16, 640, true, 0, MVD_OUT_F32, 0, 0},
      {"conv 16x8x8x1280->1280 res", 1024, 1280, 11520, 16, 8, 1280, true, 0, MVD_OUT_F32, 0, 0},
      {"conv 16x4x4x1280->1280 res", 256, 1280, 11520, 16, 4, 1280, true, 0, MVD_OUT_F32, 0, 0},
      {"conv 16x4x4x2560->1280", 256, 1280, 23040, 16, 4, 2560, false, 0, MVD_OUT_F32, 0, 0},
      {"conv 16x16x16x1280->1280 (up)", 4096, 1280, 11520, 16, 16, 1280, false, 0, MVD_OUT_F32, 0, 0},
      {"conv 16x32x32x640->640 (up)", 16384, 640, 5760, 16, 32, 640, false, 0, MVD_OUT_F32, 0, 0},
      {"conv 16x32x32x960->320", 16384, 320, 8640, 16, 32, 960, false, 0, MVD_OUT_F32, 0, 0},
      {"lin 16384x320x768 f16", 16384, 320, 768, 0, 0, 0, false, 0, MVD_OUT_F16, 0, 0},
      // N = 320 outputs as one wave of 320-column pair tiles vs the heuristic's choice (the lines above)
      {"wide320 lin 16384x320x320 f32 res", 16384, 320, 320, 0, 0, 0, true, 0, MVD_OUT_F32, 1, 320, 2},
      {"wide320 lin 16384x320x1280 f32 res", 16384, 320, 1280, 0, 0, 0, true, 0, MVD_OUT_F32, 1, 320, 2},
      {"wide320 conv 16x32x32x320->320 res", 16384, 320, 2880, 16, 32, 320, true, 0, MVD_OUT_F32, 1, 320, 2},
      {"wide320 conv 16x32x32x960->320", 16384, 320, 8640, 16, 32, 960, false, 0, MVD_OUT_F32, 1, 320, 2},
      {"wide320 lin 16384x320x768 f16", 16384, 320, 768, 0, 0, 0, false, 0, MVD_OUT_F16, 1, 320, 2},
  };
  const int n = sizeof(cs) / sizeof(cs[0]);
  for (int i = 0; i < n; ++i)
    if (only < 0 || only == i) bench(cs[i], 40);
  return 0;
}

// ---- `gemm_check tune`: sweep (tile_n, split_k) for the small-M / weight-bound shapes of the step; prints the best
static int tune_main() {
  struct T { int M, N, K, img, hw, c; bool res; };
  const T ts[] = {
      {256, 1280, 1280, 0, 0, 0, true},   {256, 1280, 2560, 0, 0, 0, false},  {256, 1280, 5120, 0, 0, 0, true},
      {256, 1280, 768, 0, 0, 0, false},   {256, 1280, 11520, 16, 4, 1280, true}, {256, 1280, 23040, 16, 4, 2560, false},
      {256, 1280, 11520, 0, 0, 0, false}, {1024, 1280, 1280, 0, 0, 0, true},  {1024, 1280, 5120, 0, 0, 0, true},
      {1024, 1280, 2560, 0, 0, 0, false}, {1024, 1280, 1920, 0, 0, 0, false}, {1024, 1280, 640, 0, 0, 0, false},
      {1024, 1280, 768, 0, 0, 0, false},  {1024, 640, 5760, 0, 0, 0, false},  {1024, 1280, 11520, 16, 8, 1280, true},
      {1024, 1280, 23040, 16, 8, 2560, false}, {1024, 1280, 17280, 16, 8, 1920, false}, {1024, 1280, 5760, 16, 8, 640, false},
      {4096, 640, 640, 0, 0, 0, true},    {4096, 640, 2560, 0, 0, 0, true},   {4096, 640, 5760, 16, 16, 640, true},
      {4096, 640, 11520, 16, 16, 1280, false}, {4096, 1280, 11520, 16, 16, 1280, false}, {4096, 320, 2880, 0, 0, 0, false},
      {4096, 640, 768, 0, 0, 0, false},   {16384, 320, 320, 0, 0, 0, true},   {16384, 320, 1280, 0, 0, 0, true},
      {16384, 320, 2880, 16, 32, 320, true}, {16384, 320, 768, 0, 0, 0, false},
  };
  const int bns[] = {64, 96, 128, 160, 192, 256};
  const int sps[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16};
  for (const T& t : ts) {
    double best = 1e30, dflt = 0; int bbn = 0, bsp = 0;
    {
      BCase c{"", t.M, t.N, t.K, t.img, t.hw, t.c, t.res, 0, MVD_OUT_F32, 0, 0};
      dflt = bench(c, 20, true);
    }
    for (int bn : bns) {
      if (t.N % bn != 0 && bn != 256) continue;
      const int tiles = ((t.M + 127) / 128) * ((t.N + bn - 1) / bn);
      for (int sp : sps) {
        if (sp > 1 && tiles * sp > 148) continue;
        if (sp > (t.K + 63) / 64) continue;
        BCase c{"", t.M, t.N, t.K, t.img, t.hw, t.c, t.res, 0, MVD_OUT_F32, sp, bn};
        const double us = bench(c, 20, true);
        if (us < best) { best = us; bbn = bn; bsp = sp; }
      }
    }
    printf("TUNE %s M=%d N=%d K=%d res=%d : default %.1f us ; best %.1f us tile_n=%d split_k=%d\n", t.img ? "conv" : "lin", t.M, t.N, t.K,
           t.res ? 1 : 0, dflt, best, bbn, bsp);
    fflush(stdout);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "tune") == 0) return tune_main();
  if (argc > 1 && strcmp(argv[1], "bench") == 0) return bench_main(argc > 2 ? atoi(argv[2]) : -1);
  int only = argc > 1 ? atoi(argv[1]) : -1;
  std::vector<Case> cases;
  { Case c; c.name = "plain 128x128x64"; c.M = 128; c.N = 128; c.K = 64; cases.push_back(c); }
  { Case c; c.name = "plain 256x128x256"; c.M = 256; c.N = 128; c.K = 256; cases.push_back(c); }
  { Case c; c.name = "tails 300x200x736 bias+res"; c.M = 300; c.N = 200; c.K = 736; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "bn64 200x48x320"; c.M = 200; c.N = 48; c.K = 320; c.bias = true; cases.push_back(c); }
  { Case c; c.name = "bn256 512x512x1280 rowbias"; c.M = 512; c.N = 512; c.K = 1280; c.tile_n = 256; c.rowbias = true; c.rows_per_group = 64; cases.push_back(c); }
  { Case c; c.name = "f16 gelu 384x256x256"; c.M = 384; c.N = 256; c.K = 256; c.bias = true; c.act = MVD_ACT_GELU; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "f16 silu bn64 130x64x128"; c.M = 130; c.N = 64; c.K = 128; c.act = MVD_ACT_SILU; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "geglu bn128 256x512x320"; c.M = 256; c.N = 512; c.K = 320; c.bias = true; c.act = MVD_ACT_GEGLU; c.out_mode = MVD_OUT_F16; c.tile_n = 128; cases.push_back(c); }
  { Case c; c.name = "geglu bn256 256x1024x320"; c.M = 256; c.N = 1024; c.K = 320; c.bias = true; c.act = MVD_ACT_GEGLU; c.out_mode = MVD_OUT_F16; c.tile_n = 256; cases.push_back(c); }
  { Case c; c.name = "splitk4 128x256x2048 bias+res"; c.M = 128; c.N = 256; c.K = 2048; c.split_k = 4; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "splitk3 200x130x1000"; c.M = 200; c.N = 130; c.K = 1000; c.split_k = 3; c.lda_pad = 0; cases.push_back(c); }
  { Case c; c.name = "conv 2x32x32x64->64"; c.a_mode = MVD_A_CONV3X3; c.n_img = 2; c.H = c.W = 32; c.C = 64; c.N = 64; c.K = 9 * 64; c.M = 2 * 32 * 32; c.bias = true; cases.push_back(c); }
  { Case c; c.name = "conv 3x16x16x128->160 res"; c.a_mode = MVD_A_CONV3X3; c.n_img = 3; c.H = c.W = 16; c.C = 128; c.N = 160; c.K = 9 * 128; c.M = 3 * 256; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "conv 3x8x8x192->128 rowbias"; c.a_mode = MVD_A_CONV3X3; c.n_img = 3; c.H = c.W = 8; c.C = 192; c.N = 128; c.K = 9 * 192; c.M = 3 * 64; c.rowbias = true; c.rows_per_group = 64; cases.push_back(c); }
  { Case c; c.name = "conv 5x4x4x64->64 splitk"; c.a_mode = MVD_A_CONV3X3; c.n_img = 5; c.H = c.W = 4; c.C = 64; c.N = 64; c.K = 9 * 64; c.M = 5 * 16; c.split_k = 3; cases.push_back(c); }
  { Case c; c.name = "conv stem 2x32x32x16->320"; c.a_mode = MVD_A_CONV3X3; c.n_img = 2; c.H = c.W = 32; c.C = 16; c.N = 320; c.K = 9 * 16; c.M = 2048; c.bias = true; cases.push_back(c); }
  { Case c; c.name = "conv head 2x32x32x320->5"; c.a_mode = MVD_A_CONV3X3; c.n_img = 2; c.H = c.W = 32; c.C = 320; c.N = 5; c.K = 9 * 320; c.M = 2048; c.bias = true; cases.push_back(c); }
  { Case c; c.name = "conv 1x64x64x64->64"; c.a_mode = MVD_A_CONV3X3; c.n_img = 1; c.H = c.W = 64; c.C = 64; c.N = 64; c.K = 9 * 64; c.M = 4096; cases.push_back(c); }
  { Case c; c.name = "qkv heads 2x64 d40"; c.M = 128; c.N = 3 * 8 * 40; c.K = 320; c.out_mode = MVD_OUT_QKV_HEADS; c.heads = 8; c.dhead = 40; c.dpad = 64; c.seq = 64; cases.push_back(c); }
  { Case c; c.name = "qkv heads 3x16 d160"; c.M = 48; c.N = 3 * 8 * 160; c.K = 1280; c.out_mode = MVD_OUT_QKV_HEADS; c.heads = 8; c.dhead = 160; c.dpad = 192; c.seq = 16; cases.push_back(c); }
  { Case c; c.name = "persistent 20480x320x320 bias+res"; c.M = 20480; c.N = 320; c.K = 320; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "persistent f16 20000x960x320"; c.M = 20000; c.N = 960; c.K = 320; c.out_mode = MVD_OUT_F16; c.bias = true; cases.push_back(c); }
  { Case c; c.name = "auto-split 256x1280x11520 bias+res"; c.M = 256; c.N = 1280; c.K = 11520; c.split_k = 0; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "auto-split gelu f16 1024x1280x1280"; c.M = 1024; c.N = 1280; c.K = 1280; c.split_k = 0; c.bias = true; c.act = MVD_ACT_GELU; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "split16 rowbias 130x640x8192"; c.M = 130; c.N = 640; c.K = 8192; c.split_k = 16; c.rowbias = true; c.rows_per_group = 16; cases.push_back(c); }
  { Case c; c.name = "conv 16x8x8x1280->1280 auto-split res"; c.a_mode = MVD_A_CONV3X3; c.n_img = 16; c.H = c.W = 8; c.C = 1280; c.N = 1280; c.K = 9 * 1280; c.M = 1024; c.split_k = 0; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "geglu 4096x2560x320"; c.M = 4096; c.N = 2560; c.K = 320; c.bias = true; c.act = MVD_ACT_GEGLU; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "f16 out tails 300x200x320 bias"; c.M = 300; c.N = 200; c.K = 320; c.bias = true; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "odd N 130x37x128 bias+res"; c.M = 130; c.N = 37; c.K = 128; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "split5 silu f16 256x320x2560 rowbias"; c.M = 256; c.N = 320; c.K = 2560; c.split_k = 5; c.bias = c.rowbias = true; c.rows_per_group = 16; c.act = MVD_ACT_SILU; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "qkv heads 16x1024 d40 persistent"; c.M = 16384; c.N = 3 * 8 * 40; c.K = 320; c.out_mode = MVD_OUT_QKV_HEADS; c.heads = 8; c.dhead = 40; c.dpad = 64; c.seq = 1024; cases.push_back(c); }
  { Case c; c.name = "qkv heads d8 (generic scatter)"; c.M = 256; c.N = 3 * 8 * 8; c.K = 64; c.out_mode = MVD_OUT_QKV_HEADS; c.heads = 8; c.dhead = 8; c.dpad = 64; c.seq = 64; cases.push_back(c); }
  { Case c; c.name = "persistent res 40000x160x256"; c.M = 40000; c.N = 160; c.K = 256; c.bias = c.residual = true; cases.push_back(c); }
  // narrow tiles for the one / two m-tile layers (tools/tune_gemm.py picks them for 256x1280x1280)
  { Case c; c.name = "bn32 256x1280x1280 bias+res"; c.M = 256; c.N = 1280; c.K = 1280; c.tile_n = 32; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "bn48 200x1280x768 f16"; c.M = 200; c.N = 1280; c.K = 768; c.tile_n = 48; c.bias = true; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "bn32 split4 256x1280x2560"; c.M = 256; c.N = 1280; c.K = 2560; c.tile_n = 32; c.split_k = 3; c.bias = true; cases.push_back(c); }
  // 320-column pair tiles (two N = 160 MMAs into one single-buffered accumulator, TMA epilogue)
  { Case c; c.name = "wide320 512x320x1280 bias+res"; c.M = 512; c.N = 320; c.K = 1280; c.tile_n = 320; c.cta_pair = 2; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "wide320 odd tiles 384x320x320 f16"; c.M = 384; c.N = 320; c.K = 320; c.tile_n = 320; c.cta_pair = 2; c.bias = true; c.out_mode = MVD_OUT_F16; cases.push_back(c); }
  { Case c; c.name = "wide320 40960x320x320 res (3 units per pair)"; c.M = 40960; c.N = 320; c.K = 320; c.tile_n = 320; c.cta_pair = 2; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "wide320 conv 16x32x32x320->320 res"; c.a_mode = MVD_A_CONV3X3; c.n_img = 16; c.H = c.W = 32; c.C = 320; c.N = 320; c.K = 9 * 320; c.M = 16384; c.tile_n = 320; c.cta_pair = 2; c.bias = c.residual = true; cases.push_back(c); }
  { Case c; c.name = "wide320 N tail 300x300x640"; c.M = 300; c.N = 300; c.K = 640; c.tile_n = 320; c.cta_pair = 2; c.bias = true; cases.push_back(c); }
  for (size_t i = 0; i < cases.size(); ++i)
    if (only < 0 || only == static_cast<int>(i)) run(cases[i]);
  printf("%s (%d failing)\n", g_fail ? "GEMM CHECK FAILED" : "GEMM CHECK PASSED", g_fail);
  return g_fail ? 1 : 0;
}
