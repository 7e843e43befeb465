// Host-side helpers shared by every translation unit of libmvd_b200.so:
// error reporting for the C ABI, launch counting, and TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvd_b200.h"

namespace mvd {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define MVD_CUDA_CHECK(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::mvd::set_error(MVD_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

// fp16 row-major matrix [rows, ld] of which [rows, cols] is addressable; box = box_cols x box_rows, 128B swizzle.
int make_tmap_2d(CUtensorMap* out, const void* base, int cols, int rows, int ld, int box_cols, int box_rows);
// fp16 3-D tensor [d2, d1, d0] with strides (ld1, ld2 elements); box (b0, b1, b2), 128B swizzle.
int make_tmap_3d(CUtensorMap* out, const void* base, int d0, int d1, int d2, long long ld1, long long ld2, int b0,
                 int b1, int b2);
// fp16 NHWC image batch [n, h, w, c]; box (bc, bw, bh, bn), 128B swizzle, OOB -> 0.
int make_tmap_nhwc(CUtensorMap* out, const void* base, int n, int h, int w, int c, int bc, int bw, int bh, int bn);

// generic 2-D map: `elem_bytes` 2 (fp16) or 4 (fp32); swizzle_bytes 0 / 128.  Used for epilogue TMA stores / residual loads.
int make_tmap_2d_ex(CUtensorMap* out, const void* base, int elem_bytes, long long cols, long long rows, long long ld,
                    int box_cols, int box_rows, int swizzle_bytes);

}  // namespace mvd
