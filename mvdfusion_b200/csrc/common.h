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

// Launch with programmatic dependent launch enabled (MVD_NO_PDL=1 in the environment turns the attribute off) and an
// optional thread-block cluster along x.  Works under stream capture: the edge becomes a programmatic graph dependency.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define MVD_LAUNCH(kernel, grid, block, smem, stream, ...) \
  MVD_CUDA_CHECK(::mvd::launch_kernel(kernel, dim3(grid), dim3(block), smem, stream, 1, __VA_ARGS__))

// dynamic shared memory of a kernel as a typed array (a macro so that tests/native/cpu_emul can substitute a host buffer)
#define MVD_DYNAMIC_SHARED(type, name) extern __shared__ type name[]
#define MVD_DYNAMIC_SHARED_ALIGNED16(type, name) extern __shared__ __align__(16) type name[]

// fp16 row-major matrix [rows, ld] of which [rows, cols] is addressable; box = box_cols x box_rows, 128B swizzle.
int make_tmap_2d(CUtensorMap* out, const void* base, int cols, int rows, int ld, int box_cols, int box_rows);
// fp16 3-D tensor [d2, d1, d0] with strides (ld1, ld2 elements); box (b0, b1, b2), 128B swizzle.
int make_tmap_3d(CUtensorMap* out, const void* base, int d0, int d1, int d2, long long ld1, long long ld2, int b0,
                 int b1, int b2);
// fp16 NHWC image batch [n, h, w, c] with `pitch` elements between pixels (0 = c); box = (bc channels, bw, bh, bn) PIXELS VISITED,
// every `stride`-th pixel in x and y (element strides; 0 / 1 = dense); 128B swizzle, OOB -> 0.
int make_tmap_nhwc(CUtensorMap* out, const void* base, int n, int h, int w, int c, int bc, int bw, int bh, int bn, int pitch = 0,
                   int stride = 1);

// generic 2-D map: `elem_bytes` 2 (fp16) or 4 (fp32); swizzle_bytes 0 / 64 / 128.  Used for epilogue TMA stores / residual loads.
int make_tmap_2d_ex(CUtensorMap* out, const void* base, int elem_bytes, long long cols, long long rows, long long ld,
                    int box_cols, int box_rows, int swizzle_bytes);

}  // namespace mvd
