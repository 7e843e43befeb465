#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mvd {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static const bool on = getenv("MVD_NO_PDL") == nullptr;
  return on;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, const cuuint32_t* estr_in = nullptr) {
  EncodeTiledFn fn = get_encode();
  if (fn == nullptr) return set_error(MVD_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; estr_in != nullptr && i < rank; ++i) estr[i] = estr_in[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(MVD_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu/%llu, box %u/%u)",
                     static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0],
                     box[1]);
  return MVD_OK;
}

int make_tmap_2d_ex(CUtensorMap* out, const void* base, int elem_bytes, long long cols, long long rows, long long ld,
                    int box_cols, int box_rows, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode();
  if (fn == nullptr) return set_error(MVD_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(MVD_ECUDA, "cuTensorMapEncodeTiled(2d_ex) failed (CUresult %d, dims %lld/%lld ld %lld box %d/%d)",
                     static_cast<int>(r), cols, rows, ld, box_cols, box_rows);
  return MVD_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int cols, int rows, int ld, int box_cols, int box_rows) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  return encode(out, base, 2, dims, strides, box);
}

int make_tmap_3d(CUtensorMap* out, const void* base, int d0, int d1, int d2, long long ld1, long long ld2, int b0,
                 int b1, int b2) {
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld1) * 2, static_cast<cuuint64_t>(ld2) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(b0), static_cast<cuuint32_t>(b1), static_cast<cuuint32_t>(b2)};
  return encode(out, base, 3, dims, strides, box);
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, int n, int h, int w, int c, int bc, int bw, int bh, int bn, int pitch,
                   int stride) {
  // pitch: elements between consecutive pixels (>= c: the image may be a column window of a wider buffer); stride: the box visits
  // every `stride`-th pixel in x and y (a strided convolution reads its taps straight from the full-resolution image)
  const cuuint64_t pp = static_cast<cuuint64_t>(pitch > 0 ? pitch : c);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {pp * 2, static_cast<cuuint64_t>(w) * pp * 2, static_cast<cuuint64_t>(h) * w * pp * 2};
  const cuuint32_t es = static_cast<cuuint32_t>(stride > 1 ? stride : 1);
  cuuint32_t box[4] = {static_cast<cuuint32_t>(bc), static_cast<cuuint32_t>(bw) * es, static_cast<cuuint32_t>(bh) * es,
                       static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[4] = {1, es, es, 1};
  return encode(out, base, 4, dims, strides, box, estr);
}

}  // namespace mvd

extern "C" const char* mvd_last_error(void) { return mvd::g_err; }
extern "C" int mvd_abi_version(void) { return 17; }
extern "C" long long mvd_launch_count(void) { return mvd::g_launches.load(std::memory_order_relaxed); }
