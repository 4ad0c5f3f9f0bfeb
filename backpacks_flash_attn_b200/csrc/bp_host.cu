// Host-side helpers: error reporting, device check, TMA tensor-map encoding via the runtime's
// driver-entry-point lookup (so the library has no link-time dependency on libcuda and loads on a
// machine without a GPU).
#include "bp_host.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

namespace bp {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return BP_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    } else {
      cudaGetLastError();
    }
  });
  return fn;
}

int encode_tensor_map(CUtensorMap* out, int dtype, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(BP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMapDataType dt = dtype == BP_DTYPE_F16    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : dtype == BP_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                    : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(BP_ERR_INVALID_ARGUMENT,
                "cuTensorMapEncodeTiled failed (CUresult %d): rank %d base %p dims [%llu %llu %llu %llu] "
                "stride1 %llu box [%u %u %u %u]",
                static_cast<int>(r), rank, base, (unsigned long long)gdim[0],
                (unsigned long long)(rank > 1 ? gdim[1] : 0), (unsigned long long)(rank > 2 ? gdim[2] : 0),
                (unsigned long long)(rank > 3 ? gdim[3] : 0), (unsigned long long)(rank > 1 ? gstr[0] : 0), bx[0],
                rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
  }
  return BP_OK;
}

}  // namespace bp

extern "C" {

int bp_abi_version(void) { return BP_ABI_VERSION; }

const char* bp_last_error(void) { return bp::g_err; }

int bp_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return bp::fail(BP_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return bp::fail(BP_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  }
  if (major != 10)
    return bp::fail(BP_ERR_ARCH, "libbackpack_b200 is built for sm_100a only; device %d is sm_%dx", dev, major);
  return BP_OK;
}

}  // extern "C"
