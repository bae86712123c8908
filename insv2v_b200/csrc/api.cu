// Error reporting + host-side TMA descriptor encoding for libivv_b200.so.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "../../include/ivv.h"
#include "common.cuh"
#include <stdlib.h>

namespace ivv {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
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
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  IVV_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver (no CUDA driver / GPU?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i];
      IVV_REQUIRE(strides_bytes[i] % 16 == 0, "TMA stride %llu of dim %d is not a multiple of 16 bytes",
                  (unsigned long long)strides_bytes[i], i);
    }
    IVV_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box dim %d = %u out of range", i, box[i]);
  }
  IVV_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address %p is not 16-byte aligned", base);
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IVV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return 0;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("IVV_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace ivv

extern "C" int ivv_abi_version(void) { return IVV_ABI_VERSION; }
extern "C" const char* ivv_last_error(void) { return ivv::g_err; }
