#include "host_common.h"

#include <mutex>
#include <string.h>

#include "../../include/vjepa2_b200.h"

namespace vj {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  encode_tiled_fn enc = get_encode();
  VJ_CHECK(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
  VJ_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3];
  cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) {
    VJ_CHECK((strides_bytes[i] & 15) == 0, "TMA stride %llu not a multiple of 16 bytes",
             (unsigned long long)strides_bytes[i]);
    gstr[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dtype == VJ_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VJ_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
           (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
           rank > 1 ? box[1] : 0);
  return 0;
}

}  // namespace vj

extern "C" const char* vj_last_error(void) { return vj::g_err; }
extern "C" int vj_abi_version(void) { return 3; }
extern "C" int vj_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  VJ_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  VJ_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm) *sm = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return 0;
}
