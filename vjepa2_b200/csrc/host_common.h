// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptor
// encoding (driver entry point fetched through the runtime -- no link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace vj {

void set_error(const char* fmt, ...);
int sm_count();

#define VJ_CHECK(cond, ...)        \
  do {                             \
    if (!(cond)) {                 \
      vj::set_error(__VA_ARGS__);  \
      return -1;                   \
    }                              \
  } while (0)

#define VJ_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      vj::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

#define VJ_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      vj::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -3;                                                                           \
    }                                                                                      \
  } while (0)

// Tiled tensor map (dtype VJ_BF16 / VJ_F32), up to 3 dims.  dims[0] is the contiguous dimension;
// strides_bytes[i] is the byte stride of dims[i+1].  swizzle_bytes in {32, 64, 128} must equal box[0]*elt.
int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return make_tmap(out, base, 0 /*VJ_BF16*/, rank, dims, strides_bytes, box, swizzle_bytes);
}

}  // namespace vj
