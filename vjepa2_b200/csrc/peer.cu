// Device side of the data-parallel gradient all-reduce over peer-mapped (symmetric) memory (app/vjepa/train.py:279-281:
// DistributedDataParallel's gradient mean).  The bytes move on the COPY ENGINES over NVLink (cudaMemcpyAsync between
// peer-mapped buffers, issued by vjepa2_b200/train.py: PeerGradReducer), so the persistent one-CTA-per-SM GEMM /
// attention kernels of the backward pass keep every SM: an NCCL kernel cannot co-reside with them (they hold the whole
// register file), so each NCCL bucket stalls the statically scheduled compute kernel it meets -- measured on 2 B200s,
// overlapped NCCL cost as much as an exposed one.  Only two tiny kernels run on SMs:
//   * vj_peer_barrier: one warp; lane p publishes `epoch` into rank p's flag word and waits for rank p's word here
//     (release / acquire at system scope over NVLink).  32 registers, no shared memory: fits beside a GEMM CTA.
//   * vj_sum_into: dst += src_0 + src_1 + ... over one slice (fixed order, so every rank ends with identical bits).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vjepa2_b200.h"
#include "host_common.h"

namespace vj {

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

#ifndef VJ_PEER_WATCHDOG_CYCLES
#define VJ_PEER_WATCHDOG_CYCLES 40000000000ll   // ~20 s: a peer that never arrives traps instead of hanging the GPU
#endif

__global__ void __maxnreg__(32) peer_barrier_kernel(const vj_ptr_list flags, const int rank, const int world,
                                                          const unsigned epoch) {
  const int p = threadIdx.x;
  if (p >= world || p == rank) return;
  unsigned* theirs = reinterpret_cast<unsigned*>(flags.ptr[p]) + rank;      // my word in rank p's flag array
  const unsigned* mine = reinterpret_cast<const unsigned*>(flags.ptr[rank]) + p;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const long long t0 = clock64();
  unsigned v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (clock64() - t0 > VJ_PEER_WATCHDOG_CYCLES) {
      printf("vj_peer_barrier: rank %d waited too long for rank %d (epoch %u, seen %u)\n", rank, p, epoch, v);
      __trap();
    }
    __nanosleep(200);
  } while (true);
}

template <int NSRC>
__global__ void __launch_bounds__(256) sum_into_kernel(float* __restrict__ dst, const vj_ptr_list srcs, const long long n4,
                                                       const int n_src) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float4* d = reinterpret_cast<float4*>(dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = d[i];
    if (NSRC > 0) {
#pragma unroll
      for (int k = 0; k < NSRC; ++k) {
        const float4 b = __ldcs(reinterpret_cast<const float4*>(srcs.ptr[k]) + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
    } else {
      for (int k = 0; k < n_src; ++k) {
        const float4 b = __ldcs(reinterpret_cast<const float4*>(srcs.ptr[k]) + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
    }
    d[i] = a;
  }
}

}  // namespace vj

extern "C" int vj_peer_barrier(const vj_ptr_list* flags, int rank, int world, uint32_t epoch, void* stream) {
  using namespace vj;
  VJ_CHECK(flags != nullptr, "vj_peer_barrier: null flag list");
  VJ_CHECK(world >= 1 && world <= VJ_MAX_PEERS && rank >= 0 && rank < world, "vj_peer_barrier: bad rank %d / world %d", rank,
           world);
  for (int p = 0; p < world; ++p) VJ_CHECK(flags->ptr[p] != nullptr, "vj_peer_barrier: flag pointer of rank %d is null", p);
  if (world == 1) return 0;
  peer_barrier_kernel<<<1, 32, 0, STREAM(stream)>>>(*flags, rank, world, epoch);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_sum_into(float* dst, const vj_ptr_list* srcs, int n_src, int64_t n, void* stream) {
  using namespace vj;
  VJ_CHECK(dst && srcs, "vj_sum_into: null pointer");
  VJ_CHECK(n_src >= 0 && n_src <= VJ_MAX_PEERS, "vj_sum_into: n_src=%d not in 0..%d", n_src, VJ_MAX_PEERS);
  VJ_CHECK(n >= 0 && n % 4 == 0, "vj_sum_into: n=%lld must be a multiple of 4", (long long)n);
  VJ_CHECK((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "vj_sum_into: dst must be 16-byte aligned");
  for (int k = 0; k < n_src; ++k)
    VJ_CHECK(srcs->ptr[k] && (reinterpret_cast<uintptr_t>(srcs->ptr[k]) & 15) == 0, "vj_sum_into: source %d null or misaligned", k);
  if (n == 0 || n_src == 0) return 0;
  const long long n4 = n / 4;
  long long want = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 4;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaStream_t st = STREAM(stream);
  switch (n_src) {
    case 1: sum_into_kernel<1><<<grid, 256, 0, st>>>(dst, *srcs, n4, n_src); break;
    case 3: sum_into_kernel<3><<<grid, 256, 0, st>>>(dst, *srcs, n4, n_src); break;
    case 7: sum_into_kernel<7><<<grid, 256, 0, st>>>(dst, *srcs, n4, n_src); break;
    default: sum_into_kernel<0><<<grid, 256, 0, st>>>(dst, *srcs, n4, n_src); break;
  }
  VJ_LAUNCH_CHECK();
  return 0;
}
