// Device-side multiblock-3D mask collator: src/masks/multiseq_multiblock3d.py:129-239 on the GPU, RNG-call-identical.
//
// The reference's contract is its random stream: per draw one generator seeded with the draw counter yields three
// float32 uniforms (temporal scale, spatial scale, aspect ratio -> block extent, :129-155), then the GLOBAL torch CPU
// generator yields, per sample and block, randint top, left, start in that order (:157-161).  Both are MT19937
// (at::mt19937): torch.rand(1) = (y & 0xFFFFFF) * 2^-24, torch.randint(0, n, (1,)) = y % n (one 32-bit output each,
// also when n == 1).  The global generator's state lives in device memory here (626 words: 624 state, left, next,
// uploaded once from torch.get_rng_state()) and is advanced by the kernel, so a sequence of calls consumes exactly the
// stream the reference's collator would.  The draws are sequential by definition (thread 0); the visibility grid,
// the ordered compaction into sorted index lists, the batch-wide minimum (:199-213) and the complement variants
// (:214-231) are done by the 1024 threads of ONE CTA.  The kept / hidden counts K_enc, K_pred go to a small device
// buffer the host reads one step ahead (they set the launch geometry of the step).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vjepa2_b200.h"
#include "host_common.h"

namespace vj {

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

constexpr int MT_N = 624, MT_M = 397;
constexpr int MASK_THREADS = 1024;
constexpr int MAX_BLOCKS = 64;

__device__ __forceinline__ unsigned mt_temper(unsigned y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__device__ __forceinline__ unsigned mt_mix(unsigned a, unsigned b, unsigned c) {
  const unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// at::mt19937::operator(): refill when `left` runs out, then one tempered word
__device__ unsigned mt_next(unsigned* st, int& left, int& next) {
  if (--left == 0) {
    for (int i = 0; i < MT_N; ++i) st[i] = mt_mix(st[i], st[(i + 1) % MT_N], st[(i + MT_M) % MT_N]);
    left = MT_N;
    next = 0;
  }
  return mt_temper(st[next++]);
}

// float32 uniform exactly as torch.rand(1).item() (a double holding a 24-bit fraction)
__device__ __forceinline__ double mt_uniform(unsigned y) { return (double)(y & 0xffffffu) * (1.0 / 16777216.0); }
// lo + u * (hi - lo) with Python's double rounding (no FMA contraction)
__device__ __forceinline__ double lerp_rn(double lo, double hi, double u) {
  return __dadd_rn(lo, __dmul_rn(u, __dadd_rn(hi, -lo)));
}

__global__ void __launch_bounds__(MASK_THREADS) mask_collate_kernel(unsigned* __restrict__ rng, const vj_mask_spec sp,
                                                                    const unsigned seed, const int B,
                                                                    long long* __restrict__ enc, long long* __restrict__ pred,
                                                                    int* __restrict__ counts,
                                                                    unsigned char* __restrict__ vis) {
  __shared__ unsigned s_mt[MT_N];
  __shared__ int s_draw[3 * MAX_BLOCKS];
  __shared__ int s_ext[3];
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = sp.frames * sp.rows * sp.cols;
  const int plane = sp.rows * sp.cols;
  for (int i = tid; i < MT_N; i += MASK_THREADS) s_mt[i] = rng[i];
  int left = 0, next = 0;
  if (tid == 0) {
    left = (int)rng[MT_N];
    next = (int)rng[MT_N + 1];
    // ---- block extent from the draw's own generator: the first three outputs of mt19937(seed) need words
    // 0..3 and 397..399 of the seeded state only (init_with_uint32 + the first three steps of the refill)
    unsigned lo[4], hi[3], s = seed;
    for (int j = 0; j < MT_M + 3; ++j) {
      if (j > 0) s = 1812433253u * (s ^ (s >> 30)) + (unsigned)j;
      if (j < 4) lo[j] = s;
      if (j >= MT_M) hi[j - MT_M] = s;
    }
    double u[3];
    for (int i = 0; i < 3; ++i) u[i] = mt_uniform(mt_temper(mt_mix(lo[i], lo[i + 1], hi[i])));
    const double ts = lerp_rn(sp.temporal_lo, sp.temporal_hi, u[0]);
    int bf = (int)__dmul_rn((double)sp.frames, ts);
    bf = bf < 1 ? 1 : bf;
    const double ss = lerp_rn(sp.spatial_lo, sp.spatial_hi, u[1]);
    const int area = (int)__dmul_rn((double)plane, ss);
    const double ar = lerp_rn(sp.aspect_lo, sp.aspect_hi, u[2]);
    int br = (int)rint(__dsqrt_rn(__dmul_rn((double)area, ar)));      // Python round(): half to even
    int bc = (int)rint(__dsqrt_rn(__ddiv_rn((double)area, ar)));
    br = br < sp.rows ? br : sp.rows;
    bc = bc < sp.cols ? bc : sp.cols;
    s_ext[0] = bf; s_ext[1] = br; s_ext[2] = bc;
  }
  __syncthreads();
  const int bf = s_ext[0], br = s_ext[1], bc = s_ext[2];
  const int per = (N + MASK_THREADS - 1) / MASK_THREADS;
  const int c0 = tid * per < N ? tid * per : N, c1 = (tid + 1) * per < N ? (tid + 1) * per : N;

  // block-wide sum of one int per thread, result in every thread (also leaves the per-warp sums in s_warp)
  auto block_sum = [&](int v) -> int {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                       // s_warp / s_total of the previous use are no longer read
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
      if (lane == 0) s_total = w;
    }
    __syncthreads();
    return s_total;
  };

  // ---- phase 1: per sample, draw block corners until the context is not empty; visibility bytes to scratch
  int min_e = N, min_p = N;
  for (int b = 0; b < B; ++b) {
    unsigned char* v = vis + (size_t)b * N;
    int total;
    do {
      if (tid == 0)
        for (int k = 0; k < sp.num_blocks; ++k) {
          s_draw[3 * k + 0] = (int)(mt_next(s_mt, left, next) % (unsigned)(sp.rows - br + 1));    // top
          s_draw[3 * k + 1] = (int)(mt_next(s_mt, left, next) % (unsigned)(sp.cols - bc + 1));    // left
          s_draw[3 * k + 2] = (int)(mt_next(s_mt, left, next) % (unsigned)(sp.frames - bf + 1));  // start
        }
      __syncthreads();
      int cnt = 0;
      for (int c = c0; c < c1; ++c) {
        const int f = c / plane, rem = c - f * plane, r = rem / sp.cols, cc = rem - r * sp.cols;
        bool keep = f < sp.context_frames;
        for (int k = 0; k < sp.num_blocks; ++k) {
          const int top = s_draw[3 * k], lft = s_draw[3 * k + 1], stt = s_draw[3 * k + 2];
          keep = keep && !(f >= stt && f < stt + bf && r >= top && r < top + br && cc >= lft && cc < lft + bc);
        }
        v[c] = keep ? 1 : 0;
        cnt += keep ? 1 : 0;
      }
      total = block_sum(cnt);
    } while (total == 0);
    min_e = total < min_e ? total : min_e;
    min_p = (N - total) < min_p ? (N - total) : min_p;
  }
  int Ke = min_e, Kp = min_p;
  if (sp.max_keep > 0 && sp.max_keep < Ke) Ke = sp.max_keep;
  const int Ke_sel = Ke, Kp_sel = Kp;          // how many sorted kept / hidden ids of each sample survive truncation
  if (sp.full_complement) Kp = N - Ke;
  else if (sp.pred_full_complement) Ke = N - Kp;

  // ---- phase 2: ordered compaction (exclusive scan of the visibility bytes), dense [B][K] int64 output
  for (int b = 0; b < B; ++b) {
    const unsigned char* v = vis + (size_t)b * N;
    int cnt = 0;
    for (int c = c0; c < c1; ++c) cnt += v[c];
    // exclusive prefix over threads: warp scan + scan of the warp sums
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    int vr = inc - cnt + (warp > 0 ? s_warp[warp - 1] : 0);     // visible cells before c0
    long long* eo = enc + (size_t)b * Ke;
    long long* po = pred + (size_t)b * Kp;
    for (int c = c0; c < c1; ++c) {
      const bool keep = v[c] != 0;
      const int hr = c - vr;                                     // hidden cells before c
      const bool e_sel = keep && vr < Ke_sel, p_sel = !keep && hr < Kp_sel;
      if (sp.full_complement) {            // predict every token the truncated context does not hold (:214-221)
        if (e_sel) eo[vr] = c; else po[c - (vr < Ke_sel ? vr : Ke_sel)] = c;
      } else if (sp.pred_full_complement) {                      // (:222-231)
        if (p_sel) po[hr] = c; else eo[c - (hr < Kp_sel ? hr : Kp_sel)] = c;
      } else {
        if (e_sel) eo[vr] = c;
        if (p_sel) po[hr] = c;
      }
      vr += keep ? 1 : 0;
    }
  }
  __syncthreads();
  for (int i = tid; i < MT_N; i += MASK_THREADS) rng[i] = s_mt[i];
  if (tid == 0) {
    rng[MT_N] = (unsigned)left;
    rng[MT_N + 1] = (unsigned)next;
    counts[0] = Ke;
    counts[1] = Kp;
  }
}

}  // namespace vj

extern "C" size_t vj_mask_collate_scratch(const vj_mask_spec* sp, int64_t B) {
  if (!sp || B <= 0) return 0;
  return (size_t)B * sp->frames * sp->rows * sp->cols;
}

extern "C" int vj_mask_collate(uint32_t* rng_state, const vj_mask_spec* sp, uint32_t seed, int64_t B, int64_t* masks_enc,
                               int64_t* masks_pred, int32_t* counts, void* scratch, void* stream) {
  using namespace vj;
  VJ_CHECK(rng_state && sp && masks_enc && masks_pred && counts && scratch, "vj_mask_collate: null pointer");
  VJ_CHECK(B > 0 && B < (1 << 20), "vj_mask_collate: bad batch %lld", (long long)B);
  VJ_CHECK(sp->frames > 0 && sp->rows > 0 && sp->cols > 0 && (int64_t)sp->frames * sp->rows * sp->cols < (1ll << 24),
           "vj_mask_collate: bad token grid %d x %d x %d", sp->frames, sp->rows, sp->cols);
  VJ_CHECK(sp->num_blocks > 0 && sp->num_blocks <= MAX_BLOCKS, "vj_mask_collate: num_blocks %d not in 1..%d",
           sp->num_blocks, MAX_BLOCKS);
  VJ_CHECK(sp->context_frames >= 1, "vj_mask_collate: context_frames must be >= 1");
  VJ_CHECK(!(sp->full_complement && sp->pred_full_complement),
           "vj_mask_collate: full_complement and pred_full_complement are exclusive");
  mask_collate_kernel<<<1, MASK_THREADS, 0, STREAM(stream)>>>(
      rng_state, *sp, seed, (int)B, reinterpret_cast<long long*>(masks_enc), reinterpret_cast<long long*>(masks_pred),
      counts, reinterpret_cast<unsigned char*>(scratch));
  VJ_LAUNCH_CHECK();
  return 0;
}
