// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma -> TMEM
// (double-buffered accumulator) -> epilogue warps -> swizzled smem staging -> TMA store / TMA reduce-add.
//
//   warp 0   TMA producer (one elected lane)
//   warp 1   MMA issuer   (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warp 2   TMEM allocate / free
//   warp 3   idle
//   warps 4-11 epilogue, two warpgroups: warp (4+q) / (8+q) owns TMEM lanes [32q, 32q+32) == output rows
//              m0+32q..+31; the 128-byte-wide column groups of the tile alternate between the groups.
//              Each warp converts its 32x(128 B) block in registers (bias, RoPE, bf16 rounding, GELU,
//              GELU', residual -- operands prefetched one step ahead), writes it to its own 4 KB swizzled
//              staging buffer and hands it to the TMA unit: full-line coalesced stores, no LSU pressure.
//              fp32 "out += acc" (weight gradients) uses cp.reduce.async.bulk.tensor (.add) -- no read at all.
//
// Tile: 128 (M) x BN (N) x 64 (K) per stage.  Operands may be K-major or MN-major (transposed storage), which
// covers forward (x W^T), dgrad (dy W) and wgrad (dy^T x) without any transpose pass.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_common.h"
#include "../../include/vjepa2_b200.h"

namespace vj {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_STG_BYTES = 4096;          // per epilogue warp, per staging buffer: 32 rows x 128 B
constexpr int EPI_INTERNAL_REDUCE = 1 << 20;  // out += acc through TMA reduce-add

// Optional in-kernel cycle accounting (build with -DVJ_GEMM_PROFILE; used by `make build/selftest_prof`):
//  [0] MMA warp total  [1] MMA wait full (TMA late)  [2] MMA wait tempty (epilogue late)
//  [3] producer wait empty  [4] epilogue wait tfull  [5] epilogue busy  [6] CTAs  [7] epilogue tmem ld+wait
#ifdef VJ_GEMM_PROFILE
__device__ unsigned long long g_gemm_prof[12];   // [8] wait_read stall [9] math+stage [10] fence+issue
#define VJ_PROF_T0(v) const long long v = clock64()
#define VJ_PROF_ADD(acc, v) acc += clock64() - v
#else
#define VJ_PROF_T0(v)
#define VJ_PROF_ADD(acc, v)
#endif

struct GemmEpi {
  const float* bias;
  const void* residual;
  const void* aux_in;
  long long ldr, ld_aux;
  const __half* rope;
  int rope_hd, rope_D;
  int flags;
  float* bias_grad;      // VJ_EPI_BIAS_GRAD: += accumulator column n_out of every row
  int n_out;             // output columns (the kernel's N counts the ones-column block as well)
};

template <int BN, bool AUX>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_PER_WARP = (AUX ? 2 : 1) * GEMM_STG_BYTES;   // out (+ pre-activation) staging
  static constexpr int STG_TOTAL = 8 * STG_PER_WARP;
  static constexpr int STAGES_RAW = (227 * 1024 - STG_TOTAL - 2048) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;   // TMEM columns between the two accumulator stages
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int OFF_STG = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_STG + STG_TOTAL;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024 /*align slack*/;
  static_assert(BN <= 256 && BN % 64 == 0, "BN must be a multiple of 64 (128-byte bf16 store groups) and <= 256");
  static_assert(STAGES >= 3, "need at least 3 pipeline stages");
};

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& mb, int& nb) {
  // bands of 16 m-blocks; inside a band m varies fastest so co-resident CTAs share the same B tile
  constexpr int GROUP_M = 16;
  const int per_band = GROUP_M * num_n;
  const int band = tile / per_band;
  const int first_m = band * GROUP_M;
  const int gm = min(GROUP_M, num_m - first_m);
  const int r = tile - band * per_band;
  nb = r / gm;
  mb = first_m + (r - nb * gm);
}

// ---------------------------------------------------------------- TMA store helpers
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- epilogue math
// Side inputs of one 32-column block of one output row, fetched ahead of the accumulator:
//   bf16 residual -> v[0..3];  fp32 residual -> v[0..7];  bf16 dGELU operand -> v[4..7];  RoPE cos/sin -> v[0..7]
struct EpiSide {
  uint4 v[8];
};

__device__ __forceinline__ void epilogue_prefetch(const GemmEpi& e, EpiSide& s, long long row, int col0, int N) {
  const int flags = e.flags;
  if (flags & VJ_EPI_RESIDUAL) {
    if (flags & VJ_EPI_RES_F32) {
      const float* rp = reinterpret_cast<const float*>(e.residual) + row * e.ldr + col0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col0 + i * 4 < N) s.v[i] = *reinterpret_cast<const uint4*>(rp + i * 4);
    } else {
      const bf16* rp = reinterpret_cast<const bf16*>(e.residual) + row * e.ldr + col0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (col0 + i * 8 < N) s.v[i] = *reinterpret_cast<const uint4*>(rp + i * 8);
    }
  }
  if (flags & VJ_EPI_DGELU) {
    const bf16* ap = reinterpret_cast<const bf16*>(e.aux_in) + row * e.ld_aux + col0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (col0 + i * 8 < N) s.v[4 + i] = *reinterpret_cast<const uint4*>(ap + i * 8);
  }
  if (flags & VJ_EPI_ROPE) {
    // two 16-column groups (a group never straddles a head: hd % 16 == 0): cos -> v[2g..2g+1], sin -> v[4+2g..]
    const __half* tr = e.rope + row * 2 * e.rope_hd;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int c = col0 + g * 16;
      if (c < N && c < 2 * e.rope_D) {
        const int d0 = (c % e.rope_D) % e.rope_hd;
        s.v[2 * g] = *reinterpret_cast<const uint4*>(tr + d0);
        s.v[2 * g + 1] = *reinterpret_cast<const uint4*>(tr + d0 + 8);
        s.v[4 + 2 * g] = *reinterpret_cast<const uint4*>(tr + e.rope_hd + d0);
        s.v[4 + 2 * g + 1] = *reinterpret_cast<const uint4*>(tr + e.rope_hd + d0 + 8);
      }
    }
  }
}

__device__ __forceinline__ void rope_pairs8(float* v, const uint4 c, const uint4 s) {
  const __half2* ch = reinterpret_cast<const __half2*>(&c);
  const __half2* sh = reinterpret_cast<const __half2*>(&s);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 cc = __half22float2(ch[i]);
    const float2 ss = __half22float2(sh[i]);
    const float x0 = v[2 * i], x1 = v[2 * i + 1];
    v[2 * i] = x0 * cc.x - x1 * ss.x;
    v[2 * i + 1] = x1 * cc.y + x0 * ss.y;
  }
}

__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact-erf GELU without the erf: h(u) = erfc(u / sqrt2) / 2 = 2^(-q(u)) with q a degree-5 polynomial in u = |x|
// (fit of -log2(erfc) on [0, 6 sqrt2], abs. error of h <= 3e-7, of gelu <= 1.2e-6, of gelu' <= 5e-7 -- three orders
// below the bf16 rounding of the result).  Phi(x) = 1 - h for x >= 0, h for x < 0 (no cancellation in the tail), so
// gelu(x) = max(x, 0) - |x| h.  Scalar FFMAs with immediate coefficients: 5 FFMA + 1 MUFU.EX2 + FMNMX + FFMA per
// element (the packed-fp32 form spent more on building register pairs than it saved; ncu opcode mix, profiles/).
__device__ __forceinline__ float gelu_neg_q(float u) {
  float r = fmaf(u, -0.00052044867f, 0.0073974645f);
  r = fmaf(r, u, -0.052561168f);
  r = fmaf(r, u, -0.45925471f);
  r = fmaf(r, u, -1.1510913f);
  return fmaf(r, u, -1.0f);                              // -(1 + u R(u)): the "/ 2" is the leading -1
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fabsf(x);
  return fmaf(-u, mufu_ex2(gelu_neg_q(u)), fmaxf(x, 0.f));
}
// gelu'(x) = Phi(x) + x * pdf(x)
__device__ __forceinline__ float dgelu_fast(float x) {
  const float u = fabsf(x);
  const float h = mufu_ex2(gelu_neg_q(u));
  const float cdf = 0.5f + copysignf(0.5f - h, x);
  return fmaf(x * 0.39894228040143268f, mufu_ex2(x * x * -0.72134752f), cdf);
}

// 32 consecutive columns of one output row: accumulator -> final values (v) and optional pre-activation (pre)
template <bool WANT_PRE>
__device__ __forceinline__ void epilogue_math(const GemmEpi& e, const uint32_t (&acc)[32], const EpiSide& s, int col0,
                                              int N, float (&v)[32], float (&pre)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  const int flags = e.flags;
  if (flags & VJ_EPI_BIAS) {
    if (col0 + 32 <= N) {                               // full block (all but a ragged last one): no per-load checks
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (col0 + i < N) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + i));
          v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
      }
    }
  }
  if (flags & VJ_EPI_ROPE) {
    // mirror the reference: the Linear output is rounded to bf16 before the rotation (modules.py:330-365)
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int c = col0 + g * 16;
      if (c < N && c < 2 * e.rope_D) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[g * 16 + i] = bf16_round(v[g * 16 + i]);
        rope_pairs8(v + g * 16, s.v[2 * g], s.v[4 + 2 * g]);
        rope_pairs8(v + g * 16 + 8, s.v[2 * g + 1], s.v[4 + 2 * g + 1]);
      }
    }
  }
  if (WANT_PRE) {
#pragma unroll
    for (int i = 0; i < 32; ++i) pre[i] = v[i];
  }
  if (flags & VJ_EPI_ROUND_BF16) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
  }
  if (flags & VJ_EPI_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
  }
  if (flags & VJ_EPI_DGELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = s.v[4 + i];
      v[i * 8] *= dgelu_fast(bf16_lo(u.x)); v[i * 8 + 1] *= dgelu_fast(bf16_hi(u.x));
      v[i * 8 + 2] *= dgelu_fast(bf16_lo(u.y)); v[i * 8 + 3] *= dgelu_fast(bf16_hi(u.y));
      v[i * 8 + 4] *= dgelu_fast(bf16_lo(u.z)); v[i * 8 + 5] *= dgelu_fast(bf16_hi(u.z));
      v[i * 8 + 6] *= dgelu_fast(bf16_lo(u.w)); v[i * 8 + 7] *= dgelu_fast(bf16_hi(u.w));
    }
  }
  if (flags & VJ_EPI_RESIDUAL) {
    if (flags & VJ_EPI_RES_F32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = s.v[i];
        v[i * 4] += __uint_as_float(u.x); v[i * 4 + 1] += __uint_as_float(u.y);
        v[i * 4 + 2] += __uint_as_float(u.z); v[i * 4 + 3] += __uint_as_float(u.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = s.v[i];
        v[i * 8] += bf16_lo(u.x); v[i * 8 + 1] += bf16_hi(u.x); v[i * 8 + 2] += bf16_lo(u.y); v[i * 8 + 3] += bf16_hi(u.y);
        v[i * 8 + 4] += bf16_lo(u.z); v[i * 8 + 5] += bf16_hi(u.z); v[i * 8 + 6] += bf16_lo(u.w); v[i * 8 + 7] += bf16_hi(u.w);
      }
    }
  }
}

// ---------------------------------------------------------------- 16-column epilogue unit (16-epilogue-warp variant)
// Half the registers of the 32-column unit (accumulator 16 + results 16 + pre-activation 16 + side operands 16), so
// that 16 epilogue warps fit the 120 registers setmaxnreg can give them (see Gemm2Cfg).
//   bf16 residual -> v[0..1];  fp32 residual -> v[0..3];  bf16 dGELU operand -> v[2..3];  RoPE cos -> v[0..1], sin -> v[2..3]
struct EpiSide16 {
  uint4 v[4];
};

__device__ __forceinline__ void epilogue_fetch16(const GemmEpi& e, EpiSide16& s, long long row, int col0, int N) {
  const int flags = e.flags;
  if (flags & VJ_EPI_RESIDUAL) {
    if (flags & VJ_EPI_RES_F32) {
      const float* rp = reinterpret_cast<const float*>(e.residual) + row * e.ldr + col0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (col0 + i * 4 < N) s.v[i] = *reinterpret_cast<const uint4*>(rp + i * 4);
    } else {
      const bf16* rp = reinterpret_cast<const bf16*>(e.residual) + row * e.ldr + col0;
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (col0 + i * 8 < N) s.v[i] = *reinterpret_cast<const uint4*>(rp + i * 8);
    }
  }
  if (flags & VJ_EPI_DGELU) {
    const bf16* ap = reinterpret_cast<const bf16*>(e.aux_in) + row * e.ld_aux + col0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (col0 + i * 8 < N) s.v[2 + i] = *reinterpret_cast<const uint4*>(ap + i * 8);
  }
  if ((flags & VJ_EPI_ROPE) && col0 < N && col0 < 2 * e.rope_D) {
    const __half* tr = e.rope + row * 2 * e.rope_hd + (col0 % e.rope_D) % e.rope_hd;     // a 16-column unit never straddles a head
    s.v[0] = *reinterpret_cast<const uint4*>(tr);
    s.v[1] = *reinterpret_cast<const uint4*>(tr + 8);
    s.v[2] = *reinterpret_cast<const uint4*>(tr + e.rope_hd);
    s.v[3] = *reinterpret_cast<const uint4*>(tr + e.rope_hd + 8);
  }
}

template <bool WANT_PRE>
__device__ __forceinline__ void epilogue_math16(const GemmEpi& e, const uint32_t (&acc)[16], const EpiSide16& s, int col0,
                                                int N, float (&v)[16], float (&pre)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
  const int flags = e.flags;
  if (flags & VJ_EPI_BIAS) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      if (col0 + i < N) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    }
  }
  if ((flags & VJ_EPI_ROPE) && col0 < N && col0 < 2 * e.rope_D) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = bf16_round(v[i]);
    rope_pairs8(v, s.v[0], s.v[2]);
    rope_pairs8(v + 8, s.v[1], s.v[3]);
  }
  if (WANT_PRE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pre[i] = v[i];
  }
  if (flags & VJ_EPI_ROUND_BF16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = bf16_round(v[i]);
  }
  if (flags & VJ_EPI_GELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i]);
  }
  if (flags & VJ_EPI_DGELU) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 u = s.v[2 + i];
      v[i * 8] *= dgelu_fast(bf16_lo(u.x)); v[i * 8 + 1] *= dgelu_fast(bf16_hi(u.x));
      v[i * 8 + 2] *= dgelu_fast(bf16_lo(u.y)); v[i * 8 + 3] *= dgelu_fast(bf16_hi(u.y));
      v[i * 8 + 4] *= dgelu_fast(bf16_lo(u.z)); v[i * 8 + 5] *= dgelu_fast(bf16_hi(u.z));
      v[i * 8 + 6] *= dgelu_fast(bf16_lo(u.w)); v[i * 8 + 7] *= dgelu_fast(bf16_hi(u.w));
    }
  }
  if (flags & VJ_EPI_RESIDUAL) {
    if (flags & VJ_EPI_RES_F32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = s.v[i];
        v[i * 4] += __uint_as_float(u.x); v[i * 4 + 1] += __uint_as_float(u.y);
        v[i * 4 + 2] += __uint_as_float(u.z); v[i * 4 + 3] += __uint_as_float(u.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint4 u = s.v[i];
        v[i * 8] += bf16_lo(u.x); v[i * 8 + 1] += bf16_hi(u.x); v[i * 8 + 2] += bf16_lo(u.y); v[i * 8 + 3] += bf16_hi(u.y);
        v[i * 8 + 4] += bf16_lo(u.z); v[i * 8 + 5] += bf16_hi(u.z); v[i * 8 + 6] += bf16_lo(u.w); v[i * 8 + 7] += bf16_hi(u.w);
      }
    }
  }
}

// 16 columns of row `lane` into a [32 rows][128 B] swizzled staging block: bf16 -> chunks 2u, 2u+1; fp32 -> chunks 4u..4u+3
__device__ __forceinline__ void stage_bf16x16(uint8_t* stg, int lane, int unit, const float (&v)[16]) {
  uint8_t* rowp = stg + lane * 128;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[i * 8], v[i * 8 + 1]); u.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
    u.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]); u.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
    *reinterpret_cast<uint4*>(rowp + (((unit * 2 + i) ^ (lane & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void stage_f32x16(uint8_t* stg, int lane, int unit, const float (&v)[16]) {
  uint8_t* rowp = stg + lane * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(rowp + (((unit * 4 + i) ^ (lane & 7)) << 4)) =
        make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
}

// row `lane` of a [32 rows][128 B] staging block, 128-byte swizzle: 16-B chunk c of row r lives at c ^ (r & 7)
__device__ __forceinline__ void stage_bf16x32(uint8_t* stg, int lane, int half, const float (&v)[32]) {
  uint8_t* rowp = stg + lane * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[i * 8], v[i * 8 + 1]); u.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
    u.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]); u.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
    *reinterpret_cast<uint4*>(rowp + (((half * 4 + i) ^ (lane & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void stage_f32x32(uint8_t* stg, int lane, const float (&v)[32]) {
  uint8_t* rowp = stg + lane * 128;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(rowp + ((i ^ (lane & 7)) << 4)) = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
}

template <int BN, bool A_MN, bool B_MN, bool AUX>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux, GemmEpi epi, int M,
            int N, int K, int splits) {
  using Cfg = GemmCfg<BN, AUX>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
  // split-K (reduce-add outputs only): work item w = split * num_tiles + tile handles k-blocks [kb0, kb1)
  const int kb_per = (num_kb + splits - 1) / splits;
  const int num_work = num_tiles * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 8);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long prof_a = 0;
      (void)prof_a;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
        const int split = work / num_tiles, tile = work - split * num_tiles;
        int mb, nb;
        tile_coords(tile, num_m, num_n, mb, nb);
        const int m0 = mb * GEMM_BM, n0 = nb * BN;
        const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        // a narrow last N tile only fetches the B rows / column groups it needs
        const int n_live = min(BN, ((N - n0) + 15) & ~15);
        const bool half_b = !B_MN && n_live <= BN / 2;
        const int b_boxes = B_MN ? (n_live + 63) / 64 : 0;
        const uint32_t tx_bytes = Cfg::A_BYTES + (B_MN ? b_boxes * 8192 : (half_b ? Cfg::B_BYTES / 2 : Cfg::B_BYTES));
        for (int kb = kb0; kb < kb1; ++kb) {
          VJ_PROF_T0(tw);
          mbar_wait(&empty[stage], phase ^ 1);
          VJ_PROF_ADD(prof_a, tw);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full[stage], tx_bytes);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full[stage], kb * GEMM_BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, &full[stage], m0 + c * 64, kb * GEMM_BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, half_b ? &tmBh : &tmB, &full[stage], kb * GEMM_BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              if (c < b_boxes) tma_load_2d(sb + c * 8192, &tmB, &full[stage], n0 + c * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
#ifdef VJ_GEMM_PROFILE
      atomicAdd(&g_gemm_prof[3], (unsigned long long)prof_a);
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    // The last N tile may be narrower than BN: its MMAs are issued with N = the live columns rounded up to 16, so
    // e.g. N = 1408 costs 5 x 256 + 1 x 128 columns of tensor time instead of 6 x 256 (TMA still zero-fills the box).
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    long long prof_full = 0, prof_te = 0;
    (void)prof_full; (void)prof_te;
    VJ_PROF_T0(t_all);
    for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++local) {
      const int split = work / num_tiles;
      const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      int mb_, nb_;
      tile_coords(work - split * num_tiles, num_m, num_n, mb_, nb_);
      const int n_live = min(BN, ((N - nb_ * BN) + 15) & ~15);
      const uint32_t idesc = make_idesc(GEMM_BM, n_live, A_MN, B_MN);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      VJ_PROF_T0(t1);
      mbar_wait(&tempty[as], aphase ^ 1);
      VJ_PROF_ADD(prof_te, t1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
      for (int kb = kb0; kb < kb1; ++kb) {
        VJ_PROF_T0(t2);
        mbar_wait(&full[stage], phase);
        VJ_PROF_ADD(prof_full, t2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t ad = A_MN ? desc_mnmajor<128>(sa, 8192) : desc_kmajor<128>(sa);
          const uint64_t bd = B_MN ? desc_mnmajor<128>(sb, 8192) : desc_kmajor<128>(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adk = desc_advance(ad, A_MN ? k * 2048 : k * 32);
            const uint64_t bdk = desc_advance(bd, B_MN ? k * 2048 : k * 32);
            umma_bf16(d_tmem, adk, bdk, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == kb1 - 1) umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
#ifdef VJ_GEMM_PROFILE
    if (lane == 0) {
      atomicAdd(&g_gemm_prof[0], (unsigned long long)(clock64() - t_all));
      atomicAdd(&g_gemm_prof[1], (unsigned long long)prof_full);
      atomicAdd(&g_gemm_prof[2], (unsigned long long)prof_te);
      atomicAdd(&g_gemm_prof[6], 1ull);
    }
#endif
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue
    const int q = warp & 3;
    const int ew = warp - 4;                         // 0..7
    const int eg = ew >> 2;                          // warpgroup 0: even column groups, 1: odd
    const bool out_f32 = (epi.flags & VJ_EPI_OUT_F32) != 0;
    constexpr bool want_aux = AUX;
    const bool reduce = (epi.flags & EPI_INTERNAL_REDUCE) != 0;
    const int gw = out_f32 ? 32 : 64;                // columns per 128-byte store group
    const int ngroups = BN / gw;
    uint8_t* stg_out = smem + Cfg::OFF_STG + ew * Cfg::STG_PER_WARP;
    uint8_t* stg_aux = stg_out + GEMM_STG_BYTES;
    int local = 0;
    long long prof_w = 0, prof_b = 0, prof_ld = 0, prof_wr = 0, prof_ms = 0, prof_fi = 0;
    (void)prof_w; (void)prof_b; (void)prof_ld; (void)prof_wr; (void)prof_ms; (void)prof_fi;
    for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++local) {
      const int tile = work % num_tiles;
      int mb, nb;
      tile_coords(tile, num_m, num_n, mb, nb);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int row0 = mb * GEMM_BM + q * 32;
      const long long row = (long long)row0 + lane;
      const bool row_ok = row < M;
      const int n0 = nb * BN;
      const int nlim = min(N, n0 + BN);
      EpiSide side;
#pragma unroll
      for (int i = 0; i < 8; ++i) side.v[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && n0 + eg * gw < nlim) epilogue_prefetch(epi, side, row, n0 + eg * gw, nlim);
      VJ_PROF_T0(t3);
      mbar_wait(&tfull[as], aphase);
      VJ_PROF_ADD(prof_w, t3);
      VJ_PROF_T0(t4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;
#pragma unroll 1
      for (int g = eg; g < ngroups; g += 2) {
        const int gcol = n0 + g * gw;
        if (gcol >= N) break;                       // warp-uniform
        // the previous TMA store of this warp must have finished READING the staging buffers
        VJ_PROF_T0(t6);
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        VJ_PROF_ADD(prof_wr, t6);
        const int halves = out_f32 ? 1 : 2;
#pragma unroll 1
        for (int hh = 0; hh < halves; ++hh) {
          const int col0 = gcol + hh * 32;
          if (col0 >= nlim) break;                  // warp-uniform
          uint32_t acc[32];
          VJ_PROF_T0(t5);
          tmem_ld32(taddr + g * gw + hh * 32, acc);
          tmem_ld_wait();
          VJ_PROF_ADD(prof_ld, t5);
          VJ_PROF_T0(t7);
          const EpiSide cur = side;
          // prefetch the side inputs of the next 32-column block this warp will process
          {
            int ncol = col0 + 32;
            if (hh + 1 >= halves) ncol = gcol + 2 * gw;
            if (row_ok && ncol < nlim) epilogue_prefetch(epi, side, row, ncol, nlim);
          }
          float v[32], pre[32];
          if (want_aux) {
            epilogue_math<true>(epi, acc, cur, col0, nlim, v, pre);
            stage_bf16x32(stg_aux, lane, hh, pre);
          } else {
            epilogue_math<false>(epi, acc, cur, col0, nlim, v, pre);
          }
          if (out_f32) stage_f32x32(stg_out, lane, v);
          else stage_bf16x32(stg_out, lane, hh, v);
          VJ_PROF_ADD(prof_ms, t7);
        }
        VJ_PROF_T0(t8);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (reduce) tma_reduce_add_2d(&tmOut, stg_out, gcol, row0);
          else tma_store_2d(&tmOut, stg_out, gcol, row0);
          if (want_aux) tma_store_2d(&tmAux, stg_aux, gcol, row0);
          bulk_commit();
        }
        VJ_PROF_ADD(prof_fi, t8);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      VJ_PROF_ADD(prof_b, t4);
    }
    if (lane == 0) bulk_wait_all0();                // smem must outlive the last bulk stores
#ifdef VJ_GEMM_PROFILE
    if (warp == 4 && lane == 0) {
      atomicAdd(&g_gemm_prof[4], (unsigned long long)prof_w);
      atomicAdd(&g_gemm_prof[5], (unsigned long long)prof_b);
      atomicAdd(&g_gemm_prof[7], (unsigned long long)prof_ld);
      atomicAdd(&g_gemm_prof[8], (unsigned long long)prof_wr);
      atomicAdd(&g_gemm_prof[9], (unsigned long long)prof_ms);
      atomicAdd(&g_gemm_prof[10], (unsigned long long)prof_fi);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN, bool AUX>
static int launch_gemm(const vj_gemm_args* g, int flags, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, AUX>;
  CUtensorMap tmA, tmB, tmBh, tmOut, tmAux;
  {
    const uint64_t dimsK[2] = {(uint64_t)g->K, (uint64_t)g->M};
    const uint64_t dimsM[2] = {(uint64_t)g->M, (uint64_t)g->K};
    const uint64_t str[1] = {(uint64_t)g->lda * 2};
    const uint32_t boxK[2] = {64, GEMM_BM};
    const uint32_t boxM[2] = {64, GEMM_BK};
    int r = make_tmap(&tmA, g->a, VJ_BF16, 2, A_MN ? dimsM : dimsK, str, A_MN ? boxM : boxK, 128);
    if (r) return r;
  }
  {
    const uint64_t dimsK[2] = {(uint64_t)g->K, (uint64_t)g->N};
    const uint64_t dimsN[2] = {(uint64_t)g->N, (uint64_t)g->K};
    const uint64_t str[1] = {(uint64_t)g->ldb * 2};
    const uint32_t boxK[2] = {64, (uint32_t)BN};
    const uint32_t boxN[2] = {64, GEMM_BK};
    int r = make_tmap(&tmB, g->b, VJ_BF16, 2, B_MN ? dimsN : dimsK, str, B_MN ? boxN : boxK, 128);
    if (r) return r;
    tmBh = tmB;
    if (!B_MN) {   // half-height box for a narrow last N tile
      const uint32_t boxH[2] = {64, (uint32_t)BN / 2};
      r = make_tmap(&tmBh, g->b, VJ_BF16, 2, dimsK, str, boxH, 128);
      if (r) return r;
    }
  }
  {
    const bool f32 = (flags & VJ_EPI_OUT_F32) != 0;
    const uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)g->M};
    const uint64_t str[1] = {(uint64_t)g->ldo * (f32 ? 4 : 2)};
    const uint32_t box[2] = {f32 ? 32u : 64u, 32u};
    int r = make_tmap(&tmOut, g->out, f32 ? VJ_F32 : VJ_BF16, 2, dims, str, box, 128);
    if (r) return r;
    tmAux = tmOut;
    if (flags & VJ_EPI_AUX_OUT) {
      const uint64_t stra[1] = {(uint64_t)g->ld_aux * 2};
      const uint32_t boxa[2] = {64u, 32u};
      r = make_tmap(&tmAux, g->aux_out, VJ_BF16, 2, dims, stra, boxa, 128);
      if (r) return r;
    }
  }
  GemmEpi e;
  e.bias = g->bias; e.residual = g->residual; e.aux_in = g->aux_in;
  e.ldr = g->ldr; e.ld_aux = g->ld_aux; e.flags = flags;
  e.rope = reinterpret_cast<const __half*>(g->rope_table); e.rope_hd = g->rope_hd; e.rope_D = g->rope_D;
  e.bias_grad = nullptr; e.n_out = (int)g->N;

  auto kern = gemm_kernel<BN, A_MN, B_MN, AUX>;
  static bool attr_set = false;
  if (!attr_set) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int num_m = (int)((g->M + GEMM_BM - 1) / GEMM_BM);
  const int num_n = (int)((g->N + BN - 1) / BN);
  const int tiles = num_m * num_n;
  const int num_kb = (int)((g->K + GEMM_BK - 1) / GEMM_BK);
  // split-K for reduce-add outputs (weight gradients: few output tiles, very long K): pick the split count that
  // fills whole waves of the persistent grid; an epilogue pass costs about as much as 6 k-blocks of mainloop
  int splits = 1;
  if ((flags & EPI_INTERNAL_REDUCE) && !(flags & VJ_EPI_BIAS)) {
    const int sms = sm_count();
    double best = 0.0;
    for (int sp = 1; sp <= 32; ++sp) {
      const int kb_per = (num_kb + sp - 1) / sp;
      if (sp > 1 && kb_per < 8) break;
      const int sp_eff = (num_kb + kb_per - 1) / kb_per;
      const long long items = (long long)tiles * sp_eff;
      const long long waves = (items + sms - 1) / sms;
      const double score = (double)items / (double)(waves * sms) * kb_per / (kb_per + 6.0);
      if (score > best * 1.02) { best = score; splits = sp_eff; }
    }
  }
  const long long work = (long long)tiles * splits;
  const int grid = work < sm_count() ? (int)work : sm_count();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmBh, tmOut, tmAux, e, (int)g->M, (int)g->N,
                                                       (int)g->K, splits);
  VJ_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================================
// CTA-pair variant (cta_group::2): two CTAs of a cluster (the two SMs of a TPC) compute one 256 x BN tile.  CTA r owns
// output rows m0 + 128 r .. +127: it loads ITS 128 rows of A and HALF of the B tile (n_live / 2 rows); the leader CTA
// (rank 0) issues tcgen05.mma.cta_group::2 with M = 256, which reads A and the two B halves from both CTAs' shared
// memory and writes each CTA's 128 x BN accumulator half into that CTA's TMEM.  Per SM and k-block the tensor core
// does the same work as in the 1-CTA kernel while shared memory holds / feeds 32 KB instead of 48 KB of operands, so
// the ring is deeper and operand traffic (smem reads, L2 -> smem) per FLOP drops by a third.
//   full[s]   (leader)      1 arrival (leader's expect_tx) + the TMA bytes of BOTH CTAs
//   empty[s]  (each CTA)    tcgen05.commit multicast to both CTAs
//   tfull[a]  (each CTA)    tcgen05.commit multicast to both CTAs
//   tempty[a] (leader)      16 arrivals: the 8 epilogue warps of each CTA (the peer's arrive remotely)
// =====================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// default (.release.cta) semantics: the arrive only has to follow this warp's tcgen05.ld + wait::ld in program order
// (the data is already in registers); a .release.cluster arrive compiles to MEMBAR.ALL + ERRBAR and cost 11 % of the
// epilogue warps' stall samples (profiles/r01c_ncu_gemm2_gelu_summary.txt)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier in the leader CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same smem offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// EW = epilogue warps per CTA: 8 (default) or 16.  The 16-warp variant is for shapes whose epilogue, not the MMA
// mainloop, sets the pace (short K: the predictor's K = 384 GEMMs): four warps per scheduler instead of two hide the
// tcgen05.ld / LDG / MUFU latencies of the per-row epilogue chain (ncu of the 8-warp epilogue: issue slots 41 % busy).
// 640 threads start at 96 registers (61 440 for the CTA); setmaxnreg hands registers from the producer / MMA warpgroup (-> 32, frees 8 192) to the four
// epilogue warpgroups (-> 112, takes exactly those 8 192), which therefore work on 16-column units.  Costs staging space: EW x 4 KB (x 2 with AUX).
template <bool AUX, int EW = 8>
struct Gemm2Cfg {
  static constexpr int THREADS = (4 + EW) * 32;
  static constexpr int BN = 256;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;          // this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;         // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;          // 32 KB
  static constexpr int STG_PER_WARP = (AUX ? 2 : 1) * GEMM_STG_BYTES;
  static constexpr int STG_TOTAL = EW * STG_PER_WARP;
  static constexpr int STAGES_RAW = (227 * 1024 - STG_TOTAL - 2048) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_STRIDE = 256;
  static constexpr int TMEM_COLS = 512;
  static constexpr int OFF_STG = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_STG + STG_TOTAL;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

template <bool A_MN, bool B_MN, bool AUX, int EW>
__global__ void __launch_bounds__((4 + EW) * 32, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux, GemmEpi epi, int M,
             int N, int K, int splits) {
  using Cfg = Gemm2Cfg<AUX, EW>;
  constexpr int BN = Cfg::BN;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);      // 256-row blocks
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per = (num_kb + splits - 1) / splits;
  const int num_work = num_tiles * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * EW);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer's barriers exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (both CTAs)
    if constexpr (EW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");     // register hand-over (see Gemm2Cfg)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = pair; work < num_work; work += num_pairs) {
        const int split = work / num_tiles, tile = work - split * num_tiles;
        int mb, nb;
        tile_coords(tile, num_m, num_n, mb, nb);
        const int m0 = mb * 2 * GEMM_BM + (int)rank * GEMM_BM, n0 = nb * BN;
        const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        const int n_live = min(BN, ((N - n0) + 15) & ~15);
        const int n_half = n_live >> 1;                 // B rows / columns each CTA supplies
        const int nb0 = n0 + (int)rank * n_half;
        const int b_boxes = B_MN ? (n_half + 63) / 64 : 0;
        const uint32_t my_bytes = Cfg::A_BYTES + (B_MN ? b_boxes * 8192 : Cfg::B_BYTES);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (leader) mbar_expect_tx(&full[stage], 2 * my_bytes);
          const uint32_t fb = mapa_shared(smem_u32(&full[stage]), 0);
          if (!A_MN) {
            tma_load_2d_pair(sa, &tmA, fb, kb * GEMM_BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c) tma_load_2d_pair(sa + c * 8192, &tmA, fb, m0 + c * 64, kb * GEMM_BK);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, &tmB, fb, kb * GEMM_BK, nb0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 2 / 64; ++c)
              if (c < b_boxes) tma_load_2d_pair(sb + c * 8192, &tmB, fb, nb0 + c * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only)
    if constexpr (EW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int work = pair; work < num_work; work += num_pairs, ++local) {
        const int split = work / num_tiles;
        const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        int mb_, nb_;
        tile_coords(work - split * num_tiles, num_m, num_n, mb_, nb_);
        const int n_live = min(BN, ((N - nb_ * BN) + 15) & ~15);
        const uint32_t idesc = make_idesc(2 * GEMM_BM, n_live, A_MN, B_MN);
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t sb = sa + Cfg::A_BYTES;
            const uint64_t ad = A_MN ? desc_mnmajor<128>(sa, 8192) : desc_kmajor<128>(sa);
            const uint64_t bd = B_MN ? desc_mnmajor<128>(sb, 8192) : desc_kmajor<128>(sb);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              const uint64_t adk = desc_advance(ad, A_MN ? k * 2048 : k * 32);
              const uint64_t bdk = desc_advance(bd, B_MN ? k * 2048 : k * 32);
              umma_bf16_pair(d_tmem, adk, bdk, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            }
            umma_commit_pair(&empty[stage]);
            if (kb == kb1 - 1) umma_commit_pair(&tfull[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    if constexpr (EW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");     // warps 2, 3: same warpgroup as 0, 1
  } else {
    // ------------------------------------------------ epilogue (both CTAs, each on its own 128 rows)
    if constexpr (EW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int q = warp & 3;
    const int ew = warp - 4;
    const int eg = ew >> 2;
    const bool out_f32 = (epi.flags & VJ_EPI_OUT_F32) != 0;
    constexpr bool want_aux = AUX;
    const bool reduce = (epi.flags & EPI_INTERNAL_REDUCE) != 0;
    const int gw = out_f32 ? 32 : 64;
    const int ngroups = BN / gw;
    uint8_t* stg_out = smem + Cfg::OFF_STG + ew * Cfg::STG_PER_WARP;
    uint8_t* stg_aux = stg_out + GEMM_STG_BYTES;
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty[0]), 0);
    const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty[1]), 0);
    if constexpr (EW == 16) {
      // ---- 16 epilogue warps, 16-column units: warpgroup eg owns the 128-byte column groups g = eg, eg + 4, ...
      int local16 = 0;
      for (int work = pair; work < num_work; work += num_pairs, ++local16) {
        const int tile = work % num_tiles;
        int mb, nb;
        tile_coords(tile, num_m, num_n, mb, nb);
        const int as = local16 & 1;
        const uint32_t aphase = (local16 >> 1) & 1;
        const int row0 = mb * 2 * GEMM_BM + (int)rank * GEMM_BM + q * 32;
        const long long row = (long long)row0 + lane;
        const bool row_ok = row < M;
        const int n0 = nb * BN;
        const int nlim = min(N, n0 + BN);
        mbar_wait(&tfull[as], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;
        if (row0 < M) {
#pragma unroll 1
          for (int g = eg; g < ngroups; g += 4) {
            const int gcol = n0 + g * gw;
            if (gcol >= N) break;
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            const int units = gw / 16;
#pragma unroll 1
            for (int u = 0; u < units; ++u) {
              const int col0 = gcol + u * 16;
              if (col0 >= nlim) break;
              EpiSide16 side;
#pragma unroll
              for (int i = 0; i < 4; ++i) side.v[i] = make_uint4(0u, 0u, 0u, 0u);
              if (row_ok) epilogue_fetch16(epi, side, row, col0, nlim);
              uint32_t acc[16];
              tmem_ld16(taddr + g * gw + u * 16, acc);
              tmem_ld_wait();
              float v[16], pre[16];
              if (want_aux) {
                epilogue_math16<true>(epi, acc, side, col0, nlim, v, pre);
                stage_bf16x16(stg_aux, lane, u, pre);
              } else {
                epilogue_math16<false>(epi, acc, side, col0, nlim, v, pre);
              }
              if ((epi.flags & VJ_EPI_BIAS_GRAD) && row_ok && epi.n_out >= col0 && epi.n_out < col0 + 16)
                atomicAdd(epi.bias_grad + row, epi.n_out == col0 ? v[0] : v[8]);
              if (out_f32) stage_f32x16(stg_out, lane, u, v);
              else stage_bf16x16(stg_out, lane, u, v);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (gcol >= epi.n_out) {}
              else if (reduce) tma_reduce_add_2d(&tmOut, stg_out, gcol, row0);
              else tma_store_2d(&tmOut, stg_out, gcol, row0);
              if (want_aux) tma_store_2d(&tmAux, stg_aux, gcol, row0);
              bulk_commit();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(as ? tempty_leader1 : tempty_leader0);
      }
    } else {
    int local = 0;
    for (int work = pair; work < num_work; work += num_pairs, ++local) {
      const int tile = work % num_tiles;
      int mb, nb;
      tile_coords(tile, num_m, num_n, mb, nb);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const int row0 = mb * 2 * GEMM_BM + (int)rank * GEMM_BM + q * 32;
      const long long row = (long long)row0 + lane;
      const bool row_ok = row < M;
      const int n0 = nb * BN;
      const int nlim = min(N, n0 + BN);
      EpiSide side;
#pragma unroll
      for (int i = 0; i < 8; ++i) side.v[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && n0 + eg * gw < nlim) epilogue_prefetch(epi, side, row, n0 + eg * gw, nlim);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;
      if (row0 < M) {                                   // warp-uniform: a fully out-of-range 32-row slab stores nothing
#pragma unroll 1
        for (int g = eg; g < ngroups; g += 2) {
          const int gcol = n0 + g * gw;
          if (gcol >= N) break;
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          const int halves = out_f32 ? 1 : 2;
#pragma unroll 1
          for (int hh = 0; hh < halves; ++hh) {
            const int col0 = gcol + hh * 32;
            if (col0 >= nlim) break;
            uint32_t acc[32];
            tmem_ld32(taddr + g * gw + hh * 32, acc);
            tmem_ld_wait();
            const EpiSide cur = side;
            {
              int ncol = col0 + 32;
              if (hh + 1 >= halves) ncol = gcol + 2 * gw;
              if (row_ok && ncol < nlim) epilogue_prefetch(epi, side, row, ncol, nlim);
            }
            float v[32], pre[32];
            if (want_aux) {
              epilogue_math<true>(epi, acc, cur, col0, nlim, v, pre);
              stage_bf16x32(stg_aux, lane, hh, pre);
            } else {
              epilogue_math<false>(epi, acc, cur, col0, nlim, v, pre);
            }
            if ((epi.flags & VJ_EPI_BIAS_GRAD) && row_ok && epi.n_out >= col0 && epi.n_out < col0 + 32) {
              // the ones-column of B: accumulator column n_out of this row is sum_k A[k][row] (n_out % 8 == 0)
              const int o = epi.n_out - col0;
              const float bg = o == 0 ? v[0] : o == 8 ? v[8] : o == 16 ? v[16] : v[24];
              atomicAdd(epi.bias_grad + row, bg);
            }
            if (out_f32) stage_f32x32(stg_out, lane, v);
            else stage_bf16x32(stg_out, lane, hh, v);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (gcol >= epi.n_out) {}                      // the ones-block has no output columns
            else if (reduce) tma_reduce_add_2d(&tmOut, stg_out, gcol, row0);
            else tma_store_2d(&tmOut, stg_out, gcol, row0);
            if (want_aux) tma_store_2d(&tmAux, stg_aux, gcol, row0);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(as ? tempty_leader1 : tempty_leader0);
    }
    }
    if (lane == 0) bulk_wait_all0();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // both CTAs are done with both TMEM halves and all barriers
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
}

static int max_pairs_cached(const void* kern, int smem_bytes, int threads) {
  // co-resident CTA pairs for this kernel (74 on a B200: one pair per TPC)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (unsigned)sm_count());
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = sm_count() / 2;
  }
  return n;
}

template <bool A_MN, bool B_MN, bool AUX, int EW>
static int launch_gemm2(const vj_gemm_args* g, int flags, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<AUX, EW>;
  constexpr int BN = Cfg::BN;
  CUtensorMap tmA, tmB, tmOut, tmAux;
  // VJ_EPI_BIAS_GRAD: the MMAs run over N + 8 columns of B (the last 8 are the ones-block); the output keeps N
  const int64_t Nmma = g->N + ((flags & VJ_EPI_BIAS_GRAD) ? 8 : 0);
  {
    const uint64_t dimsK[2] = {(uint64_t)g->K, (uint64_t)g->M};
    const uint64_t dimsM[2] = {(uint64_t)g->M, (uint64_t)g->K};
    const uint64_t str[1] = {(uint64_t)g->lda * 2};
    const uint32_t boxK[2] = {64, GEMM_BM};
    const uint32_t boxM[2] = {64, GEMM_BK};
    int r = make_tmap(&tmA, g->a, VJ_BF16, 2, A_MN ? dimsM : dimsK, str, A_MN ? boxM : boxK, 128);
    if (r) return r;
  }
  {
    const uint64_t dimsK[2] = {(uint64_t)g->K, (uint64_t)Nmma};
    const uint64_t dimsN[2] = {(uint64_t)Nmma, (uint64_t)g->K};
    const uint64_t str[1] = {(uint64_t)g->ldb * 2};
    const uint32_t boxK[2] = {64, (uint32_t)BN / 2};
    const uint32_t boxN[2] = {64, GEMM_BK};
    int r = make_tmap(&tmB, g->b, VJ_BF16, 2, B_MN ? dimsN : dimsK, str, B_MN ? boxN : boxK, 128);
    if (r) return r;
  }
  {
    const bool f32 = (flags & VJ_EPI_OUT_F32) != 0;
    const uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)g->M};
    const uint64_t str[1] = {(uint64_t)g->ldo * (f32 ? 4 : 2)};
    const uint32_t box[2] = {f32 ? 32u : 64u, 32u};
    int r = make_tmap(&tmOut, g->out, f32 ? VJ_F32 : VJ_BF16, 2, dims, str, box, 128);
    if (r) return r;
    tmAux = tmOut;
    if (flags & VJ_EPI_AUX_OUT) {
      const uint64_t stra[1] = {(uint64_t)g->ld_aux * 2};
      const uint32_t boxa[2] = {64u, 32u};
      r = make_tmap(&tmAux, g->aux_out, VJ_BF16, 2, dims, stra, boxa, 128);
      if (r) return r;
    }
  }
  GemmEpi e;
  e.bias = g->bias; e.residual = g->residual; e.aux_in = g->aux_in;
  e.ldr = g->ldr; e.ld_aux = g->ld_aux; e.flags = flags;
  e.rope = reinterpret_cast<const __half*>(g->rope_table); e.rope_hd = g->rope_hd; e.rope_D = g->rope_D;
  e.bias_grad = g->bias_grad; e.n_out = (int)g->N;

  auto kern = gemm2_kernel<A_MN, B_MN, AUX, EW>;
  static int pairs = 0;
  if (pairs == 0) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    pairs = max_pairs_cached(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, Cfg::THREADS);
  }
  const int num_m = (int)((g->M + 2 * GEMM_BM - 1) / (2 * GEMM_BM));
  const int num_n = (int)((Nmma + BN - 1) / BN);
  const int tiles = num_m * num_n;
  const int num_kb = (int)((g->K + GEMM_BK - 1) / GEMM_BK);
  int splits = 1;
  if ((flags & EPI_INTERNAL_REDUCE) && !(flags & VJ_EPI_BIAS)) {
    double best = 0.0;
    for (int sp = 1; sp <= 32; ++sp) {
      const int kb_per = (num_kb + sp - 1) / sp;
      if (sp > 1 && kb_per < 8) break;
      const int sp_eff = (num_kb + kb_per - 1) / kb_per;
      const long long items = (long long)tiles * sp_eff;
      const long long waves = (items + pairs - 1) / pairs;
      const double score = (double)items / (double)(waves * pairs) * kb_per / (kb_per + 6.0);
      if (score > best * 1.02) { best = score; splits = sp_eff; }
    }
  }
  const long long work = (long long)tiles * splits;
  const int npairs = work < pairs ? (int)work : pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (unsigned)npairs);
  cfg.blockDim = dim3((unsigned)Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  VJ_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmAux, e, (int)g->M, (int)Nmma, (int)g->K, splits));
  return 0;
}

// CTA-pair kernel: on by default for problems with enough 256-row blocks; VJ_GEMM_2CTA=0 forces the 1-CTA kernels
static int g_pair_mode = -1;
static int pair_mode() {
  if (g_pair_mode < 0) {
    const char* s = getenv("VJ_GEMM_2CTA");
    g_pair_mode = (s && s[0] >= '0' && s[0] <= '2') ? s[0] - '0' : 1;      // 0 off, 1 auto, 2 every shape (tests)
  }
  return g_pair_mode;
}
// 16 epilogue warps (Gemm2Cfg): VJ_GEMM_EPI16 = 0 never, 1 every K-major-A shape (tests), unset = short-K shapes
static int g_epi16_mode = -1;
static int epi16_mode() {
  if (g_epi16_mode < 0) {
    const char* s = getenv("VJ_GEMM_EPI16");
    g_epi16_mode = (s && (s[0] == '0' || s[0] == '1')) ? s[0] - '0' : 2;
  }
  return g_epi16_mode;
}
static bool use_wide_epilogue(const vj_gemm_args* g, int flags) {
  const int mode = epi16_mode();
  if (mode != 2) return mode == 1;
  // measured on one B200 (profiles/r02y_epi16_ab.txt): K = 384 -- qkv + RoPE 0.119 -> 0.097 ms, fc1 + GELU + aux 0.153 ->
  // 0.118, fc2-dgrad * GELU' 0.153 -> 0.141; K = 1408 -- qkv + RoPE 0.451 -> 0.429, proj + residual 0.167 -> 0.160,
  // GELU' 0.230 -> 0.225, but plain / bias-only shapes lose 2-3 % (one pipeline stage less) and the AUX variant 9 % (three
  // stages): short K always, mid K only with a side-operand epilogue and without AUX
  if (g->K <= 512) return true;
  const bool side = (flags & (VJ_EPI_ROPE | VJ_EPI_DGELU)) != 0 || ((flags & VJ_EPI_RESIDUAL) && !(flags & EPI_INTERNAL_REDUCE));
  return side && !(flags & VJ_EPI_AUX_OUT) && g->K <= 2048;
}
static bool use_pair_kernel(long long M, long long N) {
  const int mode = pair_mode();
  return mode == 2 || (mode == 1 && M >= 1024 && N >= 128);
}

// Pick the N tile.  Measured on B200 (profiles/r01_*): per-FLOP speed of the mainloop is ~1.0 at BN=256,
// ~0.86 at 192 and ~0.7 at 128 (smaller tiles re-read A more often and give the single MMA-issuing thread
// less time per k-block), so a wide tile wins unless it leaves many dead columns or a ragged last wave.
static int pick_bn(long long N, long long M) {
  const int cands[] = {256, 192, 128};
  const double speed[] = {1.0, 0.86, 0.70};
  int best = 128;
  double best_cost = 1e30;
  const long long num_m = (M + 127) / 128;
  const long long sms = sm_count();
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long tiles_n = (N + bn - 1) / bn;
    const long long tiles = tiles_n * num_m;
    const long long waves = (tiles + sms - 1) / sms;              // wave quantisation of the persistent grid
    const double cost = (double)waves * bn / speed[i];
    if (cost < best_cost * 0.999) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace vj

extern "C" int vj_gemm_set_epi16_mode(int mode) {
  const int old = vj::epi16_mode();
  if (mode >= 0 && mode <= 2) vj::g_epi16_mode = mode;
  return old;
}

extern "C" int vj_gemm_set_pair_mode(int mode) {
  const int old = vj::pair_mode();
  if (mode >= 0 && mode <= 2) vj::g_pair_mode = mode;
  return old;
}

extern "C" int vj_gemm(const vj_gemm_args* g, void* stream_) {
  using namespace vj;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VJ_CHECK(g != nullptr, "vj_gemm: null args");
  VJ_CHECK(g->M > 0 && g->N > 0 && g->K > 0, "vj_gemm: empty problem M=%lld N=%lld K=%lld", (long long)g->M,
           (long long)g->N, (long long)g->K);
  VJ_CHECK(g->M < (1ll << 31) && g->N < (1ll << 31) && g->K < (1ll << 31), "vj_gemm: dimension too large");
  VJ_CHECK(g->N % 8 == 0, "vj_gemm: N=%lld must be a multiple of 8", (long long)g->N);
  VJ_CHECK(g->lda % 8 == 0 && g->ldb % 8 == 0, "vj_gemm: lda/ldb must be multiples of 8 elements (TMA 16-B pitch)");
  VJ_CHECK(g->ldo % 8 == 0, "vj_gemm: ldo must be a multiple of 8");
  VJ_CHECK(g->a && g->b && g->out, "vj_gemm: null operand");
  int flags = g->flags;
  if (flags & VJ_EPI_BIAS) VJ_CHECK(g->bias != nullptr, "vj_gemm: bias flag without pointer");
  if (flags & VJ_EPI_RESIDUAL) VJ_CHECK(g->residual != nullptr && g->ldr % 8 == 0, "vj_gemm: bad residual");
  if (flags & VJ_EPI_AUX_OUT) VJ_CHECK(g->aux_out != nullptr && g->ld_aux % 8 == 0, "vj_gemm: bad aux_out");
  if (flags & VJ_EPI_DGELU) VJ_CHECK(g->aux_in != nullptr && g->ld_aux % 8 == 0, "vj_gemm: bad aux_in");
  if (flags & VJ_EPI_AUX_OUT) VJ_CHECK(!(flags & VJ_EPI_OUT_F32), "vj_gemm: AUX_OUT needs a bf16 output");
  VJ_CHECK(!((flags & VJ_EPI_DGELU) && (flags & VJ_EPI_RES_F32)), "vj_gemm: DGELU with an fp32 residual is not supported");
  if (flags & VJ_EPI_ROPE) {
    VJ_CHECK(!(flags & (VJ_EPI_RESIDUAL | VJ_EPI_DGELU | VJ_EPI_GELU)), "vj_gemm: ROPE combines with BIAS only");
    VJ_CHECK(g->rope_table && (g->rope_hd == 32 || g->rope_hd == 64 || g->rope_hd == 80) && g->rope_D > 0 && g->rope_D % g->rope_hd == 0 &&
                 g->N == 3 * (int64_t)g->rope_D && g->rope_D % 16 == 0,
             "vj_gemm: bad ROPE arguments (hd=%d D=%d N=%lld)", g->rope_hd, g->rope_D, (long long)g->N);
  }
  if (flags & VJ_EPI_BIAS_GRAD) {
    VJ_CHECK(g->bias_grad != nullptr, "vj_gemm: BIAS_GRAD without bias_grad pointer");
    VJ_CHECK(g->a_mn_major && g->b_mn_major && (flags & VJ_EPI_OUT_F32) &&
                 !(flags & (VJ_EPI_BIAS | VJ_EPI_GELU | VJ_EPI_DGELU | VJ_EPI_ROUND_BF16 | VJ_EPI_AUX_OUT | VJ_EPI_ROPE)),
             "vj_gemm: BIAS_GRAD is a weight-gradient option (A and B MN-major, fp32 output, no other epilogue)");
    VJ_CHECK(g->ldb >= g->N + 8, "vj_gemm: BIAS_GRAD needs 8 pad columns of ones in B (ldb=%lld, N=%lld)", (long long)g->ldb,
             (long long)g->N);
    VJ_CHECK(use_pair_kernel(g->M, g->N), "vj_gemm: BIAS_GRAD is only implemented in the CTA-pair kernel (M >= 1024)");
  }
  // fp32 "out += acc": residual aliases out with the same pitch -> TMA reduce-add, no read in the epilogue
  if ((flags & VJ_EPI_RESIDUAL) && g->residual == g->out) {
    VJ_CHECK((flags & VJ_EPI_RES_F32) && (flags & VJ_EPI_OUT_F32) && g->ldr == g->ldo &&
                 !(flags & (VJ_EPI_GELU | VJ_EPI_DGELU | VJ_EPI_ROUND_BF16 | VJ_EPI_AUX_OUT)),
             "vj_gemm: in-place residual is only supported as a plain fp32 accumulate");
    flags = (flags & ~(VJ_EPI_RESIDUAL | VJ_EPI_RES_F32)) | EPI_INTERNAL_REDUCE;
  }
  const bool amn = g->a_mn_major != 0, bmn = g->b_mn_major != 0;
  VJ_CHECK(!(amn && !bmn), "vj_gemm: (A MN-major, B K-major) is not instantiated");
  if (use_pair_kernel(g->M, g->N)) {
    const bool wide = use_wide_epilogue(g, flags);
    if (flags & VJ_EPI_AUX_OUT) {
      VJ_CHECK(!amn && !bmn, "vj_gemm: AUX_OUT is only instantiated for K-major operands");
      return wide ? launch_gemm2<false, false, true, 16>(g, flags, stream) : launch_gemm2<false, false, true, 8>(g, flags, stream);
    }
    if (!amn && !bmn)
      return wide ? launch_gemm2<false, false, false, 16>(g, flags, stream) : launch_gemm2<false, false, false, 8>(g, flags, stream);
    if (!amn && bmn)
      return wide ? launch_gemm2<false, true, false, 16>(g, flags, stream) : launch_gemm2<false, true, false, 8>(g, flags, stream);
    return launch_gemm2<true, true, false, 8>(g, flags, stream);
  }
  const int bn = pick_bn(g->N, g->M);
  if (flags & VJ_EPI_AUX_OUT) {
    VJ_CHECK(!amn && !bmn, "vj_gemm: AUX_OUT is only instantiated for K-major operands");
    if (bn == 256) return launch_gemm<256, false, false, true>(g, flags, stream);
    if (bn == 192) return launch_gemm<192, false, false, true>(g, flags, stream);
    return launch_gemm<128, false, false, true>(g, flags, stream);
  }
#define VJ_GEMM_CASE(BN_, A_, B_) \
  if (bn == BN_ && amn == A_ && bmn == B_) return launch_gemm<BN_, A_, B_, false>(g, flags, stream);
  VJ_GEMM_CASE(256, false, false)
  VJ_GEMM_CASE(192, false, false)
  VJ_GEMM_CASE(128, false, false)
  VJ_GEMM_CASE(256, false, true)
  VJ_GEMM_CASE(192, false, true)
  VJ_GEMM_CASE(128, false, true)
  VJ_GEMM_CASE(256, true, true)
  VJ_GEMM_CASE(192, true, true)
  VJ_GEMM_CASE(128, true, true)
#undef VJ_GEMM_CASE
  set_error("vj_gemm: no kernel for BN=%d a_mn=%d b_mn=%d", bn, (int)amn, (int)bmn);
  return -1;
}

#ifdef VJ_GEMM_PROFILE
extern "C" int vj_gemm_prof_read(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, vj::g_gemm_prof, 12 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[12] = {0};
    cudaMemcpyToSymbol(vj::g_gemm_prof, z, sizeof(z));
  }
  return 0;
}
#endif
