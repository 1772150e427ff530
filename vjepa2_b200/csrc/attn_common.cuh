// Shared pieces of the attention kernels: head-dimension chunking and MMA issue helpers.
//
// A head of HD bf16 values is stored in shared memory as one or two "chunks", each chunk being a
// [rows][chunk_width] sub-tile whose row pitch equals its TMA/UMMA swizzle width:
//     HD = 32 : one 32-wide chunk  (64-byte rows,  SWIZZLE_64B)
//     HD = 64 : one 64-wide chunk  (128-byte rows, SWIZZLE_128B)
//     HD = 80 : 64-wide + 16-wide  (128-byte rows SWIZZLE_128B + 32-byte rows SWIZZLE_32B)   -- ViT-H
// Contractions over the head dimension walk the chunks k-step by k-step; products whose N extent is the head
// dimension are issued as one tcgen05.mma per chunk into adjacent TMEM columns.
#pragma once
#include "common.cuh"

namespace vj {

template <int HD>
struct HDim {
  static_assert(HD == 32 || HD == 64 || HD == 80, "head_dim 32, 64 or 80");
  static constexpr int NCH = (HD == 80) ? 2 : 1;
  __host__ __device__ static constexpr int width(int c) { return HD == 80 ? (c == 0 ? 64 : 16) : HD; }
  __host__ __device__ static constexpr int start(int c) { return (HD == 80 && c == 1) ? 64 : 0; }
  __host__ __device__ static constexpr int swb(int c) { return width(c) * 2; }
  __host__ __device__ static constexpr int off(int c, int rows) { return start(c) * 2 * rows; }  // bytes
};

struct TMapPair {
  CUtensorMap m[2];
};

__device__ __forceinline__ uint64_t swz_of(int swb) { return swb == 128 ? SWZ_128B : (swb == 64 ? SWZ_64B : SWZ_32B); }
__device__ __forceinline__ uint64_t desc_k_rt(uint32_t addr, int swb) {
  return make_smem_desc(addr, 16, swb * 8, swz_of(swb));
}
__device__ __forceinline__ uint64_t desc_mn_rt(uint32_t addr, uint32_t lbo, int swb) {
  return make_smem_desc(addr, lbo, swb * 8, swz_of(swb));
}

// D[128 x N] (+)= A[128 x HD] * B[N x HD]^T, both operands K-major chunked tiles (a_rows / b_rows rows each).
template <int HD>
__device__ __forceinline__ void mma_over_hd(uint32_t d_tmem, uint32_t a_base, int a_rows, uint32_t b_base, int b_rows,
                                            uint32_t idesc) {
  using H = HDim<HD>;
#pragma unroll
  for (int c = 0; c < H::NCH; ++c) {
    const uint64_t ad = desc_k_rt(a_base + H::off(c, a_rows), H::swb(c));
    const uint64_t bd = desc_k_rt(b_base + H::off(c, b_rows), H::swb(c));
#pragma unroll
    for (int k = 0; k < H::width(c) / 16; ++k)
      umma_bf16(d_tmem, desc_advance(ad, k * 32), desc_advance(bd, k * 32), idesc, (c | k) != 0 ? 1u : 0u);
  }
}

// D[128 x HD] (+)= A[128 x KT] * B[KT x HD]: B is an MN-major chunked tile with KT rows; the A descriptor of
// k-step k comes from adesc(k).  One MMA stream per head-dim chunk (N = chunk width).
template <int HD, bool A_MN, int KT, class ADesc>
__device__ __forceinline__ void mma_into_hd(uint32_t d_tmem, ADesc adesc, uint32_t b_base, bool accumulate) {
  using H = HDim<HD>;
#pragma unroll
  for (int c = 0; c < H::NCH; ++c) {
    const uint32_t idesc = make_idesc(128, H::width(c), A_MN, true);
    const uint64_t bd = desc_mn_rt(b_base + H::off(c, KT), H::width(c) * 2 * KT, H::swb(c));
#pragma unroll
    for (int k = 0; k < KT / 16; ++k)
      umma_bf16(d_tmem + H::start(c), adesc(k), desc_advance(bd, k * 16 * H::swb(c)), idesc,
                (accumulate || k != 0) ? 1u : 0u);
  }
}

// one TMA load per chunk of a [rows][HD] tile out of a (features, S, B) tensor
template <int HD>
__device__ __forceinline__ void tma_load_head_tile(void* dst, const TMapPair* maps, uint64_t* bar, int rows, int col0,
                                                   int row0, int b) {
  using H = HDim<HD>;
#pragma unroll
  for (int c = 0; c < H::NCH; ++c)
    tma_load_3d(reinterpret_cast<uint8_t*>(dst) + H::off(c, rows), &maps->m[c], bar, col0 + H::start(c), row0, b);
}

// TMEM <-> registers for N consecutive columns of the calling warp's lanes (N in {16, 32, 40})
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : VJ_R32(r, 0)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::VJ_W32(r, 0),
               "r"(taddr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[N]) {
  static_assert(N == 8 || N == 16 || N == 32 || N == 40, "unsupported column count");
  if constexpr (N == 8) {
    tmem_ld8(taddr, r);
  } else if constexpr (N == 16) {
    tmem_ld16(taddr, r);
  } else if constexpr (N == 32) {
    tmem_ld32(taddr, r);
  } else {
    uint32_t a[32], b[8];
    tmem_ld32(taddr, a);
    tmem_ld8(taddr + 32, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[32 + i] = b[i];
  }
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t (&r)[N]) {
  static_assert(N == 8 || N == 16 || N == 32 || N == 40, "unsupported column count");
  if constexpr (N == 8) {
    tmem_ld8(taddr, r);
  } else if constexpr (N == 16) {
    tmem_st16(taddr, r);
  } else if constexpr (N == 32) {
    tmem_st32(taddr, r);
  } else {
    uint32_t a[32], b[8];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = r[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = r[32 + i];
    tmem_st32(taddr, a);
    tmem_st8(taddr + 32, b);
  }
}

// host side: the per-chunk tensor maps of a (features, S, B) bf16 tensor for tiles of `rows` rows
template <int HD>
static inline int make_head_tmaps(TMapPair* out, const void* base, uint64_t features, uint64_t S, uint64_t B,
                                  uint32_t rows) {
  using H = HDim<HD>;
  const uint64_t dims[3] = {features, S, B};
  const uint64_t strides[2] = {features * 2, S * features * 2};
  for (int c = 0; c < 2; ++c) {
    const int cc = c < H::NCH ? c : 0;
    const uint32_t box[3] = {(uint32_t)H::width(cc), rows, 1};
    int r = make_tmap(&out->m[c], base, 0 /*bf16*/, 3, dims, strides, box, H::swb(cc));
    if (r) return r;
  }
  return 0;
}

// exp2 of a packed pair on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, max rel. error 7.5e-5 --
// 50x below the bf16 rounding P gets anyway): x = n + f, n = round(x), f in [-0.5, 0.5]; 2^x = p(f) * 2^n with the
// exponent added straight into the float bits.  x is clamped at -126 (covers the -inf of masked keys: 2^-126 ~ 0).
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& p0, float& p1) {
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const uint64_t x = f32x2_pack(x0, x1);
  const uint64_t magic = f32x2_pack(12582912.0f, 12582912.0f);          // 1.5 * 2^23: rounds to integer
  const uint64_t nmagic = f32x2_pack(-12582912.0f, -12582912.0f);
  const uint64_t fx = f32x2_add(x, magic);                               // low mantissa bits = n (two's complement)
  const uint64_t xr = f32x2_add(fx, nmagic);                             // n as float
  const uint64_t f = f32x2_fma(xr, f32x2_pack(-1.0f, -1.0f), x);         // x - n
  uint64_t p = f32x2_fma(f, f32x2_pack(0.0551716685f, 0.0551716685f), f32x2_pack(0.2426111251f, 0.2426111251f));
  p = f32x2_fma(p, f, f32x2_pack(0.6932609677f, 0.6932609677f));
  p = f32x2_fma(p, f, f32x2_pack(0.9999280572f, 0.9999280572f));
  float q0, q1, n0, n1;
  f32x2_unpack(p, q0, q1);
  f32x2_unpack(fx, n0, n1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(n0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(n1) << 23));
}


}  // namespace vj
