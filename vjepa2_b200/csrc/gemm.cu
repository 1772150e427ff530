// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma -> TMEM
// (double-buffered accumulator) -> epilogue warps (tcgen05.ld, fused bias / GELU / dGELU / residual).
//
//   warp 0   TMA producer (one elected lane)
//   warp 1   MMA issuer   (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warp 2   TMEM allocate / free
//   warp 3   idle
//   warps 4-11 epilogue, two warpgroups: warp (4+q) / (8+q) owns TMEM lanes [32q, 32q+32) == output rows
//              m0+32q..+31; group 0 takes the even 32-column chunks of the tile, group 1 the odd ones.
//              Residual / dGELU operands of the next chunk are prefetched while the current one is computed.
//
// Tile: 128 (M) x BN (N) x 64 (K) per stage.  Operands may be K-major or MN-major (transposed
// storage), which covers forward (x W^T), dgrad (dy W) and wgrad (dy^T x) without any transpose pass.
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_common.h"
#include "../../include/vjepa2_b200.h"

namespace vj {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;

struct GemmEpi {
  void* out;
  const float* bias;
  const void* residual;
  void* aux_out;
  const void* aux_in;
  long long ldo, ldr, ld_aux;
  const __half* rope;
  int rope_hd, rope_D;
  int flags;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;   // TMEM columns between the two accumulator stages
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN <= 256 && BN % 16 == 0, "UMMA N must be a multiple of 16 and <= 256");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-B alignment");
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

// Side inputs of one 32-column chunk of one output row, fetched ahead of the accumulator:
//   bf16 residual -> v[0..3];  fp32 residual -> v[0..7];  bf16 dGELU operand -> v[4..7]
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

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): one rcp + one ex2 instead of erff's two branches.
// Returns erf(z) and e = exp(-z*z) (reused for the Gaussian pdf in gelu').
__device__ __forceinline__ float erf_as(float z, float& e) {
  const float az = fabsf(z);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
  e = exp2f(-az * az * 1.4426950408889634f);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * e;
  return copysignf(r, z);
}
__device__ __forceinline__ float gelu_fast(float x) {
  float e;
  return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f, e));
}
__device__ __forceinline__ float dgelu_fast(float x) {
  float e;
  const float cdf = 0.5f * (1.0f + erf_as(x * 0.70710678118654752f, e));
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

template <int BN>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi& e, const uint32_t (&acc)[32], const EpiSide& s,
                                               long long row, int col0, int N) {
  // 32 consecutive columns of one output row
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  const int flags = e.flags;
  if (flags & VJ_EPI_BIAS) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (col0 + i < N) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
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
  if (flags & VJ_EPI_AUX_OUT) {
    bf16* ap = reinterpret_cast<bf16*>(e.aux_out) + row * e.ld_aux + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      if (col0 + i < N) {
        uint4 u;
        u.x = pack_bf16x2(v[i], v[i + 1]); u.y = pack_bf16x2(v[i + 2], v[i + 3]);
        u.z = pack_bf16x2(v[i + 4], v[i + 5]); u.w = pack_bf16x2(v[i + 6], v[i + 7]);
        *reinterpret_cast<uint4*>(ap + i) = u;
      }
    }
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
  if (flags & VJ_EPI_OUT_F32) {
    float* op = reinterpret_cast<float*>(e.out) + row * e.ldo + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (col0 + i < N) *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  } else {
    bf16* op = reinterpret_cast<bf16*>(e.out) + row * e.ldo + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      if (col0 + i < N) {
        uint4 u;
        u.x = pack_bf16x2(v[i], v[i + 1]); u.y = pack_bf16x2(v[i + 2], v[i + 3]);
        u.z = pack_bf16x2(v[i + 4], v[i + 5]); u.w = pack_bf16x2(v[i + 6], v[i + 7]);
        *reinterpret_cast<uint4*>(op + i) = u;
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmEpi epi, int M,
            int N, int K) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
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

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb, nb;
        tile_coords(tile, num_m, num_n, mb, nb);
        const int m0 = mb * GEMM_BM, n0 = nb * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full[stage], kb * GEMM_BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < GEMM_BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, &full[stage], m0 + c * 64, kb * GEMM_BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full[stage], kb * GEMM_BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, &full[stage], n0 + c * 64, kb * GEMM_BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(GEMM_BM, BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
      for (int kb = 0; kb < num_kb; ++kb) {
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
            umma_bf16(d_tmem, adk, bdk, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue
    const int q = warp & 3;
    const int eg = (warp - 4) >> 2;                 // warpgroup 0: even chunks, 1: odd chunks
    constexpr int NCH = (BN + 31) / 32;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      int mb, nb;
      tile_coords(tile, num_m, num_n, mb, nb);
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1;
      const long long row = (long long)mb * GEMM_BM + q * 32 + lane;
      const bool row_ok = row < M;
      const int n0 = nb * BN;
      const int nlim = min(N, n0 + BN);
      EpiSide side;
#pragma unroll
      for (int i = 0; i < 8; ++i) side.v[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && n0 + eg * 32 < nlim) epilogue_prefetch(epi, side, row, n0 + eg * 32, nlim);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::ACC_STRIDE;
#pragma unroll 1
      for (int c = eg; c < NCH; c += 2) {
        const int col0 = n0 + c * 32;
        if (col0 >= N) break;                       // warp-uniform
        uint32_t acc[32];
        if (BN % 32 != 0 && c == NCH - 1) {         // 16-column tail of the N tile (BN = 176)
          uint32_t lo[16];
          tmem_ld16(taddr + c * 32, lo);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) { acc[i] = lo[i]; acc[16 + i] = 0u; }
        } else {
          tmem_ld32(taddr + c * 32, acc);
          tmem_ld_wait();
        }
        const EpiSide cur = side;
        if (row_ok && col0 + 64 < nlim && c + 2 < NCH) epilogue_prefetch(epi, side, row, col0 + 64, nlim);
        if (row_ok) epilogue_chunk<BN>(epi, acc, cur, row, col0, nlim);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const vj_gemm_args* g, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dimsK[2] = {(uint64_t)g->K, (uint64_t)g->M};
    const uint64_t dimsM[2] = {(uint64_t)g->M, (uint64_t)g->K};
    const uint64_t str[1] = {(uint64_t)g->lda * 2};
    const uint32_t boxK[2] = {64, GEMM_BM};
    const uint32_t boxM[2] = {64, GEMM_BK};
    int r = make_tmap_bf16(&tmA, g->a, 2, A_MN ? dimsM : dimsK, str, A_MN ? boxM : boxK, 128);
    if (r) return r;
  }
  {
    const uint64_t dimsK[2] = {(uint64_t)g->K, (uint64_t)g->N};
    const uint64_t dimsN[2] = {(uint64_t)g->N, (uint64_t)g->K};
    const uint64_t str[1] = {(uint64_t)g->ldb * 2};
    const uint32_t boxK[2] = {64, (uint32_t)BN};
    const uint32_t boxN[2] = {64, GEMM_BK};
    int r = make_tmap_bf16(&tmB, g->b, 2, B_MN ? dimsN : dimsK, str, B_MN ? boxN : boxK, 128);
    if (r) return r;
  }
  GemmEpi e;
  e.out = g->out; e.bias = g->bias; e.residual = g->residual; e.aux_out = g->aux_out; e.aux_in = g->aux_in;
  e.ldo = g->ldo; e.ldr = g->ldr; e.ld_aux = g->ld_aux; e.flags = g->flags;
  e.rope = reinterpret_cast<const __half*>(g->rope_table); e.rope_hd = g->rope_hd; e.rope_D = g->rope_D;

  auto kern = gemm_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int num_m = (int)((g->M + GEMM_BM - 1) / GEMM_BM);
  const int num_n = (int)((g->N + BN - 1) / BN);
  const int tiles = num_m * num_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, e, (int)g->M, (int)g->N, (int)g->K);
  VJ_LAUNCH_CHECK();
  return 0;
}

// Pick the N tile.  Measured on B200 (profiles/r01_*): per-FLOP speed of the mainloop is ~1.0 at BN=256,
// ~0.85 at 176/192 and ~0.7 at 128 (smaller tiles re-read A more often and give the single MMA-issuing
// thread less time per k-block), so a wide tile wins unless it leaves many dead columns.
static int pick_bn(long long N, bool b_mn, long long M) {
  const int cands_k[] = {256, 192, 176, 128};
  const int cands_mn[] = {256, 192, 192, 128};
  const double speed[] = {1.0, 0.86, 0.85, 0.70};
  const int* c = b_mn ? cands_mn : cands_k;
  int best = 128;
  double best_cost = 1e30;
  const long long num_m = (M + 127) / 128;
  const long long sms = sm_count();
  for (int i = 0; i < 4; ++i) {
    const int bn = c[i];
    const long long tiles_n = (N + bn - 1) / bn;
    const long long tiles = tiles_n * num_m;
    const long long waves = (tiles + sms - 1) / sms;              // wave quantisation of the persistent grid
    const double cost = (double)waves * bn / speed[i];
    if (cost < best_cost * 0.999) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace vj

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
  if (g->flags & VJ_EPI_BIAS) VJ_CHECK(g->bias != nullptr, "vj_gemm: bias flag without pointer");
  if (g->flags & VJ_EPI_RESIDUAL) VJ_CHECK(g->residual != nullptr && g->ldr % 8 == 0, "vj_gemm: bad residual");
  if (g->flags & VJ_EPI_AUX_OUT) VJ_CHECK(g->aux_out != nullptr && g->ld_aux % 8 == 0, "vj_gemm: bad aux_out");
  if (g->flags & VJ_EPI_DGELU) VJ_CHECK(g->aux_in != nullptr && g->ld_aux % 8 == 0, "vj_gemm: bad aux_in");
  VJ_CHECK(!((g->flags & VJ_EPI_DGELU) && (g->flags & VJ_EPI_RES_F32)), "vj_gemm: DGELU with an fp32 residual is not supported");
  if (g->flags & VJ_EPI_ROPE) {
    VJ_CHECK(!(g->flags & (VJ_EPI_RESIDUAL | VJ_EPI_DGELU | VJ_EPI_GELU)), "vj_gemm: ROPE combines with BIAS only");
    VJ_CHECK(g->rope_table && (g->rope_hd == 32 || g->rope_hd == 64) && g->rope_D > 0 && g->rope_D % g->rope_hd == 0 &&
                 g->N == 3 * (int64_t)g->rope_D && g->rope_D % 16 == 0,
             "vj_gemm: bad ROPE arguments (hd=%d D=%d N=%lld)", g->rope_hd, g->rope_D, (long long)g->N);
  }
  const bool amn = g->a_mn_major != 0, bmn = g->b_mn_major != 0;
  VJ_CHECK(!(amn && !bmn), "vj_gemm: (A MN-major, B K-major) is not instantiated");
  const int bn = pick_bn(g->N, bmn, g->M);
#define VJ_GEMM_CASE(BN_, A_, B_) \
  if (bn == BN_ && amn == A_ && bmn == B_) return launch_gemm<BN_, A_, B_>(g, stream);
  VJ_GEMM_CASE(256, false, false)
  VJ_GEMM_CASE(192, false, false)
  VJ_GEMM_CASE(176, false, false)
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
