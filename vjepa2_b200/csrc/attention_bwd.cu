// Flash-attention backward on tcgen05 for sm_100a.
//
// One CTA = one (batch, head, 128-key tile); it loops over 128-query tiles.  320 threads:
//   warps 0-7  compute: thread (r = tid & 127, half = tid >> 7) owns key row r (TMEM lane r) and 64 of the 128
//              query columns of S^T / dP^T, and half of the columns of query row r of dQ
//   warp  8    TMA producer (K,V once; Q_i / dO_i double-buffered) + TMEM allocate/free
//   warp  9    MMA issuer, five GEMMs per query tile:
//                S^T  = K Q_i^T          (128 keys x 128 q, K=d)        -> TMEM [0,128)
//                dP^T = V dO_i^T         (128 keys x 128 q, K=d)        -> TMEM [128,256)
//                dV  += P^T dO_i         (128 keys x d,   K=128 q)      -> TMEM [256,256+d)
//                dK  += dS^T Q_i         (128 keys x d,   K=128 q)      -> TMEM [256+d,256+2d)
//                dQ_i = dS K             (128 q x d,      K=128 keys)   -> TMEM [256+2d,256+3d)
//   P^T and dS^T are written to smem as bf16 K-major tiles; the same dS^T buffer is consumed as an MN-major A
//   operand for dQ (no transpose).  dQ partial tiles are staged as fp32 in smem and accumulated across key tiles
//   by the TMA unit (cp.reduce.async.bulk.tensor .add) into an fp32 scratch buffer -- no per-element atomics --
//   which a small follow-up kernel converts to bf16 (applying the adjoint RoPE map on the way).
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>

#include "attn_common.cuh"
#include "../../include/vjepa2_b200.h"

namespace vj {

// Optional in-kernel cycle accounting (-DVJ_ATTN_PROFILE), compute warp 0 of every CTA:
//  [0] CTA total [1] until first S^T/dP^T (prologue) [2] wait sdp [3] tmem ld [4] math [5] wait dq_full
//  [6] stage dQ [7] P/dS store+fence [8] barriers [9] CTAs [10] epilogue (last drain + dK/dV)
#ifdef VJ_ATTN_PROFILE
__device__ unsigned long long g_attn_bwd_prof[16];
#define BP_T0(v) const long long v = clock64()
#define BP_ADD(acc, v) acc += clock64() - v
#else
#define BP_T0(v)
#define BP_ADD(acc, v)
#endif

template <int HD>
struct AttnBwdCfg {
  static constexpr int BT = 128;                       // keys per CTA == queries per iteration
  static constexpr int TILE_BYTES = BT * HD * 2;       // K, V, Q_i, dO_i tiles
  static constexpr int PT_BYTES = BT * BT * 2;         // 32 KB: two [128 rows][128 B] atoms
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;     // 2 stages
  static constexpr int OFF_DO = OFF_Q + 2 * TILE_BYTES; // 2 stages
  static constexpr int OFF_PT = OFF_DO + 2 * TILE_BYTES;
  static constexpr int OFF_DS = OFF_PT + PT_BYTES;
  // the fp32 dQ staging tile (128 rows x HD x 4 B <= 40 KB) ALIASES P^T / dS^T: both are dead once the MMAs of
  // the iteration have retired (dq_full), and are only rewritten after the TMA has finished reading the staging
  static constexpr int OFF_DQ = OFF_PT;
  static constexpr int OFF_LSE = OFF_DS + PT_BYTES;    // 128 floats lse + 128 floats delta
  static constexpr int OFF_BAR = OFF_LSE + 1024;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int COL_ST = 0, COL_DPT = 128, COL_DV = 256, COL_DK = 256 + HD, COL_DQ = 256 + 2 * HD;
  static constexpr int TMEM_COLS = 512;
  static constexpr bool DQ_DENSE = (HD % 32) != 0;     // HD = 80: un-swizzled [128][HD] fp32 staging, one TMA box
  static_assert(HD == 80 || HD == 64 || HD == 32, "head_dim 80, 64 or 32");
  static_assert(COL_DQ + HD <= 512, "TMEM budget");
  static_assert(BT * HD * 4 <= 2 * PT_BYTES, "dQ staging must fit the aliased region");
  static_assert(TILE_BYTES % 1024 == 0, "tiles must keep 1024-B alignment");
};

// delta[b][h][q] = sum_i dO[q, h, i] * O[q, h, i], and the zero-fill of the fp32 dQ accumulators in the same launch.
// A warp owns 32 consecutive query rows of one (sample, head): lane = row, so each lane streams its head's HD bf16 of
// O and dO (whole cache lines, all loads issued up front) and the 32 results are one coalesced store (the old
// thread-per-(row, head) mapping stored with stride S: 0.35 of HBM peak under ncu).  Every thread then clears its
// share of dq_acc with coalesced 16-byte stores, which replaces a cudaMemsetAsync per attention backward.
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                                         float* __restrict__ delta, float4* __restrict__ dq_acc4,
                                                         long long n_acc4, int B, int S, int H) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n_qt = (S + 31) / 32;
  const long long n_warps = (long long)B * H * n_qt;
  if (warp < n_warps) {
    const int qt = (int)(warp % n_qt);
    const long long bh = warp / n_qt;
    const int h = (int)(bh % H);
    const long long b = bh / H;
    const int q = qt * 32 + lane;
    if (q < S) {
      const long long row = b * S + q;
      const bf16* o = out + row * (long long)H * HD + h * HD;
      const bf16* d = dout + row * (long long)H * HD + h * HD;
      uint4 a[HD / 8], c[HD / 8];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        a[i] = __ldg(reinterpret_cast<const uint4*>(o) + i);
        c[i] = __ldg(reinterpret_cast<const uint4*>(d) + i);
      }
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i)
        acc += bf16_lo(a[i].x) * bf16_lo(c[i].x) + bf16_hi(a[i].x) * bf16_hi(c[i].x) + bf16_lo(a[i].y) * bf16_lo(c[i].y) +
               bf16_hi(a[i].y) * bf16_hi(c[i].y) + bf16_lo(a[i].z) * bf16_lo(c[i].z) + bf16_hi(a[i].z) * bf16_hi(c[i].z) +
               bf16_lo(a[i].w) * bf16_lo(c[i].w) + bf16_hi(a[i].w) * bf16_hi(c[i].w);
      delta[bh * S + q] = acc;
    }
  }
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_acc4; i += nthreads) dq_acc4[i] = z;
}

__device__ __forceinline__ void bulk_commit_bwd() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0_bwd() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0_bwd() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// adjoint of the RoPE pair map on 8 consecutive head dims (cos/sin: 8 fp16 each, vj_rope_table layout)
__device__ __forceinline__ void rope_adjoint8(float* g, const uint4 c, const uint4 s) {
  const __half2* ch = reinterpret_cast<const __half2*>(&c);
  const __half2* sh = reinterpret_cast<const __half2*>(&s);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 cc = __half22float2(ch[i]);
    const float2 ss = __half22float2(sh[i]);
    const float g0 = g[2 * i], g1 = g[2 * i + 1];
    g[2 * i] = cc.x * g0 + ss.y * g1;
    g[2 * i + 1] = -ss.x * g0 + cc.y * g1;
  }
}

// dq accumulators fp32 [B*S][D] -> bf16 q-third of dqkv [B*S][3D] (+ adjoint RoPE if a table is given)
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ dq_acc,
                                                              bf16* __restrict__ dqkv, long long rows, int D, int hd,
                                                              const __half* __restrict__ rope) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vec_per_row = D >> 3;
  if (t >= rows * vec_per_row) return;
  const long long row = t / vec_per_row;
  const int c = (int)(t - row * vec_per_row) * 8;
  const float4 a = *reinterpret_cast<const float4*>(dq_acc + row * D + c);
  const float4 b = *reinterpret_cast<const float4*>(dq_acc + row * D + c + 4);
  float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (rope) {
    const __half* tr = rope + row * 2 * hd + (c % hd);
    rope_adjoint8(g, *reinterpret_cast<const uint4*>(tr), *reinterpret_cast<const uint4*>(tr + hd));
  }
  uint4 u;
  u.x = pack_bf16x2(g[0], g[1]); u.y = pack_bf16x2(g[2], g[3]);
  u.z = pack_bf16x2(g[4], g[5]); u.w = pack_bf16x2(g[6], g[7]);
  *reinterpret_cast<uint4*>(dqkv + row * 3 * (long long)D + c) = u;
}

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// 320 threads: warps 0-7 compute (thread: key row r = tid & 127 == TMEM lane, half = tid >> 7 -> 64 of the 128 query
// columns of S^T / dP^T, and half of the columns of the dQ / dK / dV rows), warp 8 TMA producer + TMEM, warp 9 MMA.
template <int HD>
__global__ void __launch_bounds__(320, 1)
attn_bwd_kernel(const __grid_constant__ TMapPair tmQKV, const __grid_constant__ TMapPair tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const float* __restrict__ lse,
                const float* __restrict__ delta, bf16* __restrict__ dqkv, const __half* __restrict__ rope, int S, int H,
                int D, float scale, float scale_log2) {
  using Cfg = AttnBwdCfg<HD>;
  constexpr int HO = HD / 2;                               // output columns per compute thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sDO = smem + Cfg::OFF_DO;
  uint8_t* sPT = smem + Cfg::OFF_PT;
  uint8_t* sDS = smem + Cfg::OFF_DS;
  uint8_t* sDQ = smem + Cfg::OFF_DQ;                        // fp32 dQ tile staging: [HD/32 atoms][128 rows][128 B]
  float* s_lse = reinterpret_cast<float*>(smem + Cfg::OFF_LSE);
  float* s_delta = s_lse + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars;          // 1
  uint64_t* qdo_full = bars + 1;     // 2
  uint64_t* qdo_empty = bars + 3;    // 2
  uint64_t* sdp_full = bars + 5;     // 1
  uint64_t* pds_full = bars + 6;     // 1 (256 arrivals)
  uint64_t* dq_full = bars + 7;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int k0 = blockIdx.x * Cfg::BT;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_q = (S + Cfg::BT - 1) / Cfg::BT;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 8);               // one arrival per compute warp
    mbar_init(dq_full, 1);
    mbar_fence_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQKV.m[0]);
      tma_prefetch_desc(&tmDO.m[0]);
      mbar_expect_tx(kv_full, 2 * Cfg::TILE_BYTES);
      tma_load_head_tile<HD>(sK, &tmQKV, kv_full, Cfg::BT, D + h * HD, k0, b);
      tma_load_head_tile<HD>(sV, &tmQKV, kv_full, Cfg::BT, 2 * D + h * HD, k0, b);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[st], 2 * Cfg::TILE_BYTES);
        tma_load_head_tile<HD>(sQ + st * Cfg::TILE_BYTES, &tmQKV, &qdo_full[st], Cfg::BT, h * HD, i * Cfg::BT, b);
        tma_load_head_tile<HD>(sDO + st * Cfg::TILE_BYTES, &tmDO, &qdo_full[st], Cfg::BT, h * HD, i * Cfg::BT, b);
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t id_s = make_idesc(128, 128, false, false);
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint64_t pt_k = desc_kmajor<128>(smem_u32(sPT));
    const uint64_t ds_k = desc_kmajor<128>(smem_u32(sDS));
    const uint64_t ds_mn = desc_mnmajor<128>(smem_u32(sDS), 16384);
    mbar_wait(kv_full, 0);
    for (int i = 0; i < n_q; ++i) {
      const int st = i & 1;
      mbar_wait(&qdo_full[st], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQ + st * Cfg::TILE_BYTES);
      const uint32_t do_addr = smem_u32(sDO + st * Cfg::TILE_BYTES);
      if (elect_one()) {
        mma_over_hd<HD>(tmem_base + Cfg::COL_ST, k_addr, Cfg::BT, q_addr, Cfg::BT, id_s);     // S^T  = K Q^T
        mma_over_hd<HD>(tmem_base + Cfg::COL_DPT, v_addr, Cfg::BT, do_addr, Cfg::BT, id_s);   // dP^T = V dO^T
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_full, i & 1);
      tc_fence_after();
      if (elect_one()) {
        // dV += P^T dO ; dK += dS^T Q   (A: K-major 2-atom tiles, k-step k lives in atom k/4)
        mma_into_hd<HD, false, Cfg::BT>(tmem_base + Cfg::COL_DV,
                                        [&](int k) { return desc_advance(pt_k, (k >> 2) * 16384 + (k & 3) * 32); },
                                        do_addr, i != 0);
        mma_into_hd<HD, false, Cfg::BT>(tmem_base + Cfg::COL_DK,
                                        [&](int k) { return desc_advance(ds_k, (k >> 2) * 16384 + (k & 3) * 32); },
                                        q_addr, i != 0);
        // dQ_i = dS K   (A: the same dS^T buffer read MN-major)
        mma_into_hd<HD, true, Cfg::BT>(tmem_base + Cfg::COL_DQ, [&](int k) { return desc_advance(ds_mn, k * 2048); },
                                       k_addr, false);
        umma_commit(&qdo_empty[st]);
        umma_commit(dq_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    const int r = threadIdx.x & 127;
    const int half = threadIdx.x >> 7;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const bool key_ok = (k0 + r) < S;
    const float* lse_bh = lse + ((long long)b * H + h) * S;
    const float* delta_bh = delta + ((long long)b * H + h) * S;
    for (int i = 0; i < n_q; ++i) {
      if (half == 0) {
        const int q = i * Cfg::BT + r;
        s_lse[r] = q < S ? lse_bh[q] : INFINITY;
        s_delta[r] = q < S ? delta_bh[q] : 0.f;
      } else if (r == 0 && i > 0) {
        bulk_wait_read0_bwd();                            // dQ staging of iteration i-1 has been read by the TMA
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc2 = 0; cc2 < 2; ++cc2) {
        const int c = half * 2 + cc2;                      // 32-column chunk of the 128 query columns
        uint32_t sv[32], dv[32];
        tmem_ld32(lane_addr + Cfg::COL_ST + c * 32, sv);
        tmem_ld32(lane_addr + Cfg::COL_DPT + c * 32, dv);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(sv[j]), scale_log2, -s_lse[c * 32 + j]));
          float p1 = ex2_approx(fmaf(__uint_as_float(sv[j + 1]), scale_log2, -s_lse[c * 32 + j + 1]));
          if (!key_ok) { p0 = 0.f; p1 = 0.f; }
          const float d0 = p0 * (__uint_as_float(dv[j]) - s_delta[c * 32 + j]);
          const float d1 = p1 * (__uint_as_float(dv[j + 1]) - s_delta[c * 32 + j + 1]);
          pp[j >> 1] = pack_bf16x2(p0, p1);
          dd[j >> 1] = pack_bf16x2(d0, d1);
        }
        // 4 x 16-byte chunks per buffer; chunk cc of the row -> atom cc/8, position (cc%8) ^ (r&7)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c * 4 + u;
          const uint32_t off = (cc >> 3) * 16384 + r * 128 + (((cc & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(sPT + off) = make_uint4(pp[u * 4], pp[u * 4 + 1], pp[u * 4 + 2], pp[u * 4 + 3]);
          *reinterpret_cast<uint4*>(sDS + off) = make_uint4(dd[u * 4], dd[u * 4 + 1], dd[u * 4 + 2], dd[u * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(pds_full);
      // dQ_i tile: lane r == query row r; stage scale*dQ as fp32 and let the TMA reduce-add it into dq_acc
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      {
        uint32_t o[HO];
        tmem_ld_n<HO>(lane_addr + Cfg::COL_DQ + half * HO, o);
        tmem_ld_wait();
        if constexpr (Cfg::DQ_DENSE) {
          float* rowp = reinterpret_cast<float*>(sDQ) + r * HD + half * HO;     // dense [128][HD] fp32
#pragma unroll
          for (int u = 0; u < HO / 4; ++u)
            *reinterpret_cast<float4*>(rowp + u * 4) =
                make_float4(__uint_as_float(o[u * 4]) * scale, __uint_as_float(o[u * 4 + 1]) * scale,
                            __uint_as_float(o[u * 4 + 2]) * scale, __uint_as_float(o[u * 4 + 3]) * scale);
        } else {
          // 128-byte swizzled atoms of 32 fp32: HD = 64 -> atom = half (8 chunks); HD = 32 -> one atom, 4 chunks each
          uint8_t* rowp = sDQ + (HO == 32 ? half * 16384 : 0) + r * 128;
          const int cbase = HO == 32 ? 0 : half * 4;
#pragma unroll
          for (int u = 0; u < HO / 4; ++u)
            *reinterpret_cast<float4*>(rowp + (((cbase + u) ^ (r & 7)) << 4)) =
                make_float4(__uint_as_float(o[u * 4]) * scale, __uint_as_float(o[u * 4 + 1]) * scale,
                            __uint_as_float(o[u * 4 + 2]) * scale, __uint_as_float(o[u * 4 + 3]) * scale);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (threadIdx.x == 128) {                            // same thread that waits on the bulk group above
        if constexpr (Cfg::DQ_DENSE) {
          tma_reduce_add_3d(&tmDQ, sDQ, h * HD, i * Cfg::BT, b);
        } else {
#pragma unroll
          for (int a = 0; a < HD / 32; ++a)
            tma_reduce_add_3d(&tmDQ, sDQ + a * 16384, h * HD + a * 32, i * Cfg::BT, b);
        }
        bulk_commit_bwd();
      }
    }
    if (threadIdx.x == 128) bulk_wait_all0_bwd();
    // dK / dV epilogue (all MMAs retired: dq_full of the last iteration was committed after them)
    const int key = k0 + r;
    bf16* dk_row = dqkv + ((long long)b * S + key) * 3 * D + D + h * HD + half * HO;
    bf16* dv_row = dk_row + D;
    constexpr int NV = HO / 8;
    uint32_t a[HO], c[HO];
    tmem_ld_n<HO>(lane_addr + Cfg::COL_DK + half * HO, a);
    tmem_ld_n<HO>(lane_addr + Cfg::COL_DV + half * HO, c);
    tmem_ld_wait();
    if (key_ok) {
      const __half* tr = rope ? rope + ((long long)b * S + key) * 2 * HD + half * HO : nullptr;
#pragma unroll
      for (int jj = 0; jj < NV; ++jj) {
        const int j = jj * 8;
        uint4 u, w;
        float gk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) gk[e] = __uint_as_float(a[j + e]) * scale;
        if (rope) rope_adjoint8(gk, *reinterpret_cast<const uint4*>(tr + j), *reinterpret_cast<const uint4*>(tr + HD + j));
        u.x = pack_bf16x2(gk[0], gk[1]);
        u.y = pack_bf16x2(gk[2], gk[3]);
        u.z = pack_bf16x2(gk[4], gk[5]);
        u.w = pack_bf16x2(gk[6], gk[7]);
        w.x = pack_bf16x2(__uint_as_float(c[j]), __uint_as_float(c[j + 1]));
        w.y = pack_bf16x2(__uint_as_float(c[j + 2]), __uint_as_float(c[j + 3]));
        w.z = pack_bf16x2(__uint_as_float(c[j + 4]), __uint_as_float(c[j + 5]));
        w.w = pack_bf16x2(__uint_as_float(c[j + 6]), __uint_as_float(c[j + 7]));
        *reinterpret_cast<uint4*>(dk_row + j) = u;
        *reinterpret_cast<uint4*>(dv_row + j) = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// Pipelined persistent variant (head_dim 64 / 32).  The kernel above runs each query tile as one dependent chain
//   S^T,dP^T (MMA) -> P^T,dS^T (compute) -> dV,dK,dQ (MMA) -> dQ drain (compute)
// and the tensor pipe idles through both compute phases (17 % active in ncu).  With one CTA per SM there are
// 200 registers per thread to spend, so here a compute thread copies its 64 S^T and 64 dP^T values out of TMEM
// at once and releases the TMEM tiles (sdp_free): the MMA warp computes S^T,dP^T of tile i+1 while the
// compute warps are still busy with tile i, and dV,dK,dQ of tile i while they start on tile i+1.  The dQ drain of
// tile i-1 is folded into tile i (after the exp math, when dq_full(i-1) has long fired), with its own fp32
// staging buffer so the TMA reduce-add overlaps the next tile.  Per element: one FFMA2 + MUFU.EX2 for P, one
// FFMA2 + FMUL2 for dS (the 1/sqrt(d) factor of dQ and dK is folded into dS), per-column -lse / -delta*scale
// come from smem as 128-bit broadcast loads.
template <int HD>
struct AttnBwd2Cfg {
  static constexpr int BT = 128;
  static constexpr int TILE_BYTES = BT * HD * 2;
  static constexpr int PT_BYTES = BT * BT * 2;
  static constexpr int DQ_BYTES = BT * HD * 4;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int QDO_STAGES = 3;                   // Q_i / dO_i ring: tile i+1 must already be in flight while
                                                         // the MMAs of tile i-1 still read theirs
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;
  static constexpr int OFF_DO = OFF_Q + QDO_STAGES * TILE_BYTES;
  static constexpr int OFF_PT = OFF_DO + QDO_STAGES * TILE_BYTES;
  static constexpr int OFF_DS = OFF_PT + PT_BYTES;
  static constexpr int OFF_DQ = OFF_DS + PT_BYTES;       // fp32 dQ staging: [HD/32 atoms][128 rows][128 B]
  static constexpr int OFF_LSE = OFF_DQ + DQ_BYTES;      // [tile parity][-lse | -delta*scale][128] floats
  static constexpr int OFF_BAR = OFF_LSE + 2 * 2 * 128 * 4;
  // no alignment slack: the dynamic window starts 1024-B aligned (checked at kernel entry), and at head_dim 64
  // the layout needs all but 768 bytes of the 227 KB
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int COL_ST = 0, COL_DPT = 128, COL_DV = 256, COL_DK = 256 + HD, COL_DQ = 256 + 2 * HD;
  static constexpr int TMEM_COLS = 512;
  static_assert(HD == 64 || HD == 32, "pipelined backward: head_dim 64 or 32");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// POLY: share of the P = exp2(.) evaluations done as a polynomial on the FMA pipe instead of MUFU.EX2 (16384
// exponentials per 128 x 128 tile = 1024 MUFU cycles of a ~1475-cycle math phase): 0 none (default), 1 = one pair in
// 8, 2 = one in 4, 4 = one in 2.  Measured on B200 (VJ_ATTN_BWD_POLY): every non-zero share is 1-6 % SLOWER, i.e. the
// phase is bound by the issue / LDS mix, not by MUFU throughput; kept as an experiment switch.
// NP = compute threads per key row: each owns 128 / NP query columns of the S^T / dP^T tile (NP * 4 compute warps).
// NP = 2 is the default; NP = 4 (16 warps, 96 registers) doubles the warps per scheduler (ncu of NP = 2 at d = 32: issue
// slots 29 % busy, 2 warps per scheduler) and shortens the math phase, but not the tile (see launch_attn_bwd).
template <int HD, int POLY, int NP>
__global__ void __launch_bounds__(NP * 128 + 64, 1)
attn_bwd2_kernel(const __grid_constant__ TMapPair tmQKV, const __grid_constant__ TMapPair tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const float* __restrict__ lse,
                 const float* __restrict__ delta, bf16* __restrict__ dqkv, const __half* __restrict__ rope, int S, int H,
                 int D, float scale, float scale_log2, int n_kt, int n_items) {
  using Cfg = AttnBwd2Cfg<HD>;
  constexpr int HO = HD / NP;                              // output columns per compute thread
  constexpr int CW = 128 / NP;                             // S^T / dP^T columns per compute thread
  constexpr int NCT = NP * 128;                            // compute threads
  constexpr int W_PROD = NP * 4, W_MMA = NP * 4 + 1;
  constexpr int NST = Cfg::QDO_STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) {                    // swizzled tiles need 1024-B aligned bases
    if (threadIdx.x == 0) printf("vjepa2_b200: attn_bwd2: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sDO = smem + Cfg::OFF_DO;
  uint8_t* sPT = smem + Cfg::OFF_PT;
  uint8_t* sDS = smem + Cfg::OFF_DS;
  uint8_t* sDQ = smem + Cfg::OFF_DQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars;                  // 1: K, V of the work item landed
  uint64_t* kv_empty = bars + 1;             // 1: the item's last MMAs retired (K, V smem free)
  uint64_t* qdo_full = bars + 2;             // NST
  uint64_t* qdo_empty = qdo_full + NST;      // NST
  uint64_t* sdp_full = qdo_empty + NST;      // 1: S^T, dP^T of tile G are in TMEM
  uint64_t* sdp_free = sdp_full + 1;         // 1 (8 warp arrivals): ... and have been copied to registers
  uint64_t* pds_full = sdp_free + 1;         // 1 (8 warp arrivals): P^T, dS^T of tile G are in smem
  uint64_t* dq_full = pds_full + 1;          // 1: dV, dK, dQ MMAs of tile G retired (P^T/dS^T smem free, dQ_G in TMEM)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_full + 1);

  // Persistent CTA: work items (key tile, head, sample) = blockIdx.x, +gridDim.x, ...; the query tiles of all its
  // items are numbered G = 0, 1, 2, ... and ring stages / barrier phases follow G straight through item boundaries.
  const int warp = threadIdx.x >> 5;
  const int n_q = (S + Cfg::BT - 1) / Cfg::BT;
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < NST; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, NP * 4);
    mbar_init(pds_full, NP * 4);
    mbar_init(dq_full, 1);
    mbar_fence_init();
  }
  if (warp == W_PROD) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_PROD) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQKV.m[0]);
      tma_prefetch_desc(&tmDO.m[0]);
      int G = 0;
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int k0 = (item % n_kt) * Cfg::BT, h = (item / n_kt) % H, b = item / (n_kt * H);
        mbar_wait(kv_empty, (it & 1) ^ 1);
        mbar_expect_tx(kv_full, 2 * Cfg::TILE_BYTES);
        tma_load_head_tile<HD>(sK, &tmQKV, kv_full, Cfg::BT, D + h * HD, k0, b);
        tma_load_head_tile<HD>(sV, &tmQKV, kv_full, Cfg::BT, 2 * D + h * HD, k0, b);
        for (int i = 0; i < n_q; ++i, ++G) {
          const int st = G % NST;
          mbar_wait(&qdo_empty[st], ((G / NST) & 1) ^ 1);
          mbar_expect_tx(&qdo_full[st], 2 * Cfg::TILE_BYTES);
          tma_load_head_tile<HD>(sQ + st * Cfg::TILE_BYTES, &tmQKV, &qdo_full[st], Cfg::BT, h * HD, i * Cfg::BT, b);
          tma_load_head_tile<HD>(sDO + st * Cfg::TILE_BYTES, &tmDO, &qdo_full[st], Cfg::BT, h * HD, i * Cfg::BT, b);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ---------------------------------------------------------------- MMA issuer
    //   S,dP(0) | sdp_free(0): S,dP(1) | pds_full(0): dV,dK,dQ(0) | sdp_free(1): S,dP(2) | pds_full(1): dV,dK,dQ(1) ...
    // (at the last tile of an item the order flips: the next item's S,dP need its K, V, whose smem is only
    //  released by this item's last dV,dK,dQ)
    constexpr uint32_t id_s = make_idesc(128, 128, false, false);
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint64_t pt_k = desc_kmajor<128>(smem_u32(sPT));
    const uint64_t ds_k = desc_kmajor<128>(smem_u32(sDS));
    const uint64_t ds_mn = desc_mnmajor<128>(smem_u32(sDS), 16384);
    auto issue_sdp = [&](int G, bool first_of_item, int it) {
      const int st = G % NST;
      if (first_of_item) mbar_wait(kv_full, it & 1);
      mbar_wait(&qdo_full[st], (G / NST) & 1);
      if (G > 0) mbar_wait(sdp_free, (G - 1) & 1);          // tile G-1 left TMEM for the compute warps' registers
      tc_fence_after();
      if (elect_one()) {
        mma_over_hd<HD>(tmem_base + Cfg::COL_ST, k_addr, Cfg::BT, smem_u32(sQ + st * Cfg::TILE_BYTES), Cfg::BT, id_s);
        mma_over_hd<HD>(tmem_base + Cfg::COL_DPT, v_addr, Cfg::BT, smem_u32(sDO + st * Cfg::TILE_BYTES), Cfg::BT, id_s);
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    int G = 0;
    for (int it = 0; it < n_my; ++it) {
      if (it == 0) issue_sdp(0, true, 0);
      for (int i = 0; i < n_q; ++i, ++G) {
        const int st = G % NST;
        const bool last = (i == n_q - 1);
        if (!last) issue_sdp(G + 1, false, it);
        mbar_wait(pds_full, G & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q_addr = smem_u32(sQ + st * Cfg::TILE_BYTES);
          const uint32_t do_addr = smem_u32(sDO + st * Cfg::TILE_BYTES);
          // dV += P^T dO ; dK += dS^T Q   (A: K-major 2-atom tiles, k-step k lives in atom k/4)
          mma_into_hd<HD, false, Cfg::BT>(tmem_base + Cfg::COL_DV,
                                          [&](int k) { return desc_advance(pt_k, (k >> 2) * 16384 + (k & 3) * 32); },
                                          do_addr, i != 0);
          mma_into_hd<HD, false, Cfg::BT>(tmem_base + Cfg::COL_DK,
                                          [&](int k) { return desc_advance(ds_k, (k >> 2) * 16384 + (k & 3) * 32); },
                                          q_addr, i != 0);
          // dQ_i = dS K   (A: the same dS^T buffer read MN-major)
          mma_into_hd<HD, true, Cfg::BT>(tmem_base + Cfg::COL_DQ, [&](int k) { return desc_advance(ds_mn, k * 2048); },
                                         k_addr, false);
          umma_commit(&qdo_empty[st]);
          umma_commit(dq_full);
          if (last) umma_commit(kv_empty);
        }
        __syncwarp();
        if (last && it + 1 < n_my) issue_sdp(G + 1, true, it + 1);
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    const int r = threadIdx.x & 127;                        // key row == TMEM lane
    const int part = threadIdx.x >> 7;                      // query columns [CW*part, CW*part + CW)
    const int lane = threadIdx.x & 31;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t s_col = smem_u32(smem + Cfg::OFF_LSE);
    // this thread's CW columns of the bf16 P^T / dS^T tiles: 64-column atoms of 128-byte rows, 16-byte chunks
    const int pds_atom = (part * CW) >> 6, pds_cb = ((part * CW) & 63) >> 3;
    const uint32_t pt_row = smem_u32(sPT) + pds_atom * 16384 + r * 128;
    const uint32_t ds_row = smem_u32(sDS) + pds_atom * 16384 + r * 128;
    const int swz = r & 7;
    const uint64_t sc2 = f32x2_pack(scale_log2, scale_log2), s2 = f32x2_pack(scale, scale);
    long long bp2 = 0, bp3 = 0, bp4 = 0, bp5 = 0, bp6 = 0, bp7 = 0, bp8 = 0, bp1 = 0, bp10 = 0, bp11 = 0, bp12 = 0, bp13 = 0;
    (void)bp1; (void)bp2; (void)bp3; (void)bp4; (void)bp5; (void)bp6; (void)bp7; (void)bp8; (void)bp10; (void)bp11;
    (void)bp12; (void)bp13;
    BP_T0(bp_all);

    // dQ tile (TMEM) -> fp32 staging tile; lane r == query row r of the tile
    auto stage_dq = [&]() {
      uint32_t o[HO];
      tmem_ld_n<HO>(lane_addr + Cfg::COL_DQ + part * HO, o);
      tmem_ld_wait();
      // 128-byte swizzled atoms of 32 fp32 columns; this thread's HO columns start at column part * HO
      const uint32_t rowp = smem_u32(sDQ) + ((part * HO) >> 5) * 16384 + r * 128;
      const int cbase = ((part * HO) & 31) >> 2;
#pragma unroll
      for (int u = 0; u < HO / 4; ++u)
        st_shared_v4(rowp + (((cbase + u) ^ swz) << 4), o[u * 4], o[u * 4 + 1], o[u * 4 + 2], o[u * 4 + 3]);
    };

    int G = 0;
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int k0 = (item % n_kt) * Cfg::BT, h = (item / n_kt) % H, b = item / (n_kt * H);
      const float* lse_bh = lse + ((long long)b * H + h) * S;
      const float* delta_bh = delta + ((long long)b * H + h) * S;
      auto reduce_dq = [&](int j) {                         // one thread: TMA reduce-add the staged tile into dq_acc
#pragma unroll
        for (int a = 0; a < HD / 32; ++a) tma_reduce_add_3d(&tmDQ, sDQ + a * 16384, h * HD + a * 32, j * Cfg::BT, b);
        bulk_commit_bwd();
      };
      // per-column -lse and -delta*scale of query tile j -> smem (parity of its global index); written one tile
      // ahead, published by the barrier inside the previous tile
      // (the global loads are issued a phase before their values are stored to smem, so their latency hides
      // behind the TMEM loads and the exp math)
      // (fetch only issues the loads -- no arithmetic on the values, an in-order warp would stall on it)
      float col_l = 0.f, col_d = 0.f;
      auto fetch_cols = [&](int j) {
        if (part == 0 && j < n_q) {
          const int q = j * Cfg::BT + r;
          if (q < S) {
            col_l = __ldg(lse_bh + q);
            col_d = __ldg(delta_bh + q);
          } else {
            col_l = INFINITY;
            col_d = 0.f;
          }
        }
      };
      auto put_cols = [&](int j, int Gj) {
        if (part == 0 && j < n_q) {
          const uint32_t dst = s_col + (Gj & 1) * 1024;
          st_shared_f32(dst + r * 4, -col_l);
          st_shared_f32(dst + 512 + r * 4, -col_d * scale);
        }
      };
      BP_T0(c13);
      fetch_cols(0);
      put_cols(0, G);
      asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
      BP_ADD(bp13, c13);
      for (int i = 0; i < n_q; ++i, ++G) {
        const uint32_t s_par = s_col + (G & 1) * 1024;
        fetch_cols(i + 1);
        BP_T0(c2);
        mbar_wait(sdp_full, G & 1);
        if (i == 0) { BP_ADD(bp1, c2); } else { BP_ADD(bp2, c2); }
        tc_fence_after();
        BP_T0(c3);
        uint32_t sv[CW], dv[CW];
#pragma unroll
        for (int c = 0; c < CW; c += 32) {
          tmem_ld32(lane_addr + Cfg::COL_ST + part * CW + c, reinterpret_cast<uint32_t(&)[32]>(sv[c]));
          tmem_ld32(lane_addr + Cfg::COL_DPT + part * CW + c, reinterpret_cast<uint32_t(&)[32]>(dv[c]));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);               // S^T, dP^T of the next tile may be computed now
        BP_ADD(bp3, c3);
        BP_T0(c4);
        // P = exp2(s*scale*log2e - lse), dS = P * (dP - delta) * scale, on packed fp32 pairs.  Rows of keys >= S
        // and columns of queries >= S need no masking: their K / V / Q / dO rows are TMA zero-fill and lse = +inf
        // there, so every product they reach is an exact zero or lands in a dK / dV row that is never stored.
        uint32_t pp[CW / 2], dd[CW / 2];
#pragma unroll
        for (int j = 0; j < CW; j += 4) {
          uint32_t nl[4], nd[4];
          ld_shared_v4(s_par + (part * CW + j) * 4, nl);
          ld_shared_v4(s_par + 512 + (part * CW + j) * 4, nd);
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const uint64_t x = f32x2_fma(f32x2_pack(__uint_as_float(sv[j + e]), __uint_as_float(sv[j + e + 1])), sc2,
                                         f32x2_pack(__uint_as_float(nl[e]), __uint_as_float(nl[e + 1])));
            float x0, x1;
            f32x2_unpack(x, x0, x1);
            float p0, p1;
            const bool poly = POLY >= 4 ? e == 2 : POLY == 2 ? (e == 2 && (j & 4)) : POLY == 1 ? (e == 2 && (j & 12) == 12)
                                                                                  : false;
            if (poly) {
              exp2_poly_pair(x0, x1, p0, p1);
            } else {
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            const uint64_t t = f32x2_fma(f32x2_pack(__uint_as_float(dv[j + e]), __uint_as_float(dv[j + e + 1])), s2,
                                         f32x2_pack(__uint_as_float(nd[e]), __uint_as_float(nd[e + 1])));
            const uint64_t d = f32x2_mul(f32x2_pack(p0, p1), t);
            float d0, d1;
            f32x2_unpack(d, d0, d1);
            pp[(j + e) >> 1] = pack_bf16x2(p0, p1);
            dd[(j + e) >> 1] = pack_bf16x2(d0, d1);
          }
        }
        BP_ADD(bp4, c4);
        {
          // the staging tile is free once the TMA has read the dQ tile staged a whole tile ago; this barrier also
          // publishes the column constants of the next tile
          BP_T0(c8);
          put_cols(i + 1, G + 1);
          if (threadIdx.x == 128) bulk_wait_read0_bwd();
          asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
          BP_ADD(bp8, c8);
        }
        if (i > 0) {
          // MMAs of the previous tile retired: P^T / dS^T smem may be overwritten, its dQ waits in TMEM
          BP_T0(c5);
          mbar_wait(dq_full, (G - 1) & 1);
          BP_ADD(bp5, c5);
          tc_fence_after();
          BP_T0(c6);
          stage_dq();
          BP_ADD(bp6, c6);
        }
        BP_T0(c7);
        // CW / 8 16-byte chunks per buffer: chunk c of a 128-byte row lands at c ^ (r & 7)
#pragma unroll
        for (int u = 0; u < CW / 8; ++u) {
          st_shared_v4(pt_row + (((pds_cb + u) ^ swz) << 4), pp[u * 4], pp[u * 4 + 1], pp[u * 4 + 2], pp[u * 4 + 3]);
          st_shared_v4(ds_row + (((pds_cb + u) ^ swz) << 4), dd[u * 4], dd[u * 4 + 1], dd[u * 4 + 2], dd[u * 4 + 3]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
        BP_ADD(bp7, c7);
        if (i > 0) {
          BP_T0(c9);
          asm volatile("bar.sync 2, %0;" ::"n"(NCT) : "memory");   // every thread's part of the staged dQ tile is fenced
          if (threadIdx.x == 128) reduce_dq(i - 1);        // (the thread that waits on the bulk group above)
          BP_ADD(bp8, c9);
        }
      }
      // ---- item epilogue: last dQ tile, then dK / dV (the 1/sqrt(d) factor is already in dS)
      // (dK / dV are pulled out of TMEM first: the TMA is still reading the dQ tile staged a moment ago)
      BP_T0(c10);
      mbar_wait(dq_full, (G - 1) & 1);
      BP_ADD(bp11, c10);
      tc_fence_after();
      constexpr int NV = HO / 8;
      uint32_t a[HO], c[HO];
      tmem_ld_n<HO>(lane_addr + Cfg::COL_DK + part * HO, a);
      tmem_ld_n<HO>(lane_addr + Cfg::COL_DV + part * HO, c);
      tmem_ld_wait();
      BP_T0(c12);
      if (threadIdx.x == 128) bulk_wait_read0_bwd();
      asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
      BP_ADD(bp12, c12);
      stage_dq();
      tc_fence_before();                                    // the next item's MMAs overwrite dQ / dK / dV in TMEM
      // (fence + barrier come BEFORE the dK / dV global stores: the proxy fence is a full MEMBAR and would wait
      //  for those stores to be acknowledged)
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, %0;" ::"n"(NCT) : "memory");
      if (threadIdx.x == 128) reduce_dq(n_q - 1);
      const int key = k0 + r;
      if (key < S) {
        bf16* dk_row = dqkv + ((long long)b * S + key) * 3 * D + D + h * HD + part * HO;
        bf16* dv_row = dk_row + D;
        const __half* tr = rope ? rope + ((long long)b * S + key) * 2 * HD + part * HO : nullptr;
#pragma unroll
        for (int jj = 0; jj < NV; ++jj) {
          const int j = jj * 8;
          uint4 u, w;
          float gk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) gk[e] = __uint_as_float(a[j + e]);
          if (rope) rope_adjoint8(gk, *reinterpret_cast<const uint4*>(tr + j), *reinterpret_cast<const uint4*>(tr + HD + j));
          u.x = pack_bf16x2(gk[0], gk[1]);
          u.y = pack_bf16x2(gk[2], gk[3]);
          u.z = pack_bf16x2(gk[4], gk[5]);
          u.w = pack_bf16x2(gk[6], gk[7]);
          w.x = pack_bf16x2(__uint_as_float(c[j]), __uint_as_float(c[j + 1]));
          w.y = pack_bf16x2(__uint_as_float(c[j + 2]), __uint_as_float(c[j + 3]));
          w.z = pack_bf16x2(__uint_as_float(c[j + 4]), __uint_as_float(c[j + 5]));
          w.w = pack_bf16x2(__uint_as_float(c[j + 6]), __uint_as_float(c[j + 7]));
          *reinterpret_cast<uint4*>(dk_row + j) = u;
          *reinterpret_cast<uint4*>(dv_row + j) = w;
        }
      }
      BP_ADD(bp10, c10);
    }
    if (threadIdx.x == 128) bulk_wait_all0_bwd();           // all reduce-adds have landed before the kernel ends
#ifdef VJ_ATTN_PROFILE
    if (threadIdx.x == 0) {
      atomicAdd(&g_attn_bwd_prof[0], (unsigned long long)(clock64() - bp_all));
      atomicAdd(&g_attn_bwd_prof[1], (unsigned long long)bp1);
      atomicAdd(&g_attn_bwd_prof[2], (unsigned long long)bp2);
      atomicAdd(&g_attn_bwd_prof[3], (unsigned long long)bp3);
      atomicAdd(&g_attn_bwd_prof[4], (unsigned long long)bp4);
      atomicAdd(&g_attn_bwd_prof[5], (unsigned long long)bp5);
      atomicAdd(&g_attn_bwd_prof[6], (unsigned long long)bp6);
      atomicAdd(&g_attn_bwd_prof[7], (unsigned long long)bp7);
      atomicAdd(&g_attn_bwd_prof[8], (unsigned long long)bp8);
      atomicAdd(&g_attn_bwd_prof[9], (unsigned long long)n_my);
      atomicAdd(&g_attn_bwd_prof[10], (unsigned long long)bp10);
      atomicAdd(&g_attn_bwd_prof[11], (unsigned long long)bp11);
      atomicAdd(&g_attn_bwd_prof[12], (unsigned long long)bp12);
      atomicAdd(&g_attn_bwd_prof[13], (unsigned long long)bp13);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_PROD) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int HD>
static int launch_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                           void* scratch, const void* rope, int B, int S, int H, cudaStream_t stream) {
  using Cfg = AttnBwdCfg<HD>;
  const int D = H * HD;
  float* dq_acc = reinterpret_cast<float*>(scratch);
  float* delta = dq_acc + (size_t)B * S * D;
  {
    const long long n_warps = (long long)B * H * ((S + 31) / 32);
    attn_delta_kernel<HD><<<(unsigned)((n_warps + 7) / 8), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(out), reinterpret_cast<const bf16*>(dout), delta, reinterpret_cast<float4*>(dq_acc),
        (long long)B * S * D / 4, B, S, H);
    VJ_LAUNCH_CHECK();
  }
  TMapPair tmQKV, tmDO;
  {
    int r = make_head_tmaps<HD>(&tmQKV, qkv, (uint64_t)3 * D, (uint64_t)S, (uint64_t)B, Cfg::BT);
    if (r) return r;
    r = make_head_tmaps<HD>(&tmDO, dout, (uint64_t)D, (uint64_t)S, (uint64_t)B, Cfg::BT);
    if (r) return r;
  }
  CUtensorMap tmDQ;
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)S, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)D * 4, (uint64_t)S * D * 4};
    const uint32_t box[3] = {Cfg::DQ_DENSE ? (uint32_t)HD : 32u, Cfg::BT, 1};
    int r = make_tmap(&tmDQ, dq_acc, VJ_F32, 3, dims, strides, box, Cfg::DQ_DENSE ? 0 : 128);
    if (r) return r;
  }
  dim3 grid((S + Cfg::BT - 1) / Cfg::BT, H, B);
  const float scale = 1.0f / sqrtf((float)HD);
  if constexpr (HD == 80) {
    auto kern = attn_bwd_kernel<HD>;
    static bool attr_set = false;
    if (!attr_set) {
      VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      attr_set = true;
    }
    kern<<<grid, 320, Cfg::SMEM_BYTES, stream>>>(tmQKV, tmDO, tmDQ, lse, delta, reinterpret_cast<bf16*>(dqkv),
                                                reinterpret_cast<const __half*>(rope), S, H, D, scale,
                                                scale * 1.4426950408889634f);
  } else {
    using Cfg2 = AttnBwd2Cfg<HD>;
    static int poly = -1;
    if (poly < 0) {
      const char* e = getenv("VJ_ATTN_BWD_POLY");
      poly = (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 0;   // measured: no share wins here (profiles/)
    }
    static int parts = -1;
    if (parts < 0) {
      // compute threads per key row: 2 (8 warps, default) or 4 (16 warps).  Measured on B200 (profiles/r02m_*): the math
      // phase drops from ~1400 to ~1160 cycles per 128 x 128 tile (MUFU floor 1024) but barriers and waits grow by as
      // much: 0.586 vs 0.576 ms at d = 32, S = 1448 -- the phases are serialised by the tile barriers, not by warps
      const char* e = getenv("VJ_ATTN_BWD_PARTS");
      parts = (e && e[0] == '4') ? 4 : 2;
    }
    auto kern = parts == 4 ? (poly == 0 ? attn_bwd2_kernel<HD, 0, 4> : poly == 1 ? attn_bwd2_kernel<HD, 1, 4>
                              : poly <= 3 ? attn_bwd2_kernel<HD, 2, 4> : attn_bwd2_kernel<HD, 4, 4>)
                           : (poly == 0 ? attn_bwd2_kernel<HD, 0, 2> : poly == 1 ? attn_bwd2_kernel<HD, 1, 2>
                              : poly <= 3 ? attn_bwd2_kernel<HD, 2, 2> : attn_bwd2_kernel<HD, 4, 2>);
    static bool attr_set[2][5] = {{false, false, false, false, false}, {false, false, false, false, false}};
    if (!attr_set[parts == 4][poly]) {
      VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES));
      attr_set[parts == 4][poly] = true;
    }
    const int n_kt = (S + Cfg2::BT - 1) / Cfg2::BT;
    const long long n_items = (long long)n_kt * H * B;
    VJ_CHECK(n_items < (1ll << 30), "vj_attn_bwd: too many (key tile, head, sample) work items");
    const int pgrid = (int)(n_items < sm_count() ? n_items : sm_count());      // persistent: one CTA per SM
    kern<<<pgrid, parts * 128 + 64, Cfg2::SMEM_BYTES, stream>>>(tmQKV, tmDO, tmDQ, lse, delta, reinterpret_cast<bf16*>(dqkv),
                                                                reinterpret_cast<const __half*>(rope), S, H, D, scale,
                                                                scale * 1.4426950408889634f, n_kt, (int)n_items);
  }
  VJ_LAUNCH_CHECK();
  {
    const long long n = (long long)B * S * (D / 8);
    attn_dq_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        dq_acc, reinterpret_cast<bf16*>(dqkv), (long long)B * S, D, HD, reinterpret_cast<const __half*>(rope));
    VJ_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace vj

#ifdef VJ_ATTN_PROFILE
extern "C" int vj_attn_bwd_prof_read(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, vj::g_attn_bwd_prof, 16 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(vj::g_attn_bwd_prof, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" size_t vj_attn_bwd_scratch(int B, int S, int H, int head_dim) {
  return ((size_t)B * S * H * head_dim + (size_t)B * H * S) * sizeof(float);
}

extern "C" int vj_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                           void* scratch, const void* rope_table, int B, int S, int H, int head_dim, void* stream) {
  using namespace vj;
  VJ_CHECK(qkv && out && dout && lse && dqkv && scratch, "vj_attn_bwd: null pointer");
  VJ_CHECK(B > 0 && S > 0 && H > 0, "vj_attn_bwd: bad shape B=%d S=%d H=%d", B, S, H);
  VJ_CHECK(B <= 65535 && H <= 65535, "vj_attn_bwd: B/H exceed grid limits");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (head_dim == 64) return launch_attn_bwd<64>(qkv, out, dout, lse, dqkv, scratch, rope_table, B, S, H, st);
  if (head_dim == 32) return launch_attn_bwd<32>(qkv, out, dout, lse, dqkv, scratch, rope_table, B, S, H, st);
  if (head_dim == 80) return launch_attn_bwd<80>(qkv, out, dout, lse, dqkv, scratch, rope_table, B, S, H, st);
  set_error("vj_attn_bwd: head_dim %d not supported (32, 64, 80)", head_dim);
  return -1;
}
