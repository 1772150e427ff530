// Flash-attention forward on tcgen05 for sm_100a (non-causal, no mask, scale 1/sqrt(d)).
// TMEM (256 columns): S ping-pong [0,64) [64,128), O at [128,128+d).  ~83 KB smem -> 2 CTAs / SM so one CTA's
// exp phase overlaps the other's MMAs.  Q/K/V tiles are read straight out of the fused qkv activation
// [B*S][3D] with a 3-D tensor map (d, S, B) -- no head-major relayout pass.  Roles: see attn_fwd_kernel.
#include "common.cuh"
#include "host_common.h"
#include <stdlib.h>

#include "attn_common.cuh"
#include "../../include/vjepa2_b200.h"

namespace vj {

// Optional in-kernel cycle accounting (-DVJ_ATTN_PROFILE, `make build/selftest_prof`), warp 0 / MMA warp:
//  [0] softmax loop total [1] wait S [2] tmem ld S [3] max/exp/pack [4] wait PV(j-1) [5] O rescale
//  [6] P store + fence + arrive [7] MMA wait P [8] MMA wait K/V [9] CTAs [10] MMA total
#ifdef VJ_ATTN_PROFILE
__device__ unsigned long long g_attn_prof[16];
#define AP_T0(v) const long long v = clock64()
#define AP_ADD(acc, v) acc += clock64() - v
#else
#define AP_T0(v)
#define AP_ADD(acc, v)
#endif

template <int HD>
struct AttnFwdCfg {
  static constexpr int BM = 128, BN = 64, KV_STAGES = (HD == 80) ? 2 : 3;   // two CTAs must fit one SM
  static constexpr int Q_BYTES = BM * HD * 2;
  static constexpr int KV_BYTES = BN * HD * 2;
  static constexpr int P_BYTES = BM * BN * 2;        // 16 KB, 128-B rows
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;
  static constexpr int OFF_X = OFF_P + 2 * P_BYTES;  // (P is double-buffered) row-max / row-sum exchange: [2 parity][2 half][128] floats
  static constexpr int OFF_BAR = OFF_X + 2 * 2 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int O_COL = 128;
  static constexpr int THREADS = 320;                // 8 softmax warps + producer + MMA
  static_assert(HD == 80 || HD == 64 || HD == 32, "head_dim 80, 64 or 32");
  static_assert(Q_BYTES % 1024 == 0 && KV_BYTES % 1024 == 0, "tiles must keep 1024-B alignment");
};

__device__ __forceinline__ void pair_barrier(int q) {   // the two warps that share TMEM lane quarter q
  asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
}

// One CTA = one (batch, head, 128-query tile).  320 threads:
//   warps 0-7  softmax: thread (r = tid & 127, half = tid >> 7) owns columns [32*half, 32*half+32) of row r of
//              every 64-key S tile (TMEM lane r); the two halves of a row exchange their maxima through smem
//   warp  8    TMA producer (Q once, K/V ring of 3 stages) + TMEM allocate/free
//   warp  9    MMA issuer:  S_j = Q K_j^T  (M128 x N64 x K=d),  O += P_j V_j  (M128 x N=d x K64)
// The running max is only advanced when it grows by more than 2^8 (lazy rescale), so the O correction in TMEM
// is rare after the first tiles; exp2 goes straight to MUFU (ex2.approx), the kernel's limiting pipe.
template <int HD>
__global__ void __launch_bounds__(320, 2)
attn_fwd_kernel(const __grid_constant__ TMapPair tmQ, const __grid_constant__ TMapPair tmKV,
                bf16* __restrict__ out, float* __restrict__ lse, int S, int H, int D, float scale_log2) {
  using Cfg = AttnFwdCfg<HD>;
  constexpr int HO = HD / 2;                         // O columns per softmax thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sP = smem + Cfg::OFF_P;          // two 16 KB buffers (tile parity)
  float* xch = reinterpret_cast<float*>(smem + Cfg::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  constexpr int NST = Cfg::KV_STAGES;
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // NST
  uint64_t* k_empty = k_full + NST;        // NST
  uint64_t* v_full = k_empty + NST;        // NST
  uint64_t* v_empty = v_full + NST;        // NST
  uint64_t* s_full = v_empty + NST;        // 2
  uint64_t* p_full = s_full + 2;           // 2 (8 arrivals each: one per softmax warp)
  uint64_t* pv_done = p_full + 2;          // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * Cfg::BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (S + Cfg::BN - 1) / Cfg::BN;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);           // one arrival per softmax warp
      mbar_init(&pv_done[i], 1);
    }
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
      tma_prefetch_desc(&tmQ.m[0]);
      tma_prefetch_desc(&tmKV.m[0]);
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      tma_load_head_tile<HD>(sQ, &tmQ, q_full, Cfg::BM, h * HD, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % NST;
        const uint32_t ph = (j / NST) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::KV_BYTES);
        tma_load_head_tile<HD>(sK + st * Cfg::KV_BYTES, &tmKV, &k_full[st], Cfg::BN, D + h * HD, j * Cfg::BN, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], Cfg::KV_BYTES);
        tma_load_head_tile<HD>(sV + st * Cfg::KV_BYTES, &tmKV, &v_full[st], Cfg::BN, 2 * D + h * HD, j * Cfg::BN, b);
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_qk = make_idesc(128, Cfg::BN, false, false);
    const uint64_t pd0 = desc_kmajor<128>(smem_u32(sP));
    mbar_wait(q_full, 0);
    long long ap_p = 0, ap_kv = 0;
    (void)ap_p; (void)ap_kv;
    AP_T0(ap_all);
    auto issue_qk = [&](int j) {
      const int st = j % NST;
      AP_T0(a0);
      mbar_wait(&k_full[st], (j / NST) & 1);
      AP_ADD(ap_kv, a0);
      tc_fence_after();
      if (elect_one()) {
        mma_over_hd<HD>(tmem_base + (j & 1) * Cfg::BN, smem_u32(sQ), Cfg::BM, smem_u32(sK + st * Cfg::KV_BYTES), Cfg::BN,
                        idesc_qk);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      // S_{j+1} overwrites the buffer of S_{j-1}: every softmax warp finished reading it before p_full(j-1)
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int st = j % NST;
      AP_T0(a1);
      mbar_wait(&v_full[st], (j / NST) & 1);
      AP_ADD(ap_kv, a1);
      AP_T0(a2);
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      AP_ADD(ap_p, a2);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t pd = desc_advance(pd0, (j & 1) * Cfg::P_BYTES);
        mma_into_hd<HD, false, Cfg::BN>(tmem_base + Cfg::O_COL, [&](int k) { return desc_advance(pd, k * 32); },
                                        smem_u32(sV + st * Cfg::KV_BYTES), j != 0);
        umma_commit(&v_empty[st]);
        umma_commit(&pv_done[j & 1]);
      }
      __syncwarp();
    }
#ifdef VJ_ATTN_PROFILE
    if (lane == 0) {
      atomicAdd(&g_attn_prof[7], (unsigned long long)ap_p);
      atomicAdd(&g_attn_prof[8], (unsigned long long)ap_kv);
      atomicAdd(&g_attn_prof[10], (unsigned long long)(clock64() - ap_all));
      atomicAdd(&g_attn_prof[9], 1ull);
    }
#endif
  } else {
    // ---------------------------------------------------------------- softmax / correction / epilogue
    const int r = threadIdx.x & 127;                        // query row in tile == TMEM lane
    const int half = threadIdx.x >> 7;                      // which 32 of the 64 S columns
    const int q = warp & 3;                                 // TMEM lane quarter (shared with warp ^ 4)
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t prow = smem_u32(sP) + r * 128;
    const uint32_t xch_mine = smem_u32(xch) + (half * 128 + r) * 4;
    const uint32_t xch_peer = smem_u32(xch) + ((half ^ 1) * 128 + r) * 4;
    long long ap1 = 0, ap2 = 0, ap3 = 0, ap4 = 0, ap5 = 0, ap6 = 0;
    (void)ap1; (void)ap2; (void)ap3; (void)ap4; (void)ap5; (void)ap6;
    AP_T0(ap_tot);
    uint32_t sr[32];
    mbar_wait(&s_full[0], 0);
    tc_fence_after();
    tmem_ld32(lane_addr + half * 32, sr);
    tmem_ld_wait_regs(sr);
    for (int j = 0; j < n_tiles; ++j) {
      // sr holds this thread's 32 columns of S_j
      AP_T0(b3);
      const int valid = S - j * Cfg::BN - half * 32;        // columns of this half that are real keys
      if (valid < 32) {                                     // only in the last tile
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) sr[i] = 0xff800000u;              // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
      }
      const float mloc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const uint32_t xoff = (j & 1) * 1024;
      st_shared_f32(xch_mine + xoff, mloc);
      pair_barrier(q);
      const float m_tile = fmaxf(mloc, ld_shared_f32(xch_peer + xoff)) * scale_log2;
      // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger
      float alpha = 1.0f;
      if (m_tile > m_run + 8.0f) {
        alpha = ex2_approx(m_run - m_tile);                 // 0 on the first tile
        m_run = m_tile;
      }
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sr[i]), scale_log2, -m_run));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sr[i + 1]), scale_log2, -m_run));
        const float p2 = ex2_approx(fmaf(__uint_as_float(sr[i + 2]), scale_log2, -m_run));
        const float p3 = ex2_approx(fmaf(__uint_as_float(sr[i + 3]), scale_log2, -m_run));
        rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3;
        pk[i >> 1] = pack_bf16x2(p0, p1);
        pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
      }
      l_run = l_run * alpha + ((rs0 + rs1) + (rs2 + rs3));
      AP_ADD(ap3, b3);
      // start fetching S_{j+1} (issued by the MMA warp a whole tile ago) while P_j is being published
      const bool more = j + 1 < n_tiles;
      if (more) {
        AP_T0(b1);
        mbar_wait(&s_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        AP_ADD(ap1, b1);
        tc_fence_after();
        tmem_ld32(lane_addr + ((j + 1) & 1) * Cfg::BN + half * 32, sr);
      }
      if (j >= 2) {                                         // P buffer (j & 1) was last read by PV_{j-2}
        AP_T0(b4);
        mbar_wait(&pv_done[j & 1], ((j - 2) >> 1) & 1);
        AP_ADD(ap4, b4);
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        // O correction needs exclusive access to the accumulator: PV_{j-1} must have landed
        AP_T0(b5);
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t o[HO];
        tmem_ld_n<HO>(lane_addr + Cfg::O_COL + half * HO, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HO; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_n<HO>(lane_addr + Cfg::O_COL + half * HO, o);
        tmem_st_wait();
        AP_ADD(ap5, b5);
      }
      AP_T0(b6);
      // P half-row -> smem, K-major 128-B swizzled rows: 16-B chunk c lands at c ^ (r & 7)
      const uint32_t pdst = prow + (j & 1) * Cfg::P_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        st_shared_v4(pdst + (((half * 4 + c) ^ (r & 7)) << 4), pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
      AP_ADD(ap6, b6);
      if (more) {
        AP_T0(b2);
        tmem_ld_wait_regs(sr);
        AP_ADD(ap2, b2);
      }
    }
#ifdef VJ_ATTN_PROFILE
    if (threadIdx.x == 0) {
      atomicAdd(&g_attn_prof[0], (unsigned long long)(clock64() - ap_tot));
      atomicAdd(&g_attn_prof[1], (unsigned long long)ap1);
      atomicAdd(&g_attn_prof[2], (unsigned long long)ap2);
      atomicAdd(&g_attn_prof[3], (unsigned long long)ap3);
      atomicAdd(&g_attn_prof[4], (unsigned long long)ap4);
      atomicAdd(&g_attn_prof[5], (unsigned long long)ap5);
      atomicAdd(&g_attn_prof[6], (unsigned long long)ap6);
    }
#endif
    // combine the two partial row sums (same reference max in both halves)
    const uint32_t xoff = (n_tiles & 1) * 1024;
    st_shared_f32(xch_mine + xoff, l_run);
    pair_barrier(q);
    const float l_tot = l_run + ld_shared_f32(xch_peer + xoff);
    mbar_wait(&pv_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
    tc_fence_after();
    const int qrow = q0 + r;
    const float inv_l = 1.0f / l_tot;
    bf16* orow = out + ((long long)b * S + qrow) * D + h * HD + half * HO;
    {
      uint32_t o[HO];
      tmem_ld_n<HO>(lane_addr + Cfg::O_COL + half * HO, o);
      tmem_ld_wait();
      if (qrow < S) {
#pragma unroll
        for (int i = 0; i < HO; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + i) = u;
        }
      }
    }
    if (half == 0 && qrow < S) lse[((long long)b * H + h) * S + qrow] = m_run + log2f(l_tot);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// Dual-stream persistent variant (head_dim 64 / 32).
//
// The kernel above is a single dependent chain per CTA (S -> max -> exp -> P -> PV -> ...): with two CTAs per SM
// only two chains share the MUFU pipe, which idles whenever both sit in their latency-bound phases, and every
// CTA pays ~3 us of start-up (launch, TMEM alloc, first Q/K fetch) for as little as 8 KV tiles of work.  Here:
//  * the even and the odd KV tiles are two INDEPENDENT online-softmax streams: softmax warpgroup g (128 threads,
//    thread = full 64-column row, no cross-thread max exchange) owns S buffer g, P buffer g and its own O
//    accumulator g in TMEM; the (m, l, O) partial results are merged once per work item.  Four chains per SM.
//  * CTAs are persistent: each walks work items (query tile, head, sample) = blockIdx.x, +gridDim.x, ...; KV tiles
//    are numbered globally across items (T = 0, 1, 2, ...; stream = T & 1; ring stage = T % 3), so the TMA
//    rings, the MMA issue order and the barrier phases run straight through item boundaries and the next
//    item's Q/K/V arrive while the current item's last tiles are still in the softmax.
//  * one warp issues every QK_T and another every PV_T, both in global tile order: each TMA ring is consumed in
//    order by exactly one agent, so a parity wait is never more than one phase away from the barrier's phase
//    (per-stream issuers skip every other fill of a stage and their parity waits can pass vacuously).
// TMEM: S0 [0,64) S1 [64,128) O0 [128,128+d) O1 [128+d,128+2d).
// Register budget: 2 CTAs x 384 threads -> 80 at launch; setmaxnreg moves registers from the producer/MMA
// warpgroup (32) to the softmax warpgroups (104).
template <int HD>
struct AttnFwd2Cfg {
  static constexpr int BM = 128, BN = 64, KV_STAGES = 3;
  static constexpr int Q_BYTES = BM * HD * 2;
  static constexpr int KV_BYTES = BN * HD * 2;
  static constexpr int P_BYTES = BM * BN * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;     // one P buffer per stream
  static constexpr int OFF_X = OFF_P + 2 * P_BYTES;              // merge exchange: [item parity][m|l][2 streams][128] floats
  static constexpr int OFF_BAR = OFF_X + 2 * 2 * 2 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int O_COL = 128;
  static constexpr int THREADS = 384;                            // 2 softmax warpgroups + {K producer, QK issuer, V producer, PV issuer}
  static_assert(HD == 64 || HD == 32, "dual-stream kernel: head_dim 64 or 32");
  static_assert(O_COL + 2 * HD <= TMEM_COLS, "TMEM budget");
};

// POLY8: how many of every 8 probability pairs (16 keys) take the polynomial instead of MUFU.EX2 (0, 2, 3 or 4)
template <int HD, int POLY8>
__global__ void __launch_bounds__(384, 2)
attn_fwd2_kernel(const __grid_constant__ TMapPair tmQ, const __grid_constant__ TMapPair tmKV,
                 bf16* __restrict__ out, float* __restrict__ lse, int S, int H, int D, float scale_log2, int n_qt,
                 int n_items) {
  using Cfg = AttnFwd2Cfg<HD>;
  constexpr int HO = HD / 2;                         // output columns per thread in the merge
  constexpr int NST = Cfg::KV_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sP = smem + Cfg::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars;                 // 1: Q of the item landed
  uint64_t* q_empty = bars + 1;            // 1: last QK of the item finished reading Q
  uint64_t* k_full = bars + 2;             // NST
  uint64_t* k_empty = k_full + NST;        // NST
  uint64_t* v_full = k_empty + NST;        // NST
  uint64_t* v_empty = v_full + NST;        // NST
  uint64_t* s_full = v_empty + NST;        // 2: S_T of stream T&1 landed in TMEM
  uint64_t* s_free = s_full + 2;           // 2: the stream copied S_T to registers (4 warp arrivals)
  uint64_t* p_full = s_free + 2;           // 2: P_T is in smem (4 warp arrivals)
  uint64_t* pv_done = p_full + 2;          // 2: O_g += P_T V_T finished
  uint64_t* o_free = pv_done + 2;          // 1: both accumulators of the item were read out (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (S + Cfg::BN - 1) / Cfg::BN;
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // work items of this CTA


  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(o_free, 8);
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

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 8) {
      // -------------------------------------------------------------- TMA producer: Q of each item + the K ring
      if (elect_one()) {
        tma_prefetch_desc(&tmQ.m[0]);
        tma_prefetch_desc(&tmKV.m[0]);
        int T = 0;
        for (int i = 0; i < n_my; ++i) {
          const int item = blockIdx.x + i * gridDim.x;
          const int q0 = (item % n_qt) * Cfg::BM, h = (item / n_qt) % H, b = item / (n_qt * H);
          mbar_wait(q_empty, (i & 1) ^ 1);
          mbar_expect_tx(q_full, Cfg::Q_BYTES);
          tma_load_head_tile<HD>(sQ, &tmQ, q_full, Cfg::BM, h * HD, q0, b);
          for (int t = 0; t < n_tiles; ++t, ++T) {
            const int st = T % NST;
            mbar_wait(&k_empty[st], ((T / NST) & 1) ^ 1);
            mbar_expect_tx(&k_full[st], Cfg::KV_BYTES);
            tma_load_head_tile<HD>(sK + st * Cfg::KV_BYTES, &tmKV, &k_full[st], Cfg::BN, D + h * HD, t * Cfg::BN, b);
          }
        }
      }
    } else if (warp == 10) {
      // -------------------------------------------------------------- TMA producer: the V ring (own warp: QK runs
      // up to four tiles ahead of PV, so K refills must not queue behind V refills)
      if (elect_one()) {
        int T = 0;
        for (int i = 0; i < n_my; ++i) {
          const int item = blockIdx.x + i * gridDim.x;
          const int h = (item / n_qt) % H, b = item / (n_qt * H);
          for (int t = 0; t < n_tiles; ++t, ++T) {
            const int st = T % NST;
            mbar_wait(&v_empty[st], ((T / NST) & 1) ^ 1);
            mbar_expect_tx(&v_full[st], Cfg::KV_BYTES);
            tma_load_head_tile<HD>(sV + st * Cfg::KV_BYTES, &tmKV, &v_full[st], Cfg::BN, 2 * D + h * HD, t * Cfg::BN, b);
          }
        }
      }
    } else if (warp == 9) {
      // -------------------------------------------------------------- QK^T issuer: S_T = Q K_T^T, in global tile order
      // (the only consumer of the K ring, so its parity waits follow the ring's phases one by one).  It runs as
      // far ahead of the softmax as the two S buffers and the K ring allow.
      constexpr uint32_t idesc_qk = make_idesc(128, Cfg::BN, false, false);
      long long mp_q = 0, mp_sf = 0, mp_k = 0;
      (void)mp_q; (void)mp_sf; (void)mp_k;
      AP_T0(mp_all);
      int T = 0;
      for (int i = 0; i < n_my; ++i) {
        AP_T0(m0);
        mbar_wait(q_full, i & 1);
        AP_ADD(mp_q, m0);
        for (int t = 0; t < n_tiles; ++t, ++T) {
          const int st = T % NST;
          AP_T0(m1);
          if (T >= 2) mbar_wait(&s_free[T & 1], ((T >> 1) - 1) & 1);   // S_{T-2} is in the stream's registers
          AP_ADD(mp_sf, m1);
          AP_T0(m2);
          mbar_wait(&k_full[st], (T / NST) & 1);
          AP_ADD(mp_k, m2);
          tc_fence_after();
          if (elect_one()) {
            mma_over_hd<HD>(tmem_base + (T & 1) * Cfg::BN, smem_u32(sQ), Cfg::BM, smem_u32(sK + st * Cfg::KV_BYTES),
                            Cfg::BN, idesc_qk);
            umma_commit(&k_empty[st]);
            umma_commit(&s_full[T & 1]);
            if (t == n_tiles - 1) umma_commit(q_empty);
          }
          __syncwarp();
        }
      }
#ifdef VJ_ATTN_PROFILE
      if (lane == 0) {
        atomicAdd(&g_attn_prof[8], (unsigned long long)mp_k);
        atomicAdd(&g_attn_prof[10], (unsigned long long)(clock64() - mp_all));
        atomicAdd(&g_attn_prof[11], (unsigned long long)mp_q);
        atomicAdd(&g_attn_prof[12], (unsigned long long)mp_sf);
      }
#endif
    } else {
      // -------------------------------------------------------------- PV issuer: O_{T&1} += P_T V_T, in global tile
      // order (the only consumer of the V ring)
      const uint64_t pd0 = desc_kmajor<128>(smem_u32(sP));
      long long mp_v = 0, mp_p = 0, mp_o = 0;
      (void)mp_v; (void)mp_p; (void)mp_o;
      AP_T0(mp_all);
      int T = 0;
      for (int i = 0; i < n_my; ++i) {
        for (int t = 0; t < n_tiles; ++t, ++T) {
          const int g = T & 1, st = T % NST;
          AP_T0(m3);
          if (t < 2 && i > 0) mbar_wait(o_free, (i - 1) & 1);       // previous item's accumulators were read out
          AP_ADD(mp_o, m3);
          AP_T0(m4);
          mbar_wait(&v_full[st], (T / NST) & 1);
          AP_ADD(mp_v, m4);
          AP_T0(m5);
          mbar_wait(&p_full[g], (T >> 1) & 1);
          AP_ADD(mp_p, m5);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t pd = desc_advance(pd0, g * Cfg::P_BYTES);
            mma_into_hd<HD, false, Cfg::BN>(tmem_base + Cfg::O_COL + g * HD,
                                            [&](int k) { return desc_advance(pd, k * 32); },
                                            smem_u32(sV + st * Cfg::KV_BYTES), t >= 2);
            umma_commit(&v_empty[st]);
            umma_commit(&pv_done[g]);
          }
          __syncwarp();
        }
      }
#ifdef VJ_ATTN_PROFILE
      if (lane == 0) {
        atomicAdd(&g_attn_prof[7], (unsigned long long)mp_p);
        atomicAdd(&g_attn_prof[15], (unsigned long long)(clock64() - mp_all));
        atomicAdd(&g_attn_prof[13], (unsigned long long)mp_v);
        atomicAdd(&g_attn_prof[14], (unsigned long long)mp_o);
      }
#endif
    }
  } else {
    // ---------------------------------------------------------------- softmax stream g
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int r = threadIdx.x & 127;                        // query row in tile == TMEM lane
    const int g = threadIdx.x >> 7;                         // stream: global KV tiles T with T % 2 == g
    const int q = warp & 3;                                 // TMEM lane quarter
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_addr = lane_addr + g * Cfg::BN;
    const uint32_t o_addr = lane_addr + Cfg::O_COL + g * HD;
    const uint32_t prow = smem_u32(sP) + g * Cfg::P_BYTES + r * 128;
    const uint32_t xbase = smem_u32(smem + Cfg::OFF_X);
    const int swz = r & 7;
    long long ap1 = 0, ap2 = 0, ap3 = 0, ap4 = 0, ap5 = 0, ap6 = 0;
    (void)ap1; (void)ap2; (void)ap3; (void)ap4; (void)ap5; (void)ap6;
    AP_T0(ap_tot);
    for (int i = 0; i < n_my; ++i) {
      const int item = blockIdx.x + i * gridDim.x;
      const int q0 = (item % n_qt) * Cfg::BM, h = (item / n_qt) % H, b = item / (n_qt * H);
      const int T0 = i * n_tiles;
      float m_run = -INFINITY, l_run = 0.f;
      int n_mine = 0;
      for (int t = (g - T0) & 1; t < n_tiles; t += 2, ++n_mine) {
        const int c = (T0 + t) >> 1;                        // this stream's running tile count = barrier phase
        AP_T0(b1);
        mbar_wait(&s_full[g], c & 1);
        AP_ADD(ap1, b1);
        tc_fence_after();
        uint32_t sa[32], sb[32];
        AP_T0(b2);
        tmem_ld32(s_addr, sa);
        tmem_ld32(s_addr + 32, sb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);             // QK_{T+2} may overwrite the S buffer now
        AP_ADD(ap2, b2);
        AP_T0(b3);
        const int valid = S - t * Cfg::BN;                  // real keys in this tile
        if (valid < Cfg::BN) {                              // only in the last tile of an item
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            if (k >= valid) sa[k] = 0xff800000u;            // -inf
            if (k + 32 >= valid) sb[k] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[k]), __uint_as_float(sb[k])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sa[k + 1]), __uint_as_float(sb[k + 1])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sa[k + 2]), __uint_as_float(sb[k + 2])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sa[k + 3]), __uint_as_float(sb[k + 3])));
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
        // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger
        float alpha = 1.0f;
        if (m_tile > m_run + 8.0f) {
          alpha = ex2_approx(m_run - m_tile);               // 0 on the first tile
          m_run = m_tile;
        }
        AP_ADD(ap3, b3);
#ifndef VJ_FWD2_LATE_PVWAIT
        AP_T0(b4);
        if (n_mine > 0) {
          // PV of this stream's previous tile: P buffer free again, O accumulator quiescent
          mbar_wait(&pv_done[g], (c - 1) & 1);
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HD; c0 += 8) {            // rare path: small chunks keep the S registers resident
              uint32_t o[8];
              tmem_ld8(o_addr + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
              tmem_st8(o_addr + c0, o);
            }
            tmem_st_wait();
          }
        }
        AP_ADD(ap4, b4);
        AP_T0(b5);
        // exp2(s * scale - m) on packed fp32 pairs (FFMA2 / FADD2 halve the non-MUFU issue slots); every 8 keys
        // make one 16-byte chunk of the P row, stored right away so the smem writes hide under the MUFU work
        const uint64_t sc2 = f32x2_pack(scale_log2, scale_log2), nm2 = f32x2_pack(-m_run, -m_run);
        uint64_t rsum[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t* src = ch < 4 ? &sa[ch * 8] : &sb[(ch - 4) * 8];
          uint32_t pk[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t x =
                f32x2_fma(f32x2_pack(__uint_as_float(src[2 * k]), __uint_as_float(src[2 * k + 1])), sc2, nm2);
            float x0, x1;
            f32x2_unpack(x, x0, x1);
            float p0, p1;
            // pairs of a 16-key group (two chunks) that go to the FMA pipe: chosen at compile time, spread over the group
            const bool kPoly = (POLY8 >= 4) ? (k & 1) == 1
                                 : (POLY8 == 3) ? ((ch & 1) ? (k & 1) == 1 : k == 3)
                                 : (POLY8 == 2) ? k == 3 : false;
            if (kPoly) {
              exp2_poly_pair(x0, x1, p0, p1);
            } else {
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            rsum[k] = f32x2_add(rsum[k], f32x2_pack(p0, p1));
            pk[k] = pack_bf16x2(p0, p1);
          }
          // K-major 128-B swizzled rows: 16-B chunk ch lands at ch ^ (r & 7)
          st_shared_v4(prow + ((ch ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        {
          const uint64_t u = f32x2_add(f32x2_add(rsum[0], rsum[1]), f32x2_add(rsum[2], rsum[3]));
          float u0, u1;
          f32x2_unpack(u, u0, u1);
          l_run = l_run * alpha + (u0 + u1);
        }
        AP_ADD(ap5, b5);
        AP_T0(b6);
#else
        AP_T0(b5);
        // exp2(s * scale - m) on packed fp32 pairs: FFMA2 / FADD2 halve the non-MUFU issue slots.  All 64
        // probabilities are packed to bf16 in registers first: the P buffer / O accumulator are only needed
        // afterwards, so the wait for this stream's previous PV hides behind the exp phase.
        const uint64_t sc2 = f32x2_pack(scale_log2, scale_log2), nm2 = f32x2_pack(-m_run, -m_run);
        uint64_t rsum[4] = {0ull, 0ull, 0ull, 0ull};
        uint32_t pk[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const uint32_t* src = k < 16 ? &sa[2 * k] : &sb[2 * (k - 16)];
          const uint64_t x = f32x2_fma(f32x2_pack(__uint_as_float(src[0]), __uint_as_float(src[1])), sc2, nm2);
          float x0, x1;
          f32x2_unpack(x, x0, x1);
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          rsum[k & 3] = f32x2_add(rsum[k & 3], f32x2_pack(p0, p1));
          pk[k] = pack_bf16x2(p0, p1);
        }
        {
          const uint64_t u = f32x2_add(f32x2_add(rsum[0], rsum[1]), f32x2_add(rsum[2], rsum[3]));
          float u0, u1;
          f32x2_unpack(u, u0, u1);
          l_run = l_run * alpha + (u0 + u1);
        }
        AP_ADD(ap5, b5);
        AP_T0(b4);
        if (n_mine > 0) {
          // PV of this stream's previous tile: P buffer free again, O accumulator quiescent
          mbar_wait(&pv_done[g], (c - 1) & 1);
          if (__any_sync(0xffffffffu, alpha != 1.0f)) {
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HD; c0 += 8) {            // rare path
              uint32_t o[8];
              tmem_ld8(o_addr + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
              tmem_st8(o_addr + c0, o);
            }
            tmem_st_wait();
          }
        }
        AP_ADD(ap4, b4);
        AP_T0(b6);
        // P row -> smem, K-major 128-B swizzled rows: 16-B chunk ch (8 keys) lands at ch ^ (r & 7)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          st_shared_v4(prow + ((ch ^ swz) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
#endif
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        AP_ADD(ap6, b6);
      }
      // ---- merge the two streams: thread (r, g) finishes output columns [g*HO, g*HO+HO) of row r
      const int n_peer = n_tiles - n_mine;
      const uint32_t xm = xbase + (i & 1) * 2048;
      st_shared_f32(xm + (g * 128 + r) * 4, m_run);
      st_shared_f32(xm + (256 + g * 128 + r) * 4, l_run);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float m_peer = ld_shared_f32(xm + ((g ^ 1) * 128 + r) * 4);
      const float l_peer = ld_shared_f32(xm + (256 + (g ^ 1) * 128 + r) * 4);
      const float m_all = fmaxf(m_run, m_peer);
      const float w_mine = n_mine > 0 ? ex2_approx(m_run - m_all) : 0.f;
      const float w_peer = n_peer > 0 ? ex2_approx(m_peer - m_all) : 0.f;
      const float l_tot = l_run * w_mine + l_peer * w_peer;
      const float inv_l = 1.0f / l_tot;
      const float w0 = (g == 0 ? w_mine : w_peer) * inv_l, w1 = (g == 0 ? w_peer : w_mine) * inv_l;
      const bool has0 = (g == 0 ? n_mine : n_peer) > 0, has1 = (g == 0 ? n_peer : n_mine) > 0;
      {
        // the last PV of each stream in this item: global tiles T_end and T_end - 1
        const int T_end = T0 + n_tiles - 1;
        mbar_wait(&pv_done[T_end & 1], (T_end >> 1) & 1);
        if (n_tiles > 1) mbar_wait(&pv_done[(T_end - 1) & 1], ((T_end - 1) >> 1) & 1);
      }
      tc_fence_after();
      const int qrow = q0 + r;
      bf16* orow = out + ((long long)b * S + qrow) * D + h * HD + g * HO;
      {
        uint32_t o0[HO], o1[HO];
        if (has0) {
          tmem_ld_n<HO>(lane_addr + Cfg::O_COL + g * HO, o0);
        } else {
#pragma unroll
          for (int k = 0; k < HO; ++k) o0[k] = 0u;
        }
        if (has1) {
          tmem_ld_n<HO>(lane_addr + Cfg::O_COL + HD + g * HO, o1);
        } else {
#pragma unroll
          for (int k = 0; k < HO; ++k) o1[k] = 0u;
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);                 // the next item's first PVs may overwrite O0 / O1
        if (qrow < S) {
          float f[HO];
#pragma unroll
          for (int k = 0; k < HO; ++k) f[k] = __uint_as_float(o0[k]) * w0 + __uint_as_float(o1[k]) * w1;
#pragma unroll
          for (int k = 0; k < HO; k += 8) {
            uint4 u;
            u.x = pack_bf16x2(f[k], f[k + 1]);
            u.y = pack_bf16x2(f[k + 2], f[k + 3]);
            u.z = pack_bf16x2(f[k + 4], f[k + 5]);
            u.w = pack_bf16x2(f[k + 6], f[k + 7]);
            *reinterpret_cast<uint4*>(orow + k) = u;
          }
        }
      }
      if (g == 0 && qrow < S) lse[((long long)b * H + h) * S + qrow] = m_all + log2f(l_tot);
    }
#ifdef VJ_ATTN_PROFILE
    if (threadIdx.x == 0) {
      atomicAdd(&g_attn_prof[0], (unsigned long long)(clock64() - ap_tot));
      atomicAdd(&g_attn_prof[1], (unsigned long long)ap1);
      atomicAdd(&g_attn_prof[2], (unsigned long long)ap2);
      atomicAdd(&g_attn_prof[3], (unsigned long long)ap3);
      atomicAdd(&g_attn_prof[4], (unsigned long long)ap4);
      atomicAdd(&g_attn_prof[5], (unsigned long long)ap5);
      atomicAdd(&g_attn_prof[6], (unsigned long long)ap6);
      atomicAdd(&g_attn_prof[9], (unsigned long long)n_my);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int HD, int POLY8>
static int launch_attn_fwd2(const void* qkv, void* out, float* lse, int B, int S, int H, cudaStream_t stream) {
  using Cfg = AttnFwd2Cfg<HD>;
  const int D = H * HD;
  TMapPair tmQ, tmKV;
  int r = make_head_tmaps<HD>(&tmQ, qkv, (uint64_t)3 * D, (uint64_t)S, (uint64_t)B, Cfg::BM);
  if (r) return r;
  r = make_head_tmaps<HD>(&tmKV, qkv, (uint64_t)3 * D, (uint64_t)S, (uint64_t)B, Cfg::BN);
  if (r) return r;
  auto kern = attn_fwd2_kernel<HD, POLY8>;
  static bool attr_set = false;
  if (!attr_set) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int n_qt = (S + Cfg::BM - 1) / Cfg::BM;
  const long long n_items = (long long)n_qt * H * B;
  VJ_CHECK(n_items < (1ll << 30), "vj_attn_fwd: too many (query tile, head, sample) work items");
  const int grid = (int)(n_items < 2ll * sm_count() ? n_items : 2ll * sm_count());   // persistent: 2 CTAs per SM
  const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmQ, tmKV, reinterpret_cast<bf16*>(out), lse, S, H, D, scale_log2,
                                                        n_qt, (int)n_items);
  VJ_LAUNCH_CHECK();
  return 0;
}

template <int HD>
static int launch_attn_fwd(const void* qkv, void* out, float* lse, int B, int S, int H, cudaStream_t stream) {
  using Cfg = AttnFwdCfg<HD>;
  const int D = H * HD;
  TMapPair tmQ, tmKV;
  int r = make_head_tmaps<HD>(&tmQ, qkv, (uint64_t)3 * D, (uint64_t)S, (uint64_t)B, Cfg::BM);
  if (r) return r;
  r = make_head_tmaps<HD>(&tmKV, qkv, (uint64_t)3 * D, (uint64_t)S, (uint64_t)B, Cfg::BN);
  if (r) return r;
  auto kern = attn_fwd_kernel<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    VJ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((S + Cfg::BM - 1) / Cfg::BM, H, B);
  const float scale_log2 = (1.0f / sqrtf((float)HD)) * 1.4426950408889634f;
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmQ, tmKV, reinterpret_cast<bf16*>(out), lse, S, H, D, scale_log2);
  VJ_LAUNCH_CHECK();
  return 0;
}

}  // namespace vj

extern "C" int vj_attn_fwd(const void* qkv, void* out, float* lse, int B, int S, int H, int head_dim, void* stream) {
  using namespace vj;
  VJ_CHECK(qkv && out && lse, "vj_attn_fwd: null pointer");
  VJ_CHECK(B > 0 && S > 0 && H > 0, "vj_attn_fwd: bad shape B=%d S=%d H=%d", B, S, H);
  VJ_CHECK(B <= 65535 && H <= 65535, "vj_attn_fwd: B/H exceed grid limits");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // share of the softmax exponentials evaluated on the FMA pipe instead of MUFU (VJ_ATTN_POLY = 0, 2, 3, 4 of 8)
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("VJ_ATTN_POLY");
    poly = (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 2;   // 2 of 8 measured best (+3 %)
  }
#define VJ_FWD2(HD_)                                                                           \
  if (head_dim == HD_) {                                                                       \
    if (poly == 0) return launch_attn_fwd2<HD_, 0>(qkv, out, lse, B, S, H, st);               \
    if (poly <= 2) return launch_attn_fwd2<HD_, 2>(qkv, out, lse, B, S, H, st);               \
    if (poly == 3) return launch_attn_fwd2<HD_, 3>(qkv, out, lse, B, S, H, st);               \
    return launch_attn_fwd2<HD_, 4>(qkv, out, lse, B, S, H, st);                              \
  }
  VJ_FWD2(64)
  VJ_FWD2(32)
#undef VJ_FWD2
  if (head_dim == 80) return launch_attn_fwd<80>(qkv, out, lse, B, S, H, st);
  set_error("vj_attn_fwd: head_dim %d not supported (32, 64, 80)", head_dim);
  return -1;
}

#ifdef VJ_ATTN_PROFILE
extern "C" int vj_attn_prof_read(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, vj::g_attn_prof, 16 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(vj::g_attn_prof, z, sizeof(z));
  }
  return 0;
}
#endif
