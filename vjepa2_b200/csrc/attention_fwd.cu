// Flash-attention forward on tcgen05 for sm_100a (non-causal, no mask, scale 1/sqrt(d)).
// TMEM (256 columns): S ping-pong [0,64) [64,128), O at [128,128+d).  ~83 KB smem -> 2 CTAs / SM so one CTA's
// exp phase overlaps the other's MMAs.  Q/K/V tiles are read straight out of the fused qkv activation
// [B*S][3D] with a 3-D tensor map (d, S, B) -- no head-major relayout pass.  Roles: see attn_fwd_kernel.
#include "common.cuh"
#include "host_common.h"
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
  static constexpr int BM = 128, BN = 64, KV_STAGES = 3;
  static constexpr int Q_BYTES = BM * HD * 2;
  static constexpr int KV_BYTES = BN * HD * 2;
  static constexpr int P_BYTES = BM * BN * 2;        // 16 KB, 128-B rows
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + KV_STAGES * KV_BYTES;
  static constexpr int OFF_X = OFF_P + P_BYTES;      // row-max / row-sum exchange: [2 parity][2 half][128] floats
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
  uint8_t* sP = smem + Cfg::OFF_P;
  float* xch = reinterpret_cast<float*>(smem + Cfg::OFF_X);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars;           // 1
  uint64_t* k_full = bars + 1;       // 3
  uint64_t* k_empty = bars + 4;      // 3
  uint64_t* v_full = bars + 7;       // 3
  uint64_t* v_empty = bars + 10;     // 3
  uint64_t* s_full = bars + 13;      // 2
  uint64_t* p_full = bars + 15;      // 1 (8 arrivals: one per softmax warp)
  uint64_t* pv_done = bars + 16;     // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * Cfg::BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (S + Cfg::BN - 1) / Cfg::BN;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(p_full, 8);                 // one arrival per softmax warp
    mbar_init(pv_done, 1);
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
        const int st = j % 3;
        const uint32_t ph = (j / 3) & 1;
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
    const uint64_t pd = desc_kmajor<128>(smem_u32(sP));
    mbar_wait(q_full, 0);
    long long ap_p = 0, ap_kv = 0;
    (void)ap_p; (void)ap_kv;
    AP_T0(ap_all);
    auto issue_qk = [&](int j) {
      const int st = j % 3;
      AP_T0(a0);
      mbar_wait(&k_full[st], (j / 3) & 1);
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
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int st = j % 3;
      AP_T0(a1);
      mbar_wait(&v_full[st], (j / 3) & 1);
      AP_ADD(ap_kv, a1);
      AP_T0(a2);
      mbar_wait(p_full, j & 1);
      AP_ADD(ap_p, a2);
      tc_fence_after();
      if (elect_one()) {
        mma_into_hd<HD, false, Cfg::BN>(tmem_base + Cfg::O_COL, [&](int k) { return desc_advance(pd, k * 32); },
                                        smem_u32(sV + st * Cfg::KV_BYTES), j != 0);
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
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
    uint8_t* prow = sP + r * 128;
    long long ap1 = 0, ap2 = 0, ap3 = 0, ap4 = 0, ap5 = 0, ap6 = 0;
    (void)ap1; (void)ap2; (void)ap3; (void)ap4; (void)ap5; (void)ap6;
    AP_T0(ap_tot);
    for (int j = 0; j < n_tiles; ++j) {
      AP_T0(b1);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      AP_ADD(ap1, b1);
      tc_fence_after();
      uint32_t sr[32];
      AP_T0(b2);
      tmem_ld32(lane_addr + (j & 1) * Cfg::BN + half * 32, sr);
      tmem_ld_wait();
      AP_ADD(ap2, b2);
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
      float* xm = xch + (j & 1) * 256;
      xm[half * 128 + r] = mloc;
      pair_barrier(q);
      const float m_tile = fmaxf(mloc, xm[(half ^ 1) * 128 + r]) * scale_log2;
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
      if (j > 0) {
        // PV_{j-1} finished: P buffer is free and O holds tiles < j
        AP_T0(b4);
        mbar_wait(pv_done, (j - 1) & 1);
        AP_ADD(ap4, b4);
        AP_T0(b5);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
          uint32_t o[HO];
          tmem_ld_n<HO>(lane_addr + Cfg::O_COL + half * HO, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < HO; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_n<HO>(lane_addr + Cfg::O_COL + half * HO, o);
          tmem_st_wait();
        }
        AP_ADD(ap5, b5);
      }
      AP_T0(b6);
      // P half-row -> smem, K-major 128-B swizzled rows: 16-B chunk c lands at c ^ (r & 7)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        *reinterpret_cast<uint4*>(prow + (((half * 4 + c) ^ (r & 7)) << 4)) = u;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      AP_ADD(ap6, b6);
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
    float* xl = xch + (n_tiles & 1) * 256;
    xl[half * 128 + r] = l_run;
    pair_barrier(q);
    const float l_tot = l_run + xl[(half ^ 1) * 128 + r];
    mbar_wait(pv_done, (n_tiles - 1) & 1);
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
  if (head_dim == 64) return launch_attn_fwd<64>(qkv, out, lse, B, S, H, st);
  if (head_dim == 32) return launch_attn_fwd<32>(qkv, out, lse, B, S, H, st);
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
