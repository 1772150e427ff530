// Bandwidth-class kernels of the step: LayerNorm fwd/bwd, 3-axis RoPE, apply_masks row gather /
// scatter, tubelet im2col, bias-gradient column sums, fused gather+L1 loss, predictor token ranks,
// flat EMA / AdamW / grad check.  All HBM-bound: 128-bit coalesced accesses, warp-shuffle reductions,
// deterministic two-stage reductions (no fp32 atomics except the documented scatter-add).
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_common.h"
#include "../../include/vjepa2_b200.h"

namespace vj {

// ---------------------------------------------------------------- 8-wide typed load/store
__device__ __forceinline__ void load8(const void* base, int dtype, long long elem_off, float (&v)[8]) {
  if (dtype == VJ_BF16) {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + elem_off);
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void store8(void* base, int dtype, long long elem_off, const float (&v)[8]) {
  if (dtype == VJ_BF16) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + elem_off) = u;
  } else {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__device__ __forceinline__ void rope_load8(const __half* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// ---------------------------------------------------------------- LayerNorm forward / backward
// One warp per row, rows taken in a grid-stride loop by a persistent grid.  The row stays in registers in its
// STORAGE format (bf16 rows: NV uint4 per lane) and is converted on the fly in each pass, which keeps the
// kernels at ~40-60 registers -> high occupancy, many 128-bit loads in flight per SM.
constexpr int LN_MAXV = 8;  // vectors of 8 per lane -> D <= 2048

template <bool F32> struct RawVec;
template <> struct RawVec<false> {                       // 8 bf16
  uint4 u;
  __device__ __forceinline__ void load(const void* base, long long off) {
    u = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + off);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
  }
};
template <> struct RawVec<true> {                        // 8 fp32
  float4 a, b;
  __device__ __forceinline__ void load(const void* base, long long off) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    a = p[0]; b = p[1];
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <bool F32>
__device__ __forceinline__ void store8t(void* base, long long off, const float (&v)[8]) {
  if (F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + off) = u;
  }
}

// Forward: the NEXT row of a warp is fetched before the current one is reduced (two rows of 128-bit loads in flight
// per lane), so the load latency of row i+1 hides behind the two warp reductions and the store of row i.
// ldy >= D is the row pitch of y; the pad columns [D, ldy) are written with 1.0 -- the wgrad GEMM that consumes y as
// its MN-major B operand then produces the bias gradient as one extra output column (VJ_EPI_BIAS_GRAD).
template <bool XF32, bool YF32, int NV>
__global__ void __launch_bounds__(256, 3) ln_fwd_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, void* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     long long rows, int D, long long ldy, float eps) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * 8;
  const int nvec = D >> 3;
  const int npad = (int)(ldy - D) >> 3;
  const float inv_d = 1.0f / D;
  RawVec<XF32> nxt[NV];
  if (wid < rows) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) nxt[i].load(x, wid * D + vi * 8);
    }
  }
  for (long long row = wid; row < rows; row += nw) {
    RawVec<XF32> raw[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) raw[i] = nxt[i];
    if (row + nw < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) nxt[i].load(x, (row + nw) * D + vi * 8);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
        float v[8];
        raw[i].get(v);
        s += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
        float v[8];
        raw[i].get(v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; ss = fmaf(d, d, ss); }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float v[8], o[8];
        raw[i].get(v);
        if (gamma) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (beta) {
            b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
            b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
          }
          o[0] = fmaf((v[0] - mean) * rstd, g0.x, b0.x); o[1] = fmaf((v[1] - mean) * rstd, g0.y, b0.y);
          o[2] = fmaf((v[2] - mean) * rstd, g0.z, b0.z); o[3] = fmaf((v[3] - mean) * rstd, g0.w, b0.w);
          o[4] = fmaf((v[4] - mean) * rstd, g1.x, b1.x); o[5] = fmaf((v[5] - mean) * rstd, g1.y, b1.y);
          o[6] = fmaf((v[6] - mean) * rstd, g1.z, b1.z); o[7] = fmaf((v[7] - mean) * rstd, g1.w, b1.w);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd;
        }
        store8t<YF32>(y, row * ldy + vi * 8, o);
      }
    }
    if (lane < npad) {
      const float ones[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      store8t<YF32>(y, row * ldy + D + lane * 8, ones);
    }
  }
}

// Backward, one pass over dy / x / dres:  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres), g = dy * gamma,
// AND the column sums dbeta = sum_r dy, dgamma = sum_r dy * xhat, dbias = sum_r dres (the bias gradient of the Linear
// that produced the residual branch: dres is d(fc2 output) in LN2's backward and d(proj output) in LN1's) -- no second
// read of dy and x, no separate bias-gradient launches.
// CTA = RG row groups x WPR warps.  A thread owns the SAME 8 columns of every row its group visits, so the column sums
// stay in 24 registers for the whole kernel; the two row sums go warp shuffle -> shared memory across the WPR warps
// (double buffered: one __syncthreads per iteration, R rows per group and iteration).  Per-CTA column partials go to
// scratch and are added in CTA order by ln_bwd_final_kernel (no atomics: deterministic).
constexpr int LNB_THREADS = 384;
constexpr int LNB_MAX_CTAS = 384;

// (A variant that stages the rows through a shared-memory ring filled by cp.async.bulk was measured slower -- 201 vs
// 150 us at 49152 x 1408: the kernel is close to instruction-issue-bound (~200 instructions per thread and row), and
// the extra LDS + mbarrier polling cost more than the deeper prefetch gained.)
template <bool DYF32, bool XF32, bool DXF32, int R>
__global__ void __launch_bounds__(LNB_THREADS, 2) ln_bwd_fused_kernel(
    const void* __restrict__ dy, const void* __restrict__ x, const float* __restrict__ gamma,
    const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const void* __restrict__ dres,
    void* __restrict__ dx, float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ dbias, long long rows, int D, int WPR, int RG, int iters) {
  __shared__ float s_red[2][LNB_THREADS / 32][2 * R];
  __shared__ float s_col[3][LN_MAXV * 256];
  const int t = threadIdx.x, lane = t & 31, wl = t >> 5;
  const int rg = wl / WPR, w = wl - rg * WPR;
  const int ct = w * 32 + lane;
  const int nvec = D >> 3;
  const bool live = rg < RG && ct < nvec;
  const int col = ct * 8;
  const float inv_d = 1.0f / D;
  const bool want_cols = dgamma || dbeta || dbias;
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = 1.f;
  if (live && gamma) load8(gamma, VJ_F32, col, gm);
  float acc_b[8], acc_g[8], acc_r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_b[j] = acc_g[j] = acc_r[j] = 0.f;
  int buf = 0;
  for (int it = 0; it < iters; ++it, buf ^= 1) {
    const long long base = (((long long)it * gridDim.x + blockIdx.x) * RG + rg) * R;
    RawVec<DYF32> rdy[R];
    RawVec<XF32> rx[R];
    RawVec<DXF32> rr[R];
    float mu[R], rs[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = base + u;
      const bool ok = live && row < rows;
      mu[u] = 0.f; rs[u] = 0.f;
      if (ok) {
        rdy[u].load(dy, row * D + col);
        rx[u].load(x, row * D + col);
        if (dres) rr[u].load(dres, row * D + col);
        mu[u] = __ldg(mean_in + row);
        rs[u] = __ldg(rstd_in + row);
      }
    }
    float s1[R], s2[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      s1[u] = s2[u] = 0.f;
      if (live && base + u < rows) {
        float dv[8], xv[8];
        rdy[u].get(dv);
        rx[u].get(xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = dv[j] * gm[j];
          s1[u] += g;
          s2[u] = fmaf(g, (xv[j] - mu[u]) * rs[u], s2[u]);
        }
      }
      s1[u] = warp_sum(s1[u]);
      s2[u] = warp_sum(s2[u]);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < R; ++u) { s_red[buf][wl][2 * u] = s1[u]; s_red[buf][wl][2 * u + 1] = s2[u]; }
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const long long row = base + u;
        if (row < rows) {
          float c1 = 0.f, c2 = 0.f;
          for (int k = 0; k < WPR; ++k) { c1 += s_red[buf][rg * WPR + k][2 * u]; c2 += s_red[buf][rg * WPR + k][2 * u + 1]; }
          c1 *= inv_d; c2 *= inv_d;
          float dv[8], xv[8], o[8];
          rdy[u].get(dv);
          rx[u].get(xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - mu[u]) * rs[u];
            o[j] = rs[u] * (dv[j] * gm[j] - c1 - xh * c2);
            acc_b[j] += dv[j];
            acc_g[j] = fmaf(dv[j], xh, acc_g[j]);
          }
          if (dres) {
            float r[8];
            rr[u].get(r);
#pragma unroll
            for (int j = 0; j < 8; ++j) { o[j] += r[j]; acc_r[j] += r[j]; }
          }
          store8t<DXF32>(dx, row * D + col, o);
        }
      }
    }
  }
  if (!want_cols) return;
  // ---- column sums of this CTA: row groups add into shared memory one after the other (fixed order)
  for (int g = 0; g < RG; ++g) {
    if (live && rg == g) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (g == 0) { s_col[0][col + j] = acc_b[j]; s_col[1][col + j] = acc_g[j]; s_col[2][col + j] = acc_r[j]; }
        else { s_col[0][col + j] += acc_b[j]; s_col[1][col + j] += acc_g[j]; s_col[2][col + j] += acc_r[j]; }
      }
    }
    __syncthreads();
  }
  const int n3 = 3 * D;
  float* mine = part + (long long)blockIdx.x * n3;
  for (int j = t; j < n3; j += (int)blockDim.x) mine[j] = s_col[j / D][j - (j / D) * D];
}

// Adds the per-CTA column partials of ln_bwd_fused_kernel ([P][3][D], CTA order) into dbeta / dgamma / dbias.
// block (16 columns, 16 partial groups): every thread sums P/16 partials with independent loads, the 16 group sums
// are combined in shared memory in group order -- deterministic, ~2 us.
__global__ void __launch_bounds__(256) ln_bwd_final_kernel(const float* __restrict__ part, int P, int D,
                                                           float* __restrict__ dbeta, float* __restrict__ dgamma,
                                                           float* __restrict__ dbias) {
  __shared__ float sm[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n3 = 3 * D;
  const int j = blockIdx.x * 16 + tx;
  float s = 0.f;
  if (j < n3) {
    const int per = (P + 15) / 16;
    const int k0 = ty * per, k1 = min(P, k0 + per);
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      const float a = __ldcg(part + (long long)k * n3 + j), b = __ldcg(part + (long long)(k + 1) * n3 + j);
      const float c = __ldcg(part + (long long)(k + 2) * n3 + j), d = __ldcg(part + (long long)(k + 3) * n3 + j);
      s += a; s += b; s += c; s += d;
    }
    for (; k < k1; ++k) s += __ldcg(part + (long long)k * n3 + j);
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < n3) {
    float r = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) r += sm[g][tx];
    const int q = j / D, c = j - q * D;
    float* out = q == 0 ? dbeta : q == 1 ? dgamma : dbias;
    if (out) out[c] += r;
  }
}

static inline int ln_pick_nv(long long D) {
  const long long need = (D / 8 + 31) / 32;
  return need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : need <= 6 ? 6 : 8;
}
static inline unsigned ln_grid(long long rows) {
  const long long want = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;      // 8 blocks x 8 warps per SM in flight, grid-stride beyond
  return (unsigned)(want < cap ? want : cap);
}

// ---------------------------------------------------------------- column reductions
// partial[chunk][c] = sum over the chunk's rows of f(r,c);  WITH_XHAT: also dy*xhat (LN dgamma).
// block (32, 8): thread owns 2 adjacent columns, warps stride over rows.  grid (ceil(D/64), chunks).
constexpr int COL_CHUNKS_MAX = 128;

__device__ __forceinline__ float2 load2(const void* base, int dtype, long long off) {
  if (dtype == VJ_BF16) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const bf16*>(base) + off);
    return make_float2(bf16_lo(u), bf16_hi(u));
  }
  return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + off);
}

template <bool WITH_XHAT>
__global__ void __launch_bounds__(256) colreduce_kernel(const void* __restrict__ dy, int dy_dtype,
                                                        const void* __restrict__ x, int x_dtype,
                                                        const float* __restrict__ mean,
                                                        const float* __restrict__ rstd,
                                                        float* __restrict__ part_sum,   // [chunks][D]
                                                        float* __restrict__ part_xh,    // [chunks][D]
                                                        long long rows, int D, long long rows_per_chunk) {
  __shared__ float sm[2][8][64];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * 64 + tx * 2;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  if (col < D) {
    for (long long r = r0 + ty; r < r1; r += 8) {
      const float2 d = load2(dy, dy_dtype, r * D + col);
      a0 += d.x; a1 += d.y;
      if (WITH_XHAT) {
        const float2 xv = load2(x, x_dtype, r * D + col);
        const float m = mean[r], rs = rstd[r];
        b0 += d.x * (xv.x - m) * rs;
        b1 += d.y * (xv.y - m) * rs;
      }
    }
  }
  sm[0][ty][tx * 2] = a0; sm[0][ty][tx * 2 + 1] = a1;
  if (WITH_XHAT) { sm[1][ty][tx * 2] = b0; sm[1][ty][tx * 2 + 1] = b1; }
  __syncthreads();
  const int t = ty * 32 + tx;
  if (t < 64) {
    const int c = blockIdx.x * 64 + t;
    if (c < D) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sm[0][w][t];
      part_sum[(long long)blockIdx.y * D + c] = s;
    }
  } else if (WITH_XHAT && t < 128) {
    const int tt = t - 64;
    const int c = blockIdx.x * 64 + tt;
    if (c < D) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sm[1][w][tt];
      part_xh[(long long)blockIdx.y * D + c] = s;
    }
  }
}

__global__ void colreduce_final_kernel(const float* __restrict__ part, float* __restrict__ out, int chunks, int D,
                                       int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += part[(long long)k * D + c];
  out[c] = accumulate ? out[c] + s : s;
}

static inline int col_chunks(long long rows) {
  long long c = (rows + 127) / 128;
  if (c < 1) c = 1;
  if (c > COL_CHUNKS_MAX) c = COL_CHUNKS_MAX;
  return (int)c;
}

// Wide variant (D % 8 == 0): a thread owns 8 adjacent columns (one 16-byte load per bf16 row, two per fp32 row), a
// warp covers 256 columns = 512 contiguous bytes per row, four rows in flight per thread.  The last CTA of a column
// group to finish (ticket counter, self-resetting) adds the chunk partials in chunk order and writes / accumulates
// the result, so there is no separate finalize launch and the sum order stays deterministic.
constexpr int COLW_GROUPS_MAX = 256;                   // D <= 65536
__device__ unsigned int g_colw_tickets[COLW_GROUPS_MAX];

template <bool WITH_XHAT>
__global__ void __launch_bounds__(256) colreduce_wide_kernel(const void* __restrict__ dy, int dy_dtype,
                                                             const void* __restrict__ x, int x_dtype,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd,
                                                             float* __restrict__ part_sum,   // [chunks][D]
                                                             float* __restrict__ part_xh,    // [chunks][D]
                                                             float* __restrict__ out_sum,    // [D] or null
                                                             float* __restrict__ out_xh,     // [D] or null
                                                             long long rows, int D, long long rows_per_chunk,
                                                             int accumulate) {
  __shared__ float sm[WITH_XHAT ? 2 : 1][8][256 + 8];
  __shared__ unsigned int s_ticket;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col = blockIdx.x * 256 + tx * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = b[k] = 0.f;
  if (col < D) {
    long long r = r0 + ty;
    for (; r + 24 < r1; r += 32) {                      // four rows (8 apart) in flight
      float d[4][8], xv[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(dy, dy_dtype, (r + 8 * u) * D + col, d[u]);
      if (WITH_XHAT) {
#pragma unroll
        for (int u = 0; u < 4; ++u) load8(x, x_dtype, (r + 8 * u) * D + col, xv[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (WITH_XHAT) {
          const float m = __ldg(mean + r + 8 * u), rs = __ldg(rstd + r + 8 * u);
#pragma unroll
          for (int k = 0; k < 8; ++k) { a[k] += d[u][k]; b[k] += d[u][k] * (xv[u][k] - m) * rs; }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] += d[u][k];
        }
      }
    }
    for (; r < r1; r += 8) {
      float d[8], xv[8];
      load8(dy, dy_dtype, r * D + col, d);
      if (WITH_XHAT) {
        load8(x, x_dtype, r * D + col, xv);
        const float m = __ldg(mean + r), rs = __ldg(rstd + r);
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[k] += d[k]; b[k] += d[k] * (xv[k] - m) * rs; }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += d[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sm[0][ty][tx * 8 + k] = a[k];
    if (WITH_XHAT) sm[WITH_XHAT ? 1 : 0][ty][tx * 8 + k] = b[k];
  }
  __syncthreads();
  const int t = ty * 32 + tx;
  const int c = blockIdx.x * 256 + t;
  if (c < D) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[0][w][t];
    part_sum[(long long)blockIdx.y * D + c] = s;
    if (WITH_XHAT) {
      float q = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) q += sm[WITH_XHAT ? 1 : 0][w][t];
      part_xh[(long long)blockIdx.y * D + c] = q;
    }
  }
  // ---- last CTA of this column group adds the partials: warp ty takes chunks ty, ty+8, ... (fixed order)
  __threadfence();
  __syncthreads();
  if (t == 0) s_ticket = atomicAdd(&g_colw_tickets[blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  __threadfence();
  const int chunks = (int)gridDim.y;
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = b[k] = 0.f;
  if (col < D) {
#pragma unroll 4
    for (int k = ty; k < chunks; k += 8) {
      const float4* ps = reinterpret_cast<const float4*>(part_sum + (long long)k * D + col);
      const float4 u = __ldcg(ps), v = __ldcg(ps + 1);
      a[0] += u.x; a[1] += u.y; a[2] += u.z; a[3] += u.w; a[4] += v.x; a[5] += v.y; a[6] += v.z; a[7] += v.w;
      if (WITH_XHAT) {
        const float4* px = reinterpret_cast<const float4*>(part_xh + (long long)k * D + col);
        const float4 p = __ldcg(px), q = __ldcg(px + 1);
        b[0] += p.x; b[1] += p.y; b[2] += p.z; b[3] += p.w; b[4] += q.x; b[5] += q.y; b[6] += q.z; b[7] += q.w;
      }
    }
  }
  __syncthreads();                                       // everyone is done reading sm from the first phase
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sm[0][ty][tx * 8 + k] = a[k];
    if (WITH_XHAT) sm[WITH_XHAT ? 1 : 0][ty][tx * 8 + k] = b[k];
  }
  __syncthreads();
  if (c < D) {
    if (out_sum) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sm[0][w][t];
      out_sum[c] = accumulate ? out_sum[c] + s : s;
    }
    if (WITH_XHAT && out_xh) {
      float q = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) q += sm[WITH_XHAT ? 1 : 0][w][t];
      out_xh[c] = accumulate ? out_xh[c] + q : q;
    }
  }
  if (t == 0) g_colw_tickets[blockIdx.x] = 0u;          // ready for the next launch on this stream
}

static inline bool colw_ok(int64_t D) { return D % 8 == 0 && (D + 255) / 256 <= COLW_GROUPS_MAX; }

// ---------------------------------------------------------------- RoPE
// Table layout (fp16): [row][2][hd] -- cos then sin PER ELEMENT d of the head: angle index of element d in
// its segment is (d mod seg) mod (seg/2) (the reference tiles sin/cos, modules.py:40-41); pass-through
// dims carry cos = 1, sin = 0.  With it:  out[2k]   = x[2k]  *c[2k]   - x[2k+1]*s[2k]
//                                         out[2k+1] = x[2k+1]*c[2k+1] + x[2k]  *s[2k+1]
__global__ void rope_table_kernel(const long long* __restrict__ ids, long long n, long long period, int Hp, int Wp,
                                  int hd, int seg, __half* __restrict__ table) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * hd) return;
  const long long row = t / hd;
  const int d = (int)(t - row * hd);
  float c = 1.f, s = 0.f;
  if (d < 3 * seg) {
    const int half = seg >> 1;
    const int axis = d / seg;
    const int j = (d - axis * seg) % half;
    const long long id = ids ? ids[row] : (row % period);
    const long long tpf = (long long)Hp * Wp;
    const long long f = id / tpf;
    const long long rem = id - tpf * f;
    const long long yy = rem / Wp;
    const long long xx = rem - Wp * yy;
    const double pos = (double)(axis == 0 ? f : (axis == 1 ? yy : xx));
    const double omega = 1.0 / pow(10000.0, (double)j / (double)half);
    double sd, cd;
    sincos(pos * omega, &sd, &cd);
    c = (float)cd;
    s = (float)sd;
  }
  table[row * 2 * hd + d] = __float2half_rn(c);
  table[row * 2 * hd + hd + d] = __float2half_rn(s);
}

// one thread per 8 consecutive features of the q or k part of one row
__global__ void __launch_bounds__(256) rope_apply_kernel(bf16* __restrict__ qkv, long long rows, int D, int hd,
                                                         const __half* __restrict__ table, int transpose) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vec_per_row = (2 * D) >> 3;
  if (t >= rows * vec_per_row) return;
  const long long row = t / vec_per_row;
  const int col = (int)(t - row * vec_per_row) * 8;      // in [0, 2D)
  const int d0 = (col % D) % hd;
  bf16* p = qkv + row * 3 * (long long)D + col;
  float v[8], o[8], c[8], s[8];
  load8(p, VJ_BF16, 0, v);
  rope_load8(table + row * 2 * hd + d0, c);
  rope_load8(table + row * 2 * hd + hd + d0, s);
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    if (!transpose) {
      o[k] = v[k] * c[k] - v[k + 1] * s[k];
      o[k + 1] = v[k + 1] * c[k + 1] + v[k] * s[k + 1];
    } else {
      o[k] = c[k] * v[k] + s[k + 1] * v[k + 1];
      o[k + 1] = -s[k] * v[k] + c[k + 1] * v[k + 1];
    }
  }
  store8(p, VJ_BF16, 0, o);
}

// ---------------------------------------------------------------- gather / scatter
__global__ void __launch_bounds__(256) gather_rows_kernel(const void* __restrict__ src, int src_dtype,
                                                          void* __restrict__ dst, int dst_dtype,
                                                          const long long* __restrict__ index,
                                                          const float* __restrict__ fill, long long n_out, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + warp;
  if (r >= n_out) return;
  const long long s = index[r];
  const int nvec = D >> 3;
  for (int vi = lane; vi < nvec; vi += 32) {
    float v[8];
    if (s >= 0) {
      load8(src, src_dtype, s * D + vi * 8, v);
    } else if (fill) {
      load8(fill, VJ_F32, vi * 8, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    store8(dst, dst_dtype, r * D + vi * 8, v);
  }
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const void* __restrict__ src, int src_dtype,
                                                               float* __restrict__ dst,
                                                               const long long* __restrict__ index, long long n_src,
                                                               int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + warp;
  if (r >= n_src) return;
  const long long d = index[r];
  if (d < 0) return;
  const int nvec = D >> 3;
  for (int vi = lane; vi < nvec; vi += 32) {
    float v[8];
    load8(src, src_dtype, r * D + vi * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dst + d * D + vi * 8 + j, v[j]);
  }
}

__global__ void mask_to_rows_kernel(const long long* __restrict__ masks, long long* __restrict__ out, long long B,
                                    long long K, long long N) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * K) return;
  out[t] = (t / K) * N + masks[t];
}

// ---------------------------------------------------------------- tubelet im2col
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ clips,
                                                     const long long* __restrict__ ids, bf16* __restrict__ cols,
                                                     int B, int C, int T, int H, int W, int tub, int p, long long K,
                                                     int reps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + warp;
  if (r >= (long long)B * reps * K) return;
  const int Hp = H / p, Wp = W / p;
  const long long bp = r / K;
  const long long b = bp % B;             // apply_masks(concat=True) stacks the masks along the batch
  const long long n = ids ? ids[r] : (r - bp * K);
  const int t = (int)(n / (Hp * Wp));
  const int rem = (int)(n - (long long)t * Hp * Wp);
  const int hh = rem / Wp, ww = rem - hh * Wp;
  const int p8 = p >> 3;
  const int nvec = C * tub * p * p8;
  const long long KK = (long long)C * tub * p * p;
  for (int vi = lane; vi < nvec; vi += 32) {
    const int kw8 = vi % p8;
    int q = vi / p8;
    const int kh = q % p; q /= p;
    const int kt = q % tub;
    const int c = q / tub;
    const float* src = clips + ((((long long)b * C + c) * T + (t * tub + kt)) * H + (hh * p + kh)) * W + ww * p + kw8 * 8;
    float v[8];
    load8(src, VJ_F32, 0, v);
    store8(cols, VJ_BF16, r * KK + (long long)vi * 8, v);
  }
}

// ---------------------------------------------------------------- fused gather + L1 loss (+ grad)
__global__ void __launch_bounds__(256) l1_loss_kernel(const bf16* __restrict__ z, const float* __restrict__ h,
                                                      const long long* __restrict__ idx, bf16* __restrict__ dz,
                                                      float grad_scale, const float* __restrict__ grad_scale_mul,
                                                      float* __restrict__ partial, long long B, long long K,
                                                      long long N, int D) {
  __shared__ float wsum[8];
  if (grad_scale_mul) grad_scale *= grad_scale_mul[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + warp;
  float acc = 0.f;
  if (r < B * K) {
    const long long b = r / K;
    const long long hrow = b * N + idx[r];
    const int nvec = D >> 3;
    for (int vi = lane; vi < nvec; vi += 32) {
      float zv[8], hv[8], g[8];
      load8(z, VJ_BF16, r * D + vi * 8, zv);
      load8(h, VJ_F32, hrow * D + vi * 8, hv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = zv[j] - hv[j];
        acc += fabsf(d);
        g[j] = d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f);
      }
      if (dz) store8(dz, VJ_BF16, r * D + vi * 8, g);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += wsum[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024) l1_final_kernel(const float* __restrict__ partial, long long n,
                                                        float* __restrict__ loss_accum, float loss_scale) {
  __shared__ double sm[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    loss_accum[0] += (float)(t * (double)loss_scale);
  }
}

// ---------------------------------------------------------------- predictor token ranks
__global__ void __launch_bounds__(256) argsort_rank_kernel(const long long* __restrict__ ids, int* __restrict__ rank,
                                                           long long S) {
  __shared__ long long tile[256];
  const long long b = blockIdx.y;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long* row = ids + b * S;
  const long long mine = i < S ? row[i] : 0;
  int cnt = 0;
  for (long long j0 = 0; j0 < S; j0 += 256) {
    const long long j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < S ? row[j] : 0x7fffffffffffffffll;
    __syncthreads();
    const int lim = (int)min((long long)256, S - j0);
    for (int k = 0; k < lim; ++k) {
      const long long o = tile[k];
      cnt += (o < mine) || (o == mine && (j0 + k) < i);
    }
    __syncthreads();
  }
  if (i < S) rank[b * S + i] = cnt;
}

// predictor.py:206-217,240-242 in one pass: element i of cat(masks_x, masks_y)[b] gets its stable
// ascending rank; emits every index the assemble / extract steps (and their adjoints) need.
__global__ void __launch_bounds__(256) pred_indices_kernel(const long long* __restrict__ mx,
                                                           const long long* __restrict__ my, long long Kc,
                                                           long long Kp, long long* __restrict__ ids_sorted,
                                                           long long* __restrict__ asm_idx,
                                                           long long* __restrict__ tgt_pos,
                                                           long long* __restrict__ ctx_pos,
                                                           long long* __restrict__ seq_to_tgt) {
  __shared__ long long tile[256];
  const long long S = Kc + Kp;
  const long long b = blockIdx.y;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long* rx = mx + b * Kc;
  const long long* ry = my + b * Kp;
  const long long mine = i < S ? (i < Kc ? rx[i] : ry[i - Kc]) : 0;
  int cnt = 0;
  for (long long j0 = 0; j0 < S; j0 += 256) {
    const long long j = j0 + threadIdx.x;
    tile[threadIdx.x] = j < S ? (j < Kc ? rx[j] : ry[j - Kc]) : 0x7fffffffffffffffll;
    __syncthreads();
    const int lim = (int)min((long long)256, S - j0);
    for (int k = 0; k < lim; ++k) {
      const long long o = tile[k];
      cnt += (o < mine) || (o == mine && (j0 + k) < i);
    }
    __syncthreads();
  }
  if (i >= S) return;
  const long long pos = b * S + cnt;
  ids_sorted[pos] = mine;
  if (i < Kc) {
    asm_idx[pos] = b * Kc + i;
    seq_to_tgt[pos] = -1;
    ctx_pos[b * Kc + i] = pos;
  } else {
    asm_idx[pos] = -1;
    seq_to_tgt[pos] = b * Kp + (i - Kc);
    tgt_pos[b * Kp + (i - Kc)] = pos;
  }
}

// torch.cuda.amp.GradScaler.update() on device scalars (train.py:451): backoff on inf, growth every
// `interval` clean steps; refreshes inv_scale = 1/(scale*world) and clears found_inf.
__global__ void scaler_update_kernel(float* scale, float* inv_scale, int* growth_tracker, float* found_inf,
                                     float growth, float backoff, int interval, float world) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float s = scale[0];
  if (found_inf[0] != 0.f) {
    s *= backoff;
    growth_tracker[0] = 0;
  } else {
    const int t = growth_tracker[0] + 1;
    if (t >= interval) { s *= growth; growth_tracker[0] = 0; } else growth_tracker[0] = t;
  }
  scale[0] = s;
  inv_scale[0] = 1.0f / (s * world);
  found_inf[0] = 0.f;
}

// ---------------------------------------------------------------- flat optimizer kernels
__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ tgt, const float* __restrict__ src,
                                                  bf16* __restrict__ shadow, long long n4, float m, float om) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 t = reinterpret_cast<float4*>(tgt)[i];
    const float4 s = __ldg(reinterpret_cast<const float4*>(src) + i);
    t.x = __fmaf_rn(om, s.x, __fmul_rn(t.x, m));
    t.y = __fmaf_rn(om, s.y, __fmul_rn(t.y, m));
    t.z = __fmaf_rn(om, s.z, __fmul_rn(t.z, m));
    t.w = __fmaf_rn(om, s.w, __fmul_rn(t.w, m));
    reinterpret_cast<float4*>(tgt)[i] = t;
    if (shadow) {
      uint2 u;
      u.x = pack_bf16x2(t.x, t.y);
      u.y = pack_bf16x2(t.z, t.w);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
}

__global__ void __launch_bounds__(256) grad_check_kernel(const float* __restrict__ g, long long n4,
                                                         float* __restrict__ found_inf) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) found_inf[0] = 1.0f;
}

struct AdamArgs {
  float lr_decay;      // 1 - lr*wd
  float one_minus_b1, b2, one_minus_b2;
  float step_size;     // lr / bias_c1
  float inv_sqrt_bc2;  // 1 / sqrt(bias_c2)
  float eps;
  float lr;            // for device-side bias corrections
};

// torch keeps a per-parameter step count that an inf-skipped step (GradScaler.step) does NOT advance.  The host only
// knows the number of step() calls; the number of skipped ones lives on the device: t = step - skipped.
__global__ void adam_prepare_kernel(float* bias_c, int* skipped, const float* found_inf, int step, double b1, double b2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int sk = skipped[0];
  int t = step - sk;
  if (t < 1) t = 1;
  if (found_inf && found_inf[0] != 0.f) skipped[0] = sk + 1;      // this step will be skipped by adamw_kernel
  bias_c[0] = (float)(1.0 - pow(b1, (double)t));
  bias_c[1] = (float)(1.0 - pow(b2, (double)t));
}

// one block per 4 tiles of 1024 elements; thread handles one float4 per tile
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    bf16* __restrict__ shadow, const uint8_t* __restrict__ flags,
                                                    long long ntiles, AdamArgs a, const float* __restrict__ inv_scale,
                                                    const float* __restrict__ found_inf,
                                                    const float* __restrict__ dev_bias_c) {
  if (found_inf && found_inf[0] != 0.f) return;
  const float is = inv_scale ? inv_scale[0] : 1.0f;
  if (dev_bias_c) {                                     // bias corrections of the device-side step count
    a.step_size = a.lr / __ldg(dev_bias_c);
    a.inv_sqrt_bc2 = 1.0f / sqrtf(__ldg(dev_bias_c + 1));
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const long long tile = (long long)blockIdx.x * 4 + t;
    if (tile >= ntiles) return;
    const uint8_t f = flags[tile];
    if (f & 2) continue;  // frozen: parameter never received a gradient (torch skips grad=None)
    const float decay = (f & 1) ? a.lr_decay : 1.0f;
    const long long i = tile * 256 + threadIdx.x;
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define VJ_ADAM1(c)                                                        \
    {                                                                      \
      const float gr = gg.c * is;                                          \
      pp.c *= decay;                                                       \
      mm.c = mm.c + a.one_minus_b1 * (gr - mm.c);                          \
      vv.c = vv.c * a.b2 + a.one_minus_b2 * gr * gr;                       \
      const float denom = sqrtf(vv.c) * a.inv_sqrt_bc2 + a.eps;            \
      pp.c -= a.step_size * (mm.c / denom);                                \
    }
    VJ_ADAM1(x) VJ_ADAM1(y) VJ_ADAM1(z) VJ_ADAM1(w)
#undef VJ_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      uint2 u;
      u.x = pack_bf16x2(pp.x, pp.y);
      u.y = pack_bf16x2(pp.z, pp.w);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
}

__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n4,
                                                   long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 s = __ldg(reinterpret_cast<const float4*>(src) + i);
    uint2 u;
    u.x = pack_bf16x2(s.x, s.y);
    u.y = pack_bf16x2(s.z, s.w);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  // tail (n not a multiple of 4)
  const long long tail0 = n4 * 4;
  const long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

static inline int flat_grid(long long n4) {
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace vj

using namespace vj;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

static int check_rowvec(const char* who, int64_t rows, int64_t D) {
  VJ_CHECK(rows >= 0 && D > 0, "%s: bad shape rows=%lld D=%lld", who, (long long)rows, (long long)D);
  VJ_CHECK(D % 8 == 0, "%s: D=%lld must be a multiple of 8", who, (long long)D);
  return 0;
}

template <bool XF32, bool YF32>
static void ln_fwd_launch(int nv, unsigned grid, cudaStream_t st, const void* x, const float* gamma, const float* beta,
                          void* y, float* mean, float* rstd, long long rows, int D, long long ldy, float eps) {
#define VJ_LN_CASE(NV_) \
  case NV_: ln_fwd_kernel<XF32, YF32, NV_><<<grid, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, D, ldy, eps); break;
  switch (nv) { VJ_LN_CASE(1) VJ_LN_CASE(2) VJ_LN_CASE(4) VJ_LN_CASE(6) VJ_LN_CASE(8) }
#undef VJ_LN_CASE
}

extern "C" int vj_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                                int y_dtype, float* mean, float* rstd, int64_t rows, int64_t D, int64_t ldy, float eps,
                                void* stream) {
  if (check_rowvec("vj_layernorm_fwd", rows, D)) return -1;
  VJ_CHECK(D <= LN_MAXV * 256, "vj_layernorm_fwd: D=%lld exceeds %d", (long long)D, LN_MAXV * 256);
  VJ_CHECK(x && y, "vj_layernorm_fwd: null pointer");
  if (ldy == 0) ldy = D;
  VJ_CHECK(ldy >= D && (ldy - D) % 8 == 0 && ldy - D <= 256, "vj_layernorm_fwd: bad output pitch %lld for D=%lld",
           (long long)ldy, (long long)D);
  VJ_CHECK(ldy == D || x != y, "vj_layernorm_fwd: a padded output cannot be written in place");
  if (rows == 0) return 0;
  const int nv = ln_pick_nv(D);
  const unsigned grid = ln_grid(rows);
  cudaStream_t st = STREAM(stream);
  const bool xf = x_dtype == VJ_F32, yf = y_dtype == VJ_F32;
  if (!xf && !yf) ln_fwd_launch<false, false>(nv, grid, st, x, gamma, beta, y, mean, rstd, rows, (int)D, ldy, eps);
  else if (!xf && yf) ln_fwd_launch<false, true>(nv, grid, st, x, gamma, beta, y, mean, rstd, rows, (int)D, ldy, eps);
  else if (xf && !yf) ln_fwd_launch<true, false>(nv, grid, st, x, gamma, beta, y, mean, rstd, rows, (int)D, ldy, eps);
  else ln_fwd_launch<true, true>(nv, grid, st, x, gamma, beta, y, mean, rstd, rows, (int)D, ldy, eps);
  VJ_LAUNCH_CHECK();
  return 0;
}

// launch geometry of ln_bwd_fused_kernel: WPR warps cover one row (8 columns per thread), RG rows groups per CTA
struct LnBwdGeo { int WPR, RG, R, grid, iters; };
static LnBwdGeo ln_bwd_geo(long long rows, long long D, bool any_f32) {
  LnBwdGeo g;
  g.WPR = (int)((D / 8 + 31) / 32);
  g.RG = (LNB_THREADS / 32) / g.WPR;
  if (g.RG < 1) g.RG = 1;
  g.R = any_f32 ? 1 : 2;
  const long long per_cta = (long long)g.RG * g.R;
  long long want = (rows + per_cta - 1) / per_cta;
  long long cap = 2ll * sm_count();
  if (cap > LNB_MAX_CTAS) cap = LNB_MAX_CTAS;
  g.grid = (int)(want < cap ? want : cap);
  if (g.grid < 1) g.grid = 1;
  g.iters = (int)((rows + per_cta * g.grid - 1) / (per_cta * g.grid));
  return g;
}

template <bool DYF32, bool XF32, bool DXF32>
static int ln_bwd_launch(const LnBwdGeo& g, cudaStream_t st, const void* dy, const void* x, const float* gamma,
                         const float* mean, const float* rstd, const void* dres, void* dx, float* part, float* dgamma,
                         float* dbeta, float* dbias, long long rows, int D) {
  constexpr int R = (DYF32 || XF32 || DXF32) ? 1 : 2;
  ln_bwd_fused_kernel<DYF32, XF32, DXF32, R><<<g.grid, g.WPR * g.RG * 32, 0, st>>>(
      dy, x, gamma, mean, rstd, dres, dx, part, dgamma, dbeta, dbias, rows, D, g.WPR, g.RG, g.iters);
  return 0;
}

extern "C" size_t vj_layernorm_bwd_scratch(int64_t rows, int64_t D) {
  (void)rows;
  return (size_t)LNB_MAX_CTAS * 3 * (size_t)D * sizeof(float);
}

extern "C" int vj_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* gamma,
                                const float* mean, const float* rstd, const void* dres, void* dx, int dx_dtype,
                                float* dgamma, float* dbeta, float* dbias, void* scratch, int64_t rows, int64_t D,
                                void* stream) {
  if (check_rowvec("vj_layernorm_bwd", rows, D)) return -1;
  VJ_CHECK(D <= LN_MAXV * 256, "vj_layernorm_bwd: D=%lld exceeds %d", (long long)D, LN_MAXV * 256);
  VJ_CHECK(dy && x && mean && rstd && dx, "vj_layernorm_bwd: null pointer");
  VJ_CHECK(!dbias || dres, "vj_layernorm_bwd: dbias is the column sum of dres, which is null");
  if (dgamma || dbeta || dbias) VJ_CHECK(scratch != nullptr, "vj_layernorm_bwd: scratch required for dgamma / dbeta / dbias");
  if (rows == 0) return 0;
  const bool f32 = dy_dtype == VJ_F32 || x_dtype == VJ_F32 || dx_dtype == VJ_F32;
  const LnBwdGeo g = ln_bwd_geo(rows, D, f32);
  cudaStream_t st = STREAM(stream);
  float* part = reinterpret_cast<float*>(scratch);
  const int key = (dy_dtype == VJ_F32 ? 4 : 0) | (x_dtype == VJ_F32 ? 2 : 0) | (dx_dtype == VJ_F32 ? 1 : 0);
#define VJ_LNB(A_, B_, C_) \
  if (ln_bwd_launch<A_, B_, C_>(g, st, dy, x, gamma, mean, rstd, dres, dx, part, dgamma, dbeta, dbias, rows, (int)D)) return -2
  switch (key) {   // the dtype combinations the engine produces (encoder bf16 stream, predictor fp32 stream)
    case 0: VJ_LNB(false, false, false); break;
    case 1: VJ_LNB(false, false, true); break;
    case 2: VJ_LNB(false, true, false); break;
    case 3: VJ_LNB(false, true, true); break;
    case 4: VJ_LNB(true, false, false); break;
    case 5: VJ_LNB(true, false, true); break;
    case 6: VJ_LNB(true, true, false); break;
    default: VJ_LNB(true, true, true); break;
  }
#undef VJ_LNB
  VJ_LAUNCH_CHECK();
  if (dgamma || dbeta || dbias) {
    ln_bwd_final_kernel<<<(unsigned)((3 * D + 15) / 16), 256, 0, st>>>(part, g.grid, (int)D, dbeta, dgamma, dbias);
    VJ_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" size_t vj_colsum_scratch(int64_t rows, int64_t D) { return (size_t)col_chunks(rows) * (size_t)D * sizeof(float); }

extern "C" int vj_colsum(const void* x, int x_dtype, float* out, int accumulate, void* scratch, int64_t rows,
                         int64_t D, void* stream) {
  VJ_CHECK(rows > 0 && D > 0 && D % 2 == 0, "vj_colsum: bad shape rows=%lld D=%lld", (long long)rows, (long long)D);
  VJ_CHECK(x && out && scratch, "vj_colsum: null pointer");
  const int chunks = col_chunks(rows);
  const long long rpc = (rows + chunks - 1) / chunks;
  float* ps = reinterpret_cast<float*>(scratch);
  if (colw_ok(D)) {
    dim3 grid((unsigned)((D + 255) / 256), (unsigned)chunks), block(32, 8);
    colreduce_wide_kernel<false><<<grid, block, 0, STREAM(stream)>>>(x, x_dtype, nullptr, 0, nullptr, nullptr, ps, nullptr,
                                                                    out, nullptr, rows, (int)D, rpc, accumulate);
    VJ_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((unsigned)((D + 63) / 64), (unsigned)chunks), block(32, 8);
  colreduce_kernel<false><<<grid, block, 0, STREAM(stream)>>>(x, x_dtype, nullptr, 0, nullptr, nullptr, ps, nullptr,
                                                             rows, (int)D, rpc);
  VJ_LAUNCH_CHECK();
  colreduce_final_kernel<<<(unsigned)((D + 255) / 256), 256, 0, STREAM(stream)>>>(ps, out, chunks, (int)D, accumulate);
  VJ_LAUNCH_CHECK();
  return 0;
}

static int rope_seg(int head_dim) { return 2 * ((head_dim / 3) / 2); }

extern "C" int vj_rope_table(const int64_t* ids, int64_t n, int64_t period, int Hp, int Wp, int head_dim, void* table,
                             void* stream) {
  VJ_CHECK(table && n > 0 && Hp > 0 && Wp > 0, "vj_rope_table: bad arguments");
  VJ_CHECK(ids != nullptr || period > 0, "vj_rope_table: ids == NULL needs period > 0");
  VJ_CHECK(head_dim % 8 == 0 && rope_seg(head_dim) > 0, "vj_rope_table: head_dim %d unsupported", head_dim);
  const long long total = (long long)n * head_dim;
  rope_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const long long*>(ids), n, period, Hp, Wp, head_dim, rope_seg(head_dim),
      reinterpret_cast<__half*>(table));
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_rope_apply(void* qkv, int64_t rows, int64_t D, int heads, int head_dim, const void* table,
                             int transpose, void* stream) {
  VJ_CHECK(qkv && table && rows > 0, "vj_rope_apply: bad arguments");
  VJ_CHECK((int64_t)heads * head_dim == D && head_dim % 8 == 0, "vj_rope_apply: D=%lld != heads*head_dim (%d*%d)",
           (long long)D, heads, head_dim);
  const long long total = rows * ((2 * D) >> 3);
  rope_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<bf16*>(qkv), rows, (int)D, head_dim, reinterpret_cast<const __half*>(table), transpose);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_gather_rows(const void* src, int src_dtype, void* dst, int dst_dtype, const int64_t* index,
                              const float* fill, int64_t n_out, int64_t D, void* stream) {
  if (check_rowvec("vj_gather_rows", n_out, D)) return -1;
  VJ_CHECK(dst && index, "vj_gather_rows: null pointer");
  if (n_out == 0) return 0;
  gather_rows_kernel<<<(unsigned)((n_out + 7) / 8), 256, 0, STREAM(stream)>>>(
      src, src_dtype, dst, dst_dtype, reinterpret_cast<const long long*>(index), fill, n_out, (int)D);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_scatter_add_rows(const void* src, int src_dtype, float* dst, const int64_t* index, int64_t n_src,
                                   int64_t D, void* stream) {
  if (check_rowvec("vj_scatter_add_rows", n_src, D)) return -1;
  VJ_CHECK(src && dst && index, "vj_scatter_add_rows: null pointer");
  if (n_src == 0) return 0;
  scatter_add_rows_kernel<<<(unsigned)((n_src + 7) / 8), 256, 0, STREAM(stream)>>>(
      src, src_dtype, dst, reinterpret_cast<const long long*>(index), n_src, (int)D);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_mask_to_rows(const int64_t* masks, int64_t* out, int64_t B, int64_t K, int64_t N, void* stream) {
  VJ_CHECK(masks && out && B > 0 && K > 0, "vj_mask_to_rows: bad arguments");
  mask_to_rows_kernel<<<(unsigned)((B * K + 255) / 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const long long*>(masks), reinterpret_cast<long long*>(out), B, K, N);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_im2col_tubelets(const float* clips, const int64_t* ids, void* cols, int B, int C, int T, int H,
                                  int W, int tubelet, int patch, int64_t K, int reps, void* stream) {
  VJ_CHECK(clips && cols && B > 0 && C > 0, "vj_im2col_tubelets: bad arguments");
  VJ_CHECK(patch % 8 == 0 && T % tubelet == 0 && H % patch == 0 && W % patch == 0 && W % 4 == 0,
           "vj_im2col_tubelets: geometry T=%d H=%d W=%d tubelet=%d patch=%d unsupported", T, H, W, tubelet, patch);
  if (!ids) { K = (int64_t)(T / tubelet) * (H / patch) * (W / patch); reps = 1; }
  VJ_CHECK(reps >= 1, "vj_im2col_tubelets: reps must be >= 1");
  im2col_kernel<<<(unsigned)(((int64_t)B * reps * K + 7) / 8), 256, 0, STREAM(stream)>>>(
      clips, reinterpret_cast<const long long*>(ids), reinterpret_cast<bf16*>(cols), B, C, T, H, W, tubelet, patch, K,
      reps);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t vj_l1_scratch(int64_t B, int64_t K, int64_t D) {
  (void)D;
  return (size_t)((B * K + 7) / 8) * sizeof(float);
}

extern "C" int vj_l1_loss(const void* z, const float* h, const int64_t* idx, float* loss_accum, void* dz,
                          float loss_scale, float grad_scale, const float* grad_scale_mul, void* scratch, int64_t B,
                          int64_t K, int64_t N, int64_t D, void* stream) {
  if (check_rowvec("vj_l1_loss", B * K, D)) return -1;
  VJ_CHECK(z && h && idx && loss_accum && scratch && B * K > 0, "vj_l1_loss: null pointer / empty");
  const long long blocks = (B * K + 7) / 8;
  l1_loss_kernel<<<(unsigned)blocks, 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const bf16*>(z), h, reinterpret_cast<const long long*>(idx), reinterpret_cast<bf16*>(dz),
      grad_scale, grad_scale_mul, reinterpret_cast<float*>(scratch), B, K, N, (int)D);
  VJ_LAUNCH_CHECK();
  l1_final_kernel<<<1, 1024, 0, STREAM(stream)>>>(reinterpret_cast<const float*>(scratch), blocks, loss_accum, loss_scale);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_argsort_rank(const int64_t* ids, int32_t* rank, int64_t B, int64_t S, void* stream) {
  VJ_CHECK(ids && rank && B > 0 && S > 0, "vj_argsort_rank: bad arguments");
  dim3 grid((unsigned)((S + 255) / 256), (unsigned)B);
  argsort_rank_kernel<<<grid, 256, 0, STREAM(stream)>>>(reinterpret_cast<const long long*>(ids), rank, S);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_ema_update(float* tgt, const float* src, void* tgt_bf16, int64_t n, float m, float one_minus_m,
                             void* stream) {
  VJ_CHECK(tgt && src && n > 0 && n % 4 == 0, "vj_ema_update: n=%lld must be a positive multiple of 4", (long long)n);
  ema_kernel<<<flat_grid(n / 4), 256, 0, STREAM(stream)>>>(tgt, src, reinterpret_cast<bf16*>(tgt_bf16), n / 4, m,
                                                         one_minus_m);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_grad_check(const float* g, int64_t n, float* found_inf, void* stream) {
  VJ_CHECK(g && found_inf && n > 0 && n % 4 == 0, "vj_grad_check: n=%lld must be a positive multiple of 4", (long long)n);
  grad_check_kernel<<<flat_grid(n / 4), 256, 0, STREAM(stream)>>>(g, n / 4, found_inf);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_adamw_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, void* p_bf16,
                             const uint8_t* tile_flags, int64_t n, float lr, float beta1, float beta2, float eps,
                             float wd, float bias_c1, float bias_c2, const float* dev_bias_c, const float* inv_scale,
                             const float* found_inf, void* stream) {
  VJ_CHECK(p && g && exp_avg && exp_avg_sq && tile_flags, "vj_adamw_step: null pointer");
  VJ_CHECK(n > 0 && n % 1024 == 0, "vj_adamw_step: n=%lld must be a positive multiple of 1024", (long long)n);
  AdamArgs a;
  a.lr_decay = (float)(1.0 - (double)lr * (double)wd);
  a.one_minus_b1 = (float)(1.0 - (double)beta1);
  a.b2 = beta2;
  a.one_minus_b2 = (float)(1.0 - (double)beta2);
  a.step_size = (float)((double)lr / (double)bias_c1);
  a.inv_sqrt_bc2 = (float)(1.0 / sqrt((double)bias_c2));
  a.eps = eps;
  a.lr = lr;
  const long long ntiles = n / 1024;
  adamw_kernel<<<(unsigned)((ntiles + 3) / 4), 256, 0, STREAM(stream)>>>(p, g, exp_avg, exp_avg_sq,
                                                                        reinterpret_cast<bf16*>(p_bf16), tile_flags,
                                                                        ntiles, a, inv_scale, found_inf, dev_bias_c);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_adam_prepare(float* bias_c, int32_t* skipped, const float* found_inf, int step, double beta1,
                               double beta2, void* stream) {
  VJ_CHECK(bias_c && skipped, "vj_adam_prepare: null pointer");
  VJ_CHECK(step >= 1, "vj_adam_prepare: step=%d must be >= 1", step);
  adam_prepare_kernel<<<1, 32, 0, STREAM(stream)>>>(bias_c, skipped, found_inf, step, beta1, beta2);
  VJ_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  const long long n4 = n >> 2;
  const float4 v4 = make_float4(v, v, v, v);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) reinterpret_cast<float4*>(p)[i] = v4;
  const long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

extern "C" int vj_fill_f32(float* p, int64_t n, float value, void* stream) {
  VJ_CHECK(p != nullptr && n >= 0, "vj_fill_f32: bad arguments");
  VJ_CHECK((reinterpret_cast<uintptr_t>(p) & 15) == 0, "vj_fill_f32: pointer must be 16-byte aligned");
  if (n == 0) return 0;
  fill_f32_kernel<<<flat_grid(n / 4 + 1), 256, 0, STREAM(stream)>>>(p, n, value);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  VJ_CHECK(src && dst && n > 0, "vj_cast_f32_bf16: bad arguments");
  cast_kernel<<<flat_grid(n / 4), 256, 0, STREAM(stream)>>>(src, reinterpret_cast<bf16*>(dst), n / 4, n);
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_pred_indices(const int64_t* masks_x, const int64_t* masks_y, int64_t B, int64_t Kc, int64_t Kp,
                               int64_t* ids_sorted, int64_t* asm_idx, int64_t* tgt_pos, int64_t* ctx_pos,
                               int64_t* seq_to_tgt, void* stream) {
  VJ_CHECK(masks_x && masks_y && ids_sorted && asm_idx && tgt_pos && ctx_pos && seq_to_tgt, "vj_pred_indices: null pointer");
  VJ_CHECK(B > 0 && Kc > 0 && Kp > 0 && B <= 65535, "vj_pred_indices: bad shape B=%lld Kc=%lld Kp=%lld", (long long)B,
           (long long)Kc, (long long)Kp);
  dim3 grid((unsigned)((Kc + Kp + 255) / 256), (unsigned)B);
  pred_indices_kernel<<<grid, 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const long long*>(masks_x), reinterpret_cast<const long long*>(masks_y), Kc, Kp,
      reinterpret_cast<long long*>(ids_sorted), reinterpret_cast<long long*>(asm_idx),
      reinterpret_cast<long long*>(tgt_pos), reinterpret_cast<long long*>(ctx_pos),
      reinterpret_cast<long long*>(seq_to_tgt));
  VJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int vj_scaler_update(float* scale, float* inv_scale, int32_t* growth_tracker, float* found_inf,
                                float growth, float backoff, int interval, float world, void* stream) {
  VJ_CHECK(scale && inv_scale && growth_tracker && found_inf, "vj_scaler_update: null pointer");
  scaler_update_kernel<<<1, 32, 0, STREAM(stream)>>>(scale, inv_scale, growth_tracker, found_inf, growth, backoff,
                                                    interval, world);
  VJ_LAUNCH_CHECK();
  return 0;
}
