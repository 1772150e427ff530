// Stand-alone device self-test + micro-benchmark for the tcgen05 kernels (no Python / torch).
// Checks vj_gemm (all operand-major modes, epilogues, ragged shapes) and vj_attn_fwd / vj_attn_bwd against
// naive CUDA-core kernels, then times the ViT-g shapes.  Build: make selftest.  Run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/vjepa2_b200.h"

typedef __nv_bfloat16 bf16;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)
#define VJ(x)                                                        \
  do {                                                               \
    int r = (x);                                                     \
    if (r != 0) {                                                    \
      printf("vj error %d: %s (%s:%d)\n", r, vj_last_error(), __FILE__, __LINE__); \
      exit(3);                                                       \
    }                                                                \
  } while (0)

static int g_fail = 0;

__global__ void fill_kernel(bf16* p, long long n, unsigned seed, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = __float2bfloat16(((x & 0xFFFF) / 65535.0f - 0.5f) * 2.0f * scale);
  }
}
__global__ void fill_f32_kernel(float* p, long long n, unsigned seed, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u ^ seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = ((x & 0xFFFF) / 65535.0f - 0.5f) * 2.0f * scale;
  }
}

// naive reference: out[m,n] = epi(sum_k A(m,k) B(n,k))
__global__ void ref_gemm_kernel(const bf16* A, const bf16* B, float* out, int M, int N, int K, long long lda,
                                long long ldb, int a_mn, int b_mn) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = __bfloat162float(a_mn ? A[(long long)k * lda + m] : A[(long long)m * lda + k]);
    const float b = __bfloat162float(b_mn ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k]);
    acc += a * b;
  }
  out[(long long)m * N + n] = acc;
}

static float bf16_round_h(float v) { return __bfloat162float(__float2bfloat16(v)); }
static float gelu_h(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678f)); }
static float dgelu_h(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678f)) + x * 0.3989422804f * expf(-0.5f * x * x);
}

template <typename T>
static T* dalloc(size_t n) {
  T* p;
  CK(cudaMalloc(&p, n * sizeof(T)));
  return p;
}
static void fill(bf16* p, long long n, unsigned seed, float scale) { fill_kernel<<<512, 256>>>(p, n, seed, scale); }
static void fillf(float* p, long long n, unsigned seed, float scale) { fill_f32_kernel<<<512, 256>>>(p, n, seed, scale); }

static void test_gemm(const char* name, int M, int N, int K, int a_mn, int b_mn, int flags) {
  const long long lda = a_mn ? M : K, ldb = b_mn ? N : K;
  bf16* A = dalloc<bf16>((size_t)M * K);
  bf16* B = dalloc<bf16>((size_t)N * K);
  float* bias = dalloc<float>(N);
  bf16* aux_in = dalloc<bf16>((size_t)M * N);
  bf16* aux_out = dalloc<bf16>((size_t)M * N);
  const bool of32 = flags & VJ_EPI_OUT_F32, rf32 = flags & VJ_EPI_RES_F32;
  void* out = of32 ? (void*)dalloc<float>((size_t)M * N) : (void*)dalloc<bf16>((size_t)M * N);
  void* res = rf32 ? (void*)dalloc<float>((size_t)M * N) : (void*)dalloc<bf16>((size_t)M * N);
  float* ref = dalloc<float>((size_t)M * N);
  fill(A, (long long)M * K, 1, 1.0f);
  fill(B, (long long)N * K, 2, 1.0f);
  fillf(bias, N, 3, 1.0f);
  fill(aux_in, (long long)M * N, 4, 2.0f);
  if (rf32) fillf((float*)res, (long long)M * N, 5, 1.0f); else fill((bf16*)res, (long long)M * N, 5, 1.0f);
  CK(cudaMemset(out, 0xFF, (size_t)M * N * (of32 ? 4 : 2)));
  // in-place accumulate variant: residual aliases out
  const bool alias = (flags & VJ_EPI_RESIDUAL) && of32 && rf32 && !(flags & (VJ_EPI_ROUND_BF16 | VJ_EPI_GELU | VJ_EPI_DGELU));
  if (alias) CK(cudaMemcpy(out, res, (size_t)M * N * 4, cudaMemcpyDeviceToDevice));
  vj_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = A; g.b = B; g.out = out; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldo = N;
  g.a_mn_major = a_mn; g.b_mn_major = b_mn; g.flags = flags; g.bias = bias;
  g.residual = alias ? out : res; g.ldr = N; g.aux_out = aux_out; g.aux_in = aux_in; g.ld_aux = N;
  VJ(vj_gemm(&g, 0));
  dim3 grid((N + 127) / 128, M);
  ref_gemm_kernel<<<grid, 128>>>(A, B, ref, M, N, K, lda, ldb, a_mn, b_mn);
  CK(cudaDeviceSynchronize());
  std::vector<float> hr((size_t)M * N), hb(N), hres((size_t)M * N), ho((size_t)M * N);
  std::vector<bf16> hauxi((size_t)M * N), hauxo((size_t)M * N);
  CK(cudaMemcpy(hr.data(), ref, hr.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), bias, N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hauxi.data(), aux_in, hauxi.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hauxo.data(), aux_out, hauxo.size() * 2, cudaMemcpyDeviceToHost));
  if (rf32) CK(cudaMemcpy(hres.data(), res, hres.size() * 4, cudaMemcpyDeviceToHost));
  else {
    std::vector<bf16> t((size_t)M * N);
    CK(cudaMemcpy(t.data(), res, t.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < t.size(); ++i) hres[i] = __bfloat162float(t[i]);
  }
  if (of32) CK(cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost));
  else {
    std::vector<bf16> t((size_t)M * N);
    CK(cudaMemcpy(t.data(), out, t.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < t.size(); ++i) ho[i] = __bfloat162float(t[i]);
  }
  double max_err = 0, max_aux = 0;
  long long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      const size_t i = (size_t)m * N + n;
      float v = hr[i];
      if (flags & VJ_EPI_BIAS) v += hb[n];
      const float pre = v;
      if (flags & VJ_EPI_ROUND_BF16) v = bf16_round_h(v);
      if (flags & VJ_EPI_GELU) v = gelu_h(v);
      if (flags & VJ_EPI_DGELU) v *= dgelu_h(__bfloat162float(hauxi[i]));
      if (flags & VJ_EPI_RESIDUAL) v += hres[i];
      const double tol = 2e-2 * fabs(v) + 2e-2 * sqrt((double)K) * 0.02 + (of32 ? 1e-3 : 0);
      const double err = fabs((double)ho[i] - v);
      if (!(err <= tol)) ++bad;
      if (err > max_err) max_err = err;
      if (flags & VJ_EPI_AUX_OUT) {
        const double ea = fabs((double)__bfloat162float(hauxo[i]) - pre);
        if (ea > max_aux) max_aux = ea;
        if (!(ea <= 1e-2 * fabs(pre) + 1e-2)) ++bad;
      }
    }
  printf("[gemm] %-34s M=%d N=%d K=%d amn=%d bmn=%d flags=%3d  max_err=%.4g aux_err=%.4g bad=%lld  %s\n", name, M, N,
         K, a_mn, b_mn, flags, max_err, max_aux, bad, bad ? "FAIL" : "ok");
  if (bad) g_fail = 1;
  cudaFree(A); cudaFree(B); cudaFree(bias); cudaFree(aux_in); cudaFree(aux_out); cudaFree(out); cudaFree(res); cudaFree(ref);
}

// ------------------------------------------------------------------ attention reference (one thread per q row)
__global__ void ref_attn_fwd_kernel(const bf16* qkv, float* out, float* lse2, int B, int S, int H, int hd) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  if (q >= S) return;
  const int D = H * hd;
  const bf16* Q = qkv + ((long long)b * S + q) * 3 * D + h * hd;
  const float scale = rsqrtf((float)hd);
  float m = -INFINITY, l = 0.f, acc[128];
  for (int i = 0; i < hd; ++i) acc[i] = 0.f;
  for (int k = 0; k < S; ++k) {
    const bf16* Kp = qkv + ((long long)b * S + k) * 3 * D + D + h * hd;
    const bf16* Vp = Kp + D;
    float s = 0.f;
    for (int i = 0; i < hd; ++i) s += __bfloat162float(Q[i]) * __bfloat162float(Kp[i]);
    s *= scale;
    const float mn = fmaxf(m, s);
    const float a = expf(m - mn), p = expf(s - mn);
    l = l * a + p;
    for (int i = 0; i < hd; ++i) acc[i] = acc[i] * a + p * __bfloat162float(Vp[i]);
    m = mn;
  }
  float* o = out + ((long long)b * S + q) * D + h * hd;
  for (int i = 0; i < hd; ++i) o[i] = acc[i] / l;
  lse2[((long long)b * H + h) * S + q] = (m + logf(l)) * 1.4426950408889634f;
}

// reference backward: one thread per (b,h,row) for dq, and per (b,h,key) for dk/dv (O(S^2 d) each)
__global__ void ref_attn_bwd_kernel(const bf16* qkv, const float* o_ref, const bf16* dout, const float* lse2,
                                    float* dqkv, int B, int S, int H, int hd) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  if (t >= S) return;
  const int D = H * hd;
  const float scale = rsqrtf((float)hd);
  const float L2E = 1.4426950408889634f;
  // ---- dq for query row t
  {
    const bf16* Q = qkv + ((long long)b * S + t) * 3 * D + h * hd;
    const bf16* dO = dout + ((long long)b * S + t) * D + h * hd;
    const float* O = o_ref + ((long long)b * S + t) * D + h * hd;
    float delta = 0.f;
    for (int i = 0; i < hd; ++i) delta += __bfloat162float(dO[i]) * O[i];
    const float l2 = lse2[((long long)b * H + h) * S + t];
    float dq[128];
    for (int i = 0; i < hd; ++i) dq[i] = 0.f;
    for (int k = 0; k < S; ++k) {
      const bf16* Kp = qkv + ((long long)b * S + k) * 3 * D + D + h * hd;
      const bf16* Vp = Kp + D;
      float s = 0.f, dp = 0.f;
      for (int i = 0; i < hd; ++i) {
        s += __bfloat162float(Q[i]) * __bfloat162float(Kp[i]);
        dp += __bfloat162float(dO[i]) * __bfloat162float(Vp[i]);
      }
      const float p = exp2f(s * scale * L2E - l2);
      const float ds = p * (dp - delta) * scale;
      for (int i = 0; i < hd; ++i) dq[i] += ds * __bfloat162float(Kp[i]);
    }
    float* o = dqkv + ((long long)b * S + t) * 3 * D + h * hd;
    for (int i = 0; i < hd; ++i) o[i] = dq[i];
  }
  // ---- dk, dv for key row t
  {
    const bf16* Kp = qkv + ((long long)b * S + t) * 3 * D + D + h * hd;
    const bf16* Vp = Kp + D;
    float dk[128], dv[128];
    for (int i = 0; i < hd; ++i) dk[i] = dv[i] = 0.f;
    for (int q = 0; q < S; ++q) {
      const bf16* Q = qkv + ((long long)b * S + q) * 3 * D + h * hd;
      const bf16* dO = dout + ((long long)b * S + q) * D + h * hd;
      const float* O = o_ref + ((long long)b * S + q) * D + h * hd;
      float s = 0.f, dp = 0.f, delta = 0.f;
      for (int i = 0; i < hd; ++i) {
        s += __bfloat162float(Q[i]) * __bfloat162float(Kp[i]);
        dp += __bfloat162float(dO[i]) * __bfloat162float(Vp[i]);
        delta += __bfloat162float(dO[i]) * O[i];
      }
      const float p = exp2f(s * scale * L2E - lse2[((long long)b * H + h) * S + q]);
      const float ds = p * (dp - delta) * scale;
      for (int i = 0; i < hd; ++i) {
        dk[i] += ds * __bfloat162float(Q[i]);
        dv[i] += p * __bfloat162float(dO[i]);
      }
    }
    float* o = dqkv + ((long long)b * S + t) * 3 * D + D + h * hd;
    for (int i = 0; i < hd; ++i) { o[i] = dk[i]; o[D + i] = dv[i]; }
  }
}

static void compare(const char* what, const std::vector<float>& got, const std::vector<float>& ref, double rtol,
                    double atol) {
  double max_err = 0, max_ref = 0;
  long long bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double e = fabs((double)got[i] - ref[i]);
    if (e > max_err) max_err = e;
    if (fabs(ref[i]) > max_ref) max_ref = fabs(ref[i]);
    if (!(e <= atol + rtol * fabs(ref[i]))) ++bad;
  }
  printf("        %-10s max_err=%.4g (max |ref|=%.4g) bad=%lld/%zu  %s\n", what, max_err, max_ref, bad, ref.size(),
         bad ? "FAIL" : "ok");
  if (bad) g_fail = 1;
}

static std::vector<float> d2h_bf16(const bf16* p, size_t n) {
  std::vector<bf16> t(n);
  CK(cudaMemcpy(t.data(), p, n * 2, cudaMemcpyDeviceToHost));
  std::vector<float> f(n);
  for (size_t i = 0; i < n; ++i) f[i] = __bfloat162float(t[i]);
  return f;
}
static std::vector<float> d2h_f32(const float* p, size_t n) {
  std::vector<float> f(n);
  CK(cudaMemcpy(f.data(), p, n * 4, cudaMemcpyDeviceToHost));
  return f;
}

static void test_attn(int B, int S, int H, int hd, bool bwd) {
  const int D = H * hd;
  const size_t nq = (size_t)B * S * 3 * D, no = (size_t)B * S * D, nl = (size_t)B * H * S;
  bf16* qkv = dalloc<bf16>(nq);
  bf16* out = dalloc<bf16>(no);
  float* lse = dalloc<float>(nl);
  float* oref = dalloc<float>(no);
  float* lref = dalloc<float>(nl);
  fill(qkv, nq, 7, 2.0f);
  CK(cudaMemset(out, 0xFF, no * 2));
  VJ(vj_attn_fwd(qkv, out, lse, B, S, H, hd, 0));
  dim3 grid((S + 63) / 64, H, B);
  ref_attn_fwd_kernel<<<grid, 64>>>(qkv, oref, lref, B, S, H, hd);
  CK(cudaDeviceSynchronize());
  printf("[attn] B=%d S=%d H=%d hd=%d\n", B, S, H, hd);
  compare("out", d2h_bf16(out, no), d2h_f32(oref, no), 2e-2, 1e-2);
  compare("lse", d2h_f32(lse, nl), d2h_f32(lref, nl), 1e-3, 1e-2);
  if (bwd) {
    bf16* dout = dalloc<bf16>(no);
    bf16* dqkv = dalloc<bf16>(nq);
    float* dref = dalloc<float>(nq);
    fill(dout, no, 9, 1.0f);
    const size_t sb = vj_attn_bwd_scratch(B, S, H, hd);
    void* scratch = dalloc<char>(sb);
    CK(cudaMemset(dqkv, 0xFF, nq * 2));
    VJ(vj_attn_bwd(qkv, out, dout, lse, dqkv, scratch, nullptr, B, S, H, hd, 0));
    ref_attn_bwd_kernel<<<grid, 64>>>(qkv, oref, dout, lref, dref, B, S, H, hd);
    CK(cudaDeviceSynchronize());
    std::vector<float> got = d2h_bf16(dqkv, nq), ref = d2h_f32(dref, nq);
    // split per q/k/v for readable reporting
    std::vector<float> g3[3], r3[3];
    for (size_t row = 0; row < (size_t)B * S; ++row)
      for (int w = 0; w < 3; ++w)
        for (int c = 0; c < D; ++c) {
          g3[w].push_back(got[row * 3 * D + w * D + c]);
          r3[w].push_back(ref[row * 3 * D + w * D + c]);
        }
    compare("dq", g3[0], r3[0], 3e-2, 3e-2);
    compare("dk", g3[1], r3[1], 3e-2, 3e-2);
    compare("dv", g3[2], r3[2], 3e-2, 3e-2);
    cudaFree(dout); cudaFree(dqkv); cudaFree(dref); cudaFree(scratch);
  }
  cudaFree(qkv); cudaFree(out); cudaFree(lse); cudaFree(oref); cudaFree(lref);
}

// ------------------------------------------------------------------ GEMM stress: CTA-pair kernel, back-to-back launches
__global__ void diff_words_kernel(const unsigned* a, const unsigned* b, long long n, unsigned long long* bad) {
  unsigned long long c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += a[i] != b[i];
  if (c) atomicAdd(bad, c);
}

// The pair kernel's cross-CTA handshakes (remote accumulator-release arrives, multicast commits) under the
// conditions a training step creates: consecutive launches whose CTAs overlap on the SMs, no host sync in between.
// Every launch must reproduce the 1-CTA kernel's output bit for bit (same k-block order, same epilogue).
static void stress_gemm(const char* name, int M, int N, int K, int a_mn, int b_mn, int flags, int reps) {
  const bool of32 = flags & VJ_EPI_OUT_F32;
  const size_t ob = (size_t)M * N * (of32 ? 4 : 2);
  bf16* A = dalloc<bf16>((size_t)M * K);
  bf16* B = dalloc<bf16>((size_t)N * K);
  float* bias = dalloc<float>(N);
  bf16* side = dalloc<bf16>((size_t)M * N);
  char* ref = dalloc<char>(ob);
  char* out[2] = {dalloc<char>(ob), dalloc<char>(ob)};
  bf16* aux[2] = {dalloc<bf16>((size_t)M * N), dalloc<bf16>((size_t)M * N)};
  unsigned long long* bad = dalloc<unsigned long long>(1);
  CK(cudaMemset(bad, 0, 8));
  fill(A, (long long)M * K, 11, 1.0f); fill(B, (long long)N * K, 12, 0.05f); fillf(bias, N, 13, 1.0f);
  fill(side, (long long)M * N, 14, 1.0f);
  vj_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = A; g.b = B; g.M = M; g.N = N; g.K = K; g.lda = a_mn ? M : K; g.ldb = b_mn ? N : K; g.ldo = N;
  g.a_mn_major = a_mn; g.b_mn_major = b_mn; g.flags = flags; g.bias = bias; g.residual = side; g.ldr = N;
  g.aux_in = side; g.ld_aux = N;
  const int old = vj_gemm_set_pair_mode(0);
  g.out = ref; g.aux_out = aux[0];
  VJ(vj_gemm(&g, 0));
  CK(cudaDeviceSynchronize());
  vj_gemm_set_pair_mode(2);
  for (int r = 0; r < reps; ++r) {
    g.out = out[r & 1]; g.aux_out = aux[r & 1];
    VJ(vj_gemm(&g, 0));
    diff_words_kernel<<<296, 256>>>((const unsigned*)out[r & 1], (const unsigned*)ref, (long long)(ob / 4), bad);
  }
  CK(cudaDeviceSynchronize());
  vj_gemm_set_pair_mode(old);
  unsigned long long hb = 0;
  CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
  printf("[stress gemm] %-26s M=%d N=%d K=%d  %d back-to-back pair launches vs 1-CTA kernel: %llu differing words  %s\n", name,
         M, N, K, reps, hb, hb ? "FAIL" : "ok");
  if (hb) g_fail = 1;
  cudaFree(A); cudaFree(B); cudaFree(bias); cudaFree(side); cudaFree(ref); cudaFree(out[0]); cudaFree(out[1]);
  cudaFree(aux[0]); cudaFree(aux[1]); cudaFree(bad);
}

// ------------------------------------------------------------------ bandwidth-class micro benchmarks (algorithmic GB/s)
static void bench_ln(long long rows, int D, int xdt, int ydt) {
  const size_t xs = xdt == VJ_F32 ? 4 : 2, ys = ydt == VJ_F32 ? 4 : 2;
  void* x = dalloc<char>((size_t)rows * D * xs);
  void* y = dalloc<char>((size_t)rows * D * ys);
  void* dy = dalloc<char>((size_t)rows * D * 2);
  void* dres = dalloc<char>((size_t)rows * D * xs);
  void* dx = dalloc<char>((size_t)rows * D * xs);
  float* gamma = dalloc<float>(D); float* beta = dalloc<float>(D);
  float* mean = dalloc<float>(rows); float* rstd = dalloc<float>(rows);
  float* dg = dalloc<float>(3 * D);
  void* scratch = dalloc<char>(vj_layernorm_bwd_scratch(rows, D));
  if (xdt == VJ_F32) fillf((float*)x, rows * D, 1, 1.0f); else fill((bf16*)x, rows * D, 1, 1.0f);
  fill((bf16*)dy, rows * D, 2, 1.0f);
  if (xdt == VJ_F32) fillf((float*)dres, rows * D, 3, 1.0f); else fill((bf16*)dres, rows * D, 3, 1.0f);
  fillf(gamma, D, 4, 1.0f); fillf(beta, D, 5, 1.0f);
  CK(cudaMemset(dg, 0, 3 * D * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20;
  float ms;
  for (int i = 0; i < 3; ++i) VJ(vj_layernorm_fwd(x, xdt, gamma, beta, y, ydt, mean, rstd, rows, D, 0, 1e-6f, 0));
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) VJ(vj_layernorm_fwd(x, xdt, gamma, beta, y, ydt, mean, rstd, rows, D, 0, 1e-6f, 0));
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  double bytes = (double)rows * D * (xs + ys);
  printf("[bench ln fwd] rows=%lld D=%d x%zu y%zu  %.1f us  %.0f GB/s\n", rows, D, xs, ys, ms * 1e3, bytes / ms * 1e-6);
  for (int i = 0; i < 3; ++i)
    VJ(vj_layernorm_bwd(dy, VJ_BF16, x, xdt, gamma, mean, rstd, dres, dx, xdt, dg, dg + D, dg + 2 * D, scratch, rows, D, 0));
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i)
    VJ(vj_layernorm_bwd(dy, VJ_BF16, x, xdt, gamma, mean, rstd, dres, dx, xdt, dg, dg + D, dg + 2 * D, scratch, rows, D, 0));
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  bytes = (double)rows * D * (2 + 3 * xs);
  printf("[bench ln bwd] rows=%lld D=%d (dx + dgamma/dbeta/dbias)  %.1f us  %.0f GB/s\n", rows, D, ms * 1e3, bytes / ms * 1e-6);
  cudaFree(x); cudaFree(y); cudaFree(dy); cudaFree(dres); cudaFree(dx); cudaFree(gamma); cudaFree(beta); cudaFree(mean);
  cudaFree(rstd); cudaFree(dg); cudaFree(scratch);
}

static void bench_colsum(long long rows, int D) {
  bf16* x = dalloc<bf16>((size_t)rows * D);
  float* out = dalloc<float>(D);
  void* scratch = dalloc<char>(vj_colsum_scratch(rows, D));
  fill(x, rows * D, 1, 1.0f);
  CK(cudaMemset(out, 0, D * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) VJ(vj_colsum(x, VJ_BF16, out, 1, scratch, rows, D, 0));
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) VJ(vj_colsum(x, VJ_BF16, out, 1, scratch, rows, D, 0));
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  printf("[bench colsum] rows=%lld D=%d bf16  %.1f us  %.0f GB/s\n", rows, D, ms * 1e3, (double)rows * D * 2 / ms * 1e-6);
  cudaFree(x); cudaFree(out); cudaFree(scratch);
}

#ifdef VJ_GEMM_PROFILE
extern "C" int vj_gemm_prof_read(unsigned long long* out8, int reset);
#endif

static void bench_gemm(const char* name, long long M, long long N, long long K, int a_mn, int b_mn, int flags) {
  bf16* A = dalloc<bf16>((size_t)M * K);
  bf16* B = dalloc<bf16>((size_t)N * K);
  const bool of32 = flags & VJ_EPI_OUT_F32;
  void* out = dalloc<char>((size_t)M * N * (of32 ? 4 : 2));
  void* side = dalloc<char>((size_t)M * N * 4);   // residual / aux operand (separate buffer unless accumulating)
  CK(cudaMemset(side, 0, (size_t)M * N * 4));
  float* bias = dalloc<float>(N);
  fill(A, M * K, 1, 1.0f); fill(B, N * K, 2, 0.05f); fillf(bias, N, 3, 1.0f);
  CK(cudaMemset(out, 0, (size_t)M * N * (of32 ? 4 : 2)));
  vj_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = A; g.b = B; g.out = out; g.M = M; g.N = N; g.K = K; g.lda = a_mn ? M : K; g.ldb = b_mn ? N : K; g.ldo = N;
  const bool accumulate = (flags & VJ_EPI_RESIDUAL) && of32 && (flags & VJ_EPI_RES_F32) && !(flags & VJ_EPI_ROUND_BF16);
  g.a_mn_major = a_mn; g.b_mn_major = b_mn; g.flags = flags; g.bias = bias; g.residual = accumulate ? out : side; g.ldr = N;
  g.aux_out = side; g.aux_in = side; g.ld_aux = N;
  if (flags & VJ_EPI_ROPE) { g.rope_table = side; g.rope_hd = 64; g.rope_D = (int)(N / 3); }   // zero table: timing only
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) VJ(vj_gemm(&g, 0));
  CK(cudaDeviceSynchronize());
  const int iters = 10;
#ifdef VJ_GEMM_PROFILE
  unsigned long long pr[12];
  vj_gemm_prof_read(pr, 1);
#endif
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) VJ(vj_gemm(&g, 0));
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("[bench gemm] %-28s M=%lld N=%lld K=%lld  %.3f ms  %.1f TFLOP/s\n", name, M, N, K, ms,
         2.0 * M * N * K / ms * 1e-9);
#ifdef VJ_GEMM_PROFILE
  vj_gemm_prof_read(pr, 1);
  const double n = (double)pr[6];
  printf("      per-CTA cycles: mma_total %.0f | mma wait TMA %.1f%% | mma wait epilogue %.1f%% | producer wait slot %.1f%%"
         " | epilogue wait acc %.0f busy %.0f (tmem ld+wait %.0f)\n",
         pr[0] / n, 100.0 * pr[1] / pr[0], 100.0 * pr[2] / pr[0], 100.0 * pr[3] / pr[0], pr[4] / n, pr[5] / n, pr[7] / n);
  printf("      epilogue warp 4: store-read wait %.0f | math+stage %.0f | fence+TMA issue %.0f\n", pr[8] / n, pr[9] / n, pr[10] / n);
#endif
  cudaFree(A); cudaFree(B); cudaFree(out); cudaFree(bias); cudaFree(side);
}

#ifdef VJ_ATTN_PROFILE
extern "C" int vj_attn_prof_read(unsigned long long* out16, int reset);
extern "C" int vj_attn_bwd_prof_read(unsigned long long* out16, int reset);
#endif

static void bench_attn(int B, int S, int H, int hd, bool bwd) {
  const int D = H * hd;
  const size_t nq = (size_t)B * S * 3 * D, no = (size_t)B * S * D, nl = (size_t)B * H * S;
  bf16* qkv = dalloc<bf16>(nq);
  bf16* out = dalloc<bf16>(no);
  bf16* dout = dalloc<bf16>(no);
  bf16* dqkv = dalloc<bf16>(nq);
  float* lse = dalloc<float>(nl);
  void* scratch = dalloc<char>(vj_attn_bwd_scratch(B, S, H, hd));
  fill(qkv, nq, 7, 1.0f); fill(dout, no, 9, 1.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) VJ(vj_attn_fwd(qkv, out, lse, B, S, H, hd, 0));
  CK(cudaDeviceSynchronize());
  const int iters = 5;
#ifdef VJ_ATTN_PROFILE
  unsigned long long ap[16];
  vj_attn_prof_read(ap, 1);
#endif
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) VJ(vj_attn_fwd(qkv, out, lse, B, S, H, hd, 0));
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  const double fl = 4.0 * B * H * (double)S * S * hd;
  printf("[bench attn fwd] B=%d S=%d H=%d hd=%d  %.3f ms  %.1f TFLOP/s\n", B, S, H, hd, ms, fl / ms * 1e-9);
#ifdef VJ_ATTN_PROFILE
  vj_attn_prof_read(ap, 1);
  {
    // dual-stream kernel (hd 64/32): stream 0 of each CTA reports; it owns every other KV tile
    const double n = (double)ap[9], tiles = (double)((((S + 63) / 64) + 1) / 2);
    printf("      stream 0, per own KV tile (cycles): total %.0f = wait S %.0f + tmem ld/release %.0f + max %.0f + wait PV/rescale %.0f"
           " + exp/pack/store %.0f + fence/arrive %.0f\n",
           ap[0] / n / tiles, ap[1] / n / tiles, ap[2] / n / tiles, ap[3] / n / tiles, ap[4] / n / tiles, ap[5] / n / tiles,
           ap[6] / n / tiles);
    const double nt = n * (double)((S + 63) / 64);
    printf("      QK warp, per KV tile (cycles): total %.0f = wait K %.0f + wait Q %.0f + wait S-drain %.0f + issue/other | PV warp: total %.0f"
           " = wait P %.0f + wait V %.0f + wait O-drain %.0f + issue/other\n",
           ap[10] / nt, ap[8] / nt, ap[11] / nt, ap[12] / nt, ap[15] / nt, ap[7] / nt, ap[13] / nt, ap[14] / nt);
  }
#endif
  if (bwd) {
    for (int i = 0; i < 2; ++i) VJ(vj_attn_bwd(qkv, out, dout, lse, dqkv, scratch, nullptr, B, S, H, hd, 0));
    CK(cudaDeviceSynchronize());
#ifdef VJ_ATTN_PROFILE
    { unsigned long long z[16]; vj_attn_bwd_prof_read(z, 1); }
#endif
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) VJ(vj_attn_bwd(qkv, out, dout, lse, dqkv, scratch, nullptr, B, S, H, hd, 0));
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("[bench attn bwd] B=%d S=%d H=%d hd=%d  %.3f ms  %.1f TFLOP/s (2.5x fwd flops)\n", B, S, H, hd, ms,
           2.5 * fl / ms * 1e-9);
#ifdef VJ_ATTN_PROFILE
    {
      unsigned long long bp[16];
      vj_attn_bwd_prof_read(bp, 1);
      const double n = (double)bp[9], it = (double)((S + 127) / 128);
      if (n > 0)
        printf("      per work item (cycles): total %.0f = first S/dP wait %.0f + %d x [barriers %.0f + wait S/dP %.0f + tmem ld %.0f + math %.0f"
               " + wait dQ %.0f + stage dQ %.0f + P/dS store %.0f] + epilogue %.0f\n",
               bp[0] / n, bp[1] / n, (int)it, bp[8] / n / it, bp[2] / n / it, bp[3] / n / it, bp[4] / n / it, bp[5] / n / it,
               bp[6] / n / it, bp[7] / n / it, bp[10] / n);
      if (n > 0)
        printf("      epilogue: wait last MMAs %.0f, wait staging TMA + barrier %.0f; item start (column constants + barrier) %.0f\n",
               bp[11] / n, bp[12] / n, bp[13] / n);
    }
#endif
  }
  cudaFree(qkv); cudaFree(out); cudaFree(dout); cudaFree(dqkv); cudaFree(lse); cudaFree(scratch);
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  const bool all = !strcmp(what, "all");
  int sm, maj, min;
  VJ(vj_device_info(&sm, &maj, &min));
  printf("device: %d SMs, sm_%d%d\n", sm, maj, min);
  if (all || !strcmp(what, "gemm")) {
    test_gemm("fwd tiny", 128, 128, 64, 0, 0, 0);
    test_gemm("fwd 1 tile K=256", 128, 256, 256, 0, 0, 0);
    test_gemm("fwd ragged", 300, 264, 200, 0, 0, 0);
    test_gemm("fwd bias", 515, 1408, 1408, 0, 0, VJ_EPI_BIAS);
    test_gemm("fwd BN176 bias+res bf16", 1000, 4224, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | VJ_EPI_ROUND_BF16);
    test_gemm("fwd bias+gelu+aux", 777, 6144, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_GELU | VJ_EPI_AUX_OUT | VJ_EPI_ROUND_BF16);
    test_gemm("fwd f32 out + f32 res", 260, 384, 1536, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | VJ_EPI_OUT_F32 | VJ_EPI_RES_F32 | VJ_EPI_ROUND_BF16);
    test_gemm("dgrad (B MN) tiny", 128, 128, 64, 0, 1, 0);
    test_gemm("dgrad (B MN) ragged", 300, 1408, 4224, 0, 1, 0);
    test_gemm("dgrad dgelu", 515, 6144, 1408, 0, 1, VJ_EPI_DGELU);
    test_gemm("dgrad N=384 K=1152", 999, 384, 1152, 0, 1, 0);
    test_gemm("wgrad (A,B MN) tiny", 128, 128, 64, 1, 1, VJ_EPI_OUT_F32);
    test_gemm("wgrad ragged accumulate", 1408, 384, 1000, 1, 1, VJ_EPI_OUT_F32 | VJ_EPI_RES_F32 | VJ_EPI_RESIDUAL);
    test_gemm("wgrad 4224x1408 tok 3000", 4224, 1408, 3000, 1, 1, VJ_EPI_OUT_F32);
    test_gemm("wgrad 1408x6144 tok 520", 1408, 6144, 520, 1, 1, VJ_EPI_OUT_F32);
  }
  if (all || !strcmp(what, "attn")) {
    test_attn(1, 128, 1, 64, false);
    test_attn(2, 200, 3, 64, false);
    test_attn(1, 1000, 2, 64, false);
    test_attn(2, 72, 2, 32, false);
    test_attn(1, 700, 3, 32, false);
    test_attn(2, 200, 2, 80, false);
    test_attn(1, 1000, 3, 80, false);
    test_attn(5, 330, 24, 64, false);    // 360 work items (> 2 per SM): persistent CTAs walk several items, odd tile count
    test_attn(7, 64, 50, 32, false);     // single-tile items, 350 of them
  }
  if (all || !strcmp(what, "attnbwd")) {
    test_attn(1, 128, 1, 64, true);
    test_attn(2, 200, 3, 64, true);
    test_attn(1, 520, 2, 64, true);
    test_attn(2, 72, 2, 32, true);
    test_attn(1, 700, 3, 32, true);
    test_attn(2, 200, 2, 80, true);
    test_attn(1, 520, 3, 80, true);
    test_attn(4, 330, 20, 64, true);     // 240 work items on 148 persistent CTAs, 3 query tiles each
    test_attn(6, 100, 40, 32, true);     // single-tile items
  }
  if (all || !strcmp(what, "bench")) {
    bench_gemm("qkv fwd (ViT-g target)", 49152, 4224, 1408, 0, 0, VJ_EPI_BIAS);
    bench_gemm("proj fwd", 49152, 1408, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | VJ_EPI_ROUND_BF16);
    bench_gemm("fc1 fwd gelu", 49152, 6144, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_GELU | VJ_EPI_ROUND_BF16);
    bench_gemm("fc2 fwd", 49152, 1408, 6144, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | VJ_EPI_ROUND_BF16);
    bench_gemm("fc1 dgrad (ctx)", 12096, 1408, 6144, 0, 1, 0);
    bench_gemm("fc2 dgrad dgelu (ctx)", 12096, 6144, 1408, 0, 1, VJ_EPI_DGELU);
    bench_gemm("fc1 wgrad (ctx)", 6144, 1408, 12096, 1, 1, VJ_EPI_OUT_F32 | VJ_EPI_RES_F32 | VJ_EPI_RESIDUAL);
    bench_gemm("qkv wgrad (ctx)", 4224, 1408, 12096, 1, 1, VJ_EPI_OUT_F32 | VJ_EPI_RES_F32 | VJ_EPI_RESIDUAL);
    bench_gemm("proj wgrad (ctx)", 1408, 1408, 12096, 1, 1, VJ_EPI_OUT_F32 | VJ_EPI_RES_F32 | VJ_EPI_RESIDUAL);
    bench_gemm("pred qkv wgrad", 1152, 384, 36000, 1, 1, VJ_EPI_OUT_F32 | VJ_EPI_RES_F32 | VJ_EPI_RESIDUAL);
    bench_gemm("pred proj wgrad", 384, 384, 36000, 1, 1, VJ_EPI_OUT_F32 | VJ_EPI_RES_F32 | VJ_EPI_RESIDUAL);
    bench_gemm("8192^3", 8192, 8192, 8192, 0, 0, 0);
    bench_attn(24, 2048, 22, 64, false);
  }
  if (!strcmp(what, "benchepi")) {   // what each fused epilogue costs on top of the bare GEMM
    const int RB = VJ_EPI_ROUND_BF16;
    bench_gemm("qkv  none", 49152, 4224, 1408, 0, 0, 0);
    bench_gemm("qkv  bias", 49152, 4224, 1408, 0, 0, VJ_EPI_BIAS);
    bench_gemm("qkv  bias+rope", 49152, 4224, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_ROPE);
    bench_gemm("fc1  bias", 49152, 6144, 1408, 0, 0, VJ_EPI_BIAS);
    bench_gemm("fc1  bias+gelu", 49152, 6144, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_GELU | RB);
    bench_gemm("fc1  bias+gelu+aux", 49152, 6144, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_GELU | RB | VJ_EPI_AUX_OUT);
    bench_gemm("proj bias", 49152, 1408, 1408, 0, 0, VJ_EPI_BIAS);
    bench_gemm("proj bias+res", 49152, 1408, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | RB);
    bench_gemm("fc2  bias", 49152, 1408, 6144, 0, 0, VJ_EPI_BIAS);
    bench_gemm("fc2  bias+res", 49152, 1408, 6144, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | RB);
    bench_gemm("dgelu none", 16384, 6144, 1408, 0, 1, 0);
    bench_gemm("dgelu", 16384, 6144, 1408, 0, 1, VJ_EPI_DGELU);
  }
  if (!strcmp(what, "stressgemm")) {   // CTA-pair GEMM: back-to-back launches, bitwise against the 1-CTA kernel
    const int reps = argc > 2 ? atoi(argv[2]) : 200;
    const int RB = VJ_EPI_ROUND_BF16;
    stress_gemm("qkv fwd bias", 12096, 4224, 1408, 0, 0, VJ_EPI_BIAS, reps);
    stress_gemm("fc1 fwd gelu+aux", 12096, 6144, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_GELU | RB | VJ_EPI_AUX_OUT, reps);
    stress_gemm("fc2 fwd +res", 12096, 1408, 6144, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | RB, reps);
    stress_gemm("proj fwd +res (short K)", 49152, 1408, 1408, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | RB, reps / 2);
    stress_gemm("fc2 dgrad dgelu", 12096, 6144, 1408, 0, 1, VJ_EPI_DGELU, reps);
    stress_gemm("fc1 wgrad", 6144, 1408, 12096, 1, 1, VJ_EPI_OUT_F32, reps);
    stress_gemm("pred fc1 (K=384)", 36000, 1536, 384, 0, 0, VJ_EPI_BIAS | VJ_EPI_GELU | RB, reps);
    stress_gemm("ragged", 3001, 1416, 520, 0, 0, VJ_EPI_BIAS, reps);
  }
  if (!strcmp(what, "benchbw")) {      // bandwidth class at step shapes (B=24: target 49152 rows, context ~12096, predictor ~36000)
    bench_ln(49152, 1408, VJ_BF16, VJ_BF16);
    bench_ln(12096, 1408, VJ_BF16, VJ_BF16);
    bench_ln(12096, 1024, VJ_BF16, VJ_BF16);
    bench_ln(12096, 1280, VJ_BF16, VJ_BF16);
    bench_ln(36000, 384, VJ_F32, VJ_BF16);
    bench_colsum(12096, 6144);
    bench_colsum(12096, 4224);
    bench_colsum(36000, 1536);
  }
  if (!strcmp(what, "benchpred")) {     // the predictor's short-K GEMMs (dim 384, hidden 1536, ~55k tokens per step pass)
    const int RB = VJ_EPI_ROUND_BF16;
    bench_gemm("pred qkv bias+rope", 55296, 1152, 384, 0, 0, VJ_EPI_BIAS | VJ_EPI_ROPE);
    bench_gemm("pred proj +res f32", 55296, 384, 384, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | VJ_EPI_RES_F32 | VJ_EPI_OUT_F32 | RB);
    bench_gemm("pred fc1 gelu+aux", 55296, 1536, 384, 0, 0, VJ_EPI_BIAS | VJ_EPI_GELU | RB | VJ_EPI_AUX_OUT);
    bench_gemm("pred fc2 +res f32", 55296, 384, 1536, 0, 0, VJ_EPI_BIAS | VJ_EPI_RESIDUAL | VJ_EPI_RES_F32 | VJ_EPI_OUT_F32 | RB);
    bench_gemm("pred fc2 dgrad dgelu", 55296, 1536, 384, 0, 1, VJ_EPI_DGELU);
    bench_gemm("pred fc1 dgrad", 55296, 384, 1536, 0, 1, 0);
    bench_gemm("pred qkv dgrad", 55296, 384, 1152, 0, 1, 0);
    bench_gemm("pred proj dgrad", 55296, 384, 384, 0, 1, 0);
  }
  if (!strcmp(what, "stressattn")) {   // back-to-back launches (CTAs of consecutive launches overlap on the SMs)
    const int reps = argc > 2 ? atoi(argv[2]) : 20;
    for (int i = 0; i < reps; ++i) {
      bench_attn(24, 2048, 22, 64, false);
      bench_attn(24, 504, 22, 64, i % 4 == 0);
      bench_attn(24, 1448, 12, 32, i % 4 == 1);
      bench_attn(3, 200, 5, 64, false);
    }
  }
  if (!strcmp(what, "benchattn")) {
    bench_attn(6, 8192, 22, 64, false);
    bench_attn(24, 2048, 22, 64, false);
    bench_attn(24, 2048, 12, 32, false);
    bench_attn(24, 504, 22, 64, true);
  }
  if (all || !strcmp(what, "benchbwd")) {
    bench_attn(24, 504, 22, 64, true);
    bench_attn(24, 1448, 12, 32, true);
    bench_attn(24, 504, 16, 80, true);
    bench_attn(24, 2048, 16, 80, false);
  }
  printf(g_fail ? "SELFTEST FAILED\n" : "SELFTEST PASSED\n");
  return g_fail;
}
