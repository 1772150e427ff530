/* vjepa2_b200 -- C ABI of the B200 (sm_100a) kernels behind the V-JEPA 2 pre-training step.
 *
 * The reference (weipeilun/vjepa2) is pure Python/PyTorch and has no FFI of its own: the seam it
 * offers is Python name lookup + nn.Module duck typing (app/vjepa/utils.py:159-190).  This header
 * is therefore the boundary a maintainer binds (ctypes); every entry point
 * cites the reference call site whose vendor kernel(s) it replaces.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - returns 0 on success, negative on error; vj_last_error() gives a thread-local message.
 *   - enqueues on `stream` (a cudaStream_t passed as void*) and never synchronises, allocates or
 *     frees caller memory.  Scratch is passed in by the caller where needed.
 *   - all pointers are device pointers unless named host_*; matrices are row-major.
 *   - no CPU fallback: on a machine without an sm_100 device every launch fails loudly.
 *   - threading: safe from any host thread; calls for ONE device must be enqueued on one stream at a time per
 *     entry point family (vj_colsum keeps self-resetting per-column-group tickets in device memory, so two of them
 *     must not run CONCURRENTLY on the same device; back-to-back on one stream is the intended use).  vj_mask_collate
 *     and vj_peer_barrier keep their state in caller-owned device memory.  Everything else keeps no device state.
 *   - tuning switches read once from the environment (all default to the measured-best setting):
 *     VJ_GEMM_2CTA = 0 | 1 | 2   1-CTA kernels only | CTA-pair (cta_group::2) kernel for M >= 1024 | always;
 *     VJ_GEMM_EPI16 = 0 | 1      CTA-pair kernel: 8 epilogue warps always | 16 for every K-major-A shape (unset: by shape);
 *     VJ_ATTN_POLY = 0..4        eighths of the forward softmax exponentials evaluated on the FMA pipe (default 2);
 *     VJ_ATTN_BWD_POLY = 0..4    same for the backward recomputation (default 0);
 *     VJ_ATTN_BWD_PARTS = 2 | 4  compute threads per key row of the attention backward (default 2).
 *     Host side (vjepa2_b200/train.py): VJ_DDP_COMM = peer | nccl, VJ_DDP_SYNC = overlap | end, VJ_DDP_BUCKET_MB.
 */
#ifndef VJEPA2_B200_H
#define VJEPA2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VJ_BF16 = 0, VJ_F32 = 1 };

/* ------------------------------------------------------------------ library */
const char* vj_last_error(void);
int vj_abi_version(void);
/* sm count / compute capability of the current device; <0 if no usable device */
int vj_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ GEMM (tcgen05 + TMEM + TMA)
 * out[M,N] = epilogue( A[M,K] * B[N,K]^T ), bf16 operands, fp32 accumulation in tensor memory.
 * Replaces nn.Linear / nn.Conv3d-as-GEMM forward, dgrad and wgrad:
 *   modules.py:330 (qkv), :380 (proj), :78-81 (fc1/fc2), patch_embed.py:51, predictor.py:182,244.
 * Operand storage: a_mn_major==0 -> A stored [M][K] (K contiguous, ld=lda);
 *                  a_mn_major==1 -> A stored [K][M] (M contiguous, ld=lda)   (same for B / N).
 *   forward  y = x W^T      : A=x  [M][K],        B=W  [N][K]            (0,0)
 *   dgrad    dx = dy W      : A=dy [M][N'],       B=W  [N'][K'] as [K][N] (0,1)
 *   wgrad    dW = dy^T x    : A=dy [tok][N'] (1), B=x  [tok][K'] (1)
 * Epilogue order: v = acc; +bias[n]; aux_out=bf16(v); v=bf16round(v) if ROUND; gelu; *gelu'(aux_in);
 *                 +residual; store (bf16 or fp32).
 */
enum {
  VJ_EPI_BIAS = 1,       /* v += bias[n] (fp32 [N]) */
  VJ_EPI_GELU = 2,       /* v = gelu_erf(v)  (nn.GELU default, modules.py:73) */
  VJ_EPI_DGELU = 4,      /* v *= gelu'(aux_in[m,n])  (backward through the activation) */
  VJ_EPI_RESIDUAL = 8,   /* v += residual[m,n] */
  VJ_EPI_OUT_F32 = 16,   /* out is fp32 (else bf16) */
  VJ_EPI_RES_F32 = 32,   /* residual is fp32 (else bf16) */
  VJ_EPI_ROUND_BF16 = 64,/* round v to bf16 before gelu/residual (mirrors autocast's bf16 Linear output) */
  VJ_EPI_AUX_OUT = 128,  /* also write the pre-activation v as bf16 to aux_out */
  VJ_EPI_ROPE = 256,     /* qkv projection: apply the 3-axis RoPE map to columns < 2*rope_D (q and k thirds)
                            after bias, using rope_table[m] (vj_rope_table layout); v passes through */
  VJ_EPI_BIAS_GRAD = 512 /* weight-gradient GEMM (A, B MN-major, fp32 out): B has 8 more columns than N, all 1.0 (ldb >=
                            N + 8, written by vj_layernorm_fwd's padded output), so accumulator column N of row m is
                            sum_k A[k][m] -- the bias gradient of the same Linear; bias_grad[m] += it.  Costs 8 extra
                            MMA columns instead of a column-sum pass over A.  CTA-pair kernel only (M >= 1024). */
};

typedef struct {
  const void* a;       /* bf16 */
  const void* b;       /* bf16 */
  void* out;           /* [M][N], ld = ldo */
  int64_t M, N, K;
  int64_t lda, ldb, ldo; /* leading dimensions in elements */
  int32_t a_mn_major, b_mn_major;
  int32_t flags;
  const float* bias;
  const void* residual;  /* ld = ldr; may alias out (fp32 accumulate) */
  int64_t ldr;
  void* aux_out;         /* bf16, ld = ld_aux */
  const void* aux_in;    /* bf16, ld = ld_aux */
  int64_t ld_aux;
  const void* rope_table; /* fp16 [M][2][rope_hd] (VJ_EPI_ROPE) */
  int32_t rope_hd;        /* head dim (32, 64 or 80) */
  int32_t rope_D;         /* model width: N == 3*rope_D */
  float* bias_grad;       /* fp32 [M] (VJ_EPI_BIAS_GRAD), accumulated */
} vj_gemm_args;

int vj_gemm(const vj_gemm_args* a, void* stream);

/* ------------------------------------------------------------------ LayerNorm
 * nn.LayerNorm(eps=1e-6) norm1/norm2/norm/predictor_norm (modules.py:558,562;
 * vision_transformer.py:211; predictor.py:233) and F.layer_norm without affine, eps 1e-5
 * (train.py:417; gamma/beta NULL).  Statistics in fp32.  mean/rstd may be NULL in forward.
 * ldy: row pitch of y in elements (0 = D).  With ldy > D (a multiple of 8 more) the pad columns [D, ldy) of every
 * row are set to 1.0: vj_gemm's VJ_EPI_BIAS_GRAD reads them as the ones-column that yields a bias gradient. */
int vj_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype,
                     float* mean, float* rstd, int64_t rows, int64_t D, int64_t ldy, float eps, void* stream);
/* dx = LN'(dy) (+ dres if given, same dtype as dx), one pass over dy / x / dres that also ACCUMULATES (+=) the column
 * sums dgamma = sum dy*xhat, dbeta = sum dy and dbias = sum dres (fp32 [D] each, any may be NULL): dres is the
 * gradient of the residual branch's Linear output, so dbias is that Linear's bias gradient (fc2 in norm2's backward,
 * proj in norm1's).  Deterministic (per-CTA partials in scratch, added in CTA order by a second small kernel).
 * scratch: at least vj_layernorm_bwd_scratch(rows, D) bytes when any column sum is requested. */
size_t vj_layernorm_bwd_scratch(int64_t rows, int64_t D);
int vj_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* gamma,
                     const float* mean, const float* rstd, const void* dres, void* dx, int dx_dtype,
                     float* dgamma, float* dbeta, float* dbias, void* scratch, int64_t rows, int64_t D, void* stream);

/* ------------------------------------------------------------------ 3-axis RoPE
 * rotate_queries_or_keys + separate_positions (modules.py:26-50, 311-365).
 * vj_rope_table: token ids (int64, as produced by MaskCollator; NULL = unmasked sequence, id(row) =
 *   row % period, the torch.arange of modules.py:337-341) -> fp16 table [n][2][head_dim]: per ELEMENT d of the
 *   head cos(theta) then sin(theta) with theta = pos_axis(d) * 10000^(-j/(seg/2)), seg = 2*((head_dim/3)/2),
 *   axis(d) = d/seg (frame, height, width), j = (d mod seg) mod (seg/2) -- the TILED angle layout of the
 *   reference; dims >= 3*seg pass through (cos 1, sin 0).
 * vj_rope_apply: in place on the q and k thirds of qkv [rows][3*D] (bf16); transpose != 0 applies the adjoint
 *   (backward).  The per-pair map [[c(2k), -s(2k)],[s(2k+1), c(2k+1)]] is NOT a rotation (modules.py:40-50).
 * The same table feeds the fused paths: VJ_EPI_ROPE in vj_gemm (forward) and vj_attn_bwd (adjoint). */
int vj_rope_table(const int64_t* ids, int64_t n, int64_t period, int Hp, int Wp, int head_dim, void* table,
                  void* stream);
int vj_rope_apply(void* qkv, int64_t rows, int64_t D, int heads, int head_dim, const void* table, int transpose,
                  void* stream);

/* ------------------------------------------------------------------ attention (tcgen05 flash fwd/bwd)
 * F.scaled_dot_product_attention, non-causal, no mask, dropout 0, scale 1/sqrt(d) (modules.py:369).
 * qkv: [B*S][3*D] bf16, feature = which*D + head*d + i (modules.py:330-331), q/k already rotated.
 * out: [B*S][D] bf16 (head-major features = x.transpose(1,2).reshape, modules.py:379).
 * lse: [B][H][S] fp32, log2-domain log-sum-exp (saved for backward).
 * head_dim in {32, 64, 80} (80 = ViT-H: handled as a 64 + 16 split of every head-dim operand).  S >= 1. */
int vj_attn_fwd(const void* qkv, void* out, float* lse, int B, int S, int H, int head_dim, void* stream);
/* dqkv: [B*S][3*D] bf16: gradient w.r.t. the rotated q/k and v; if rope_table != NULL (vj_rope_table layout,
 * rows = B*S) the adjoint RoPE map is applied to dq and dk on the way out, giving the gradient w.r.t. the
 * un-rotated qkv projection output.
 * scratch: vj_attn_bwd_scratch bytes (fp32 dq accumulators + delta). */
size_t vj_attn_bwd_scratch(int B, int S, int H, int head_dim);
int vj_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* scratch,
                const void* rope_table, int B, int S, int H, int head_dim, void* stream);

/* ------------------------------------------------------------------ row gather / scatter
 * apply_masks (masks/utils.py:9-21): dst[r,:] = src[index[r],:]; index[r] < 0 -> fill[:] (or 0).
 * Bit-exact copy when dtypes match; converts otherwise.  index holds ABSOLUTE source rows. */
int vj_gather_rows(const void* src, int src_dtype, void* dst, int dst_dtype, const int64_t* index,
                   const float* fill, int64_t n_out, int64_t D, void* stream);
/* dst[index[r],:] += src[r,:]  (fp32 atomics; adjoint of gather for duplicate-capable indices) */
int vj_scatter_add_rows(const void* src, int src_dtype, float* dst, const int64_t* index, int64_t n_src,
                        int64_t D, void* stream);
/* absolute gather index for apply_masks: out[b*K+k] = b*N + masks[b*K+k] */
int vj_mask_to_rows(const int64_t* masks, int64_t* out, int64_t B, int64_t K, int64_t N, void* stream);

/* ------------------------------------------------------------------ patch embed im2col
 * PatchEmbed3D (patch_embed.py:49-52): Conv3d k=s=(tub,p,p) as im2col (+ vj_gemm).
 * clips fp32 [B][C][T][H][W]; ids int64 [B*reps][K] token ids to keep, row block j*B..(j+1)*B-1 being
 * mask j of apply_masks(concat=True) (NULL = all T/tub*H/p*W/p tokens, K and reps ignored);
 * cols bf16 [B*reps*K][C*tub*p*p], K order (c,kt,kh,kw).  Gathering BEFORE the GEMM computes only
 * the kept tubelets (the reference embeds all tokens, then gathers: vision_transformer.py:188-192). */
int vj_im2col_tubelets(const float* clips, const int64_t* ids, void* cols, int B, int C, int T, int H, int W,
                       int tubelet, int patch, int64_t K, int reps, void* stream);

/* ------------------------------------------------------------------ reductions / loss
 * out[D] (+)= sum_r x[r,:]   (bias gradients).  scratch >= vj_colsum_scratch bytes. */
size_t vj_colsum_scratch(int64_t rows, int64_t D);
int vj_colsum(const void* x, int x_dtype, float* out, int accumulate, void* scratch, int64_t rows, int64_t D,
              void* stream);
/* loss_fn (train.py:425-435) for one mask: loss_accum += loss_scale * sum |z - h[idx]|,
 * dz = grad_scale * sign(z - h[idx]) (bf16; NULL to skip).  z bf16 [B][K][D]; h fp32 [B][N][D];
 * idx int64 [B][K].  grad_scale_mul: optional device scalar multiplied into grad_scale (the
 * GradScaler loss scale, train.py:445).  scratch >= vj_l1_scratch bytes.  Deterministic two-stage
 * reduction. */
size_t vj_l1_scratch(int64_t B, int64_t K, int64_t D);
int vj_l1_loss(const void* z, const float* h, const int64_t* idx, float* loss_accum, void* dz, float loss_scale,
               float grad_scale, const float* grad_scale_mul, void* scratch, int64_t B, int64_t K, int64_t N,
               int64_t D, void* stream);

/* ------------------------------------------------------------------ predictor token order
 * predictor.py:210-217,240-241: rank[b][i] = position of element i in the ascending (stable) order
 * of ids[b][:]  (== argsort(argsort(ids))). */
int vj_argsort_rank(const int64_t* ids, int32_t* rank, int64_t B, int64_t S, void* stream);
/* All index tensors of VisionTransformerPredictor.forward in one pass (predictor.py:206-217,240-242).
 * S = Kc + Kp; element i of cat(masks_x, masks_y)[b] has stable ascending rank r:
 *   ids_sorted[b*S + r] = id            asm_idx[b*S + r]    = b*Kc + i (context) or -1 (mask token)
 *   ctx_pos[b*Kc + i]   = b*S + r       tgt_pos[b*Kp + k]   = b*S + r  (i = Kc + k)
 *   seq_to_tgt[b*S + r] = b*Kp + k (target) or -1 */
int vj_pred_indices(const int64_t* masks_x, const int64_t* masks_y, int64_t B, int64_t Kc, int64_t Kp,
                    int64_t* ids_sorted, int64_t* asm_idx, int64_t* tgt_pos, int64_t* ctx_pos,
                    int64_t* seq_to_tgt, void* stream);

/* ------------------------------------------------------------------ flat optimizer kernels
 * EMA (train.py:457-465): tgt = fma(1-m, src, tgt*m) over a flat fp32 buffer (the rounding order of
 * torch._foreach_mul_ followed by _foreach_add_(alpha=1-m)); optionally refreshes the bf16 shadow of
 * the target weights.  m and 1-m are both passed (computed in double on the host). */
int vj_ema_update(float* tgt, const float* src, void* tgt_bf16, int64_t n, float m, float one_minus_m,
                  void* stream);
/* found_inf[0] = 1.0f if any g*inv_scale is non-finite (GradScaler.unscale_ check, train.py:446). */
int vj_grad_check(const float* g, int64_t n, float* found_inf, void* stream);
/* torch.optim.AdamW (app/vjepa/utils.py:239) over a flat buffer, grads multiplied by *inv_scale,
 * skipped entirely when *found_inf != 0 (GradScaler.step).  tile_flags: one byte per 1024
 * elements: bit0 = weight decay applies, bit1 = frozen (no grad ever produced).  Also refreshes
 * the bf16 shadow.  bias corrections 1-b1^t, 1-b2^t: host scalars, or -- when dev_bias_c is non-null -- the two
 * device floats written by vj_adam_prepare (the host scalars are then ignored). */
int vj_adamw_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, void* p_bf16,
                  const uint8_t* tile_flags, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
                  float bias_c1, float bias_c2, const float* dev_bias_c, const float* inv_scale,
                  const float* found_inf, void* stream);
/* torch's optimizer step count does not advance on a step GradScaler skips (train.py:447-450).  Keeps that count on
 * the device without a host sync: t = step - *skipped (step = number of calls so far, this one included); writes
 * bias_c[0..1] = 1-b1^t, 1-b2^t (double precision) and, if *found_inf != 0, counts this step as skipped.
 * Call once per step after vj_grad_check and before the vj_adamw_step launches. */
int vj_adam_prepare(float* bias_c, int32_t* skipped, const float* found_inf, int step, double beta1, double beta2,
                    void* stream);
/* GradScaler.update() on device scalars (train.py:451): scale *= backoff if *found_inf else grows by
 * `growth` every `interval` clean steps; writes inv_scale = 1/(scale*world), clears found_inf. */
int vj_scaler_update(float* scale, float* inv_scale, int32_t* growth_tracker, float* found_inf, float growth,
                     float backoff, int interval, float world, void* stream);
/* p[0..n) = value (optimizer.zero_grad() of the flat gradient buffers, train.py:454; loss accumulator reset) */
int vj_fill_f32(float* p, int64_t n, float value, void* stream);
/* fp32 -> bf16 flat cast (weight shadow refresh) */
int vj_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------ device-side MaskCollator
 * src/masks/multiseq_multiblock3d.py:129-239 (_MaskGenerator.__call__) as one kernel, RNG-call-identical to the
 * reference: the block extent comes from mt19937(seed) (seed = the generator's draw counter, :170-178), the block
 * corners from the GLOBAL torch CPU generator, whose MT19937 state lives on the device in rng_state
 * (VJ_MASK_RNG_WORDS uint32: state[624], left, next -- the fields of torch.get_rng_state()) and is advanced in place.
 * masks_enc / masks_pred: device buffers of capacity B*frames*rows*cols int64, written DENSE as [B][K_enc] and
 * [B][K_pred] (sorted token ids, truncated to the batch minimum / max_keep, :199-213; complement variants :214-231).
 * counts[0..1] = K_enc, K_pred (device ints; the host reads them one step ahead of the step that uses the masks).
 * scratch: vj_mask_collate_scratch bytes.  inv_block (:236-239) is a swap of the two outputs on the caller's side. */
#define VJ_MASK_RNG_WORDS 626
typedef struct {
  int32_t frames, rows, cols;      /* token grid: duration, height, width (:104-105) */
  int32_t num_blocks;              /* npred */
  int32_t context_frames;          /* max_context_duration (:116-118) */
  int32_t max_keep;                /* <= 0: none */
  int32_t full_complement, pred_full_complement;
  double temporal_lo, temporal_hi; /* temporal_pred_mask_scale */
  double spatial_lo, spatial_hi;   /* spatial_pred_mask_scale */
  double aspect_lo, aspect_hi;     /* aspect_ratio */
} vj_mask_spec;
size_t vj_mask_collate_scratch(const vj_mask_spec* spec, int64_t B);
int vj_mask_collate(uint32_t* rng_state, const vj_mask_spec* spec, uint32_t seed, int64_t B, int64_t* masks_enc,
                    int64_t* masks_pred, int32_t* counts, void* scratch, void* stream);

/* ------------------------------------------------------------------ data-parallel gradient all-reduce, device side
 * app/vjepa/train.py:279-281 (DistributedDataParallel's gradient mean).  The bytes move on the copy engines between
 * peer-mapped (symmetric) buffers, issued by the host side (vjepa2_b200/train.py: PeerGradReducer); these are the two
 * kernels that need an SM.  vj_ptr_list carries up to VJ_MAX_PEERS device pointers by value.
 * vj_peer_barrier: flags.ptr[p] = rank p's flag array (uint32[VJ_MAX_PEERS], zero-initialised, mapped into this
 * process); publishes `epoch` (monotonically increasing per call, same sequence on every rank) to every peer and waits
 * until every peer has published an epoch >= it.  One warp; traps after ~20 s instead of hanging.
 * vj_sum_into: dst[i] += srcs.ptr[0][i] + ... + srcs.ptr[n_src-1][i] (fp32, in that order), n % 4 == 0. */
#define VJ_MAX_PEERS 16
typedef struct {
  void* ptr[VJ_MAX_PEERS];
} vj_ptr_list;
int vj_peer_barrier(const vj_ptr_list* flags, int rank, int world, uint32_t epoch, void* stream);
int vj_sum_into(float* dst, const vj_ptr_list* srcs, int n_src, int64_t n, void* stream);

/* Tuning / test switch of vj_gemm's kernel choice (same as the VJ_GEMM_2CTA environment variable): 0 = 1-CTA kernels
 * only, 1 = automatic (CTA-pair kernel for M >= 1024), 2 = CTA-pair kernel for every shape.  Returns the old mode. */
int vj_gemm_set_pair_mode(int mode);
/* Same for the CTA-pair kernel's epilogue width (VJ_GEMM_EPI16): 0 = 8 epilogue warps always, 1 = 16 epilogue warps for
 * every K-major-A shape, 2 = automatic (short K, or K <= 2048 with a side-operand epilogue).  Returns the old mode. */
int vj_gemm_set_epi16_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* VJEPA2_B200_H */
