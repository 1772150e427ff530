"""Tensor-level wrappers over the C ABI.  PyTorch supplies device memory and streams only; every
function here enqueues hand-written sm_100a kernels on the current CUDA stream."""
from __future__ import annotations

import ctypes
import os

import torch

from . import _cabi as C

BF16, F32 = torch.bfloat16, torch.float32


def _dt(t: torch.Tensor) -> int:
    if t.dtype == BF16:
        return C.VJ_BF16
    if t.dtype == F32:
        return C.VJ_F32
    raise TypeError(f"vjepa2_b200: unsupported dtype {t.dtype}")


def _p(t):
    return None if t is None else t.data_ptr()


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vjepa2_b200: expected CUDA tensors (there is no CPU path)")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# kernels launched per C-ABI call (bench.py reports the total as `gpu_launches`); layernorm_bwd = dx kernel + one
# column-reduction kernel (every model width is a multiple of 8), colsum = one column-reduction kernel
KERNELS_PER_CALL = {
    "vj_gemm": 1, "vj_layernorm_fwd": 1, "vj_layernorm_bwd": 2, "vj_rope_table": 1, "vj_rope_apply": 1,
    "vj_attn_fwd": 1, "vj_attn_bwd": 3, "vj_gather_rows": 1, "vj_scatter_add_rows": 1, "vj_mask_to_rows": 1,
    "vj_im2col_tubelets": 1, "vj_colsum": 1, "vj_l1_loss": 2, "vj_argsort_rank": 1, "vj_pred_indices": 1,
    "vj_ema_update": 1, "vj_grad_check": 1, "vj_adamw_step": 1, "vj_adam_prepare": 1, "vj_scaler_update": 1, "vj_cast_f32_bf16": 1,
}
LAUNCHES = 0
_real_check = C.check


def _counting_check(rc, what=""):
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(what, 1)
    _real_check(rc, what)


# ------------------------------------------------------------------------------ GEMM
_gemm_args = C.GemmArgs()


def gemm(a, b, out, M, N, K, *, a_mn=False, b_mn=False, bias=None, gelu=False, dgelu_aux=None, residual=None,
         aux_out=None, round_bf16=False, rope=None, bias_grad=None, st=None):
    """out[M,N] = epi(A[M,K] @ B[N,K]^T).  a/b: 2-D bf16 views whose last dim is contiguous
    (a: [M,K] or, if a_mn, [K,M]; b likewise).  out: bf16 or fp32 [M,N]."""
    g = _gemm_args
    flags = 0
    if bias is not None:
        flags |= C.EPI_BIAS
    if gelu:
        flags |= C.EPI_GELU
    if dgelu_aux is not None:
        flags |= C.EPI_DGELU
    if residual is not None:
        flags |= C.EPI_RESIDUAL
        if residual.dtype == F32:
            flags |= C.EPI_RES_F32
    if out.dtype == F32:
        flags |= C.EPI_OUT_F32
    if round_bf16:
        flags |= C.EPI_ROUND_BF16
    if aux_out is not None:
        flags |= C.EPI_AUX_OUT
    if rope is not None:                    # (table, head_dim, D): fused 3-axis RoPE on the q/k thirds
        flags |= C.EPI_ROPE
        g.rope_table, g.rope_hd, g.rope_D = rope[0].data_ptr(), rope[1], rope[2]
    if bias_grad is not None:               # wgrad only: b carries 8 pad columns of ones (layernorm_fwd padded output)
        flags |= C.EPI_BIAS_GRAD
    g.bias_grad = _p(bias_grad)
    g.a, g.b, g.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldo = a.stride(0), b.stride(0), out.stride(0)
    g.a_mn_major, g.b_mn_major, g.flags = int(a_mn), int(b_mn), flags
    g.bias = _p(bias)
    g.residual = _p(residual)
    g.ldr = residual.stride(0) if residual is not None else 0
    aux = aux_out if aux_out is not None else dgelu_aux
    g.aux_out = _p(aux_out)
    g.aux_in = _p(dgelu_aux)
    g.ld_aux = aux.stride(0) if aux is not None else 0
    _counting_check(C.load().vj_gemm(ctypes.byref(g), st if st is not None else stream()), "vj_gemm")
    return out


def bias_grad_pad(D, *out_dims):
    """Pad columns (0 or 8) the LayerNorm output of width D needs so that the wgrad GEMMs reading it produce their bias
    gradients (VJ_EPI_BIAS_GRAD): only when the ones-block rides in the last, partly filled 256-column tile (no extra
    tile) and every wgrad is large enough for the CTA-pair kernel."""
    if D % 256 == 0 or D % 256 + 8 > 256 or any(m < 1024 for m in out_dims):
        return 0
    if os.environ.get("VJ_BIAS_GRAD_PAD", "1") == "0":          # A/B switch: column-sum kernels instead
        return 0
    if C.load().vj_gemm_set_pair_mode(-1) == 0:
        return 0
    return 8


# ------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x, gamma, beta, y, mean=None, rstd=None, eps=1e-6, st=None):
    """y may be wider than x ([rows, D + pad], pad % 8 == 0): the pad columns are set to 1 (bias-gradient ones-column
    of the wgrad GEMM that reads y as its B operand)."""
    rows, D = x.shape
    ldy = y.shape[1] if y.dim() == 2 else D
    _counting_check(C.load().vj_layernorm_fwd(x.data_ptr(), _dt(x), _p(gamma), _p(beta), y.data_ptr(), _dt(y), _p(mean),
                                      _p(rstd), rows, D, ldy, eps, st if st is not None else stream()),
            "vj_layernorm_fwd")
    return y


def _scratch(nbytes, device, alloc):
    if alloc is not None:
        return alloc((nbytes,), torch.uint8)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dres=None, dgamma=None, dbeta=None, dbias=None, st=None, alloc=None):
    """dbias (fp32 [D], +=): column sum of dres = bias gradient of the Linear whose output the residual branch added."""
    rows, D = x.shape
    scratch = None
    if dgamma is not None or dbeta is not None or dbias is not None:
        scratch = _scratch(C.load().vj_layernorm_bwd_scratch(rows, D), x.device, alloc)
    if dres is not None and dres.dtype != dx.dtype:
        raise TypeError("layernorm_bwd: dres must have dx's dtype")
    _counting_check(C.load().vj_layernorm_bwd(dy.data_ptr(), _dt(dy), x.data_ptr(), _dt(x), _p(gamma), mean.data_ptr(),
                                      rstd.data_ptr(), _p(dres), dx.data_ptr(), _dt(dx), _p(dgamma), _p(dbeta),
                                      _p(dbias), _p(scratch), rows, D, st if st is not None else stream()),
            "vj_layernorm_bwd")
    return dx


# ------------------------------------------------------------------------------ RoPE
def rope_seg(head_dim: int) -> int:
    return 2 * ((head_dim // 3) // 2)


def rope_table(ids, n, period, Hp, Wp, head_dim, device, st=None, alloc=None, out=None):
    """ids: int64 [n] flattened token ids, or None for id(row) = row % period -> fp16 table [n, 2, head_dim]
    (per-element cos then sin, see include/vjepa2_b200.h).  out: optional pre-allocated table (rows of a larger one)."""
    shape = (n, 2, head_dim)
    if out is not None:
        table = out
    else:
        table = alloc(shape, torch.float16) if alloc is not None else torch.empty(shape, dtype=torch.float16, device=device)
    _counting_check(C.load().vj_rope_table(_p(ids), n, period, Hp, Wp, head_dim, table.data_ptr(),
                                           st if st is not None else stream()), "vj_rope_table")
    return table


def rope_apply(qkv, D, heads, head_dim, table, transpose=False, st=None):
    rows = qkv.shape[0]
    _counting_check(C.load().vj_rope_apply(qkv.data_ptr(), rows, D, heads, head_dim, table.data_ptr(),
                                           int(transpose), st if st is not None else stream()), "vj_rope_apply")
    return qkv


# ------------------------------------------------------------------------------ attention
def attn_fwd(qkv, out, lse, B, S, H, head_dim, st=None):
    _counting_check(C.load().vj_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, head_dim,
                                 st if st is not None else stream()), "vj_attn_fwd")
    return out


def attn_bwd(qkv, out, dout, lse, dqkv, B, S, H, head_dim, st=None, alloc=None, rope=None):
    """rope: optional fp16 table (rope_table layout): the adjoint RoPE map is fused into the dq / dk outputs."""
    scratch = _scratch(C.load().vj_attn_bwd_scratch(B, S, H, head_dim), qkv.device, alloc)
    _counting_check(C.load().vj_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                         dqkv.data_ptr(), scratch.data_ptr(), _p(rope), B, S, H, head_dim,
                                         st if st is not None else stream()), "vj_attn_bwd")
    return dqkv


# ------------------------------------------------------------------------------ gather / scatter / im2col
def gather_rows(src, dst, index, fill=None, st=None):
    n_out, D = dst.shape
    _counting_check(C.load().vj_gather_rows(_p(src), _dt(src) if src is not None else C.VJ_F32, dst.data_ptr(), _dt(dst),
                                    index.data_ptr(), _p(fill), n_out, D, st if st is not None else stream()),
            "vj_gather_rows")
    return dst


def scatter_add_rows(src, dst, index, st=None):
    n_src, D = src.shape
    _counting_check(C.load().vj_scatter_add_rows(src.data_ptr(), _dt(src), dst.data_ptr(), index.data_ptr(), n_src, D,
                                         st if st is not None else stream()), "vj_scatter_add_rows")
    return dst


def mask_to_rows(masks, N, st=None):
    B, K = masks.shape
    out = torch.empty(B * K, dtype=torch.int64, device=masks.device)
    _counting_check(C.load().vj_mask_to_rows(masks.data_ptr(), out.data_ptr(), B, K, N, st if st is not None else stream()),
            "vj_mask_to_rows")
    return out


def im2col_tubelets(clips, ids, tubelet, patch, st=None, alloc=None, out=None):
    """clips fp32 [B,C,T,H,W]; ids int64 [B,K] or None -> bf16 [B*K, C*tubelet*patch*patch].
    out: optional pre-allocated destination (rows of a larger matrix)."""
    B, Cc, T, H, W = clips.shape
    if ids is not None:
        reps, K = ids.shape[0] // B, ids.shape[1]
    else:
        reps, K = 1, (T // tubelet) * (H // patch) * (W // patch)
    shape = (B * reps * K, Cc * tubelet * patch * patch)
    if out is not None:
        cols = out
    else:
        cols = alloc(shape, BF16) if alloc is not None else torch.empty(shape, dtype=BF16, device=clips.device)
    _counting_check(C.load().vj_im2col_tubelets(clips.data_ptr(), _p(ids), cols.data_ptr(), B, Cc, T, H, W, tubelet, patch, K,
                                        reps, st if st is not None else stream()), "vj_im2col_tubelets")
    return cols


# ------------------------------------------------------------------------------ reductions / loss
def colsum(x, out, accumulate=True, st=None, alloc=None):
    rows, D = x.shape
    scratch = _scratch(C.load().vj_colsum_scratch(rows, D), x.device, alloc)
    _counting_check(C.load().vj_colsum(x.data_ptr(), _dt(x), out.data_ptr(), int(accumulate), scratch.data_ptr(), rows, D,
                               st if st is not None else stream()), "vj_colsum")
    return out


def l1_loss(z, h, idx, loss_accum, dz, loss_scale, grad_scale, grad_scale_mul=None, st=None, alloc=None):
    """z bf16 [B,K,D]; h fp32 [B,N,D]; idx int64 [B,K].  loss_accum (fp32 [1]) += loss_scale*sum|z-h[idx]|."""
    B, K, D = z.shape
    N = h.shape[1]
    scratch = _scratch(C.load().vj_l1_scratch(B, K, D), z.device, alloc)
    _counting_check(C.load().vj_l1_loss(z.data_ptr(), h.data_ptr(), idx.data_ptr(), loss_accum.data_ptr(), _p(dz),
                                loss_scale, grad_scale, _p(grad_scale_mul), scratch.data_ptr(), B, K, N, D,
                                st if st is not None else stream()), "vj_l1_loss")
    return loss_accum


def argsort_rank(ids, st=None):
    B, S = ids.shape
    rank = torch.empty(B, S, dtype=torch.int32, device=ids.device)
    _counting_check(C.load().vj_argsort_rank(ids.data_ptr(), rank.data_ptr(), B, S, st if st is not None else stream()),
            "vj_argsort_rank")
    return rank


def pred_indices(masks_x, masks_y, st=None, alloc=None):
    B, Kc = masks_x.shape
    Kp = masks_y.shape[1]
    S = Kc + Kp
    dev = masks_x.device
    i64 = torch.int64
    if alloc is None:
        alloc = lambda shape, dtype: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731
    ids_sorted = alloc((B, S), i64)
    asm_idx = alloc((B * S,), i64)
    tgt_pos = alloc((B * Kp,), i64)
    ctx_pos = alloc((B * Kc,), i64)
    seq_to_tgt = alloc((B * S,), i64)
    _counting_check(C.load().vj_pred_indices(masks_x.data_ptr(), masks_y.data_ptr(), B, Kc, Kp, ids_sorted.data_ptr(),
                                     asm_idx.data_ptr(), tgt_pos.data_ptr(), ctx_pos.data_ptr(),
                                     seq_to_tgt.data_ptr(), st if st is not None else stream()), "vj_pred_indices")
    return ids_sorted, asm_idx, tgt_pos, ctx_pos, seq_to_tgt


# ------------------------------------------------------------------------------ flat optimizer kernels
def ema_update(tgt, src, tgt_bf16, m, st=None):
    n = tgt.numel()
    _counting_check(C.load().vj_ema_update(tgt.data_ptr(), src.data_ptr(), _p(tgt_bf16), n, float(m), float(1.0 - m),
                                   st if st is not None else stream()), "vj_ema_update")


def grad_check(g, found_inf, st=None):
    _counting_check(C.load().vj_grad_check(g.data_ptr(), g.numel(), found_inf.data_ptr(),
                                   st if st is not None else stream()), "vj_grad_check")


def adamw_step(p, g, m, v, p_bf16, tile_flags, lr, beta1, beta2, eps, wd, step, inv_scale=None, found_inf=None,
               st=None, dev_bias=None):
    """dev_bias: the two device floats of adam_prepare (step count that skips inf-skipped steps, like torch's);
    without it the bias corrections come from the host-side `step`."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    _counting_check(C.load().vj_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _p(p_bf16),
                                   tile_flags.data_ptr(), p.numel(), lr, beta1, beta2, eps, wd, bc1, bc2,
                                   _p(dev_bias), _p(inv_scale), _p(found_inf), st if st is not None else stream()),
            "vj_adamw_step")


def adam_prepare(bias_c, skipped, found_inf, step, beta1, beta2, st=None):
    _counting_check(C.load().vj_adam_prepare(bias_c.data_ptr(), skipped.data_ptr(), _p(found_inf), int(step), float(beta1),
                                             float(beta2), st if st is not None else stream()), "vj_adam_prepare")


def scaler_update(scale, inv_scale, growth_tracker, found_inf, world=1.0, growth=2.0, backoff=0.5, interval=2000,
                  st=None):
    _counting_check(C.load().vj_scaler_update(scale.data_ptr(), inv_scale.data_ptr(), growth_tracker.data_ptr(),
                                      found_inf.data_ptr(), growth, backoff, interval, float(world),
                                      st if st is not None else stream()), "vj_scaler_update")


def fill_f32(t, value=0.0, st=None):
    """t[...] = value for a contiguous fp32 tensor (zero_grad of the flat gradient buffers)."""
    _counting_check(C.load().vj_fill_f32(t.data_ptr(), t.numel(), float(value), st if st is not None else stream()), "vj_fill_f32")
    return t


def cast_f32_bf16(src, dst, st=None):
    _counting_check(C.load().vj_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(),
                                      st if st is not None else stream()), "vj_cast_f32_bf16")
    return dst
