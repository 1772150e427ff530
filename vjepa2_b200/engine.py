"""Explicit forward / backward pipelines of the encoder and predictor on the sm_100a kernels.

This is the host-side "runtime" of the hot path: it sequences the C-ABI kernels (vjepa2_b200.ops) for
VisionTransformer.forward (vision_transformer.py:161-213), Block.forward (modules.py:556-563),
RoPEAttention.forward (modules.py:326-382), MLP.forward (modules.py:77-83) and
VisionTransformerPredictor.forward (predictor.py:166-246), and their hand-derived backward passes.
No autograd graph is built here; vjepa2_b200.vision_transformer / predictor wrap these functions in
torch.autograd.Function for drop-in use, and vjepa2_b200.train calls them directly.

dtype flow follows bf16 autocast in the reference (SURVEY 8a notes): bf16 residual stream in the
encoder, fp32 residual stream in the predictor, LayerNorm statistics in fp32, bf16 GEMM operands with
fp32 accumulation, fp32 final-norm output, fp32 parameter gradients.
"""
from __future__ import annotations

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


# ------------------------------------------------------------------------------------------------
# handles: raw views the kernels consume, built once per (flat store, grad buffer)
# ------------------------------------------------------------------------------------------------
class BlockH:
    __slots__ = ("n1w", "n1b", "qkv_w", "qkv_b", "proj_w", "proj_b", "n2w", "n2b", "fc1_w", "fc1_b", "fc2_w",
                 "fc2_b", "hidden", "params")

    def __init__(self, blk, fs):
        a, m = blk.attn, blk.mlp
        self.n1w, self.n1b = blk.norm1.weight.data, blk.norm1.bias.data
        self.qkv_w, self.qkv_b = fs.w16(a.qkv.weight), a.qkv.bias.data
        self.proj_w, self.proj_b = fs.w16(a.proj.weight), a.proj.bias.data
        self.n2w, self.n2b = blk.norm2.weight.data, blk.norm2.bias.data
        self.fc1_w, self.fc1_b = fs.w16(m.fc1.weight), m.fc1.bias.data
        self.fc2_w, self.fc2_b = fs.w16(m.fc2.weight), m.fc2.bias.data
        self.hidden = m.fc1.weight.shape[0]
        self.params = list(blk.parameters())


class BlockG:
    __slots__ = ("n1w", "n1b", "qkv_w", "qkv_b", "proj_w", "proj_b", "n2w", "n2b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")

    def __init__(self, blk, fs, gbuf):
        a, m = blk.attn, blk.mlp
        gv = lambda p: fs.grad_view(gbuf, p)  # noqa: E731
        self.n1w, self.n1b = gv(blk.norm1.weight), gv(blk.norm1.bias)
        self.qkv_w, self.qkv_b = gv(a.qkv.weight), gv(a.qkv.bias)
        self.proj_w, self.proj_b = gv(a.proj.weight), gv(a.proj.bias)
        self.n2w, self.n2b = gv(blk.norm2.weight), gv(blk.norm2.bias)
        self.fc1_w, self.fc1_b = gv(m.fc1.weight), gv(m.fc1.bias)
        self.fc2_w, self.fc2_b = gv(m.fc2.weight), gv(m.fc2.bias)


# ------------------------------------------------------------------------------------------------
# one transformer block
# ------------------------------------------------------------------------------------------------
def block_forward(h: BlockH, x, B, S, heads, hd, cos, sin, save, st):
    """x: residual stream [B*S, D] (bf16 encoder / fp32 predictor).  Returns (x_out, saved)."""
    M, D = x.shape
    dev = x.device
    Hm = h.hidden
    mean1 = rstd1 = mean2 = rstd2 = hpre = None
    if save:
        stats = torch.empty(4, M, dtype=F32, device=dev)
        mean1, rstd1, mean2, rstd2 = stats[0], stats[1], stats[2], stats[3]
        hpre = torch.empty(M, Hm, dtype=BF16, device=dev)
    ln1 = torch.empty(M, D, dtype=BF16, device=dev)
    ops.layernorm_fwd(x, h.n1w, h.n1b, ln1, mean1, rstd1, 1e-6, st)
    qkv = torch.empty(M, 3 * D, dtype=BF16, device=dev)
    ops.gemm(ln1, h.qkv_w, qkv, M, 3 * D, D, bias=h.qkv_b, st=st)
    ops.rope_apply(qkv, D, heads, hd, cos, sin, False, st)
    att = torch.empty(M, D, dtype=BF16, device=dev)
    lse = torch.empty(B * heads * S, dtype=F32, device=dev)
    ops.attn_fwd(qkv, att, lse, B, S, heads, hd, st)
    x1 = torch.empty_like(x)
    ops.gemm(att, h.proj_w, x1, M, D, D, bias=h.proj_b, residual=x, round_bf16=True, st=st)
    ln2 = torch.empty(M, D, dtype=BF16, device=dev)
    ops.layernorm_fwd(x1, h.n2w, h.n2b, ln2, mean2, rstd2, 1e-6, st)
    act = torch.empty(M, Hm, dtype=BF16, device=dev)
    ops.gemm(ln2, h.fc1_w, act, M, Hm, D, bias=h.fc1_b, gelu=True, round_bf16=True, aux_out=hpre, st=st)
    x2 = torch.empty_like(x)
    ops.gemm(act, h.fc2_w, x2, M, D, Hm, bias=h.fc2_b, residual=x1, round_bf16=True, st=st)
    saved = (x, mean1, rstd1, ln1, qkv, att, lse, x1, mean2, rstd2, ln2, hpre, act) if save else None
    return x2, saved


def _as_bf16(t, st):
    if t.dtype == BF16:
        return t
    out = torch.empty(t.shape, dtype=BF16, device=t.device)
    ops.cast_f32_bf16(t, out, st)
    return out


def block_backward(h: BlockH, g: BlockG, saved, dx2, B, S, heads, hd, cos, sin, st):
    """dx2: gradient w.r.t. the block output [M, D] (dtype of the residual stream).  Parameter gradients
    are ACCUMULATED into g (fp32).  Returns the gradient w.r.t. the block input."""
    x, mean1, rstd1, ln1, qkv, att, lse, x1, mean2, rstd2, ln2, hpre, act = saved
    M, D = x.shape
    dev = x.device
    Hm = h.hidden
    d2 = _as_bf16(dx2, st)
    # ---- MLP: x2 = x1 + fc2(gelu(fc1(LN2(x1))))
    dh = torch.empty(M, Hm, dtype=BF16, device=dev)
    ops.gemm(d2, h.fc2_w, dh, M, Hm, D, b_mn=True, dgelu_aux=hpre, st=st)                 # dgrad fc2 * gelu'
    ops.gemm(d2, act, g.fc2_w, D, Hm, M, a_mn=True, b_mn=True, residual=g.fc2_w, st=st)   # wgrad fc2 (+=)
    ops.colsum(d2, g.fc2_b, True, st)
    dln2 = torch.empty(M, D, dtype=BF16, device=dev)
    ops.gemm(dh, h.fc1_w, dln2, M, D, Hm, b_mn=True, st=st)                               # dgrad fc1
    ops.gemm(dh, ln2, g.fc1_w, Hm, D, M, a_mn=True, b_mn=True, residual=g.fc1_w, st=st)   # wgrad fc1
    ops.colsum(dh, g.fc1_b, True, st)
    dx1 = torch.empty_like(dx2)
    ops.layernorm_bwd(dln2, x1, h.n2w, mean2, rstd2, dx1, dres=dx2, dgamma=g.n2w, dbeta=g.n2b, st=st)
    # ---- attention: x1 = x + proj(attn(rope(qkv(LN1(x)))))
    d1 = _as_bf16(dx1, st)
    datt = torch.empty(M, D, dtype=BF16, device=dev)
    ops.gemm(d1, h.proj_w, datt, M, D, D, b_mn=True, st=st)
    ops.gemm(d1, att, g.proj_w, D, D, M, a_mn=True, b_mn=True, residual=g.proj_w, st=st)
    ops.colsum(d1, g.proj_b, True, st)
    dqkv = torch.empty(M, 3 * D, dtype=BF16, device=dev)
    ops.attn_bwd(qkv, att, datt, lse, dqkv, B, S, heads, hd, st)
    ops.rope_apply(dqkv, D, heads, hd, cos, sin, True, st)
    dln1 = torch.empty(M, D, dtype=BF16, device=dev)
    ops.gemm(dqkv, h.qkv_w, dln1, M, D, 3 * D, b_mn=True, st=st)
    ops.gemm(dqkv, ln1, g.qkv_w, 3 * D, D, M, a_mn=True, b_mn=True, residual=g.qkv_w, st=st)
    ops.colsum(dqkv, g.qkv_b, True, st)
    dx0 = torch.empty_like(dx2)
    ops.layernorm_bwd(dln1, x, h.n1w, mean1, rstd1, dx0, dres=dx1, dgamma=g.n1w, dbeta=g.n1b, st=st)
    return dx0


# ------------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------------
class EncoderRT:
    """Kernel-facing view of a VisionTransformer bound to its flat store."""

    def __init__(self, model, fs):
        self.fs = fs
        self.D = model.embed_dim
        self.heads = model.num_heads
        self.hd = self.D // self.heads
        self.patch, self.tubelet = model.patch_size, model.tubelet_size
        pw = model.patch_embed.proj.weight
        self.pe_k = pw[0].numel()
        self.pe_w = fs.w16(pw).view(self.D, self.pe_k)
        self.pe_b = model.patch_embed.proj.bias.data
        self.pe_params = [pw, model.patch_embed.proj.bias]
        self.blocks = [BlockH(b, fs) for b in model.blocks]
        self.norm_w, self.norm_b = model.norm.weight.data, model.norm.bias.data
        self.norm_params = [model.norm.weight, model.norm.bias]
        self.model = model
        self._grads = {}

    def __deepcopy__(self, memo):
        return None

    def grads(self, gbuf):
        key = gbuf.data_ptr()
        g = self._grads.get(key)
        if g is None:
            m, fs = self.model, self.fs
            g = dict(blocks=[BlockG(b, fs, gbuf) for b in m.blocks],
                     pe_w=fs.grad_view(gbuf, m.patch_embed.proj.weight).view(self.D, self.pe_k),
                     pe_b=fs.grad_view(gbuf, m.patch_embed.proj.bias),
                     norm_w=fs.grad_view(gbuf, m.norm.weight), norm_b=fs.grad_view(gbuf, m.norm.bias))
            if len(self._grads) > 4:
                self._grads.clear()
            self._grads[key] = g
        return g


def encoder_forward(rt: EncoderRT, clips, ids, grid_hw, save, out_layers=None):
    """clips fp32 [B,C,T,H,W]; ids None or int64 [B*reps, K] kept-token ids.
    Returns (out fp32 [B*reps, S, D], saved) -- or the list of normed intermediate outputs if out_layers."""
    st = ops.stream()
    B = clips.shape[0]
    dev = clips.device
    Hp, Wp = grid_hw
    D, heads, hd = rt.D, rt.heads, rt.hd
    cols = ops.im2col_tubelets(clips, ids, rt.tubelet, rt.patch, st)
    M = cols.shape[0]
    if ids is not None:
        Bp, S = ids.shape
    else:
        Bp, S = B, M // B
    x = torch.empty(M, D, dtype=BF16, device=dev)
    ops.gemm(cols, rt.pe_w, x, M, D, rt.pe_k, bias=rt.pe_b, st=st)
    cos, sin = ops.rope_table(ids, M, S, Hp, Wp, hd, dev, st)
    blocks_saved = []
    outs = []
    for i, h in enumerate(rt.blocks):
        x, sv = block_forward(h, x, Bp, S, heads, hd, cos, sin, save, st)
        if save:
            blocks_saved.append(sv)
        if out_layers is not None and i in out_layers:
            o = torch.empty(M, D, dtype=F32, device=dev)
            ops.layernorm_fwd(x, rt.norm_w, rt.norm_b, o, None, None, 1e-6, st)
            outs.append(o.view(Bp, S, D))
    if out_layers is not None:
        return outs, None
    out = torch.empty(M, D, dtype=F32, device=dev)
    mean = rstd = None
    if save:
        mean = torch.empty(M, dtype=F32, device=dev)
        rstd = torch.empty(M, dtype=F32, device=dev)
    ops.layernorm_fwd(x, rt.norm_w, rt.norm_b, out, mean, rstd, 1e-6, st)
    saved = (cols, cos, sin, blocks_saved, x, mean, rstd, Bp, S) if save else None
    return out.view(Bp, S, D), saved


def encoder_backward(rt: EncoderRT, saved, dout, gbuf, on_block_done=None):
    """dout: gradient w.r.t. the encoder output [B', S, D] (bf16 or fp32).  Accumulates parameter
    gradients into gbuf (flat fp32).  The input clip needs no gradient."""
    st = ops.stream()
    cols, cos, sin, blocks_saved, x_last, mean, rstd, Bp, S = saved
    D, heads, hd = rt.D, rt.heads, rt.hd
    M = Bp * S
    g = rt.grads(gbuf)
    dy = dout.reshape(M, D)
    dx = torch.empty(M, D, dtype=BF16, device=dy.device)
    ops.layernorm_bwd(dy, x_last, rt.norm_w, mean, rstd, dx, dres=None, dgamma=g["norm_w"], dbeta=g["norm_b"], st=st)
    if on_block_done is not None:
        on_block_done(len(rt.blocks))           # final norm params are done
    for i in range(len(rt.blocks) - 1, -1, -1):
        dx = block_backward(rt.blocks[i], g["blocks"][i], blocks_saved[i], dx, Bp, S, heads, hd, cos, sin, st)
        blocks_saved[i] = None                  # release activations as we go
        if on_block_done is not None:
            on_block_done(i)
    ops.gemm(dx, cols, g["pe_w"], D, rt.pe_k, M, a_mn=True, b_mn=True, residual=g["pe_w"], st=st)
    ops.colsum(dx, g["pe_b"], True, st)
    if on_block_done is not None:
        on_block_done(-1)                       # patch-embed params are done


# ------------------------------------------------------------------------------------------------
# predictor
# ------------------------------------------------------------------------------------------------
class PredictorRT:
    def __init__(self, model, fs):
        self.fs = fs
        self.model = model
        self.D_in = model.predictor_embed.weight.shape[1]
        self.D = model.predictor_embed.weight.shape[0]
        self.heads = model.predictor_blocks[0].attn.num_heads
        self.hd = self.D // self.heads
        self.grid = model.grid_height
        self.embed_w, self.embed_b = fs.w16(model.predictor_embed.weight), model.predictor_embed.bias.data
        self.proj_w, self.proj_b = fs.w16(model.predictor_proj.weight), model.predictor_proj.bias.data
        self.norm_w, self.norm_b = model.predictor_norm.weight.data, model.predictor_norm.bias.data
        self.mask_tokens = [p.data.view(-1) for p in model.mask_tokens]
        self.blocks = [BlockH(b, fs) for b in model.predictor_blocks]
        self._grads = {}

    def __deepcopy__(self, memo):
        return None

    def grads(self, gbuf):
        key = gbuf.data_ptr()
        g = self._grads.get(key)
        if g is None:
            m, fs = self.model, self.fs
            gv = lambda p: fs.grad_view(gbuf, p)  # noqa: E731
            g = dict(blocks=[BlockG(b, fs, gbuf) for b in m.predictor_blocks],
                     embed_w=gv(m.predictor_embed.weight), embed_b=gv(m.predictor_embed.bias),
                     proj_w=gv(m.predictor_proj.weight), proj_b=gv(m.predictor_proj.bias),
                     norm_w=gv(m.predictor_norm.weight), norm_b=gv(m.predictor_norm.bias),
                     mask_tokens=[gv(p).view(-1) for p in m.mask_tokens])
            if len(self._grads) > 4:
                self._grads.clear()
            self._grads[key] = g
        return g


def predictor_forward(rt: PredictorRT, z, masks_x, masks_y, mask_index, save):
    """z: context-encoder output [B, Kc, D_in] (fp32 or bf16); masks_x [B,Kc], masks_y [B,Kp] int64.
    Returns (pred bf16 [B, Kp, D_in], saved)."""
    st = ops.stream()
    dev = z.device
    B, Kc, Din = z.shape
    Kp = masks_y.shape[1]
    S = Kc + Kp
    D, heads, hd = rt.D, rt.heads, rt.hd
    z16 = _as_bf16(z.reshape(B * Kc, Din), st)
    emb = torch.empty(B * Kc, D, dtype=BF16, device=dev)
    ops.gemm(z16, rt.embed_w, emb, B * Kc, D, Din, bias=rt.embed_b, st=st)
    ids_sorted, asm_idx, tgt_pos, ctx_pos, seq_to_tgt = ops.pred_indices(masks_x, masks_y, st)
    mi = mask_index % len(rt.mask_tokens)
    x = torch.empty(B * S, D, dtype=F32, device=dev)
    ops.gather_rows(emb, x, asm_idx, fill=rt.mask_tokens[mi], st=st)
    cos, sin = ops.rope_table(ids_sorted, B * S, S, rt.grid, rt.grid, hd, dev, st)
    blocks_saved = []
    for h in rt.blocks:
        x, sv = block_forward(h, x, B, S, heads, hd, cos, sin, save, st)
        if save:
            blocks_saved.append(sv)
    # LayerNorm is row-wise, so norm(x)[targets] == norm(x[targets]) (predictor.py:233,240-242)
    xg = torch.empty(B * Kp, D, dtype=F32, device=dev)
    ops.gather_rows(x, xg, tgt_pos, st=st)
    y16 = torch.empty(B * Kp, D, dtype=BF16, device=dev)
    mean = rstd = None
    if save:
        mean = torch.empty(B * Kp, dtype=F32, device=dev)
        rstd = torch.empty(B * Kp, dtype=F32, device=dev)
    ops.layernorm_fwd(xg, rt.norm_w, rt.norm_b, y16, mean, rstd, 1e-6, st)
    out = torch.empty(B * Kp, Din, dtype=BF16, device=dev)
    ops.gemm(y16, rt.proj_w, out, B * Kp, Din, D, bias=rt.proj_b, st=st)
    saved = None
    if save:
        saved = (z16, cos, sin, blocks_saved, xg, mean, rstd, y16, tgt_pos, ctx_pos, seq_to_tgt, mi, B, Kc, Kp)
    return out.view(B, Kp, Din), saved


def predictor_backward(rt: PredictorRT, saved, dout, gbuf):
    """dout bf16 [B, Kp, D_in].  Accumulates parameter grads into gbuf; returns d(z) bf16 [B, Kc, D_in]."""
    st = ops.stream()
    z16, cos, sin, blocks_saved, xg, mean, rstd, y16, tgt_pos, ctx_pos, seq_to_tgt, mi, B, Kc, Kp = saved
    dev = z16.device
    S = Kc + Kp
    D, heads, hd, Din = rt.D, rt.heads, rt.hd, rt.D_in
    g = rt.grads(gbuf)
    do = _as_bf16(dout.reshape(B * Kp, Din), st)
    # predictor_proj
    dy16 = torch.empty(B * Kp, D, dtype=BF16, device=dev)
    ops.gemm(do, rt.proj_w, dy16, B * Kp, D, Din, b_mn=True, st=st)
    ops.gemm(do, y16, g["proj_w"], Din, D, B * Kp, a_mn=True, b_mn=True, residual=g["proj_w"], st=st)
    ops.colsum(do, g["proj_b"], True, st)
    # predictor_norm on the target rows, then scatter back into the sorted sequence
    dxg = torch.empty(B * Kp, D, dtype=F32, device=dev)
    ops.layernorm_bwd(dy16, xg, rt.norm_w, mean, rstd, dxg, dres=None, dgamma=g["norm_w"], dbeta=g["norm_b"], st=st)
    dx = torch.empty(B * S, D, dtype=F32, device=dev)
    ops.gather_rows(dxg, dx, seq_to_tgt, fill=None, st=st)          # context rows get zeros
    for i in range(len(rt.blocks) - 1, -1, -1):
        dx = block_backward(rt.blocks[i], g["blocks"][i], blocks_saved[i], dx, B, S, heads, hd, cos, sin, st)
        blocks_saved[i] = None
    # mask token: sum of the gradients of every target slot (predictor.py:195-197)
    dtg = torch.empty(B * Kp, D, dtype=F32, device=dev)
    ops.gather_rows(dx, dtg, tgt_pos, st=st)
    ops.colsum(dtg, g["mask_tokens"][mi], True, st)
    # predictor_embed
    demb = torch.empty(B * Kc, D, dtype=BF16, device=dev)
    ops.gather_rows(dx, demb, ctx_pos, st=st)
    dz = torch.empty(B * Kc, Din, dtype=BF16, device=dev)
    ops.gemm(demb, rt.embed_w, dz, B * Kc, Din, D, b_mn=True, st=st)
    ops.gemm(demb, z16, g["embed_w"], D, Din, B * Kc, a_mn=True, b_mn=True, residual=g["embed_w"], st=st)
    ops.colsum(demb, g["embed_b"], True, st)
    return dz.view(B, Kc, Din)
