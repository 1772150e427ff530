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
from .workspace import TorchAlloc

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
def _segs(B, S):
    """Normalise the sequence layout of a token matrix: a list of (samples, tokens per sample) segments whose
    rows follow each other.  Everything but attention is row-wise, so differently masked copies of a batch
    (different kept-token counts) run through ONE set of LayerNorm / GEMM launches; only attention is launched
    per segment."""
    if isinstance(B, (list, tuple)):
        return list(B)
    return [(B, S)]


def _attn_fwd_segs(qkv, att, lse, segs, heads, hd, st):
    r0 = l0 = 0
    for (b, s) in segs:
        n, nl = b * s, b * heads * s
        ops.attn_fwd(qkv[r0:r0 + n], att[r0:r0 + n], lse[l0:l0 + nl], b, s, heads, hd, st)
        r0 += n
        l0 += nl


def _attn_bwd_segs(qkv, att, datt, lse, dqkv, segs, heads, hd, st, ws, rope):
    r0 = l0 = 0
    for (b, s) in segs:
        n, nl = b * s, b * heads * s
        mk = ws.mark()
        ops.attn_bwd(qkv[r0:r0 + n], att[r0:r0 + n], datt[r0:r0 + n], lse[l0:l0 + nl], dqkv[r0:r0 + n], b, s, heads, hd,
                     st, ws.tmp, rope[r0:r0 + n] if rope is not None else None)
        ws.release(mk)                          # one segment's attention scratch is dead once the call is enqueued
        r0 += n
        l0 += nl


def block_forward(h: BlockH, x, x_out, B, S, heads, hd, rope, save, st, ws):
    """x: residual stream [sum(B_i*S_i), D] (bf16 encoder / fp32 predictor); x_out: where the block output goes.
    B, S: samples and tokens per sample, or B = list of (B_i, S_i) segments (see _segs) and S ignored.
    save=True keeps everything backward needs in ws.act; otherwise temporaries live in ws.tmp (caller
    brackets the call with mark/release).  Returns saved tuple or None."""
    segs = _segs(B, S)
    M, D = x.shape
    Hm = h.hidden
    A = ws.act if save else ws.tmp
    mean1 = rstd1 = mean2 = rstd2 = hpre = None
    if save:
        stats = A((4, M), F32)
        mean1, rstd1, mean2, rstd2 = stats[0], stats[1], stats[2], stats[3]
        hpre = A((M, Hm), BF16)
    # saved LayerNorm outputs carry 8 pad columns of ones when the qkv / fc1 weight-gradient GEMMs can turn them into
    # the bias gradients (ops.bias_grad_pad); the forward GEMMs read the [:, :D] view
    pad = ops.bias_grad_pad(D, 3 * D, Hm) if save else 0
    ln1 = A((M, D + pad), BF16)
    ops.layernorm_fwd(x, h.n1w, h.n1b, ln1, mean1, rstd1, 1e-6, st)
    qkv = A((M, 3 * D), BF16)
    ops.gemm(ln1[:, :D], h.qkv_w, qkv, M, 3 * D, D, bias=h.qkv_b, rope=(rope, hd, D), st=st)     # qkv + fused 3-axis RoPE
    att = A((M, D), BF16)
    lse = A((sum(b * s for b, s in segs) * heads,), F32)
    _attn_fwd_segs(qkv, att, lse, segs, heads, hd, st)
    x1 = A((M, D), x.dtype)
    ops.gemm(att, h.proj_w, x1, M, D, D, bias=h.proj_b, residual=x, round_bf16=True, st=st)
    ln2 = A((M, D + pad), BF16)
    ops.layernorm_fwd(x1, h.n2w, h.n2b, ln2, mean2, rstd2, 1e-6, st)
    act = A((M, Hm), BF16)
    ops.gemm(ln2[:, :D], h.fc1_w, act, M, Hm, D, bias=h.fc1_b, gelu=True, round_bf16=True, aux_out=hpre, st=st)
    ops.gemm(act, h.fc2_w, x_out, M, D, Hm, bias=h.fc2_b, residual=x1, round_bf16=True, st=st)
    return (x, mean1, rstd1, ln1, qkv, att, lse, x1, mean2, rstd2, ln2, hpre, act) if save else None


def _as_bf16(t, st, ws):
    if t.dtype == BF16:
        return t
    out = ws.tmp(tuple(t.shape), BF16)
    ops.cast_f32_bf16(t, out, st)
    return out


def block_backward(h: BlockH, g: BlockG, saved, dx2, dx0, B, S, heads, hd, rope, st, ws):
    """dx2: gradient w.r.t. the block output [M, D] (dtype of the residual stream); dx0: output buffer for
    the gradient w.r.t. the block input (may not alias dx2).  Parameter gradients are ACCUMULATED into g
    (fp32).  Temporaries come from ws.tmp and are released before returning."""
    x, mean1, rstd1, ln1, qkv, att, lse, x1, mean2, rstd2, ln2, hpre, act = saved
    segs = _segs(B, S)
    M, D = x.shape
    Hm = h.hidden
    T = ws.tmp
    mk = ws.mark()
    d2 = _as_bf16(dx2, st, ws)
    # ---- MLP: x2 = x1 + fc2(gelu(fc1(LN2(x1))))
    dh = T((M, Hm), BF16)
    ops.gemm(d2, h.fc2_w, dh, M, Hm, D, b_mn=True, dgelu_aux=hpre, st=st)                 # dgrad fc2 * gelu'
    ops.gemm(d2, act, g.fc2_w, D, Hm, M, a_mn=True, b_mn=True, residual=g.fc2_w, st=st)   # wgrad fc2 (+=)
    dln2 = T((M, D), BF16)
    ops.gemm(dh, h.fc1_w, dln2, M, D, Hm, b_mn=True, st=st)                               # dgrad fc1
    padded = ln2.shape[1] > D                 # ones-columns present: the wgrad GEMM also yields the bias gradient
    if padded:
        ops.gemm(dh, ln2[:, :D], g.fc1_w, Hm, D, M, a_mn=True, b_mn=True, residual=g.fc1_w, bias_grad=g.fc1_b, st=st)
    else:
        ops.gemm(dh, ln2, g.fc1_w, Hm, D, M, a_mn=True, b_mn=True, residual=g.fc1_w, st=st)   # wgrad fc1
        ops.colsum(dh, g.fc1_b, True, st, T)
    dx1 = T((M, D), dx2.dtype)
    # norm2 backward also sums its dres = d(fc2 output) over the rows: the fc2 bias gradient
    ops.layernorm_bwd(dln2, x1, h.n2w, mean2, rstd2, dx1, dres=dx2, dgamma=g.n2w, dbeta=g.n2b, dbias=g.fc2_b, st=st, alloc=T)
    # ---- attention: x1 = x + proj(attn(rope(qkv(LN1(x)))))
    d1 = _as_bf16(dx1, st, ws)
    datt = T((M, D), BF16)
    ops.gemm(d1, h.proj_w, datt, M, D, D, b_mn=True, st=st)
    ops.gemm(d1, att, g.proj_w, D, D, M, a_mn=True, b_mn=True, residual=g.proj_w, st=st)
    dqkv = T((M, 3 * D), BF16)
    _attn_bwd_segs(qkv, att, datt, lse, dqkv, segs, heads, hd, st, ws, rope)              # + fused adjoint RoPE
    dln1 = T((M, D), BF16)
    ops.gemm(dqkv, h.qkv_w, dln1, M, D, 3 * D, b_mn=True, st=st)
    if padded:
        ops.gemm(dqkv, ln1[:, :D], g.qkv_w, 3 * D, D, M, a_mn=True, b_mn=True, residual=g.qkv_w, bias_grad=g.qkv_b, st=st)
    else:
        ops.gemm(dqkv, ln1, g.qkv_w, 3 * D, D, M, a_mn=True, b_mn=True, residual=g.qkv_w, st=st)
        ops.colsum(dqkv, g.qkv_b, True, st, T)
    # norm1 backward: dres = d(proj output) -> proj bias gradient
    ops.layernorm_bwd(dln1, x, h.n1w, mean1, rstd1, dx0, dres=dx1, dgamma=g.n1w, dbeta=g.n1b, dbias=g.proj_b, st=st, alloc=T)
    ws.release(mk)
    return dx0


def _run_blocks_forward(blocks, x, B, S, heads, hd, rope, save, st, ws, on_block=None):
    """Runs the block stack.  save=True: every block output is a fresh ws.act tensor (it is the next block's
    saved input).  save=False: two ping-pong residual buffers, per-block temporaries released immediately."""
    saved_all = []
    if save:
        for i, h in enumerate(blocks):
            x_out = ws.act(tuple(x.shape), x.dtype)
            saved_all.append(block_forward(h, x, x_out, B, S, heads, hd, rope, True, st, ws))
            x = x_out
            if on_block is not None:
                on_block(i, x)
        return x, saved_all
    pp = (ws.tmp(tuple(x.shape), x.dtype), ws.tmp(tuple(x.shape), x.dtype))
    for i, h in enumerate(blocks):
        x_out = pp[i & 1]
        mk = ws.mark()
        block_forward(h, x, x_out, B, S, heads, hd, rope, False, st, ws)
        ws.release(mk)
        x = x_out
        if on_block is not None:
            on_block(i, x)
    return x, None


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


def encoder_forward(rt: EncoderRT, clips, ids, grid_hw, save, ws=None, out_layers=None):
    """clips fp32 [B,C,T,H,W]; ids: None, int64 [B*reps, K] kept-token ids, or a LIST of int64 [B, K_i] (one entry
    per mask; the masked copies share every LayerNorm / GEMM launch, see _segs).
    Returns (out, saved): out fp32 [B*reps, S, D], or a list of [B, K_i, D] views for a list of ids -- or the list
    of normed intermediate outputs if out_layers.
    With an Arena `ws`, the output lives in ws.act (save) / ws.tmp and is only valid until the arena is reset."""
    st = ops.stream()
    B = clips.shape[0]
    dev = clips.device
    if ws is None:
        ws = TorchAlloc(dev)
    Hp, Wp = grid_hw
    D, heads, hd = rt.D, rt.heads, rt.hd
    A = ws.act if save else ws.tmp
    multi = isinstance(ids, (list, tuple))
    if multi:
        if out_layers is not None:
            raise NotImplementedError("vjepa2_b200: out_layers with a list of masks")
        segs = [(B, int(m.shape[1])) for m in ids]
        M = sum(b * s for b, s in segs)
        cols = A((M, rt.pe_k), BF16)
        rope = A((M, 2, hd), torch.float16)
        r0 = 0
        for m, (b, s) in zip(ids, segs):
            ops.im2col_tubelets(clips, m, rt.tubelet, rt.patch, st, out=cols[r0:r0 + b * s])
            ops.rope_table(m, b * s, s, Hp, Wp, hd, dev, st, out=rope[r0:r0 + b * s])
            r0 += b * s
        Bp, S = segs, None
    else:
        cols = ops.im2col_tubelets(clips, ids, rt.tubelet, rt.patch, st, A)
        M = cols.shape[0]
        if ids is not None:
            Bp, S = ids.shape
        else:
            Bp, S = B, M // B
        rope = ops.rope_table(ids, M, S, Hp, Wp, hd, dev, st, A)
    x = A((M, D), BF16)
    ops.gemm(cols, rt.pe_w, x, M, D, rt.pe_k, bias=rt.pe_b, st=st)
    outs = []

    def collect(i, xi):
        if out_layers is not None and i in out_layers:
            o = torch.empty(M, D, dtype=F32, device=dev)
            ops.layernorm_fwd(xi, rt.norm_w, rt.norm_b, o, None, None, 1e-6, st)
            outs.append(o.view(Bp, S, D))

    x, blocks_saved = _run_blocks_forward(rt.blocks, x, Bp, S, heads, hd, rope, save, st, ws,
                                          collect if out_layers is not None else None)
    if out_layers is not None:
        return outs, None
    out = A((M, D), F32)
    mean = rstd = None
    if save:
        mean = A((M,), F32)
        rstd = A((M,), F32)
    ops.layernorm_fwd(x, rt.norm_w, rt.norm_b, out, mean, rstd, 1e-6, st)
    saved = (cols, rope, blocks_saved, x, mean, rstd, Bp, S) if save else None
    if multi:
        views, r0 = [], 0
        for (b, s) in segs:
            views.append(out[r0:r0 + b * s].view(b, s, D))
            r0 += b * s
        return views, saved
    return out.view(Bp, S, D), saved


def encoder_backward(rt: EncoderRT, saved, dout, gbuf, ws=None, on_block_done=None):
    """dout: gradient w.r.t. the encoder output, [B', S, D] or the row-concatenated [M, D] of a multi-mask
    forward (bf16 or fp32).  Accumulates parameter gradients into gbuf (flat fp32).  The clip needs no gradient."""
    st = ops.stream()
    cols, rope, blocks_saved, x_last, mean, rstd, Bp, S = saved
    D, heads, hd = rt.D, rt.heads, rt.hd
    M = x_last.shape[0]
    if ws is None:
        ws = TorchAlloc(dout.device)
    g = rt.grads(gbuf)
    dy = dout.reshape(M, D)
    outer = ws.mark()
    pp = (ws.tmp((M, D), BF16), ws.tmp((M, D), BF16))
    dx = pp[0]
    ops.layernorm_bwd(dy, x_last, rt.norm_w, mean, rstd, dx, dres=None, dgamma=g["norm_w"], dbeta=g["norm_b"], st=st,
                      alloc=ws.tmp)
    if on_block_done is not None:
        on_block_done(len(rt.blocks))           # final norm params are done
    k = 0
    for i in range(len(rt.blocks) - 1, -1, -1):
        k ^= 1
        dx = block_backward(rt.blocks[i], g["blocks"][i], blocks_saved[i], dx, pp[k], Bp, S, heads, hd, rope, st, ws)
        blocks_saved[i] = None                  # release activations as we go (torch allocator path)
        if on_block_done is not None:
            on_block_done(i)
    ops.gemm(dx, cols, g["pe_w"], D, rt.pe_k, M, a_mn=True, b_mn=True, residual=g["pe_w"], st=st)
    ops.colsum(dx, g["pe_b"], True, st, ws.tmp)
    ws.release(outer)
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


def predictor_forward(rt: PredictorRT, z, masks_x, masks_y, mask_index, save, ws=None):
    """z: context-encoder output [B, Kc, D_in] (fp32 or bf16); masks_x [B,Kc], masks_y [B,Kp] int64 -- or three
    equally long LISTS of those (one entry per mask; the entries of z must be consecutive row blocks of one
    matrix, as encoder_forward returns them).  Returns (pred bf16 [B, Kp, D_in] or a list of such views, saved)."""
    st = ops.stream()
    multi = isinstance(z, (list, tuple))
    zs = list(z) if multi else [z]
    mxs = list(masks_x) if multi else [masks_x]
    mys = list(masks_y) if multi else [masks_y]
    dev = zs[0].device
    if ws is None:
        ws = TorchAlloc(dev)
    B, Din = zs[0].shape[0], zs[0].shape[2]
    Kcs = [int(t.shape[1]) for t in zs]
    Kps = [int(t.shape[1]) for t in mys]
    Mc, Mp = B * sum(Kcs), B * sum(Kps)
    segs = [(B, kc + kp) for kc, kp in zip(Kcs, Kps)]
    Ms = sum(b * s for b, s in segs)
    D, heads, hd = rt.D, rt.heads, rt.hd
    A = ws.act if save else ws.tmp
    if len(zs) == 1:
        z2 = zs[0].reshape(Mc, Din)
    else:
        step = zs[0].element_size() * Din
        p0 = zs[0].data_ptr()
        for t, kc_prev in zip(zs[1:], Kcs[:-1]):
            p0 += B * kc_prev * step
            if t.data_ptr() != p0 or not t.is_contiguous():
                raise ValueError("vjepa2_b200: predictor_forward expects the z entries to be consecutive row blocks")
        z2 = torch.as_strided(zs[0], (Mc, Din), (Din, 1))
    if z2.dtype == BF16:
        z16 = z2
    else:
        z16 = A((Mc, Din), BF16)
        ops.cast_f32_bf16(z2, z16, st)
    emb = A((Mc, D), BF16)
    ops.gemm(z16, rt.embed_w, emb, Mc, D, Din, bias=rt.embed_b, st=st)
    mi = mask_index % len(rt.mask_tokens)
    x = A((Ms, D), F32)
    rope = A((Ms, 2, hd), torch.float16)
    idx = []
    c0 = s0 = 0
    for mx, my, kc, kp, (b, sq) in zip(mxs, mys, Kcs, Kps, segs):
        ids_sorted, asm_idx, tgt_pos, ctx_pos, seq_to_tgt = ops.pred_indices(mx, my, st, A)
        ops.gather_rows(emb[c0:c0 + b * kc], x[s0:s0 + b * sq], asm_idx, fill=rt.mask_tokens[mi], st=st)
        ops.rope_table(ids_sorted, b * sq, sq, rt.grid, rt.grid, hd, dev, st, out=rope[s0:s0 + b * sq])
        idx.append((tgt_pos, ctx_pos, seq_to_tgt))
        c0 += b * kc
        s0 += b * sq
    x, blocks_saved = _run_blocks_forward(rt.blocks, x, segs, None, heads, hd, rope, save, st, ws)
    # LayerNorm is row-wise, so norm(x)[targets] == norm(x[targets]) (predictor.py:233,240-242)
    xg = A((Mp, D), F32)
    s0 = t0 = 0
    for (tgt_pos, _, _), kp, (b, sq) in zip(idx, Kps, segs):
        ops.gather_rows(x[s0:s0 + b * sq], xg[t0:t0 + b * kp], tgt_pos, st=st)
        s0 += b * sq
        t0 += b * kp
    y16 = A((Mp, D), BF16)
    mean = rstd = None
    if save:
        mean = A((Mp,), F32)
        rstd = A((Mp,), F32)
    ops.layernorm_fwd(xg, rt.norm_w, rt.norm_b, y16, mean, rstd, 1e-6, st)
    out = A((Mp, Din), BF16)
    ops.gemm(y16, rt.proj_w, out, Mp, Din, D, bias=rt.proj_b, st=st)
    saved = None
    if save:
        saved = (z16, rope, blocks_saved, xg, mean, rstd, y16, idx, mi, B, Kcs, Kps)
    if multi:
        views, t0 = [], 0
        for kp in Kps:
            views.append(out[t0:t0 + B * kp].view(B, kp, Din))
            t0 += B * kp
        return views, saved
    return out.view(B, Kps[0], Din), saved


def predictor_backward(rt: PredictorRT, saved, dout, gbuf, ws=None, dz_out=None):
    """dout bf16 [B, Kp, D_in] (or the row-concatenated [sum B*Kp_i, D_in] of a multi-mask forward).
    Accumulates parameter grads into gbuf; returns d(z) bf16 [B, Kc, D_in] ([sum B*Kc_i, D_in] for multi-mask),
    written to dz_out if given, else allocated from ws.act so it outlives this call's temporaries."""
    st = ops.stream()
    z16, rope, blocks_saved, xg, mean, rstd, y16, idx, mi, B, Kcs, Kps = saved
    dev = z16.device
    if ws is None:
        ws = TorchAlloc(dev)
    segs = [(B, kc + kp) for kc, kp in zip(Kcs, Kps)]
    Mc, Mp = B * sum(Kcs), B * sum(Kps)
    Ms = sum(b * s for b, s in segs)
    D, heads, hd, Din = rt.D, rt.heads, rt.hd, rt.D_in
    g = rt.grads(gbuf)
    dz = dz_out if dz_out is not None else ws.act((Mc, Din), BF16)
    T = ws.tmp
    outer = ws.mark()
    do = _as_bf16(dout.reshape(Mp, Din), st, ws)
    # predictor_proj
    dy16 = T((Mp, D), BF16)
    ops.gemm(do, rt.proj_w, dy16, Mp, D, Din, b_mn=True, st=st)
    ops.gemm(do, y16, g["proj_w"], Din, D, Mp, a_mn=True, b_mn=True, residual=g["proj_w"], st=st)
    ops.colsum(do, g["proj_b"], True, st, T)
    # predictor_norm on the target rows, then scatter back into the sorted sequence
    dxg = T((Mp, D), F32)
    ops.layernorm_bwd(dy16, xg, rt.norm_w, mean, rstd, dxg, dres=None, dgamma=g["norm_w"], dbeta=g["norm_b"], st=st,
                      alloc=T)
    pp = (T((Ms, D), F32), T((Ms, D), F32))
    dx = pp[0]
    s0 = t0 = 0
    for (_, _, seq_to_tgt), kp, (b, sq) in zip(idx, Kps, segs):
        ops.gather_rows(dxg[t0:t0 + b * kp], dx[s0:s0 + b * sq], seq_to_tgt, fill=None, st=st)   # context rows get zeros
        s0 += b * sq
        t0 += b * kp
    k = 0
    for i in range(len(rt.blocks) - 1, -1, -1):
        k ^= 1
        dx = block_backward(rt.blocks[i], g["blocks"][i], blocks_saved[i], dx, pp[k], segs, None, heads, hd, rope, st, ws)
        blocks_saved[i] = None
    # mask token: sum of the gradients of every target slot (predictor.py:195-197); predictor_embed inputs
    dtg = T((Mp, D), F32)
    demb = T((Mc, D), BF16)
    s0 = t0 = c0 = 0
    for (tgt_pos, ctx_pos, _), kc, kp, (b, sq) in zip(idx, Kcs, Kps, segs):
        ops.gather_rows(dx[s0:s0 + b * sq], dtg[t0:t0 + b * kp], tgt_pos, st=st)
        ops.gather_rows(dx[s0:s0 + b * sq], demb[c0:c0 + b * kc], ctx_pos, st=st)
        s0 += b * sq
        t0 += b * kp
        c0 += b * kc
    ops.colsum(dtg, g["mask_tokens"][mi], True, st, T)
    # predictor_embed
    ops.gemm(demb, rt.embed_w, dz, Mc, Din, D, b_mn=True, st=st)
    ops.gemm(demb, z16, g["embed_w"], D, Din, Mc, a_mn=True, b_mn=True, residual=g["embed_w"], st=st)
    ops.colsum(demb, g["embed_b"], True, st, T)
    ws.release(outer)
    if len(Kcs) == 1:
        return dz.view(B, Kcs[0], Din)
    return dz.view(Mc, Din)
