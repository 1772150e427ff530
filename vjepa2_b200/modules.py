"""Parameter containers mirroring src/models/utils/modules.py of the reference (MLP :67-83,
RoPEAttention :261-382, Block :500-563) and src/models/utils/patch_embed.py (PatchEmbed3D :26-52).

The classes keep the reference's module tree and parameter names (`norm1`, `attn.qkv`, `attn.proj`,
`norm2`, `mlp.fc1`, `mlp.fc2`, `patch_embed.proj`) so checkpoints load unchanged.  Inside a
VisionTransformer / VisionTransformerPredictor they are weight containers: the owning model drives
vjepa2_b200.engine over its flat parameter store.  Called directly (the reference's
`blk(x, mask=..., T=..., H_patches=..., W_patches=...)`, `attn(x, mask=...)`, `mlp(x)`, `patch_embed(x)`
signatures) each `forward` runs the SAME kernels for that one module, with its own cached bf16 operand copies:
`Block.forward` is differentiable (one autograd node, engine.block_forward / block_backward); `RoPEAttention`,
`MLP` and `PatchEmbed3D` are forward-only (torch.no_grad) -- training goes through Block or the owning model.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


BF16, F32 = torch.bfloat16, torch.float32


class _KernelModule(nn.Module):
    """bf16 operand copies of the module's own fp32 parameters, refreshed when a parameter is rewritten or moved."""

    def w16(self, p):
        cache = self.__dict__.setdefault("_w16_cache", {})
        stamp = (p.data_ptr(), p._version)
        ent = cache.get(id(p))
        if ent is None or ent[0] != stamp:
            from . import ops
            if not p.is_cuda:
                raise RuntimeError("vjepa2_b200: the module must be on a CUDA device (there is no CPU path)")
            t = torch.empty(p.shape, dtype=BF16, device=p.device)
            ops.cast_f32_bf16(p.data.contiguous(), t)
            ent = (stamp, t)
            cache[id(p)] = ent
        return ent[1]

    def grad_view(self, gbuf, p):          # engine.BlockG protocol: gbuf is a dict id(param) -> fp32 gradient tensor
        return gbuf[id(p)]

    def _no_autograd(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(q.requires_grad for q in self.parameters())):
            raise NotImplementedError(
                f"vjepa2_b200: {type(self).__name__}.forward is forward-only; call it under torch.no_grad().  Gradients "
                "flow through Block.forward or the owning VisionTransformer / VisionTransformerPredictor.")


def _rows(x):
    if x.dim() != 3:
        raise ValueError("vjepa2_b200: expected [B, N, C] tokens")
    B, N, C = x.shape
    return x.contiguous().view(B * N, C), B, N


def _rope_for(attn, mask, B, N, T, H_patches, W_patches, device, st):
    """Token ids -> RoPE table, with the reference's defaults (modules.py:311-341): ids = arange when mask is None, the
    row / column split uses grid_size unless H_patches and W_patches are given."""
    from . import ops
    Hp = attn.grid_size if H_patches is None or W_patches is None else H_patches
    Wp = attn.grid_size if H_patches is None or W_patches is None else W_patches
    ids = None
    if mask is not None:
        ids = mask.to(device=device, dtype=torch.int64).contiguous()
        if tuple(ids.shape) != (B, N):
            raise ValueError("vjepa2_b200: mask must be [B, N] token ids")
    elif T is not None and H_patches is not None and W_patches is not None and int(T * H_patches * W_patches) != N:
        raise ValueError("vjepa2_b200: T*H_patches*W_patches must equal the sequence length when mask is None")
    return ops.rope_table(ids, B * N, N, int(Hp), int(Wp), attn.head_dim, device, st)


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    """Same sampling law as src/utils/tensors.py:14-47 (inverse-CDF truncated normal, absolute bounds)."""
    def norm_cdf(x):
        return (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0

    with torch.no_grad():
        lo, up = norm_cdf((a - mean) / std), norm_cdf((b - mean) / std)
        tensor.uniform_(2 * lo - 1, 2 * up - 1)
        tensor.erfinv_()
        tensor.mul_(std * math.sqrt(2.0))
        tensor.add_(mean)
        tensor.clamp_(min=a, max=b)
    return tensor


class MLP(_KernelModule):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        if act_layer is not nn.GELU or drop != 0.0:
            raise NotImplementedError("vjepa2_b200: MLP supports exact GELU and drop=0 (the pre-training configs)")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        """modules.py:77-83 under bf16 autocast: fc1 -> exact GELU -> fc2, bf16 in / out (no residual)."""
        from . import ops
        self._no_autograd(x)
        x2, B, N = _rows(x)
        st = ops.stream()
        M, D, Hm, Do = x2.shape[0], self.fc1.in_features, self.fc1.out_features, self.fc2.out_features
        a = x2 if x2.dtype == BF16 else ops.cast_f32_bf16(x2, torch.empty(M, D, dtype=BF16, device=x.device), st)
        h = torch.empty(M, Hm, dtype=BF16, device=x.device)
        ops.gemm(a, self.w16(self.fc1.weight), h, M, Hm, D, bias=self.fc1.bias.data, gelu=True, round_bf16=True, st=st)
        y = torch.empty(M, Do, dtype=BF16, device=x.device)
        ops.gemm(h, self.w16(self.fc2.weight), y, M, Do, Hm, bias=self.fc2.bias.data, st=st)
        return y.view(B, N, Do)


class RoPEAttention(_KernelModule):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0, use_sdpa=True,
                 grid_size=14, is_causal=False):
        super().__init__()
        if not qkv_bias or qk_scale is not None or attn_drop != 0.0 or proj_drop != 0.0 or is_causal:
            raise NotImplementedError("vjepa2_b200: RoPEAttention supports qkv_bias=True, default scale, no dropout, "
                                      "non-causal (the pre-training configs)")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        if self.head_dim not in (32, 64, 80):
            raise NotImplementedError(f"vjepa2_b200: head_dim {self.head_dim} not supported (32, 64 and 80 are)")
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop_prob = proj_drop
        self.proj_drop = nn.Dropout(proj_drop)
        self.use_sdpa = use_sdpa
        self.d_dim = self.h_dim = self.w_dim = int(2 * ((self.head_dim // 3) // 2))
        self.grid_size = grid_size
        self.is_causal = is_causal

    def forward(self, x, mask=None, attn_mask=None, T=None, H_patches=None, W_patches=None):
        """modules.py:326-382 under bf16 autocast: qkv Linear (+ fused 3-axis RoPE on q, k) -> SDPA -> proj, bf16 out."""
        from . import ops
        if attn_mask is not None:
            raise NotImplementedError("vjepa2_b200: attn_mask is not supported (None in every pre-training call)")
        self._no_autograd(x)
        x2, B, N = _rows(x)
        st = ops.stream()
        M, D, hd = x2.shape[0], self.qkv.in_features, self.head_dim
        rope = _rope_for(self, mask, B, N, T, H_patches, W_patches, x.device, st)
        a = x2 if x2.dtype == BF16 else ops.cast_f32_bf16(x2, torch.empty(M, D, dtype=BF16, device=x.device), st)
        qkv = torch.empty(M, 3 * D, dtype=BF16, device=x.device)
        ops.gemm(a, self.w16(self.qkv.weight), qkv, M, 3 * D, D, bias=self.qkv.bias.data, rope=(rope, hd, D), st=st)
        att = torch.empty(M, D, dtype=BF16, device=x.device)
        lse = torch.empty(B * self.num_heads * N, dtype=F32, device=x.device)
        ops.attn_fwd(qkv, att, lse, B, N, self.num_heads, hd, st)
        y = torch.empty(M, D, dtype=BF16, device=x.device)
        ops.gemm(att, self.w16(self.proj.weight), y, M, D, D, bias=self.proj.bias.data, st=st)
        return y.view(B, N, D)


class _BlockFn(torch.autograd.Function):
    """One autograd node for a directly called Block: engine.block_forward / block_backward on the module's own
    operand copies; parameter gradients come back as ordinary `.grad`s."""

    @staticmethod
    def forward(ctx, blk, x2, B, N, rope, *params):
        from . import engine, ops
        from .workspace import TorchAlloc
        h = engine.BlockH(blk, blk)
        need = any(ctx.needs_input_grad)        # (grad mode is off inside Function.forward)
        ws = TorchAlloc(x2.device)
        out = torch.empty_like(x2)
        saved = engine.block_forward(h, x2, out, B, N, blk.attn.num_heads, blk.attn.head_dim, rope, need, ops.stream(), ws)
        ctx.blk, ctx.h, ctx.saved, ctx.rope, ctx.BN = blk, h, saved, rope, (B, N)
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import engine, ops
        from .workspace import TorchAlloc
        blk, (B, N) = ctx.blk, ctx.BN
        params = list(blk.parameters())
        gbuf = {id(p): torch.zeros(p.shape, dtype=F32, device=dout.device) for p in params}
        g = engine.BlockG(blk, blk, gbuf)
        dx = torch.empty_like(dout)
        engine.block_backward(ctx.h, g, ctx.saved, dout.contiguous(), dx, B, N, blk.attn.num_heads, blk.attn.head_dim,
                              ctx.rope, ops.stream(), TorchAlloc(dout.device))
        ctx.saved = None
        return (None, dx, None, None, None) + tuple(gbuf[id(p)] if p.requires_grad else None for p in params)


class Block(_KernelModule):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, wide_silu=True, norm_layer=nn.LayerNorm, use_sdpa=True,
                 is_causal=False, grid_size=16, use_rope=False, **kwargs):
        super().__init__()
        if not use_rope:
            raise NotImplementedError("vjepa2_b200: only use_rope=True blocks are implemented (all train configs)")
        if drop_path != 0.0:
            raise NotImplementedError("vjepa2_b200: drop_path > 0 is not implemented (rate 0 in all train configs)")
        self.norm1 = norm_layer(dim)
        self.attn = RoPEAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                  use_sdpa=use_sdpa, is_causal=is_causal, grid_size=grid_size, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, mask=None, attn_mask=None, T=None, H_patches=None, W_patches=None):
        """modules.py:556-563 under bf16 autocast: x + attn(norm1(x)), then + mlp(norm2(.)).  x: [B, N, C] residual
        stream, bf16 (encoder) or fp32 (predictor); the output keeps x's dtype.  mask: [B, N] token ids or None."""
        from . import ops
        if attn_mask is not None:
            raise NotImplementedError("vjepa2_b200: attn_mask is not supported (None in every pre-training call)")
        if x.dtype not in (BF16, F32):
            raise TypeError("vjepa2_b200: Block expects a bf16 or fp32 residual stream")
        for n in (self.norm1, self.norm2):
            if not isinstance(n, nn.LayerNorm) or abs(n.eps - 1e-6) > 1e-12:
                raise NotImplementedError("vjepa2_b200: Block norms must be nn.LayerNorm(eps=1e-6)")
        x2, B, N = _rows(x)
        rope = _rope_for(self.attn, mask, B, N, T, H_patches, W_patches, x.device, ops.stream())
        return _BlockFn.apply(self, x2, B, N, rope, *self.parameters()).view(B, N, -1)


class PatchEmbed3D(_KernelModule):
    def __init__(self, patch_size=16, tubelet_size=2, in_chans=3, embed_dim=768):
        super().__init__()
        self.patch_size = patch_size
        self.tubelet_size = tubelet_size
        self.proj = nn.Conv3d(in_channels=in_chans, out_channels=embed_dim,
                              kernel_size=(tubelet_size, patch_size, patch_size),
                              stride=(tubelet_size, patch_size, patch_size))

    def forward(self, x, **kwargs):
        """patch_embed.py:49-52 under bf16 autocast: Conv3d k = s = (tubelet, p, p) as im2col + GEMM;
        fp32 clip [B, C, T, H, W] -> bf16 tokens [B, T/tubelet * H/p * W/p, D], t-major then h then w."""
        from . import ops
        self._no_autograd(x)
        if x.dim() != 5:
            raise ValueError("vjepa2_b200: expected a video tensor [B, C, T, H, W]")
        st = ops.stream()
        cols = ops.im2col_tubelets(x.contiguous().float(), None, self.tubelet_size, self.patch_size, st)
        w = self.w16(self.proj.weight)
        D, K = w.shape[0], cols.shape[1]
        y = torch.empty(cols.shape[0], D, dtype=BF16, device=x.device)
        ops.gemm(cols, w.view(D, K), y, cols.shape[0], D, K, bias=self.proj.bias.data, st=st)
        return y.view(x.shape[0], -1, D)


def init_weights_(module: nn.Module, init_std: float):
    """_init_weights of both reference models (vision_transformer.py:130-146, predictor.py:149-155)."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Conv3d)):
            trunc_normal_(m.weight, std=init_std)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


def rescale_blocks_(blocks):
    """_rescale_blocks (vision_transformer.py:148-154)."""
    for layer_id, layer in enumerate(blocks):
        layer.attn.proj.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
        layer.mlp.fc2.weight.data.div_(math.sqrt(2.0 * (layer_id + 1)))
