"""Drop-in for src/models/predictor.py of the reference (VisionTransformerPredictor :18-246,
vit_predictor :249-253): same constructor arguments, parameter names and
forward(x, masks_x, masks_y, mask_index=1, has_cls=False) signature, executed on sm_100a kernels.

Scope: use_rope=True, use_mask_tokens=True, exact-GELU MLP, has_cls=False, return_all_tokens=False,
chop_last_n_tokens=0 -- what the pre-training configs use; anything else raises.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from . import engine
from .modules import Block, init_weights_, rescale_blocks_, trunc_normal_
from .vision_transformer import _FlatModule


class _PredictorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, z, masks_x, masks_y, mask_index, *params):
        rt = model.runtime()
        out, saved = engine.predictor_forward(rt, z, masks_x, masks_y, mask_index, save=True)
        ctx.rt, ctx.saved, ctx.z_dtype = rt, saved, z.dtype
        ctx.mi = mask_index % len(rt.mask_tokens)
        return out

    @staticmethod
    def backward(ctx, dout):
        rt = ctx.rt
        fs = rt.fs
        gbuf = torch.zeros(fs.total, dtype=torch.float32, device=dout.device)
        dz = engine.predictor_backward(rt, ctx.saved, dout.contiguous(), gbuf)
        mi = ctx.mi
        ctx.saved = None
        used = {id(p) for p in fs.params} - {id(p) for k, p in enumerate(rt.model.mask_tokens) if k != mi}
        grads = tuple(fs.grad_view(gbuf, p) if (p.requires_grad and id(p) in used) else None for p in fs.params)
        return (None, dz.to(ctx.z_dtype), None, None, None) + grads


class VisionTransformerPredictor(_FlatModule):
    _RT = engine.PredictorRT

    def __init__(self, img_size=(224, 224), patch_size=16, num_frames=1, tubelet_size=2, embed_dim=768,
                 predictor_embed_dim=384, depth=6, num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm, init_std=0.02,
                 uniform_power=False, use_mask_tokens=False, num_mask_tokens=2, zero_init_mask_tokens=True,
                 use_silu=False, wide_silu=True, use_activation_checkpointing=False, return_all_tokens=False,
                 chop_last_n_tokens=0, use_rope=False, **kwargs):
        super().__init__()
        if not use_rope or not use_mask_tokens or use_silu:
            raise NotImplementedError("vjepa2_b200: predictor supports use_rope=True, use_mask_tokens=True, "
                                      "use_silu=False (the pre-training configs)")
        if return_all_tokens or chop_last_n_tokens or drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("vjepa2_b200: return_all_tokens / chop_last_n_tokens / dropout are out of scope")
        self.return_all_tokens = return_all_tokens
        self.chop_last_n_tokens = chop_last_n_tokens
        self.predictor_embed = nn.Linear(embed_dim, predictor_embed_dim, bias=True)
        self.num_mask_tokens = num_mask_tokens
        self.mask_tokens = nn.ParameterList(
            [nn.Parameter(torch.zeros(1, 1, predictor_embed_dim)) for _ in range(num_mask_tokens)])
        if type(img_size) is int:
            img_size = (img_size, img_size)
        self.img_height, self.img_width = img_size
        self.patch_size = patch_size
        self.num_frames = num_frames
        self.tubelet_size = tubelet_size
        self.is_video = num_frames > 1
        self.grid_height = img_size[0] // patch_size
        self.grid_width = img_size[1] // patch_size
        self.grid_depth = num_frames // tubelet_size
        self.use_activation_checkpointing = use_activation_checkpointing
        self.num_patches = self.grid_depth * self.grid_height * self.grid_width if self.is_video \
            else self.grid_height * self.grid_width
        self.uniform_power = uniform_power
        self.predictor_pos_embed = None
        self.use_rope = use_rope
        self.predictor_blocks = nn.ModuleList([
            Block(use_rope=use_rope, grid_size=self.grid_height, grid_depth=self.grid_depth, dim=predictor_embed_dim,
                  num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                  act_layer=nn.GELU, wide_silu=wide_silu, attn_drop=attn_drop_rate, drop_path=0.0,
                  norm_layer=norm_layer)
            for _ in range(depth)])
        self.predictor_norm = norm_layer(predictor_embed_dim)
        self.predictor_proj = nn.Linear(predictor_embed_dim, embed_dim, bias=True)
        self.init_std = init_std
        if not zero_init_mask_tokens:
            for mt in self.mask_tokens:
                trunc_normal_(mt, std=init_std)
        init_weights_(self, init_std)
        rescale_blocks_(self.predictor_blocks)
        self._init_flat_state()

    def forward(self, x, masks_x, masks_y, mask_index=1, has_cls=False):
        """x: context tokens [B, Kc, D]; masks_x [B, Kc], masks_y [B, Kp] token ids (disjoint).
        Returns bf16 [B, Kp, D] predictions for the masks_y positions (predictor.py:242-246)."""
        if has_cls:
            raise NotImplementedError("vjepa2_b200: has_cls=True is out of scope")
        if isinstance(masks_x, list) or isinstance(masks_y, list):
            if len(masks_x) != 1 or len(masks_y) != 1:
                raise NotImplementedError("vjepa2_b200: one (masks_x, masks_y) pair per call "
                                          "(PredictorMultiSeqWrapper calls the backbone once per mask)")
            masks_x, masks_y = masks_x[0], masks_y[0]
        dev = x.device
        masks_x = masks_x.to(device=dev, dtype=torch.int64).contiguous()
        masks_y = masks_y.to(device=dev, dtype=torch.int64).contiguous()
        x = x.contiguous()
        rt = self.runtime()
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in rt.fs.params))
        if not needs_grad:
            out, _ = engine.predictor_forward(rt, x, masks_x, masks_y, mask_index, save=False)
            return out
        return _PredictorFn.apply(self, x, masks_x, masks_y, mask_index, *rt.fs.params)


def vit_predictor(**kwargs):
    return VisionTransformerPredictor(mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
