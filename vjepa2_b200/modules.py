"""Parameter containers mirroring src/models/utils/modules.py of the reference (MLP :67-83,
RoPEAttention :261-382, Block :500-563) and src/models/utils/patch_embed.py (PatchEmbed3D :26-52).

The classes keep the reference's module tree and parameter names (`norm1`, `attn.qkv`, `attn.proj`,
`norm2`, `mlp.fc1`, `mlp.fc2`, `patch_embed.proj`) so checkpoints load unchanged.  They hold weights
only: the compute of a block is executed by vjepa2_b200.engine on the sm_100a kernels, driven by the
owning VisionTransformer / VisionTransformerPredictor.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


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


class MLP(nn.Module):
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


class RoPEAttention(nn.Module):
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


class Block(nn.Module):
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


class PatchEmbed3D(nn.Module):
    def __init__(self, patch_size=16, tubelet_size=2, in_chans=3, embed_dim=768):
        super().__init__()
        self.patch_size = patch_size
        self.tubelet_size = tubelet_size
        self.proj = nn.Conv3d(in_channels=in_chans, out_channels=embed_dim,
                              kernel_size=(tubelet_size, patch_size, patch_size),
                              stride=(tubelet_size, patch_size, patch_size))


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
