"""Encoder-only inference entry points of the reference, on the same sm_100a kernels as the training step.

  * hub constructors  vjepa2_vit_large / vjepa2_vit_huge / vjepa2_vit_giant / vjepa2_vit_giant_384
    (src/hub/backbones.py:88-168): same encoder / predictor keyword sets; weights come from a local checkpoint
    file (there is no download path in this package).
  * init_module + ClipAggregation  (evals/video_classification_frozen/modelcustom/vit_encoder_multiclip.py:40-149 and its
    `_multilevel` variant):
    the frozen-encoder feature extractor the evals call -- all clips and views go through ONE encoder call
    (batch-concatenated), then tokens are regrouped per spatial view and concatenated along time.

The encoder forward itself is `VisionTransformer.forward` (no autograd graph is recorded when gradients are
disabled or the parameters are frozen; `out_layers` returns the per-layer normalised outputs,
vision_transformer.py:204-208).  The regrouping below is slicing / concatenation of result tensors only.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import vision_transformer as vit
from .checkpoint import clean_backbone_key, load_pretrained
from .predictor import vit_predictor

ARCH_NAME_MAP = {                       # src/hub/backbones.py:14-20 (the AC model is out of scope)
    "vit_large": ("vit_large", "vitl"),
    "vit_huge": ("vit_huge", "vith"),
    "vit_giant": ("vit_giant_xformers", "vitg"),
    "vit_giant_384": ("vit_giant_xformers", "vitg-384"),
}


def _make_vjepa2_model(*, model_name="vit_large", img_size=256, patch_size=16, tubelet_size=2, num_frames=64,
                       pretrained=False, checkpoint=None, **kwargs):
    """src/hub/backbones.py:88-136.  `pretrained=True` needs `checkpoint=<path to vitl.pt / vith.pt / ...>`."""
    enc_kwargs = dict(patch_size=patch_size, img_size=(img_size, img_size), num_frames=num_frames,
                      tubelet_size=tubelet_size, use_sdpa=True, use_SiLU=False, wide_SiLU=True, uniform_power=False,
                      use_rope=True)
    enc_kwargs.update(**kwargs)
    encoder = vit.__dict__[ARCH_NAME_MAP[model_name][0]](**enc_kwargs)
    pred_kwargs = dict(img_size=(img_size, img_size), patch_size=patch_size, use_mask_tokens=True,
                       embed_dim=encoder.embed_dim, predictor_embed_dim=384, num_frames=num_frames,
                       tubelet_size=tubelet_size, depth=12, num_heads=12, num_mask_tokens=10, use_rope=True,
                       uniform_power=False, use_sdpa=True, use_silu=False, wide_silu=True)
    pred_kwargs.update(**kwargs)
    predictor = vit_predictor(**pred_kwargs)
    if pretrained:
        if checkpoint is None:
            raise RuntimeError(f"vjepa2_b200: no download path; pass checkpoint=<local {ARCH_NAME_MAP[model_name][1]}.pt>")
        sd = checkpoint if isinstance(checkpoint, dict) else torch.load(checkpoint, map_location="cpu",
                                                                        weights_only=False)
        encoder.load_state_dict(clean_backbone_key(sd["encoder"]), strict=False)      # pos_embed keys are ignored
        predictor.load_state_dict(clean_backbone_key(sd["predictor"]), strict=False)
    return encoder, predictor


def vjepa2_vit_large(*, pretrained=False, **kwargs):
    return _make_vjepa2_model(model_name="vit_large", img_size=256, pretrained=pretrained, **kwargs)


def vjepa2_vit_huge(*, pretrained=False, **kwargs):
    return _make_vjepa2_model(model_name="vit_huge", img_size=256, pretrained=pretrained, **kwargs)


def vjepa2_vit_giant(*, pretrained=False, **kwargs):
    return _make_vjepa2_model(model_name="vit_giant", img_size=256, pretrained=pretrained, **kwargs)


def vjepa2_vit_giant_384(*, pretrained=False, **kwargs):
    return _make_vjepa2_model(model_name="vit_giant_384", img_size=384, pretrained=pretrained, **kwargs)


def init_module(resolution, frames_per_clip, checkpoint, model_kwargs, wrapper_kwargs, device="cuda"):
    """vit_encoder_multiclip.py:40-79: build the named encoder, load `checkpoint[checkpoint_key]` (prefixes stripped,
    mismatching tensors skipped), wrap it in ClipAggregation.  `checkpoint` may be a path or an already loaded dict."""
    enc_kwargs = dict(model_kwargs["encoder"])
    key = enc_kwargs.pop("checkpoint_key", "target_encoder")
    name = enc_kwargs.pop("model_name")
    if wrapper_kwargs.get("out_layers") is not None:       # vit_encoder_multiclip_multilevel.py:56-60
        enc_kwargs["out_layers"] = wrapper_kwargs["out_layers"]
    model = vit.__dict__[name](img_size=resolution, num_frames=frames_per_clip, **enc_kwargs)
    load_pretrained(model, checkpoint, checkpoint_key=key, strict=False)
    model.to(device)
    for p in model.parameters():
        p.requires_grad = False                   # frozen-encoder evals (eval.py runs it under no_grad)
    return ClipAggregation(model, tubelet_size=model.tubelet_size, **wrapper_kwargs)


class ClipAggregation(nn.Module):
    """Process each clip independently and concatenate all tokens (vit_encoder_multiclip.py:82-149).

    x: list over clips of lists over spatial views of [B, C, F, H, W] tensors.  Returns a list over views of
    [B, num_clips * T * S, D] token tensors (clips concatenated along time)."""

    def __init__(self, model, tubelet_size=2, max_frames=128, use_pos_embed=False, out_layers=None):
        super().__init__()
        if use_pos_embed:
            raise NotImplementedError("vjepa2_b200: the 1-D temporal sincos embedding of ClipAggregation is out of "
                                      "scope (use_pos_embed=False in the RoPE eval configs)")
        self.model = model
        self.tubelet_size = tubelet_size
        self.embed_dim = model.embed_dim
        self.num_heads = model.num_heads
        self.pos_embed = None

    def forward(self, x, clip_indices=None):
        num_clips = len(x)
        num_views = len(x[0])
        B, C, F, H, W = x[0][0].size()
        xs = torch.cat([torch.cat(xi, dim=0) for xi in x], dim=0)     # [clips * views * B, C, F, H, W]
        out = self.model(xs)
        if isinstance(out, list):                  # out_layers (vit_encoder_multiclip_multilevel.py:115-116): the
            out = torch.cat(out, dim=1)            # selected layers' tokens are concatenated along the token axis
        return self._regroup(out, B, F, num_clips, num_views)

    def _regroup(self, outputs, B, F, num_clips, num_views):
        _, N, D = outputs.size()
        T = F // self.tubelet_size
        S = N // T
        eff_B = B * num_views
        per_view = [[] for _ in range(num_views)]
        for i in range(num_clips):
            o = outputs[i * eff_B:(i + 1) * eff_B]
            for j in range(num_views):
                per_view[j].append(o[j * B:(j + 1) * B].reshape(B, T, S, D))
        return [torch.cat(v, dim=1).flatten(1, 2) for v in per_view]
