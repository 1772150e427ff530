"""Harness around the UNMODIFIED reference (weipeilun/vjepa2) for the like-for-like bar and the tolerance oracle.

Not product code: nothing under vjepa2_b200/ imports this file.  Users: tests/ (full-width parity against the
reference's own modules, on the B200 in fp32 and under bf16 autocast), bench.py (`torch_cuda_baseline`: the
reference's PyTorch-eager CUDA path timed on the same box; `--impl reference`: its CPU path) and
oracle/make_golden_fullwidth.py (compact full-width golden vectors).

The reference is pure Python.  Its sources are NOT part of this repository: `stage()` copies
/root/reference/{src,app,configs/train} into the git-ignored baseline/_ref/ (build container only) so the
tree travels to the GPU box with the gpurun snapshot; `find_ref_root()` looks at $VJEPA_REF, baseline/_ref,
/root/reference in that order.  The one missing import on the path, `timm.models.layers.drop_path`
(src/models/utils/modules.py:9), is dead code at drop-path rate 0 (modules.py:546 -> nn.Identity) and is
stubbed in memory.

The reference's step is a closure inside app/vjepa/train.py:main() (train.py:409-471), which cannot run
without decord and a video dataset; `RefStep.step` restates that closure around the reference's own
encoder / predictor / wrappers / apply_masks / init_opt objects.
"""
from __future__ import annotations

import copy
import os
import shutil
import sys
import types
from functools import partial

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"
PARTS = ("src", "app", os.path.join("configs", "train"))

# name -> (embed_dim, depth, heads, mlp_ratio): vision_transformer.py:275-316
WIDTHS = {
    "vit_large": (1024, 24, 16, 4.0),
    "vit_huge": (1280, 32, 16, 4.0),
    "vit_giant_xformers": (1408, 40, 22, 48 / 11),
}


def stage(force=False):
    """Copy the reference's python packages into baseline/_ref (git-ignored).  No-op without /root/reference."""
    if not os.path.isdir(SOURCE):
        return os.path.isdir(os.path.join(STAGED, "src"))
    for part in PARTS:
        dst = os.path.join(STAGED, part)
        if os.path.isdir(dst):
            if not force:
                continue
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SOURCE, part), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


def find_ref_root():
    for cand in (os.environ.get("VJEPA_REF"), STAGED, SOURCE):
        if cand and os.path.isdir(os.path.join(cand, "src", "models")):
            return cand
    return None


_NS = None


def import_reference(root=None):
    """Import the reference's hot-path modules; returns a namespace (cached).  Raises if no tree is found."""
    global _NS
    if _NS is not None:
        return _NS
    root = root or find_ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (VJEPA_REF, baseline/_ref, /root/reference)")
    if "timm" not in sys.modules:
        t, tm, tl = (types.ModuleType(n) for n in ("timm", "timm.models", "timm.models.layers"))
        tl.drop_path = lambda x, drop_prob=0.0, training=False: x
        t.models, tm.layers = tm, tl
        sys.modules.update({"timm": t, "timm.models": tm, "timm.models.layers": tl})
    import logging
    level = logging.getLogger().level
    sys.path.insert(0, root)
    try:
        from app.vjepa.utils import init_opt, init_video_model
        from src.masks.multiseq_multiblock3d import MaskCollator
        from src.masks.utils import apply_masks
        from src.models.predictor import VisionTransformerPredictor
        from src.models.vision_transformer import VisionTransformer
        from src.utils.wrappers import MultiSeqWrapper, PredictorMultiSeqWrapper
        import src.models.vision_transformer as video_vit
    finally:
        sys.path.remove(root)
        logging.getLogger().setLevel(max(level, logging.WARNING))      # app/vjepa/utils.py sets INFO on the root logger
    _NS = types.SimpleNamespace(
        root=root, init_opt=init_opt, init_video_model=init_video_model, MaskCollator=MaskCollator,
        apply_masks=apply_masks, VisionTransformer=VisionTransformer, video_vit=video_vit,
        VisionTransformerPredictor=VisionTransformerPredictor, MultiSeqWrapper=MultiSeqWrapper,
        PredictorMultiSeqWrapper=PredictorMultiSeqWrapper)
    return _NS


def build_models(R, model_name, crop=256, frames=16, depth=None, pred_depth=12, pred_heads=12, pred_dim=384,
                 num_mask_tokens=6, activation_checkpointing=False, device="cpu"):
    """Reference encoder + predictor in the shipped pre-training configuration
    (configs/train/vitg16/pretrain-256px-16f.yaml model: block), optionally with fewer blocks."""
    import torch.nn as nn
    D, full_depth, heads, ratio = WIDTHS[model_name]
    ln = partial(nn.LayerNorm, eps=1e-6)
    enc = R.VisionTransformer(img_size=crop, patch_size=16, num_frames=frames, tubelet_size=2, embed_dim=D,
                              depth=depth or full_depth, num_heads=heads, mlp_ratio=ratio, qkv_bias=True, norm_layer=ln,
                              uniform_power=True, use_sdpa=True, use_silu=False, wide_silu=False, use_rope=True,
                              use_activation_checkpointing=activation_checkpointing)
    pred = R.VisionTransformerPredictor(img_size=crop, patch_size=16, num_frames=frames, tubelet_size=2, embed_dim=D,
                                        predictor_embed_dim=pred_dim, depth=pred_depth, num_heads=pred_heads,
                                        mlp_ratio=4, qkv_bias=True, norm_layer=ln, uniform_power=True,
                                        use_mask_tokens=True, num_mask_tokens=num_mask_tokens,
                                        zero_init_mask_tokens=True, use_rope=True, use_sdpa=True, use_silu=False,
                                        wide_silu=False, use_activation_checkpointing=activation_checkpointing)
    return R.MultiSeqWrapper(enc).to(device), R.PredictorMultiSeqWrapper(pred).to(device)


class RefStep:
    """The reference's train_step (app/vjepa/train.py:409-471) around the reference's own objects.

    mixed_precision=True: bf16 autocast + GradScaler, as every shipped config (train.py:96-103, 438-451);
    False: plain fp32 (the tight oracle).  `step(..., grads=True)` also returns the unscaled gradients
    (read between scaler.unscale_ and scaler.step)."""

    def __init__(self, R, encoder, predictor, opt, mixed_precision, target_encoder=None):
        import torch
        self.R, self.encoder, self.predictor = R, encoder, predictor
        self.target_encoder = target_encoder if target_encoder is not None else copy.deepcopy(encoder)   # train.py:210
        for p in self.target_encoder.parameters():
            p.requires_grad = False
        self.mixed = mixed_precision
        self.optimizer, self.scaler, self.scheduler, self.wd_scheduler = R.init_opt(
            encoder=encoder, predictor=predictor, iterations_per_epoch=opt["ipe"], start_lr=opt["start_lr"],
            ref_lr=opt["lr"], warmup=opt["warmup"], num_epochs=opt["epochs"], wd=opt["weight_decay"],
            final_wd=opt["final_weight_decay"], final_lr=opt["final_lr"], mixed_precision=mixed_precision,
            ipe_scale=opt["ipe_scale"], betas=opt.get("betas", (0.9, 0.999)), eps=opt.get("eps", 1e-8))
        e0, e1 = opt["ema"]
        total = int(opt["ipe"] * opt["epochs"] * opt["ipe_scale"]) + 1
        self.momentum = (e0 + i * (e1 - e0) / (opt["ipe"] * opt["epochs"] * opt["ipe_scale"]) for i in range(total))
        self.device_type = next(encoder.parameters()).device.type
        self.torch = torch

    def forward_only(self, clips, masks_enc, masks_pred):
        """h (normalised target features), z (predictions), loss -- train.py:414-435."""
        torch, R = self.torch, self.R
        import torch.nn.functional as F
        with torch.autocast(self.device_type, dtype=torch.bfloat16, enabled=self.mixed):
            with torch.no_grad():
                h = [F.layer_norm(hi, (hi.size(-1),)) for hi in self.target_encoder(clips)]
            z_enc = self.encoder(clips, masks_enc)
            z = self.predictor(z_enc, masks_enc, masks_pred)
            hm = [R.apply_masks(hi, mi, concat=False) for hi, mi in zip(h, masks_pred)]
            terms = [torch.mean(torch.abs(zij - hij)) for zi, hi in zip(z, hm) for zij, hij in zip(zi, hi)]
            loss = sum(terms) / len(terms)
        return h, z_enc, z, loss

    def step(self, clips, masks_enc, masks_pred, grads=False, keep=False):
        torch = self.torch
        lr, wd = self.scheduler.step(), self.wd_scheduler.step()
        h, z_enc, z, loss = self.forward_only(clips, masks_enc, masks_pred)
        if self.mixed:
            self.scaler.scale(loss).backward()
            self.scaler.unscale_(self.optimizer)
        else:
            loss.backward()
        captured = None
        if grads:
            captured = ({n: p.grad.detach().clone() for n, p in self.encoder.backbone.named_parameters() if p.grad is not None},
                        {n: p.grad.detach().clone() for n, p in self.predictor.backbone.named_parameters() if p.grad is not None})
        if self.mixed:
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            self.optimizer.step()
        self.optimizer.zero_grad()
        m = next(self.momentum)
        with torch.no_grad():
            tk = list(self.target_encoder.parameters())
            torch._foreach_mul_(tk, m)
            torch._foreach_add_(tk, list(self.encoder.parameters()), alpha=1 - m)
        out = dict(loss=float(loss.detach()), lr=lr, wd=wd)
        if grads:
            out["grads"] = captured
        if keep:
            out.update(h=h, z_enc=z_enc, z=z)
        return out


if __name__ == "__main__":
    print("staged:", stage(force="--force" in sys.argv), "->", find_ref_root())
