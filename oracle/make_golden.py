"""Generate tests/golden/*.pt by running the REAL reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py            # writes tests/golden/ref_golden.pt

Test infrastructure.  The reference needs `timm.models.layers.drop_path`, which is not
installed; it is dead code at drop-path rate 0 (modules.py:546 -> nn.Identity), so a stub
module is injected.  Inputs and weights are produced by seeded generators that live in
oracle/vjepa_oracle.py, so the tests can rebuild them bit-identically and only the
reference OUTPUTS need to be committed.
"""
import copy
import os
import sys
import types
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VJEPA_REF", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

_timm = types.ModuleType("timm")
_timm_models = types.ModuleType("timm.models")
_timm_layers = types.ModuleType("timm.models.layers")
_timm_layers.drop_path = lambda x, p=0.0, training=False: x
sys.modules.update({"timm": _timm, "timm.models": _timm_models, "timm.models.layers": _timm_layers})

import vjepa_oracle as O  # noqa: E402
from src.masks.multiseq_multiblock3d import MaskCollator  # noqa: E402
from src.masks.utils import apply_masks  # noqa: E402
from src.models.predictor import VisionTransformerPredictor  # noqa: E402
from src.models.utils.modules import rotate_queries_or_keys  # noqa: E402
from src.models.vision_transformer import VisionTransformer  # noqa: E402
from src.utils.schedulers import CosineWDSchedule, WarmupCosineSchedule  # noqa: E402
from app.vjepa.utils import init_opt  # noqa: E402

# ---- the tiny geometry every golden test uses ------------------------------------------
TINY = dict(img=96, frames=8, patch=16, tubelet=2, dim=128, depth=2, heads=2, mlp_ratio=4.0,
            pred_dim=64, pred_depth=2, pred_heads=2, num_mask_tokens=2)


def tiny_clips(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, TINY["frames"], TINY["img"], TINY["img"], generator=g)


def tiny_masks(B, seed=5):
    """Two disjoint ascending index sets per sample, same length across the batch."""
    g = torch.Generator().manual_seed(seed)
    N = (TINY["frames"] // 2) * (TINY["img"] // 16) ** 2
    me, mp = [], []
    for _ in range(B):
        perm = torch.randperm(N, generator=g)
        me.append(perm[:40].sort().values)
        mp.append(perm[40:40 + 72].sort().values)
    return torch.stack(me), torch.stack(mp)


def build_ref_models():
    t = TINY
    enc = VisionTransformer(img_size=t["img"], patch_size=t["patch"], num_frames=t["frames"],
                            tubelet_size=t["tubelet"], embed_dim=t["dim"], depth=t["depth"],
                            num_heads=t["heads"], mlp_ratio=t["mlp_ratio"], qkv_bias=True,
                            norm_layer=partial(nn.LayerNorm, eps=1e-6), use_rope=True)
    pred = VisionTransformerPredictor(img_size=t["img"], patch_size=t["patch"], num_frames=t["frames"],
                                      tubelet_size=t["tubelet"], embed_dim=t["dim"],
                                      predictor_embed_dim=t["pred_dim"], depth=t["pred_depth"],
                                      num_heads=t["pred_heads"], use_mask_tokens=True,
                                      num_mask_tokens=t["num_mask_tokens"], zero_init_mask_tokens=True,
                                      use_rope=True, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6))
    w_enc = O.init_encoder_weights(t["dim"], t["depth"], t["mlp_ratio"], seed=0, rand_bias=True)
    w_pred = O.init_predictor_weights(t["dim"], t["pred_dim"], t["pred_depth"], t["num_mask_tokens"],
                                      seed=1, rand_bias=True)
    # strict load: proves the oracle's parameter names / shapes equal the reference's
    enc.load_state_dict(w_enc, strict=True)
    pred.load_state_dict(w_pred, strict=True)
    return enc, pred, w_enc, w_pred


def main():
    torch.set_num_threads(8)
    G = {}
    G2 = {}     # second file (tests/golden/ref_golden_infer.pt): optimizer moments + encoder-only inference paths

    # 1. rotate_queries_or_keys (modules.py:26-50) fwd + autograd grad, fp64 and fp32
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 17, 20, generator=g, dtype=torch.float64, requires_grad=True)
    pos = torch.randint(0, 16, (2, 3, 17), generator=g).to(torch.float64)
    y = rotate_queries_or_keys(x, pos)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (gx,) = torch.autograd.grad(y, x, gy)
    G["rope.x"], G["rope.pos"], G["rope.y"], G["rope.gy"], G["rope.gx"] = x.detach(), pos, y.detach(), gy, gx

    # 2. apply_masks (masks/utils.py:9-21), incl. duplicate / unsorted indices
    xm = torch.randn(3, 12, 8, generator=g)
    m1 = torch.randint(0, 12, (3, 5), generator=g)
    m2 = torch.randint(0, 12, (3, 7), generator=g)
    G["am.x"], G["am.m1"], G["am.m2"] = xm, m1, m2
    G["am.cat"] = apply_masks(xm, [m1, m1.flip(1)])
    G["am.list1"] = apply_masks(xm, [m2], concat=False)[0]

    # 3. tiny encoder / predictor forward, fp32, reference modules
    enc, pred, w_enc, w_pred = build_ref_models()
    clips = tiny_clips(2)
    me, mp = tiny_masks(2)
    with torch.no_grad():
        G["enc.full"] = enc(clips)
        G["enc.masked"] = enc(clips, me)
        G["enc.patch_embed"] = enc.patch_embed(clips)
        blk_in = G["enc.patch_embed"]
        G["enc.block0"] = enc.blocks[0](blk_in, mask=None, attn_mask=None, T=4, H_patches=6, W_patches=6)
        z = G["enc.masked"]
        G["pred.out"] = pred(z, me, mp, mask_index=0)
        G["pred.out_idx1"] = pred(z, me, mp, mask_index=1)

    # 4. MaskCollator (multiseq_multiblock3d.py) with the shipped mask config, 16f x 256px
    cfgs = [dict(m) for m in O.DEFAULT_MASK_CFG]
    coll = MaskCollator(cfgs_mask=cfgs, dataset_fpcs=[16], crop_size=(256, 256), patch_size=(16, 16), tubelet_size=2)
    torch.manual_seed(239)
    for it in range(3):
        batch = [(torch.zeros(1), 0, [list(range(16))]) for _ in range(6)]
        out = coll(batch)
        _, menc, mpred = out[0]
        for j in range(2):
            G[f"mask.it{it}.enc{j}"] = menc[j]
            G[f"mask.it{it}.pred{j}"] = mpred[j]
    # 64f x 384px cooldown geometry, one draw
    coll2 = MaskCollator(cfgs_mask=cfgs, dataset_fpcs=[64], crop_size=(384, 384), patch_size=(16, 16), tubelet_size=2)
    torch.manual_seed(7)
    out = coll2([(torch.zeros(1), 0, [list(range(64))]) for _ in range(2)])
    for j in range(2):
        G[f"mask384.enc{j}"] = out[0][1][j]
        G[f"mask384.pred{j}"] = out[0][2][j]

    # 5. schedulers (schedulers.py:41-93) on a dummy optimizer
    dummy = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=1.0)
    sch = WarmupCosineSchedule(dummy, warmup_steps=4, start_lr=1e-4, ref_lr=5.25e-4, final_lr=1e-5, T_max=20)
    wds = CosineWDSchedule(dummy, ref_wd=0.04, final_wd=0.4, T_max=20)
    G["sched.lr"] = torch.tensor([sch.step() for _ in range(24)], dtype=torch.float64)
    G["sched.wd"] = torch.tensor([wds.step() for _ in range(24)], dtype=torch.float64)

    # 6. two full training steps with the reference modules (train.py:409-471 restated around
    #    imported modules; fp32, mixed_precision=False), B=2, two masks
    target = copy.deepcopy(enc)
    for p in target.parameters():
        p.requires_grad = False
    opt_cfg = dict(ipe=10, epochs=2, ipe_scale=1.25, warmup=0.2, start_lr=1e-4, lr=5.25e-4, final_lr=1e-5,
                   weight_decay=0.04, final_weight_decay=0.4, ema=(0.99, 1.0), loss_exp=1.0)
    optimizer, _, scheduler, wd_scheduler = init_opt(
        encoder=enc, predictor=pred, iterations_per_epoch=opt_cfg["ipe"], start_lr=opt_cfg["start_lr"],
        ref_lr=opt_cfg["lr"], warmup=opt_cfg["warmup"], num_epochs=opt_cfg["epochs"],
        wd=opt_cfg["weight_decay"], final_wd=opt_cfg["final_weight_decay"], final_lr=opt_cfg["final_lr"],
        mixed_precision=False, ipe_scale=opt_cfg["ipe_scale"])
    ipe, ne, ipes = opt_cfg["ipe"], opt_cfg["epochs"], opt_cfg["ipe_scale"]
    momentum = (opt_cfg["ema"][0] + i * (opt_cfg["ema"][1] - opt_cfg["ema"][0]) / (ipe * ne * ipes)
                for i in range(int(ipe * ne * ipes) + 1))
    me2, mp2 = tiny_masks(2, seed=6)
    masks_enc, masks_pred = [me, me2[:, :24]], [mp, mp2[:, :88]]
    for it in range(2):
        scheduler.step()
        wd_scheduler.step()
        with torch.no_grad():
            h = F.layer_norm(target(clips), (TINY["dim"],))
        zs = [pred(enc(clips, m_e), m_e, m_p, mask_index=0) for m_e, m_p in zip(masks_enc, masks_pred)]
        hs = apply_masks(h, masks_pred, concat=False)
        loss = sum(torch.mean(torch.abs(zj - hj)) for zj, hj in zip(zs, hs)) / len(zs)
        loss.backward()
        if it == 0:
            G["step.h"] = h
            G["step.z0"] = zs[0].detach()
            for k in ["blocks.0.attn.qkv.weight", "blocks.1.mlp.fc2.bias", "patch_embed.proj.weight",
                      "norm.weight", "blocks.0.norm1.bias"]:
                G["step.genc." + k] = dict(enc.named_parameters())[k].grad.clone()
            for k in ["predictor_embed.weight", "mask_tokens.0", "predictor_blocks.1.attn.proj.weight",
                      "predictor_proj.bias", "predictor_norm.weight"]:
                G["step.gpred." + k] = dict(pred.named_parameters())[k].grad.clone()
            assert dict(pred.named_parameters())["mask_tokens.1"].grad is None
        optimizer.step()
        optimizer.zero_grad()
        m = next(momentum)
        with torch.no_grad():
            pk, pq = list(target.parameters()), list(enc.parameters())
            torch._foreach_mul_(pk, m)
            torch._foreach_add_(pk, pq, alpha=1 - m)
        G[f"step.loss{it}"] = loss.detach()
    for k in ["blocks.0.attn.qkv.weight", "blocks.1.mlp.fc2.bias", "norm.weight"]:
        G["step.after.enc." + k] = enc.state_dict()[k].clone()
        G["step.after.tgt." + k] = target.state_dict()[k].clone()
    for k in ["predictor_embed.weight", "mask_tokens.0", "mask_tokens.1", "predictor_proj.bias"]:
        G["step.after.pred." + k] = pred.state_dict()[k].clone()
    # AdamW moments after the two steps (what a checkpoint's "opt" entry carries, train.py:320)
    for k in ["blocks.0.attn.qkv.weight", "blocks.1.mlp.fc2.bias", "norm.weight"]:
        st_ = optimizer.state[dict(enc.named_parameters())[k]]
        G2["opt.enc.exp_avg." + k], G2["opt.enc.exp_avg_sq." + k] = st_["exp_avg"].clone(), st_["exp_avg_sq"].clone()
        G2["opt.step"] = st_["step"].clone()
    assert dict(pred.named_parameters())["mask_tokens.1"] not in optimizer.state

    # 7. ViT-H geometry in miniature: head_dim 80 (RoPE segments 26/26/26 + 2 pass-through dims)
    encH = VisionTransformer(img_size=TINY["img"], patch_size=16, num_frames=TINY["frames"], tubelet_size=2,
                             embed_dim=160, depth=2, num_heads=2, mlp_ratio=4.0, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), use_rope=True)
    encH.load_state_dict(O.init_encoder_weights(160, 2, 4.0, seed=3, rand_bias=True), strict=True)
    with torch.no_grad():
        G["encH.full"] = encH(clips)
        G["encH.masked"] = encH(clips, me)

    # 8. encoder-only inference: out_layers (vision_transformer.py:204-208) and the evals' ClipAggregation wrappers
    #    (evals/video_classification_frozen/modelcustom/vit_encoder_multiclip{,_multilevel}.py), 2 clips x 2 views
    from evals.video_classification_frozen.modelcustom.vit_encoder_multiclip import ClipAggregation as CA
    from evals.video_classification_frozen.modelcustom.vit_encoder_multiclip_multilevel import ClipAggregation as CAML
    t = TINY
    enc8 = VisionTransformer(img_size=t["img"], patch_size=16, num_frames=t["frames"], tubelet_size=2,
                             embed_dim=t["dim"], depth=t["depth"], num_heads=t["heads"], mlp_ratio=t["mlp_ratio"],
                             qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), use_rope=True)
    enc8.load_state_dict(w_enc, strict=True)
    enc8ml = VisionTransformer(img_size=t["img"], patch_size=16, num_frames=t["frames"], tubelet_size=2,
                               embed_dim=t["dim"], depth=t["depth"], num_heads=t["heads"], mlp_ratio=t["mlp_ratio"],
                               qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), use_rope=True,
                               out_layers=[0, 1])
    enc8ml.load_state_dict(w_enc, strict=True)
    views = [[tiny_clips(2, seed=20 + 2 * i + j) for j in range(2)] for i in range(2)]
    with torch.no_grad():
        outs = enc8ml(clips)
        G2["infer.out_layers.0"], G2["infer.out_layers.1"] = outs[0], outs[1]
        for j, o in enumerate(CA(enc8, tubelet_size=2)(views)):
            G2[f"infer.agg.view{j}"] = o
        for j, o in enumerate(CAML(enc8ml, tubelet_size=2)(views)):
            G2[f"infer.aggml.view{j}"] = o

    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "ref_golden.pt")
    torch.save({k: v.contiguous() for k, v in G.items()}, path)
    print("wrote", path, os.path.getsize(path), "bytes;", len(G), "tensors; torch", torch.__version__)
    path2 = os.path.join(out_dir, "ref_golden_infer.pt")
    torch.save({k: v.contiguous() for k, v in G2.items()}, path2)
    print("wrote", path2, os.path.getsize(path2), "bytes;", len(G2), "tensors")


if __name__ == "__main__":
    main()
