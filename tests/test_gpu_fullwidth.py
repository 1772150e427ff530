"""Full-width GPU parity: BASELINE.json's real widths (ViT-L 1024/16 heads, ViT-H 1280/head_dim 80, ViT-g 1408/22
heads/Hm 6144; predictor 384/12 heads; 16 x 256^2 clips -> 2048 tokens; the shipped multiblock mask config) against

  (1) compact golden vectors of the REAL reference run in fp32 on CPU (tests/golden/ref_fullwidth.pt, written by
      oracle/make_golden_fullwidth.py) -- always available, and
  (2) the real reference modules run live on the same B200, in fp32 (tight oracle) AND under bf16 autocast (the
      reference's own precision, SURVEY 8c oracle (ii)); needs the reference tree (baseline/_ref, staged by
      __graft_entry__.build()).  Our error against fp32 is reported next to autocast's own error against fp32.

Tolerances (north_star): activations rel. Frobenius error <= 1e-2, loss within 1e-3; gradients: matrices <= 3e-2,
vectors (biases / LN affine: long cancelling sums) <= 5e-2 -- and never worse than 1.5x what the reference's own
bf16-autocast path shows against its fp32 self (+1e-3 floor)."""
import json
import os
import sys
from functools import partial

import pytest
import torch
import torch.nn as nn

import fullwidth_common as FW

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_harness as H  # noqa: E402

pytestmark = pytest.mark.gpu
HAVE_REF = H.find_ref_root() is not None
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference tree not staged (baseline/_ref)")
REPORT = {}


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "ref_fullwidth.pt"), map_location="cpu")


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    yield
    if REPORT:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "fullwidth_parity.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)


def build_ours(name, dev, depth=None, pred_depth=None, w=None):
    from vjepa2_b200.predictor import vit_predictor
    from vjepa2_b200.vision_transformer import VisionTransformer
    D, heads, ratio, d0 = FW.CASES[name]
    enc = VisionTransformer(img_size=FW.CROP, patch_size=16, num_frames=FW.FRAMES, tubelet_size=2, embed_dim=D,
                            depth=depth or d0, num_heads=heads, mlp_ratio=ratio, qkv_bias=True,
                            norm_layer=partial(nn.LayerNorm, eps=1e-6), uniform_power=True, use_sdpa=True, use_rope=True)
    pred = vit_predictor(img_size=FW.CROP, patch_size=16, num_frames=FW.FRAMES, tubelet_size=2, embed_dim=D,
                         predictor_embed_dim=FW.PRED["dim"], depth=pred_depth or FW.PRED["depth"],
                         num_heads=FW.PRED["heads"], uniform_power=True, use_mask_tokens=True,
                         num_mask_tokens=FW.PRED["num_mask_tokens"], zero_init_mask_tokens=True, use_rope=True,
                         use_sdpa=True)
    if w is not None:
        enc.load_state_dict(w[0], strict=True)
        pred.load_state_dict(w[1], strict=True)
    return enc.to(dev), pred.to(dev)


def our_masks():
    from vjepa2_b200.masks import MaskCollator
    return FW.draw_masks(MaskCollator)


def ours_forward(enc, pred, tgt, clips, me, mp):
    """h, z_enc[j], z[j] through the public module API (no grad)."""
    with torch.no_grad():
        h = torch.nn.functional.layer_norm(tgt(clips), (tgt.embed_dim,))
        z_enc = [enc(clips, m) for m in me]
        z = [pred(ze, m, p, mask_index=0) for ze, m, p in zip(z_enc, me, mp)]
    return h, z_enc, z


def ours_step(name, dev, w, clips, me, mp, depth=None, pred_depth=None):
    """One fused train step; returns loss, unscaled gradient dicts, activations of the same weights."""
    from vjepa2_b200.train import JepaTrainStep
    enc, pred = build_ours(name, dev, depth, pred_depth, w)
    step = JepaTrainStep(enc, pred, **FW.OPT)
    cd = clips.to(dev)
    med, mpd = [m.to(dev) for m in me], [m.to(dev) for m in mp]
    acts = ours_forward(enc, pred, step.target_encoder, cd, med, mpd)
    loss, _, _ = step.step([cd], [med], [mpd])
    loss = float(loss.item())
    efs, pfs = step.enc_rt.fs, step.pred_rt.fs
    g_enc = {n: (efs.grad_view(efs.g32, p) / 65536.0).cpu() for n, p in enc.named_parameters()}
    g_pred = {n: (pfs.grad_view(pfs.g32, p) / 65536.0).cpu() for n, p in pred.named_parameters()}
    return loss, g_enc, g_pred, acts


def grad_tol(g):
    return 5e-2 if g.dim() == 1 or g.numel() == g.shape[-1] else 3e-2


@pytest.mark.parametrize("name", list(FW.CASES))
def test_fullwidth_step_vs_reference_golden(name, dev, gold):
    """(1): loss, sampled activation rows / all row norms, gradient norms and strided gradient samples of the real
    reference's fp32 step at full width."""
    p = name + "."
    me, mp = our_masks()
    for j in range(2):       # the mirror's collator draws the reference's indices bit-exactly on the 8x16x16 grid
        assert torch.equal(me[j], gold[p + f"masks_enc.{j}"].long()) and torch.equal(mp[j], gold[p + f"masks_pred.{j}"].long())
    loss, g_enc, g_pred, (h, z_enc, z) = ours_step(name, dev, FW.weights(name), FW.clips(), me, mp)
    rep = REPORT.setdefault("golden", {}).setdefault(name, {})
    rep["loss"] = (loss, float(gold[p + "loss"]))
    assert abs(loss - float(gold[p + "loss"])) < 1e-3, (loss, float(gold[p + "loss"]))
    acts = {"h": h, **{f"z_enc.{j}": t for j, t in enumerate(z_enc)}, **{f"z.{j}": t for j, t in enumerate(z)}}
    for k, a in acts.items():
        a2 = a.float().reshape(-1, a.shape[-1]).cpu()
        rows = FW.row_sample(a2.shape[0], seed=len(k))
        e_rows = relerr(a2[rows], gold[p + k + ".rows"])
        e_norm = relerr(a2.norm(dim=1), gold[p + k + ".rownorm"])
        rep[k] = (e_rows, e_norm)
        assert e_rows < 1e-2 and e_norm < 1e-2, (k, e_rows, e_norm)
    worst = {}
    for tag, gd in (("genc", g_enc), ("gpred", g_pred)):
        for n, g in gd.items():
            key = p + f"{tag}.{n}.slice"
            if key not in gold:
                assert float(g.abs().max()) == 0.0, n            # unused mask tokens: no gradient
                continue
            e = relerr(FW.grad_slice(g), gold[key])
            en = abs(float(g.norm()) / float(gold[p + f"{tag}.{n}.norm"]) - 1.0)
            cls = tag + (".vec" if grad_tol(g) == 5e-2 else ".mat")
            worst[cls] = max(worst.get(cls, 0.0), e)
            assert e < grad_tol(g) and en < grad_tol(g), (n, e, en)
    rep["grads_worst"] = worst


def _ref_run(R, name, mixed, w, clips, me, mp, dev, depth=None, pred_depth=None):
    D, heads, ratio, d0 = FW.CASES[name]
    enc, pred = H.build_models(R, name, crop=FW.CROP, frames=FW.FRAMES, depth=depth or d0,
                               pred_depth=pred_depth or FW.PRED["depth"], pred_heads=FW.PRED["heads"],
                               pred_dim=FW.PRED["dim"], num_mask_tokens=FW.PRED["num_mask_tokens"])
    enc.backbone.load_state_dict(w[0], strict=True)
    pred.backbone.load_state_dict(w[1], strict=True)
    enc, pred = enc.to(dev), pred.to(dev)
    step = H.RefStep(R, enc, pred, FW.OPT, mixed_precision=mixed)
    out = step.step([clips.to(dev)], [[m.to(dev) for m in me]], [[m.to(dev) for m in mp]], grads=True, keep=True)
    acts = {"h": out["h"][0]}
    for j in range(len(me)):
        acts[f"z_enc.{j}"] = out["z_enc"][0][j]
        acts[f"z.{j}"] = out["z"][0][j]
    acts = {k: v.detach().float().cpu() for k, v in acts.items()}
    g_enc = {k: v.float().cpu() for k, v in out["grads"][0].items()}
    g_pred = {k: v.float().cpu() for k, v in out["grads"][1].items()}
    del step, enc, pred
    torch.cuda.empty_cache()
    return out["loss"], acts, g_enc, g_pred


def _exact_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


@needs_ref
@pytest.mark.parametrize("name", list(FW.CASES))
def test_fullwidth_vs_live_reference_fp32_and_autocast(name, dev):
    """(2): the same step against the reference's modules on this GPU.  Every quantity: our error vs the fp32
    reference, beside the error of the reference's own bf16-autocast path vs its fp32 self."""
    _exact_fp32()
    R = H.import_reference()
    w, clips = FW.weights(name), FW.clips()
    me, mp = our_masks()
    rme, rmp = FW.draw_masks(R.MaskCollator)
    assert all(torch.equal(a, b) for a, b in zip(me + mp, rme + rmp))
    l32, a32, ge32, gp32 = _ref_run(R, name, False, w, clips, me, mp, dev)
    l16, a16, ge16, gp16 = _ref_run(R, name, True, w, clips, me, mp, dev)
    loss, g_enc, g_pred, (h, z_enc, z) = ours_step(name, dev, w, clips, me, mp)
    rep = REPORT.setdefault("live", {}).setdefault(name, {})
    rep["loss"] = dict(ours=loss, ref_fp32=l32, ref_autocast=l16)
    assert abs(loss - l32) < 1e-3, (loss, l32, l16)
    acts = {"h": h, **{f"z_enc.{j}": t for j, t in enumerate(z_enc)}, **{f"z.{j}": t for j, t in enumerate(z)}}
    for k, a in acts.items():
        eo, ea = relerr(a, a32[k]), relerr(a16[k], a32[k])
        rep[k] = dict(ours=eo, autocast=ea)
        assert eo < 1e-2 and eo < 1.5 * ea + 1e-3, (k, eo, ea)
    worst = {}
    for tag, ours, r32, r16 in (("genc", g_enc, ge32, ge16), ("gpred", g_pred, gp32, gp16)):
        for n, g in ours.items():
            if n not in r32:
                assert float(g.abs().max()) == 0.0, n
                continue
            eo, ea = relerr(g, r32[n]), relerr(r16[n], r32[n])
            cls = tag + (".vec" if grad_tol(g) == 5e-2 else ".mat")
            wo = worst.setdefault(cls, dict(ours=0.0, autocast=0.0, name=""))
            if eo > wo["ours"]:
                wo.update(ours=eo, name=n)
            wo["autocast"] = max(wo["autocast"], ea)
            assert eo < grad_tol(g) and eo < 1.5 * ea + 1e-3, (n, eo, ea)
    rep["grads_worst"] = worst


@needs_ref
def test_full_depth_vitg_step_vs_live_reference(dev):
    """BASELINE config C3 at FULL depth (ViT-g/16, 40 blocks, predictor 12 blocks, 2 clips): loss and target / prediction
    features of the whole step against the reference on this GPU (fp32 and bf16 autocast)."""
    import vjepa_oracle as O
    _exact_fp32()
    R = H.import_reference()
    name = "vit_giant_xformers"
    D, heads, ratio, _ = FW.CASES[name]
    w = (O.init_encoder_weights(D, 40, ratio, seed=20), O.init_predictor_weights(D, 384, 12, 6, seed=21))
    clips = FW.clips(seed=3)
    me, mp = our_masks()
    l32, a32, ge32, _ = _ref_run(R, name, False, w, clips, me, mp, dev, depth=40, pred_depth=12)
    l16, a16, ge16, _ = _ref_run(R, name, True, w, clips, me, mp, dev, depth=40, pred_depth=12)
    loss, g_enc, _, (h, z_enc, z) = ours_step(name, dev, w, clips, me, mp, depth=40, pred_depth=12)
    rep = REPORT.setdefault("live_full_depth", {})
    rep["loss"] = dict(ours=loss, ref_fp32=l32, ref_autocast=l16)
    assert abs(loss - l32) < 1e-3, (loss, l32, l16)
    acts = {"h": h, **{f"z_enc.{j}": t for j, t in enumerate(z_enc)}, **{f"z.{j}": t for j, t in enumerate(z)}}
    for k, a in acts.items():
        eo, ea = relerr(a, a32[k]), relerr(a16[k], a32[k])
        rep[k] = dict(ours=eo, autocast=ea)
        assert eo < 1e-2 or eo < 1.5 * ea, (k, eo, ea)     # 40 blocks of bf16 rounding: bounded by the reference's own
    for n in ("blocks.0.attn.qkv.weight", "blocks.20.mlp.fc1.weight", "blocks.39.mlp.fc2.weight", "patch_embed.proj.weight"):
        eo, ea = relerr(g_enc[n], ge32[n]), relerr(ge16[n], ge32[n])
        rep["g." + n] = dict(ours=eo, autocast=ea)
        assert eo < 3e-2 or eo < 1.5 * ea, (n, eo, ea)


def _sdpa_ref(q, k, v, dout):
    """fp32 softmax(q k^T / sqrt(d)) v and its gradients, one (batch, head) at a time (S x S fp32 fits easily)."""
    B, Hh, S, d = q.shape
    outs, dqs, dks, dvs = [], [], [], []
    for b in range(B):
        for h in range(Hh):
            qq, kk, vv = (t[b, h].float().requires_grad_(True) for t in (q, k, v))
            p = torch.softmax(qq @ kk.t() * d ** -0.5, dim=-1)
            o = p @ vv
            o.backward(dout[b, h].float())
            outs.append(o.detach()), dqs.append(qq.grad), dks.append(kk.grad), dvs.append(vv.grad)
            del p, o
    f = lambda xs: torch.stack(xs).view(B, Hh, S, d)  # noqa: E731
    return f(outs), f(dqs), f(dks), f(dvs)


@pytest.mark.parametrize("S,d,heads", [(18432, 64, 2), (15525, 32, 3), (2048, 80, 2)])
def test_long_sequence_attention_vs_fp32_softmax(S, d, heads, dev):
    """Cooldown geometry (64 x 384^2: S = 18 432 encoder tokens at head_dim 64; ~15.5 k predictor tokens at head_dim 32,
    not a multiple of the tile) forward and backward against an fp32 softmax reference, and against torch SDPA bf16
    (what the reference calls, modules.py:369)."""
    from vjepa2_b200 import ops
    B, D = 1, heads * d
    g = torch.Generator().manual_seed(S + d)
    qkv = (torch.randn(B * S, 3 * D, generator=g) * 1.0).to(dev).bfloat16()
    dout = torch.randn(B * S, D, generator=g).to(dev).bfloat16()
    out = torch.empty(B * S, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B * heads * S, dtype=torch.float32, device=dev)
    dqkv = torch.empty_like(qkv)
    ops.attn_fwd(qkv, out, lse, B, S, heads, d)
    ops.attn_bwd(qkv, out, dout, lse, dqkv, B, S, heads, d)
    torch.cuda.synchronize()
    split = lambda t: t.view(B, S, 3, heads, d).permute(2, 0, 3, 1, 4)  # noqa: E731
    q, k, v = split(qkv)
    do = dout.view(B, S, heads, d).permute(0, 2, 1, 3)
    o32, dq32, dk32, dv32 = _sdpa_ref(q, k, v, do)
    got_o = out.view(B, S, heads, d).permute(0, 2, 1, 3)
    gq, gk, gv = split(dqkv)
    # torch SDPA in bf16 = the reference's kernel; its error against fp32 is the yardstick
    qs, ks, vs = (t.contiguous().requires_grad_(True) for t in (q, k, v))
    o16 = torch.nn.functional.scaled_dot_product_attention(qs, ks, vs)
    o16.backward(do.contiguous())
    rep = REPORT.setdefault("attention", {}).setdefault(f"S{S}_d{d}", {})
    for nm, ours, ref, sd in (("out", got_o, o32, o16), ("dq", gq, dq32, qs.grad), ("dk", gk, dk32, ks.grad),
                              ("dv", gv, dv32, vs.grad)):
        eo, es = relerr(ours, ref), relerr(sd, ref)
        rep[nm] = dict(ours=eo, torch_sdpa_bf16=es)
        assert eo < 1e-2 and eo < 1.5 * es + 1e-3, (nm, eo, es)
