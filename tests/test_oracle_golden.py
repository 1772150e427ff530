"""CPU: pin the oracle (oracle/vjepa_oracle.py) against vectors produced by the real reference
(oracle/make_golden.py -> tests/golden/ref_golden.pt)."""
import torch

import vjepa_oracle as O
from golden_common import GRID, NTOK, OPT_CFG, TINY, step_masks, tiny_clips, tiny_masks, tiny_weights


def close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_rope_rotate_fwd_bwd(golden):
    x = golden["rope.x"].clone().requires_grad_(True)
    y = O.rope_rotate(x, golden["rope.pos"])
    close(y, golden["rope.y"], 1e-12, 1e-12)
    (gx,) = torch.autograd.grad(y, x, golden["rope.gy"])
    close(gx, golden["rope.gx"], 1e-12, 1e-12)


def test_apply_masks(golden):
    x, m1, m2 = golden["am.x"], golden["am.m1"], golden["am.m2"]
    assert torch.equal(O.apply_masks(x, [m1, m1.flip(1)]), golden["am.cat"])
    assert torch.equal(O.apply_masks(x, [m2], concat=False)[0], golden["am.list1"])


def test_encoder_forward(golden):
    w_enc, _ = tiny_weights()
    clips = tiny_clips(2)
    me, _ = tiny_masks(2)
    pe = O.patch_embed3d(clips, w_enc["patch_embed.proj.weight"], w_enc["patch_embed.proj.bias"])
    close(pe, golden["enc.patch_embed"])
    b0 = O.block(pe, w_enc, "blocks.0.", TINY["heads"], torch.arange(NTOK), GRID, GRID)
    close(b0, golden["enc.block0"])
    close(O.vit_forward(w_enc, clips, None, TINY["depth"], TINY["heads"]), golden["enc.full"])
    close(O.vit_forward(w_enc, clips, me, TINY["depth"], TINY["heads"]), golden["enc.masked"])


def test_encoder_forward_head_dim_80(golden):
    """ViT-H's head_dim (80): RoPE segment width 26, two pass-through dims."""
    w = O.init_encoder_weights(160, 2, 4.0, seed=3, rand_bias=True)
    clips = tiny_clips(2)
    me, _ = tiny_masks(2)
    close(O.vit_forward(w, clips, None, 2, 2), golden["encH.full"])
    close(O.vit_forward(w, clips, me, 2, 2), golden["encH.masked"])


def test_predictor_forward(golden):
    _, w_pred = tiny_weights()
    me, mp = tiny_masks(2)
    for idx, key in ((0, "pred.out"), (1, "pred.out_idx1")):
        out = O.predictor_forward(w_pred, golden["enc.masked"], me, mp, TINY["pred_depth"], TINY["pred_heads"],
                                  GRID, NTOK, mask_index=idx, num_mask_tokens=TINY["num_mask_tokens"])
        close(out, golden[key])


def test_mask_generator_bit_exact(golden):
    gens = O.make_mask_generators(O.DEFAULT_MASK_CFG, (256, 256), 16)
    torch.manual_seed(239)
    for it in range(3):
        for j, gen in enumerate(gens):
            e, p = gen(6)
            assert e.dtype == torch.int64
            assert torch.equal(e, golden[f"mask.it{it}.enc{j}"])
            assert torch.equal(p, golden[f"mask.it{it}.pred{j}"])
    gens = O.make_mask_generators(O.DEFAULT_MASK_CFG, (384, 384), 64)
    torch.manual_seed(7)
    for j, gen in enumerate(gens):
        e, p = gen(2)
        assert torch.equal(e, golden[f"mask384.enc{j}"])
        assert torch.equal(p, golden[f"mask384.pred{j}"])


def test_schedules(golden):
    lr = [O.warmup_cosine_lr(s, 4, 1e-4, 5.25e-4, 1e-5, 20) for s in range(1, 25)]
    wd = [O.cosine_wd(s, 0.04, 0.4, 20) for s in range(1, 25)]
    close(torch.tensor(lr, dtype=torch.float64), golden["sched.lr"], 1e-12, 0)
    close(torch.tensor(wd, dtype=torch.float64), golden["sched.wd"], 1e-12, 0)


def test_train_step(golden, golden_infer):
    w_enc, w_pred = tiny_weights()
    st = O.StepState(w_enc, w_pred, dict(depth=TINY["depth"], heads=TINY["heads"]),
                     dict(depth=TINY["pred_depth"], heads=TINY["pred_heads"], grid_size=GRID, num_patches=NTOK,
                          num_mask_tokens=TINY["num_mask_tokens"]), OPT_CFG)
    clips = tiny_clips(2)
    masks_enc, masks_pred = step_masks()
    loss0, g_enc, g_pred, _, _ = O.train_step(st, clips, masks_enc, masks_pred, return_grads=True)
    close(torch.tensor(loss0), golden["step.loss0"], 1e-5, 1e-6)
    for k, v in golden.items():
        if k.startswith("step.genc."):
            close(g_enc[k[len("step.genc."):]], v, 1e-3, 1e-7)
        if k.startswith("step.gpred."):
            close(g_pred[k[len("step.gpred."):]], v, 1e-3, 1e-7)
    assert g_pred["mask_tokens.1"] is None
    loss1 = O.train_step(st, clips, masks_enc, masks_pred)
    close(torch.tensor(loss1), golden["step.loss1"], 1e-4, 1e-6)
    for k, v in golden.items():
        if k.startswith("step.after.enc."):
            close(st.w_enc[k[len("step.after.enc."):]], v, 1e-4, 1e-6)
        if k.startswith("step.after.tgt."):
            close(st.w_tgt[k[len("step.after.tgt."):]], v, 1e-4, 1e-6)
        if k.startswith("step.after.pred."):
            close(st.w_pred[k[len("step.after.pred."):]], v, 1e-4, 1e-6)
    # AdamW moments after the two steps (what a checkpoint's "opt" entry carries)
    assert float(golden_infer["opt.step"]) == st.step == 2
    for k, v in golden_infer.items():
        if k.startswith("opt.enc.exp_avg_sq."):
            close(st.adam_enc[k[len("opt.enc.exp_avg_sq."):]][1], v, 2e-3, 1e-12)
        elif k.startswith("opt.enc.exp_avg."):
            close(st.adam_enc[k[len("opt.enc.exp_avg."):]][0], v, 1e-3, 1e-8)


def test_encoder_out_layers(golden_infer):
    w_enc, _ = tiny_weights()
    outs = O.vit_forward(w_enc, tiny_clips(2), None, TINY["depth"], TINY["heads"], out_layers=[0, 1])
    assert len(outs) == 2
    close(outs[0], golden_infer["infer.out_layers.0"])
    close(outs[1], golden_infer["infer.out_layers.1"])
