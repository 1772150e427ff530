"""Compact FULL-WIDTH golden vectors from the REAL reference (fp32, CPU): tests/golden/ref_fullwidth.pt.

    python oracle/make_golden_fullwidth.py        # build container only (needs /root/reference or baseline/_ref)

Test infrastructure.  For every case of tests/fullwidth_common.py (ViT-L / ViT-H / ViT-g widths, 2 blocks, 16 x 256^2
clips, B = 2, predictor 384 / 12 heads, masks of the shipped mask config drawn by the reference's MaskCollator) it
runs one full reference train step (baseline/ref_harness.RefStep = train.py:409-471 around the reference's modules)
and keeps: the mask indices, the loss, ROWS sampled token rows + all per-row norms of h / encoder outputs /
predictions, the norm and a strided SLICE-element sample of every parameter gradient.  A few MB in total, so the
GPU parity test also runs where the reference tree is absent."""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "baseline")]

import fullwidth_common as FW  # noqa: E402
import ref_harness as H  # noqa: E402


def run_case(R, name, G):
    D, heads, ratio, depth = FW.CASES[name]
    enc, pred = H.build_models(R, name, crop=FW.CROP, frames=FW.FRAMES, depth=depth, pred_depth=FW.PRED["depth"],
                               pred_heads=FW.PRED["heads"], pred_dim=FW.PRED["dim"],
                               num_mask_tokens=FW.PRED["num_mask_tokens"])
    w_enc, w_pred = FW.weights(name)
    enc.backbone.load_state_dict(w_enc, strict=True)
    pred.backbone.load_state_dict(w_pred, strict=True)
    me, mp = FW.draw_masks(R.MaskCollator)
    clips = FW.clips()
    step = H.RefStep(R, enc, pred, FW.OPT, mixed_precision=False)
    t0 = time.time()
    out = step.step([clips], [me], [mp], grads=True, keep=True)
    print(f"{name}: loss {out['loss']:.6f}  K_enc {[m.shape[1] for m in me]} K_pred {[m.shape[1] for m in mp]}"
          f"  ({time.time() - t0:.1f}s)")
    p = f"{name}."
    G[p + "loss"] = torch.tensor(out["loss"], dtype=torch.float64)
    for j in range(len(me)):
        G[p + f"masks_enc.{j}"] = me[j].to(torch.int16)
        G[p + f"masks_pred.{j}"] = mp[j].to(torch.int16)
    acts = {"h": out["h"][0]}
    for j in range(len(me)):
        acts[f"z_enc.{j}"] = out["z_enc"][0][j]
        acts[f"z.{j}"] = out["z"][0][j]
    for k, a in acts.items():
        a2 = a.detach().float().reshape(-1, a.shape[-1])
        G[p + k + ".rownorm"] = a2.norm(dim=1)
        G[p + k + ".rows"] = a2[FW.row_sample(a2.shape[0], seed=len(k))].clone()
    g_enc, g_pred = out["grads"]
    for tag, gd in (("genc", g_enc), ("gpred", g_pred)):
        for n, g in gd.items():
            G[p + f"{tag}.{n}.norm"] = g.float().norm().double()
            G[p + f"{tag}.{n}.slice"] = FW.grad_slice(g)
    assert "mask_tokens.1" not in g_pred and "mask_tokens.0" in g_pred


def main():
    torch.set_num_threads(os.cpu_count())
    R = H.import_reference()
    G = {}
    for name in FW.CASES:
        run_case(R, name, G)
    path = os.path.join(ROOT, "tests", "golden", "ref_fullwidth.pt")
    torch.save(G, path)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB, {len(G)} entries")


if __name__ == "__main__":
    main()
