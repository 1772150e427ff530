"""Seeded input builders shared by the golden tests (mirror oracle/make_golden.py)."""
import torch

import vjepa_oracle as O

TINY = dict(img=96, frames=8, patch=16, tubelet=2, dim=128, depth=2, heads=2, mlp_ratio=4.0,
            pred_dim=64, pred_depth=2, pred_heads=2, num_mask_tokens=2)
GRID = TINY["img"] // 16
NTOK = (TINY["frames"] // 2) * GRID * GRID

OPT_CFG = dict(ipe=10, epochs=2, ipe_scale=1.25, warmup=0.2, start_lr=1e-4, lr=5.25e-4, final_lr=1e-5,
               weight_decay=0.04, final_weight_decay=0.4, ema=(0.99, 1.0), loss_exp=1.0)


def tiny_clips(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, TINY["frames"], TINY["img"], TINY["img"], generator=g)


def tiny_masks(B, seed=5):
    g = torch.Generator().manual_seed(seed)
    me, mp = [], []
    for _ in range(B):
        perm = torch.randperm(NTOK, generator=g)
        me.append(perm[:40].sort().values)
        mp.append(perm[40:40 + 72].sort().values)
    return torch.stack(me), torch.stack(mp)


def tiny_weights():
    t = TINY
    w_enc = O.init_encoder_weights(t["dim"], t["depth"], t["mlp_ratio"], seed=0, rand_bias=True)
    w_pred = O.init_predictor_weights(t["dim"], t["pred_dim"], t["pred_depth"], t["num_mask_tokens"],
                                      seed=1, rand_bias=True)
    return w_enc, w_pred


def step_masks():
    me, mp = tiny_masks(2)
    me2, mp2 = tiny_masks(2, seed=6)
    return [me, me2[:, :24]], [mp, mp2[:, :88]]
