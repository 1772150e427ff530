"""Full-width parity cases (BASELINE.json configs C1-C3 widths: ViT-L / ViT-H / ViT-g, 16 x 256^2 clips, 2048-token grid,
predictor 384 / 12 heads, masks from the shipped YAML mask block) with the transformer stacks cut to a few blocks so
the fp32 reference also runs on CPU in seconds.  Shared by oracle/make_golden_fullwidth.py (runs the REAL reference,
writes tests/golden/ref_fullwidth.pt) and tests/test_gpu_fullwidth.py (rebuilds the same seeded inputs)."""
import torch

import vjepa_oracle as O

# configs/train/vitg16/pretrain-256px-16f.yaml:41-67
MASK_CFG = [
    dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None, full_complement=False),
    dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None, full_complement=False),
]
OPT = dict(ipe=300, epochs=800, ipe_scale=1.25, warmup=40, start_lr=1e-4, lr=5.25e-4, final_lr=5.25e-4,
           weight_decay=0.04, final_weight_decay=0.04, ema=(0.99925, 0.99925), loss_exp=1.0)
FRAMES, CROP, BATCH = 16, 256, 2
PRED = dict(dim=384, depth=3, heads=12, num_mask_tokens=6)
# name: (embed_dim, heads, mlp_ratio, blocks kept)
CASES = {
    "vit_large": (1024, 16, 4.0, 2),
    "vit_huge": (1280, 16, 4.0, 2),            # head_dim 80
    "vit_giant_xformers": (1408, 22, 48 / 11, 2),   # D = 11 * 128, Hm = 6144
}
SLICE = 4096       # elements of every gradient kept in the golden file
ROWS = 48          # token rows of every activation kept in the golden file


def clips(seed=0):
    return torch.randn(BATCH, 3, FRAMES, CROP, CROP, generator=torch.Generator().manual_seed(seed))


def weights(name):
    D, heads, ratio, depth = CASES[name]
    w_enc = O.init_encoder_weights(D, depth, ratio, seed=10, rand_bias=True)
    w_pred = O.init_predictor_weights(D, PRED["dim"], PRED["depth"], PRED["num_mask_tokens"], seed=11, rand_bias=True)
    return w_enc, w_pred


def draw_masks(collator_cls):
    """One draw of the shipped mask config, config seed 239 -- works with the reference's MaskCollator and the mirror."""
    torch.manual_seed(239)
    c = collator_cls(cfgs_mask=MASK_CFG, dataset_fpcs=[FRAMES], crop_size=(CROP, CROP), patch_size=(16, 16), tubelet_size=2)
    gens = getattr(c, "samplers", None) or c.mask_generators
    pairs = [g(BATCH) for g in gens[FRAMES]]
    return [e for e, _ in pairs], [p for _, p in pairs]


def row_sample(n_rows, seed):
    return torch.randperm(n_rows, generator=torch.Generator().manual_seed(seed))[:ROWS].sort().values


def grad_slice(g):
    g = g.detach().float().reshape(-1)
    step = max(1, g.numel() // SLICE)
    return g[::step][:SLICE].clone()
