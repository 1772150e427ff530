"""CPU oracle for the V-JEPA 2 pre-training step  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (fp32 / fp64, CPU) *restatement* of the reference
algorithm on the hot path.  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  Nothing under `vjepa2_b200/` imports it, and the
product path never falls back to it.

Parity status: PINNED.  `oracle/make_golden.py` imports the real reference
modules from /root/reference (with a 3-line stub for the absent `timm`
dependency), runs them on seeded inputs and commits input/output vectors to
`tests/golden/`; `tests/test_oracle_golden.py` checks every function here
against those vectors (the reference's own test-suite has shape checks only,
SURVEY.md section 4 / 8c).

All file:line citations are into /root/reference.  Weights are passed as a
flat dict keyed by the reference `state_dict()` names so a reference checkpoint
feeds this file unchanged.
"""
from __future__ import annotations

import math
from multiprocessing import Value

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# configs (src/models/vision_transformer.py:275-314, src/models/predictor.py:249)
# ----------------------------------------------------------------------------
ENCODER_CFGS = {
    # name: (embed_dim, depth, heads, mlp_ratio)
    "vit_tiny_test": (128, 2, 2, 4.0),       # test-only size (head_dim 64)
    "vit_large": (1024, 24, 16, 4.0),
    "vit_huge": (1280, 32, 16, 4.0),
    "vit_giant_xformers": (1408, 40, 22, 48 / 11),
}


def mlp_hidden(dim, mlp_ratio):
    return int(dim * mlp_ratio)  # modules.py:551


# ----------------------------------------------------------------------------
# apply_masks  (src/masks/utils.py:9-21)
# ----------------------------------------------------------------------------
def apply_masks(x, masks, concat=True):
    out = []
    for m in masks:
        idx = m.unsqueeze(-1).expand(-1, -1, x.size(-1))
        out.append(torch.gather(x, 1, idx))
    return torch.cat(out, 0) if concat else out


# ----------------------------------------------------------------------------
# PatchEmbed3D (src/models/utils/patch_embed.py:26-52): Conv3d with
# kernel == stride == (tubelet, p, p) restated as im2col + GEMM.
# ----------------------------------------------------------------------------
def im2col_tubelets(x, tubelet=2, patch=16):
    B, C, T, H, W = x.shape
    t, h, w = T // tubelet, H // patch, W // patch
    x = x.reshape(B, C, t, tubelet, h, patch, w, patch)
    # token order: t-major, then h, then w; K order (c, kt, kh, kw)
    x = x.permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B, t * h * w, C * tubelet * patch * patch)
    return x


def patch_embed3d(x, weight, bias, tubelet=2, patch=16):
    cols = im2col_tubelets(x, tubelet, patch)
    return cols @ weight.reshape(weight.shape[0], -1).t() + bias


# ----------------------------------------------------------------------------
# RoPE (src/models/utils/modules.py:26-50, 285-287, 311-365)
# ----------------------------------------------------------------------------
def separate_positions(ids, Hp, Wp):
    tpf = int(Hp * Wp)
    f = ids // tpf
    r = ids - tpf * f
    y = r // Wp
    x = r - Wp * y
    return f, y, x


def rope_segment_width(head_dim):
    return int(2 * ((head_dim // 3) // 2))  # modules.py:285-287


def rope_angles(pos, seg):
    """theta[..., j] = pos * 10000^(-j/(seg/2)), j in [0, seg/2)  (modules.py:31-34)."""
    half = seg // 2
    omega = torch.arange(half, dtype=torch.float64) / (seg / 2.0)
    omega = 1.0 / 10000 ** omega
    return pos.to(torch.float64).unsqueeze(-1) * omega


def rope_rotate(x, pos):
    """x [..., N, seg]; pos [..., N] (broadcastable).  modules.py:26-50:
    sin/cos tiled x2, pair rotation interleaved."""
    seg = x.shape[-1]
    th = rope_angles(pos, seg)
    sin = torch.cat([th.sin(), th.sin()], -1).to(x.dtype)
    cos = torch.cat([th.cos(), th.cos()], -1).to(x.dtype)
    y = x.unflatten(-1, (-1, 2))
    y1, y2 = y.unbind(-1)
    y = torch.stack((-y2, y1), -1).flatten(-2)
    return x * cos + y * sin


def rope_qk(q, ids, Hp, Wp):
    """q [B, heads, N, d]; ids [B, N] or [N] token ids."""
    d = q.shape[-1]
    s = rope_segment_width(d)
    f, y, x = separate_positions(ids, Hp, Wp)
    if ids.dim() == 2:
        f, y, x = f[:, None], y[:, None], x[:, None]
    parts = [
        rope_rotate(q[..., 0:s], f),
        rope_rotate(q[..., s:2 * s], y),
        rope_rotate(q[..., 2 * s:3 * s], x),
    ]
    if 3 * s < d:
        parts.append(q[..., 3 * s:])
    return torch.cat(parts, -1)


def rope_attention(x, w, pfx, heads, ids, Hp, Wp):
    """RoPEAttention.forward (modules.py:326-382)."""
    B, N, C = x.shape
    d = C // heads
    qkv = x @ w[pfx + "qkv.weight"].t() + w[pfx + "qkv.bias"]
    qkv = qkv.unflatten(-1, (3, heads, d)).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = rope_qk(q, ids, Hp, Wp)
    k = rope_qk(k, ids, Hp, Wp)
    att = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    att = att.softmax(-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return o @ w[pfx + "proj.weight"].t() + w[pfx + "proj.bias"]


def mlp(x, w, pfx):
    """MLP.forward (modules.py:77-83): fc1 -> exact-erf GELU -> fc2."""
    h = x @ w[pfx + "fc1.weight"].t() + w[pfx + "fc1.bias"]
    h = F.gelu(h)
    return h @ w[pfx + "fc2.weight"].t() + w[pfx + "fc2.bias"]


def layer_norm(x, weight=None, bias=None, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def block(x, w, pfx, heads, ids, Hp, Wp):
    """Block.forward (modules.py:556-563)."""
    y = layer_norm(x, w[pfx + "norm1.weight"], w[pfx + "norm1.bias"])
    x = x + rope_attention(y, w, pfx + "attn.", heads, ids, Hp, Wp)
    y = layer_norm(x, w[pfx + "norm2.weight"], w[pfx + "norm2.bias"])
    return x + mlp(y, w, pfx + "mlp.")


# ----------------------------------------------------------------------------
# VisionTransformer.forward (src/models/vision_transformer.py:161-213), use_rope=True
# ----------------------------------------------------------------------------
def vit_forward(w, x, masks, depth, heads, tubelet=2, patch=16, out_layers=None):
    """VisionTransformer.forward (vision_transformer.py:161-213); out_layers -> list of norm(x) after those blocks."""
    if masks is not None and not isinstance(masks, list):
        masks = [masks]
    _, _, T, H, W = x.shape
    Hp, Wp = H // patch, W // patch
    x = patch_embed3d(x, w["patch_embed.proj.weight"], w["patch_embed.proj.bias"], tubelet, patch)
    if masks is not None:
        x = apply_masks(x, masks)
        ids = torch.cat(masks, 0)
    else:
        ids = torch.arange(x.shape[1])
    outs = []
    for i in range(depth):
        x = block(x, w, f"blocks.{i}.", heads, ids, Hp, Wp)
        if out_layers is not None and i in out_layers:
            outs.append(layer_norm(x, w["norm.weight"], w["norm.bias"]))
    if out_layers is not None:
        return outs
    return layer_norm(x, w["norm.weight"], w["norm.bias"])


# ----------------------------------------------------------------------------
# VisionTransformerPredictor.forward (src/models/predictor.py:166-246), use_rope=True
# ----------------------------------------------------------------------------
def predictor_forward(w, x, masks_x, masks_y, depth, heads, grid_size, num_patches,
                      mask_index=1, num_mask_tokens=2):
    if not isinstance(masks_x, list):
        masks_x = [masks_x]
    if not isinstance(masks_y, list):
        masks_y = [masks_y]
    B = len(x) // len(masks_x)
    x = x @ w["predictor_embed.weight"].t() + w["predictor_embed.bias"]
    _, N_ctxt, D = x.shape
    mask_index = mask_index % num_mask_tokens
    pred_tokens = w[f"mask_tokens.{mask_index}"].repeat(B, num_patches, 1)
    pred_tokens = apply_masks(pred_tokens, masks_y)
    x = x.repeat(len(masks_x), 1, 1)
    x = torch.cat([x, pred_tokens], 1)
    mx = torch.cat(masks_x, 0)
    my = torch.cat(masks_y, 0)
    masks = torch.cat([mx, my], 1)
    argsort = torch.argsort(masks, dim=1)
    masks = torch.stack([masks[i, row] for i, row in enumerate(argsort)], 0)
    x = torch.stack([x[i, row, :] for i, row in enumerate(argsort)], 0)
    for i in range(depth):
        x = block(x, w, f"predictor_blocks.{i}.", heads, masks, grid_size, grid_size)
    x = layer_norm(x, w["predictor_norm.weight"], w["predictor_norm.bias"])
    rev = torch.argsort(argsort, dim=1)
    x = torch.stack([x[i, row, :] for i, row in enumerate(rev)], 0)
    x = x[:, N_ctxt:]
    return x @ w["predictor_proj.weight"].t() + w["predictor_proj.bias"]


# ----------------------------------------------------------------------------
# loss / EMA / optimizer / schedules (app/vjepa/train.py:409-471, app/vjepa/utils.py:207-255,
# src/utils/schedulers.py:41-93)
# ----------------------------------------------------------------------------
def target_forward(w_tgt, clips, depth, heads, tubelet=2, patch=16):
    """forward_target (train.py:414-418): no-grad encoder + non-affine LN eps 1e-5."""
    with torch.no_grad():
        h = vit_forward(w_tgt, clips, None, depth, heads, tubelet, patch)
        return F.layer_norm(h, (h.size(-1),))


def jepa_loss(z_list, h, masks_pred, loss_exp=1.0):
    """loss_fn (train.py:425-435) for one fpc group: z_list[j] vs gathered targets."""
    hs = apply_masks(h, masks_pred, concat=False)
    loss, n = 0.0, 0
    for zj, hj in zip(z_list, hs):
        loss = loss + torch.mean(torch.abs(zj - hj) ** loss_exp) / loss_exp
        n += 1
    return loss / n


def ema_update(w_tgt, w_ctx, m):
    """train.py:457-465: p_k = m*p_k + (1-m)*p_q over every encoder parameter."""
    for k in w_tgt:
        w_tgt[k].mul_(m).add_(w_ctx[k], alpha=1 - m)


def warmup_cosine_lr(step, warmup_steps, start_lr, ref_lr, final_lr, T_max_total):
    """WarmupCosineSchedule.step() value after `step` calls (schedulers.py:41-68)."""
    T_max = T_max_total - warmup_steps
    if step < warmup_steps:
        progress = float(step) / float(max(1, warmup_steps))
        return start_lr + progress * (ref_lr - start_lr)
    progress = float(step - warmup_steps) / float(max(1, T_max))
    return max(final_lr, final_lr + (ref_lr - final_lr) * 0.5 * (1.0 + math.cos(math.pi * progress)))


def cosine_wd(step, ref_wd, final_wd, T_max):
    """CosineWDSchedule.step() value after `step` calls (schedulers.py:71-93)."""
    progress = step / T_max
    new_wd = final_wd + (ref_wd - final_wd) * 0.5 * (1.0 + math.cos(math.pi * progress))
    return max(final_wd, new_wd) if final_wd <= ref_wd else min(final_wd, new_wd)


def wd_applies(name, p):
    """init_opt grouping (app/vjepa/utils.py:224-237): decay iff not bias and not 1-D."""
    return ("bias" not in name) and (p.dim() != 1)


def adamw_step(params, grads, state, step, lr, wd, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.AdamW single step restated (decoupled decay, bias-corrected).
    params/grads/state: dict name -> tensor (state[name] = (exp_avg, exp_avg_sq)).
    Parameters whose grad is None are skipped, as torch does."""
    b1, b2 = betas
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        m, v = state[k]
        decay = wd if wd_applies(k, p) else 0.0
        p.mul_(1 - lr * decay)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


# ----------------------------------------------------------------------------
# Mask generator (src/masks/multiseq_multiblock3d.py:79-239), RNG-call-identical
# ----------------------------------------------------------------------------
class MaskGenerator:
    def __init__(self, crop_size=(224, 224), num_frames=16, spatial_patch_size=(16, 16),
                 temporal_patch_size=2, spatial_pred_mask_scale=(0.2, 0.8),
                 temporal_pred_mask_scale=(1.0, 1.0), aspect_ratio=(0.3, 3.0), npred=1,
                 max_context_frames_ratio=1.0, max_keep=None):
        if not isinstance(crop_size, tuple):
            crop_size = (crop_size,) * 2
        if not isinstance(spatial_patch_size, tuple):
            spatial_patch_size = (spatial_patch_size,) * 2
        self.height = crop_size[0] // spatial_patch_size[0]
        self.width = crop_size[1] // spatial_patch_size[1]
        self.duration = num_frames // temporal_patch_size
        self.aspect_ratio = aspect_ratio
        self.spatial_scale = spatial_pred_mask_scale
        self.temporal_scale = temporal_pred_mask_scale
        self.npred = npred
        self.max_context_duration = max(1, int(self.duration * max_context_frames_ratio))
        self.max_keep = max_keep
        self._itr_counter = Value("i", -1)

    def step(self):
        i = self._itr_counter
        with i.get_lock():
            i.value += 1
            return i.value

    def _block_size(self, g):
        r = torch.rand(1, generator=g).item()
        ts = self.temporal_scale[0] + r * (self.temporal_scale[1] - self.temporal_scale[0])
        t = max(1, int(self.duration * ts))
        r = torch.rand(1, generator=g).item()
        ss = self.spatial_scale[0] + r * (self.spatial_scale[1] - self.spatial_scale[0])
        keep = int(self.height * self.width * ss)
        r = torch.rand(1, generator=g).item()
        ar = self.aspect_ratio[0] + r * (self.aspect_ratio[1] - self.aspect_ratio[0])
        h = min(int(round(math.sqrt(keep * ar))), self.height)
        w = min(int(round(math.sqrt(keep / ar))), self.width)
        return t, h, w

    def _block_mask(self, size):
        t, h, w = size
        top = torch.randint(0, self.height - h + 1, (1,))
        left = torch.randint(0, self.width - w + 1, (1,))
        start = torch.randint(0, self.duration - t + 1, (1,))
        m = torch.ones((self.duration, self.height, self.width), dtype=torch.int32)
        m[start:start + t, top:top + h, left:left + w] = 0
        if self.max_context_duration < self.duration:
            m[self.max_context_duration:, :, :] = 0
        return m

    def __call__(self, batch_size):
        g = torch.Generator()
        g.manual_seed(self.step())
        size = self._block_size(g)
        enc, pred = [], []
        min_e = min_p = self.duration * self.height * self.width
        for _ in range(batch_size):
            while True:
                me = torch.ones((self.duration, self.height, self.width), dtype=torch.int32)
                for _ in range(self.npred):
                    me *= self._block_mask(size)
                me = me.flatten()
                mp = torch.argwhere(me == 0).squeeze()
                me = torch.nonzero(me).squeeze()
                if len(me) != 0:
                    break
            min_p = min(min_p, len(mp))
            min_e = min(min_e, len(me))
            pred.append(mp)
            enc.append(me)
        if self.max_keep is not None:
            min_e = min(min_e, self.max_keep)
        enc = torch.stack([m[:min_e] for m in enc], 0)
        pred = torch.stack([m[:min_p] for m in pred], 0)
        return enc, pred


def make_mask_generators(cfgs_mask, crop_size, num_frames, patch_size=16, tubelet=2):
    """MaskCollator.__init__ for one fpc (multiseq_multiblock3d.py:28-47)."""
    gens = []
    for m in cfgs_mask:
        gens.append(MaskGenerator(
            crop_size=crop_size, num_frames=num_frames,
            spatial_patch_size=(patch_size, patch_size), temporal_patch_size=tubelet,
            spatial_pred_mask_scale=m.get("spatial_scale"),
            temporal_pred_mask_scale=m.get("temporal_scale"),
            aspect_ratio=m.get("aspect_ratio"), npred=m.get("num_blocks"),
            max_context_frames_ratio=m.get("max_temporal_keep", 1.0),
            max_keep=m.get("max_keep", None)))
    return gens


# mask block of configs/train/vit{l,h,g}16/pretrain-256px-16f.yaml:41-67
DEFAULT_MASK_CFG = [
    dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15),
         temporal_scale=(1.0, 1.0), max_temporal_keep=1.0, max_keep=None),
    dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7),
         temporal_scale=(1.0, 1.0), max_temporal_keep=1.0, max_keep=None),
]


# ----------------------------------------------------------------------------
# random-init weights with the reference's parameter names/shapes
# (vision_transformer.py:130-153, predictor.py:149-164; trunc_normal_ bounds are +-2 absolute,
# i.e. a plain normal at std 0.02)
# ----------------------------------------------------------------------------
def _block_init(w, pfx, dim, hidden, layer_id, g, std):
    w[pfx + "norm1.weight"] = torch.ones(dim)
    w[pfx + "norm1.bias"] = torch.zeros(dim)
    w[pfx + "attn.qkv.weight"] = torch.randn(3 * dim, dim, generator=g) * std
    w[pfx + "attn.qkv.bias"] = torch.zeros(3 * dim)
    w[pfx + "attn.proj.weight"] = torch.randn(dim, dim, generator=g) * std / math.sqrt(2.0 * layer_id)
    w[pfx + "attn.proj.bias"] = torch.zeros(dim)
    w[pfx + "norm2.weight"] = torch.ones(dim)
    w[pfx + "norm2.bias"] = torch.zeros(dim)
    w[pfx + "mlp.fc1.weight"] = torch.randn(hidden, dim, generator=g) * std
    w[pfx + "mlp.fc1.bias"] = torch.zeros(hidden)
    w[pfx + "mlp.fc2.weight"] = torch.randn(dim, hidden, generator=g) * std / math.sqrt(2.0 * layer_id)
    w[pfx + "mlp.fc2.bias"] = torch.zeros(dim)


def init_encoder_weights(dim, depth, mlp_ratio, seed=0, std=0.02, tubelet=2, patch=16, rand_bias=False):
    g = torch.Generator().manual_seed(seed)
    w = {}
    w["patch_embed.proj.weight"] = torch.randn(dim, 3, tubelet, patch, patch, generator=g) * std
    w["patch_embed.proj.bias"] = torch.zeros(dim)
    for i in range(depth):
        _block_init(w, f"blocks.{i}.", dim, mlp_hidden(dim, mlp_ratio), i + 1, g, std)
    w["norm.weight"] = torch.ones(dim)
    w["norm.bias"] = torch.zeros(dim)
    if rand_bias:  # tests: make biases / LN affine non-trivial
        for k in w:
            if k.endswith("bias"):
                w[k] = torch.randn(w[k].shape, generator=g) * std
            elif "norm" in k and k.endswith("weight"):
                w[k] = 1.0 + torch.randn(w[k].shape, generator=g) * 0.1
    return w


def init_predictor_weights(embed_dim, pred_dim, depth, num_mask_tokens, seed=1, std=0.02,
                           mlp_ratio=4.0, rand_bias=False):
    g = torch.Generator().manual_seed(seed)
    w = {}
    w["predictor_embed.weight"] = torch.randn(pred_dim, embed_dim, generator=g) * std
    w["predictor_embed.bias"] = torch.zeros(pred_dim)
    for k in range(num_mask_tokens):
        w[f"mask_tokens.{k}"] = torch.zeros(1, 1, pred_dim)
    for i in range(depth):
        _block_init(w, f"predictor_blocks.{i}.", pred_dim, mlp_hidden(pred_dim, mlp_ratio), i + 1, g, std)
    w["predictor_norm.weight"] = torch.ones(pred_dim)
    w["predictor_norm.bias"] = torch.zeros(pred_dim)
    w["predictor_proj.weight"] = torch.randn(embed_dim, pred_dim, generator=g) * std
    w["predictor_proj.bias"] = torch.zeros(embed_dim)
    if rand_bias:
        for k in w:
            if k.endswith("bias") or k.startswith("mask_tokens"):
                w[k] = torch.randn(w[k].shape, generator=g) * std
            elif "norm" in k and k.endswith("weight"):
                w[k] = 1.0 + torch.randn(w[k].shape, generator=g) * 0.1
    return w


# ----------------------------------------------------------------------------
# the whole step (train.py:409-471), fp32, no GradScaler arithmetic (scale cancels
# exactly in fp32 up to rounding; inf-skip never triggers on finite fp32 grads)
# ----------------------------------------------------------------------------
class StepState:
    """Everything the step closure touches, as plain tensors."""

    def __init__(self, w_enc, w_pred, enc_cfg, pred_cfg, opt_cfg):
        self.w_enc = {k: v.clone() for k, v in w_enc.items()}
        self.w_tgt = {k: v.clone() for k, v in w_enc.items()}     # copy.deepcopy(encoder), train.py:210
        self.w_pred = {k: v.clone() for k, v in w_pred.items()}
        self.enc_cfg, self.pred_cfg, self.opt = enc_cfg, pred_cfg, opt_cfg
        self.adam_enc = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in self.w_enc.items()}
        self.adam_pred = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in self.w_pred.items()}
        self.step = 0


def train_step(st: StepState, clips, masks_enc, masks_pred, return_grads=False):
    """One fpc group, len(masks_enc) masks.  Returns loss (and grads)."""
    o = st.opt
    st.step += 1
    T_max = int(o["ipe_scale"] * o["epochs"] * o["ipe"])
    lr = warmup_cosine_lr(st.step, int(o["warmup"] * o["ipe"]), o["start_lr"], o["lr"], o["final_lr"], T_max)
    wd = cosine_wd(st.step, o["weight_decay"], o["final_weight_decay"], T_max)
    depth, heads = st.enc_cfg["depth"], st.enc_cfg["heads"]
    tub, patch = st.enc_cfg.get("tubelet", 2), st.enc_cfg.get("patch", 16)

    h = target_forward(st.w_tgt, clips, depth, heads, tub, patch)

    we = {k: v.detach().requires_grad_(True) for k, v in st.w_enc.items()}
    wp = {k: v.detach().requires_grad_(True) for k, v in st.w_pred.items()}
    zs = []
    for j, (me, mp) in enumerate(zip(masks_enc, masks_pred)):
        z = vit_forward(we, clips, me, depth, heads, tub, patch)
        z = predictor_forward(wp, z, me, mp, st.pred_cfg["depth"], st.pred_cfg["heads"],
                              st.pred_cfg["grid_size"], st.pred_cfg["num_patches"],
                              mask_index=0, num_mask_tokens=st.pred_cfg["num_mask_tokens"])
        zs.append(z)
    loss = jepa_loss(zs, h, masks_pred, o.get("loss_exp", 1.0))
    loss.backward()
    g_enc = {k: v.grad for k, v in we.items()}
    g_pred = {k: v.grad for k, v in wp.items()}   # unused mask tokens stay None -> skipped
    with torch.no_grad():
        adamw_step(st.w_enc, g_enc, st.adam_enc, st.step, lr, wd)
        adamw_step(st.w_pred, g_pred, st.adam_pred, st.step, lr, wd)
        # momentum_scheduler (train.py:286-289): i-th value, float denominator
        m = o["ema"][0] + (st.step - 1) * (o["ema"][1] - o["ema"][0]) / (o["ipe"] * o["epochs"] * o["ipe_scale"])
        ema_update(st.w_tgt, st.w_enc, m)
    if return_grads:
        return float(loss.detach()), g_enc, g_pred, lr, wd
    return float(loss.detach())
