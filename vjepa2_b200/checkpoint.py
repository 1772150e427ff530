"""Checkpoint wire format of the reference, read and written around the flat buffers of the fused step.

A reference checkpoint (app/vjepa/train.py:315-333) is one `torch.save`d dict:

    encoder / predictor / target_encoder : state_dict of the (DDP-wrapped) MultiSeqWrapper, i.e. keys
                                           `module.backbone.<param>` (`backbone.<param>` without DDP)
    opt                                  : torch.optim.AdamW.state_dict() over the four groups of
                                           init_opt (app/vjepa/utils.py:224-239)
    scaler                               : torch.cuda.amp.GradScaler.state_dict() or None
    epoch, loss, batch_size, world_size, lr

`save_checkpoint` emits exactly that from a `JepaTrainStep` (moments come out of the flat Adam buffers as
per-parameter tensors in the optimizer's own index order), `load_checkpoint` (app/vjepa/utils.py:90-135)
restores it, so a run can move between the reference loop and this step in either direction.  The released
weights (`vitl.pt` ... `vitg-384.pt`, src/hub/backbones.py:22-28,129-134) load through
`load_pretrained`, which strips the `module.` / `backbone.` prefixes the same way.

Pure host logic: tensors move with `copy_`, nothing is computed here.
"""
from __future__ import annotations

import torch

GROWTH_FACTOR, BACKOFF_FACTOR, GROWTH_INTERVAL = 2.0, 0.5, 2000   # torch.cuda.amp.GradScaler() defaults


def clean_backbone_key(state_dict):
    """src/hub/backbones.py:22-28 -- drop the DDP / MultiSeqWrapper prefixes."""
    return {k.replace("module.", "").replace("backbone.", ""): v for k, v in state_dict.items()}


def _unwrap(m):
    m = m.module if hasattr(m, "module") else m
    return m.backbone if hasattr(m, "backbone") else m


def _prefix_of(m):
    pfx = ""
    if hasattr(m, "module"):
        pfx += "module."
        m = m.module
    if hasattr(m, "backbone"):
        pfx += "backbone."
    return pfx


def opt_param_groups(encoder, predictor):
    """The four AdamW groups of init_opt (app/vjepa/utils.py:224-237) as lists of (name, parameter), in the
    optimizer's index order: encoder matrices, predictor matrices, encoder bias / 1-D, predictor bias / 1-D."""
    e, p = list(_unwrap(encoder).named_parameters()), list(_unwrap(predictor).named_parameters())

    def decayed(n, t):
        return ("bias" not in n) and (t.dim() != 1)

    return [[(n, t) for n, t in e if decayed(n, t)], [(n, t) for n, t in p if decayed(n, t)],
            [(n, t) for n, t in e if not decayed(n, t)], [(n, t) for n, t in p if not decayed(n, t)]]


def build_opt_state_dict(groups, moments, step, lr, wd, betas=(0.9, 0.999), eps=1e-8):
    """AdamW.state_dict() for `groups` (from opt_param_groups).  `moments(p)` returns (exp_avg, exp_avg_sq)
    views of parameter p, or None for a parameter that never received a gradient (torch keeps no state for
    those: the unused predictor mask tokens, train.py:280 `find_unused_parameters`).  The group metadata is
    produced by a real torch.optim.AdamW so that it matches the installed torch's key set."""
    meta = torch.optim.AdamW(
        [{"params": [t for _, t in groups[0]]}, {"params": [t for _, t in groups[1]]},
         {"params": [t for _, t in groups[2]], "WD_exclude": True, "weight_decay": 0},
         {"params": [t for _, t in groups[3]], "WD_exclude": True, "weight_decay": 0}],
        betas=betas, eps=eps).state_dict()
    state, idx = {}, 0
    for gi, grp in enumerate(groups):
        pg = meta["param_groups"][gi]
        pg["lr"] = lr
        if not pg.get("WD_exclude", False):
            pg["weight_decay"] = wd                       # CosineWDSchedule writes it into the group (schedulers.py:88-91)
        for _, t in grp:
            mv = moments(t) if step > 0 else None
            if mv is not None:
                state[idx] = {"step": torch.tensor(float(step)), "exp_avg": mv[0].detach().clone(),
                              "exp_avg_sq": mv[1].detach().clone()}
            idx += 1
    meta["state"] = state
    return meta


def restore_opt_state(groups, opt_sd, moments):
    """Inverse of build_opt_state_dict: copy the per-parameter moments of an AdamW.state_dict() into the views
    `moments(p)` returns.  Returns the optimizer step count (max over parameters; 0 for a fresh optimizer)."""
    saved = opt_sd["param_groups"]
    if len(saved) != len(groups):
        raise ValueError(f"checkpoint optimizer has {len(saved)} param groups, expected {len(groups)}")
    step = 0
    for gi, grp in enumerate(groups):
        ids = saved[gi]["params"]
        if len(ids) != len(grp):
            raise ValueError(f"optimizer group {gi}: checkpoint has {len(ids)} parameters, the model {len(grp)}")
        for pid, (name, t) in zip(ids, grp):
            mv = moments(t)
            s = opt_sd["state"].get(pid)
            if s is None:
                if mv is not None:
                    mv[0].zero_()
                    mv[1].zero_()
                continue
            if tuple(s["exp_avg"].shape) != tuple(t.shape):
                raise ValueError(f"optimizer state of {name}: shape {tuple(s['exp_avg'].shape)} != {tuple(t.shape)}")
            mv[0].copy_(s["exp_avg"])
            mv[1].copy_(s["exp_avg_sq"])
            step = max(step, int(float(s["step"])))
    return step


DEFAULT_PREFIX = "module.backbone."


def _wrapped_state_dict(m, prefix):
    sd = _unwrap(m).state_dict()
    return {prefix + k: v.detach().cpu().clone() for k, v in sd.items()}


def _moments_of(step, skip_frozen=True):
    def moments(t):
        for rt in (step.enc_rt, step.pred_rt):
            fs = rt.fs
            if id(t) in fs.index:
                m, v = fs._view(fs.exp_avg, t), fs._view(fs.exp_avg_sq, t)
                # torch keeps no state for a parameter that never received a gradient (unused mask tokens)
                if skip_frozen and fs.is_frozen(t) and not bool(v.any()):
                    return None
                return m, v
        raise KeyError("parameter is not owned by this train step")
    return moments


def scaler_state_dict(step):
    """GradScaler.state_dict() (None when mixed precision is off, train.py:322)."""
    if not step.mixed_precision:
        return None
    return {"scale": float(step.scale.item()), "growth_factor": GROWTH_FACTOR, "backoff_factor": BACKOFF_FACTOR,
            "growth_interval": GROWTH_INTERVAL, "_growth_tracker": int(step.growth_tracker.item())}


def save_checkpoint(path, step, epoch, *, loss=0.0, batch_size=None, world_size=None, lr=None, encoder=None,
                    predictor=None, target_encoder=None):
    """train.py:315-333.  `step` is the JepaTrainStep; pass the (possibly wrapped) modules to reproduce their key
    prefixes, otherwise `module.backbone.` is written: app/vjepa/train.py always wraps the three models in
    DistributedDataParallel(MultiSeqWrapper(...)) (train.py:279-281) and its load_checkpoint does a strict
    load_state_dict on those (app/vjepa/utils.py:104-118), so that is the only prefix the reference loop accepts."""
    groups = opt_param_groups(step.encoder, step.predictor)
    cur_lr, cur_wd = step.last_lr_wd
    save_dict = {
        "encoder": _wrapped_state_dict(step.encoder, _prefix_of(encoder) if encoder is not None else DEFAULT_PREFIX),
        "predictor": _wrapped_state_dict(step.predictor,
                                         _prefix_of(predictor) if predictor is not None else DEFAULT_PREFIX),
        "opt": build_opt_state_dict(groups, _moments_of(step), step.optimizer_steps(), cur_lr, cur_wd, step.betas, step.eps),
        "scaler": scaler_state_dict(step),
        "target_encoder": _wrapped_state_dict(step.target_encoder,
                                              _prefix_of(target_encoder) if target_encoder is not None else DEFAULT_PREFIX),
        "epoch": epoch,
        "loss": loss,
        "batch_size": batch_size,
        "world_size": step.world if world_size is None else world_size,
        "lr": lr,
    }
    for k in ("opt",):
        for s in save_dict[k]["state"].values():
            s["exp_avg"], s["exp_avg_sq"] = s["exp_avg"].cpu(), s["exp_avg_sq"].cpu()
    torch.save(save_dict, path)
    return save_dict


def load_checkpoint(r_path, step, *, fast_forward=True, ipe=None):
    """app/vjepa/utils.py:90-135 plus the scheduler fast-forward of train.py:309-313: restores encoder, predictor,
    target encoder, AdamW moments / step count and the GradScaler, then (fast_forward) advances the LR / WD /
    momentum schedules by `epoch * ipe` iterations.  Returns the epoch."""
    ckpt = r_path if isinstance(r_path, dict) else torch.load(r_path, map_location="cpu", weights_only=False)
    epoch = ckpt["epoch"]
    for key, model in (("encoder", step.encoder), ("predictor", step.predictor),
                       ("target_encoder", step.target_encoder)):
        model.load_state_dict(clean_backbone_key(ckpt[key]))         # strict, like the reference
    step.reload_weights()
    groups = opt_param_groups(step.encoder, step.predictor)
    step.set_optimizer_steps(restore_opt_state(groups, ckpt["opt"], _moments_of(step, skip_frozen=False)))
    sc = ckpt.get("scaler")
    if sc is not None and step.mixed_precision:
        step.set_scaler(float(sc["scale"]), int(sc["_growth_tracker"]))
    if fast_forward:
        step.fast_forward(int(epoch) * int(ipe if ipe is not None else step.ipe))
    return epoch


def load_pretrained(model, path_or_dict, checkpoint_key="target_encoder", strict=False):
    """Released / pre-trained weights into an encoder or predictor (src/hub/backbones.py:129-134;
    evals/video_classification_frozen/modelcustom/vit_encoder_multiclip.py:58-70): take `checkpoint_key`, strip the
    prefixes, keep the model's own tensor where a key is missing or has another shape (`pos_embed` of the
    sincos checkpoints), load non-strictly.  Returns the torch `load_state_dict` message."""
    ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu",
                                                                          weights_only=False)
    sd = ckpt[checkpoint_key] if checkpoint_key in ckpt else ckpt
    sd = clean_backbone_key(sd)
    target = _unwrap(model)
    for k, v in target.state_dict().items():
        if k in sd and tuple(sd[k].shape) != tuple(v.shape):
            sd[k] = v
    return target.load_state_dict(sd, strict=strict)
