"""Mask application and generation: src/masks/utils.py:9-21 (apply_masks) and
src/masks/multiseq_multiblock3d.py:16-239 (MaskCollator, _MaskGenerator).

apply_masks runs on the device as a 128-bit row-copy gather (bit-exact).  The generator is host-side
integer work, kept RNG-call-identical to the reference (same torch CPU generator calls in the same
order), so identical RNG state in gives bit-identical indices out.
"""
from __future__ import annotations

import math
from multiprocessing import Value

import torch

from . import ops


class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rows):
        B, N, D = x.shape
        out = torch.empty(rows.numel(), D, dtype=x.dtype, device=x.device)
        ops.gather_rows(x.reshape(B * N, D), out, rows)
        ctx.save_for_backward(rows)
        ctx.shape = (B, N, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        (rows,) = ctx.saved_tensors
        B, N, D = ctx.shape
        dx = torch.zeros(B * N, D, dtype=torch.float32, device=dout.device)
        ops.scatter_add_rows(dout.contiguous(), dx, rows)   # duplicate indices are legal -> atomics
        return dx.view(B, N, D).to(dout.dtype), None


def apply_masks(x, masks, concat=True):
    """x [B, N, D] (fp32 or bf16, CUDA); masks: list of int64 [B, K] -> [B*len(masks), K, D] or list."""
    if not x.is_cuda:
        raise RuntimeError("vjepa2_b200.apply_masks: CUDA tensors only (no CPU path)")
    if x.dtype not in (torch.float32, torch.bfloat16) or x.shape[-1] % 8 != 0:
        raise NotImplementedError("vjepa2_b200.apply_masks: fp32/bf16 payload with feature dim % 8 == 0")
    B, N, D = x.shape
    x = x.contiguous()
    outs = []
    for m in masks:
        m = m.to(device=x.device, dtype=torch.int64).contiguous()
        rows = ops.mask_to_rows(m, N)
        if torch.is_grad_enabled() and x.requires_grad:
            o = _GatherFn.apply(x, rows)
        else:
            o = torch.empty(rows.numel(), D, dtype=x.dtype, device=x.device)
            ops.gather_rows(x.view(B * N, D), o, rows)
        outs.append(o.view(m.shape[0], m.shape[1], D))
    if not concat:
        return outs
    return torch.cat(outs, dim=0)


class MaskCollator(object):
    def __init__(self, cfgs_mask, dataset_fpcs, crop_size=(224, 224), patch_size=(16, 16), tubelet_size=2):
        self.mask_generators = dict()
        for fpc in dataset_fpcs:
            self.mask_generators[fpc] = []
            for m in cfgs_mask:
                self.mask_generators[fpc].append(_MaskGenerator(
                    crop_size=crop_size, num_frames=fpc, spatial_patch_size=patch_size,
                    temporal_patch_size=tubelet_size, spatial_pred_mask_scale=m.get("spatial_scale"),
                    temporal_pred_mask_scale=m.get("temporal_scale"), aspect_ratio=m.get("aspect_ratio"),
                    npred=m.get("num_blocks"), max_context_frames_ratio=m.get("max_temporal_keep", 1.0),
                    max_keep=m.get("max_keep", None), full_complement=m.get("full_complement", False),
                    pred_full_complement=m.get("pred_full_complement", False), inv_block=m.get("inv_block", False)))

    def step(self):
        for fpc in self.mask_generators:
            for g in self.mask_generators[fpc]:
                g.step()

    def __call__(self, batch):
        filtered = {fpc: [] for fpc in self.mask_generators}
        for sample in batch:
            filtered[len(sample[-1][-1])] += [sample]
        out = []
        for fpc, fpc_batch in filtered.items():
            if len(fpc_batch) == 0:
                continue
            collated = torch.utils.data.default_collate(fpc_batch)
            enc, pred = [], []
            for g in self.mask_generators[fpc]:
                me, mp = g(len(fpc_batch))
                enc.append(me)
                pred.append(mp)
            out += [(collated, enc, pred)]
        return out

    def draw(self, fpc, batch_size):
        """Masks only (what the step consumes): ([masks_enc per cfg], [masks_pred per cfg])."""
        enc, pred = [], []
        for g in self.mask_generators[fpc]:
            me, mp = g(batch_size)
            enc.append(me)
            pred.append(mp)
        return enc, pred


class _MaskGenerator(object):
    def __init__(self, crop_size=(224, 224), num_frames=16, spatial_patch_size=(16, 16), temporal_patch_size=2,
                 spatial_pred_mask_scale=(0.2, 0.8), temporal_pred_mask_scale=(1.0, 1.0), aspect_ratio=(0.3, 3.0),
                 npred=1, max_context_frames_ratio=1.0, max_keep=None, inv_block=False, full_complement=False,
                 pred_full_complement=False):
        if not isinstance(crop_size, tuple):
            crop_size = (crop_size,) * 2
        if not isinstance(spatial_patch_size, tuple):
            spatial_patch_size = (spatial_patch_size,) * 2
        self.crop_size = crop_size
        self.height, self.width = [crop_size[i] // spatial_patch_size[i] for i in (0, 1)]
        self.duration = num_frames // temporal_patch_size
        self.full_complement = full_complement
        self.pred_full_complement = pred_full_complement
        self.aspect_ratio = aspect_ratio
        self.spatial_pred_mask_scale = spatial_pred_mask_scale
        self.temporal_pred_mask_scale = temporal_pred_mask_scale
        self.npred = npred
        self.max_context_duration = max(1, int(self.duration * max_context_frames_ratio))
        self.max_keep = max_keep
        self._itr_counter = Value("i", -1)   # shared across DataLoader workers, as in the reference
        self.inv_block = inv_block

    def step(self):
        i = self._itr_counter
        with i.get_lock():
            i.value += 1
            return i.value

    def _sample_block_size(self, generator, temporal_scale, spatial_scale, aspect_ratio_scale):
        r = torch.rand(1, generator=generator).item()
        t = max(1, int(self.duration * (temporal_scale[0] + r * (temporal_scale[1] - temporal_scale[0]))))
        r = torch.rand(1, generator=generator).item()
        keep = int(self.height * self.width * (spatial_scale[0] + r * (spatial_scale[1] - spatial_scale[0])))
        r = torch.rand(1, generator=generator).item()
        ar = aspect_ratio_scale[0] + r * (aspect_ratio_scale[1] - aspect_ratio_scale[0])
        h = min(int(round(math.sqrt(keep * ar))), self.height)
        w = min(int(round(math.sqrt(keep / ar))), self.width)
        return (t, h, w)

    def _sample_block_mask(self, b_size):
        t, h, w = b_size
        top = torch.randint(0, self.height - h + 1, (1,))
        left = torch.randint(0, self.width - w + 1, (1,))
        start = torch.randint(0, self.duration - t + 1, (1,))
        mask = torch.ones((self.duration, self.height, self.width), dtype=torch.int32)
        mask[start:start + t, top:top + h, left:left + w] = 0
        if self.max_context_duration < self.duration:
            mask[self.max_context_duration:, :, :] = 0
        return mask

    def __call__(self, batch_size):
        g = torch.Generator()
        g.manual_seed(self.step())
        p_size = self._sample_block_size(g, self.temporal_pred_mask_scale, self.spatial_pred_mask_scale,
                                         self.aspect_ratio)
        masks_p, masks_e = [], []
        total = self.duration * self.height * self.width
        min_keep_enc = min_keep_pred = total
        for _ in range(batch_size):
            while True:
                mask_e = torch.ones((self.duration, self.height, self.width), dtype=torch.int32)
                for _ in range(self.npred):
                    mask_e *= self._sample_block_mask(p_size)
                mask_e = mask_e.flatten()
                mask_p = torch.argwhere(mask_e == 0).squeeze()
                mask_e = torch.nonzero(mask_e).squeeze()
                if len(mask_e) != 0:
                    break
            min_keep_pred = min(min_keep_pred, len(mask_p))
            min_keep_enc = min(min_keep_enc, len(mask_e))
            masks_p.append(mask_p)
            masks_e.append(mask_e)
        if self.max_keep is not None:
            min_keep_enc = min(min_keep_enc, self.max_keep)
        masks_e = [cm[:min_keep_enc] for cm in masks_e]
        masks_p = [cm[:min_keep_pred] for cm in masks_p]
        if self.full_complement:
            masks_p = [torch.tensor(sorted(set(range(total)) - set(cm.tolist())), dtype=cm.dtype) for cm in masks_e]
        elif self.pred_full_complement:
            masks_e = [torch.tensor(sorted(set(range(total)) - set(cm.tolist())), dtype=cm.dtype) for cm in masks_p]
        masks_e = torch.utils.data.default_collate(masks_e)
        masks_p = torch.utils.data.default_collate(masks_p)
        if self.inv_block:
            return masks_p, masks_e
        return masks_e, masks_p
