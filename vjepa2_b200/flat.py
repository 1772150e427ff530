"""Flat parameter storage.

All parameters of one model live in ONE contiguous fp32 buffer (each parameter padded to a 1024-element
tile), mirrored by one bf16 shadow buffer (the tensor-core operands), one fp32 gradient buffer and the
two Adam moment buffers.  nn.Parameters are views into the fp32 buffer, so `state_dict()`,
`load_state_dict()` and `zip(encoder.parameters(), target_encoder.parameters())` (train.py:461) keep
working, while EMA / AdamW / grad-check / all-reduce become single flat kernels or collectives instead
of 484-tensor foreach loops (train.py:464-465, app/vjepa/utils.py:239).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

TILE = 1024
FLAG_WD = 1       # weight decay applies (app/vjepa/utils.py:224-237: not bias, not 1-D)
FLAG_FROZEN = 2   # parameter never receives a gradient (torch skips grad=None)


class FlatStore:
    def __init__(self, module: nn.Module, device):
        self.device = torch.device(device)
        named = list(module.named_parameters())
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.offsets, self.numels = [], []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            self.numels.append(p.numel())
            off += (p.numel() + TILE - 1) // TILE * TILE
        self.total = off
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self.p32 = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        flags = torch.zeros(self.total // TILE, dtype=torch.uint8)
        for i, (n, p) in enumerate(named):
            o, k = self.offsets[i], self.numels[i]
            self.p32[o:o + k].copy_(p.detach().reshape(-1))
            p.data = self.p32[o:o + k].view(p.shape)
            if ("bias" not in n) and (p.dim() != 1):
                flags[o // TILE:(o + k + TILE - 1) // TILE] = FLAG_WD
        self.flags_host = flags
        self.flags = flags.to(self.device)
        self.p16 = torch.empty(self.total, dtype=torch.bfloat16, device=self.device)
        self.g32 = None
        self.exp_avg = None
        self.exp_avg_sq = None
        self._versions = None
        self.refresh_shadows()

    def __deepcopy__(self, memo):
        # copy.deepcopy(encoder) (train.py:210) must not alias or half-copy the store: the copy rebuilds its own
        return None

    # ---------------------------------------------------------------- views
    def _view(self, buf, p):
        i = self.index[id(p)]
        o, k = self.offsets[i], self.numels[i]
        return buf[o:o + k].view(p.shape)

    def w16(self, p):
        """bf16 shadow of parameter p (same shape)."""
        return self._view(self.p16, p)

    def grad_view(self, gbuf, p):
        return self._view(gbuf, p)

    def range_of(self, params):
        """[start, end) of the flat range spanned by `params` (must be contiguous in the store)."""
        idx = sorted(self.index[id(p)] for p in params)
        last = idx[-1]
        end = self.offsets[last] + (self.numels[last] + TILE - 1) // TILE * TILE
        return self.offsets[idx[0]], end

    # ---------------------------------------------------------------- state
    def valid(self) -> bool:
        p0, pl = self.params[0], self.params[-1]
        return (p0.data_ptr() == self.p32.data_ptr() + 4 * self.offsets[0]
                and pl.data_ptr() == self.p32.data_ptr() + 4 * self.offsets[-1])

    def refresh_shadows(self):
        ops.cast_f32_bf16(self.p32, self.p16)
        self._versions = [p._version for p in self.params]

    def shadows_stale(self) -> bool:
        v = self._versions
        for i, p in enumerate(self.params):
            if p._version != v[i]:
                return True
        return False

    def mark_fresh(self):
        self._versions = [p._version for p in self.params]

    def ensure_grads(self):
        if self.g32 is None:
            self.g32 = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        return self.g32

    def ensure_adam(self):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros(self.total, dtype=torch.float32, device=self.device)
            self.exp_avg_sq = torch.zeros(self.total, dtype=torch.float32, device=self.device)

    def is_frozen(self, p) -> bool:
        """Current state of the device flag byte (train.FrozenTokenSync updates it on the device); one small D2H read."""
        return bool(int(self.flags[self.offsets[self.index[id(p)]] // TILE].item()) & FLAG_FROZEN)

    def set_frozen(self, params, frozen=True):
        """Mark parameters that never get a gradient (e.g. unused predictor mask tokens, a1/a17)."""
        for p in params:
            i = self.index[id(p)]
            o, k = self.offsets[i], self.numels[i]
            sl = slice(o // TILE, (o + k + TILE - 1) // TILE)
            if frozen:
                self.flags_host[sl] |= FLAG_FROZEN
            else:
                self.flags_host[sl] &= ~FLAG_FROZEN & 0xFF
        self.flags.copy_(self.flags_host)          # in place: kernels / FrozenTokenSync keep pointing at this tensor
