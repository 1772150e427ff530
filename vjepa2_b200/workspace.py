"""Activation / scratch memory for the engine.

Sequence lengths change every step (multiblock masks are truncated to the batch minimum,
multiseq_multiblock3d.py:211-215), which makes a caching allocator keep hitting cudaMalloc and fragmenting.
The fused step therefore carves every activation and temporary out of two bump arenas:

  act  -- what a forward pass saves for its backward; reset once the (group, mask) pair is done
  tmp  -- temporaries with stack (mark / release) lifetime

Both grow on demand (a fresh, larger chunk; old views stay valid until the next reset) and settle at the
high-water mark of the largest step, after which no allocator call happens inside a step.
`TorchAlloc` is the same interface on top of torch.empty for the autograd drop-in path.
"""
from __future__ import annotations

import torch

_ALIGN = 256
_GROW_CAP = 8 << 30


class TorchAlloc:
    def __init__(self, device):
        self.device = device

    def act(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    tmp = act

    def mark(self):
        return None

    def release(self, m):
        pass


class _Bump:
    def __init__(self, device, nbytes):
        self.device = device
        self.chunks = [torch.empty(max(nbytes, _ALIGN), dtype=torch.uint8, device=device)]
        self.ci, self.off, self.used_hw, self.used = 0, 0, 0, 0

    def alloc(self, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        nbytes = (n * dtype.itemsize + _ALIGN - 1) // _ALIGN * _ALIGN
        c = self.chunks[self.ci]
        if self.off + nbytes > c.numel():
            # look for room in a later chunk, else grow
            self.ci += 1
            while self.ci < len(self.chunks) and self.chunks[self.ci].numel() < nbytes:
                self.ci += 1
            if self.ci == len(self.chunks):
                # geometric growth, capped: a 60 GB arena (64f x 384px activations) must not double to 120 GB
                total = sum(ch.numel() for ch in self.chunks)
                self.chunks.append(torch.empty(max(nbytes, min(total, _GROW_CAP)), dtype=torch.uint8,
                                               device=self.device))
            self.off = 0
            c = self.chunks[self.ci]
        out = c[self.off:self.off + n * dtype.itemsize].view(dtype).view(shape)
        self.off += nbytes
        self.used += nbytes
        if self.used > self.used_hw:
            self.used_hw = self.used
        return out

    def mark(self):
        return (self.ci, self.off, self.used)

    def release(self, m):
        self.ci, self.off, self.used = m

    def reset(self):
        if len(self.chunks) > 1:        # consolidate to one chunk at the high-water size
            total = sum(ch.numel() for ch in self.chunks)
            self.chunks = None
            self.chunks = [torch.empty(total, dtype=torch.uint8, device=self.device)]
        self.ci, self.off, self.used = 0, 0, 0


class Arena:
    def __init__(self, device, act_bytes=1 << 30, tmp_bytes=1 << 28):
        self.device = device
        self._act = _Bump(device, act_bytes)
        self._tmp = _Bump(device, tmp_bytes)

    def act(self, shape, dtype):
        return self._act.alloc(shape, dtype)

    def tmp(self, shape, dtype):
        return self._tmp.alloc(shape, dtype)

    def mark(self):
        return self._tmp.mark()

    def release(self, m):
        self._tmp.release(m)

    def reset(self):
        self._act.reset()
        self._tmp.reset()

    def high_water(self):
        return self._act.used_hw, self._tmp.used_hw
