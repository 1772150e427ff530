"""Mask application and generation: the counterparts of src/masks/utils.py:9-21 (apply_masks) and
src/masks/multiseq_multiblock3d.py:16-239 (MaskCollator and its per-config generator).

apply_masks runs on the device as a 128-bit row-copy gather (bit-exact).  The multiblock-3D sampler is host-side
integer work; its contract with the reference is the RANDOM STREAM, not the code: per draw one seeded generator
yields three uniforms (temporal scale, spatial scale, aspect ratio), then the GLOBAL torch CPU generator yields,
per sample and block, `randint` top, left, start in that order (multiseq_multiblock3d.py:172-196).  Identical RNG
state in therefore gives bit-identical index tensors out (golden-pinned in tests/test_cpu_host.py).  Everything
around those calls -- a boolean visibility grid instead of multiplied int32 masks, specs as dataclasses -- is
this package's own.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
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


@dataclass(frozen=True)
class TokenGrid:
    """Token lattice of a clip after tubelet / patch embedding: frames x rows x columns."""
    frames: int
    rows: int
    cols: int

    @property
    def size(self):
        return self.frames * self.rows * self.cols


@dataclass(frozen=True)
class BlockMaskSpec:
    """One entry of the YAML `mask:` list (configs/train/*/pretrain-*.yaml)."""
    spatial_scale: tuple = (0.2, 0.8)
    temporal_scale: tuple = (1.0, 1.0)
    aspect_ratio: tuple = (0.3, 3.0)
    num_blocks: int = 1
    max_temporal_keep: float = 1.0
    max_keep: int | None = None
    full_complement: bool = False
    pred_full_complement: bool = False
    inv_block: bool = False

    @classmethod
    def from_cfg(cls, cfg):
        known = {f: cfg[f] for f in cls.__dataclass_fields__ if cfg.get(f) is not None}
        return cls(**known)


def _lerp(lo_hi, u):
    return lo_hi[0] + u * (lo_hi[1] - lo_hi[0])


class BlockMaskSampler:
    """Draws (kept, masked) token-id matrices for one mask spec on one token grid."""

    def __init__(self, grid: TokenGrid, spec: BlockMaskSpec):
        self.grid, self.spec = grid, spec
        self.context_frames = max(1, int(grid.frames * spec.max_temporal_keep))
        self._draws = Value("i", -1)              # shared by DataLoader workers: draw k is seeded with k

    def step(self):
        with self._draws.get_lock():
            self._draws.value += 1
            return self._draws.value

    # -- the three seeded uniforms of a draw -> block extent (frames, rows, cols), shared by the whole batch
    def _block_extent(self, seed):
        gen = torch.Generator().manual_seed(seed)
        u_t, u_s, u_a = (torch.rand(1, generator=gen).item() for _ in range(3))
        g, sp = self.grid, self.spec
        frames = max(1, int(g.frames * _lerp(sp.temporal_scale, u_t)))
        area = int(g.rows * g.cols * _lerp(sp.spatial_scale, u_s))
        aspect = _lerp(sp.aspect_ratio, u_a)
        rows = min(int(round(math.sqrt(area * aspect))), g.rows)
        cols = min(int(round(math.sqrt(area / aspect))), g.cols)
        return frames, rows, cols

    # -- one sample: the union of num_blocks boxes is hidden; three global-RNG randints per box (top, left, start)
    def _visible(self, extent):
        g = self.grid
        bf, br, bc = extent
        vis = torch.ones(g.frames, g.rows, g.cols, dtype=torch.bool)
        for _ in range(self.spec.num_blocks):
            top = int(torch.randint(0, g.rows - br + 1, (1,)))
            left = int(torch.randint(0, g.cols - bc + 1, (1,)))
            start = int(torch.randint(0, g.frames - bf + 1, (1,)))
            vis[start:start + bf, top:top + br, left:left + bc] = False
        if self.context_frames < g.frames:
            vis[self.context_frames:] = False
        return vis.flatten()

    def __call__(self, batch_size):
        extent = self._block_extent(self.step())
        kept, hidden = [], []
        for _ in range(batch_size):
            vis = self._visible(extent)
            while not bool(vis.any()):              # nothing left for the encoder: draw this sample again
                vis = self._visible(extent)
            ids = torch.arange(self.grid.size)
            kept.append(ids[vis])
            hidden.append(ids[~vis])
        # a batch is rectangular: truncate every sample to the shortest one (and to max_keep)
        n_kept = min(len(k) for k in kept)
        n_hidden = min(len(h) for h in hidden)
        if self.spec.max_keep is not None:
            n_kept = min(n_kept, self.spec.max_keep)
        kept = [k[:n_kept] for k in kept]
        hidden = [h[:n_hidden] for h in hidden]
        everything = torch.ones(self.grid.size, dtype=torch.bool)
        if self.spec.full_complement:               # predict every token the (truncated) context does not hold
            hidden = [_complement(everything, k) for k in kept]
        elif self.spec.pred_full_complement:
            kept = [_complement(everything, h) for h in hidden]
        kept, hidden = torch.stack(kept), torch.stack(hidden)
        return (hidden, kept) if self.spec.inv_block else (kept, hidden)


def _complement(everything, ids):
    rest = everything.clone()
    rest[ids] = False
    return torch.nonzero(rest).squeeze(1)


class MaskCollator(object):
    """DataLoader collate_fn (multiseq_multiblock3d.py:16-76): groups samples by frames-per-clip, collates each group
    and attaches one (masks_enc, masks_pred) pair per mask spec.  Same constructor arguments and output structure."""

    def __init__(self, cfgs_mask, dataset_fpcs, crop_size=(224, 224), patch_size=(16, 16), tubelet_size=2):
        crop = crop_size if isinstance(crop_size, tuple) else (crop_size,) * 2
        patch = patch_size if isinstance(patch_size, tuple) else (patch_size,) * 2
        specs = [BlockMaskSpec.from_cfg(c) for c in cfgs_mask]
        self.samplers = {
            fpc: [BlockMaskSampler(TokenGrid(fpc // tubelet_size, crop[0] // patch[0], crop[1] // patch[1]), sp)
                  for sp in specs]
            for fpc in dataset_fpcs}

    def step(self):
        for group in self.samplers.values():
            for sampler in group:
                sampler.step()

    def draw(self, fpc, batch_size):
        """Masks only (what the step consumes): ([masks_enc per spec], [masks_pred per spec])."""
        pairs = [sampler(batch_size) for sampler in self.samplers[fpc]]
        return [e for e, _ in pairs], [p for _, p in pairs]

    def __call__(self, batch):
        by_fpc = {fpc: [] for fpc in self.samplers}
        for sample in batch:
            by_fpc[len(sample[-1][-1])].append(sample)      # frames per clip = length of the last clip-index list
        out = []
        for fpc, group in by_fpc.items():
            if group:
                enc, pred = self.draw(fpc, len(group))
                out.append((torch.utils.data.default_collate(group), enc, pred))
        return out


# ---------------------------------------------------------------------------------------------- device-side collator
def _mt_words_from_torch_state(state: torch.Tensor) -> torch.Tensor:
    """torch.get_rng_state() (CPUGeneratorImplState: seed u64, left i32, seeded i32, next u64, state u64[624], ...)
    -> the 626 uint32 words vj_mask_collate keeps on the device (state[624], left, next), as int32 storage."""
    raw = state.numpy()
    left = int(raw[8:12].view("<i4")[0])
    nxt = int(raw[16:24].view("<u8")[0])
    words = raw[24:24 + 624 * 8].view("<u8").astype("<u4")
    import numpy as np
    return torch.from_numpy(np.concatenate([words, np.array([left, nxt], dtype="<u4")]).view("<i4").copy())


def _torch_state_from_mt_words(words: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Inverse of _mt_words_from_torch_state: patch left / next / state[] into a copy of a torch CPU RNG state."""
    out = like.clone()
    raw = out.numpy()
    w = words.cpu().numpy().view("<u4")
    raw[8:12].view("<i4")[0] = int(w[624])
    raw[16:24].view("<u8")[0] = int(w[625])
    raw[24:24 + 624 * 8].view("<u8")[:] = w[:624].astype("<u8")
    return out


class DeviceMaskCollator:
    """MaskCollator whose sampling runs on the GPU (csrc/maskgen.cu), RNG-call-identical to the reference
    (multiseq_multiblock3d.py:129-239): same constructor arguments, and -- started from the same global torch CPU
    generator state and the same draw counters -- bit-identical (masks_enc, masks_pred) index tensors, produced in
    device memory.  The Mersenne-Twister state of the global generator is uploaded once (`seed_from_torch`) and then
    lives on the device; `torch_rng_state()` reads it back in torch's own format.

    The kept / hidden counts set the launch geometry of the step that consumes the masks, so they have to reach the
    host: `enqueue` launches the kernels of one draw on a side stream and starts a 16-byte async copy of the counts
    into pinned memory; `collect` (called one step later) waits on that event only and returns views of the dense
    device buffers.  No host synchronisation is added to the step."""

    def __init__(self, cfgs_mask, dataset_fpcs, crop_size=(224, 224), patch_size=(16, 16), tubelet_size=2,
                 device=None, depth=2):
        from . import _cabi as C
        crop = crop_size if isinstance(crop_size, tuple) else (crop_size,) * 2
        patch = patch_size if isinstance(patch_size, tuple) else (patch_size,) * 2
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("vjepa2_b200.DeviceMaskCollator: CUDA device only (the host-side sampler is MaskCollator)")
        self.specs = [BlockMaskSpec.from_cfg(c) for c in cfgs_mask]
        self.grids = {fpc: TokenGrid(fpc // tubelet_size, crop[0] // patch[0], crop[1] // patch[1]) for fpc in dataset_fpcs}
        self._cspecs = {}
        for fpc, g in self.grids.items():
            self._cspecs[fpc] = []
            for sp in self.specs:
                cs = C.MaskSpec(
                    frames=g.frames, rows=g.rows, cols=g.cols, num_blocks=sp.num_blocks,
                    context_frames=max(1, int(g.frames * sp.max_temporal_keep)),
                    max_keep=sp.max_keep if sp.max_keep is not None else 0,
                    full_complement=int(sp.full_complement), pred_full_complement=int(sp.pred_full_complement),
                    temporal_lo=sp.temporal_scale[0], temporal_hi=sp.temporal_scale[1],
                    spatial_lo=sp.spatial_scale[0], spatial_hi=sp.spatial_scale[1],
                    aspect_lo=sp.aspect_ratio[0], aspect_hi=sp.aspect_ratio[1])
                self._cspecs[fpc].append(cs)
        self._draws = {fpc: [-1] * len(self.specs) for fpc in dataset_fpcs}     # per-generator draw counters (:121-126)
        self._rng = torch.zeros(C.MASK_RNG_WORDS, dtype=torch.int32, device=self.device)
        self._seeded = False
        self._like = None
        self._stream = torch.cuda.Stream(self.device)
        self._depth = depth
        self._slots = []           # ring of output buffer sets
        self._turn = 0
        self._pending = []

    # -- RNG state plumbing
    def seed_from_torch(self, state=None):
        """Take over the global torch CPU generator stream (default: its current state)."""
        state = torch.get_rng_state() if state is None else state
        self._like = state.clone()
        words = _mt_words_from_torch_state(state)
        with torch.cuda.stream(self._stream):
            self._rng.copy_(words.to(self.device))
        self._stream.synchronize()
        self._seeded = True

    def torch_rng_state(self):
        """The device generator's state in torch.get_rng_state() format (for torch.set_rng_state / checkpoints)."""
        self._stream.synchronize()
        return _torch_state_from_mt_words(self._rng, self._like)

    def step(self):
        for fpc in self._draws:
            self._draws[fpc] = [d + 1 for d in self._draws[fpc]]

    # -- one draw = one kernel per mask spec, in the reference's order (generator 0 for the whole batch, then 1, ...)
    def _slot(self, fpc, B):
        """Output buffers rotate over depth + 1 sets: `depth` draws in flight plus the one the current step reads."""
        n = self.grids[fpc].size
        if len(self._slots) < self._depth + 1:
            self._slots.append(None)
            i = len(self._slots) - 1
        else:
            i = self._turn % (self._depth + 1)
        self._turn += 1
        s = self._slots[i]
        if s is None or s["cap"] < B * n:
            s = dict(cap=B * n,
                     enc=[torch.empty(B * n, dtype=torch.int64, device=self.device) for _ in self.specs],
                     pred=[torch.empty(B * n, dtype=torch.int64, device=self.device) for _ in self.specs],
                     counts=torch.zeros(2 * len(self.specs), dtype=torch.int32, device=self.device),
                     counts_host=torch.zeros(2 * len(self.specs), dtype=torch.int32).pin_memory(),
                     scratch=torch.empty(B * n, dtype=torch.uint8, device=self.device),
                     event=torch.cuda.Event())
            self._slots[i] = s
        return s

    def enqueue(self, fpc, batch_size):
        import ctypes
        from . import _cabi as C
        if not self._seeded:
            self.seed_from_torch()
        if len(self._pending) >= self._depth:
            raise RuntimeError("DeviceMaskCollator: too many outstanding draws; call collect() first")
        lib = C.load()
        s = self._slot(fpc, batch_size)
        # the set being overwritten was last read by work already queued on the caller's stream
        self._stream.wait_stream(torch.cuda.current_stream(self.device))
        st = self._stream.cuda_stream
        for j, cs in enumerate(self._cspecs[fpc]):
            self._draws[fpc][j] += 1
            seed = self._draws[fpc][j] & 0xFFFFFFFF
            C.check(lib.vj_mask_collate(self._rng.data_ptr(), ctypes.byref(cs), seed, batch_size, s["enc"][j].data_ptr(),
                                        s["pred"][j].data_ptr(), s["counts"][2 * j:].data_ptr(), s["scratch"].data_ptr(),
                                        st), "vj_mask_collate")
        with torch.cuda.stream(self._stream):
            s["counts_host"].copy_(s["counts"], non_blocking=True)
            s["event"].record(self._stream)
        self._pending.append((s, fpc, batch_size))

    def collect(self):
        """([masks_enc per spec], [masks_pred per spec]) of the oldest outstanding draw: int64 [B, K] device tensors.
        The returned views stay valid until `depth` further draws have been enqueued."""
        s, fpc, B = self._pending.pop(0)
        s["event"].synchronize()
        torch.cuda.current_stream(self.device).wait_event(s["event"])
        k = s["counts_host"].tolist()
        enc, pred = [], []
        for j, sp in enumerate(self.specs):
            e = s["enc"][j][:B * k[2 * j]].view(B, k[2 * j])
            p = s["pred"][j][:B * k[2 * j + 1]].view(B, k[2 * j + 1])
            enc.append(p if sp.inv_block else e)
            pred.append(e if sp.inv_block else p)
        return enc, pred

    def draw(self, fpc, batch_size):
        """Synchronous convenience (tests): enqueue + collect."""
        self.enqueue(fpc, batch_size)
        return self.collect()
