"""The V-JEPA 2 pre-training step (app/vjepa/train.py:409-471) as one fused, autograd-free pipeline.

    h   = LN(target_encoder(clips))                         (no grad, :414-418)
    z   = predictor(encoder(clips, masks_enc), masks_enc, masks_pred)   (:420-423)
    L   = mean_j mean |z_j - h[masks_pred_j]|               (:425-435)
    GradScaler-scaled backward, unscale, inf-skip AdamW step (:444-454)
    target <- m * target + (1 - m) * encoder                (:456-465)

Differences in *mechanism* (not in result): each (fpc group, mask) pair runs forward and backward back
to back so only one pair's activations are alive; parameter gradients accumulate in a flat fp32 buffer
written directly by the wgrad GEMM epilogues; the optimizer, the inf check, the GradScaler update and the
EMA are flat kernels over that buffer; nothing synchronises with the host (the reference's
`float(loss)` is left to the caller).  Data parallelism: per-block gradient ranges are all-reduced with
NCCL as soon as the last backward pass has produced them (DDP semantics, train.py:279-281).
"""
from __future__ import annotations

import copy
import ctypes

import torch
import torch.distributed as dist

from . import engine, ops
from .predictor import vit_predictor
from .schedulers import CosineWDSchedule, WarmupCosineSchedule, momentum_schedule
from .workspace import Arena
from .wrappers import MultiSeqWrapper, PredictorMultiSeqWrapper
from . import vision_transformer as video_vit


def init_video_model(device, patch_size=16, max_num_frames=16, tubelet_size=2, model_name="vit_base", crop_size=224,
                     pred_depth=6, pred_num_heads=None, pred_embed_dim=384, uniform_power=False,
                     use_mask_tokens=False, num_mask_tokens=2, zero_init_mask_tokens=True, use_sdpa=False,
                     use_rope=False, use_silu=False, use_pred_silu=False, wide_silu=False,
                     use_activation_checkpointing=False):
    """app/vjepa/utils.py:138-204 (same arguments; same name lookup seam)."""
    encoder = video_vit.__dict__[model_name](
        img_size=crop_size, patch_size=patch_size, num_frames=max_num_frames, tubelet_size=tubelet_size,
        uniform_power=uniform_power, use_sdpa=use_sdpa, use_silu=use_silu, wide_silu=wide_silu,
        use_activation_checkpointing=use_activation_checkpointing, use_rope=use_rope)
    encoder = MultiSeqWrapper(encoder)
    predictor = vit_predictor(
        img_size=crop_size, use_mask_tokens=use_mask_tokens, patch_size=patch_size, num_frames=max_num_frames,
        tubelet_size=tubelet_size, embed_dim=encoder.backbone.embed_dim, predictor_embed_dim=pred_embed_dim,
        depth=pred_depth, num_heads=encoder.backbone.num_heads if pred_num_heads is None else pred_num_heads,
        uniform_power=uniform_power, num_mask_tokens=num_mask_tokens, zero_init_mask_tokens=zero_init_mask_tokens,
        use_rope=use_rope, use_sdpa=use_sdpa, use_silu=use_pred_silu, wide_silu=wide_silu,
        use_activation_checkpointing=use_activation_checkpointing)
    predictor = PredictorMultiSeqWrapper(predictor)
    encoder.to(device)
    predictor.to(device)
    return encoder, predictor


def _world_of(group):
    """Ranks in `group` (None = the default group); group=False opts out of data parallelism inside a distributed job."""
    if group is False or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class GradBucketer:
    """Bucketed, asynchronous all-reduce over ranges of a flat gradient buffer (device agnostic, so the
    N>1 logic is testable on CPU with gloo).  Ranges are reduced in the order they are submitted."""

    def __init__(self, group=None):
        self.group = group
        self.world = _world_of(group)
        self._works = []

    def submit(self, flat, start, end):
        if self.world == 1 or end <= start:
            return
        self._works.append(dist.all_reduce(flat[start:end], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for w in self._works:
            w.wait()
        self._works = []


class _NoReducer:
    """VJ_DDP_COMM=none: measures what the replicas cost WITHOUT any gradient exchange (never a training mode)."""

    def submit(self, flat, start, end):
        pass

    def wait(self):
        pass


class PeerGradReducer:
    """Gradient all-reduce (train.py:279-281) over peer-mapped memory with the COPY ENGINES doing the transfers.

    Why not NCCL here: the backward pass is a train of persistent one-CTA-per-SM kernels that hold the whole register
    file, so an NCCL kernel and a GEMM cannot share an SM; every overlapped NCCL bucket stalls the (statically
    scheduled) compute kernel it meets.  Measured on 2 B200s (profiles/r02e_*): 303 ms at N=1, 314 ms with the
    overlapped NCCL all-reduce, 314 ms with one exposed all-reduce after backward (6.8 ms stand-alone).

    Here the flat fp32 gradient buffers live in symmetric memory (torch.distributed._symmetric_memory: every rank maps
    every other rank's buffer).  Per bucket [lo, hi), on a side stream, rank r owning slice r of the bucket:
        wait(compute produced the bucket) -> barrier -> PULL slice r of every peer into local staging (DMA, NVLink)
        -> g[slice r] += staged (vj_sum_into, the only SM work: ~1 ms per step in total) -> PUSH slice r into every
        peer's buffer (DMA) -> barrier.
    The barrier is a one-warp kernel (vj_peer_barrier: release/acquire flags in symmetric memory) that fits beside a
    GEMM CTA.  Slice r is summed by rank r only, in rank order, and broadcast: all ranks hold identical bits."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm
        from . import _cabi as C
        self._symm, self._C = symm, C
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > C.MAX_PEERS:
            raise RuntimeError(f"PeerGradReducer: world size {self.world} > {C.MAX_PEERS}")
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device, priority=-1)
        self._flags = symm.empty(C.MAX_PEERS, dtype=torch.int32, device=self.device)
        self._flags.zero_()
        torch.cuda.synchronize(self.device)
        hdl = symm.rendezvous(self._flags, self.group)
        self._flag_list = C.PtrList()
        for p in range(self.world):
            self._flag_list.ptr[p] = int(hdl.buffer_ptrs[p])
        self._flags_hdl = hdl
        hdl.barrier()                               # every rank's flags are zeroed and mapped before the first epoch
        torch.cuda.synchronize(self.device)
        self._epoch = 0
        self._bufs = {}                             # data_ptr -> (local tensor, [peer views])
        self._staging = None
        self._done = None

    def alloc(self, numel):
        """A zeroed flat fp32 gradient buffer in symmetric memory, mapped by every rank (collective call)."""
        t = self._symm.empty(int(numel), dtype=torch.float32, device=self.device)
        t.zero_()
        torch.cuda.synchronize(self.device)
        hdl = self._symm.rendezvous(t, self.group)
        views = [t if p == self.rank else hdl.get_buffer(p, (int(numel),), torch.float32) for p in range(self.world)]
        self._bufs[t.data_ptr()] = (t, views, hdl)
        return t

    def _barrier(self):
        self._epoch += 1
        self._C.check(self._C.load().vj_peer_barrier(ctypes.byref(self._flag_list), self.rank, self.world,
                                                     self._epoch & 0x7FFFFFFF, self.stream.cuda_stream), "vj_peer_barrier")

    def submit(self, flat, start, end):
        """All-reduce flat[start:end] (flat from alloc()); ordered after the work queued so far on the current stream."""
        if self.world == 1 or end <= start:
            return
        local, views, _ = self._bufs[flat.data_ptr()]
        W, r = self.world, self.rank
        n = end - start
        per = ((n + W - 1) // W + 1023) // 1024 * 1024         # slice length (multiple of 1024 elements)
        lo = min(end, start + r * per)
        hi = min(end, lo + per)
        m = hi - lo
        if self._staging is None or self._staging.shape[1] < per:
            self.stream.synchronize()
            self._staging = torch.empty(W - 1, per, dtype=torch.float32, device=self.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        st = self.stream
        st.wait_event(ready)
        with torch.cuda.stream(st):
            self._barrier()                                     # the bucket is final on every rank
            if m > 0:
                peers = [p for p in range(W) if p != r]
                for k, p in enumerate(peers):                    # pull (copy engine)
                    self._staging[k, :m].copy_(views[p][lo:hi], non_blocking=True)
                srcs = self._C.PtrList()
                for k in range(W - 1):
                    srcs.ptr[k] = self._staging[k].data_ptr()
                mm = (m + 3) // 4 * 4                            # slices are 1024-aligned inside 1024-padded buffers
                self._C.check(self._C.load().vj_sum_into(local[lo:].data_ptr(), ctypes.byref(srcs), W - 1, mm,
                                                         st.cuda_stream), "vj_sum_into")
                for p in peers:                                  # push (copy engine)
                    views[p][lo:hi].copy_(local[lo:hi], non_blocking=True)
            self._barrier()                                     # every slice of the bucket has landed everywhere
            self._done = torch.cuda.Event()
            self._done.record(st)

    def wait(self):
        if self._done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done)
            self._done = None


def make_bucket_hook(ranges, bucket_bytes, submit, last_key=-1):
    """on_block_done callback of the last backward pass.  `ranges[key]` = [lo, hi) of the flat fp32 gradient range that
    is complete once engine.encoder_backward reports `key` (final norm, blocks depth-1 .. 0, then patch embed = last_key):
    the encoder finishes from the END of the flat store towards its start, so consecutive finished ranges are adjacent and
    are merged until a bucket holds at least bucket_bytes, then handed to submit(lo, hi)."""
    state = {"lo": None, "hi": None}

    def hook(key):
        lo, hi = ranges[key]
        if state["hi"] is None:
            state["lo"], state["hi"] = lo, hi
        else:
            state["lo"], state["hi"] = min(state["lo"], lo), max(state["hi"], hi)
        if key == last_key or (state["hi"] - state["lo"]) * 4 >= bucket_bytes:
            submit(state["lo"], state["hi"])
            state["lo"] = state["hi"] = None
    return hook


class FrozenTokenSync:
    """Which predictor mask tokens get an optimizer update this step (device agnostic, so testable with gloo).

    The reference wraps the predictor in DDP(find_unused_parameters=True) (train.py:280): a mask token no rank used
    keeps grad None and torch's AdamW skips it (no update, no weight decay), while a token ANY rank used receives the
    all-reduced gradient and is updated on EVERY rank.  With several dataset_fpcs, ranks may see different fpc
    groups in one step, so the used set is reduced over the ranks (MAX, on the device -- no host sync) and written into
    the per-tile flag bytes the AdamW kernel reads."""

    def __init__(self, flags, token_tiles, group=None):
        """flags: uint8 [tiles] (FlatStore.flags); token_tiles: list over tokens of (first tile, n tiles)."""
        self.flags, self.group = flags, group
        self.world = _world_of(group)
        dev = flags.device
        tiles = [t for t0, n in token_tiles for t in range(t0, t0 + n)]
        owner = [k for k, (t0, n) in enumerate(token_tiles) for _ in range(n)]
        self.n_tokens = len(token_tiles)
        self.tiles = torch.tensor(tiles, dtype=torch.int64, device=dev)
        self.owner = torch.tensor(owner, dtype=torch.int64, device=dev)
        self.base = (flags[self.tiles] & ~2 & 0xFF).clone()          # the non-frozen bits (weight-decay bit) of those tiles
        self._key = None

    def update(self, used_local):
        """used_local: set of token indices this rank's step uses.  Returns the device tensor of globally used tokens
        (world > 1) or None when nothing had to change."""
        key = tuple(sorted(used_local))
        if self.world == 1 and key == self._key:
            return None
        self._key = key
        host = torch.zeros(self.n_tokens, dtype=torch.int32)
        host[list(key)] = 1
        if self.flags.is_cuda:
            host = host.pin_memory()
        used = host.to(self.flags.device, non_blocking=True)
        if self.world > 1:
            dist.all_reduce(used, op=dist.ReduceOp.MAX, group=self.group)
        frozen_bit = ((1 - used[self.owner]) * 2).to(torch.uint8)
        self.flags[self.tiles] = self.base | frozen_bit
        return used


class HostFeeder:
    """Pinned-host -> device staging of a step's inputs on a side stream, double buffered, so the H2D copies of
    step i+1 (train.py:393-402: clips + mask indices) overlap the compute of step i.

        feeder.prefetch(clips, masks_enc, masks_pred)      # host (pinned) tensors of the NEXT step
        clips_d, me_d, mp_d = feeder.get()                  # device tensors of the oldest prefetched step
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._q = []
        self._clip_bufs = {}      # (slot, group index) -> persistent device buffer
        self._slot = 0
        self._free = [None, None]  # event: compute finished reading this slot

    def prefetch(self, clips, masks_enc, masks_pred):
        if len(self._q) >= 2:
            raise RuntimeError("vjepa2_b200: HostFeeder is double buffered -- at most two prefetched steps may be "
                               "outstanding (call get() first)")
        slot = self._slot
        self._slot ^= 1
        st = self.stream
        if self._free[slot] is not None:
            st.wait_event(self._free[slot])
        with torch.cuda.stream(st):
            dc = []
            for gi, c in enumerate(clips):
                buf = self._clip_bufs.get((slot, gi))
                if buf is None or buf.shape != c.shape:
                    buf = torch.empty(c.shape, dtype=torch.float32, device=self.device)
                    self._clip_bufs[(slot, gi)] = buf
                buf.copy_(c, non_blocking=True)
                dc.append(buf)
            me = [[m.to(self.device, non_blocking=True) for m in g] for g in masks_enc]
            mp = [[m.to(self.device, non_blocking=True) for m in g] for g in masks_pred]
            ev = torch.cuda.Event()
            ev.record(st)
        self._q.append((slot, ev, dc, me, mp))

    def get(self):
        slot, ev, dc, me, mp = self._q.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for g in me + mp:
            for m in g:
                m.record_stream(cur)
        self._pending_slot = slot
        return dc, me, mp

    def release(self):
        """Call after the step that consumed the last get() has been enqueued."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[self._pending_slot] = ev


def _unwrap(m):
    return m.backbone if hasattr(m, "backbone") else m


class JepaTrainStep:
    """Owns the optimizer state and runs train_step().  Construct once; call step() per iteration."""

    def __init__(self, encoder, predictor, target_encoder=None, *, ipe=300, epochs=800, ipe_scale=1.25, warmup=40,
                 start_lr=1e-4, lr=5.25e-4, final_lr=5.25e-4, weight_decay=0.04, final_weight_decay=0.04,
                 ema=(0.99925, 0.99925), betas=(0.9, 0.999), eps=1e-8, loss_exp=1.0, mixed_precision=True,
                 loss_scaling=True, process_group=None, overlap_target=None, grad_sync=None, bucket_mb=None):
        if loss_exp != 1.0:
            raise NotImplementedError("vjepa2_b200: loss_exp must be 1.0 (L1), as in every shipped config")
        if not mixed_precision:
            # train.py:96-103: mixed_precision=False means dtype=float32 and no autocast.  The kernels compute with bf16
            # tensor-core operands only, so a float32 config must not silently train in bf16.
            raise NotImplementedError("vjepa2_b200: mixed_precision=False (fp32 training) is out of scope; every shipped "
                                      "pre-training config uses dtype bfloat16.  For bf16 compute without GradScaler "
                                      "loss scaling pass loss_scaling=False.")
        self.encoder = _unwrap(encoder)
        self.predictor = _unwrap(predictor)
        if target_encoder is None:
            target_encoder = copy.deepcopy(self.encoder)          # train.py:210
        self.target_encoder = _unwrap(target_encoder)
        for p in self.target_encoder.parameters():
            p.requires_grad = False                                # train.py:282-283
        self.betas, self.eps = betas, eps
        self.ipe = ipe
        self.last_lr_wd = (start_lr, weight_decay)
        T_max = int(ipe_scale * epochs * ipe)
        self.scheduler = WarmupCosineSchedule(int(warmup * ipe), start_lr, lr, T_max, final_lr)
        self.wd_scheduler = CosineWDSchedule(weight_decay, T_max, final_weight_decay)
        self.momentum = momentum_schedule(ema, ipe, epochs, ipe_scale)
        self.bucketer = GradBucketer(process_group)
        self.world = self.bucketer.world
        import os
        # gradient all-reduce schedule (train.py:279-281): "overlap" = per-bucket async all-reduce launched from the last
        # backward pass; "end" = one all-reduce per model after backward (no SM contention with the persistent kernels)
        self.grad_sync = grad_sync or os.environ.get("VJ_DDP_SYNC", "overlap")
        if self.grad_sync not in ("overlap", "end"):
            raise ValueError("grad_sync must be 'overlap' or 'end'")
        self._bucket_mb = bucket_mb if bucket_mb is not None else os.environ.get("VJ_DDP_BUCKET_MB")
        # optimizer step() calls so far; the steps GradScaler skipped (found_inf) are counted on the device
        # (self.skipped), so the bias corrections use torch's count, applied_steps - skipped, without a host sync
        self.applied_steps = 0
        # the step manages bf16 shadows itself (AdamW / EMA kernels rewrite them)
        for m in (self.encoder, self.predictor, self.target_encoder):
            m._manual_shadows = False
        self.enc_rt = self.encoder.runtime()
        self.pred_rt = self.predictor.runtime()
        self.tgt_rt = self.target_encoder.runtime()
        for m in (self.encoder, self.predictor, self.target_encoder):
            m._manual_shadows = True
        dev = self.enc_rt.fs.device
        # gradient all-reduce transport: "peer" = copy engines over peer-mapped symmetric memory (PeerGradReducer, the
        # default on CUDA with world > 1), "nccl" = torch.distributed all_reduce (GradBucketer; also the gloo path on CPU)
        self.grad_comm = os.environ.get("VJ_DDP_COMM", "peer") if (self.world > 1 and dev.type == "cuda") else "nccl"
        if self.grad_comm not in ("peer", "nccl", "none"):
            raise ValueError("VJ_DDP_COMM must be 'peer' or 'nccl' ('none' = no gradient exchange at all: a timing "
                             "diagnostic that lets the replicas diverge)")
        # bucket size: per-block ranges are merged up to this many bytes (NCCL: one bucket per block; the peer path
        # pays two cross-GPU barriers per bucket, so it takes larger ones)
        self.bucket_bytes = int(self._bucket_mb if self._bucket_mb is not None else (64 if self.grad_comm == "peer" else 0)) << 20
        self.peer = None
        if self.grad_comm == "peer":
            if self.enc_rt.fs.g32 is not None or self.pred_rt.fs.g32 is not None:
                raise RuntimeError("vjepa2_b200: gradient buffers already exist; build JepaTrainStep before the first "
                                   "backward so they can live in symmetric memory (or set VJ_DDP_COMM=nccl)")
            # symmetric memory needs peer access between all ranks of the group (one NVLink / NVSwitch box); every rank
            # must take the same path, so the outcome of the set-up is agreed on (MIN over ranks) before it is used
            peer, bufs, why = None, [], ""
            try:
                peer = PeerGradReducer(process_group, dev)
                bufs = [peer.alloc(fs.total) for fs in (self.enc_rt.fs, self.pred_rt.fs)]
            except Exception as ex:                                # noqa: BLE001 -- any failure means "use NCCL"
                peer, why = None, f"{type(ex).__name__}: {str(ex)[:200]}"
            ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
            if int(ok.item()) == 1:
                self.peer = peer
                self.enc_rt.fs.g32, self.pred_rt.fs.g32 = bufs
            else:
                import warnings
                warnings.warn("vjepa2_b200: peer-mapped gradient all-reduce unavailable on this group (" + (why or "another "
                              "rank failed") + "); using the NCCL all-reduce", RuntimeWarning)
                self.grad_comm = "nccl"
                self.bucket_bytes = int(self._bucket_mb if self._bucket_mb is not None else 0) << 20
        self.enc_rt.fs.ensure_grads()
        self.pred_rt.fs.ensure_grads()
        self.enc_rt.fs.ensure_adam()
        self.pred_rt.fs.ensure_adam()
        if self.enc_rt.fs.total != self.tgt_rt.fs.total:
            raise RuntimeError("target encoder layout differs from the encoder's")
        f32 = torch.float32
        self.loss_accum = torch.zeros(1, dtype=f32, device=dev)
        # loss_scaling=False (not a reference option; tests): bf16 compute, scale 1, and -- like the reference's plain
        # optimizer.step() branch (train.py:452) -- the optimizer never skips a step
        init_scale = 65536.0 if loss_scaling else 1.0               # torch.cuda.amp.GradScaler() default
        self.mixed_precision = mixed_precision
        self.loss_scaling = bool(loss_scaling)
        self.scale = torch.full((1,), init_scale, dtype=f32, device=dev)
        self.inv_scale = torch.full((1,), 1.0 / (init_scale * self.world), dtype=f32, device=dev)
        self.found_inf = torch.zeros(1, dtype=f32, device=dev)
        self.growth_tracker = torch.zeros(1, dtype=torch.int32, device=dev)
        self.skipped = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bias_c = torch.ones(2, dtype=f32, device=dev)
        pfs = self.pred_rt.fs
        self.frozen_sync = FrozenTokenSync(
            pfs.flags, [(pfs.offsets[pfs.index[id(t)]] // 1024, (t.numel() + 1023) // 1024) for t in self.predictor.mask_tokens],
            process_group)
        if self.world > 1:
            self.broadcast_state()                                  # DDP broadcasts rank 0's parameters at construction
        self.ws = Arena(dev)                                        # activations / temporaries (no allocator in-step)
        # Optional: the target-encoder forward (train.py:414-418) does not depend on the context pass, so it can run
        # on a second stream with its own arena, the prologue / tail of every persistent kernel of one stream filled
        # by the other stream's next kernel (joined by an event before the loss reads h; forked after the previous
        # step's EMA has rewritten the target weights).  Bit-identical results.  Measured on B200 (profiles/
        # r01c_overlap_target_ab.txt): 78.13 vs 78.05 clips/s -- the step is power-capped, the filled gaps come back
        # as a lower SM clock (1.46 vs 1.62 GHz) -- so it is OFF by default (VJ_OVERLAP_TARGET=1 / overlap_target=True).
        if overlap_target is None:
            import os
            overlap_target = os.environ.get("VJ_OVERLAP_TARGET", "0") == "1"
        self.overlap_target = bool(overlap_target)
        self._side = torch.cuda.Stream(dev) if self.overlap_target else None
        self.ws_t = None
        if self.overlap_target:
            with torch.cuda.stream(self._side):
                self.ws_t = Arena(dev, act_bytes=1 << 20, tmp_bytes=1 << 28)
        # flat ranges of the per-block buckets (reverse order of completion in backward)
        fs = self.enc_rt.fs
        self._enc_ranges = {i: fs.range_of(h.params) for i, h in enumerate(self.enc_rt.blocks)}
        self._enc_ranges[len(self.enc_rt.blocks)] = fs.range_of(self.enc_rt.norm_params)
        self._enc_ranges[-1] = fs.range_of(self.enc_rt.pe_params)

    # ------------------------------------------------------------------------------------------
    def _reducer(self):
        if self.grad_comm == "none":
            return _NoReducer()
        return self.peer if self.peer is not None else self.bucketer

    def _bucket_hook(self, efs):
        red = self._reducer()
        return make_bucket_hook(self._enc_ranges, self.bucket_bytes, lambda lo, hi: red.submit(efs.g32, lo, hi))

    def _set_frozen_mask_tokens(self, n_groups):
        """Group i uses mask token i % num_mask_tokens (wrappers.py:40, predictor.py:195); tokens no rank uses in this
        step are skipped by the optimizer (see FrozenTokenSync)."""
        n = len(self.predictor.mask_tokens)
        self.frozen_sync.update({i % n for i in range(n_groups)})

    def broadcast_state(self, src=0):
        """DDP construction semantics (train.py:279-281): every rank starts from rank `src`'s parameters (encoder,
        predictor, target encoder) -- and, after a checkpoint load, its optimizer state."""
        if self.world == 1:
            return
        g = self.bucketer.group
        for fs in (self.enc_rt.fs, self.pred_rt.fs, self.tgt_rt.fs):
            dist.broadcast(fs.p32, src, group=g)
            if fs.exp_avg is not None:
                dist.broadcast(fs.exp_avg, src, group=g)
                dist.broadcast(fs.exp_avg_sq, src, group=g)
            fs.refresh_shadows()
        for t in (self.scale, self.inv_scale, self.growth_tracker, self.skipped):
            dist.broadcast(t, src, group=g)

    # ------------------------------------------------------------------------------------------ resume support
    def reload_weights(self):
        """After load_state_dict / any in-place edit of the fp32 parameters: rebuild the bf16 operand shadows."""
        for rt in (self.enc_rt, self.pred_rt, self.tgt_rt):
            if not rt.fs.valid():
                raise RuntimeError("vjepa2_b200: parameters were re-allocated (e.g. .to()); build a new JepaTrainStep")
            rt.fs.refresh_shadows()
        self.broadcast_state()

    def optimizer_steps(self):
        """torch's optimizer step count: step() calls minus the ones GradScaler skipped (one 4-byte D2H read)."""
        return self.applied_steps - int(self.skipped.item())

    def set_optimizer_steps(self, n):
        self.applied_steps = int(n)
        self.skipped.zero_()

    def set_scaler(self, scale, growth_tracker=0):
        """GradScaler.load_state_dict (app/vjepa/utils.py:121-122)."""
        self.scale.fill_(float(scale))
        self.inv_scale.fill_(1.0 / (float(scale) * self.world))
        self.growth_tracker.fill_(int(growth_tracker))

    def fast_forward(self, n_steps):
        """train.py:309-313: advance the LR / WD / momentum schedules by n_steps iterations without training."""
        for _ in range(int(n_steps)):
            lr = self.scheduler.step()
            wd = self.wd_scheduler.step()
            next(self.momentum)
            self.last_lr_wd = (lr, wd)

    ACT_BUDGET_BYTES = 70 << 30          # saved activations of one forward/backward pass (of 180 GB HBM)

    def _mask_passes(self, mes, mps):
        """Greedy split of a group's masks into passes whose saved activations fit ACT_BUDGET_BYTES: per token and
        block the engine keeps ~(16 D + 4 Hm) bytes (LN outputs, qkv, attention out, fc1 pre/post, residual stream)."""
        e, p = self.encoder, self.predictor
        per_tok_enc = len(e.blocks) * (16 * e.embed_dim + 4 * e.blocks[0].mlp.fc1.out_features)
        pd = p.predictor_embed.out_features
        per_tok_pred = len(p.predictor_blocks) * (20 * pd + 4 * p.predictor_blocks[0].mlp.fc1.out_features)
        passes, lo, acc = [], 0, 0
        for j, (me, mp) in enumerate(zip(mes, mps)):
            need = int(1.1 * (me.numel() * per_tok_enc + (me.numel() + mp.numel()) * per_tok_pred))
            if j > lo and acc + need > self.ACT_BUDGET_BYTES:
                passes.append((lo, j))
                lo, acc = j, 0
            acc += need
        passes.append((lo, len(mes)))
        return passes

    def step(self, clips, masks_enc, masks_pred):
        """clips: list (one fp32 [B,3,T,H,W] tensor per fpc group); masks_enc / masks_pred: list over
        groups of lists over masks of int64 [B, K].  Returns (loss [1] fp32 device tensor, lr, wd)."""
        self.applied_steps += 1
        new_lr = self.scheduler.step()
        new_wd = self.wd_scheduler.step()
        self.last_lr_wd = (new_lr, new_wd)
        st = ops.stream()
        enc_rt, pred_rt, tgt_rt = self.enc_rt, self.pred_rt, self.tgt_rt
        efs, pfs, tfs = enc_rt.fs, pred_rt.fs, tgt_rt.fs
        enc = self.encoder
        for fs in (efs, pfs, tfs):
            if fs._versions is None:                              # load_state_dict since the last step: masters changed
                fs.refresh_shadows()
        self._set_frozen_mask_tokens(len(clips))
        ops.fill_f32(efs.g32, 0.0, st)                            # optimizer.zero_grad() (train.py:454)
        ops.fill_f32(pfs.g32, 0.0, st)
        ops.fill_f32(self.loss_accum, 0.0, st)
        ws = self.ws
        ws.reset()
        n_pairs = sum(len(m) for m in masks_enc)
        p = enc.patch_size

        pair_no = 0
        for i, c in enumerate(clips):
            c = c.contiguous()
            _, _, T, H, W = c.shape
            grid = (H // p, W // p) if enc.handle_nonsquare_inputs else (enc.grid_size, enc.grid_size)
            # ---- target (train.py:414-418): no-grad encoder + non-affine LayerNorm, eps 1e-5
            ev_h = None
            if self.overlap_target:
                main = torch.cuda.current_stream()
                ev_fork = torch.cuda.Event()
                ev_fork.record(main)                      # previous EMA / loss reads of the old h are ahead of this
                self._side.wait_event(ev_fork)
                with torch.cuda.stream(self._side):
                    if i == 0:
                        self.ws_t.reset()
                    h, _ = engine.encoder_forward(tgt_rt, c, None, grid, save=False, ws=self.ws_t)
                    Bq, N, D = h.shape
                    h2 = h.view(Bq * N, D)
                    ops.layernorm_fwd(h2, None, None, h2, None, None, 1e-5, ops.stream())
                    ev_h = torch.cuda.Event()
                    ev_h.record(self._side)
            else:
                h, _ = engine.encoder_forward(tgt_rt, c, None, grid, save=False, ws=ws)   # stays on the tmp stack
                Bq, N, D = h.shape
                h2 = h.view(Bq * N, D)
                ops.layernorm_fwd(h2, None, None, h2, None, None, 1e-5, st)     # in place (row-local)
            # ---- all masks of the group in ONE pass when their saved activations fit the budget: the masked copies are
            #      row blocks of the same token matrix, so every LayerNorm / GEMM / reduction launch covers them all
            #      (attention runs per mask); the reference loops over the masks (wrappers.py:15-43), which is the same
            #      arithmetic row by row.  The 64f x 384px geometry (62 GB for mask 0 alone) goes mask by mask.
            all_me = [m.contiguous() for m in masks_enc[i]]
            all_mp = [m.contiguous() for m in masks_pred[i]]
            for lo, hi in self._mask_passes(all_me, all_mp):
                mes, mps = all_me[lo:hi], all_mp[lo:hi]
                pair_no += len(mes)
                last = pair_no == n_pairs
                # ---- context + predictor forward (train.py:420-423)
                zs, sv_e = engine.encoder_forward(enc_rt, c, mes, grid, save=True, ws=ws)
                preds, sv_p = engine.predictor_forward(pred_rt, zs, mes, mps, i, save=True, ws=ws)
                # ---- loss (train.py:425-435) and its gradient, GradScaler-scaled (train.py:445)
                if ev_h is not None:
                    torch.cuda.current_stream().wait_event(ev_h)      # join: h is complete
                    ev_h = None
                Din = preds[0].shape[2]
                rows = sum(pr.shape[0] * pr.shape[1] for pr in preds)
                dz = ws.act((rows, Din), preds[0].dtype)
                r0 = 0
                for pr, mp in zip(preds, mps):
                    n = pr.shape[0] * pr.shape[1]
                    inv = 1.0 / (n_pairs * pr.numel())
                    mk = ws.mark()
                    ops.l1_loss(pr, h, mp, self.loss_accum, dz[r0:r0 + n], inv, inv, self.scale, st, ws.tmp)
                    ws.release(mk)
                    r0 += n
                # ---- backward
                dzenc = engine.predictor_backward(pred_rt, sv_p, dz, pfs.g32, ws=ws)
                del sv_p
                hook = None
                if last and self.world > 1 and self.grad_sync == "overlap":
                    self._reducer().submit(pfs.g32, 0, pfs.total)
                    hook = self._bucket_hook(efs)
                engine.encoder_backward(enc_rt, sv_e, dzenc, efs.g32, ws=ws, on_block_done=hook)
                del sv_e, zs, preds, dz, dzenc
                ws._act.reset()                                   # this pass's activations are dead
        if self.world > 1 and self.grad_sync == "end":                # one collective per model after backward
            self._reducer().submit(pfs.g32, 0, pfs.total)
            self._reducer().submit(efs.g32, 0, efs.total)
        self._reducer().wait()

        # ---- unscale + inf check + AdamW (train.py:446-451; app/vjepa/utils.py:239), flat kernels
        ops.grad_check(efs.g32, self.found_inf, st)
        ops.grad_check(pfs.g32, self.found_inf, st)
        b1, b2 = self.betas
        skip_flag = self.found_inf if self.loss_scaling else None       # no GradScaler: never skip (train.py:452)
        ops.adam_prepare(self.bias_c, self.skipped, skip_flag, self.applied_steps, b1, b2, st)
        for fs in (efs, pfs):
            ops.adamw_step(fs.p32, fs.g32, fs.exp_avg, fs.exp_avg_sq, fs.p16, fs.flags, new_lr, b1, b2, self.eps,
                           new_wd, self.applied_steps, self.inv_scale, skip_flag, st, dev_bias=self.bias_c)
        if self.loss_scaling:
            ops.scaler_update(self.scale, self.inv_scale, self.growth_tracker, self.found_inf, float(self.world), st=st)
        else:
            self.found_inf.zero_()
        # ---- EMA of the target encoder (train.py:456-465) + its bf16 shadow
        m = next(self.momentum)
        ops.ema_update(tfs.p32, efs.p32, tfs.p16, m, st)
        return self.loss_accum, new_lr, new_wd
