#!/usr/bin/env python
"""Benchmark of the V-JEPA 2 pre-training step (BASELINE.json: clips/s & tokens/s per GPU, ViT-g/16,
16x256x256 clips, batch 24 per GPU, bf16 tensor-core operands; % of bf16 tensor-core peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model vit_giant_xformers]

One "step" = one full train step (target fwd, 2x masked context fwd+bwd, 2x predictor fwd+bwd, L1 loss,
GradScaler check, AdamW, EMA) on one synthetic batch with a fresh multiblock-3D mask draw.  Prints ONE JSON
line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the public API with pinned
host buffers (H2D of clips + masks and D2H of the loss inside the timed region).  `--impl reference` times
the CPU restatement of the reference step (oracle/, kind "port") on the host cores.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {  # name: (embed_dim, depth, heads, mlp_hidden)
    "vit_large": (1024, 24, 16, 4096),
    "vit_huge": (1280, 32, 16, 5120),
    "vit_giant_xformers": (1408, 40, 22, 6144),
}
PRED = dict(dim=384, depth=12, heads=12, hidden=1536)
# configs/train/vitg16/pretrain-256px-16f.yaml
OPT = dict(ipe=300, epochs=800, ipe_scale=1.25, warmup=40, start_lr=1e-4, lr=5.25e-4, final_lr=5.25e-4,
           weight_decay=0.04, final_weight_decay=0.04, ema=(0.99925, 0.99925))
MASK_CFG = [
    dict(aspect_ratio=(0.75, 1.5), num_blocks=8, spatial_scale=(0.15, 0.15), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None, full_complement=False),
    dict(aspect_ratio=(0.75, 1.5), num_blocks=2, spatial_scale=(0.7, 0.7), temporal_scale=(1.0, 1.0),
         max_temporal_keep=1.0, max_keep=None, full_complement=False),
]
FRAMES, CROP, PATCH, TUB = 16, 256, 16, 2
NTOK = (FRAMES // TUB) * (CROP // PATCH) ** 2


def block_flops(S, D, Hm):
    """forward FLOPs of one transformer block for one sample of S tokens."""
    return 8 * S * D * D + 4 * S * D * Hm + 4 * S * S * D


def step_flops(model, B, k_enc, k_pred, ntok=None):
    """Algorithmic FLOPs of one train step for B clips (backward = 2x forward, no recompute counted)."""
    D, depth, _, Hm = MODELS[model]
    pe = 2 * 1536 * D
    NTOK = ntok if ntok is not None else globals()["NTOK"]
    f = depth * block_flops(NTOK, D, Hm) + NTOK * pe                       # target forward
    for ke, kp in zip(k_enc, k_pred):
        f += 3 * depth * block_flops(ke, D, Hm) + 2 * ke * pe                # context fwd+bwd (+patch embed fwd, wgrad)
        S = ke + kp
        f += 3 * (PRED["depth"] * block_flops(S, PRED["dim"], PRED["hidden"]) + 2 * ke * D * PRED["dim"]
                  + 2 * kp * PRED["dim"] * D)
    return B * f


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        busy = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def make_masks(collator, B, steps, frames=None):
    out = []
    for _ in range(steps):
        enc, pred = collator.draw(frames or FRAMES, B)
        out.append((enc, pred))
    return out


# ---------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """Use every host core: torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


def ref_harness():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_harness as H
    return H


def cpu_reference_step_time(model, max_seconds=240.0, steps=1, warmup=0, batch=1, log=None):
    """Times the reference's CPU path of the step (fp32, torch CPU, all host threads) on 1-clip samples of the
    workload.  kind "reference": the reference's own modules (baseline/_ref, see baseline/ref_harness.py) around the
    restated step closure; kind "port": oracle/vjepa_oracle.py when the reference tree is absent."""
    import torch
    D, depth, heads, Hm = MODELS[model]
    H = ref_harness()
    kind = "reference" if H.find_ref_root() is not None else "port"
    torch.manual_seed(0)
    if kind == "reference":
        R = H.import_reference()
        enc, pred = H.build_models(R, model, crop=CROP, frames=FRAMES, pred_depth=PRED["depth"], pred_heads=PRED["heads"],
                                   pred_dim=PRED["dim"], num_mask_tokens=6)
        rs = H.RefStep(R, enc, pred, dict(OPT, loss_exp=1.0), mixed_precision=False)
        coll = R.MaskCollator(cfgs_mask=MASK_CFG, dataset_fpcs=[FRAMES], crop_size=(CROP, CROP), patch_size=(PATCH, PATCH),
                              tubelet_size=TUB)
        gens = coll.mask_generators[FRAMES]
        run = lambda c, me, mp: rs.step([c], [me], [mp])["loss"]  # noqa: E731
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import vjepa_oracle as O
        w_enc = O.init_encoder_weights(D, depth, Hm / D, seed=0)
        w_pred = O.init_predictor_weights(D, PRED["dim"], PRED["depth"], 6, seed=1)
        st = O.StepState(w_enc, w_pred, dict(depth=depth, heads=heads),
                         dict(depth=PRED["depth"], heads=PRED["heads"], grid_size=CROP // PATCH, num_patches=NTOK,
                              num_mask_tokens=6), dict(OPT, loss_exp=1.0))
        del w_enc, w_pred
        gens = O.make_mask_generators(O.DEFAULT_MASK_CFG, (CROP, CROP), FRAMES)
        run = lambda c, me, mp: O.train_step(st, c, me, mp)  # noqa: E731
    torch.manual_seed(239)
    g = torch.Generator().manual_seed(0)
    times, flops = [], []
    t_begin = time.time()
    for it in range(warmup + steps):
        clips = torch.randn(batch, 3, FRAMES, CROP, CROP, generator=g)
        masks = [gen(batch) for gen in gens]
        me, mp = [m[0] for m in masks], [m[1] for m in masks]
        t0 = time.time()
        loss = run(clips, me, mp)
        dt = time.time() - t0
        if log:
            log(f"cpu {kind} step {it}: {dt:.1f}s loss {loss:.4f}")
        if it >= warmup:
            times.append(dt)
            flops.append(step_flops(model, batch, [m.shape[1] for m in me], [m.shape[1] for m in mp]))
        if time.time() - t_begin > max_seconds and times:
            break
    return times, flops, batch, kind


def cpu_sample_text(kind, n_steps, batch):
    what = ("the reference's own modules (baseline/_ref) + restated train.py:409-471" if kind == "reference"
            else "oracle/vjepa_oracle.py")
    return (f"{n_steps} step(s) of 1 clip (same model, masks and optimizer; the GPU arm runs batch {batch}), "
            f"{what} on torch CPU fp32")


# ---------------------------------------------------------------------------------------------- reference on the GPU
def torch_cuda_baseline(model, B, masks_host, clips_host, dev, warmup=2, steps=3, log=None):
    """The like-for-like bar (SURVEY 8d "also report"): the UNMODIFIED reference's PyTorch-eager CUDA path --
    train.py:409-471 around the reference's modules, bf16 autocast + GradScaler, same batch, same mask draws, same
    box -- with and without use_activation_checkpointing (vision_transformer.py:198-201; the shipped configs set
    it to true).  None of this repo's kernels are on that path."""
    import torch
    H = ref_harness()
    if H.find_ref_root() is None:
        return dict(unavailable="reference tree not staged (baseline/_ref)")
    R = H.import_reference()
    out = dict(what="reference modules (PyTorch eager: cuDNN conv3d, cuBLASLt, SDPA, ATen), bf16 autocast + GradScaler, "
                    f"batch {B}, same masks; CUDA events", unit="clips/s", torch=torch.__version__)
    clips = clips_host.to(dev)
    for tag, ac in (("activation_checkpointing", True), ("no_checkpointing", False)):
        rs = None
        try:
            torch.manual_seed(0)
            with torch.device(dev):
                enc, pred = H.build_models(R, model, crop=CROP, frames=FRAMES, pred_depth=PRED["depth"],
                                           pred_heads=PRED["heads"], pred_dim=PRED["dim"], num_mask_tokens=6,
                                           activation_checkpointing=ac, device=dev)
            rs = H.RefStep(R, enc, pred, dict(OPT, loss_exp=1.0), mixed_precision=True)
            del enc, pred
            md = [([m.to(dev) for m in e], [m.to(dev) for m in p]) for e, p in masks_host[:warmup + steps]]
            for i in range(warmup):
                rs.step([clips], [md[i][0]], [md[i][1]])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(warmup, warmup + steps):
                loss = rs.step([clips], [md[i][0]], [md[i][1]])["loss"]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = dict(value=B / (ms / 1e3), ms_per_step=ms, steps=steps, warmup=warmup, loss=loss,
                            peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
            if log:
                log(f"reference eager CUDA ({tag}): {ms:.1f} ms/step = {B / (ms / 1e3):.2f} clips/s, loss {loss:.4f}")
        except torch.OutOfMemoryError as ex:
            out[tag] = dict(value=None, error="out of memory: " + str(ex)[:120])
        except Exception as ex:  # a baseline failure must not take our numbers down
            out[tag] = dict(value=None, error=f"{type(ex).__name__}: {str(ex)[:200]}")
        del rs
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
    return out


# ---------------------------------------------------------------------------------------------- other configs
ALL_CONFIGS = {   # BASELINE.json `configs` 1-4: (model, frames, crop, batch per GPU)
    "C1 ViT-L/16 16x256x256 B=24": ("vit_large", 16, 256, 24),
    "C2 ViT-H/16 16x256x256 B=24": ("vit_huge", 16, 256, 24),
    "C3 ViT-g/16 16x256x256 B=24": ("vit_giant_xformers", 16, 256, 24),
    "C4 ViT-g/16 cooldown 64x384x384 B=6": ("vit_giant_xformers", 64, 384, 6),
}


def quick_config(T, MaskCollator, model, frames, crop, B, dev, peaks, log, warmup=2, steps=3):
    """Short device-timed run of the same fused step on another BASELINE.json geometry (fresh masks per step)."""
    import torch
    ntok = (frames // TUB) * (crop // PATCH) ** 2
    torch.manual_seed(0)
    with torch.device(dev):
        encoder, predictor = T.init_video_model(
            device=dev, patch_size=PATCH, max_num_frames=frames, tubelet_size=TUB, model_name=model, crop_size=crop,
            pred_depth=PRED["depth"], pred_num_heads=PRED["heads"], pred_embed_dim=PRED["dim"], uniform_power=True,
            use_mask_tokens=True, num_mask_tokens=6, zero_init_mask_tokens=True, use_sdpa=True, use_rope=True,
            use_activation_checkpointing=True)
    step = T.JepaTrainStep(encoder, predictor, **OPT)
    coll = MaskCollator(cfgs_mask=MASK_CFG, dataset_fpcs=[frames], crop_size=(crop, crop), patch_size=(PATCH, PATCH),
                        tubelet_size=TUB)
    torch.manual_seed(239)
    masks = make_masks(coll, B, warmup + steps, frames)
    clips = torch.randn(B, 3, frames, crop, crop, generator=torch.Generator().manual_seed(1000)).to(dev)
    md = [([m.to(dev) for m in e], [m.to(dev) for m in p]) for e, p in masks]
    big = max(range(len(masks)), key=lambda i: sum(int(m.numel()) for m in masks[i][0]) * 4
              + sum(int(m.numel()) for m in masks[i][1]))
    step.step([clips], [md[big][0]], [md[big][1]])
    for i in range(warmup):
        step.step([clips], [md[i][0]], [md[i][1]])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warmup, warmup + steps):
        loss, _, _ = step.step([clips], [md[i][0]], [md[i][1]])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fl = sum(step_flops(model, B, [m.shape[1] for m in masks[i][0]], [m.shape[1] for m in masks[i][1]], ntok)
             for i in range(warmup, warmup + steps)) / steps
    tf = fl / (ms / 1e3) / 1e12
    res = dict(value=B / (ms / 1e3), unit="clips/s", ms_per_step=ms, steps=steps, warmup=warmup + 1, tokens_per_clip=ntok,
               tflops_per_gpu=tf, frac_of_measured_sustained=tf / peaks["sustained"], loss=float(loss.item()))
    log(f"all-configs {model} {frames}x{crop}^2 B={B}: {ms:.1f} ms/step = {res['value']:.2f} clips/s, {tf:.0f} TFLOP/s")
    del step, encoder, predictor, clips, md
    return res


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vit_giant_xformers", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--frames", type=int, default=16, help="frames per clip (16 = pretrain configs, 64 = cooldown)")
    ap.add_argument("--crop", type=int, default=256, help="crop size (256 pretrain, 384 cooldown)")
    ap.add_argument("--rank-local-masks", action="store_true",
                    help="seed the mask stream with 239 + rank instead of the reference's rank-independent seed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the reference's PyTorch-eager CUDA arm")
    ap.add_argument("--no-all-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--profile-ops", action="store_true",
                    help="after the timed region, run one extra step with CUDA events around every C-ABI call and "
                         "print a per-op time table to stderr")
    args = ap.parse_args()
    global FRAMES, CROP, NTOK
    FRAMES, CROP = args.frames, args.crop
    NTOK = (FRAMES // TUB) * (CROP // PATCH) ** 2

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    pretty = {"vit_giant_xformers": "ViT-g/16", "vit_large": "ViT-L/16", "vit_huge": "ViT-H/16"}[args.model]
    metric = f"clips/sec ({pretty} V-JEPA 2 pretrain step, " \
             f"{args.frames}x{args.crop}x{args.crop} clips, fwd+bwd+AdamW+EMA)"
    config = dict(workload=f"{args.model} pretrain step (configs/train/vitg16/"
                           f"{'pretrain-256px-16f' if args.frames == 16 else 'cooldown-384px-64f'}.yaml shapes), "
                           f"batch {args.batch}/GPU, multiblock3d masks (8x0.15 + 2x0.7), predictor depth 12 / 384",
                  global_batch=args.batch * world, tokens_per_clip=NTOK, parallelism=f"dp{world}",
                  l2_policy="working set (>= 2 GB of bf16 weights + activations per step) exceeds the 126 MB L2; no flush",
                  mask_stream="rank-local seed 239 + rank" if args.rank_local_masks else
                  "config seed 239 on every rank (reference: train.py:147-148), fresh draw per step; clips rank-local")

    def log(msg):
        print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    # ------------------------------------------------------------------ reference arm (the reference's CPU path)
    if args.impl == "reference":
        if rank != 0:
            return
        cores = host_threads()
        times, flops, b, kind = cpu_reference_step_time(args.model, max_seconds=200.0, steps=max(1, args.steps),
                                                        warmup=min(1, args.warmup), batch=1, log=log)
        ms = 1e3 * sum(times) / len(times)
        val = b / (ms / 1e3)
        line = dict(metric=metric, value=val, unit="clips/s", impl="reference", n_gpus=args.gpus, device="cpu",
                    steps=len(times),
                    steps_requested=args.steps, warmup=min(1, args.warmup), ms_per_step=ms, higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=config,
                    tokens_per_s=val * NTOK, tflops=sum(flops) / sum(times) / 1e12,
                    cpu_baseline=dict(value=val, unit="clips/s", cores=cores, kind=kind,
                                      sample=cpu_sample_text(kind, len(times), args.batch)),
                    e2e=dict(value=val, unit="clips/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from vjepa2_b200 import ops
    from vjepa2_b200 import train as T
    from vjepa2_b200.masks import MaskCollator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = read_peaks()

    t0 = time.time()
    torch.manual_seed(0)                               # identical init on every rank (DDP broadcasts rank 0's)
    with torch.device(dev):
        encoder, predictor = T.init_video_model(
            device=dev, patch_size=PATCH, max_num_frames=FRAMES, tubelet_size=TUB, model_name=args.model,
            crop_size=CROP, pred_depth=PRED["depth"], pred_num_heads=PRED["heads"], pred_embed_dim=PRED["dim"],
            uniform_power=True, use_mask_tokens=True, num_mask_tokens=6, zero_init_mask_tokens=True, use_sdpa=True,
            use_rope=True, use_activation_checkpointing=True)
    step = T.JepaTrainStep(encoder, predictor, **OPT)     # world > 1: broadcasts rank 0's parameters (DDP construction)
    if world > 1:
        config["grad_allreduce"] = (f"{step.grad_comm} ({'copy engines over peer-mapped symmetric memory' if step.grad_comm == 'peer' else 'torch.distributed all_reduce'}), "
                                    f"{step.grad_sync}, buckets >= {step.bucket_bytes >> 20} MB")
    log(f"models built in {time.time() - t0:.1f}s; encoder params "
        f"{sum(p.numel() for p in step.encoder.parameters()) / 1e6:.1f}M")

    B = args.batch
    n_e2e = 0 if args.no_e2e else args.steps      # e2e replays the SAME mask draws as the timed region
    total = args.warmup + args.steps + 1
    collator = MaskCollator(cfgs_mask=MASK_CFG, dataset_fpcs=[FRAMES], crop_size=(CROP, CROP), patch_size=(PATCH, PATCH),
                            tubelet_size=TUB)
    # Mask stream: the reference seeds EVERY rank with the same config seed (app/vjepa/train.py:147-148, no rank offset)
    # and its collator runs in DataLoader workers whose seeds derive from that generator (base_seed + worker_id), so all
    # ranks draw the SAME (masks_enc, masks_pred) each iteration -- K_enc / K_pred, hence the per-step work, are equal
    # across ranks.  --rank-local-masks gives each rank its own stream instead (then every step waits for the rank
    # with the largest draw: dp_imbalance).  Clips are rank-local in both modes.
    torch.manual_seed(239 + (rank if args.rank_local_masks else 0))
    masks_host = make_masks(collator, B, total)
    gclip = torch.Generator().manual_seed(1000 + rank)
    clips_host = torch.randn(B, 3, FRAMES, CROP, CROP, generator=gclip).pin_memory()
    clips_dev = clips_host.to(dev, non_blocking=True)
    masks_dev = [([m.to(dev) for m in e], [m.to(dev) for m in p]) for e, p in masks_host]
    torch.cuda.synchronize()

    def run_step(i, clips):
        e, p = masks_dev[i]
        return step.step([clips], [e], [p])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # allocator warm-up: the workspace arenas grow to the high-water mark of the largest step seen so far (K_enc /
    # K_pred change with every mask draw), so run the largest planned draw once before the regular warm-up steps --
    # every rank does, so the collectives stay matched -- and no timed step has to grow an arena
    big = max(range(total), key=lambda i: sum(int(m.numel()) for m in masks_host[i][0]) * 4
              + sum(int(m.numel()) for m in masks_host[i][1]))
    run_step(big, clips_dev)
    for i in range(args.warmup):
        run_step(i, clips_dev)
    barrier()
    if rank == 0 and world == 1:   # single-process only: an extra step on one rank would dead-lock the collectives
        # host-side enqueue cost of one step (GPU idle at start, nothing awaited): must stay < GPU time
        t_enq = time.perf_counter()
        run_step(args.warmup - 1 if args.warmup else 0, clips_dev)
        t_enq = time.perf_counter() - t_enq
        log(f"host enqueue time of one step (incl. launch-queue back-pressure): {t_enq * 1e3:.1f} ms")
        # pure host cost: same step with every C-ABI entry point stubbed out (no kernel is launched)
        from vjepa2_b200 import _cabi as _C

        class _Stub:
            def __getattr__(self, name):
                real = getattr(_real_lib, name)
                if name.endswith("_scratch"):
                    return real
                return lambda *a: 0
        _real_lib = _C.load()
        torch.cuda.synchronize()
        _C._lib = _Stub()
        t_py = time.perf_counter()
        run_step(args.warmup - 1 if args.warmup else 0, clips_dev)
        t_py = time.perf_counter() - t_py
        _C._lib = _real_lib
        torch.cuda.synchronize()
        log(f"pure host (Python + ctypes marshalling, kernels stubbed) time of one step: {t_py * 1e3:.1f} ms")
    barrier()

    # ---- timed region: device-resident inputs, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ops.LAUNCHES
    ev0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        loss, _, _ = run_step(i, clips_dev)
    ev1.record()
    torch.cuda.synchronize()
    launches = ops.LAUNCHES - launches0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    barrier()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    loss_val = float(loss.item())
    flops_timed = sum(step_flops(args.model, B, [m.shape[1] for m in masks_host[i][0]],
                                 [m.shape[1] for m in masks_host[i][1]])
                      for i in range(args.warmup, args.warmup + args.steps))
    # data-parallel imbalance: every rank draws its own masks (reference: one collator per rank), K_enc / K_pred are
    # truncated to the rank-local batch minimum, and the step synchronises at the gradient all-reduce -- so a step costs
    # the SLOWEST rank's work.  ratio = sum_steps max_rank(flops) / sum_steps mean_rank(flops) bounds the weak-scaling
    # efficiency any implementation of the reference's algorithm can reach on these draws (1 / ratio).
    imbalance = None
    if world > 1:
        mine = torch.tensor([step_flops(args.model, B, [m.shape[1] for m in masks_host[i][0]],
                                        [m.shape[1] for m in masks_host[i][1]])
                             for i in range(args.warmup, args.warmup + args.steps)], dtype=torch.float64, device=dev)
        allf = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allf, mine)
        allf = torch.stack(allf)
        ratio = float(allf.max(dim=0).values.sum() / allf.mean(dim=0).sum())
        imbalance = dict(max_over_mean_step_flops=ratio, efficiency_bound=1.0 / ratio,
                         mask_stream="rank-local (239 + rank)" if args.rank_local_masks else
                         "shared: every rank seeded with the config seed, as app/vjepa/train.py:147-148 does",
                         note="with rank-local mask draws K_enc / K_pred (truncated to the rank-local batch minimum) "
                              "differ across ranks and the all-reduce makes every step wait for the slowest rank; the "
                              "reference's seeding gives every rank the same draw, hence a ratio of 1")
    clips_s = world * B * args.steps / (ms_total / 1e3)
    tflops_gpu = flops_timed / (ms_total / 1e3) / 1e12          # per GPU (each rank does B clips)

    # ---- e2e: public API with pinned host inputs; every step's H2D (clips + masks, on a side stream, double
    #      buffered by train.HostFeeder) and D2H (loss) are inside the timed region (wall clock)
    e2e = None
    if n_e2e:
        base = args.warmup
        h2d = clips_host.numel() * 4 + sum(m.numel() * 8 for m in masks_host[base][0] + masks_host[base][1])
        pinned_masks = [([m.pin_memory() for m in e], [m.pin_memory() for m in p]) for e, p in masks_host]
        feeder = T.HostFeeder(dev)
        # copy bandwidth alone, for the record
        torch.cuda.synchronize()
        tcp = time.perf_counter()
        _tmp = clips_host.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = clips_host.numel() * 4 / (time.perf_counter() - tcp) / 1e9
        del _tmp
        barrier()
        t_start = time.perf_counter()
        feeder.prefetch([clips_host], [pinned_masks[base][0]], [pinned_masks[base][1]])
        for i in range(base, base + n_e2e):
            c, e, p = feeder.get()
            if i + 1 < base + n_e2e:
                feeder.prefetch([clips_host], [pinned_masks[i + 1][0]], [pinned_masks[i + 1][1]])
            l, _, _ = step.step(c, e, p)
            feeder.release()
            _ = float(l.item())                                 # train.py:468 host read of the loss
        torch.cuda.synchronize()
        dt = time.perf_counter() - t_start
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = dict(value=world * B * n_e2e / float(tt.item()), unit="clips/s", h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=4, steps=n_e2e, ms_per_step=1e3 * float(tt.item()) / n_e2e,
                   h2d_gbs_measured=h2d_gbs,
                   note="H2D of step i+1 overlaps compute of step i (side stream, pinned host memory)")

    # ---- e2e with the DEVICE-side mask collator (csrc/maskgen.cu): same public API, same mask draws (RNG-call-identical to
    #      the host sampler: checked below), but the index tensors are produced on the GPU one step ahead; only the
    #      clips cross PCIe.  Reported next to `e2e`, never instead of it.
    if n_e2e:
        from vjepa2_b200.masks import DeviceMaskCollator
        base = args.warmup
        dmc = DeviceMaskCollator(cfgs_mask=MASK_CFG, dataset_fpcs=[FRAMES], crop_size=(CROP, CROP), patch_size=(PATCH, PATCH),
                                 tubelet_size=TUB, device=dev)
        torch.manual_seed(239 + (rank if args.rank_local_masks else 0))
        dmc.seed_from_torch()
        for _ in range(base):                                    # replay the draws of the warm-up steps
            dmc.draw(FRAMES, B)
        chk_e, chk_p = dmc.draw(FRAMES, B)
        same = all(torch.equal(a.cpu(), b) for a, b in zip(chk_e + chk_p, masks_host[base][0] + masks_host[base][1]))
        if not same:
            raise SystemExit("bench.py: device-side mask collator diverged from the host sampler")
        feeder2 = T.HostFeeder(dev)
        barrier()
        t_start = time.perf_counter()
        dmc.enqueue(FRAMES, B)                                   # draw base + 1 ... (the check above consumed draw `base`)
        feeder2.prefetch([clips_host], [[]], [[]])
        for i in range(n_e2e):
            c, _, _ = feeder2.get()
            me_d, mp_d = dmc.collect()
            if i + 1 < n_e2e:
                dmc.enqueue(FRAMES, B)
                feeder2.prefetch([clips_host], [[]], [[]])
            l, _, _ = step.step(c, [me_d], [mp_d])
            feeder2.release()
            _ = float(l.item())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t_start
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e["device_mask_collator"] = dict(
            value=world * B * n_e2e / float(tt.item()), unit="clips/s", ms_per_step=1e3 * float(tt.item()) / n_e2e,
            h2d_bytes_per_step=clips_host.numel() * 4, d2h_bytes_per_step=4 + 16,
            note="masks drawn on the GPU by vj_mask_collate one step ahead (bit-identical to the host sampler, checked); "
                 "D2H = loss + the four K counts")

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): per-launch CUDA events, one instrumented step
    #      (every rank runs the step -- it contains the gradient all-reduce -- rank 0 instruments it)
    import vjepa2_b200.engine as eng
    recs = []
    real_gemm = ops.gemm

    def timed_gemm(a, b, out, M, N, K, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = real_gemm(a, b, out, M, N, K, **kw)
        e.record()
        recs.append((s, e, 2.0 * M * N * K, (M, N, K, bool(kw.get("gelu")), bool(kw.get("a_mn")), bool(kw.get("b_mn")))))
        return r

    if rank == 0:
        eng.ops.gemm = timed_gemm
    barrier()
    evs, eve = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs.record()
    run_step(total - 1, clips_dev)
    eve.record()
    torch.cuda.synchronize()
    eng.ops.gemm = real_gemm
    roof = None
    if rank == 0:
        g_ms = sum(s.elapsed_time(e) for s, e, _, _ in recs)
        g_fl = sum(f for _, _, f, _ in recs)
        # the dominant launch: fc1 forward (+bias, +GELU) of the target encoder, the largest single share of the step
        hid = MODELS[args.model][3]
        dom = [(s, e, f, k) for s, e, f, k in recs if k[3] and not k[4] and not k[5] and k[1] == hid and k[0] == args.batch * NTOK]
        if not dom:
            dom = recs
        d_ms = sum(s.elapsed_time(e) for s, e, _, _ in dom) / len(dom)
        d_fl = sum(f for _, _, f, _ in dom) / len(dom)
        Md, Nd, Kd = dom[0][3][:3]
        achieved = d_fl / (d_ms / 1e3) / 1e12
        traffic, traffic_src = None, None
        try:        # dram__bytes_read.sum + dram__bytes_write.sum of ONE such launch, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01c_ncu_traffic.json")))["fc1gelu_pair"]
            if (Md, Nd, Kd) == (49152, 6144, 1408):
                traffic, traffic_src = tj["dram_bytes"], "profiles/r01c_ncu_gemm2_gelu_summary.txt (ncu --set full, same kernel, shape and epilogue)"
        except Exception:
            pass
        roof = dict(bound="tensor",
                    kernel=f"vj::gemm2_kernel<0,0,0> (tcgen05 cta_group::2 GEMM, 256x256 tile per CTA pair): fc1 forward +bias +GELU of "
                           f"the target encoder, "
                           f"M={Md} N={Nd} K={Kd}",
                    achieved=achieved, peak=peaks["sustained"], unit="TFLOP/s", frac=achieved / peaks["sustained"],
                    peak_source=f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                    flops_per_launch=d_fl, ms_per_launch=d_ms, launches=len(dom),
                    algorithmic_bytes_per_launch=2.0 * (Md * Kd + Nd * Kd + Md * Nd),
                    traffic=traffic, traffic_source=traffic_src,
                    all_gemms=dict(achieved=g_fl / (g_ms / 1e3) / 1e12, frac=g_fl / (g_ms / 1e3) / 1e12 / peaks["sustained"],
                                   launches=len(recs), gemm_ms_per_step=g_ms,
                                   step_ms_instrumented=evs.elapsed_time(eve),
                                   gemm_share_of_step=g_ms / evs.elapsed_time(eve)))
    barrier()

    if args.profile_ops and rank == 0 and world == 1:
        import collections
        agg = collections.defaultdict(list)
        names = ["gemm", "layernorm_fwd", "layernorm_bwd", "attn_fwd", "attn_bwd", "gather_rows", "im2col_tubelets",
                 "colsum", "l1_loss", "pred_indices", "rope_table", "cast_f32_bf16", "adamw_step", "adam_prepare", "ema_update",
                 "grad_check", "scaler_update"]
        saved = {n: getattr(ops, n) for n in names}

        def wrap(name, fn):
            def w(*a, **kw):
                s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                r = fn(*a, **kw)
                e0.record()
                key = name
                if name == "gemm":
                    M, N, K = a[3], a[4], a[5]
                    key = f"gemm {'wgrad' if kw.get('a_mn') else ('dgrad' if kw.get('b_mn') else 'fwd')} " \
                          f"N={N} K={K}" + (" gelu" if kw.get("gelu") else "") + (" dgelu" if kw.get("dgelu_aux") is not None else "") + \
                          (" rope" if kw.get("rope") is not None else "") + (" +res" if kw.get("residual") is not None else "")
                    agg[key].append((s0, e0, 2.0 * M * N * K))
                else:
                    agg[key].append((s0, e0, 0.0))
                return r
            return w

        for n in names:
            setattr(ops, n, wrap(n, saved[n]))
        evs, eve = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evs.record()
        run_step(total - 1, clips_dev)
        eve.record()
        torch.cuda.synchronize()
        for n in names:
            setattr(ops, n, saved[n])
        rows = []
        for k, v in agg.items():
            ms = sum(s0.elapsed_time(e0) for s0, e0, _ in v)
            fl = sum(f for _, _, f in v)
            rows.append((ms, k, len(v), fl))
        tot = sum(r[0] for r in rows)
        log(f"per-op profile of one step: {evs.elapsed_time(eve):.1f} ms wall, {tot:.1f} ms inside ops")
        for ms, k, n, fl in sorted(rows, reverse=True):
            log(f"  {ms:8.2f} ms {100 * ms / tot:5.1f}%  n={n:4d}  {k}" + (f"  {fl / ms / 1e9:7.1f} TFLOP/s" if fl else ""))

    # ---- the other BASELINE.json configs (C1 ViT-L, C2 ViT-H, C4 ViT-g cooldown 64 x 384^2), short runs of the same
    #      code path, so those rows are driver-measured too (rank 0, N=1 only)
    extras = rank == 0 and world == 1
    all_cfg = None
    if extras:
        del step, encoder, predictor
        step = encoder = predictor = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if extras and not args.no_all_configs:
        all_cfg = {}
        for tag, (mname, fr, cr, bb) in ALL_CONFIGS.items():
            if (mname, fr, cr, bb) == (args.model, args.frames, args.crop, args.batch):
                continue
            try:
                all_cfg[tag] = quick_config(T, MaskCollator, mname, fr, cr, bb, dev, peaks, log)
            except Exception as ex:
                all_cfg[tag] = dict(value=None, error=f"{type(ex).__name__}: {str(ex)[:200]}")
            gc.collect()
            torch.cuda.empty_cache()

    # ---- the reference's own PyTorch-eager CUDA path on this GPU (the like-for-like bar; rank 0, N=1 only)
    tcb = None
    if extras and not args.no_torch_baseline:
        tcb = torch_cuda_baseline(args.model, B, masks_host, clips_host, dev, log=log)
        if tcb.get("activation_checkpointing", {}).get("value"):
            tcb["speedup_vs_shipped_config"] = clips_s / tcb["activation_checkpointing"]["value"]
        if tcb.get("no_checkpointing", {}).get("value"):
            tcb["speedup_vs_no_checkpointing"] = clips_s / tcb["no_checkpointing"]["value"]

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same step on the host cores
    cpu = None
    if extras and not args.no_cpu_baseline:
        cores = host_threads()
        try:
            times, _, b, kind = cpu_reference_step_time(args.model, max_seconds=120.0, steps=1, warmup=0, batch=1, log=log)
            cpu = dict(value=b / (sum(times) / len(times)), unit="clips/s", cores=cores, kind=kind,
                       sample=cpu_sample_text(kind, len(times), B))
        except Exception as ex:  # the CPU leg must never take the GPU numbers down with it
            cpu = dict(value=None, unit="clips/s", cores=cores, kind="port", sample=f"failed: {ex}")

    if rank == 0:
        line = dict(metric=metric, value=clips_s, unit="clips/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=config, tokens_per_s=clips_s * NTOK, clips_per_s_per_gpu=clips_s / world,
                    tflops_per_gpu=tflops_gpu, frac_of_nominal_2250=tflops_gpu / 2250.0,
                    frac_of_measured_sustained=tflops_gpu / peaks["sustained"],
                    frac_of_measured_burst=tflops_gpu / peaks["burst"], loss=loss_val, clocks=clocks,
                    gpu_launches=launches, dp_imbalance=imbalance, e2e=e2e, roofline=roof, cpu_baseline=cpu, torch_cuda_baseline=tcb,
                    all_configs=all_cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
