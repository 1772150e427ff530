"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one process per GPU, NCCL).

Checks the data-parallel semantics of app/vjepa/train.py:279-281 on real GPUs: N ranks x B clips must equal one rank
x N*B clips -- loss (mean of the rank losses), gradients (mean over ranks) and the weights / target weights / Adam
moments after the optimizer + EMA -- and every rank must end the step with identical weights, even though rank > 0
starts from DIFFERENT initial weights (DDP broadcasts rank 0's at construction)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

from test_gpu_models import build_models  # noqa: E402
from golden_common import OPT_CFG, NTOK  # noqa: E402


def relerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from vjepa2_b200.train import JepaTrainStep
    B = 2
    g = torch.Generator().manual_seed(77)
    clips = torch.randn(world * B, 3, 8, 96, 96, generator=g)
    me, mp = [], []
    for K_e, K_p in ((40, 72), (24, 88)):
        e, p = [], []
        for _ in range(world * B):
            perm = torch.randperm(NTOK, generator=g)
            e.append(perm[:K_e].sort().values)
            p.append(perm[K_e:K_e + K_p].sort().values)
        me.append(torch.stack(e))
        mp.append(torch.stack(p))

    enc, pred, _, _ = build_models(dev)
    if rank > 0:                                   # a replica that was seeded differently: must be overwritten by rank 0's
        with torch.no_grad():
            for p in list(enc.parameters()) + list(pred.parameters()):
                p.add_(0.01 * rank)
    step = JepaTrainStep(enc, pred, **OPT_CFG)
    sl = slice(rank * B, (rank + 1) * B)
    steps = 2
    losses, g_first = [], None
    for it in range(steps):
        loss, _, _ = step.step([clips[sl].to(dev)], [[m[sl].to(dev) for m in me]], [[m[sl].to(dev) for m in mp]])
        losses.append(loss.clone())
        if it == 0:
            g_first = (step.enc_rt.fs.g32.clone(), step.pred_rt.fs.g32.clone())
    efs, pfs, tfs = step.enc_rt.fs, step.pred_rt.fs, step.tgt_rt.fs
    # every rank holds the same state
    for t in (efs.p32, pfs.p32, tfs.p32, efs.exp_avg, efs.g32, pfs.g32):
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, t), "replicas diverged"
    lsum = torch.stack(losses).view(-1).clone()
    dist.all_reduce(lsum)
    lmean = lsum / world

    if rank == 0:
        enc1, pred1, _, _ = build_models(dev)
        one = JepaTrainStep(enc1, pred1, process_group=False, **OPT_CFG)      # single-process arm: no collectives
        l1, g1 = [], None
        for it in range(steps):
            loss, _, _ = one.step([clips.to(dev)], [[m.to(dev) for m in me]], [[m.to(dev) for m in mp]])
            l1.append(float(loss.item()))
            if it == 0:
                g1 = (one.enc_rt.fs.g32.clone(), one.pred_rt.fs.g32.clone())
        # step 1 (identical weights on both arms): loss and gradients agree to fp32 summation-order noise.  The rank
        # buffers hold the SUM of the ranks' scaled gradients (the mean lives in inv_scale = 1 / (scale * world)).
        assert abs(lmean[0].item() - l1[0]) < 1e-5, ("loss", lmean[0].item(), l1[0])
        ge = relerr(g_first[0] / world, g1[0])
        gp = relerr(g_first[1] / world, g1[1])
        assert ge < 1e-3 and gp < 1e-3, ("grads", ge, gp)
        # step 2 starts from the updated weights: Adam's first update is ~lr * sign(g), so elements whose gradient sits at
        # the summation-noise floor may move the other way -- weights agree to a fraction of lr, the loss to 1e-4
        assert abs(lmean[1].item() - l1[1]) < 1e-4, ("loss step 2", lmean[1].item(), l1[1])
        e1, p1, t1 = one.enc_rt.fs, one.pred_rt.fs, one.tgt_rt.fs
        for nm, a, b in (("enc", efs.p32, e1.p32), ("pred", pfs.p32, p1.p32), ("tgt", tfs.p32, t1.p32)):
            assert relerr(a, b) < 1e-3, (nm, relerr(a, b))
        assert relerr(efs.exp_avg_sq, e1.exp_avg_sq) < 2e-2
        print(f"ddp_worker ok: world {world}, losses {l1}, step-1 grad err enc {ge:.2e} pred {gp:.2e}, "
              f"weights err {relerr(efs.p32, e1.p32):.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
