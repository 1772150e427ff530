"""Feasibility probe (2 GPUs): peer-mapped memory between ranks, copy-engine bandwidth, NCCL all-reduce time."""
import os, sys, time
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
def log(*a):
    print(f"[r{rank}]", *a, flush=True)

n = 1 << 28   # 1 GiB fp32
# ---- NCCL all-reduce alone
g = torch.ones(1035 * 1000 * 1000, dtype=torch.float32, device=dev)   # ~4.14 GB
for _ in range(2):
    dist.all_reduce(g)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    dist.all_reduce(g)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
log(f"NCCL all_reduce 4.14 GB fp32: {ms:.2f} ms -> algbw {g.numel()*4/ms/1e6:.0f} GB/s")
del g

# ---- classic CUDA IPC through torch storage sharing
try:
    buf = torch.zeros(n, dtype=torch.float32, device=dev)
    h = buf.untyped_storage()._share_cuda_()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    peer = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*hs[peer])
    pbuf = torch.empty(0, dtype=torch.float32, device=dev).set_(st, 0, (n,), (1,))
    log("IPC mapped peer buffer:", pbuf.shape, pbuf.device, hex(pbuf.data_ptr()))
    src = torch.full((n,), float(rank + 1), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    for nbytes in (100 << 20, 1 << 30):
        k = nbytes // 4
        pbuf[:k].copy_(src[:k]); torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            pbuf[:k].copy_(src[:k])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        log(f"P2P push {nbytes>>20} MiB via tensor.copy_: {ms:.3f} ms -> {nbytes/ms/1e6:.0f} GB/s")
    dist.barrier(); torch.cuda.synchronize()
    log("peer wrote into my buffer:", float(buf[0]), float(buf[n - 1]), "expected", float(((rank - 1) % world) + 1))
    # does the copy use SMs?  run a long compute kernel filling all SMs and a concurrent copy on a side stream
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    side = torch.cuda.Stream()
    def mm(k=40):
        for _ in range(k):
            a @ a
    mm(5); torch.cuda.synchronize()
    e0.record(); mm(); e1.record(); torch.cuda.synchronize()
    t_alone = e0.elapsed_time(e1)
    e0.record()
    with torch.cuda.stream(side):
        for _ in range(8):
            pbuf.copy_(src)
    mm()
    e1.record(); torch.cuda.synchronize()
    log(f"matmul loop alone {t_alone:.1f} ms, with 8 GiB concurrent P2P push {e0.elapsed_time(e1):.1f} ms")
except Exception as ex:
    import traceback; traceback.print_exc()
    log("IPC path failed:", repr(ex))

# ---- torch symmetric memory
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    pb = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    pb.fill_(float(rank))
    hdl.barrier()
    log("symm_mem ok; my buffer now holds", float(t[0]))
except Exception as ex:
    log("symm_mem failed:", repr(ex)[:300])
dist.barrier()
dist.destroy_process_group()
