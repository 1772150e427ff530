# round 2: where do the +16 ms at N=2 come from?  Variants of the gradient all-reduce schedule on 2 GPUs of one box.
set -x
cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02b_pytest_multi.log 2>&1; echo "multi rc=$?"
tail -3 gpurun_out/r02b_pytest_multi.log
timeout 300 python bench.py --gpus 1 $B > gpurun_out/r02b_n1.json 2> gpurun_out/r02b_n1.log
timeout 300 $TR --master-port 29601 bench.py --gpus 2 $B > gpurun_out/r02b_n2_overlap.json 2> gpurun_out/r02b_n2_overlap.log
VJ_DDP_SYNC=end timeout 300 $TR --master-port 29602 bench.py --gpus 2 $B > gpurun_out/r02b_n2_end.json 2> gpurun_out/r02b_n2_end.log
NCCL_MAX_CTAS=4 timeout 300 $TR --master-port 29603 bench.py --gpus 2 $B > gpurun_out/r02b_n2_overlap_cta4.json 2> gpurun_out/r02b_n2_overlap_cta4.log
NCCL_MAX_CTAS=2 VJ_DDP_BUCKET_MB=512 timeout 300 $TR --master-port 29604 bench.py --gpus 2 $B > gpurun_out/r02b_n2_overlap_cta2_b512.json 2> gpurun_out/r02b_n2_overlap_cta2_b512.log
VJ_DDP_BUCKET_MB=1024 timeout 300 $TR --master-port 29605 bench.py --gpus 2 $B > gpurun_out/r02b_n2_overlap_b1024.json 2> gpurun_out/r02b_n2_overlap_b1024.log
for f in gpurun_out/r02b_n*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['clocks'])"; done
