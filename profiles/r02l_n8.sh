# 8 GPUs of one box: peer-mapped copy-engine all-reduce (default) and the NCCL path, same mask stream on every rank
cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 bench.py --gpus 8 $B 2> gpurun_out/r02l_n8_peer.log | tail -1 > gpurun_out/r02l_n8_peer.json
VJ_DDP_COMM=nccl timeout 300 $TR --master-port 29702 bench.py --gpus 8 $B 2> gpurun_out/r02l_n8_nccl.log | tail -1 > gpurun_out/r02l_n8_nccl.json
timeout 200 python bench.py --gpus 1 $B 2> gpurun_out/r02l_n1.log | tail -1 > gpurun_out/r02l_n1.json
for f in gpurun_out/r02l_n*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['clocks'], d['config'].get('grad_allreduce'), d.get('dp_imbalance',{}) and d['dp_imbalance'].get('max_over_mean_step_flops'))"; done
grep -i "warn\|error\|Traceback" gpurun_out/r02l_n8_peer.log | head
