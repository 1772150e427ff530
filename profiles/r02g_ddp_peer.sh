# 2 GPUs: gradient all-reduce over peer-mapped memory + copy engines (PeerGradReducer) against the NCCL path
set -x
cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02g_pytest_multi.log 2>&1; echo "multi rc=$?"
tail -25 gpurun_out/r02g_pytest_multi.log
timeout 200 python bench.py --gpus 1 $B 2> gpurun_out/r02g_n1.log | tail -1 > gpurun_out/r02g_n1.json
timeout 200 $TR --master-port 29601 bench.py --gpus 2 $B 2> gpurun_out/r02g_n2_peer.log | tail -1 > gpurun_out/r02g_n2_peer.json
VJ_DDP_SYNC=end timeout 200 $TR --master-port 29602 bench.py --gpus 2 $B 2> gpurun_out/r02g_n2_peer_end.log | tail -1 > gpurun_out/r02g_n2_peer_end.json
VJ_DDP_COMM=nccl timeout 200 $TR --master-port 29603 bench.py --gpus 2 $B 2> gpurun_out/r02g_n2_nccl.log | tail -1 > gpurun_out/r02g_n2_nccl.json
for f in gpurun_out/r02g_n*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['clocks'], d['config'].get('grad_allreduce'))"; done
tail -5 gpurun_out/r02g_n2_peer.log
