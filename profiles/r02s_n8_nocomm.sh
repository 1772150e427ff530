cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
VJ_DDP_COMM=none timeout 300 $TR --master-port 29721 bench.py --gpus 8 $B 2> gpurun_out/r02s_n8_none.log | tail -1 > gpurun_out/r02s_n8_none.json
timeout 300 $TR --master-port 29722 bench.py --gpus 8 $B 2> gpurun_out/r02s_n8_peer.log | tail -1 > gpurun_out/r02s_n8_peer.json
VJ_DDP_BUCKET_MB=64 timeout 300 $TR --master-port 29723 bench.py --gpus 8 $B 2> gpurun_out/r02s_n8_peer_b64.log | tail -1 > gpurun_out/r02s_n8_peer_b64.json
for f in gpurun_out/r02s_n8*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['clocks'], d['config'].get('grad_allreduce'))"; done
