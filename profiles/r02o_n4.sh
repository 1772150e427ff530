cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 $TR --master-port 29711 bench.py --gpus 4 $B 2> gpurun_out/r02o_n4_peer.log | tail -1 > gpurun_out/r02o_n4_peer.json
timeout 200 python bench.py --gpus 1 $B 2> gpurun_out/r02o_n1.log | tail -1 > gpurun_out/r02o_n1.json
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
for f in gpurun_out/r02o_n*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['clocks'], d['config'].get('grad_allreduce'))"; done
