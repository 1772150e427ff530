# 2 GPUs, mask stream seeded as the reference does (same seed on every rank): what is left of the +11 ms at N=2?
set -x
cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 $B 2> gpurun_out/r02e_n1.log | tail -1 > gpurun_out/r02e_n1.json
timeout 300 $TR --master-port 29601 bench.py --gpus 2 $B 2> gpurun_out/r02e_n2_overlap.log | tail -1 > gpurun_out/r02e_n2_overlap.json
VJ_DDP_SYNC=end timeout 300 $TR --master-port 29602 bench.py --gpus 2 $B 2> gpurun_out/r02e_n2_end.log | tail -1 > gpurun_out/r02e_n2_end.json
timeout 300 $TR --master-port 29603 bench.py --gpus 2 $B --rank-local-masks 2> gpurun_out/r02e_n2_ranklocal.log | tail -1 > gpurun_out/r02e_n2_ranklocal.json
for f in gpurun_out/r02e_n*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['clocks'], (d.get('dp_imbalance') or {}).get('max_over_mean_step_flops'))"; done
