cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for v in 1 0 1 0; do
i=$((i+1))
VJ_SUM_SMALL=$v timeout 300 $TR --master-port 2973$i bench.py --gpus 8 $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sum_small=$v', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
VJ_DDP_COMM=none timeout 300 $TR --master-port 29739 bench.py --gpus 8 $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('none', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -1
