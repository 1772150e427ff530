cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mb in 64 192 64 192; do
VJ_DDP_BUCKET_MB=$mb timeout 300 $TR --master-port 297$mb bench.py --gpus 2 $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bucket $mb', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
VJ_DDP_COMM=none timeout 300 $TR --master-port 29750 bench.py --gpus 2 $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('none', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
