cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
for rep in 1 2; do
for v in 2 0 3 4; do
  VJ_ATTN_POLY=$v python bench.py $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('poly=$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
done
