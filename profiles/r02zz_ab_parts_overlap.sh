cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
run() { env "$@" python bench.py $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
run X=0
run VJ_ATTN_BWD_PARTS=4
run VJ_OVERLAP_TARGET=1
done
