# A/B on one box: bias gradients from the wgrad GEMM's ones-column vs the column-sum kernels
cd /root/repo
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
for i in 1 2 3; do
  for v in 1 0; do
    VJ_BIAS_GRAD_PAD=$v python bench.py $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pad=$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['gpu_launches'])"
  done
done
