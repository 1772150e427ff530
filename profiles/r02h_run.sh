# fused LayerNorm backward (dx + dgamma/dbeta + bias gradient in one pass), prefetching LayerNorm forward
set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02h_pytest_gpu.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs --profile-ops > gpurun_out/r02h_ops_profile.log 2>&1
grep -E "layernorm|colsum|per-op profile" gpurun_out/r02h_ops_profile.log
tail -1 gpurun_out/r02h_ops_profile.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])"
