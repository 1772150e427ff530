set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02j_pytest_gpu.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs 2>/dev/null | tail -1 > gpurun_out/r02j_n1.json
python -c "import json; d=json.load(open('gpurun_out/r02j_n1.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['gpu_launches'])"
