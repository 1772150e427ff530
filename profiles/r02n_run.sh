set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02n_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02n_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02n_bench_default.json 2> gpurun_out/r02n_bench_default.log; echo "bench rc=$?"
tail -4 gpurun_out/r02n_bench_default.log
python -c "
import json
d=json.loads(open('gpurun_out/r02n_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e'])
print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d['all_configs'].items()})
print(d['torch_cuda_baseline'])
"
