# round-2 final state: build check, full -m gpu suite, smoke, default bench line, reference arm
set -x
cd /root/repo
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02zz_build_smoke.log 2>&1; echo "build+smoke rc=$?"
tail -2 gpurun_out/r02zz_build_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02zz_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02zz_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02zz_bench_default.json 2> gpurun_out/r02zz_bench_default.log; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02zz_bench_reference.json 2> gpurun_out/r02zz_bench_reference.log; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02zz_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['tflops_per_gpu'], d['frac_of_measured_sustained'], d['gpu_launches'])
print('e2e', d['e2e']['value'], d['e2e']['device_mask_collator']['value'])
print({k:round(v['value'],2) for k,v in d['all_configs'].items()})
print({k:(round(v['value'],2) if isinstance(v,dict) else v) for k,v in d['torch_cuda_baseline'].items() if k in ('activation_checkpointing','no_checkpointing')})
r=json.loads(open('gpurun_out/r02zz_bench_reference.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
"
