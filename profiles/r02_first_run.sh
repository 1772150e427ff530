# round 2, first GPU call: full -m gpu suite (incl. full-width parity vs the live reference), default bench line with the
# new torch_cuda_baseline / all_configs keys, reference arm.
set -x
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02a_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.log; echo "bench rc=$?"
tail -3 gpurun_out/r02a_bench_default.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.log; echo "ref rc=$?"
