# round 2, second session first call: full -m gpu suite, default bench line (torch_cuda_baseline / all_configs keys),
# per-op profile, ncu dram bytes for the bandwidth class in-step, ncu --set full of attn_bwd2<32>, in-kernel counters.
set -x
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -5 gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.log; echo "bench rc=$?"
tail -3 gpurun_out/r02c_bench_default.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs --profile-ops > gpurun_out/r02c_ops_profile.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'ln_|colreduce|gather_rows|scatter_add|im2col|l1_loss|adamw|ema_kernel|grad_check|cast_|delta|dq_convert|rope_table|pred_indices' \
  -s 1500 -c 1200 --csv --log-file gpurun_out/r02c_bw_class.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs > gpurun_out/r02c_ncu_bw.log 2>&1
wc -l gpurun_out/r02c_bw_class.csv
cd vjepa2_b200/csrc
timeout 120 ./build/selftest_prof benchbwd > /root/repo/gpurun_out/r02c_prof_benchbwd.log 2>&1
timeout 120 ./build/selftest_prof benchattn > /root/repo/gpurun_out/r02c_prof_benchattn.log 2>&1
timeout 120 ./build/selftest benchbwd > /root/repo/gpurun_out/r02c_benchbwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2_kernel -s 10 -c 1 -f -o /root/repo/gpurun_out/r02c_attnbwd32 ./build/selftest benchbwd > /root/repo/gpurun_out/r02c_ncu_attnbwd32.log 2>&1
tail -2 /root/repo/gpurun_out/r02c_ncu_attnbwd32.log
