# round-2 final: launch list of one step (ncu, serialised), per-op profile, bandwidth-class DRAM counters
set -x
cd /root/repo
F="--no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
timeout 300 python bench.py --steps 1 --warmup 3 $F > gpurun_out/r02r_plain.log 2>&1 || exit 3
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 8600 -c 2300 --csv --log-file gpurun_out/r02r_launches_vitg_step.csv python bench.py --steps 1 --warmup 3 $F > gpurun_out/r02r_ncu_bench.log 2>&1
wc -l gpurun_out/r02r_launches_vitg_step.csv
timeout 300 python bench.py --steps 3 --warmup 3 $F --profile-ops > gpurun_out/r02r_ops_profile.log 2>&1
grep -c "bench r0" gpurun_out/r02r_ops_profile.log
