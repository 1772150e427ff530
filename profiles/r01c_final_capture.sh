# Final round-1 measurements (run on one B200 through gpurun): default bench line, reference arm, ncu launch list of one
# step, ncu --set full of the CTA-pair GEMM (fc1 fwd + GELU) and of the attention forward (poly share 2/8).
set -x
cd /root/repo
timeout 600 python bench.py > gpurun_out/r01c_bench_default.json 2> gpurun_out/r01c_bench_default.log || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01c_bench_reference.json 2> gpurun_out/r01c_bench_reference.log || exit 2
timeout 300 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_plain_bench.log 2>&1 || exit 3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6600 -c 2500 --csv --log-file gpurun_out/r01c_launches_vitg_step.csv python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/r01c_launches_vitg_step.csv
cd vjepa2_b200/csrc
timeout 100 ./build/selftest benchattn > /root/repo/gpurun_out/st_benchattn.log 2>&1 || exit 4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2_kernel -s 14 -c 1 -f -o /root/repo/gpurun_out/r01c_attnfwd ./build/selftest benchattn > /root/repo/gpurun_out/ncu_attnfwd.log 2>&1
tail -1 /root/repo/gpurun_out/ncu_attnfwd.log
