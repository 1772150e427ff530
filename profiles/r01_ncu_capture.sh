cd /root/repo/vjepa2_b200/csrc
timeout 100 ./build/selftest bench > /root/repo/gpurun_out/st_plain.log 2>&1 && timeout 100 ./build/selftest benchbwd >> /root/repo/gpurun_out/st_plain.log 2>&1 || exit 1
for spec in "gemm_kernel 3 qkv bench" "gemm_kernel 29 fc1gelu bench" "attn_fwd2_kernel 2 attnfwd bench" "attn_bwd2_kernel 2 attnbwd benchbwd"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o /root/repo/gpurun_out/r01b_$3 ./build/selftest $4 > /root/repo/gpurun_out/ncu_$3.log 2>&1
  tail -1 /root/repo/gpurun_out/ncu_$3.log
done
cd /root/repo
timeout 300 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_plain_bench.log 2>&1 || exit 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 2600 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r01b.csv
