cd /root/repo/vjepa2_b200/csrc
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_fwd_kernel -s 3 -c 1 -f -o /root/repo/gpurun_out/r02q_lnfwd ./build/selftest benchbw > /root/repo/gpurun_out/r02q_ncu_lnfwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_fused_kernel -s 3 -c 1 -f -o /root/repo/gpurun_out/r02q_lnbwd ./build/selftest benchbw > /root/repo/gpurun_out/r02q_ncu_lnbwd.log 2>&1
tail -2 /root/repo/gpurun_out/r02q_ncu_lnbwd.log
