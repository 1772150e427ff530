cd vjepa2_b200/csrc
timeout 280 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 55 -c 1 -f -o /root/repo/gpurun_out/r01c_gemm2_gelu ./build/selftest benchepi > /root/repo/gpurun_out/ncu_gemm2_gelu.log 2>&1
tail -3 /root/repo/gpurun_out/ncu_gemm2_gelu.log
