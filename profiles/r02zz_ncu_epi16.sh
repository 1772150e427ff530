cd /root/repo/vjepa2_b200/csrc
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 30 -c 1 -f -o /root/repo/gpurun_out/r02zz_gemm2_epi16 ./build/selftest benchpred > /root/repo/gpurun_out/r02zz_ncu_epi16.log 2>&1
VJ_GEMM_EPI16=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 30 -c 1 -f -o /root/repo/gpurun_out/r02zz_gemm2_epi8 ./build/selftest benchpred > /root/repo/gpurun_out/r02zz_ncu_epi8.log 2>&1
tail -1 /root/repo/gpurun_out/r02zz_ncu_epi8.log
