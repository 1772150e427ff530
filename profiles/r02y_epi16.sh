cd /root/repo/vjepa2_b200/csrc
echo "== correctness: pair kernel + 16 epilogue warps forced"
VJ_GEMM_2CTA=2 VJ_GEMM_EPI16=1 timeout 100 ./build/selftest gemm 2>&1 | grep -E "FAIL|PASSED|FAILED|watchdog" | head -5
for v in 0 1 0 1; do echo "== benchpred VJ_GEMM_EPI16=$v"; VJ_GEMM_EPI16=$v timeout 60 ./build/selftest benchpred 2>&1 | grep "bench gemm\|watchdog" | head -12; done
for v in 0 1; do echo "== benchepi VJ_GEMM_EPI16=$v"; VJ_GEMM_EPI16=$v timeout 60 ./build/selftest benchepi 2>&1 | grep "bench gemm\|watchdog" | head -14; done
