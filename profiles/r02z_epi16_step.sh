cd /root/repo/vjepa2_b200/csrc
echo "== correctness: pair kernel + 16 epilogue warps forced"
VJ_GEMM_2CTA=2 VJ_GEMM_EPI16=1 timeout 100 ./build/selftest gemm 2>&1 | grep -E "FAIL|PASSED|FAILED|watchdog" | head -5
cd /root/repo
VJ_GEMM_EPI16=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k gemm 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -1
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-baseline --no-all-configs"
for rep in 1 2 3; do
for v in auto 0; do
  if [ $v = auto ]; then unset VJ_GEMM_EPI16; else export VJ_GEMM_EPI16=0; fi
  python bench.py $B 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('epi16=$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
done
