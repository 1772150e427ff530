cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or rope" 2>&1 | tail -3
cd vjepa2_b200/csrc
for v in 1 0; do echo "== VJ_GEMM_SLAB=$v"; VJ_GEMM_SLAB=$v ./build/selftest benchepi 2>&1 | grep "bench gemm"; done
