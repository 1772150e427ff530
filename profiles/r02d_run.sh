# device-side MaskCollator tests, pair-GEMM stress (short watchdog build), CPU-side suite sanity on the box
set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mask_collator" > gpurun_out/r02d_pytest_mask.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_pytest_mask.log
tail -15 gpurun_out/r02d_pytest_mask.log
cd vjepa2_b200/csrc
timeout 300 ./build/selftest_stress stressgemm 200 > /root/repo/gpurun_out/r02d_stressgemm.log 2>&1; echo "stress rc=$?"
cat /root/repo/gpurun_out/r02d_stressgemm.log
