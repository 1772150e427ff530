cd /root/repo/vjepa2_b200/csrc
for p in 4 2; do
  echo "== VJ_ATTN_BWD_PARTS=$p"
  VJ_ATTN_BWD_PARTS=$p timeout 120 ./build/selftest attnbwd 2>&1 | grep -E "attn\]|FAIL|PASSED|FAILED" | tail -12
  VJ_ATTN_BWD_PARTS=$p timeout 120 ./build/selftest benchbwd 2>&1 | grep "bwd\]"
  VJ_ATTN_BWD_PARTS=$p timeout 120 ./build/selftest_prof benchbwd 2>&1 | grep -A2 "bwd\]" | grep -v "^--"
done
for q in 1 2 4; do echo "== parts 4 poly $q"; VJ_ATTN_BWD_POLY=$q timeout 120 ./build/selftest benchbwd 2>&1 | grep "bwd\]"; done
