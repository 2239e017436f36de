D=gpurun_out/${1:-gt}; mkdir -p $D; rm -f $D/trace.log
for c in "100 32 32 256 128 1 0" "100 64 64 64 128 1 0" "100 32 32 128 768 1 0"; do
  DAWN_TC_TRACE=1 timeout 120 python tools/tc_selftest.py $c >> $D/trace.log 2>&1
  DAWN_TC_TRACE=1 DAWN_TC_SHIFT=64 timeout 120 python tools/tc_selftest.py $c >> $D/trace.log 2>&1
done
grep -E "trace" $D/trace.log
