D=gpurun_out/${1:-gt}; mkdir -p $D; rm -f $D/trace.log
for c in "100 64 64 64 128 1 0" "100 32 32 128 768 1 0" "100 32 32 256 128 1 0"; do
  echo "== case $c" >> $D/trace.log
  DAWN_TC_TRACE=1 timeout 120 python tools/tc_selftest.py $c >> $D/trace.log 2>&1
done
grep -E "trace|==" $D/trace.log | sed 's/producer t0.*loader wait slot [0-9]* | //'
timeout 300 python bench.py --no-cpu-baseline --no-clip > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -6
