D=gpurun_out/${1:-c3}; mkdir -p $D; rm -f $D/trace.log
for c in "100 64 64 64 64 3 1" "100 32 32 128 128 3 1" "100 64 64 128 64 3 1"; do
  DAWN_SELFTEST_CONV3=1 DAWN_TC_TRACE=1 timeout 120 python tools/tc_selftest.py $c >> $D/trace.log 2>&1
done
grep -E "trace|case" $D/trace.log
timeout 400 python -m pytest tests/test_unet_gpu.py -q -x -k "golden or submodule or cfg2" > $D/pytest.log 2>&1; tail -3 $D/pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-clip > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | grep -E "ms/step|conv3|conv_other"
