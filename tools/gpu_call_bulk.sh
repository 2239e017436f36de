# bulk-async epilogue stores: gemm/conv selftests (vs mma.sync), traces, parity subset, bench
D=gpurun_out/${1:-bulk}; mkdir -p $D; rm -f $D/trace.log
timeout 600 python tools/tc_selftest.py > $D/selftest.log 2>&1; cat $D/selftest.log | cut -c1-150
for c in "100 64 64 64 128 1 0" "100 32 32 128 768 1 0"; do DAWN_TC_TRACE=1 timeout 120 python tools/tc_selftest.py $c >> $D/trace.log 2>&1; done
DAWN_SELFTEST_CONV3=1 DAWN_TC_TRACE=1 timeout 120 python tools/tc_selftest.py 100 64 64 64 64 3 1 >> $D/trace.log 2>&1
grep -E "trace" $D/trace.log | sed 's/producer t0.*loader wait slot [0-9]* | //'
timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_lfg_gpu.py -q -x -k "golden or submodule or cfg2 or lfg" > $D/pytest.log 2>&1; tail -3 $D/pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-clip > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -16
