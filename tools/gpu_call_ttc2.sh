# tcgen05 temporal kernel iteration: selftest (all cases, trace for the big ones), GPU suite, bench with the kernel on / off
D=gpurun_out/${1:-ttc2}; mkdir -p $D
timeout 900 python tools/ttc_selftest.py > $D/selftest.log 2>&1
cat $D/selftest.log
if grep -q "TIMEOUT\|NO OUTPUT\|rc=-" $D/selftest.log; then echo "selftest failed: keeping DAWN_TA_TC=0 for the rest"; export DAWN_TA_TC=0; fi
timeout 300 python bench.py --no-cpu-baseline > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -18
DAWN_TA_TC=0 timeout 300 python bench.py --no-cpu-baseline > $D/bench_off.json 2> $D/bench_off.err
python tools/show_bench.py $D/bench_off.json 2>/dev/null | head -1
python tools/show_bench.py $D/bench_off.json 2>/dev/null | grep temporal
( timeout 900 python -m pytest tests -m gpu -x -q -s > $D/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $D/pytest_gpu.log )
grep -E "passed|failed|error|exit|FAILED|Error|cfg3" $D/pytest_gpu.log | tail -12
