# single-query-tile mode of the tcgen05 temporal attention: selftest (time-boxed per case; stops at the first hang), then the GPU suite
# and a short bench only if the selftest is clean
D=gpurun_out/${1:-single}; mkdir -p $D
timeout 900 python tools/ttc_selftest.py > $D/selftest.log 2>&1
cat $D/selftest.log
if grep -q "TIMEOUT\|NO OUTPUT\|rc=-" $D/selftest.log; then echo "selftest failed"; exit 1; fi
( timeout 900 python -m pytest tests -m gpu -x -q > $D/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $D/pytest_gpu.log )
tail -4 $D/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -18
