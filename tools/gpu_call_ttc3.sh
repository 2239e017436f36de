D=gpurun_out/${1:-ttc3}; mkdir -p $D
timeout 900 python tools/ttc_selftest.py > $D/selftest.log 2>&1
cat $D/selftest.log
if grep -q "TIMEOUT\|NO OUTPUT\|rc=-" $D/selftest.log; then echo "selftest failed: keeping DAWN_TA_TC=0 for the rest"; export DAWN_TA_TC=0; fi
timeout 300 python bench.py --no-cpu-baseline > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -18
