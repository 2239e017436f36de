# selftest (stops at the first hang) + short bench, no pytest
D=gpurun_out/${1:-single2}; mkdir -p $D
timeout 900 python tools/ttc_selftest.py > $D/selftest.log 2>&1
grep -v "^ttc.*, [0-9], 40\|^ttc (1,\|^ttc (64" $D/selftest.log | cut -c1-400
if grep -q "TIMEOUT\|NO OUTPUT\|rc=-" $D/selftest.log; then echo "selftest failed"; exit 1; fi
timeout 300 python bench.py --no-cpu-baseline --no-clip > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -18 | grep "ms/step\|temporal"
