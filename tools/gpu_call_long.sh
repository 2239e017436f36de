# long-sequence selection change: the edge-geometry tests (incl. the 250-frame two-segment case), the reference-golden tests and a short bench
D=gpurun_out/${1:-long}; mkdir -p $D
( timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -s -k "edge or golden or window" > $D/pytest.log 2>&1; echo "pytest exit $?" >> $D/pytest.log )
grep -E "x tol|passed|failed|exit|Error" $D/pytest.log | tail -16
timeout 200 python bench.py --no-cpu-baseline --no-clip > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -2
