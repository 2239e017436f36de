# validation call: GPU test-suite and the default bench line
D=gpurun_out/${1:-val}; mkdir -p $D
( timeout 700 python -m pytest tests -m gpu -q -s > $D/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $D/pytest_gpu.log )
timeout 400 python bench.py > $D/bench.json 2> $D/bench.err
grep -E "passed|failed|error|exit|x tol|FAILED|Error" $D/pytest_gpu.log | tail -25
python tools/show_bench.py $D/bench.json 2>/dev/null | head -20
