# validation call: GPU test-suite, default bench line, A/B of the per-clip prep kernels (e2e)
D=gpurun_out/${1:-val}; mkdir -p $D
( timeout 600 python -m pytest tests -m gpu -x -q -s > $D/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $D/pytest_gpu.log )
timeout 400 python bench.py > $D/bench.json 2> $D/bench.err
DAWN_PREP_V1=1 timeout 300 python bench.py --no-cpu-baseline > $D/bench_prep_v1.json 2> $D/bench_prep_v1.err
grep -E "passed|failed|error|exit|graph sampler|general entry" $D/pytest_gpu.log | tail -12
python tools/show_bench.py $D/bench.json | head -3; python tools/show_bench.py $D/bench_prep_v1.json | head -1
