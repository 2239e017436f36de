D=gpurun_out/${1:-quick}; mkdir -p $D
timeout 400 python -m pytest tests/test_unet_gpu.py -q -x -k "golden or submodule or cfg2" > $D/pytest.log 2>&1; tail -2 $D/pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-clip > $D/bench.json 2> $D/bench.err
python tools/show_bench.py $D/bench.json 2>/dev/null | head -16
