# halo conv fed by TMA (second conv of every ResBlock, planes written by gn_hcond): parity subset + bench, default vs DAWN_CONV3_TMA=0
D=gpurun_out/${1:-tma}; mkdir -p $D
timeout 400 python -m pytest tests/test_unet_gpu.py -q -x -k "golden or submodule or cfg2" > $D/pytest_tma.log 2>&1; tail -4 $D/pytest_tma.log
timeout 300 python bench.py --no-cpu-baseline --no-clip > $D/bench_tma.json 2> $D/bench_tma.err
python tools/show_bench.py $D/bench_tma.json 2>/dev/null | grep -E "ms/step|conv3|gn_hcond"
DAWN_CONV3_TMA=0 timeout 300 python bench.py --no-cpu-baseline --no-clip > $D/bench_def.json 2> $D/bench_def.err
python tools/show_bench.py $D/bench_def.json 2>/dev/null | grep -E "ms/step|conv3|gn_hcond"
