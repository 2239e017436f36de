# Round-2 opener (one 1-GPU call, ~6 min): (1) A/B of the prepared halo-conv experiment, (2) whole-clip pipeline timing,
# (3) source-level ncu of the two dominant kernels with the per-line stall CSV exported on the box (the .ncu-rep stays there:
#     gpurun_out/ is capped at 64 MiB).
D=gpurun_out/${1:-r2a}; mkdir -p $D
timeout 300 python bench.py --no-cpu-baseline > $D/bench_default.json 2> $D/bench_default.err
DAWN_CONV3_BSTAGES=7 timeout 300 python bench.py --no-cpu-baseline > $D/bench_bst7.json 2> $D/bench_bst7.err
DAWN_CONV3_BSTAGES=7 timeout 300 python -m pytest tests/test_unet_gpu.py -q -k "golden or submodule or cfg2" > $D/pytest_bst7.log 2>&1
DAWN_CONV3_WSTAT=1 timeout 300 python bench.py --no-cpu-baseline > $D/bench_wstat.json 2> $D/bench_wstat.err
DAWN_CONV3_WSTAT=1 timeout 300 python -m pytest tests/test_unet_gpu.py -q -k "golden or submodule or cfg2" > $D/pytest_wstat.log 2>&1
timeout 200 python -m pytest tests -m staged_gpu -q -s > $D/pytest_staged.log 2>&1
timeout 200 python tools/bench_clip.py --clips 2 > $D/bench_clip.log 2>&1
timeout 200 python tools/bench_clip.py --clips 2 --graph > $D/bench_clip_graph.log 2>&1
for K in tc_conv3_kernel temporal_fused_kernel; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o $D/$K python tools/profile_step.py 1 > $D/$K.out 2>&1
  ncu -i $D/$K.ncu-rep --page raw --csv > $D/${K}_raw.csv 2>/dev/null
  ncu -i $D/$K.ncu-rep --page source --print-source cuda,sass --csv > $D/${K}_source.csv 2>/dev/null
  rm -f $D/$K.ncu-rep
done
python tools/show_bench.py $D/bench_default.json | head -16; python tools/show_bench.py $D/bench_bst7.json | head -16; python tools/show_bench.py $D/bench_wstat.json | head -16; tail -3 $D/pytest_wstat.log
tail -3 $D/pytest_bst7.log; tail -3 $D/pytest_staged.log; tail -2 $D/bench_clip.log; tail -2 $D/bench_clip_graph.log; du -sh gpurun_out
