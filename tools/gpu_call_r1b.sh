mkdir -p gpurun_out/r1b
( timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1b/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1b/pytest_gpu.log )
timeout 400 python bench.py > gpurun_out/r1b/bench.json 2> gpurun_out/r1b/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b/launches.csv python tools/profile_step.py 3 > gpurun_out/r1b/launches.out 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k 'regex:tc_conv3_kernel|tc_gemm_kernel|temporal_fused|sla_|ca_gate|gn_hcond|attention_tc' -c 36 -f -o gpurun_out/r1b/full python tools/profile_step.py 1 > gpurun_out/r1b/full.out 2>&1
tail -3 gpurun_out/r1b/pytest_gpu.log; cat gpurun_out/r1b/bench.json | cut -c1-600; tail -2 gpurun_out/r1b/launches.out; tail -2 gpurun_out/r1b/full.out; ls -la gpurun_out/r1b
