# quick profile call: bench line (+breakdown), ncu launch list, ncu --set full of the first launches of the fused kernels (CSV only)
D=gpurun_out/${1:-prof}; mkdir -p $D
timeout 400 python bench.py > $D/bench.json 2> $D/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches.csv python tools/profile_step.py 3 > $D/launches.out 2>&1
gzip -f $D/launches.csv
timeout 300 ncu --set full --clock-control none -k 'regex:temporal_fused|sla_|ca_wt|gn_hcond|tc_conv3|gn_apply' -c 14 -f -o $D/full python tools/profile_step.py 1 > $D/full.out 2>&1
ncu -i $D/full.ncu-rep --page raw --csv > $D/full_raw.csv 2>/dev/null; rm -f $D/full.ncu-rep
cut -c1-300 $D/bench.json; tail -2 $D/launches.out; tail -2 $D/full.out; du -sh gpurun_out
