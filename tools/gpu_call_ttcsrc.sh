# one ncu --set full capture (with SASS-level sampling) of a level-0 temporal_tc launch inside a real step; the report travels back
D=gpurun_out/${1:-ttcsrc}; mkdir -p $D
timeout 500 ncu --set full --section SourceCounters --clock-control none --import-source on -k 'regex:temporal_tc_kernel' -s 1 -c 1 -f -o $D/ttc python tools/profile_step.py 1 > $D/full.out 2>&1
tail -2 $D/full.out; du -sh $D
