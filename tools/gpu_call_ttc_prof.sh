# timing + cycle trace + ncu of the tcgen05 temporal kernel at the bench shape of one level-0 layer
D=gpurun_out/${1:-ttcp}; mkdir -p $D
timeout 200 python tools/ttc_selftest.py 200 4096 40 0 200 > $D/trace.log 2>&1
timeout 200 python tools/ttc_selftest.py 280 4096 40 40 240 >> $D/trace.log 2>&1
timeout 200 python tools/ttc_selftest.py 180 4096 40 40 140 >> $D/trace.log 2>&1
cat $D/trace.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:temporal_tc_kernel -s 1 -c 1 -f -o $D/ttc python tools/ttc_selftest.py 200 4096 40 0 200 > $D/ncu.out 2>&1
ncu -i $D/ttc.ncu-rep --page raw --csv > $D/ttc_raw.csv 2>/dev/null
ncu -i $D/ttc.ncu-rep --page source --print-source cuda,sass --csv > $D/ttc_source.csv 2>/dev/null
rm -f $D/ttc.ncu-rep; tail -3 $D/ncu.out; du -sh $D
