# N-GPU call (N = $2, default 4): sharded forward parity with interior ranks (both halos), sharded sampler, sharded bench
D=gpurun_out/${1:-shard4}; N=${2:-4}; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29521 tools/shard_test.py forward ddim > $D/shard_n$N.log 2>&1; echo "exit $?" >> $D/shard_n$N.log
timeout 150 $TR --master-port 29523 bench.py --gpus $N --steps 5 --warmup 3 > $D/bench_n$N.json 2> $D/bench_n$N.err; echo "exit $?" >> $D/bench_n$N.err
grep -hE "^\[|exit|Error|error" $D/shard_n$N.log | tail -12; cut -c1-300 $D/bench_n$N.json; tail -2 $D/bench_n$N.err
