# 2-GPU call: sharded forward parity, sharded sampler (eager), sharded bench, then the graph-captured sharded sampler (NCCL inside a graph)
D=gpurun_out/${1:-shard}; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/shard_test.py forward > $D/shard_forward.log 2>&1; echo "exit $?" >> $D/shard_forward.log
timeout 150 $TR --master-port 29512 tools/shard_test.py ddim > $D/shard_ddim.log 2>&1; echo "exit $?" >> $D/shard_ddim.log
timeout 200 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > $D/bench_2gpu.json 2> $D/bench_2gpu.err; echo "exit $?" >> $D/bench_2gpu.err
timeout 100 $TR --master-port 29514 tools/shard_test.py ddim_graph > $D/shard_ddim_graph.log 2>&1; echo "exit $?" >> $D/shard_ddim_graph.log
grep -hE "^\[|exit|Error|error" $D/shard_forward.log $D/shard_ddim.log $D/shard_ddim_graph.log | tail -20; cut -c1-400 $D/bench_2gpu.json; tail -2 $D/bench_2gpu.err
