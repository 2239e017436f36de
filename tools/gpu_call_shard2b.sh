# 2-GPU call at the final HEAD: sharding pytest (torchrun inside) and the sharded bench (peer-memory GroupNorm all-reduce)
D=gpurun_out/${1:-shard2b}; mkdir -p $D
( timeout 600 python -m pytest tests/test_shard_gpu.py -m gpu -q -s > $D/pytest.log 2>&1; echo "pytest exit $?" >> $D/pytest.log )
grep -E "passed|failed|exit|x tol|\[band\]|\[ddim\]|\[shardbig\]|Error" $D/pytest.log | tail -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus 2 --steps 8 --warmup 3 > $D/bench_2gpu.json 2> $D/bench_2gpu.err; echo "exit $?"
python tools/show_bench.py $D/bench_2gpu.json | grep -E "ms/step|temporal|comm|clip" | head
