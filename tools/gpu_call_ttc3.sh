D=gpurun_out/${1:-ttc3}; mkdir -p $D
timeout 900 python tools/ttc_selftest.py > $D/selftest.log 2>&1
grep -v "^   " $D/selftest.log
if grep -q "TIMEOUT\|NO OUTPUT\|rc=-" $D/selftest.log; then echo "selftest failed"; exit 1; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --no-clip > $D/bench_2gpu.json 2> $D/bench_2gpu.err; echo "exit $?"
python tools/show_bench.py $D/bench_2gpu.json | grep -E "ms/step|temporal|comm"
timeout 300 python -m pytest tests/test_shard_gpu.py -m gpu -q -k forward 2>&1 | tail -2
