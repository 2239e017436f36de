# 4-GPU sharded bench (ranks 1 and 2 are interior shards: 280-frame windows), driver launch line, no clip pipeline line
D=gpurun_out/${1:-scale4}; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29531 bench.py --gpus 4 --steps 8 --warmup 3 --no-clip > $D/bench_4gpu.json 2> $D/bench_4gpu.err; echo "exit $?"
python tools/show_bench.py $D/bench_4gpu.json | grep -E "ms/step|temporal|comm" | head
tail -2 $D/bench_4gpu.err
