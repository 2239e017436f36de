# 8-GPU sharded bench at the final HEAD (driver launch line, no clip pipeline line: GPU-minutes)
D=gpurun_out/${1:-scale8}; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29531 bench.py --gpus 8 --steps 8 --warmup 3 --no-clip > $D/bench_8gpu.json 2> $D/bench_8gpu.err; echo "exit $?"
python tools/show_bench.py $D/bench_8gpu.json | grep -E "ms/step|temporal|comm" | head
tail -2 $D/bench_8gpu.err
