# N-GPU sharded bench (driver launch line), optionally also with the NCCL all-reduce for comparison:  gpu_call_scale.sh <tag> <N> [nccl]
D=gpurun_out/${1:-scale}; N=${2:-2}; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 > $D/bench_${N}gpu.json 2> $D/bench_${N}gpu.err; echo "exit $?" >> $D/bench_${N}gpu.err
if [ "$3" = "nccl" ]; then
  DAWN_P2P=0 timeout 400 $TR --master-port 29532 bench.py --gpus $N --steps 8 --warmup 3 --no-clip > $D/bench_${N}gpu_nccl.json 2> $D/bench_${N}gpu_nccl.err; echo "exit $?" >> $D/bench_${N}gpu_nccl.err
fi
DD=$D NN=$N python - <<'PY'
import json,os
D,N=os.environ["DD"],os.environ["NN"]
for f in (f"bench_{N}gpu.json", f"bench_{N}gpu_nccl.json"):
    try:
        d=json.loads(open(f"{D}/{f}").read().strip().splitlines()[-1])
        print(f, "ms/step %.2f value %.2f e2e %.2f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), "comm", d.get("comm"))
        print("   clip", json.dumps(d.get("clip"))[:700])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $D/bench_${N}gpu.err
