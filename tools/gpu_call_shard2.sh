# 2-GPU call: sharding pytest (torchrun inside), LFG tests with the per-tap drain, sharded bench with the peer-memory all-reduce on / off
D=gpurun_out/${1:-shard2}; mkdir -p $D
( timeout 900 python -m pytest tests/test_shard_gpu.py tests/test_lfg_gpu.py -m gpu -q -s > $D/pytest.log 2>&1; echo "pytest exit $?" >> $D/pytest.log )
grep -E "passed|failed|exit|x tol|\[band\]|\[ddim\]|\[shardbig\]|Error" $D/pytest.log | tail -30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 2 --steps 8 --warmup 3 > $D/bench_2gpu.json 2> $D/bench_2gpu.err; echo "exit $?" >> $D/bench_2gpu.err
DAWN_P2P=0 timeout 300 $TR --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 > $D/bench_2gpu_nccl.json 2> $D/bench_2gpu_nccl.err; echo "exit $?" >> $D/bench_2gpu_nccl.err
DD=$D python - <<'PY'
import json,sys
for f in ("bench_2gpu.json","bench_2gpu_nccl.json"):
    try:
        d=json.loads(open(f"%s/%s" % ("'$D'" if False else __import__("os").environ.get("DD","gpurun_out/shard2"), f)).read().strip().splitlines()[-1])
        print(f, "ms/step %.2f value %.2f" % (d["ms_per_step"], d["value"]), d.get("comm"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -2 $D/bench_2gpu.err
timeout 200 python tools/lfg_diag.py > $D/lfg_diag.log 2>&1; grep -E "bottleneck|up0|up1|prediction|bench" $D/lfg_diag.log
