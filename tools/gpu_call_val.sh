# single-GPU validation: full GPU suite, default bench (with cpu baseline + clip), reference arm
D=gpurun_out/${1:-val}; mkdir -p $D
( timeout 900 python -m pytest tests -m gpu -q -s > $D/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $D/pytest_gpu.log )
grep -E "passed|failed|error|exit|FAILED|Error|cfg3|lfg_" $D/pytest_gpu.log | tail -25
timeout 600 python bench.py > $D/bench.json 2> $D/bench.err; echo "bench exit $?"
python tools/show_bench.py $D/bench.json 2>/dev/null | head -20
DD=$D python - <<'PY'
import json,os
d=json.loads(open(os.environ.get("DD","gpurun_out/val")+"/bench.json").read().strip().splitlines()[-1])
for k in ("roofline","cpu_baseline","e2e","clip","cfg1","clocks"): print(k, json.dumps(d.get(k))[:600])
PY
tail -3 $D/bench.err
