"""Print the headline numbers and the per-category breakdown of one bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"{d['ms_per_step']:.2f} ms/step  {d['value']:.2f} {d['unit']}  launches {d.get('gpu_launches')}  e2e {d.get('e2e', {}).get('value')}")
for k, v in d.get("breakdown", {}).items():
    print(f"  {k:12s} {v['ms_per_step']:7.2f} ms  {v['launches_per_step']:3d} launches  alg {v['alg_tflops']} TF/s  {v['alg_gbs']:.0f} GB/s")
