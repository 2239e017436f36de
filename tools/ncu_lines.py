"""Aggregate ncu warp-stall samples per CUDA source line: python tools/ncu_lines.py report.ncu-rep|export.csv source.cu [top]
(the .csv form is `ncu -i report.ncu-rep --page source --print-source cuda,sass --csv` exported on the GPU box)"""
import collections
import re
import subprocess
import sys

rep, srcf = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
if rep.endswith(".csv"):
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
src = open(srcf).read().splitlines()
agg, ins = collections.Counter(), collections.Counter()
sass = re.compile(r'^"","","(0x[0-9a-f]+)","([^"]*)","(\d+)","(\d+)","(\d+)","(\d+)"')
cur = 0
for l in out.splitlines():
    m = sass.match(l)
    if m:
        agg[cur] += int(m.group(3)); ins[cur] += int(m.group(6))
        continue
    m2 = re.match(r'^"(\d+)",', l)
    if m2:
        cur = int(m2.group(1))
tot, toti = sum(agg.values()) or 1, sum(ins.values()) or 1
print("samples", tot, "warp instructions", toti)
for ln, a in agg.most_common(top):
    text = src[ln - 1].strip()[:100] if 0 < ln <= len(src) else "?"
    print(f"{100 * a / tot:5.1f}% smp {100 * ins[ln] / toti:5.1f}% ins  L{ln}: {text}")
