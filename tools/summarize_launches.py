"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown)."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.search(r"gemm_kernel<\(?(?:int\))?(\d+)>", name) or re.search(r"gemm_kernel<(\d+)>", name)
    if "tc_gemm_kernel" in name:
        m2 = re.search(r"tc_gemm_kernel<[^0-9]*(\d+)[^0-9]+(\d+)", name)
        return f"dawn::tc_gemm_kernel<EPI={m2.group(1)},BN={m2.group(2)}>" if m2 else "dawn::tc_gemm_kernel"
    if "gemm_kernel" in name:
        epi = re.search(r"gemm_kernel<[^0-9]*(\d+)", name)
        return f"dawn::gemm_kernel<EPI={epi.group(1) if epi else '?'}> (mma.sync)"
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("dawn::(anonymous namespace)::", "dawn::")
    return name[:90]


def main(path, last_n=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if last_n:
        rows = rows[-last_n:]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")
    print(f"| **total** | {len(rows)} | {tot / 1e6:.3f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
