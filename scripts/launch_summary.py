"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total ms, share, launch count.
   python scripts/launch_summary.py gpurun_out/launches.csv [--md]"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1], errors="replace") if l.startswith('"')]
rd = csv.reader(rows)
hdr = next(rd)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rd:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = re.sub(r"^(void )?(ecamp::)?(\(anonymous namespace\)::|<unnamed>::)?", "", r[ik])
    name = re.sub(r"\(.*$", "", name)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"total {total:.2f} ms over {sum(cnt.values())} launches")
print("| ms | share | launches | kernel |\n|---:|---:|---:|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| {v:.3f} | {100 * v / total:.1f}% | {cnt[k]} | `{k[:110]}` |")
