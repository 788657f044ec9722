"""Per-kernel-name aggregate of an `ncu --metrics ... --csv` log: time, DRAM bytes, achieved GB/s, pipe activity."""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1], errors="replace") if l.startswith('"')]
rd = csv.reader(rows)
hdr = next(rd)
ik, iid, im, iv, iu = hdr.index("Kernel Name"), hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = defaultdict(dict)
names = {}
for r in rd:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    u = r[iu]
    if r[im] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    if r[im].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per[r[iid]][r[im]] = v
    name = re.sub(r"^(void )?(ecamp::)?(\(anonymous namespace\)::|<unnamed>::)?", "", r[ik])
    names[r[iid]] = re.sub(r"\(.*$", "", name)[:60]
agg = defaultdict(lambda: defaultdict(float))
cnt = defaultdict(int)
for i, m in per.items():
    n = names[i]
    cnt[n] += 1
    t = m.get("gpu__time_duration.sum", 0.0)
    agg[n]["ms"] += t
    agg[n]["bytes"] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    for k, v in m.items():
        if "pct" in k:
            agg[n][k] += v * t   # time-weighted
print("| kernel | n | ms | DRAM MB/launch | GB/s | dram% | tensor% | issue% | fma% | lsu% | warps% | regs |")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    t = a["ms"]
    w = lambda k: a.get(k, 0) / t if t else 0
    i0 = next(i for i in per if names[i] == n)
    print(f"| {n} | {cnt[n]} | {t:.3f} | {a['bytes']/cnt[n]/1e6:.1f} | {a['bytes']/t/1e6:.0f} | "
          f"{w('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {w('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{w('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {w('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{w('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.0f} | {w('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {per[i0].get('launch__registers_per_thread', 0):.0f} |")
