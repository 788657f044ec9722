"""Compact table of the interesting metrics from `ncu -i x.ncu-rep --page raw --csv` output (file argument)."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpc__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hi], rows[hi + 1]
idx = {h: i for i, h in enumerate(hdr)}
extra = sys.argv[2:]
for r in rows[hi + 2:]:
    print("====", r[idx["Kernel Name"]][:80], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
    for k in KEYS + extra:
        if k in idx and r[idx[k]] not in ("", "n/a"):
            print(f"   {k}: {r[idx[k]]} {units[idx[k]]}")
    if extra == ["ALL"]:
        for k in hdr:
            print(f"   {k}: {r[idx[k]]} {units[idx[k]]}")
