"""Compact per-kernel summary of an `ncu --set full` report exported with `--page raw --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size"]
ki = hdr.index("Kernel Name")
seen = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    seen.setdefault(name, []).append(r)
for name, rs in seen.items():
    r = rs[-1]
    print(f"== {name}  (launches captured: {len(rs)}; last shown)")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:75s} {r[i]:>16s} {units[i]}")
