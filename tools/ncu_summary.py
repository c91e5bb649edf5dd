"""Compact per-kernel summary of an `ncu --set full` report exported with `--page raw --csv`.

  python tools/ncu_summary.py raw.csv                               -> text table (what profiles/*_summary.txt hold)
  python tools/ncu_summary.py raw.csv --json profiles/ncu_metrics.json --tag "M=48000 N=5120 K=1280" --source <file>
        merges one entry per captured kernel into the JSON that bench.py reads for `roofline.traffic` and the tensor-pipe
        percentages it quotes (so the bench line cannot go stale against the profiles: the entry carries the kernel name,
        the shape tag, the git SHA the library was built from and the summary file it came from)."""
import argparse
import csv
import json
import os
import re
import subprocess

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size"]
_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short_name(full: str) -> str:
    n = full.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    n = re.sub(r"\((int|bool)\)", "", full.replace("void ", "").replace("<unnamed>::", "")).split("(")[0]
    return n.strip()


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        seen.setdefault(short_name(r[ki]), []).append(r)
    return hdr, units, seen


def value(hdr, units, row, key):
    if key not in hdr:
        return None
    i = hdr.index(key)
    try:
        v = float(row[i].replace(",", ""))
    except ValueError:
        return None
    return v * _SCALE.get(units[i], 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json", default=None)
    ap.add_argument("--tag", default="")
    ap.add_argument("--source", default=None, help="the summary file under profiles/ this entry is quoted from")
    ap.add_argument("--sha", default=None)
    a = ap.parse_args()
    hdr, units, seen = load(a.csv)
    for name, rs in seen.items():
        r = rs[-1]
        print(f"== {name}  (launches captured: {len(rs)}; last shown)")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:75s} {r[i]:>16s} {units[i]}")
    if a.json:
        sha = a.sha
        if sha is None:
            try:
                sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
            except Exception:
                sha = "unknown"
        db = json.load(open(a.json)) if os.path.exists(a.json) else {}
        for name, rs in seen.items():
            r = rs[-1]
            rd, wr = value(hdr, units, r, "dram__bytes_read.sum"), value(hdr, units, r, "dram__bytes_write.sum")
            db[f"{name} | {a.tag}".strip(" |")] = {
                "kernel": name, "shape": a.tag, "git_sha": sha, "source": a.source,
                "duration_us": value(hdr, units, r, "gpu__time_duration.sum"),
                "dram_bytes": (rd + wr) if rd is not None and wr is not None else None,
                "tensor_pipe_active_pct": value(hdr, units, r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "xu_pipe_pct": value(hdr, units, r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": value(hdr, units, r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "registers_per_thread": value(hdr, units, r, "launch__registers_per_thread")}
        with open(a.json, "w") as f:
            json.dump(db, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
