#!/usr/bin/env python
"""Summarise an ncu launch list (--csv, metrics gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum])
into one row per kernel: launches, total time, share of the step, DRAM bytes and achieved DRAM GB/s.

    python profiles/summarize_launches.py gpurun_out/launches_r1b.csv "title line" > profiles/r1b_launches_wan13_step.csv

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)                 # drop the argument list
    name = re.sub(r"<.*", "", name)                  # and template arguments
    name = name.replace("void ", "").strip()
    return name or "void"


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [l for l in open(path, newline="") if l.startswith('"')]
    rd = csv.DictReader(rows)
    per_id = defaultdict(dict)
    names = {}
    for r in rd:
        v = r["Metric Value"].replace(",", "")
        try:
            val = float(v)
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per_id[r["ID"]][r["Metric Name"]] = val * scale
        names[r["ID"]] = short(r["Kernel Name"])
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per_id.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    print(f"# {title}")
    print(f"# total kernel time {total / 1e6:.2f} ms over {sum(a[0] for a in agg.values())} launches "
          f"(cold-cache, serialised under ncu: compare SHARES)")
    print("share_pct,total_ms,launches,avg_us,dram_read_MB,dram_write_MB,dram_GBps,kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbps = (a[2] + a[3]) / a[1] if a[1] > 0 else 0.0          # bytes / ns == GB/s
        print(f"{100 * a[1] / total:.2f},{a[1] / 1e6:.3f},{a[0]},{a[1] / a[0] / 1e3:.1f},{a[2] / 1e6:.1f},{a[3] / 1e6:.1f},"
              f"{gbps:.0f},{k}")


if __name__ == "__main__":
    main()
