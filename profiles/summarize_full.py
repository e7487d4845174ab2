#!/usr/bin/env python
"""Key counters of an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv`:
one column per captured launch.  python profiles/summarize_full.py profiles/r1b_attn_full_raw.csv"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "CTAs"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("sm__cycles_active.avg", "SM cycles active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe, % of active cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe, % of elapsed cycles"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU ex2) pipe, % of active"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe, % of active"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe, % of active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots, % of active"),
    ("sm__warps_active.avg.per_cycle_active", "warps active / cycle"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__sass_inst_executed_op_tmem_ldt.sum", "tcgen05.ld instructions"),
    ("smsp__sass_inst_executed_op_tmem_stt.sum", "tcgen05.st instructions"),
    ("sm__inst_executed_pipe_tc.sum", "tcgen05.mma instructions (tc pipe)"),
]


def traffic(path, workload, source):
    """--traffic: average DRAM bytes (read + write) per captured launch -> one JSON record for attn_traffic.json"""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = col[key]
        tot += sum(float(r[i]) for r in data) * scale[units[i]]
    print(json.dumps({workload: {"dram_bytes_per_launch": tot / len(data), "launches_captured": len(data),
                                 "source": source}}))


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]][:40] + " grid " + r[col["Grid Size"]] for r in data] if "Kernel Name" in col else []
    print("| counter | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|" + "---|" * len(data))
    if names:
        print("| kernel | " + " | ".join(names) + " |")
    for key, label in KEYS:
        if key not in col:
            continue
        i = col[key]
        print(f"| {label} (`{key}`, {units[i]}) | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    main()
