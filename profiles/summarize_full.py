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
    """--traffic: DRAM bytes (read + write) of each captured launch -> JSON records for attn_traffic.json; launches
    are labelled "self" (routed self-attention: the larger grid time) and "cross" (dense 512-key launch) by hand in
    the committed file, here they are listed per launch with their grid sizes."""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for r in data:
        tot = sum(float(r[col[key]]) * scale[units[col[key]]] for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.append({"grid": r[col["Grid Size"]] if "Grid Size" in col else None, "dram_bytes_per_launch": tot,
                    "source": source})
    print(json.dumps({workload: out}, indent=1))


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
