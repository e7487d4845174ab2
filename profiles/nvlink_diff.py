#!/usr/bin/env python
"""Difference of two `nvidia-smi nvlink -gt d` snapshots: KiB sent / received per GPU over all links.
    python profiles/nvlink_diff.py before.txt after.txt steps"""
import re
import sys


def parse(path):
    gpu, out = None, {}
    for line in open(path):
        m = re.match(r"GPU (\d+):", line)
        if m:
            gpu = int(m.group(1))
            out[gpu] = [0, 0]
            continue
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None:
            out[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
    return out


def main():
    a, b = parse(sys.argv[1]), parse(sys.argv[2])
    steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    print("gpu,tx_GB_total,rx_GB_total,tx_MB_per_forward,rx_MB_per_forward   (forwards in the window: %g)" % steps)
    for g in sorted(b):
        tx, rx = (b[g][0] - a.get(g, [0, 0])[0]) * 1024.0, (b[g][1] - a.get(g, [0, 0])[1]) * 1024.0
        print(f"{g},{tx / 1e9:.3f},{rx / 1e9:.3f},{tx / 1e6 / steps:.1f},{rx / 1e6 / steps:.1f}")


if __name__ == "__main__":
    main()
