#!/usr/bin/env python
"""SASS opcode histogram of one kernel of libvorta_b200.so (cuobjdump -sass), for profiles/:
    python profiles/sass_histogram.py vb_attn_fwd_kernel > profiles/r2_sass_histogram_vb_attn_fwd.txt
The tcgen05 / TMA / TMEM mnemonics that prove the Blackwell-native path (B200_PROFILING.md) are listed first."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vorta_b200", "lib", "libvorta_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCATOM", "SYNCS", "MUFU",
       "FFMA2", "FADD2", "HMMA", "HGMMA", "ATOMG", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "ELECT")


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "vb_attn_fwd_kernel"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, inside, name = collections.Counter(), False, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            inside = want in m.group(1)
            name = m.group(1) if inside else name
            continue
        if not inside:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m:
            counts[m.group(1)] += 1
    total = sum(counts.values())
    print(f"# {name}: {total} SASS instructions (cuobjdump -sass vorta_b200/lib/libvorta_b200.so)")
    print("# Blackwell-native mnemonics (full opcode with modifiers, count)")
    for key in KEY:
        rows = sorted(((op, n) for op, n in counts.items() if op.split(".")[0].startswith(key)), key=lambda x: -x[1])
        if rows:
            print(f"{key:8s} {sum(n for _, n in rows):6d}   " + ", ".join(f"{op} x{n}" for op, n in rows[:8]))
    print("# by base opcode")
    base = collections.Counter()
    for op, n in counts.items():
        base[op.split(".")[0]] += n
    for op, n in base.most_common():
        print(f"{op:12s} {n:6d}  {100.0 * n / total:5.1f} %")


if __name__ == "__main__":
    main()
