"""Same workloads through several builds of the library (VB_LIB_PATH is read at import: one subprocess per build)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, statistics, torch
sys.path.insert(0, %r)
from vorta_b200 import ops
def timed(fn, iters=10):
    fn(); torch.cuda.synchronize(); ops.timing_enable(True); ops.timing_collect()
    for _ in range(iters): fn()
    torch.cuda.synchronize(); ops.timing_enable(False)
    ms, n, fl = ops.timing_collect(); return ms / iters, fl / iters
out = []
for name, lat, tile, lw, H, branch in (("dense16k", (1, 1, 16384), (1, 1, 16384), (1, 1, 2), 37, [0] * 37),
                                       ("wan14mix", (21, 45, 80), (3, 9, 16), (3, 3, 2), 40, [0] * 9 + [1] * 18 + [2] * 13)):
    plan = ops.Plan(lat, tile, (3, 3, 3) if lat[0] > 1 else (1, 1, 1), lw, 0.5)
    S = plan.seq_len
    q, k, v = (torch.randn((1, S, H, 128), device="cuda").bfloat16().transpose(1, 2) for _ in range(3))
    rates = []
    for _ in range(3):
        ms, fl = timed(lambda: ops.routed_attention(plan, q, k, v, branch=branch))
        rates.append(fl / ms / 1e9)
    out.append("%%s %%.1f" %% (name, statistics.median(rates)))
print(" | ".join(out))
''' % ROOT


def main():
    exp = os.path.join(ROOT, "vorta_b200", "lib", "exp")
    builds = [("default (1/8 poly)", None)] + [(n, os.path.join(exp, f"libvb_{n}.so")) for n in sys.argv[1:]]
    for rnd in range(2):
        for name, path in builds:
            env = dict(os.environ)
            if path:
                env["VB_LIB_PATH"] = path
            r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=env, timeout=300)
            print(f"round {rnd} {name:22s} {r.stdout.strip() or r.stderr[-300:]}", flush=True)


if __name__ == "__main__":
    main()
