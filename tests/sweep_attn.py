"""BASELINE.json configs[4]: attention-only sweep 8k-128k tokens with routing forced to the full / sliding-tile /
coreset branches, against the measured bf16 roofline.  Run with gpurun; writes gpurun_out/sweep_attn.csv.
Kernel time = CUDA events around the tcgen05 attention launches (vb_timing_*); "total" also includes the
selection / gather passes of the sparse branches.  TFLOP/s are ALGORITHMIC: padding is not credited.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vorta_b200 import ops  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRIDS = [("8k", (8, 32, 32)), ("16k", (16, 32, 32)), ("32k", (32, 32, 32)), ("64k", (32, 32, 64)), ("128k", (32, 64, 64))]
TILE, WINDOW, LOWRES = (4, 8, 8), (3, 3, 3), (2, 2, 2)      # 256-token tiles, 27-tile window, 8-token groups (keep 4)


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d["bf16_tflops_sustained"])
    return 1590.0, 1400.0


def run(plan, q, k, v, branch, iters):
    for _ in range(2):
        ops.routed_attention(plan, q, k, v, branch=branch)
    torch.cuda.synchronize()
    ops.timing_enable(True)
    ops.timing_collect()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        ops.routed_attention(plan, q, k, v, branch=branch)
    e1.record()
    torch.cuda.synchronize()
    ops.timing_enable(False)
    kms, _, fl = ops.timing_collect()
    return e0.elapsed_time(e1) / iters, kms / iters, fl / iters


def main():
    burst, sustained = peak()
    rows = ["tokens,grid,branch,heads,density_pct,total_ms,kernel_ms,kernel_tflops,frac_of_burst_peak,frac_of_spec_2250,total_tflops"]
    for name, lat in GRIDS:
        S = lat[0] * lat[1] * lat[2]
        H = 8 if S >= 65536 else 16
        plan = ops.Plan(lat, TILE, WINDOW, LOWRES, 0.5)
        q, k, v = (torch.randn((1, S, H, 128), device="cuda").bfloat16().transpose(1, 2) for _ in range(3))
        for e, bname in ((0, "full"), (2, "sliding"), (1, "coreset")):
            iters = 3 if S >= 65536 and e == 0 else 5
            ms, kms, fl = run(plan, q, k, v, [e] * H, iters)
            dens = 100.0 * plan.flops_per_head(e) / plan.flops_per_head(0)
            tf = fl / kms / 1e9
            rows.append(f"{S},{lat[0]}x{lat[1]}x{lat[2]},{bname},{H},{dens:.1f},{ms:.3f},{kms:.3f},{tf:.1f},"
                        f"{tf / burst:.3f},{tf / 2250:.3f},{fl / ms / 1e9:.1f}")
            print(rows[-1], flush=True)
        del q, k, v, plan
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep_attn.csv"), "w") as f:
        f.write(f"# attention-only sweep, tile {TILE}, window {WINDOW}, coreset window {LOWRES} r=0.5; "
                f"measured peaks: burst {burst} / sustained {sustained} TFLOP/s\n")
        f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
