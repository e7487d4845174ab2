"""A/B timing of the attention kernel's launch modes in ONE process (run with gpurun): the modes alternate, so both see
the same clocks / power state; medians over the rounds.  VB_ATTN_GRID is read by the library at every launch."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vorta_b200 import ops  # noqa: E402

MODES = [("persistent", {}), ("per-item", {"VB_ATTN_GRID": "items"})]


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    ops.timing_enable(True)
    ops.timing_collect()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    ops.timing_enable(False)
    kms, n, fl = ops.timing_collect()
    return kms / iters, fl / iters


def ab(name, fn, iters=8, rounds=3):
    res = {m: [] for m, _ in MODES}
    flops = 0.0
    for _ in range(rounds):
        for m, env in MODES:
            for k in ("VB_ATTN_GRID",):
                os.environ.pop(k, None)
            os.environ.update(env)
            ms, flops = timed(fn, iters)
            res[m].append(ms)
    os.environ.pop("VB_ATTN_GRID", None)
    out = []
    for m, _ in MODES:
        med = statistics.median(res[m])
        out.append(f"{m} {med:8.3f} ms {flops / med / 1e9:7.1f} TFLOP/s")
    print(f"{name:44s} " + " | ".join(out), flush=True)


def main():
    cases = [("dense S=16384 H=37", (1, 1, 16384), (1, 1, 16384), (1, 1, 1), (1, 1, 2), 37, [0] * 37),
             ("wan14 full H=8", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 8, [0] * 8),
             ("wan14 coreset H=16", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 16, [1] * 16),
             ("wan14 sliding H=16", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 16, [2] * 16),
             ("wan14 layer mix 9f/18c/13s H=40", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 40,
              [0] * 9 + [1] * 18 + [2] * 13),
             ("wan14 native sliding (5,9,8) H=16", (20, 45, 80), (5, 9, 8), (3, 3, 3), (2, 3, 2), 16, [2] * 16),
             ("wan13 sliding (3,10,4) H=12", (21, 30, 52), (3, 10, 4), (3, 3, 3), (3, 3, 2), 12, [2] * 12),
             ("wan13 layer mix 4f/4c/4s H=12", (21, 30, 52), (3, 10, 4), (3, 3, 3), (3, 3, 2), 12,
              [0] * 4 + [1] * 4 + [2] * 4)]
    for name, lat, tile, win, lw, H, branch in cases:
        plan = ops.Plan(lat, tile, win, lw, 0.5)
        S = plan.seq_len
        q, k, v = (torch.randn((1, S, H, 128), device="cuda").bfloat16().transpose(1, 2) for _ in range(3))
        ab(name, lambda: ops.routed_attention(plan, q, k, v, branch=branch))
        del q, k, v, plan
        torch.cuda.empty_cache()
    q = torch.randn((1, 75600, 40, 128), device="cuda").bfloat16().transpose(1, 2)
    k, v = (torch.randn((1, 512, 40, 128), device="cuda").bfloat16().transpose(1, 2) for _ in range(2))
    ab("wan14 cross attention 40 x 75600 x 512", lambda: ops.attn_dense(q, k, v), iters=20)


if __name__ == "__main__":
    main()
