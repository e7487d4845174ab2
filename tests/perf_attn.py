"""Attention-kernel micro-benchmark (run with gpurun): forced-branch sweeps, TFLOP/s from CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vorta_b200 import ops  # noqa: E402


def bench(plan, q, k, v, branch, iters=5):
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
    kms, n, fl = ops.timing_collect()
    return e0.elapsed_time(e1) / iters, kms / iters, fl / iters


def main():
    tag = os.environ.get("VB_TAG", "default")
    cases = [("dense S=16384 H=37 (16 full waves)", (1, 1, 16384), (1, 1, 16384), (1, 1, 1), (1, 1, 2), 37, 0),
             ("dense S=32768 H=37", (1, 1, 32768), (1, 1, 32768), (1, 1, 1), (1, 1, 2), 37, 0)]
    if os.environ.get("VB_QUICK2"):
        cases = [("wan14 native sliding (5,9,8) H=16", (20, 45, 80), (5, 9, 8), (3, 3, 3), (2, 3, 2), 16, 2),
                 ("wan13 grid sliding (3,10,4) H=12", (21, 30, 52), (3, 10, 4), (3, 3, 3), (3, 3, 2), 12, 2)]
    elif os.environ.get("VB_SLIDING"):
        cases = [("wan14 grid sliding H=16", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 16, 2),
                 ("wan14 native sliding (5,9,8) H=16", (20, 45, 80), (5, 9, 8), (3, 3, 3), (2, 3, 2), 16, 2),
                 ("wan13 grid sliding (3,10,4) H=12", (21, 30, 52), (3, 10, 4), (3, 3, 3), (3, 3, 2), 12, 2),
                 ("hunyuan 129f grid sliding (3,9,16) H=12", (33, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 12, 2)]
    elif not os.environ.get("VB_QUICK"):
        cases += [("wan14 grid full H=8", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 8, 0),
                  ("wan14 grid coreset H=16", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 16, 1),
                  ("wan14 grid sliding H=16", (21, 45, 80), (3, 9, 16), (3, 3, 3), (3, 3, 2), 16, 2),
                  ("wan14 native sliding (5,9,8) H=16", (20, 45, 80), (5, 9, 8), (3, 3, 3), (2, 3, 2), 16, 2),
                  ("wan13 grid sliding (3,10,4) H=12", (21, 30, 52), (3, 10, 4), (3, 3, 3), (3, 3, 2), 12, 2),
                  ("wan13 grid coreset H=12", (21, 30, 52), (3, 10, 4), (3, 3, 3), (3, 3, 2), 12, 1),
                  ("wan13 grid full H=12", (21, 30, 52), (3, 10, 4), (3, 3, 3), (3, 3, 2), 12, 0)]
    for name, lat, tile, win, lw, H, e in cases:
        plan = ops.Plan(lat, tile, win, lw, 0.5)
        S = plan.seq_len
        q, k, v = (torch.randn((1, S, H, 128), device="cuda").bfloat16().transpose(1, 2) for _ in range(3))
        ms, kms, fl = bench(plan, q, k, v, [e] * H)
        print(f"[{tag}] {name:42s} total {ms:8.3f} ms  attn-kernel {kms:8.3f} ms  {fl / kms / 1e9:8.1f} TFLOP/s (kernel)"
              f"  {fl / ms / 1e9:8.1f} (incl. select/gather)", flush=True)
        del q, k, v, plan
        torch.cuda.empty_cache()


def cross():
    """Wan-14B cross attention: 40 heads x 75,600 queries x 512 text keys (vb_attn_dense)."""
    tag = os.environ.get("VB_TAG", "default")
    q = torch.randn((1, 75600, 40, 128), device="cuda").bfloat16().transpose(1, 2)
    k, v = (torch.randn((1, 512, 40, 128), device="cuda").bfloat16().transpose(1, 2) for _ in range(2))
    for _ in range(3):
        ops.attn_dense(q, k, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        ops.attn_dense(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * 75600 * 512 * 128 * 40
    gb = (2 * 75600 * 40 * 128 * 2 + 2 * 512 * 40 * 128 * 2) / 1e9
    print(f"[{tag}] wan14 cross attention 40 heads x 75600 x 512   {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  "
          f"{gb / ms * 1e3:8.1f} GB/s of algorithmic Q+O+K+V bytes", flush=True)


if __name__ == "__main__":
    cross()
    main()
