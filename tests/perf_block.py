"""HBM-bound kernels of the path against the measured copy bandwidth (run with gpurun): algorithmic bytes / CUDA-event
time for LayerNorm-modulate, gated residual, RMSNorm+RoPE, row gather (tile layout / coreset pooling) and coreset
selection at the Wan-14B 720p and Wan-1.3B 480p sizes.  Writes one CSV line per kernel to stdout.

Inputs are rotated over enough distinct buffers to exceed the 126 MB L2 between iterations."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vorta_b200 import ops  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fns, iters=20):
    """fns: list of closures over distinct buffers, called round-robin."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peak = peak_hbm()
    print(f"# HBM-bound kernels, algorithmic bytes / CUDA-event time; peak = {peak:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs, copy)")
    print("kernel,config,algorithmic_MB,ms,GBps,frac_of_measured_copy_peak")
    dev = "cuda"
    for name, S, dim, heads in (("wan14", 75600, 5120, 40), ("wan13", 32760, 1536, 12), ("hunyuan", 118800, 3072, 24)):
        row_bytes = S * dim * 2
        nbuf = max(2, int(300e6 // row_bytes) + 1)
        xs = [torch.randn((1, S, dim), device=dev).bfloat16() for _ in range(nbuf)]
        ys = [torch.randn((1, S, dim), device=dev).bfloat16() for _ in range(nbuf)]
        scale, shift, gate = (torch.randn((1, dim), device=dev) * 0.1 for _ in range(3))
        w16 = torch.ones(dim, device=dev).bfloat16()
        cos, sin = torch.rand((S, 64), device=dev), torch.rand((S, 64), device=dev)

        def line(kernel, mb, ms):
            gbps = mb / ms
            print(f"{kernel},{name} S={S} dim={dim},{mb:.1f},{ms:.4f},{gbps:.0f},{gbps / peak:.3f}", flush=True)

        ms = timed([lambda x=x: ops.ln_modulate(x, None, None, scale, shift, 1e-6) for x in xs])
        line("vb_ln_modulate", 2 * row_bytes / 1e6, ms)
        ms = timed([lambda x=x, y=y: ops.gate_residual(x, y, gate) for x, y in zip(xs, ys)])
        line("vb_gate_residual", 3 * row_bytes / 1e6, ms)
        ms = timed([lambda x=x: ops.rmsnorm_rope(x, w16, 1e-6, cos, sin) for x in xs])
        line("vb_rmsnorm_rope", (2 * row_bytes + S * 64 * 8) / 1e6, ms)
        del ys
        # row gather: tile-major permutation of all heads of one (S, H, 128) tensor = one read + one write
        perm = torch.randperm(S, device=dev, dtype=torch.int32)
        qs = [x.view(1, S, heads, 128).transpose(1, 2) for x in xs]
        ms = timed([lambda q=q: ops.gather_rows(q, perm) for q in qs])
        line("vb_gather_rows", 2 * row_bytes / 1e6, ms)
        # coreset selection over every head: reads Q once, writes (g-1)/g * S int32 positions per head
        lat = {"wan14": (21, 45, 80), "wan13": (21, 30, 52), "hunyuan": (33, 45, 80)}[name]
        plan = ops.Plan(lat, (3, 9, 16) if name != "wan13" else (3, 10, 4), (3, 3, 3), (3, 3, 2), 0.5)
        ms = timed([lambda q=q: ops.coreset_select(plan, q) for q in qs], iters=10)
        g = 18
        # int64 matching tables over the g-1 margins of every group + int32 kept / dropped token tables (~ S entries)
        line("vb_coreset_select", (row_bytes + heads * S * ((g - 1) / g * 8 + 4)) / 1e6, ms)
        if name == "hunyuan":      # per-head RMSNorm + RoPE + placement in the joint [video | 256 text] sequence
            w128 = torch.ones(128, device=dev).bfloat16()
            joint = torch.empty((1, S + 256, dim), device=dev, dtype=torch.bfloat16)
            ms = timed([lambda x=x: ops.headnorm_rope(x, w128, 1e-6, heads, cos, sin, out=joint) for x in xs])
            line("vb_block_headnorm_rope", (2 * row_bytes + S * 64 * 8) / 1e6, ms)
        del xs, qs, plan
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
