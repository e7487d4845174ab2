"""BASELINE configs[0] as a callable: ONE Wan2.1-T2V-1.3B routed self-attention layer (12 heads x 128, 21x30x52 =
32,760 tokens) driven through the REFERENCE's own ``WanAttnProcessorTripleEval`` on the host CPU — TEST / BENCH
INFRASTRUCTURE ONLY (used by ``bench.py``'s CPU legs and by tests; never by the product path).

The reference code is imported by ``oracle/ref_loader.py`` (from /root/reference, or from the byte-identical staged
copy under ``oracle/_ref``).  The case is storage-free: ``oracle/fixtures.py`` hashes produce the same bf16-exact
weights, activations, RoPE phases and routing scores on every machine, so the GPU arm (``vorta_b200``'s processor of
the same name, bf16) and the CPU arm (reference, fp32 up-cast of the same values) see identical
inputs and their outputs can be compared.
"""
from __future__ import annotations

import os
import time

import torch

from . import fixtures as FX

# BASELINE.json configs[0]; 21 frames do not divide the reference's (2,3,2) coreset window / (5,6,4) tile, so the
# substitutes of SURVEY.md section 8d are used: coreset window (3,3,2) (g = 18, keep 9), tile (3,10,4) (120 tokens)
CONFIG0 = dict(heads=12, latent=(21, 30, 52), tile=(3, 10, 4), window=(3, 3, 3), lowres_window=(3, 3, 2), rate=0.5,
               tau_sparse=0.3,
               name="configs[0]: one Wan2.1-T2V-1.3B routed self-attention layer (12 heads x 128, 21x30x52 = 32,760 "
                    "tokens; router scores -> top-1 -> q/k/v projections + RMSNorm + RoPE -> full / coreset / "
                    "sliding-tile attention -> output projection)")
# tests/test_bench_contract.py only (bench.py --workload tiny): same code path on a 192-token, 3-head layer
TINY = dict(heads=3, latent=(4, 6, 8), tile=(2, 3, 4), window=(3, 3, 3), lowres_window=(2, 3, 2), rate=0.5,
            tau_sparse=0.3, name="CONTRACT TEST ONLY: one 3-head routed self-attention layer, 4x6x8 = 192 tokens")


# fixed per-head routing of the like-for-like layer: heads 0-3 full, 4-7 coreset, 8-11 sliding tile (configs[0])
def branch_of_head(heads: int):
    """Thirds of the heads in head order: full, coreset, sliding tile."""
    return [min(2, 3 * h // heads) for h in range(heads)]


def routing_score(heads: int = 12) -> torch.Tensor:
    """(1, H, 3) scores whose top-1 (above tau = 0.3) is branch_of_head; bf16-exact values."""
    s = torch.full((1, heads, 3), 0.125)
    for h, e in enumerate(branch_of_head(heads)):
        s[0, h, e] = 0.75
    return s


def build_case(cfg=CONFIG0, seed: int = 71):
    """Module + inputs of the layer, fp32 tensors holding bf16-representable values."""
    lat, H = cfg["latent"], cfg["heads"]
    S = lat[0] * lat[1] * lat[2]
    attn = FX.FakeWanAttn(H, seed=seed)
    hs = FX.det_tensor((1, S, H * FX.D), seed + 1, 1.0)
    rot = FX.wan_rotary(S, seed + 2)
    return dict(cfg=cfg, S=S, attn=attn, hidden_states=hs, rotary_emb=rot, routing_score=routing_score(H))


def algorithmic_flops(cfg=CONFIG0) -> dict:
    """BASELINE.md section 3 formulas for the layer's routing (attention) + the four H*D x H*D projections."""
    lat, H, D = cfg["latent"], cfg["heads"], FX.D
    S = lat[0] * lat[1] * lat[2]
    g = cfg["lowres_window"][0] * cfg["lowres_window"][1] * cfg["lowres_window"][2]
    n_u = int(g * (1 - cfg["rate"])) - 1
    S_c = (S // g) * (1 + n_u)
    k_w = 1
    for n, t, w in zip(lat, cfg["tile"], cfg["window"]):
        k_w *= min(n // t, 2 * (w // 2) + 1)
    k_w *= cfg["tile"][0] * cfg["tile"][1] * cfg["tile"][2]
    per = [4.0 * S * S * D, 4.0 * S_c * S_c * D, 4.0 * D * S * k_w]
    attn = sum(per[e] for e in branch_of_head(H))
    proj = 4 * 2.0 * S * (H * D) ** 2
    return dict(attention=attn, projections=proj, total=attn + proj)


class ReferenceLayer:
    """The reference's Eval processor on the case (vorta/attention/wan.py:303-438 driving coreset_select.py and
    sliding_attn_flex.py:137-211).  ``prepare()`` builds the group tables and the BlockMask (Inductor compiles
    ``create_block_mask`` and ``flex_attention`` for the CPU on first use: minutes, untimed); ``__call__`` is one layer."""

    def __init__(self, case, threads: int | None = None):
        from . import ref_loader
        self.threads = int(threads or os.cpu_count() or 1)
        torch.set_num_threads(self.threads)       # torchrun exports OMP_NUM_THREADS=1: pin explicitly
        self.ref = ref_loader.load()
        self.source = ref_loader.kind()
        self.case = case
        self.kw = None

    def prepare(self):
        c = self.case["cfg"]
        t0 = time.perf_counter()
        info = self.ref.cs.get_group_info(c["latent"], c["lowres_window"], reduction_rate=c["rate"])
        bm = self.ref.saf.create_sliding_tile_attn_mask_func(c["latent"], c["window"], c["tile"], 0, 0,
                                                             torch.device("cpu"))
        self.kw = dict(lowres_group_info=info, flex_attn_mask_func=bm, window_size=c["window"], tile_size=c["tile"],
                       latent_shape=c["latent"])
        self.proc = self.ref.att.WanAttnProcessorTripleEval(check_input=True)
        return time.perf_counter() - t0

    @torch.no_grad()
    def __call__(self):
        x = self.case
        return self.proc(x["attn"], x["hidden_states"], None, None, x["rotary_emb"],
                         tau_sparse=x["cfg"]["tau_sparse"], routing_score=x["routing_score"], **self.kw)
