"""Import the reference's own modules (read-only tree at /root/reference) — TEST INFRASTRUCTURE ONLY.

Only used in the build container, by ``oracle/make_golden.py`` and by tests that are skipped when the
reference tree is absent (it does not exist on the GPU box).  Recipe from SURVEY.md appendix C: four stub
``diffusers`` modules make ``vorta.attention`` importable; ``CXX=/usr/bin/g++`` lets Inductor build its CPU
kernels for ``flex_attention``.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VORTA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vorta", "attention"))


def _real_rotary_emb(x, freqs_cis, use_real=True, use_real_unbind_dim=-1):
    """Real-valued RoPE used by the HunyuanVideo processor (reference call site: vorta/attention/hunyuan.py:97-98;
    the function lives in diffusers 0.33.1, which is not installed here — restated from its published
    behaviour for ``use_real=True, use_real_unbind_dim=-1``: pairs (x0, x1) -> (x0 cos - x1 sin, x1 cos + x0 sin))."""
    import torch
    cos, sin = freqs_cis
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


def load():
    """Returns a namespace with the reference's attention / router / ulysses symbols."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("CXX", "/usr/bin/g++")
    for name in ("diffusers", "diffusers.models", "diffusers.models.attention_processor",
                 "diffusers.models.embeddings"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["diffusers.models.attention_processor"].Attention = object   # annotation only (wan.py:18)
    sys.modules["diffusers.models.embeddings"].apply_rotary_emb = _real_rotary_emb
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import vorta.attention as att
    import vorta.attention.coreset_select as cs
    import vorta.attention.sliding_attn_flex as saf
    import vorta.attention.tile as tile
    import vorta.patch.router as router
    import vorta.ulysses as uly
    ns = types.SimpleNamespace(att=att, cs=cs, saf=saf, tile=tile, router=router, uly=uly)
    return ns
