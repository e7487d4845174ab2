"""Import the reference's own modules — TEST / BENCH INFRASTRUCTURE ONLY.

Source of the modules, in this order: the read-only tree at /root/reference (build container), else the
byte-identical files ``oracle/stage_ref.py`` staged under the git-ignored ``oracle/_ref/`` (they travel to the GPU
box with the snapshot like the built ``.so``; /root/reference does not exist there).  Used by
``oracle/make_golden.py``, by tests that are skipped when neither exists, and by ``bench.py``'s CPU legs
(``--impl reference`` and ``cpu_baseline``).  Recipe from SURVEY.md appendix C: four stub ``diffusers`` modules make
``vorta.attention`` importable; ``CXX=/usr/bin/g++`` lets Inductor build its CPU kernels for ``flex_attention``.
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref")


def _pick_root() -> str:
    root = os.environ.get("VORTA_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(root, "vorta", "attention")):
        return root
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vorta", "attention"))


def kind() -> str:
    """'tree' = /root/reference itself, 'staged' = oracle/_ref (manifest-verified copy of the same files)."""
    return "staged" if REFERENCE_ROOT == STAGED_ROOT else "tree"


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


def _pick_cxx() -> None:
    """Inductor builds its CPU kernels (create_block_mask, flex_attention) with $CXX -fopenmp.  Some images export a
    CXX whose wrapper cannot find libgomp.spec: take the first candidate that links OpenMP."""
    import shutil
    import subprocess
    cands = [os.environ.get("CXX"), "/usr/bin/g++", shutil.which("g++")]
    for cxx in [c for i, c in enumerate(cands) if c and c not in cands[:i]]:
        try:
            r = subprocess.run([cxx, "-fopenmp", "-x", "c++", "-", "-o", os.devnull], input=b"int main(){return 0;}",
                               capture_output=True, timeout=60)
        except (OSError, subprocess.TimeoutExpired):
            continue
        if r.returncode == 0:
            os.environ["CXX"] = cxx
            mod = sys.modules.get("torch._inductor.config")
            if mod is not None:                       # already imported: its default was read from the old $CXX
                mod.cpp.cxx = (None, cxx)
            return


def load():
    """Returns a namespace with the reference's attention / router / ulysses symbols."""
    if not available():
        raise RuntimeError(f"reference not found: neither /root/reference nor {STAGED_ROOT} (run oracle/stage_ref.py)")
    if kind() == "staged":
        from oracle import stage_ref
        stage_ref.verify()
    _pick_cxx()
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
