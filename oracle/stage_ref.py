"""Stage the reference's own Python files for the hot path into ``oracle/_ref/`` — TEST / BENCH INFRASTRUCTURE ONLY.

    python oracle/stage_ref.py            # build container only: needs /root/reference

The reference is pure Python (no setup.py / pyproject, so ``pip install --target baseline/_ref`` has nothing to build;
DESIGN.md section 4).  ``oracle/_ref/`` plays the role a compiled ``oracle/_ref/*.so`` plays for a C reference: it is
git-ignored (the sources never enter this repository's history) but NOT gpurun-ignored, so the UNMODIFIED files
travel to the GPU box with the snapshot, where ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the
reference's own ``WanAttnProcessorTripleEval`` on the host cores.  Only the files of the path are staged
(SURVEY.md section 8a): ``vorta/attention``, ``vorta/patch/router.py``, ``vorta/ulysses``.  A manifest with the sha256 of every
staged file is written next to them; ``verify()`` re-checks it before the files are used.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VORTA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "vorta/attention/__init__.py", "vorta/attention/coreset_select.py", "vorta/attention/hunyuan.py",
    "vorta/attention/sliding_attn_flex.py", "vorta/attention/tile.py", "vorta/attention/wan.py",
    "vorta/patch/__init__.py", "vorta/patch/router.py",
    "vorta/ulysses/__init__.py", "vorta/ulysses/parallel_states.py", "vorta/ulysses/utils.py",
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def source_available() -> bool:
    return all(os.path.isfile(os.path.join(SRC, f)) for f in FILES)


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "MANIFEST.json")) and all(
        os.path.isfile(os.path.join(DST, f)) for f in FILES)


def stage() -> str:
    """Copy the path's files byte for byte; returns the staging directory."""
    if not source_available():
        raise RuntimeError(f"reference tree not found at {SRC}")
    manifest = {}
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
        manifest[f] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(dict(source="wenhao728/VORTA (unmodified files of the routed-attention path)", files=manifest), fh,
                  indent=1, sort_keys=True)
    return DST


def verify() -> None:
    """The staged files are the ones the manifest was written for (and equal the reference tree when it is present)."""
    with open(os.path.join(DST, "MANIFEST.json")) as fh:
        manifest = json.load(fh)["files"]
    for f in FILES:
        got = _sha(os.path.join(DST, f))
        if got != manifest.get(f):
            raise RuntimeError(f"oracle/_ref/{f} does not match its manifest: staged reference files were modified")
        src = os.path.join(SRC, f)
        if os.path.isfile(src) and _sha(src) != got:
            raise RuntimeError(f"oracle/_ref/{f} differs from {src}: re-run oracle/stage_ref.py")


if __name__ == "__main__":
    print("staged", len(FILES), "reference files into", stage())
    verify()
