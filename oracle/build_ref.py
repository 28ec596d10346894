"""Recipe for ``oracle/_ref``: a byte-for-byte snapshot of the reference's model modules, taken from where they lie under
/root/reference, so that the CPU arm of bench.py (``--impl reference``, ``cpu_baseline``) times THE REFERENCE ITSELF on the
GPU box's host cores (cpu_baseline.kind == "reference") instead of the oracle port. ORACLE-side tooling: test infrastructure.

    python oracle/build_ref.py          # also run by __graft_entry__.build() whenever /root/reference is present

The reference is pure Python (nothing to compile): the four files that define the fusion path -- fcmf_framework/
{mm_modeling,roi_modeling,fcmf_pretraining,fcmf_multimodal}.py -- are copied UNMODIFIED into oracle/_ref/fcmf_framework/
together with a SHA-256 manifest. ``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history)
but not gpurun-ignored, so the snapshot travels to the GPU box like a built .so. Nothing outside bench.py's CPU legs and
tests/ may import it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FCMF_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
FILES = ("mm_modeling.py", "roi_modeling.py", "fcmf_pretraining.py", "fcmf_multimodal.py")


def build(verbose: bool = True) -> str | None:
    src = os.path.join(REF, "fcmf_framework")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/_ref: {src} not present (GPU box): keeping the snapshot that travelled with the repo")
        return OUT if os.path.isdir(os.path.join(OUT, "fcmf_framework")) else None
    dst = os.path.join(OUT, "fcmf_framework")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        manifest[f] = hashlib.sha256(open(os.path.join(dst, f), "rb").read()).hexdigest()
    open(os.path.join(dst, "__init__.py"), "w").close()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"oracle/_ref: snapshot of {len(FILES)} reference modules -> {dst}")
    return OUT


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, "fcmf_framework", f)) for f in FILES)


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
