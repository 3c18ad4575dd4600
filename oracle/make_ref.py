"""Vendor the reference's own torch-only source files for the hot path into ``oracle/_ref/`` -- TEST
INFRASTRUCTURE, run in the authoring container (``__graft_entry__.build()`` calls it whenever
``/root/reference`` is mounted):

    python -m oracle.make_ref

``oracle/_ref/`` is listed in .gitignore (reference sources never enter the history) but not in .gpurunignore, so the
copy travels to the GPU box with the snapshot.  There it lets the GPU parity tests and ``bench.py --impl reference``
run the REFERENCE ITSELF (``cpu_baseline.kind = "reference"``) instead of the restatement in ``rrdb_oracle.py``:

    xmm_superres_denoise/models/modules/rrdb_blocks.py      ResidualDenseBlock_5C, RRDB, make_layer
    xmm_superres_denoise/models/modules/generator_rrdb.py   GeneratorRRDB_SR, GeneratorRRDB_DN
    xmm_superres_denoise/transforms/normalize.py            Normalize
    xmm_superres_denoise/transforms/imageupsample.py        ImageUpsample

These four import nothing but torch / numpy.  ``models/model.py`` (lightning), ``metrics/*`` (torchmetrics, piq) and
``data/*`` (astropy) are not importable in this image and are not copied.  The files are copied byte for byte;
``MANIFEST.json`` records their sha256 so a stale copy is detectable (``ref_loader.vendored_is_current``).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC_ROOT = os.environ.get("XMM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = [
    "xmm_superres_denoise/models/modules/rrdb_blocks.py",
    "xmm_superres_denoise/models/modules/generator_rrdb.py",
    "xmm_superres_denoise/transforms/normalize.py",
    "xmm_superres_denoise/transforms/imageupsample.py",
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make_ref(verbose: bool = True) -> bool:
    """Copy the files; returns False (and leaves any existing copy alone) when the reference tree is absent."""
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        if verbose:
            print(f"make_ref: no reference tree under {SRC_ROOT}; keeping {DST} as it is")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_ROOT, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"make_ref: {len(FILES)} files -> {DST}")
    return True


if __name__ == "__main__":
    make_ref()
