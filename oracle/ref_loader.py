"""Import the reference's own classes from /root/reference by file path -- TEST INFRASTRUCTURE.

In the authoring container this is the live tree; on the GPU box (no /root/reference) it is the copy of the four
torch-only files that ``oracle/make_ref.py`` vendors into ``oracle/_ref/`` (git-ignored, not gpurun-ignored).  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz``, by the tests that cross-check the oracle restatement and
the CUDA path against the reference's own classes, and by ``bench.py --impl reference``.

``import models`` would pull ``models/model.py`` -> ``lightning`` (not installed), so the two
network files and the two transform files are loaded individually behind a stub package.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("XMM_REFERENCE_ROOT", "/root/reference")
# the copy `python -m oracle.make_ref` vendors (git-ignored; travels to the GPU box with the snapshot)
VENDORED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pkg_dir() -> str:
    """The live reference tree when it is mounted, else the vendored copy under oracle/_ref."""
    for root in (REF_ROOT, VENDORED_ROOT):
        pkg = os.path.join(root, "xmm_superres_denoise")
        if os.path.isfile(os.path.join(pkg, "models", "modules", "generator_rrdb.py")):
            return pkg
    return os.path.join(REF_ROOT, "xmm_superres_denoise")


def available() -> bool:
    return os.path.isfile(os.path.join(_pkg_dir(), "models", "modules", "generator_rrdb.py"))


def source() -> str:
    """"live" (/root/reference), "vendored" (oracle/_ref) or "absent"."""
    if not available():
        return "absent"
    return "live" if _pkg_dir().startswith(REF_ROOT) else "vendored"


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with GeneratorRRDB_DN, GeneratorRRDB_SR, RRDB, Normalize, ImageUpsample."""
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT} or {VENDORED_ROOT} (python -m oracle.make_ref)")
    _PKG = _pkg_dir()
    saved = {k: sys.modules.get(k) for k in ("models", "models.modules")}
    try:
        pkg = types.ModuleType("models")
        pkg.__path__ = []  # mark as package
        sys.modules["models"] = pkg
        blocks = _load("models.modules", os.path.join(_PKG, "models", "modules", "rrdb_blocks.py"))
        gen = _load("_xmm_ref_generator_rrdb", os.path.join(_PKG, "models", "modules", "generator_rrdb.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    norm = _load("_xmm_ref_normalize", os.path.join(_PKG, "transforms", "normalize.py"))
    ups = _load("_xmm_ref_imageupsample", os.path.join(_PKG, "transforms", "imageupsample.py"))
    ns = types.SimpleNamespace(
        GeneratorRRDB_DN=gen.GeneratorRRDB_DN, GeneratorRRDB_SR=gen.GeneratorRRDB_SR, RRDB=blocks.RRDB,
        ResidualDenseBlock_5C=blocks.ResidualDenseBlock_5C, Normalize=norm.Normalize, ImageUpsample=ups.ImageUpsample)
    return ns
