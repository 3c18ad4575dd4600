"""Import the reference's own classes from /root/reference by file path -- TEST INFRASTRUCTURE.

Only usable in the authoring container (the reference tree does not exist on the GPU box);
used by ``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the CPU tests that
cross-check the oracle restatement against the live reference when it is present.

``import models`` would pull ``models/model.py`` -> ``lightning`` (not installed), so the two
network files and the two transform files are loaded individually behind a stub package.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("XMM_REFERENCE_ROOT", "/root/reference")
_PKG = os.path.join(REF_ROOT, "xmm_superres_denoise")


def available() -> bool:
    return os.path.isfile(os.path.join(_PKG, "models", "modules", "generator_rrdb.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with GeneratorRRDB_DN, GeneratorRRDB_SR, RRDB, Normalize, ImageUpsample."""
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT}")
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
