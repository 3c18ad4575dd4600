"""CPU: the C-ABI shared library builds, loads and exports exactly what include/xmm_b200.h
declares, and the ctypes structures mirror the C structs byte for byte (no GPU calls)."""
import ctypes
import os
import re
import subprocess

import pytest

import __graft_entry__ as entry
from xmm_superres_denoise_b200 import _lib

HEADER = os.path.join(entry.ROOT, "include", "xmm_b200.h")


@pytest.fixture(scope="module")
def built():
    entry.build()
    return _lib.load()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xmm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built):
    names = _declared_functions()
    assert len(names) >= 10
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (xmm_[a-z0-9_]+)", nm))
    for n in names:
        assert n in exported, f"{n} declared in xmm_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
        assert getattr(built, n) is not None
    assert exported == set(names), f"exported but undeclared: {exported - set(names)}"


def test_ctypes_structs_match_c_layout(tmp_path, built):
    structs = {"xmm_pack_segment": _lib.PackSegment, "xmm_pack_job": _lib.PackJob,
               "xmm_conv3x3_params": _lib.Conv3x3Params, "xmm_normalize_params": _lib.NormalizeParams,
               "xmm_conv_first_params": _lib.ConvFirstParams, "xmm_conv_last_params": _lib.ConvLastParams}
    extra = getattr(_lib, "EXTRA_STRUCTS", {})
    structs.update(extra)
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in structs)
    src = tmp_path / "sz.c"
    src.write_text(f'#include <stdio.h>\n#include "{HEADER}"\nint main(){{{body} return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n, cls in structs.items():
        assert int(sizes[n]) == ctypes.sizeof(cls), n


def test_library_reports_version_and_no_device_error(built):
    import torch

    assert built.xmm_version() >= 100
    assert built.xmm_pack_blob_bytes(32, 32, 5) == 5 * 9 * 32 * 32 * 2 + 32 * 4
    if not torch.cuda.is_available():
        # no GPU here: the device check must FAIL LOUDLY, not fall back
        assert built.xmm_check_device() != 0
        assert built.xmm_last_error()
