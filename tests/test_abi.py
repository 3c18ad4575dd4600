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


def _all_structs():
    structs = {"xmm_pack_segment": _lib.PackSegment, "xmm_pack_job": _lib.PackJob,
               "xmm_conv3x3_params": _lib.Conv3x3Params, "xmm_normalize_params": _lib.NormalizeParams,
               "xmm_conv_first_params": _lib.ConvFirstParams, "xmm_conv_last_params": _lib.ConvLastParams}
    structs.update(getattr(_lib, "EXTRA_STRUCTS", {}))
    return structs


def _c_field(name: str) -> str:
    return name[:-1] if name.endswith("_") else name  # `in_` / `in` (a Python keyword)


def test_ctypes_structs_match_c_layout(tmp_path, built):
    """sizeof of every struct AND offsetof of every field: the ctypes mirror has the header's fields, under the
    header's names, in the header's order (a missing / reordered / mistyped field changes an offset)."""
    structs = _all_structs()
    body = []
    for n, cls in structs.items():
        body.append(f'printf("{n} . %zu\\n", sizeof({n}));')
        for fname, *_ in cls._fields_:
            body.append(f'printf("{n} {fname} %zu\\n", offsetof({n}, {_c_field(fname)}));')
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){%s return 0;}\n' % (
        HEADER, "".join(body)))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {(a, b): int(c) for a, b, c in (line.split() for line in out.strip().splitlines())}
    for n, cls in structs.items():
        assert got[(n, ".")] == ctypes.sizeof(cls), n
        for fname, *_ in cls._fields_:
            assert got[(n, fname)] == getattr(cls, fname).offset, f"{n}.{fname}"
    # every field of the header's structs is mirrored (count the declarators between the braces)
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for m in re.finditer(r"typedef struct \{(.*?)\}\s*(xmm_[a-z0-9_]+);", hdr, flags=re.S):
        decls = [d for d in m.group(1).split(";") if d.strip()]
        nfields = sum(len(d.split(",")) for d in decls)
        assert m.group(2) in structs, f"{m.group(2)} has no ctypes mirror"
        assert nfields == len(structs[m.group(2)]._fields_), m.group(2)


def test_integration_md_does_not_redeclare_structs():
    """INTEGRATION.md's binding snippet imports the structures from _lib instead of re-declaring them (a pasted,
    stale copy would pass an undersized struct to the library)."""
    text = open(os.path.join(entry.ROOT, "INTEGRATION.md")).read()
    assert "ctypes.Structure" not in text and "_fields_" not in text
    assert "from xmm_superres_denoise_b200._lib import" in text


def test_library_reports_version_and_no_device_error(built):
    import torch

    assert built.xmm_version() >= 100
    assert built.xmm_pack_blob_bytes(32, 32, 5) == 5 * 9 * 32 * 32 * 2 + 32 * 4
    if not torch.cuda.is_available():
        # no GPU here: the device check must FAIL LOUDLY, not fall back
        assert built.xmm_check_device() != 0
        assert built.xmm_last_error()


def test_torch_library_registers_cuda_only_ops(built):
    """TORCH_LIBRARY(xmm_b200): the custom-op face registers, reports the C library's ABI version, and has NO CPU
    kernel (a CPU tensor is a dispatcher error, not a fallback)."""
    import torch

    from xmm_superres_denoise_b200 import torch_ops

    ns = torch_ops.load()
    assert ns.abi_version() == built.xmm_version()
    for name in ("conv3x3_fwd", "conv3x3_dgrad", "normalize", "denormalize", "image_upsample", "adam_step"):
        assert hasattr(ns, name)
    with pytest.raises(NotImplementedError):
        ns.image_upsample(torch.zeros(2, 2), torch.zeros(4, 4), 2)
    schema = str(ns.conv3x3_fwd.default._schema)
    assert "Tensor(a!) out" in schema and "wblob_row" in schema
