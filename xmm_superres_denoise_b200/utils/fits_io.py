"""Minimal FITS primary-HDU reader / writer (numpy only).

The reference reads and writes its images with ``astropy.io.fits`` (data/tools.py:79-86,
utils/filehandling.py:131-247).  The XMM-Newton EPIC-pn count images, the detector masks and the files the
inference script writes are single primary HDUs -- 2880-byte blocks of 80-character header cards followed by
big-endian pixels, optionally gzip-compressed -- so this file is all the I/O the accelerated path needs; when
``astropy`` is installed the inference runner prefers it.
"""
from __future__ import annotations

import gzip
from collections import OrderedDict
from typing import Tuple

import numpy as np

_BLOCK = 2880
_DTYPES = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_BITPIX = {"uint8": 8, "int16": 16, "int32": 32, "int64": 64, "float32": -32, "float64": -64}


def _parse_value(text: str):
    text = text.strip()
    if text.startswith("'"):
        end = text.find("'", 1)
        while end != -1 and text[end:end + 2] == "''":
            end = text.find("'", end + 2)
        return text[1:end if end != -1 else None].replace("''", "'").rstrip()
    text = text.split("/")[0].strip()
    if text in ("T", "F"):
        return text == "T"
    try:
        return int(text)
    except ValueError:
        try:
            return float(text.replace("D", "E"))
        except ValueError:
            return text


def read_primary(path: str) -> Tuple[np.ndarray, "OrderedDict[str, object]"]:
    """Returns (data in native byte order, header as an ordered dict; COMMENT / HISTORY cards are skipped)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if raw[:6] != b"SIMPLE":
        raise ValueError(f"{path}: not a FITS file")
    header: "OrderedDict[str, object]" = OrderedDict()
    off, done = 0, False
    while not done:
        block = raw[off:off + _BLOCK]
        if len(block) < _BLOCK:
            raise ValueError(f"{path}: truncated FITS header")
        off += _BLOCK
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                header[key] = _parse_value(card[10:])
    naxis = int(header["NAXIS"])
    shape = [int(header[f"NAXIS{i}"]) for i in range(naxis, 0, -1)]
    dtype = _DTYPES[int(header["BITPIX"])]
    n = int(np.prod(shape)) if shape else 0
    data = np.frombuffer(raw, dtype=dtype, count=n, offset=off).reshape(shape)
    bzero, bscale = float(header.get("BZERO", 0.0)), float(header.get("BSCALE", 1.0))
    if bzero != 0.0 or bscale != 1.0:
        data = data * bscale + bzero
    return np.ascontiguousarray(data.astype(data.dtype.newbyteorder("="))), header


def _card(key: str, value, comment: str = "") -> str:
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20d}"
    elif isinstance(value, (float, np.floating)):
        v = f"{float(value):>20.14G}"
        if "." not in v and "E" not in v:
            v = f"{float(value):>20.1f}"
    else:
        s = str(value).replace("'", "''")
        v = f"'{s:<8}'"
    card = f"{key:<8}= {v}"
    if comment:
        card += f" / {comment}"
    return card[:80].ljust(80)


def write_primary(path: str, data: np.ndarray, header=None, comments=()) -> None:
    """Writes one primary HDU; ``header``: mapping key -> value or (value, comment); structural keywords are
    generated from ``data``.  ``.gz`` paths are gzip-compressed (the reference writes ``*.fits.gz``)."""
    data = np.asarray(data)
    if data.dtype.name not in _BITPIX:
        data = data.astype(np.float32)
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", _BITPIX[data.dtype.name]),
             _card("NAXIS", data.ndim)]
    for i, n in enumerate(reversed(data.shape), 1):
        cards.append(_card(f"NAXIS{i}", n))
    cards.append(_card("EXTEND", True))
    skip = {"SIMPLE", "BITPIX", "NAXIS", "EXTEND", "END"} | {f"NAXIS{i}" for i in range(1, 10)}
    for key, val in (header or {}).items():
        k = str(key).upper()[:8]
        if k in skip or k in ("COMMENT", "HISTORY", ""):
            continue
        if isinstance(val, tuple):
            cards.append(_card(k, val[0], str(val[1])))
        else:
            cards.append(_card(k, val))
    for c in comments:
        text = str(c)
        for i in range(0, max(len(text), 1), 72):
            cards.append(f"COMMENT {text[i:i + 72]}".ljust(80))
    cards.append("END".ljust(80))
    head = "".join(cards).encode("ascii", "replace")
    head += b" " * (-len(head) % _BLOCK)
    body = data.astype(data.dtype.newbyteorder(">")).tobytes()
    body += b"\0" * (-len(body) % _BLOCK)
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(head + body)
