"""Minimal FITS primary-HDU reader (numpy only) -- TEST INFRASTRUCTURE.

astropy is not installed; the reference reads its images with astropy.io.fits
(xmm_superres_denoise/data/tools.py:79-86).  A primary HDU is 2880-byte blocks of 80-char
header cards followed by big-endian data; that is all the example images and detector
masks use (optionally gzip-compressed)."""
from __future__ import annotations

import gzip

import numpy as np


def read_primary(path: str):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    header = {}
    off = 0
    done = False
    while not done:
        block = raw[off:off + 2880]
        off += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                val = card[10:].split("/")[0].strip()
                header[key] = val.strip("'").strip()
    bitpix = int(header["BITPIX"])
    naxis = int(header["NAXIS"])
    shape = [int(header[f"NAXIS{i}"]) for i in range(naxis, 0, -1)]
    dtype = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}[bitpix]
    n = int(np.prod(shape)) if shape else 0
    data = np.frombuffer(raw, dtype=dtype, count=n, offset=off).reshape(shape)
    bzero = float(header.get("BZERO", 0.0))
    bscale = float(header.get("BSCALE", 1.0))
    if bzero != 0.0 or bscale != 1.0:
        data = data * bscale + bzero
    return np.ascontiguousarray(data.astype(data.dtype.newbyteorder("="))), header
