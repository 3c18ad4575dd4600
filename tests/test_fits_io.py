"""CPU: FITS primary-HDU reader / writer and the WCS header bookkeeping of the inference output
(reference: utils/filehandling.py:131-247)."""
import gzip
import math

import numpy as np
import pytest

from xmm_superres_denoise_b200.utils import fits_io
from xmm_superres_denoise_b200.utils.filehandling import HEADER_KEYS_TO_OMIT, wcs_header


@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.uint8, np.int16, np.float64])
@pytest.mark.parametrize("gz", [False, True])
def test_fits_roundtrip(tmp_path, dtype, gz):
    rng = np.random.default_rng(0)
    data = (rng.random((411, 403)) * 100).astype(dtype)
    path = str(tmp_path / ("img.fits.gz" if gz else "img.fits"))
    hdr = {"EXPOSURE": (20000.0, "seconds"), "CRPIX1": 1, "CDELT1": 80.0, "OBJECT": "NGC 1234", "PA_PNT": 261.96,
           "GOOD": True, "LONGKEYWORD": 5}
    fits_io.write_primary(path, data, hdr, comments=["x" * 100])
    raw = (gzip.open if gz else open)(path, "rb").read()
    assert len(raw) % 2880 == 0 and raw.startswith(b"SIMPLE  =")
    back, h = fits_io.read_primary(path)
    assert back.dtype == np.dtype(dtype) and np.array_equal(back, data)
    assert h["EXPOSURE"] == 20000.0 and h["CRPIX1"] == 1 and h["OBJECT"] == "NGC 1234" and h["GOOD"] is True
    assert h["NAXIS1"] == 403 and h["NAXIS2"] == 411 and h["LONGKEYW"] == 5


def test_fits_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.fits"
    p.write_bytes(b"not a fits file" * 300)
    with pytest.raises(ValueError):
        fits_io.read_primary(str(p))


def test_wcs_header_follows_reference_rules():
    src = {"SIMPLE": True, "BITPIX": 32, "NAXIS": 2, "CRPIX1": 1.0, "CRPIX2": 1.0, "CDELT1": 80.0, "CDELT2": 80.0,
           "PA_PNT": 261.963195800781, "ONTIME03": 5.0, "EXPOSURE": 20000.0, "LTV1": 3.0, "TELESCOP": "XMM"}
    h1 = wcs_header(dict(src), "a.fits", 1, 20000.0)
    assert h1["CRPIX1"] == 7.0 and h1["CRPIX2"] == 3.0 and h1["CDELT1"] == 80.0  # pad offsets (6, 2)
    assert h1["TELESCOP"] == "XMM" and h1["EXPOSURE"] == 20000.0
    assert not any(k in h1 for k in HEADER_KEYS_TO_OMIT if k != "EXPOSURE")
    h2 = wcs_header(dict(src), "a.fits", 2, 100000.0)
    assert h2["CRPIX1"] == 2 * 7.0 + 0.5 and h2["CRPIX2"] == 2 * 3.0 + 0.5 and h2["CDELT1"] == 40.0
    rot = math.radians(90.0 - 261.963195800781)
    assert h2["CROT2"] == pytest.approx(90.0 - 261.963195800781)
    assert h2["CD1_1"] == pytest.approx(40.0 * math.cos(rot)) and h2["CD1_2"] == pytest.approx(-40.0 * math.sin(rot))
    assert h2["CD2_1"] == pytest.approx(40.0 * math.sin(rot)) and h2["CD2_2"] == pytest.approx(40.0 * math.cos(rot))
    assert h2["EXPOSURE"] == 100000.0
