"""File handling of the inference entry point (reference: xmm_superres_denoise/utils/filehandling.py:131-247
``write_xmm_file_to_fits_wcs`` and ``read_yaml``): the WCS bookkeeping of the padded / super-resolved image is kept
exactly -- CRPIX shifted by the (6, 2) pixel padding, and for the 2x output CRPIX / CDELT rescaled and a CD matrix
added from PA_PNT -- on top of the numpy-only FITS writer of ``fits_io``."""
from __future__ import annotations

import math
import os
from datetime import datetime

import numpy as np
import yaml

from . import fits_io

# structural / provenance keywords of the input header that must not be copied (filehandling.py:147-196)
HEADER_KEYS_TO_OMIT = (
    "SIMPLE", "BITPIX", "NAXIS", "NAXIS1", "NAXIS2", "EXTEND", "XPROC0", "XDAL0", "CREATOR", "DATE",
    "CTYPE1L", "CRPIX1L", "CRVAL1L", "CDELT1L", "LTV1", "LTM1_1",
    "CTYPE2L", "CRPIX2L", "CRVAL2L", "CDELT2L", "LTV2", "LTM2_2", "LTM1_2", "LTM2_1",
    *(f"ONTIME{i:02d}" for i in range(1, 13)), "EXPOSURE", "DURATION",
)


def read_yaml(file_path):
    with open(file_path, "r") as f:
        return yaml.safe_load(f)


def wcs_header(in_header, source_file_name: str, res_mult: int, exposure: float) -> dict:
    """The output header of ``write_xmm_file_to_fits_wcs`` (filehandling.py:142-225) as a plain dict."""
    header = {"IMG_FILE": (source_file_name, "Input source file")}
    if in_header is not None:
        for key, val in in_header.items():
            if key not in HEADER_KEYS_TO_OMIT:
                header[key] = val
    header["EXPOSURE"] = exposure
    if "CRPIX1" in header and "CRPIX2" in header:
        crpix1_new = float(header["CRPIX1"]) + 6  # left padding of 403 -> 416
        crpix2_new = float(header["CRPIX2"]) + 2  # top padding of 411 -> 416
        header["CRPIX1"], header["CRPIX2"] = crpix1_new, crpix2_new
        if res_mult == 2:
            header["CRPIX1"] = res_mult * crpix1_new + 0.5
            header["CRPIX2"] = res_mult * crpix2_new + 0.5
            cdelt1 = float(header["CDELT1"]) / res_mult
            cdelt2 = float(header["CDELT2"]) / res_mult
            header["CDELT1"], header["CDELT2"] = cdelt1, cdelt2
            crota2 = 90.0 - float(header["PA_PNT"])
            header["CROT2"] = crota2
            r = math.radians(crota2)
            header["CD1_1"] = cdelt1 * math.cos(r)
            header["CD1_2"] = -1.0 * cdelt2 * math.sin(r)
            header["CD2_1"] = cdelt1 * math.sin(r)
            header["CD2_2"] = cdelt2 * math.cos(r)
    return header


def write_xmm_file_to_fits_wcs(img, output_dir, source_file_name, res_mult, exposure, comment=None,
                               out_file_name=None, in_header=None) -> str:
    header = wcs_header(in_header, source_file_name, res_mult, exposure)
    comments = []
    if comment is not None:
        comments.append(comment)
    comments.append("Written by xmm_superres_denoise_b200 (B200 build of SamSweere/xmm-superres-denoise)")
    comments.append("WCS handling after utils/filehandling.py (Ivan V) of the reference")
    comments.append(f"File created on {datetime.now().strftime('%d/%m/%Y %H:%M:%S')}")
    if out_file_name is None:
        out_file_name = f"{source_file_name.replace('.fits', '')}_sr_predict"
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(str(output_dir), f"{out_file_name}.fits.gz")
    fits_io.write_primary(path, np.asarray(img, dtype=np.float32), header, comments)
    return path
