"""Synthetic XMM-like count images -- TEST / BENCH INFRASTRUCTURE.

Statistics follow the reference's example data (data/example_data/**, hand-parsed): 20 ks
images have ~0.4-1.0 counts/px, a few extended sources and tens of point sources; the
detector mask zeroes ~20 % of the 411x403 frame, which is then zero-padded to 416x416
(data/tools.py:103-126).  numpy only, seeded, closed-form sources + Poisson sampling.
"""
from __future__ import annotations

import numpy as np

LR_SHAPE = (411, 403)
EXPOSURE_LR = 20_000.0


def detector_mask(scale: int = 1) -> np.ndarray:
    """A synthetic EPIC-pn-like mask: circular field of view with chip gaps (~80 % ones)."""
    h, w = LR_SHAPE[0] * scale, LR_SHAPE[1] * scale
    yy, xx = np.mgrid[0:h, 0:w]
    cy, cx = (h - 1) / 2.0, (w - 1) / 2.0
    m = ((yy - cy) ** 2 + (xx - cx) ** 2) <= (0.5 * max(h, w)) ** 2
    for frac in (0.17, 0.33, 0.5, 0.67, 0.83):
        c = int(frac * w)
        m[:, c:c + scale] = False
    m[int(0.5 * h):int(0.5 * h) + scale, :] = False
    return m.astype(np.uint8)


def rate_map(rng: np.random.Generator, scale: int = 1) -> np.ndarray:
    """Expected counts per LR-exposure per pixel on the (411*scale, 403*scale) grid."""
    h, w = LR_SHAPE[0] * scale, LR_SHAPE[1] * scale
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    lam = np.full((h, w), 0.11, dtype=np.float32)
    for _ in range(int(rng.integers(1, 4))):
        cy, cx = rng.uniform(0.2 * h, 0.8 * h), rng.uniform(0.2 * w, 0.8 * w)
        sig = rng.uniform(10, 60) * scale
        lam += rng.uniform(0.5, 5.0) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sig * sig))
    for _ in range(int(rng.integers(20, 41))):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        sig = 1.5 * scale
        lam += rng.uniform(5, 250) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sig * sig)) / (scale * scale)
    return lam / (scale * scale) if scale > 1 else lam


def pad_to(img: np.ndarray, res: int) -> np.ndarray:
    """reshape_img_to_res (data/tools.py:103-126): floor(diff/2) before, the rest after."""
    h, w = img.shape[-2:]
    top, left = (res - h) // 2, (res - w) // 2
    out = np.zeros(img.shape[:-2] + (res, res), dtype=img.dtype)
    out[..., top:top + h, left:left + w] = img
    return out


def count_pair(seed: int, kind: str = "dn"):
    """(lr_counts int32 [416,416], hr_counts int32 [416|832]^2, exposure_lr, exposure_hr)."""
    rng = np.random.default_rng(seed)
    lam = rate_map(rng, 1)
    mask = detector_mask(1)
    lr = rng.poisson(lam).astype(np.int32) * mask
    if kind == "dn":
        t_hr = 50_000.0
        hr = rng.poisson(lam * (t_hr / EXPOSURE_LR)).astype(np.int32) * mask
        return pad_to(lr, 416), pad_to(hr, 416), EXPOSURE_LR, t_hr
    t_hr = 100_000.0
    lam2 = np.repeat(np.repeat(lam, 2, axis=0), 2, axis=1) / 4.0
    hr = rng.poisson(lam2 * (t_hr / EXPOSURE_LR)).astype(np.int32) * detector_mask(2)
    return pad_to(lr, 416), pad_to(hr, 832), EXPOSURE_LR, t_hr


def count_batch(n: int, seed: int, kind: str = "dn"):
    lrs, hrs = [], []
    t_lr = t_hr = 0.0
    for i in range(n):
        lr, hr, t_lr, t_hr = count_pair(seed * 1000 + i, kind)
        lrs.append(lr)
        hrs.append(hr)
    return np.stack(lrs)[:, None], np.stack(hrs)[:, None], t_lr, t_hr
