"""Drop-in ``create_loss`` (reference: xmm_superres_denoise/utils/loss_functions.py:11-47).

``loss_config`` is the reference's pydantic ``LossCfg`` (iterating it yields ``(field, value)`` pairs
in the order l1, poisson, psnr, ssim, ms_ssim -- config/config.py:222-227), a mapping, or this
package's ``LossCfg`` dataclass.  Terms with weight > 0 are scaled by ``sc_dict[term]['scaling']``
and their ``correction`` values summed; the correction is added only when the sum is positive
(loss_functions.py:44-45).  Returns a ``CompositeLoss`` (CUDA kernels) instead of a torchmetrics
``CompositionalMetric``; it prints, moves with the LightningModule and exposes
``__call__(preds=, target=)`` / ``update`` / ``compute`` / ``reset`` like the original.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Dict, Mapping, Optional

from ..loss import LOSS_ORDER, CompositeLoss


@dataclass
class LossCfg:
    """Stand-in for the reference's pydantic LossCfg (config/config.py:218-237) when that package is not used."""
    l1: float = 0.0
    poisson: float = 0.0
    psnr: float = 0.0
    ssim: float = 0.0
    ms_ssim: float = 0.0

    def __post_init__(self):
        total = sum(getattr(self, f.name) for f in fields(self))
        if not 0.0 < total <= 1.0:
            raise ValueError(f"loss weights must sum to a value in (0, 1], got {total}")

    def __iter__(self):
        for f in fields(self):
            yield f.name, getattr(self, f.name)


def _pairs(loss_config):
    if isinstance(loss_config, Mapping):
        return [(k, loss_config[k]) for k in LOSS_ORDER if k in loss_config]
    return [(k, v) for k, v in iter(loss_config) if k in LOSS_ORDER]


def create_loss(sc_dict: Optional[Dict[str, Dict[str, float]]], loss_config) -> CompositeLoss:
    correction = 0.0
    weights: Dict[str, float] = {}
    for loss, p in _pairs(loss_config):
        if p > 0.0:
            if sc_dict is not None and loss in sc_dict:
                p = p * sc_dict[loss]["scaling"]
                correction = correction + sc_dict[loss]["correction"]
            weights[loss] = p
    assert weights
    return CompositeLoss(weights, correction if correction > 0.0 else 0.0)
