"""Drop-in ``PoissonNLLLoss`` (reference: xmm_superres_denoise/metrics/metrics.py:9-39): running
``metric += mean(p - t*log(p + 1e-8))``, ``total += batch``; ``compute() = metric / total`` -- the
batch value is the mean NLL divided by the batch size.  The eval-only metrics of that file (MDSI,
HaarPSI, GMSD, FSIM, VGG loss: piq / torchvision models) are out of scope."""
from __future__ import annotations

import torch

from ..loss import CompositeLoss


class PoissonNLLLoss(CompositeLoss):
    higher_is_better = False
    full_state_update = False

    def __init__(self) -> None:
        super().__init__({"poisson": 1.0})
