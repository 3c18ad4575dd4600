"""Validation metric collections on the GPU (reference: xmm_superres_denoise/metrics/xmm_metric_collection.py:14-38,
69-94,123-143; used by models/model.py:87-105,121-150 and train.py:73-88).

``get_metrics`` / ``get_in_metrics`` return an :class:`XMMMetricCollection` with the reference's interface
(``update(preds=, target=)``, ``compute() -> {name: 0-dim tensor}``, ``reset()``) and key names
(``"{prefix}/{stretch_mode}/[in/]{psnr,ssim,ms_ssim,l1,l2,poisson}"``).  Every ``update`` de-normalises both images
with the dataset normaliser and re-normalises them under each scaling normaliser in ONE kernel per image
(``xmm_restretch``), then accumulates all six metrics of that normaliser from one reduction pass plus the SSIM /
MS-SSIM kernels of the training loss (loss.py).  Accumulation follows torchmetrics: L1 / L2 over all elements,
PSNR with the running target range, SSIM / MS-SSIM / Poisson as means over images.

``get_ext_metrics`` / ``get_in_ext_metrics`` (VIF, FSIM, GMSD, HaarPSI, MDSI -- piq / torchvision models) are eval-only
extras outside the hot path and raise NotImplementedError.
"""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from .. import ops
from ..loss import CompositeLoss
from ..transforms import Normalize

METRIC_NAMES = ("psnr", "ssim", "ms_ssim", "l1", "l2", "poisson")  # xmm_metric_collection.py:21-31


class MetricSet(CompositeLoss):
    """All six metrics of one normaliser, accumulated together (one pass over the images per update)."""

    def __init__(self) -> None:
        super().__init__({"l1": 1.0, "poisson": 1.0, "psnr": 1.0, "ssim": 1.0, "ms_ssim": 1.0})

    def compute_all(self) -> Dict[str, torch.Tensor]:
        a = self._synced_state()  # summed over ranks (torchmetrics: dist_reduce_fx on every state)
        if a["n"] == 0:
            raise RuntimeError("compute() called before update()")
        dr = a["max_t"] - a["min_t"]
        return {
            "psnr": 10.0 * torch.log10(dr * dr / (a["sq"] / a["n"])),
            "ssim": a["ssim"] / a["b"],
            "ms_ssim": a["ms_ssim"] / a["b"],
            "l1": a["abs"] / a["n"],
            "l2": a["sq"] / a["n"],
            "poisson": a["poisson"] / a["b"],
        }


class XMMMetricCollection(nn.Module):
    def __init__(self, metrics, dataset_normalizer: Normalize, scaling_normalizers: List[Normalize], prefix: str):
        super().__init__()
        self.metric_names = tuple(metrics)
        for n in self.metric_names:
            if n.split("/")[-1] not in METRIC_NAMES:
                raise KeyError(f"metric {n} is not built on the GPU path")
        self.dataset_normalizer = dataset_normalizer
        self.normalizer_dict = {n.stretch_mode: n for n in scaling_normalizers}
        self.prefix = f"{prefix}/"
        self._sets = {mode: MetricSet() for mode in self.normalizer_dict}

    def keys(self):
        return [f"{self.prefix}{mode}/{name}" for mode in self.normalizer_dict for name in self.metric_names]

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        if not preds.is_cuda:
            raise RuntimeError("xmm_superres_denoise_b200 metrics run on CUDA (sm_100a) only; there is no CPU path")
        src = self.dataset_normalizer.stretch_mode
        p = preds.detach().contiguous().float()
        t = target.detach().contiguous().float()
        for mode, mset in self._sets.items():
            mset.update(preds=ops.restretch(p, src, mode), target=ops.restretch(t, src, mode))

    def compute(self) -> Dict[str, torch.Tensor]:
        out = {}
        for mode, mset in self._sets.items():
            vals = mset.compute_all()
            for name in self.metric_names:
                out[f"{self.prefix}{mode}/{name}"] = vals[name.split("/")[-1]]
        return out

    def reset(self) -> None:
        for mset in self._sets.values():
            mset.reset()

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> Dict[str, torch.Tensor]:
        """torchmetrics ``MetricCollection.forward``: accumulate, and return the values of THIS batch."""
        self.update(preds=preds, target=target)
        batch = XMMMetricCollection(self.metric_names, self.dataset_normalizer, list(self.normalizer_dict.values()),
                                    self.prefix[:-1])
        batch.update(preds=preds, target=target)
        return batch.compute()


def get_metrics(dataset_normalizer: Normalize, scaling_normalizers: List[Normalize], prefix: str) -> XMMMetricCollection:
    return XMMMetricCollection(METRIC_NAMES, dataset_normalizer, scaling_normalizers, prefix)


def get_in_metrics(dataset_normalizer: Normalize, scaling_normalizers: List[Normalize],
                   prefix: str) -> XMMMetricCollection:
    return XMMMetricCollection(tuple(f"in/{n}" for n in METRIC_NAMES), dataset_normalizer, scaling_normalizers, prefix)


def _out_of_scope(*_a, **_k):
    raise NotImplementedError("the piq / VGG based extended metrics (VIF, FSIM, GMSD, MS-GMSD, HaarPSI, MDSI) are "
                              "evaluation extras outside the accelerated path; use the reference's collection for them")


get_ext_metrics = get_in_ext_metrics = _out_of_scope
