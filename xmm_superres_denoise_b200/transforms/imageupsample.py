"""Drop-in ``ImageUpsample`` (reference: xmm_superres_denoise/transforms/imageupsample.py:5-26):
nearest-neighbour upsample, then divide by ``scale_factor**2`` so the image sum is preserved.
3-D inputs (C,H,W) are treated as a single image like the reference does."""
from __future__ import annotations

import torch

from .. import ops


class ImageUpsample:
    def __init__(self, scale_factor):
        if float(scale_factor) != int(scale_factor) or int(scale_factor) < 1:
            raise ValueError(f"scale_factor must be a positive integer, got {scale_factor}")
        self.scale_factor = scale_factor

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return ops.image_upsample(x.contiguous(), int(self.scale_factor))
