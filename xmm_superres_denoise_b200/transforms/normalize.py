"""Drop-in ``Normalize`` (reference: xmm_superres_denoise/transforms/normalize.py:35-107).

Same constructor, attributes (``stretch_mode``, ``lr_max``, ``hr_max`` as 0-dim tensors,
``norm`` / ``denorm`` callables) and methods.  ``normalize_*`` / ``denormalize_*`` run one fused
CUDA kernel each (clamp, divide, stretch, clamp) instead of 4-5 elementwise passes on the CPU;
``normalize_counts`` additionally fuses the detector-mask multiply and the counts -> rate
division of the data pipeline (data/dataset.py:41-42).

Behavioural notes kept from the reference:
  * unknown ``stretch_mode`` raises ``ValueError`` (normalize.py:64);
  * ``normalize_hr_image(None)`` returns ``None`` (normalize.py:97-99);
  * ``normalize_image`` with ``max_val <= 0`` divides by the image maximum (normalize.py:73-75).
Deliberate differences:
  * the reference clamps the CALLER's tensor in place before dividing (normalize.py:70); the
    kernel reads the input once and leaves it untouched -- the returned image is identical;
  * ``denormalize_image`` accepts the 0-dim ``lr_max`` / ``hr_max`` the class itself stores, for
    which the reference raises IndexError, as well as a 1-D per-image tensor.
"""
from __future__ import annotations

import torch

from .. import ops


def _asinh(x: torch.Tensor):
    a = torch.tensor(0.02)
    return torch.asinh(x / a) / torch.asinh(1.0 / a)


def _asinh_inv(x: torch.Tensor):
    a = torch.tensor(0.02)
    return a * torch.sinh(x * torch.asinh(1.0 / a))


def _log(x: torch.Tensor):
    a = torch.tensor(1000)
    return torch.log(a * x + 1) / torch.log(a)


def _log_inv(x: torch.Tensor):
    a = torch.tensor(1000)
    return (torch.pow(a, x) - 1) / a


class Normalize:
    def __init__(self, lr_max: float, hr_max: float, stretch_mode: str = "linear"):
        assert isinstance(stretch_mode, str)
        self.stretch_mode = stretch_mode
        self.lr_max: torch.Tensor = torch.tensor(lr_max)
        self.hr_max: torch.Tensor = torch.tensor(hr_max)
        # host-side callables kept for the metric collections that call .norm/.denorm directly
        # (metrics/xmm_metric_collection.py:136-142); they are not on the kernel path.
        self.norm = self.denorm = None
        if stretch_mode == "linear":
            self.norm = self.denorm = lambda x: x
        elif stretch_mode == "sqrt":
            self.norm, self.denorm = torch.sqrt, torch.square
        elif stretch_mode == "log":
            self.norm, self.denorm = _log, _log_inv
        elif stretch_mode == "asinh":
            self.norm, self.denorm = _asinh, _asinh_inv
        else:
            raise ValueError(f"Stretching function {stretch_mode} is not implemented")

    def normalize_image(self, image: torch.Tensor, max_val: torch.Tensor) -> torch.Tensor:
        return ops.normalize(image.contiguous(), float(max_val), self.stretch_mode)

    def normalize_counts(self, counts: torch.Tensor, max_val, exposure: float = 1.0,
                         det_mask: torch.Tensor | None = None) -> torch.Tensor:
        """int32 / fp32 counts (* det_mask) / exposure -> normalised image, one kernel."""
        return ops.normalize(counts.contiguous(), float(max_val), self.stretch_mode, mask=det_mask,
                             pre_scale=1.0 / float(exposure))

    def denormalize_image(self, image: torch.Tensor, max_val: torch.Tensor):
        return ops.denormalize(image.contiguous(), torch.as_tensor(max_val), self.stretch_mode)

    def normalize_lr_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.normalize_image(image, max_val=self.lr_max)

    def normalize_hr_image(self, image: torch.Tensor | None):
        if image is None:
            return None
        return self.normalize_image(image, max_val=self.hr_max)

    def denormalize_lr_image(self, image: torch.Tensor):
        return self.denormalize_image(image, max_val=self.lr_max)

    def denormalize_hr_image(self, image):
        return self.denormalize_image(image, max_val=self.hr_max)
