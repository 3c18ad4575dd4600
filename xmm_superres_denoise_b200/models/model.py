"""``Model``: the LightningModule of the reference (xmm_superres_denoise/models/model.py:13-247) around the
tensor-core generators, for the two RRDB model names (``esr_gen`` -> GeneratorRRDB_SR, ``rrdb_denoise`` ->
GeneratorRRDB_DN).  Same constructor arguments, ``forward`` (second clamp, model.py:48-49), ``_on_step``
(model.py:72-105), ``_on_epoch_end`` (107-151), ``configure_model`` (153-186) and ``configure_optimizers`` (239-247).

When ``lightning`` is installed the class derives from ``lightning.pytorch.LightningModule`` and plugs into
``Trainer`` unchanged (the autograd path of the generators gives ordinary ``.grad`` tensors, so DDP works).  Without
it (this image has no lightning) the same class derives from ``nn.Module`` and ``self.log`` / ``self.log_dict`` write
to ``self.logged``; ``training.fit`` in this package (or train.py at the repo root) drives it.
The transformer model names of the reference (SwinFIR, DRCT, HAT, Restormer) are outside the RRDB hot path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from ..transforms import ImageUpsample

try:  # pragma: no cover - depends on the environment
    import lightning.pytorch as _l

    _Base = _l.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _Base = nn.Module
    HAVE_LIGHTNING = False

ESR_GEN, RRDB_DENOISE = "esr_gen", "rrdb_denoise"


def _name(cfg_name) -> str:
    return str(getattr(cfg_name, "value", cfg_name)).lower()


class Model(_Base):
    def __init__(self, config, lr_shape: Tuple[int, int], hr_shape: Tuple[int, int], loss, metrics=None,
                 extended_metrics=None, in_metrics=None, in_extended_metrics=None):
        super().__init__()
        self.config = config
        self.metrics = metrics
        self.ext_metrics = extended_metrics
        self.in_metrics = in_metrics
        self.in_ext_metrics = in_extended_metrics
        self.loss = loss
        self.model: Optional[nn.Module] = None
        self.hr_shape = hr_shape
        self.lr_shape = lr_shape
        self.logged: Dict[str, Tensor] = {}

    # ---- Lightning-free logging sink
    if not HAVE_LIGHTNING:
        def log(self, name, value, **_kw) -> None:  # noqa: D401
            self.logged[name] = value.detach() if isinstance(value, Tensor) else value

        def log_dict(self, values, **_kw) -> None:
            for k, v in values.items():
                self.log(k, v)

    def forward(self, x) -> Tensor:
        y = self.model(x)
        if getattr(self.model, "output_is_clamped", False):
            return y  # already in [0, 1]: the second clamp is the identity (value and gradient)
        return torch.clamp(y, min=0.0, max=1.0)

    def training_step(self, batch, batch_idx=0):
        return self._on_step(batch, "train")

    def on_validation_start(self) -> None:
        self._on_epoch_end("train")

    def validation_step(self, batch, batch_idx=0):
        self._on_step(batch, "val")

    def on_validation_epoch_end(self) -> None:
        self._on_epoch_end("val")

    def test_step(self, batch, batch_idx=0):
        self._on_step(batch, "test")

    def on_test_epoch_end(self) -> None:
        self._on_epoch_end("test")

    def _on_step(self, batch, stage) -> Optional[Tensor]:
        lr_img, hr_img = batch
        preds = self(lr_img)
        target = hr_img if hr_img is not None else preds
        bs = getattr(self.config, "batch_size", lr_img.shape[0])
        if stage == "train":
            loss = self.loss(preds=preds, target=target)
            self.log(f"{stage}/loss", loss, batch_size=bs, on_step=True, on_epoch=False)
            return loss
        self.loss.update(preds=preds, target=target)
        if self.in_metrics is not None or self.in_ext_metrics is not None:
            scale_factor = target.shape[2] / lr_img.shape[2]
            if scale_factor != 1.0:
                lr_img = ImageUpsample(scale_factor=scale_factor)(lr_img)
        if self.metrics is not None:
            self.metrics.update(preds=preds, target=target)
        if self.in_metrics is not None:
            self.in_metrics.update(preds=lr_img, target=target)
        if self.ext_metrics is not None:
            self.ext_metrics.update(preds=preds, target=target)
        if self.in_ext_metrics is not None:
            self.in_ext_metrics.update(preds=lr_img, target=target)
        return None

    def _on_epoch_end(self, stage) -> None:
        bs = getattr(self.config, "batch_size", 1)
        if stage == "train":
            self.loss.reset()
            return
        self.log(f"{stage}/loss", self.loss.compute(), batch_size=bs, on_step=False, on_epoch=True, sync_dist=True)
        self.loss.reset()
        sanity = bool(getattr(getattr(self, "trainer", None), "sanity_checking", False)) if HAVE_LIGHTNING else False
        for attr in ("metrics", "ext_metrics", "in_metrics", "in_ext_metrics"):
            coll = getattr(self, attr)
            if coll is None:
                continue
            self.log_dict(coll.compute(), batch_size=bs, on_step=False, on_epoch=True, sync_dist=True)
            coll.reset()
            if attr.startswith("in_") and not sanity:  # the input metrics never change: log them once
                setattr(self, attr, None)

    def configure_model(self) -> None:
        if self.model is not None:
            return
        from . import GeneratorRRDB_DN, GeneratorRRDB_SR

        name, m = _name(self.config.name), self.config.model
        mem = bool(getattr(self.config, "memory_efficient", False))
        if name == ESR_GEN:
            up_scale = self.hr_shape[0] / self.lr_shape[0]
            if up_scale % 2 != 0:
                raise ValueError(f"Upscaling is not a multiple of two but {up_scale}, based on in_dims "
                                 f"{self.lr_shape} and out_dims {self.hr_shape}")
            self.model = GeneratorRRDB_SR(in_channels=m.in_channels, out_channels=m.out_channels,
                                          num_filters=m.filters, num_res_blocks=m.residual_blocks,
                                          num_upsample=int(up_scale / 2), memory_efficient=mem)
        elif name == RRDB_DENOISE:
            self.model = GeneratorRRDB_DN(in_channels=m.in_channels, out_channels=m.out_channels,
                                          num_filters=m.filters, num_res_blocks=m.residual_blocks,
                                          memory_efficient=mem)
        else:
            raise NotImplementedError(f"model '{name}': only the RRDB generators (esr_gen, rrdb_denoise) are on the "
                                      "accelerated path")

    def configure_optimizers(self):
        return torch.optim.Adam(self.model.parameters(), lr=self.config.optimizer.learning_rate,
                                betas=tuple(self.config.optimizer.betas))
