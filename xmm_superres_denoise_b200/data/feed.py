"""Batch data feed on the GPU.

The reference decodes gzip'd FITS on DataLoader workers and then, per sample on the CPU: sums image + AGN +
background, multiplies by the detector mask, optionally upsamples, zero-pads 411x403 -> 416x416, and normalises
(data/dataset.py:24-49,258-270; data/tools.py:103-126).  At ~1000 images/s per GPU those per-sample passes (and their
fp32 host copies) are the bottleneck, so here the workers only hand over the RAW int32 count planes:

* :func:`load_and_combine_simulations` -- the same arithmetic for a whole batch already on the device, ONE kernel
  (``xmm_prepare_counts``), fused with counts -> rate (``/exposure``) and ``Normalize``;
* :class:`CountsFeed` -- double-buffered pinned staging + a copy stream, so batch i+1's host->device copy of the raw
  planes (4 B/pixel instead of the reference's fp32 normalised 4 B/pixel PLUS its CPU passes) overlaps batch i's
  compute.

No CPU fallback: tensors must be CUDA tensors / the device must be an sm_100a GPU.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .. import ops
from ..transforms import Normalize


def load_and_combine_simulations(res: int, img: torch.Tensor, agn: Optional[torch.Tensor] = None,
                                 background: Optional[torch.Tensor] = None, det_mask: Optional[torch.Tensor] = None,
                                 upsample: int = 1, *, normalizer: Normalize, which: str = "lr", exposure=1.0,
                                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batch version of ``_load_and_combine_simulations`` + ``reshape_img_to_res`` + ``normalize_{lr,hr}_image``.

    img / agn / background: [B,h,w] or [B,1,h,w] count planes (int32 or fp32) on the GPU; det_mask: uint8 [h,w];
    exposure: seconds (float or [B] tensor).  Returns the normalised [B,1,res,res] fp32 network input / target."""
    if which not in ("lr", "hr"):
        raise ValueError("which must be 'lr' or 'hr'")
    max_val = float(normalizer.lr_max if which == "lr" else normalizer.hr_max)
    return ops.prepare_counts([img, agn, background], res, max_val, normalizer.stretch_mode, det_mask=det_mask,
                              exposure=exposure, upsample=upsample, out=out)


def reshape_img_to_res(res: int, img: torch.Tensor) -> torch.Tensor:
    """data/tools.py:103-126 for an image (batch) that is already a tensor: zero-pad (negative difference: crop)
    the last two dimensions to res x res, floor(diff/2) before and the rest after.  Pure tensor plumbing -- inside
    the fused feed the same arithmetic is part of ``xmm_prepare_counts``."""
    h, w = img.shape[-2:]
    top, left = (res - h) // 2, (res - w) // 2
    return torch.nn.functional.pad(img, (left, res - w - left, top, res - h - top), mode="constant", value=0)


class CountsFeed:
    """Pinned, double-buffered host -> device feed of raw count batches.

    ``submit(planes...)`` copies numpy int32 planes into a pinned slot and enqueues the H2D copy + the fused prepare
    kernel on a side stream; ``get()`` makes the compute stream wait for it and returns the normalised batch.  Two
    slots: submit batch i+1 before consuming batch i to overlap copy and compute."""

    def __init__(self, batch: int, h: int, w: int, res: int, normalizer: Normalize, *, which: str = "lr",
                 nplanes: int = 1, det_mask: Optional[np.ndarray] = None, upsample: int = 1,
                 device: Optional[torch.device] = None) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("CountsFeed needs a CUDA device (sm_100a); there is no CPU path")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.shape = (batch, h, w)
        self.res, self.normalizer, self.which, self.upsample = res, normalizer, which, upsample
        self.nplanes = nplanes
        self._host = [[torch.empty(batch, h, w, dtype=torch.int32).pin_memory() for _ in range(nplanes)]
                      for _ in range(2)]
        self._dev = [[torch.empty(batch, h, w, dtype=torch.int32, device=self.device) for _ in range(nplanes)]
                     for _ in range(2)]
        out_res = res
        self._out = [torch.empty(batch, 1, out_res, out_res, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._mask = None if det_mask is None else torch.from_numpy(
            np.ascontiguousarray(det_mask.astype(np.uint8))).to(self.device)
        self._stream = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._used = [False, False]
        self._pending: list = []
        self._slot = 0
        self.h2d_bytes_per_batch = batch * h * w * 4 * nplanes

    def submit(self, planes: Sequence[np.ndarray], exposure=1.0) -> None:
        if len(planes) != self.nplanes:
            raise RuntimeError(f"expected {self.nplanes} planes")
        if len(self._pending) >= 2:
            raise RuntimeError("both staging slots are in flight: call get() first")
        s = self._slot
        self._slot ^= 1
        if self._used[s]:
            self._ready[s].synchronize()  # this slot's previous host->device copy has left the pinned buffer
        self._used[s] = True
        for hbuf, p in zip(self._host[s], planes):
            hbuf.numpy()[...] = p
        # the slot's device buffers were handed out two batches ago: everything the compute stream has enqueued so
        # far (which includes that batch's consumers) must finish before they are overwritten
        self._stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):
            for hbuf, dbuf in zip(self._host[s], self._dev[s]):
                dbuf.copy_(hbuf, non_blocking=True)
            d = self._dev[s]
            load_and_combine_simulations(self.res, d[0], d[1] if self.nplanes > 1 else None,
                                         d[2] if self.nplanes > 2 else None, self._mask, self.upsample,
                                         normalizer=self.normalizer, which=self.which, exposure=exposure,
                                         out=self._out[s])
            self._ready[s].record(self._stream)
        self._pending.append(s)

    def get(self) -> torch.Tensor:
        if not self._pending:
            raise RuntimeError("get() without a submitted batch")
        s = self._pending.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[s])
        # valid until the next-but-one submit(); clone to keep it longer
        return self._out[s]
