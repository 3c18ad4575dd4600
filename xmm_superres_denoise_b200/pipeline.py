"""Overlapped inference pipeline: host counts in -> host images out.

The reference's predict path is serial per batch (DataLoader -> ``Model.predict_step`` -> ``.cpu()``).  On a B200 the
SR-2x generator produces 177 MB of fp32 output per batch of 64 -- 4-5 ms of PCIe time next to ~56 ms of compute --
so the three legs run on three streams with double-buffered device staging:

    h2d stream     : raw int32 counts (pinned) -> device                      (batch i+1)
    compute stream : fused prepare kernel (mask / pad / rate / normalise) + generator      (batch i)
    d2h stream     : prediction -> pinned host buffer                         (batch i-1)

``submit`` is asynchronous; ``wait(slot)`` / ``synchronize`` make the host buffers safe to read / reuse.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .data import load_and_combine_simulations
from .transforms import Normalize


class InferencePipeline:
    def __init__(self, model: torch.nn.Module, normalizer: Normalize, batch: int, h: int, w: int, res: int, *,
                 det_mask: Optional[np.ndarray] = None, exposure: float = 1.0, denormalize: bool = False) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("InferencePipeline needs a CUDA device (sm_100a); there is no CPU path")
        self.model = model.eval()
        self.norm, self.res, self.exposure, self.denormalize = normalizer, res, exposure, denormalize
        self.device = next(model.parameters()).device
        self.dev_in = [torch.empty(batch, h, w, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.mask = None if det_mask is None else torch.from_numpy(
            np.ascontiguousarray(det_mask.astype(np.uint8))).to(self.device)
        self.h2d = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        self.ev_in = [torch.cuda.Event(), torch.cuda.Event()]       # counts of slot s are on the device
        self.ev_free = [torch.cuda.Event(), torch.cuda.Event()]     # compute has consumed dev_in[s]
        self.ev_done = [torch.cuda.Event(), torch.cuda.Event()]     # host output of slot s is complete
        self._used = [False, False]
        self._slot = 0
        self.h2d_bytes_per_batch = batch * h * w * 4

    def submit(self, counts_host: torch.Tensor, out_host: torch.Tensor) -> int:
        """counts_host: pinned int32 [B,h,w] (or [B,1,h,w]); out_host: pinned fp32 [B,1,H,W].  Returns the slot."""
        s = self._slot
        self._slot ^= 1
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.h2d):
            if self._used[s]:
                self.h2d.wait_event(self.ev_free[s])  # the previous batch in this slot has been read by compute
            self.dev_in[s].copy_(counts_host.reshape(self.dev_in[s].shape), non_blocking=True)
            self.ev_in[s].record(self.h2d)
        cur.wait_event(self.ev_in[s])
        x = load_and_combine_simulations(self.res, self.dev_in[s], det_mask=self.mask, normalizer=self.norm,
                                         which="lr", exposure=self.exposure)
        self.ev_free[s].record(cur)
        with torch.no_grad():
            y = self.model(x)  # models/model.py:48-49
            if not getattr(self.model, "output_is_clamped", False):  # our generators clamp in the conv_last epilogue
                y = torch.clamp(y, 0.0, 1.0)
        if self.denormalize:
            y = self.norm.denormalize_hr_image(y)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(ready)
            out_host.copy_(y, non_blocking=True)
            y.record_stream(self.d2h)
            self.ev_done[s].record(self.d2h)
        self._used[s] = True
        return s

    def wait(self, slot: int) -> None:
        self.ev_done[slot].synchronize()

    def synchronize(self) -> None:
        for s in range(2):
            if self._used[s]:
                self.ev_done[s].synchronize()
