"""Drop-in ``GeneratorRRDB_SR`` / ``GeneratorRRDB_DN`` (reference:
xmm_superres_denoise/models/modules/generator_rrdb.py:8-137).

Constructor signatures, attribute names, module tree, ``state_dict`` keys and the RNG
consumption order of the initialisation are the reference's; ``forward`` hands the batch to
``engine.RRDBEngine`` (sm_100a kernels).  Under ``torch.no_grad()`` / ``eval`` it runs the
inference sequence; with gradients enabled it runs through ``autograd_fn.GeneratorFunction``
whose backward is the hand-written data/weight-gradient kernels.
"""
from __future__ import annotations

import contextlib
import functools
import math

import torch
from torch import nn

from . import RRDB, make_layer


class _GeneratorRRDB(nn.Module):
    _kind = "trunk"
    # GeneratorRRDB_SR / _DN outputs leave the conv_last kernel already clamped to [0, 1] (generator_rrdb.py:109,136),
    # so Model.forward's second clamp (models/model.py:48-49) is the identity -- value and gradient -- and callers
    # that see this flag skip the extra elementwise pass (354 MB of traffic per batch-64 SR step).
    output_is_clamped = True

    def __init__(self, in_channels: int, out_channels: int, num_filters: int, num_res_blocks: int,
                 memory_efficient: bool = False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_filters = num_filters
        self.num_res_blocks = num_res_blocks
        self.memory_efficient = memory_efficient

        rrdb = functools.partial(RRDB, nf=self.num_filters, gc=num_filters, memory_efficient=self.memory_efficient)
        self.conv_first = nn.Conv2d(self.in_channels, self.num_filters, kernel_size=3, stride=1, padding=1)
        self.rrdb = make_layer(rrdb, self.num_res_blocks)
        self.trunk_conv = nn.Conv2d(self.num_filters, self.num_filters, kernel_size=3, stride=1, padding=1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        self.conv_last = nn.Conv2d(self.num_filters, self.out_channels, kernel_size=3, stride=1, padding=1)

        # generator_rrdb.py:56-64: bias conv_last towards positive outputs (the result is clamped)
        positive_offset_std = 0.01
        stdv = 1.0 / math.sqrt(self.conv_last.weight.size(1))
        self.conv_last.weight.data.uniform_(-stdv, stdv + positive_offset_std * stdv)
        if self.conv_last.bias is not None:
            self.conv_last.bias.data.uniform_(-stdv, stdv + positive_offset_std * stdv)
        self._engine = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None  # device buffers and packed weights are rebuilt on first use
        return state

    def _get_engine(self, train: bool = False):
        from ...engine import RRDBEngine

        if train:
            from ...engine_train import TrainEngine

            if not isinstance(self._engine, TrainEngine):
                self._engine = TrainEngine(self, self._kind)
        elif self._engine is None:
            self._engine = RRDBEngine(self, self._kind)
        return self._engine

    def forward(self, x):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        # libxmm_b200 launches on the CURRENT device / stream: make the input's GPU current for the call, so
        # `model.to("cuda:1")(x)` works without torch.cuda.set_device(1) (CPU tensors are rejected by the engine)
        with torch.cuda.device(x.device) if x.is_cuda else contextlib.nullcontext():
            if needs_grad:
                from ...autograd_fn import generator_apply

                return generator_apply(self, x)
            return self._get_engine().forward_inference(x)


class GeneratorRRDB_SR(_GeneratorRRDB):
    _kind = "sr"

    def __init__(self, in_channels: int, out_channels: int, num_filters: int, num_res_blocks: int,
                 num_upsample: int = 2, memory_efficient: bool = False):
        super().__init__(in_channels=in_channels, out_channels=out_channels, num_filters=num_filters,
                         num_res_blocks=num_res_blocks, memory_efficient=memory_efficient)
        self.num_upsample = num_upsample
        upsample_layers = []
        for _ in range(num_upsample):
            upsample_layers += [
                nn.Conv2d(num_filters, num_filters * 4, 3, 1, 1),
                nn.LeakyReLU(inplace=True),
                nn.PixelShuffle(upscale_factor=2),
            ]
        self.upsampling = nn.Sequential(*upsample_layers)
        self.HRconv = nn.Conv2d(num_filters, num_filters, 3, 1, 1, bias=True)


class GeneratorRRDB_DN(_GeneratorRRDB):
    _kind = "dn"

    def __init__(self, in_channels, out_channels, num_filters, num_res_blocks, memory_efficient=False):
        super().__init__(in_channels=in_channels, out_channels=out_channels, num_filters=num_filters,
                         num_res_blocks=num_res_blocks, memory_efficient=memory_efficient)
