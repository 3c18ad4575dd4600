"""Training loss on the CUDA kernels: the object ``create_loss`` returns.

Reference behaviour being reproduced (utils/loss_functions.py:11-47 + torchmetrics semantics):
a weighted sum of MeanAbsoluteError, PoissonNLLLoss (metrics/metrics.py:30-39: mean NLL divided
by the batch size), PeakSignalNoiseRatio, StructuralSimilarityIndexMeasure and
MultiScaleStructuralSimilarityIndexMeasure(kernel_size=13, sigma=2.5, k2=0.05), plus the summed
correction constant when it is positive.  ``forward(preds=, target=)`` returns the differentiable
value of the CURRENT batch and also merges it into the running state; ``update`` / ``compute`` /
``reset`` give the accumulated value used for validation logging (models/model.py:88,109-120).

All reductions, SSIM statistics and gradients run in libxmm_b200; the scalar glue between them is
a handful of 0-dim device tensor operations (no host synchronisation anywhere in forward/backward).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib, ops
from ._lib import MsssimFinalizeParams, SsimGradParams, SsimStatsParams

LOSS_ORDER = ("l1", "poisson", "psnr", "ssim", "ms_ssim")  # config/config.py:222-227
MS_SSIM_BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
SIGMA, K1, K2 = 2.5, 0.01, 0.05  # utils/loss_functions.py:33 (k1 is the torchmetrics default)
_STAT_FLOATS = 12


def gaussian_window(sigma: float = SIGMA) -> List[float]:
    size = int(3.5 * sigma + 0.5) * 2 + 1
    if size != 19:
        raise NotImplementedError("the SSIM kernels are built for the 19-tap window of sigma=2.5")
    g = [math.exp(-(((i - (size - 1) / 2) / sigma) ** 2) / 2) for i in range(size)]
    tot = sum(g)
    # round through fp32 exactly like torch would (window is built in the image dtype)
    t = torch.tensor(g, dtype=torch.float32)
    t = t / t.sum()
    return [float(v) for v in t]


_WINDOW = None


def _window():
    global _WINDOW
    if _WINDOW is None:
        _WINDOW = (ctypes.c_float * 19)(*gaussian_window())
    return _WINDOW


def _lib_call(fn_name: str, *args) -> None:
    _lib.check(getattr(_lib.load(), fn_name)(*args, _lib.stream_ptr()))
    ops._count(2 if fn_name in ("xmm_loss_reduce", "xmm_ssim_prepare") else 1)


class _SsimPipeline:
    """Forward + backward of SSIM (nscales=1) or MS-SSIM (nscales=5) for one (preds, target) pair."""

    def __init__(self, preds: torch.Tensor, target: torch.Tensor, nscales: int, stats0: torch.Tensor,
                 workspace: torch.Tensor) -> None:
        b, c, h, w = preds.shape
        self.b, self.c, self.nimg, self.nscales = b, c, b * c, nscales
        dev = preds.device
        if nscales > 1:
            # torchmetrics: image must stay larger than the window over all scales
            if min(h, w) // (2 ** (nscales - 1)) <= 18:
                raise ValueError(f"image {h}x{w} is too small for {nscales} MS-SSIM scales with a 19-tap window")
        elif min(h, w) <= 18:
            raise ValueError(f"image {h}x{w} is too small for a 19-tap SSIM window")
        self.p = [preds]
        self.t = [target]
        self.shapes = [(h, w)]
        self.stats = torch.zeros(nscales, _STAT_FLOATS, dtype=torch.float32, device=dev)
        self.stats[0].copy_(stats0)
        self.acc, self.tiles, self.nvalid = [], [], []
        lib = _lib.load()
        for s in range(nscales):
            hs, ws = self.shapes[s]
            if s > 0:
                hp, wp = self.shapes[s - 1]
                ps = torch.empty(b, c, hs, ws, dtype=torch.float32, device=dev)
                ts = torch.empty_like(ps)
                _lib_call("xmm_avgpool2_pair", self.p[s - 1].data_ptr(), self.t[s - 1].data_ptr(), ps.data_ptr(),
                          ts.data_ptr(), self.nimg, hp, wp)
                self.p.append(ps)
                self.t.append(ts)
                _lib_call("xmm_loss_reduce", ps.data_ptr(), ts.data_ptr(), ps.numel(), None,
                          self.stats[s].data_ptr(), workspace.data_ptr())
            _lib_call("xmm_ssim_prepare", self.p[s].data_ptr(), self.p[s].numel(), self.stats[s].data_ptr(), K1, K2)
            tiles = lib.xmm_ssim_tiles(hs, ws)
            acc = torch.empty(self.nimg * tiles * 4, dtype=torch.float32, device=dev)
            sp = SsimStatsParams()
            sp.preds, sp.target, sp.nimg, sp.h, sp.w = self.p[s].data_ptr(), self.t[s].data_ptr(), self.nimg, hs, ws
            sp.stats_dev, sp.window = self.stats[s].data_ptr(), _window()
            sp.use_sim = 1 if s == nscales - 1 else 0
            sp.acc = acc.data_ptr()
            _lib_call("xmm_ssim_stats", ctypes.byref(sp))
            self.acc.append(acc)
            self.tiles.append(tiles)
            self.nvalid.append((hs - 18) * (ws - 18))
            if s + 1 < nscales:
                self.shapes.append((hs // 2, ws // 2))
        self.value = torch.empty(1, dtype=torch.float32, device=dev)
        self.img_val = torch.empty(b, dtype=torch.float32, device=dev)
        self.kimg = torch.empty(nscales * self.nimg, dtype=torch.float32, device=dev)
        self._finalize(None, 1.0)

    def _finalize(self, gl: Optional[torch.Tensor], weight: float) -> None:
        fp = MsssimFinalizeParams()
        for s in range(self.nscales):
            fp.acc[s], fp.tiles[s], fp.nvalid[s] = self.acc[s].data_ptr(), self.tiles[s], self.nvalid[s]
            fp.betas[s] = MS_SSIM_BETAS[s] if self.nscales > 1 else 1.0
        fp.nscales, fp.batch, fp.channels, fp.k1, fp.k2 = self.nscales, self.b, self.c, K1, K2
        fp.stats_dev, fp.value, fp.img_val = self.stats.data_ptr(), self.value.data_ptr(), self.img_val.data_ptr()
        fp.kimg = self.kimg.data_ptr()
        fp.gl_dev = gl.data_ptr() if gl is not None else None
        fp.weight = weight
        _lib_call("xmm_msssim_finalize", ctypes.byref(fp))

    def backward(self, gl: torch.Tensor, weight: float, grad: torch.Tensor, accumulate: bool) -> None:
        """grad (+)= gl * weight * d value / d preds."""
        self._finalize(gl, weight)
        dev = grad.device
        coarse = None
        for s in range(self.nscales - 1, -1, -1):
            hs, ws = self.shapes[s]
            nv = self.nimg * self.nvalid[s]
            maps = torch.empty(3, nv, dtype=torch.float32, device=dev)
            sp = SsimStatsParams()
            sp.preds, sp.target, sp.nimg, sp.h, sp.w = self.p[s].data_ptr(), self.t[s].data_ptr(), self.nimg, hs, ws
            sp.stats_dev, sp.window = self.stats[s].data_ptr(), _window()
            sp.use_sim = 1 if s == self.nscales - 1 else 0
            sp.kimg = self.kimg[s * self.nimg:].data_ptr()
            sp.ga, sp.gb, sp.gc = maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr()
            _lib_call("xmm_ssim_stats", ctypes.byref(sp))
            out = grad if s == 0 else torch.empty(self.nimg * hs * ws, dtype=torch.float32, device=dev)
            gp = SsimGradParams()
            gp.preds, gp.target, gp.nimg, gp.h, gp.w = self.p[s].data_ptr(), self.t[s].data_ptr(), self.nimg, hs, ws
            gp.stats_dev, gp.window = self.stats[s].data_ptr(), _window()
            gp.ga, gp.gb, gp.gc = maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr()
            gp.coarse = coarse.data_ptr() if coarse is not None else None
            gp.grad = out.data_ptr()
            gp.accumulate = 1 if (s == 0 and accumulate) else 0
            _lib_call("xmm_ssim_grad", ctypes.byref(gp))
            coarse = out


class _CompositeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, target, owner):
        preds_c = preds.detach().contiguous().float()
        target_c = target.detach().contiguous().float()
        vals = owner._evaluate(preds_c, target_c)
        owner._last_state = vals
        ctx.owner, ctx.state = owner, vals
        ctx.save_for_backward(preds_c, target_c)
        return vals["total"].reshape(())

    @staticmethod
    def backward(ctx, gl):
        preds, target = ctx.saved_tensors
        grad = ctx.owner._gradient(preds, target, ctx.state, gl.contiguous().float().reshape(1))
        return grad, None, None


class CompositeLoss(nn.Module):
    """Weighted sum of loss terms; the public interface of torchmetrics' CompositionalMetric that the
    reference uses (``__call__(preds=, target=)``, ``update``, ``compute``, ``reset``)."""

    is_differentiable = True

    def __init__(self, weights: Dict[str, float], correction: float = 0.0) -> None:
        super().__init__()
        for k in weights:
            if k not in LOSS_ORDER:
                raise KeyError(f"unknown loss term {k}")
        self.weights = {k: float(weights[k]) for k in LOSS_ORDER if k in weights and weights[k] != 0.0}
        if not self.weights:
            raise AssertionError("create_loss needs at least one term with a positive weight")
        self.correction = float(correction)
        self._ws: Optional[torch.Tensor] = None
        self._last_state = None
        self.reset()

    def extra_repr(self) -> str:
        terms = " + ".join(f"{w:.6g}*{k}" for k, w in self.weights.items())
        return terms + (f" + {self.correction:.6g}" if self.correction > 0.0 else "")

    # ------------------------------------------------------------------ evaluation
    def _workspace(self, dev: torch.device) -> torch.Tensor:
        if self._ws is None or self._ws.device != dev:
            self._ws = torch.empty(_lib.load().xmm_loss_workspace_floats(), dtype=torch.float32, device=dev)
        return self._ws

    def _evaluate(self, preds: torch.Tensor, target: torch.Tensor) -> Dict[str, object]:
        if preds.shape != target.shape or preds.dim() != 4:
            raise RuntimeError(f"preds {tuple(preds.shape)} and target {tuple(target.shape)} must be equal (B,C,H,W)")
        _lib.require_cuda_tensor(preds, torch.float32, "loss preds")
        _lib.require_cuda_tensor(target, torch.float32, "loss target")
        dev = preds.device
        n, b = preds.numel(), preds.shape[0]
        sums = torch.empty(8, dtype=torch.float32, device=dev)
        stats0 = torch.zeros(_STAT_FLOATS, dtype=torch.float32, device=dev)
        ws = self._workspace(dev)
        _lib_call("xmm_loss_reduce", preds.data_ptr(), target.data_ptr(), n, sums.data_ptr(), stats0.data_ptr(),
                  ws.data_ptr())
        st: Dict[str, object] = {"sums": sums, "n": n, "b": b}
        terms: Dict[str, torch.Tensor] = {}
        if "l1" in self.weights:
            terms["l1"] = sums[0] / n
        if "poisson" in self.weights:
            terms["poisson"] = sums[1] / n / b  # mean NLL / batch size (metrics/metrics.py:36-39)
        if "psnr" in self.weights:
            zero = torch.zeros((), device=dev)
            dr = torch.maximum(sums[6], zero) - torch.minimum(sums[5], zero)
            mse = sums[2] / n
            terms["psnr"] = 10.0 * torch.log10(dr * dr / mse)
            st["mse"] = mse
        for name, nsc in (("ssim", 1), ("ms_ssim", 5)):
            if name in self.weights:
                pipe = _SsimPipeline(preds, target, nsc, stats0, ws)
                st[name] = pipe
                terms[name] = pipe.value[0]
        total = None
        for k in LOSS_ORDER:
            if k in terms:
                v = terms[k] * self.weights[k]
                total = v if total is None else total + v
        if self.correction > 0.0:
            total = total + self.correction
        st["terms"], st["total"] = terms, total
        return st

    def _gradient(self, preds: torch.Tensor, target: torch.Tensor, st: Dict[str, object], gl: torch.Tensor):
        n, b = st["n"], st["b"]
        dev = preds.device
        coef = torch.zeros(3, dtype=torch.float32, device=dev)
        # (fill_ / copy_ of device values only: a Python scalar assigned into a CUDA tensor is a host-to-device copy,
        # which a CUDA-graph capture of the training step rejects)
        if "l1" in self.weights:
            coef[0:1].fill_(self.weights["l1"] / n)
        if "poisson" in self.weights:
            coef[1:2].fill_(self.weights["poisson"] / (n * b))
        if "psnr" in self.weights:
            coef[2:3].copy_(((self.weights["psnr"] * (-20.0 / math.log(10.0)) / n) / st["mse"]).reshape(1))
        grad = torch.empty_like(preds)
        _lib_call("xmm_loss_grad", preds.data_ptr(), target.data_ptr(), n, coef.data_ptr(), gl.data_ptr(),
                  grad.data_ptr(), 0)
        for name in ("ssim", "ms_ssim"):
            if name in self.weights:
                st[name].backward(gl, self.weights[name], grad, accumulate=True)
        return grad

    # ------------------------------------------------------------------ torchmetrics-style interface
    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if not preds.is_cuda:
            raise RuntimeError("xmm_superres_denoise_b200 losses run on CUDA (sm_100a) only; there is no CPU path")
        if torch.is_grad_enabled() and preds.requires_grad:
            out = _CompositeLossFn.apply(preds, target, self)
            st = self._last_state
        else:
            st = self._evaluate(preds.detach().contiguous().float(), target.detach().contiguous().float())
            out = st["total"].reshape(())
        self._last_state = None
        self._merge(st)
        return out

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        with torch.no_grad():
            self._merge(self._evaluate(preds.detach().contiguous().float(), target.detach().contiguous().float()))

    def _merge(self, st: Dict[str, object]) -> None:
        s = st["sums"].detach()
        a = self._acc
        a["abs"] = a["abs"] + s[0]
        a["poisson"] = a["poisson"] + s[1] / st["n"]
        a["sq"] = a["sq"] + s[2]
        a["n"] += st["n"]
        a["b"] += st["b"]
        # PeakSignalNoiseRatio(data_range=None) tracks the target range with both states starting at 0
        zero = torch.zeros((), device=s.device)
        a["min_t"] = torch.minimum(zero if a["min_t"] is None else a["min_t"], s[5])
        a["max_t"] = torch.maximum(zero if a["max_t"] is None else a["max_t"], s[6])
        for name in ("ssim", "ms_ssim"):
            if name in st:
                a[name] = a[name] + st[name].img_val.detach().sum()

    def _synced_state(self) -> dict:
        """The accumulated state summed over the ranks of the default process group (torchmetrics declares these
        states with dist_reduce_fx="sum" / min / max and reduces them in compute(); Lightning's sync_dist relies on
        it): a mean of per-rank values is not the global metric, least of all for PSNR's running data range."""
        a = self._acc
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) or a["n"] == 0:
            return a
        dev = next((v.device for v in a.values() if isinstance(v, torch.Tensor)), None)
        if dev is None:
            return a
        keys = ["abs", "poisson", "sq", "ssim", "ms_ssim", "n", "b"]
        vec = torch.stack([torch.as_tensor(a[k], dtype=torch.float64, device=dev).reshape(()) for k in keys])
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        out = dict(a)
        for k, v in zip(keys, vec):
            out[k] = int(round(float(v))) if k in ("n", "b") else v.to(torch.float32)
        if a["min_t"] is not None:
            lo, hi = a["min_t"].clone(), a["max_t"].clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            out["min_t"], out["max_t"] = lo, hi
        return out

    def compute(self) -> torch.Tensor:
        a = self._synced_state()
        if a["n"] == 0:
            raise RuntimeError("compute() called before update()")
        terms = {}
        if "l1" in self.weights:
            terms["l1"] = a["abs"] / a["n"]
        if "poisson" in self.weights:
            terms["poisson"] = a["poisson"] / a["b"]
        if "psnr" in self.weights:
            dr = a["max_t"] - a["min_t"]
            terms["psnr"] = 10.0 * torch.log10(dr * dr / (a["sq"] / a["n"]))
        for name in ("ssim", "ms_ssim"):
            if name in self.weights:
                terms[name] = a[name] / a["b"]
        total = None
        for k in LOSS_ORDER:
            if k in terms:
                v = terms[k] * self.weights[k]
                total = v if total is None else total + v
        if self.correction > 0.0:
            total = total + self.correction
        return total

    def reset(self) -> None:
        self._acc = {"abs": 0.0, "poisson": 0.0, "sq": 0.0, "n": 0, "b": 0, "min_t": None, "max_t": None, "ssim": 0.0,
                     "ms_ssim": 0.0}
