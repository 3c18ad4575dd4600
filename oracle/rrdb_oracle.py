"""CPU oracle for the RRDB hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch restatement of the reference's algorithm for the one path this
repository accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or the timed CPU baseline.  The product (``xmm_superres_denoise_b200``) never
imports this module and has no CPU fallback.

Every function cites the reference lines it follows (paths relative to
``/root/reference/xmm_superres_denoise/``).  Written as free functions over a
``state_dict`` so that it shares nothing with the reference's ``nn.Module`` classes.

Pinning status (see DESIGN.md "Oracle"):
  * network + transforms: PINNED -- ``oracle/make_golden.py`` ran the reference's own
    ``GeneratorRRDB_DN/SR``, ``Normalize`` and ``ImageUpsample`` (imported from
    /root/reference by file path) and committed their outputs under ``tests/golden``;
    ``tests/test_oracle_golden.py`` holds this file to them.
  * Poisson term: PINNED to ``torch.nn.functional.poisson_nll_loss`` (what
    ``metrics/metrics.py:36-38`` calls).
  * MAE / PSNR / SSIM / MS-SSIM terms: the arithmetic lives in the third-party package
    ``torchmetrics`` (unpinned in the reference's Dockerfile:10; poetry.lock names 0.11.4
    but the imports need >= 1.0), which is neither vendored in the reference nor installed
    here, and the reference holds no test or golden value at that boundary.  The restatement
    below follows torchmetrics >= 1.0 ``functional/image/ssim.py``.  It is pinned to an
    INDEPENDENT implementation instead -- float64 SSIM / MS-SSIM / PSNR / MAE built on
    ``scipy.ndimage`` Gaussian filtering (tests/test_oracle_loss.py, 416x416 and 832x832) -- and a
    GPU test diffs the CUDA loss against torchmetrics itself wherever that package exists
    (tests/test_gpu_loss.py).  Against torchmetrics proper these terms remain PARITY UNPINNED.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- network
def _conv(x: Tensor, sd: Mapping[str, Tensor], name: str) -> Tensor:
    return F.conv2d(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"), stride=1, padding=1)


def rdb_forward(x: Tensor, sd: Mapping[str, Tensor], prefix: str) -> Tensor:
    """ResidualDenseBlock_5C.forward -- models/modules/rrdb_blocks.py:37-54."""
    feats = [x]
    for k in range(1, 5):
        feats.append(F.leaky_relu(_conv(torch.cat(feats, 1), sd, f"{prefix}.conv{k}"), 0.2))
    x5 = _conv(torch.cat(feats, 1), sd, f"{prefix}.conv5")
    return x5 * 0.2 + x


def rrdb_forward(x: Tensor, sd: Mapping[str, Tensor], prefix: str) -> Tensor:
    """RRDB.forward -- models/modules/rrdb_blocks.py:66-70."""
    out = x
    for r in (1, 2, 3):
        out = rdb_forward(out, sd, f"{prefix}.RDB{r}")
    return out * 0.2 + x


def num_res_blocks(sd: Mapping[str, Tensor]) -> int:
    n = 0
    while f"rrdb.{n}.RDB1.conv1.weight" in sd:
        n += 1
    return n


def trunk_forward(x: Tensor, sd: Mapping[str, Tensor]) -> Tensor:
    """_GeneratorRRDB.forward -- models/modules/generator_rrdb.py:66-69."""
    fea = _conv(x, sd, "conv_first")
    t = fea
    for i in range(num_res_blocks(sd)):
        t = rrdb_forward(t, sd, f"rrdb.{i}")
    return fea + _conv(t, sd, "trunk_conv")


def generator_dn_forward(x: Tensor, sd: Mapping[str, Tensor]) -> Tensor:
    """GeneratorRRDB_DN.forward -- models/modules/generator_rrdb.py:130-137."""
    out = _conv(trunk_forward(x, sd), sd, "conv_last") + x
    return torch.clamp(out, 0.0, 1.0)


def generator_sr_forward(x: Tensor, sd: Mapping[str, Tensor], num_upsample: int) -> Tensor:
    """GeneratorRRDB_SR.forward -- models/modules/generator_rrdb.py:103-110.

    ``upsampling`` is [Conv2d, LeakyReLU() (slope 0.01), PixelShuffle(2)] * num_upsample
    (generator_rrdb.py:91-99), so the conv of stage s is ``upsampling.{3*s}``."""
    fea = trunk_forward(x, sd)
    for s in range(num_upsample):
        fea = F.pixel_shuffle(F.leaky_relu(_conv(fea, sd, f"upsampling.{3 * s}"), 0.01), 2)
    out = _conv(F.leaky_relu(_conv(fea, sd, "HRconv"), 0.2), sd, "conv_last")
    return torch.clamp(out, 0.0, 1.0)


def model_forward(x: Tensor, sd: Mapping[str, Tensor], kind: str, num_upsample: int = 1) -> Tensor:
    """Model.forward -- models/model.py:48-49 (a second, idempotent clamp)."""
    if kind == "dn":
        y = generator_dn_forward(x, sd)
    elif kind == "sr":
        y = generator_sr_forward(x, sd, num_upsample)
    else:
        raise ValueError(kind)
    return torch.clamp(y, 0.0, 1.0)


def splitmix_uniform(n: int, seed: int) -> "np.ndarray":
    """n floats in [0,1) from splitmix64(seed, index): closed-form, so fixtures never depend on a
    library's RNG stream."""
    import numpy as np

    with np.errstate(over="ignore"):
        z = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def init_state_dict(kind: str, in_ch: int, out_ch: int, nf: int, nb: int, num_upsample: int = 1,
                    seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic parameters with the keys/shapes of the reference generators
    (generator_rrdb.py:10-64,91-101; rrdb_blocks.py:27-31), drawn at nn.Conv2d's default
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) scale; conv_last gets the positive offset of
    generator_rrdb.py:56-64.  The values are NOT the reference's RNG stream."""
    sd: Dict[str, Tensor] = {}
    counter = [0]

    def uniform(shape, lo: float, hi: float) -> Tensor:
        n = 1
        for d in shape:
            n *= d
        counter[0] += 1
        u = splitmix_uniform(n, seed * 100003 + counter[0])
        return torch.from_numpy((lo + (hi - lo) * u).astype("float32")).reshape(shape)

    def conv(name: str, cin: int, cout: int) -> None:
        s = 1.0 / math.sqrt(cin * 9)
        sd[f"{name}.weight"] = uniform((cout, cin, 3, 3), -s, s)
        sd[f"{name}.bias"] = uniform((cout,), -s, s)

    conv("conv_first", in_ch, nf)
    for i in range(nb):
        for r in (1, 2, 3):
            for k in range(1, 5):
                conv(f"rrdb.{i}.RDB{r}.conv{k}", nf * k, nf)
            conv(f"rrdb.{i}.RDB{r}.conv5", nf * 5, nf)
    conv("trunk_conv", nf, nf)
    # generator_rrdb.py:59-64: stdv = 1/sqrt(weight.size(1)) -- channels only, not fan-in
    stdv = 1.0 / math.sqrt(nf)
    sd["conv_last.weight"] = uniform((out_ch, nf, 3, 3), -stdv, 1.01 * stdv)
    sd["conv_last.bias"] = uniform((out_ch,), -stdv, 1.01 * stdv)
    if kind == "sr":
        for s in range(num_upsample):
            conv(f"upsampling.{3 * s}", nf, nf * 4)
        conv("HRconv", nf, nf)
    return sd


# --------------------------------------------------------------------------- transforms
_ASINH_A = 0.02
_LOG_A = 1000.0


def stretch(x: Tensor, mode: str) -> Tensor:
    """transforms/normalize.py:4-32,55-62."""
    if mode == "linear":
        return x
    if mode == "sqrt":
        return torch.sqrt(x)
    if mode == "asinh":
        a = torch.tensor(_ASINH_A)
        return torch.asinh(x / a) / torch.asinh(1.0 / a)
    if mode == "log":
        a = torch.tensor(_LOG_A)
        return torch.log(a * x + 1) / torch.log(a)
    raise ValueError(f"Stretching function {mode} is not implemented")


def unstretch(x: Tensor, mode: str) -> Tensor:
    if mode == "linear":
        return x
    if mode == "sqrt":
        return torch.square(x)
    if mode == "asinh":
        a = torch.tensor(_ASINH_A)
        return a * torch.sinh(x * torch.asinh(1.0 / a))
    if mode == "log":
        a = torch.tensor(_LOG_A)
        return (torch.pow(a, x) - 1) / a
    raise ValueError(f"Stretching function {mode} is not implemented")


def normalize_image(image: Tensor, max_val: float, mode: str) -> Tensor:
    """Normalize.normalize_image -- transforms/normalize.py:66-82 (out-of-place here)."""
    if max_val > 0:
        image = torch.clamp(image, 0.0, max_val) / torch.tensor(max_val)
    else:
        image = image / torch.max(image)
    return torch.clamp(stretch(image, mode), 0.0, 1.0)


def denormalize_image(image: Tensor, max_val: Tensor, mode: str) -> Tensor:
    """Normalize.denormalize_image -- transforms/normalize.py:84-92; accepts the 0-dim
    max_val the class stores as well as the 1-D per-batch tensor the code indexes (I5)."""
    mv = max_val.reshape(-1)[:, None, None, None] if max_val.ndim else max_val
    return torch.minimum(torch.clamp_min(mv * unstretch(image, mode), 0.0), mv)


def image_upsample(x: Tensor, scale: int) -> Tensor:
    """ImageUpsample.__call__ -- transforms/imageupsample.py:10-26."""
    single = x.ndim < 4
    if single:
        x = x[None]
    x = F.interpolate(x, scale_factor=scale, mode="nearest") / (scale ** 2)
    return x[0] if single else x


def combine_mask_pad(img: Tensor, agn: Optional[Tensor], bkg: Optional[Tensor], det_mask: Optional[Tensor],
                     res: int) -> Tensor:
    """_load_and_combine_simulations + reshape_img_to_res -- data/dataset.py:24-49,
    data/tools.py:103-126: sum the parts, multiply by the detector mask, zero-pad to res x res
    with floor(diff/2) before and the rest after."""
    img = img.clone().float()
    if agn is not None:
        img += agn
    if bkg is not None:
        img += bkg
    if det_mask is not None:
        img *= det_mask
    h, w = img.shape[-2:]
    top, left = (res - h) // 2, (res - w) // 2
    return F.pad(img, (left, res - w - left, top, res - h - top))


# --------------------------------------------------------------------------- loss terms
def mae(preds: Tensor, target: Tensor) -> Tensor:
    """torchmetrics MeanAbsoluteError (utils/loss_functions.py:16): sum|p-t| / numel."""
    return torch.sum(torch.abs(preds - target)) / target.numel()


def poisson_nll(preds: Tensor, target: Tensor) -> Tensor:
    """PoissonNLLLoss batch value -- metrics/metrics.py:30-39: mean NLL divided by batch size."""
    return F.poisson_nll_loss(preds, target, log_input=False, reduction="mean") / preds.shape[0]


def psnr(preds: Tensor, target: Tensor) -> Tensor:
    """torchmetrics PeakSignalNoiseRatio(data_range=None): range tracked from the target with
    states initialised at 0."""
    zero = torch.zeros((), dtype=target.dtype)
    dr = torch.maximum(target.max(), zero) - torch.minimum(target.min(), zero)
    mse = torch.sum((preds - target) ** 2) / target.numel()
    return 10.0 * torch.log10(dr ** 2 / mse)


def gaussian_window(sigma: float, dtype: torch.dtype = torch.float32) -> Tensor:
    """torchmetrics ``_gaussian``: built in the dtype of the images (fp32 in the reference)."""
    size = int(3.5 * sigma + 0.5) * 2 + 1
    dist = torch.arange((1 - size) / 2, (1 + size) / 2, 1.0, dtype=dtype)
    g = torch.exp(-((dist / sigma) ** 2) / 2)
    return g / g.sum()


def ssim_sim_cs(preds: Tensor, target: Tensor, sigma: float = 2.5, k1: float = 0.01, k2: float = 0.05):
    """torchmetrics functional/image/ssim.py::_ssim_update with gaussian_kernel=True,
    data_range=None, return_contrast_sensitivity=True; (B,1,H,W) -> per-image (sim, cs)."""
    dr = max(preds.max() - preds.min(), target.max() - target.min())
    c1, c2 = (k1 * dr) ** 2, (k2 * dr) ** 2
    g = gaussian_window(sigma, preds.dtype)
    pad = (g.numel() - 1) // 2
    c = preds.shape[1]
    kernel = (g[:, None] * g[None, :]).expand(c, 1, -1, -1)
    p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    outs = F.conv2d(torch.cat((p, t, p * p, t * t, p * t)), kernel, groups=c).split(preds.shape[0])
    mu_pp, mu_tt, mu_pt = outs[0] ** 2, outs[1] ** 2, outs[0] * outs[1]
    s_pp = torch.clamp(outs[2] - mu_pp, min=0.0)
    s_tt = torch.clamp(outs[3] - mu_tt, min=0.0)
    s_pt = outs[4] - mu_pt
    upper = 2 * s_pt + c2
    lower = s_pp + s_tt + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower)
    sim = full[..., pad:-pad, pad:-pad]
    cs = (upper / lower)[..., pad:-pad, pad:-pad]
    b = preds.shape[0]
    return sim.reshape(b, -1).mean(-1), cs.reshape(b, -1).mean(-1)


def ssim(preds: Tensor, target: Tensor, sigma: float = 2.5, k1: float = 0.01, k2: float = 0.05) -> Tensor:
    """StructuralSimilarityIndexMeasure batch value (mean over images)."""
    return ssim_sim_cs(preds, target, sigma, k1, k2)[0].mean()


MS_SSIM_BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def ms_ssim(preds: Tensor, target: Tensor, sigma: float = 2.5, k1: float = 0.01, k2: float = 0.05,
            betas: Sequence[float] = MS_SSIM_BETAS) -> Tensor:
    """MultiScaleStructuralSimilarityIndexMeasure(normalize='relu') batch value."""
    vals = []
    sim = None
    for _ in betas:
        sim, cs = ssim_sim_cs(preds, target, sigma, k1, k2)
        sim, cs = torch.relu(sim), torch.relu(cs)
        vals.append(cs)
        preds, target = F.avg_pool2d(preds, 2), F.avg_pool2d(target, 2)
    vals[-1] = sim
    stack = torch.stack(vals)
    w = torch.tensor(betas, dtype=stack.dtype).view(-1, 1)
    return torch.prod(stack ** w, dim=0).mean()


LOSS_ORDER = ("l1", "poisson", "psnr", "ssim", "ms_ssim")  # config/config.py:222-227 field order
_TERMS = {"l1": mae, "poisson": poisson_nll, "psnr": psnr, "ssim": ssim, "ms_ssim": ms_ssim}


def composite_loss(preds: Tensor, target: Tensor, weights: Mapping[str, float],
                   sc_dict: Optional[Mapping[str, Mapping[str, float]]] = None) -> Tensor:
    """create_loss(...)(preds=, target=) -- utils/loss_functions.py:11-47."""
    correction = 0.0
    total = None
    for name in LOSS_ORDER:
        p = float(weights.get(name, 0.0))
        if p > 0.0:
            if sc_dict is not None and name in sc_dict:
                p = p * sc_dict[name]["scaling"]
                correction = correction + sc_dict[name]["correction"]
            term = _TERMS[name](preds, target) * p
            total = term if total is None else total + term
    assert total is not None
    if correction > 0.0:
        total = total + correction
    return total


# scaling/correction constants -- res/configs/loss_functions.toml:17-42
SCALING = {
    "linear": {"l1": (27.404768429706774, -0.5746779939709512), "poisson": (6.583278472679395, -1.187623436471363),
               "psnr": (-0.11938872970391594, 3.6491165234001905), "ssim": (-2.97441998810232, 2.1469363474122547),
               "ms_ssim": (-2.85143997718848, 2.737382378100941)},
    "sqrt": {"l1": (9.65623792970259, -0.5189262263422172), "poisson": (12.269938650306754, -5.137423312883438),
             "psnr": (-0.121713729308666, 2.7966163583252186), "ssim": (-3.0684258975145746, 1.417919607241485),
             "ms_ssim": (-3.0165912518853695, 2.636500754147813)},
    "asinh": {"l1": (5.651952749675013, -0.4542474424913807), "poisson": (0.4388467108439022, -0.22920963707377018),
              "psnr": (-0.11042402826855124, 2.1554770318021204), "ssim": (-3.2824552765468566, 1.2020351222714591),
              "ms_ssim": (-1.6189088554314395, 1.3368949328152826)},
    "log": {"l1": (4.071661237785016, -0.4364820846905537), "poisson": (0.39835876190096803, -0.2616021989403656),
            "psnr": (-0.1108524553818867, 1.8665336437202082), "ssim": (-3.414600833162603, 1.176671447107833),
            "ms_ssim": (-2.043318348998774, 1.6309767061708214)},
}


def sc_dict_for(stretch_mode: str) -> Dict[str, Dict[str, float]]:
    return {k: {"scaling": v[0], "correction": v[1]} for k, v in SCALING[stretch_mode].items()}
