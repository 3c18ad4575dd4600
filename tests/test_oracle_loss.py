"""CPU: the loss restatement.  Poisson is pinned to torch.nn.functional.poisson_nll_loss (what the
reference calls, metrics/metrics.py:36-38).  MAE / PSNR / SSIM / MS-SSIM follow torchmetrics, which
is not installed and not vendored by the reference: PARITY UNPINNED -- these tests hold the
restatement to the published definitions and to properties that any correct version satisfies."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rrdb_oracle as O


def _pair(b=2, n=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(b, 1, n, n, generator=g)
    p = (t + 0.1 * torch.randn(b, 1, n, n, generator=g)).clamp(0, 1)
    return p, t


def test_poisson_matches_torch_functional_and_batch_quirk():
    p, t = _pair(4)
    want = F.poisson_nll_loss(p, t, log_input=False, reduction="mean") / 4
    assert torch.equal(O.poisson_nll(p, t), want)
    manual = (p - t * torch.log(p + 1e-8)).mean() / 4
    assert abs(float(manual - want)) < 1e-6


def test_mae_and_psnr_definitions():
    p, t = _pair()
    assert abs(float(O.mae(p, t)) - float((p - t).abs().mean())) < 1e-7
    dr = float(t.max())  # min state starts at 0 and targets are non-negative
    want = 10 * math.log10(dr ** 2 / float(((p - t) ** 2).mean()))
    assert abs(float(O.psnr(p, t)) - want) < 1e-4


def test_gaussian_window_is_19_taps_for_sigma_2_5():
    g = O.gaussian_window(2.5)
    assert g.numel() == 19 and abs(float(g.sum()) - 1) < 1e-6 and torch.allclose(g, g.flip(0))


def test_ssim_identity_symmetry_and_range():
    p, t = _pair(2, 64)
    sim, cs = O.ssim_sim_cs(t, t)
    assert torch.allclose(sim, torch.ones(2), atol=1e-6) and torch.allclose(cs, torch.ones(2), atol=1e-6)
    s1, s2 = O.ssim(p, t), O.ssim(t, p)
    assert abs(float(s1 - s2)) < 1e-6 and 0 < float(s1) < 1


def test_ssim_against_direct_definition():
    # independent dense evaluation of the definition at a few interior pixels
    p, t = _pair(1, 48, seed=3)
    g = O.gaussian_window(2.5).double()
    k2d = g[:, None] * g[None, :]
    dr = max(float(p.max() - p.min()), float(t.max() - t.min()))
    c1, c2 = (0.01 * dr) ** 2, (0.05 * dr) ** 2
    vals = []
    for y in range(9, 48 - 9):
        for x in range(9, 48 - 9):
            wp = p[0, 0, y - 9:y + 10, x - 9:x + 10].double()
            wt = t[0, 0, y - 9:y + 10, x - 9:x + 10].double()
            mp, mt = (k2d * wp).sum(), (k2d * wt).sum()
            spp = (k2d * wp * wp).sum() - mp * mp
            stt = (k2d * wt * wt).sum() - mt * mt
            spt = (k2d * wp * wt).sum() - mp * mt
            vals.append(((2 * mp * mt + c1) * (2 * spt + c2)) / ((mp * mp + mt * mt + c1) * (spp + stt + c2)))
    want = float(torch.stack(vals).mean())
    assert abs(float(O.ssim(p, t)) - want) < 1e-5


def test_ms_ssim_identity_scales_and_gradient():
    p, t = _pair(2, 416 // 2, seed=1)  # 208 -> scales 208,104,52,26,13? 13 < 19: reflect pad needs > 9
    with pytest.raises(RuntimeError):
        O.ms_ssim(p[..., :100, :100], t[..., :100, :100])  # too small for 5 scales with a 19-tap window
    p, t = _pair(1, 416, seed=2)
    assert abs(float(O.ms_ssim(t, t)) - 1.0) < 1e-5
    p = p.clone().requires_grad_(True)
    v = O.ms_ssim(p, t)
    assert 0 < float(v) < 1
    v.backward()
    assert torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0


def test_composite_loss_follows_create_loss():
    p, t = _pair(2, 416, seed=4)
    sc = O.sc_dict_for("sqrt")
    w = {"l1": 0.5, "poisson": 0.5}
    want = O.mae(p, t) * (0.5 * sc["l1"]["scaling"]) + O.poisson_nll(p, t) * (0.5 * sc["poisson"]["scaling"])
    # the summed correction is negative here -> dropped (utils/loss_functions.py:44-45)
    assert sc["l1"]["correction"] + sc["poisson"]["correction"] < 0
    assert torch.allclose(O.composite_loss(p, t, w, sc), want)
    w = {"psnr": 0.5, "ms_ssim": 0.5}  # shipped default mix: positive correction is added
    corr = sc["psnr"]["correction"] + sc["ms_ssim"]["correction"]
    want = O.psnr(p, t) * (0.5 * sc["psnr"]["scaling"]) + O.ms_ssim(p, t) * (0.5 * sc["ms_ssim"]["scaling"]) + corr
    assert torch.allclose(O.composite_loss(p, t, w, sc), want)
    assert torch.allclose(O.composite_loss(p, t, {"l1": 1.0}, None), O.mae(p, t))
    with pytest.raises(AssertionError):
        O.composite_loss(p, t, {"l1": 0.0}, None)
