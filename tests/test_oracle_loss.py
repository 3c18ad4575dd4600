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


# ------------------------------------------------------------------------------------------------------------------
# An INDEPENDENT float64 SSIM / MS-SSIM: scipy.ndimage.gaussian_filter(sigma, truncate=3.5) is the window torchmetrics
# builds (radius int(3.5 * sigma + 0.5) = 9 -> 19 taps of exp(-x^2 / (2 sigma^2)), normalised) and scikit-image's
# structural_similarity formulation; torchmetrics reflect-pads by 9 and crops the same 9 pixels again, so on the
# cropped region the padding mode cannot matter.  Nothing below shares code with oracle/rrdb_oracle.py.
def _scipy_sim_cs(p, t, sigma=2.5, k1=0.01, k2=0.05):
    from scipy.ndimage import gaussian_filter

    dr = max(p.max() - p.min(), t.max() - t.min())   # data_range=None: from the current tensors, whole batch
    c1, c2 = (k1 * dr) ** 2, (k2 * dr) ** 2
    r = int(3.5 * sigma + 0.5)
    sims, css = [], []
    for pi, ti in zip(p[:, 0], t[:, 0]):
        f = lambda a: gaussian_filter(a, sigma=sigma, truncate=3.5, mode="nearest")[r:-r, r:-r]  # noqa: E731
        mp, mt = f(pi), f(ti)
        spp = np.maximum(f(pi * pi) - mp * mp, 0.0)
        stt = np.maximum(f(ti * ti) - mt * mt, 0.0)
        spt = f(pi * ti) - mp * mt
        cs = (2 * spt + c2) / (spp + stt + c2)
        sims.append(float(((2 * mp * mt + c1) / (mp * mp + mt * mt + c1) * cs).mean()))
        css.append(float(cs.mean()))
    return np.array(sims), np.array(css)


def _scipy_ms_ssim(p, t, betas=(0.0448, 0.2856, 0.3001, 0.2363, 0.1333)):
    vals = []
    for _ in betas:
        sim, cs = _scipy_sim_cs(p, t)
        sim, cs = np.maximum(sim, 0.0), np.maximum(cs, 0.0)          # normalize="relu"
        vals.append(cs)
        b, c, h, w = p.shape
        p = p.reshape(b, c, h // 2, 2, w // 2, 2).mean(axis=(3, 5))  # avg_pool2d(2)
        t = t.reshape(b, c, h // 2, 2, w // 2, 2).mean(axis=(3, 5))
    vals[-1] = sim
    return float(np.prod(np.stack(vals) ** np.array(betas)[:, None], axis=0).mean())


def _count_like_pair(b, n, seed):
    """Sqrt-stretched Poisson-count images (what the loss sees in training), float64."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n] / n
    rate = 0.3 + 6.0 * np.exp(-((yy - 0.4) ** 2 + (xx - 0.6) ** 2) / 0.02) + 2.0 * np.exp(-((yy - 0.7) ** 2 + (xx - 0.2) ** 2) / 0.1)
    t = np.minimum(1.0, np.sqrt(rng.poisson(rate * 4.0, size=(b, 1, n, n)) / 60.0))
    p = np.clip(np.sqrt(rate / 15.0)[None, None] + 0.03 * rng.standard_normal((b, 1, n, n)), 0.0, 1.0)
    return p, t


@pytest.mark.parametrize("n", [416, 832])
def test_ssim_and_ms_ssim_against_independent_scipy_float64(n):
    """BASELINE's two resolutions: the restated SSIM (per-image sim, cs) and MS-SSIM agree with the scipy evaluation
    to float64 round-off (oracle run in float64) and to 2e-5 in the float32 the reference computes in."""
    p, t = _count_like_pair(2, n, seed=n)
    sim_w, cs_w = _scipy_sim_cs(p, t)
    ms_w = _scipy_ms_ssim(p, t)
    p64, t64 = torch.from_numpy(p), torch.from_numpy(t)
    sim, cs = O.ssim_sim_cs(p64, t64)
    np.testing.assert_allclose(sim.numpy(), sim_w, rtol=0, atol=1e-10)
    np.testing.assert_allclose(cs.numpy(), cs_w, rtol=0, atol=1e-10)
    assert abs(float(O.ms_ssim(p64, t64)) - ms_w) < 1e-10
    assert abs(float(O.ssim(p64, t64)) - sim_w.mean()) < 1e-10
    p32, t32 = p64.float(), t64.float()
    assert abs(float(O.ms_ssim(p32, t32)) - ms_w) < 2e-5
    assert abs(float(O.ssim(p32, t32)) - sim_w.mean()) < 2e-5
    assert 0.05 < ms_w < 0.999  # a non-trivial value


def test_psnr_and_mae_against_numpy_float64():
    p, t = _count_like_pair(3, 416, seed=5)
    mse = ((p - t) ** 2).mean()
    dr = max(t.max(), 0.0) - min(t.min(), 0.0)
    assert abs(float(O.psnr(torch.from_numpy(p), torch.from_numpy(t))) - 10 * np.log10(dr ** 2 / mse)) < 1e-9
    assert abs(float(O.mae(torch.from_numpy(p), torch.from_numpy(t))) - np.abs(p - t).mean()) < 1e-12
