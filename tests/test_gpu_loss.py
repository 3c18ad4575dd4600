"""GPU (B200): the loss kernels against the oracle restatement (values and autograd gradients).
Floating-point tolerance: loss relative error <= 1e-2 (north_star); these fp32 kernels are held to
1e-4 on values and 1e-3 relative L2 on gradients."""
import numpy as np
import pytest
import torch

from oracle import rrdb_oracle as O
from oracle.synthetic import count_batch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from xmm_superres_denoise_b200 import _lib

    _lib.check(_lib.load().xmm_check_device())
    return torch.device("cuda:0")


def _pair(b, h, w, seed, saturate=True):
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(b, 1, h, w, generator=g)
    t[t < 0.3] = 0.0  # exact zeros like masked detector regions
    p = t + 0.15 * torch.randn(b, 1, h, w, generator=g)
    p = p.clamp(0, 1) if saturate else p.clamp(0.02, 0.97)
    return p, t


def _check(dev, weights, sc, p, t, tol_val=1e-4, tol_grad=1e-3):
    from xmm_superres_denoise_b200.utils.loss_functions import create_loss

    po = p.clone().requires_grad_(True)
    want = O.composite_loss(po, t, weights, sc)
    want.backward()
    loss = create_loss(sc, weights).to(dev)
    pg = p.to(dev).requires_grad_(True)
    got = loss(preds=pg, target=t.to(dev))
    (got * 1.0).backward()
    assert got.dim() == 0
    assert abs(float(got) - float(want)) <= tol_val * max(1.0, abs(float(want))), (float(got), float(want))
    r = rel_l2(pg.grad.cpu(), po.grad)
    print(weights, "value", float(got), float(want), "grad rel", r)
    assert r < tol_grad
    return loss


def test_l1_poisson_config3(dev):
    """BASELINE.json config 3: L1 + Poisson with the paper's sqrt scaling (negative correction is dropped)."""
    p, t = _pair(4, 64, 48, 0)
    loss = _check(dev, {"l1": 0.5, "poisson": 0.5}, O.sc_dict_for("sqrt"), p, t)
    assert loss.correction == 0.0 and "l1" in repr(loss)


def test_psnr_and_default_mix(dev):
    p, t = _pair(2, 416, 416, 1)
    _check(dev, {"psnr": 1.0}, None, p, t)
    _check(dev, {"psnr": 0.5, "ms_ssim": 0.5}, O.sc_dict_for("sqrt"), p, t, tol_grad=3e-3)  # shipped default mix


@pytest.mark.parametrize("saturate", [True, False])
def test_ssim_single_scale(dev, saturate):
    p, t = _pair(3, 70, 90, 2, saturate)
    _check(dev, {"ssim": 1.0}, None, p, t, tol_grad=3e-3)


def test_ms_ssim_416_and_832(dev):
    p, t = _pair(2, 416, 416, 3)
    _check(dev, {"ms_ssim": 1.0}, None, p, t, tol_grad=3e-3)
    p, t = _pair(1, 832, 832, 4, saturate=False)
    _check(dev, {"ms_ssim": 1.0}, None, p, t, tol_grad=3e-3)


def test_config4_mix_on_synthetic_counts(dev):
    """BASELINE.json config 4: L1 + Poisson + MS-SSIM on SR-shaped (832x832) count-image targets."""
    _, hr, _, t_hr = count_batch(2, seed=11, kind="sr")
    t = O.normalize_image(torch.from_numpy(hr.astype(np.float32) / t_hr), 0.0005584, "sqrt")
    g = torch.Generator().manual_seed(5)
    p = (t * 0.8 + 0.05 + 0.05 * torch.randn(t.shape, generator=g)).clamp(0, 1)
    _check(dev, {"l1": 0.3, "poisson": 0.3, "ms_ssim": 0.4}, O.sc_dict_for("sqrt"), p, t, tol_grad=3e-3)


def test_state_update_compute_reset_and_errors(dev):
    from xmm_superres_denoise_b200.metrics import PoissonNLLLoss
    from xmm_superres_denoise_b200.utils.loss_functions import LossCfg, create_loss

    loss = create_loss(None, LossCfg(l1=0.5, poisson=0.5)).to(dev)
    batches = [_pair(2, 40, 40, s) for s in (7, 8, 9)]
    for p, t in batches:
        loss.update(preds=p.to(dev), target=t.to(dev))
    abs_sum = sum(float((p - t).abs().sum()) for p, t in batches)
    n = sum(p.numel() for p, _ in batches)
    pois = sum(float(torch.nn.functional.poisson_nll_loss(p, t, log_input=False)) for p, t in batches) / 6
    assert abs(float(loss.compute()) - (0.5 * abs_sum / n + 0.5 * pois)) < 1e-5
    loss.reset()
    with pytest.raises(RuntimeError):
        loss.compute()
    m = PoissonNLLLoss().to(dev)
    p, t = batches[0]
    v = m(preds=p.to(dev), target=t.to(dev))
    assert abs(float(v) - float(O.poisson_nll(p, t))) < 1e-6
    with pytest.raises(ValueError):
        LossCfg(l1=0.8, poisson=0.8)
    with pytest.raises(ValueError, match="too small"):
        create_loss(None, {"ms_ssim": 1.0}).to(dev)(preds=p.to(dev), target=t.to(dev))
    with pytest.raises(RuntimeError, match="no CPU path"):
        create_loss(None, {"l1": 1.0})(preds=p, target=t)
    with pytest.raises(AssertionError):
        create_loss(None, {"l1": 0.0})


def test_against_torchmetrics_when_the_box_has_it(dev):
    """The reference's loss terms ARE torchmetrics objects (utils/loss_functions.py:3-9, metrics/metrics.py:9-39).
    torchmetrics is in neither this image nor the authoring container, so the oracle for these terms is pinned to an
    independent float64 scipy implementation instead (tests/test_oracle_loss.py).  Wherever the package does exist,
    this test compares the CUDA loss directly with torchmetrics; otherwise the skip reason records its absence."""
    try:
        import torchmetrics  # noqa: F401
        from torchmetrics import MeanAbsoluteError
        from torchmetrics.image import (MultiScaleStructuralSimilarityIndexMeasure, PeakSignalNoiseRatio,
                                        StructuralSimilarityIndexMeasure)
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"torchmetrics absent on this box ({type(e).__name__}: {e}); the loss oracle stays pinned to scipy float64")
    from xmm_superres_denoise_b200.utils.loss_functions import create_loss

    p, t = _pair(2, 416, 416, 3)
    pd, td = p.to(dev), t.to(dev)
    for name, metric in (("l1", MeanAbsoluteError()), ("psnr", PeakSignalNoiseRatio()),
                         ("ssim", StructuralSimilarityIndexMeasure(kernel_size=13, sigma=2.5, k2=0.05)),
                         ("ms_ssim", MultiScaleStructuralSimilarityIndexMeasure(kernel_size=13, sigma=2.5, k2=0.05))):
        want = float(metric(p, t))
        got = float(create_loss(None, {name: 1.0}).to(dev)(preds=pd, target=td))
        assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (name, got, want)
