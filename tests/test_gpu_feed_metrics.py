"""GPU (B200): the widened rows of SURVEY.md section 8(f) -- validation metric collection (f2) and the fused data
feed (f3) -- against the oracle restatement of the reference arithmetic."""
import numpy as np
import pytest
import torch

from oracle import rrdb_oracle as O
from oracle.synthetic import count_batch, detector_mask

from helpers import rel_l2

pytestmark = pytest.mark.gpu

LR_MAX, HR_MAX = 0.0022336, 0.0005584


@pytest.fixture(scope="module")
def dev():
    from xmm_superres_denoise_b200 import _lib

    _lib.check(_lib.load().xmm_check_device())
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------ f3: data feed
@pytest.mark.parametrize("mode", ["linear", "sqrt", "asinh", "log"])
@pytest.mark.parametrize("planes,upsample", [(1, 1), (3, 1), (1, 2)])
def test_prepare_counts_matches_reference_pipeline(dev, mode, planes, upsample):
    """int32 planes -> (sum) * mask -> [upsample] -> pad 411x403 -> 416 (832) -> /exposure -> Normalize, one kernel,
    vs dataset.py:24-49 + tools.py:103-126 + normalize.py:66-82 restated by the oracle."""
    from xmm_superres_denoise_b200.data import load_and_combine_simulations
    from xmm_superres_denoise_b200.transforms import Normalize

    rng = np.random.default_rng(11 + planes + upsample)
    b, h, w = 3, 411, 403
    srcs = [rng.poisson(lam, size=(b, h, w)).astype(np.int32) for lam in (1.0, 0.05, 0.1)[:planes]]
    mask = detector_mask(1)
    exposure = 20000.0
    res = 416 * upsample
    tens = [torch.from_numpy(s) for s in srcs]
    if upsample == 1:
        combined = O.combine_mask_pad(tens[0][:, None], tens[1][:, None] if planes > 1 else None,
                                      tens[2][:, None] if planes > 2 else None,
                                      torch.from_numpy(mask.astype(np.float32)), res)
    else:  # ImageUpsample sits between the mask multiply and the padding (dataset.py:41-47)
        comb = sum(t.float() for t in tens)[:, None] * torch.from_numpy(mask.astype(np.float32))
        up = O.image_upsample(comb, upsample)
        top, left = (res - h * upsample) // 2, (res - w * upsample) // 2
        combined = torch.nn.functional.pad(up, (left, res - w * upsample - left, top, res - h * upsample - top))
    want = O.normalize_image(combined / exposure, LR_MAX, mode)
    norm = Normalize(LR_MAX, HR_MAX, mode)
    got = load_and_combine_simulations(res, tens[0].to(dev), tens[1].to(dev) if planes > 1 else None,
                                       tens[2].to(dev) if planes > 2 else None, torch.from_numpy(mask).to(dev),
                                       upsample, normalizer=norm, which="lr", exposure=exposure).cpu()
    assert got.shape == (b, 1, res, res)
    assert torch.allclose(got, want, atol=2e-6, rtol=1e-5), float((got - want).abs().max())
    assert float(got[:, :, :2].abs().max()) == 0.0  # the padded border is exactly zero


def test_prepare_counts_per_image_exposure_and_crop(dev):
    from xmm_superres_denoise_b200 import ops

    rng = np.random.default_rng(5)
    src = torch.from_numpy(rng.poisson(2.0, size=(2, 40, 52)).astype(np.float32))
    expo = torch.tensor([10.0, 40.0])
    got = ops.prepare_counts([src.to(dev)], (32, 44), 0.5, "sqrt", exposure=expo.to(dev)).cpu()
    top, left = (32 - 40) // 2, (44 - 52) // 2  # negative: crop, floor(diff/2) first (tools.py:111-117)
    crop = torch.nn.functional.pad(src[:, None], (left, 44 - 52 - left, top, 32 - 40 - top))
    want = O.normalize_image(crop / expo.view(2, 1, 1, 1), 0.5, "sqrt")
    assert torch.allclose(got, want, atol=2e-6, rtol=1e-5)


def test_counts_feed_double_buffering(dev):
    from xmm_superres_denoise_b200.data import CountsFeed
    from xmm_superres_denoise_b200.transforms import Normalize

    norm = Normalize(LR_MAX, HR_MAX, "sqrt")
    mask = detector_mask(1)
    feed = CountsFeed(2, 411, 403, 416, norm, det_mask=mask)
    rng = np.random.default_rng(0)
    batches = [rng.poisson(1.0, size=(2, 411, 403)).astype(np.int32) for _ in range(5)]
    outs = []
    feed.submit([batches[0]], exposure=20000.0)
    for i in range(5):
        if i + 1 < 5:
            feed.submit([batches[i + 1]], exposure=20000.0)
        outs.append(feed.get().clone())
    torch.cuda.synchronize()
    for bt, got in zip(batches, outs):
        comb = O.combine_mask_pad(torch.from_numpy(bt)[:, None], None, None, torch.from_numpy(mask.astype(np.float32)), 416)
        want = O.normalize_image(comb / 20000.0, LR_MAX, "sqrt")
        assert torch.allclose(got.cpu(), want, atol=2e-6, rtol=1e-5)


# ------------------------------------------------------------------------------ f2: validation metrics
def test_metric_collection_matches_oracle(dev):
    """Two updates of get_metrics (dataset normaliser sqrt; scaling normalisers linear + sqrt) vs the oracle's
    restatement of torchmetrics' accumulation (xmm_metric_collection.py:14-38,135-143)."""
    from xmm_superres_denoise_b200.metrics import get_in_metrics, get_metrics
    from xmm_superres_denoise_b200.transforms import Normalize

    ds = Normalize(LR_MAX, LR_MAX, "sqrt")
    scal = [Normalize(LR_MAX, LR_MAX, "linear"), Normalize(LR_MAX, LR_MAX, "sqrt")]
    coll = get_metrics(ds, scal, "val")
    assert sorted(coll.keys()) == sorted(f"val/{m}/{n}" for m in ("linear", "sqrt")
                                         for n in ("psnr", "ssim", "ms_ssim", "l1", "l2", "poisson"))
    batches = []
    for seed in (1, 2):
        lr, hr, t_lr, t_hr = count_batch(2, seed=seed, kind="dn")
        t = O.normalize_image(torch.from_numpy(hr.astype(np.float32) / t_hr), LR_MAX, "sqrt")
        g = torch.Generator().manual_seed(seed)
        p = (t * (0.9 + 0.2 * torch.rand(t.shape, generator=g)) + 0.01 * torch.rand(t.shape, generator=g)).clamp(0, 1)
        batches.append((p, t))
        coll.update(preds=p.to(dev), target=t.to(dev))
    got = {k: float(v) for k, v in coll.compute().items()}
    for mode in ("linear", "sqrt"):
        ps = [O.stretch(O.unstretch(p, "sqrt"), mode) for p, _ in batches]
        ts = [O.stretch(O.unstretch(t, "sqrt"), mode) for _, t in batches]
        n = sum(p.numel() for p in ps)
        nb = sum(p.shape[0] for p in ps)
        absum = sum(float((p - t).abs().double().sum()) for p, t in zip(ps, ts))
        sq = sum(float(((p - t).double() ** 2).sum()) for p, t in zip(ps, ts))
        tmin = min(0.0, min(float(t.min()) for t in ts))
        tmax = max(0.0, max(float(t.max()) for t in ts))
        want = {
            "l1": absum / n, "l2": sq / n,
            "psnr": 10.0 * np.log10((tmax - tmin) ** 2 / (sq / n)),
            "poisson": sum(float(torch.nn.functional.poisson_nll_loss(p, t, log_input=False)) for p, t in zip(ps, ts)) / nb,
            "ssim": sum(float(O.ssim(p, t)) * p.shape[0] for p, t in zip(ps, ts)) / nb,
            "ms_ssim": sum(float(O.ms_ssim(p, t)) * p.shape[0] for p, t in zip(ps, ts)) / nb,
        }
        for name, w in want.items():
            g = got[f"val/{mode}/{name}"]
            assert abs(g - w) <= 2e-4 * max(1.0, abs(w)), (mode, name, g, w)
    coll.reset()
    with pytest.raises(RuntimeError):
        coll.compute()
    inm = get_in_metrics(ds, scal, "val")
    assert "val/linear/in/psnr" in inm.keys()


def test_restretch_roundtrip(dev):
    from xmm_superres_denoise_b200 import ops

    x = torch.rand(3, 1, 37, 41, generator=torch.Generator().manual_seed(2))
    for a in ("linear", "sqrt", "asinh", "log"):
        for b in ("linear", "sqrt", "asinh", "log"):
            got = ops.restretch(x.to(dev), a, b).cpu()
            want = O.stretch(O.unstretch(x, a), b)
            assert torch.allclose(got, want, atol=3e-6, rtol=2e-5), (a, b, float((got - want).abs().max()))


# ------------------------------------------------------------------------------ f4: inference on FITS files
@pytest.mark.parametrize("kind", ["dn", "sr"])
def test_run_inference_on_fits_files(dev, tmp_path, golden_dir, kind):
    """FITS in -> FITS out through utils.run_inference_on_file (reference lines 101-200): the written prediction
    equals denormalize_hr(generator(normalize_lr(pad(counts * mask) / exposure))) computed by the oracle, the input
    copy is the padded rate image, and the WCS keywords follow filehandling.py:197-225."""
    import os

    from xmm_superres_denoise_b200.utils import fits_io
    from xmm_superres_denoise_b200.utils.run_inference_on_file import build_generator, infer_files

    g = np.load(os.path.join(golden_dir, "config1_dn_example.npz"))
    counts = g["counts"].astype(np.int32)  # one real 20 ks example image (411 x 403)
    mask = detector_mask(1)
    files = []
    for i in range(3):
        p = str(tmp_path / f"obs{i}_image.fits")
        fits_io.write_primary(p, np.roll(counts, 7 * i, axis=1), {"EXPOSURE": 20000.0, "CRPIX1": 1.0, "CRPIX2": 1.0,
                                                                   "CDELT1": 80.0, "CDELT2": 80.0, "PA_PNT": 100.0})
        files.append(p)
    cfg = {"lr_res": 416, "hr_res": 832 if kind == "sr" else 416, "dataset_lr_res": 416, "data_scaling": "sqrt",
           "lr_max": LR_MAX, "hr_max": HR_MAX if kind == "sr" else LR_MAX, "hr_exp": 100 if kind == "sr" else 50,
           "det_mask": True}
    # random init: seeds whose output on this image is not clamped to zero (most are: the comparison would then rest on
    # a handful of pixels)
    torch.manual_seed(3 if kind == "dn" else 1)
    with pytest.warns(UserWarning, match="randomly initialised"):
        gen = build_generator(cfg, {"filters": 32, "residual_blocks": 1}, None, dev)
    sd = {k: v.detach().cpu() for k, v in gen.state_dict().items()}
    res = infer_files(files, cfg, str(tmp_path / "out"), gen, det_mask=mask, batch_size=2)
    assert len(res) == 3
    for i, r in enumerate(res):
        img_in, h_in = fits_io.read_primary(r["input"])
        img_out, h_out = fits_io.read_primary(r["predict"])
        comb = O.combine_mask_pad(torch.from_numpy(np.roll(counts, 7 * i, axis=1))[None, None], None, None,
                                  torch.from_numpy(mask.astype(np.float32)), 416)
        x = O.normalize_image(comb / 20000.0, LR_MAX, "sqrt")
        want_in = O.denormalize_image(x, torch.tensor(LR_MAX), "sqrt")[0, 0].numpy()
        assert img_in.shape == (416, 416) and np.allclose(img_in, want_in, atol=1e-9, rtol=1e-5)
        with torch.no_grad():
            y = O.model_forward(x, sd, kind, 1)
        want_out = O.denormalize_image(y, torch.tensor(cfg["hr_max"]), "sqrt")[0, 0].numpy()
        assert img_out.shape == want_out.shape
        assert rel_l2(img_out, want_out) < 2e-2  # bf16 generator, de-normalised (squared) output
        assert h_in["CRPIX1"] == 7.0 and h_in["CRPIX2"] == 3.0 and h_in["EXPOSURE"] == 20000.0
        if kind == "sr":
            assert h_out["CRPIX1"] == 14.5 and h_out["CDELT1"] == 40.0 and "CD1_1" in h_out
            assert h_out["EXPOSURE"] == 100000.0


def test_model_wrapper_train_and_validation_steps(dev):
    """models.Model (the reference's LightningModule surface) on the CUDA path: a training step returns the
    differentiable loss, a validation epoch accumulates loss + metric collections and logs them by name."""
    import types

    from xmm_superres_denoise_b200.metrics import get_in_metrics, get_metrics
    from xmm_superres_denoise_b200.models import Model
    from xmm_superres_denoise_b200.transforms import Normalize
    from xmm_superres_denoise_b200.utils.loss_functions import create_loss

    cfg = types.SimpleNamespace(name="rrdb_denoise", memory_efficient=False, batch_size=2,
                                model=types.SimpleNamespace(in_channels=1, out_channels=1, filters=32, residual_blocks=1),
                                optimizer=types.SimpleNamespace(learning_rate=1e-4, betas=(0.9, 0.999)))
    ds = Normalize(LR_MAX, LR_MAX, "sqrt")
    scal = [Normalize(LR_MAX, LR_MAX, "linear")]
    loss = create_loss(None, {"l1": 0.5, "poisson": 0.5})
    m = Model(cfg, (416, 416), (416, 416), loss, get_metrics(ds, scal, "val"), None, get_in_metrics(ds, scal, "val"), None)
    m.configure_model()
    m = m.to(dev)
    lr, hr, t_lr, t_hr = count_batch(2, seed=4, kind="dn")
    x = O.normalize_image(torch.from_numpy(lr.astype(np.float32) / t_lr), LR_MAX, "sqrt").to(dev)
    y = O.normalize_image(torch.from_numpy(hr.astype(np.float32) / t_hr), LR_MAX, "sqrt").to(dev)
    opt = m.configure_optimizers()
    m.train()
    l0 = m.training_step((x, y), 0)
    l0.backward()
    opt.step()
    assert torch.isfinite(l0) and all(p.grad is not None for p in m.model.parameters())
    m.eval()
    m.on_validation_start()
    with torch.no_grad():
        m.validation_step((x, y), 0)
    m.on_validation_epoch_end()
    for k in ("val/loss", "val/linear/psnr", "val/linear/ms_ssim", "val/linear/in/l1"):
        assert k in m.logged and torch.isfinite(torch.as_tensor(m.logged[k])), k
    assert m.in_metrics is None  # logged once (model.py:136-139)


def test_inference_pipeline_overlapped_matches_serial(dev):
    """pipeline.InferencePipeline (H2D / compute / D2H on three streams, double buffered) returns, for every
    submitted batch, exactly what the serial prepare + generator + copy gives."""
    from xmm_superres_denoise_b200.data import load_and_combine_simulations
    from xmm_superres_denoise_b200.models import GeneratorRRDB_SR
    from xmm_superres_denoise_b200.pipeline import InferencePipeline
    from xmm_superres_denoise_b200.transforms import Normalize

    torch.manual_seed(0)
    model = GeneratorRRDB_SR(1, 1, 32, 1, num_upsample=1).to(dev).eval()
    norm = Normalize(LR_MAX, HR_MAX, "sqrt")
    mask = detector_mask(1)
    rng = np.random.default_rng(3)
    batches = [torch.from_numpy(rng.poisson(1.0, size=(2, 411, 403)).astype(np.int32)).pin_memory() for _ in range(5)]
    outs = [torch.empty(2, 1, 832, 832).pin_memory() for _ in range(5)]
    pipe = InferencePipeline(model, norm, 2, 411, 403, 416, det_mask=mask, exposure=20000.0)
    for c, o in zip(batches, outs):
        pipe.submit(c, o)
    pipe.synchronize()
    torch.cuda.synchronize()
    mdev = torch.from_numpy(mask).to(dev)
    for c, o in zip(batches, outs):
        with torch.no_grad():
            x = load_and_combine_simulations(416, c.to(dev), det_mask=mdev, normalizer=norm, which="lr", exposure=20000.0)
            want = torch.clamp(model(x), 0, 1).cpu()
        assert torch.equal(o, want)
