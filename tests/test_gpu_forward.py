"""GPU (B200): parity of the CUDA forward path, called through the C ABI / drop-in classes, against
the oracle and the reference-generated golden fixtures.  Tolerance for the bf16 tensor-core path is
BASELINE.json's: per-pixel relative L2 <= 1e-2, PSNR delta <= 0.05 dB; fp32 elementwise kernels are
held to a few ulp."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rrdb_oracle as O
from oracle.make_golden import LR_MAX, counts_like_input, det_input
from oracle.synthetic import count_batch, detector_mask, pad_to

from helpers import REL_L2_BF16, load_case, psnr_db, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from xmm_superres_denoise_b200 import _lib

    _lib.check(_lib.load().xmm_check_device())
    return torch.device("cuda:0")


def _model(kind, nf, nb, sd, dev):
    from xmm_superres_denoise_b200.models import GeneratorRRDB_DN, GeneratorRRDB_SR

    m = GeneratorRRDB_DN(1, 1, nf, nb) if kind == "dn" else GeneratorRRDB_SR(1, 1, nf, nb, num_upsample=1)
    m.load_state_dict(sd)
    return m.to(dev).eval()


# ------------------------------------------------------------------------------ conv kernel
@pytest.mark.parametrize("cin,in_coff,in_ctot,kc,cout", [(32, 0, 160, 32, 32), (96, 32, 160, 32, 32),
                                                         (160, 0, 160, 32, 32), (64, 64, 320, 64, 64)])
def test_conv3x3_matches_torch_conv2d(dev, cin, in_coff, in_ctot, kc, cout):
    from xmm_superres_denoise_b200 import ops
    from xmm_superres_denoise_b200.engine import WeightArena, _Blob, _Segment

    g = torch.Generator().manual_seed(cin + cout)
    b, h, w = 2, 37, 29  # ragged against the 16x8 tile
    x = torch.randn(b, h, w, in_ctot, generator=g).to(torch.bfloat16)
    wgt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    bias = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(b, h, w, cout, generator=g).to(torch.bfloat16)
    arena = WeightArena()
    wd, bd = wgt.to(dev), bias.to(dev)
    arena.add(_Blob("c", cout, kc, cin // kc, [_Segment(wd, cin, 0, 0, 0, 0, cin, 1.0)], bd))
    arena.ensure(dev)
    out = torch.full((b, h, w, cout + 32), 3.0, dtype=torch.bfloat16, device=dev)
    xd, rd = x.to(dev), res.to(dev)
    ops.conv3x3(xd, in_coff, cin, arena.ptr("c"), kc, cout, out, 32, lrelu=0.2, s0=0.2, r1=rd, r1_coff=0, s1=1.0)
    torch.cuda.synchronize()
    xin = x[..., in_coff:in_coff + cin].float().permute(0, 3, 1, 2)
    want = F.leaky_relu(F.conv2d(xin, wgt.to(torch.bfloat16).float(), bias, padding=1), 0.2) * 0.2 \
        + res.float().permute(0, 3, 1, 2)
    got = out[..., 32:].float().permute(0, 3, 1, 2).cpu()
    assert rel_l2(got, want) < 4e-3  # bf16 output rounding only
    assert torch.all(out[..., :32] == 3.0)  # channels outside the window untouched


def test_conv3x3_rejects_bad_arguments(dev):
    from xmm_superres_denoise_b200 import ops

    x = torch.zeros(1, 16, 16, 32, dtype=torch.bfloat16, device=dev)
    with pytest.raises(RuntimeError, match="not a multiple"):
        ops.conv3x3(x, 0, 24, x.data_ptr(), 32, 32, x, 0)
    with pytest.raises(RuntimeError, match="no kernel"):
        ops.conv3x3(x, 0, 32, x.data_ptr(), 32, 96, torch.zeros(1, 16, 16, 96, dtype=torch.bfloat16, device=dev), 0)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.conv3x3(x.cpu(), 0, 32, x.data_ptr(), 32, 32, x, 0)


# ------------------------------------------------------------------------------ generators
@pytest.mark.parametrize("name", ["dn_f32_nb1_rand", "sr_f32_nb1_rand", "dn_f32_nb2_counts", "sr_f32_nb2_counts"])
def test_generator_forward_matches_reference_golden(dev, golden_dir, name):
    g, kind, nf, nb, seed, counts, shape = load_case(golden_dir, name)
    sd = O.init_state_dict(kind, 1, 1, nf, nb, 1, seed=seed)
    x = (counts_like_input if counts else det_input)(shape, seed + 17)
    with torch.no_grad():
        got = torch.clamp(_model(kind, nf, nb, sd, dev)(x.to(dev)), 0, 1).cpu()
    assert got.shape == g["out"].shape and got.dtype == torch.float32
    assert rel_l2(got, g["out"]) < REL_L2_BF16
    assert float(got.min()) >= 0 and float(got.max()) <= 1


@pytest.mark.parametrize("kind", ["dn", "sr"])
def test_generator_default_size_matches_oracle(dev, kind):
    """F=32, nb=4 at the reference's 416x416 input (res/configs/models.toml, baseline_config.toml:36)."""
    sd = O.init_state_dict(kind, 1, 1, 32, 4, 1, seed=21)
    lr, _, t_lr, _ = count_batch(2, seed=3, kind=kind)
    x = O.normalize_image(torch.from_numpy(lr.astype(np.float32) / t_lr), LR_MAX, "sqrt")
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        want = O.model_forward(x, sd, kind, 1)
        got = torch.clamp(_model(kind, 32, 4, sd, dev)(x.to(dev)), 0, 1).cpu()
    assert got.shape == want.shape
    r = rel_l2(got, want)
    print(f"{kind} 416x416 nb=4 rel-L2 = {r:.3e}, PSNR(got,want) = {psnr_db(got, want):.1f} dB")
    assert r < REL_L2_BF16


def test_config1_example_image(dev, golden_dir):
    """BASELINE.json config 1: DeNoise on one real example image vs the reference's fp32 CPU output."""
    from xmm_superres_denoise_b200.transforms import Normalize

    g = np.load(os.path.join(golden_dir, "config1_dn_example.npz"))
    counts = torch.from_numpy(pad_to(g["counts"].astype(np.int32), 416)).to(dev)
    norm = Normalize(lr_max=LR_MAX, hr_max=LR_MAX, stretch_mode="sqrt")
    lr = norm.normalize_counts(counts[None], norm.lr_max, exposure=float(g["exposure"]))
    np.testing.assert_allclose(lr[0, 176:240, 176:240].cpu().numpy(), g["lr_crop"], rtol=0, atol=2e-7)
    sd = O.init_state_dict("dn", 1, 1, 32, 4, seed=int(g["seed"]))
    with torch.no_grad():
        out = torch.clamp(_model("dn", 32, 4, sd, dev)(lr[None]), 0, 1)[0, 0].cpu().numpy()
    assert rel_l2(out[176:240, 176:240], g["out_crop"]) < REL_L2_BF16
    assert rel_l2(out.reshape(26, 16, 26, 16).mean(axis=(1, 3)), g["out_blocks"]) < REL_L2_BF16
    assert abs(out.mean() - g["out_stats"][0]) < 2e-3


def test_batch_invariance_and_repack_on_weight_change(dev):
    """Size-independent properties: an image's output does not depend on its batch neighbours
    (bit-exact), and an in-place parameter update is picked up by the next forward."""
    sd = O.init_state_dict("sr", 1, 1, 32, 1, 1, seed=15)
    m = _model("sr", 32, 1, sd, dev)
    x = torch.rand(5, 1, 64, 48, device=dev)
    with torch.no_grad():
        full = m(x).clone()
        one = m(x[3:4]).clone()
        assert torch.equal(full[3:4], one)
        m.conv_last.bias.add_(0.05)
        shifted = m(x[3:4])
    inner = (one > 0.01) & (one < 0.9)
    assert float((shifted - one)[inner].mean()) == pytest.approx(0.05, abs=1e-3)


def test_standalone_rrdb_block(dev):
    from xmm_superres_denoise_b200.models.modules import RRDB

    torch.manual_seed(3)
    blk = RRDB(32, 32).to(dev)
    sd = {f"rrdb.0.{k}": v.detach().cpu() for k, v in blk.state_dict().items()}
    x = torch.randn(1, 32, 24, 40)
    with torch.no_grad():
        got = blk(x.to(dev)).cpu()
    want = O.rrdb_forward(x, sd, "rrdb.0")
    assert rel_l2(got, want) < REL_L2_BF16
    with pytest.raises(NotImplementedError):
        blk(x.to(dev).requires_grad_(True))


# ------------------------------------------------------------------------------ transforms
def test_normalize_matches_reference_golden(dev, golden_dir):
    from xmm_superres_denoise_b200.transforms import Normalize

    g = np.load(os.path.join(golden_dir, "normalize.npz"))
    vals = torch.from_numpy(g["vals"]).to(dev)
    for mode in ("linear", "sqrt", "asinh", "log"):
        n = Normalize(lr_max=LR_MAX, hr_max=0.0005584, stretch_mode=mode)
        keep = vals.clone()
        np.testing.assert_allclose(n.normalize_lr_image(vals).cpu().numpy(), g[f"lr.{mode}"], rtol=0, atol=3e-7)
        assert torch.equal(vals, keep)
        np.testing.assert_allclose(n.normalize_hr_image(vals).cpu().numpy(), g[f"hr.{mode}"], rtol=0, atol=3e-7)
        np.testing.assert_allclose(n.normalize_image(vals.abs(), torch.tensor(0.0)).cpu().numpy(), g[f"dynmax.{mode}"],
                                   rtol=0, atol=3e-7)
        unit = torch.linspace(0, 1, 33).reshape(1, 1, 3, 11).to(dev)
        np.testing.assert_allclose(n.denormalize_image(unit, torch.tensor([LR_MAX])).cpu().numpy(), g[f"denorm.{mode}"],
                                   rtol=2e-6, atol=1e-10)
        np.testing.assert_allclose(n.denormalize_lr_image(unit).cpu().numpy(), g[f"denorm.{mode}"], rtol=2e-6, atol=1e-10)


def test_normalize_counts_mask_and_ragged_sizes(dev):
    from xmm_superres_denoise_b200.transforms import Normalize

    n = Normalize(lr_max=LR_MAX, hr_max=LR_MAX, stretch_mode="sqrt")
    mask = torch.from_numpy(pad_to(detector_mask(1), 416))
    lr, _, t_lr, _ = count_batch(3, seed=9, kind="dn")
    counts = torch.from_numpy(lr)
    want = O.normalize_image(counts.float() / t_lr * mask, LR_MAX, "sqrt")
    got = n.normalize_counts(counts.to(dev), n.lr_max, exposure=t_lr, det_mask=mask.to(dev)).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-7)
    for size in (0, 1, 3, 5, 1023):  # empty and ragged (not a multiple of the 4-wide vector)
        v = torch.rand(size) * 3e-3
        got = n.normalize_lr_image(v.to(dev)).cpu()
        np.testing.assert_allclose(got.numpy(), O.normalize_image(v, LR_MAX, "sqrt").numpy(), rtol=0, atol=2e-7)


def test_image_upsample_matches_reference_golden(dev, golden_dir):
    from xmm_superres_denoise_b200.transforms import ImageUpsample

    g = np.load(os.path.join(golden_dir, "imageupsample.npz"))
    x = torch.from_numpy(g["x"]).to(dev)
    np.testing.assert_array_equal(ImageUpsample(2)(x).cpu().numpy(), g["up2"])
    np.testing.assert_array_equal(ImageUpsample(3)(x[0]).cpu().numpy(), g["up3_single"])
