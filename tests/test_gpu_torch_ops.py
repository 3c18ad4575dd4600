"""GPU (B200): the TORCH_LIBRARY(xmm_b200) custom-op face (csrc/torch_ops.cpp) launches the same kernels as the ctypes
face, bit for bit, and rejects what the C ABI rejects."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from xmm_superres_denoise_b200 import _lib

    _lib.check(_lib.load().xmm_check_device())
    return torch.device("cuda:0")


def test_conv3x3_fwd_op_matches_conv2d_and_the_ctypes_call(dev):
    """torch.ops.xmm_b200.conv3x3_fwd on a dense-block layer shape (rrdb_blocks.py:37-54: cin 96 window of a 160-channel
    buffer, LeakyReLU 0.2, residual) vs F.conv2d and vs ops.conv3x3 (ctypes) -- same launcher, equal bits."""
    from xmm_superres_denoise_b200 import ops, torch_ops
    from xmm_superres_denoise_b200.engine import WeightArena, _Blob, _Segment

    ns = torch_ops.load()
    g = torch.Generator().manual_seed(11)
    b, h, w, cin, cout, kc = 2, 48, 40, 96, 32, 32
    x = torch.randn(b, h, w, 160, generator=g).to(torch.bfloat16).to(dev)
    res = torch.randn(b, h, w, 32, generator=g).to(torch.bfloat16).to(dev)
    wgt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    arena = WeightArena()
    arena.add(_Blob("c", cout, kc, cin // kc, [_Segment(wgt, cin, 0, 0, 0, 0, cin, 1.0)], bias))
    arena.add(_Blob("c.row", cout, kc, cin // kc, [_Segment(wgt, cin, 0, 0, 0, 0, cin, 1.0)], bias, tap_order=1))
    arena.ensure(dev)
    out_a = torch.zeros(b, h, w, 64, dtype=torch.bfloat16, device=dev)
    out_b = torch.zeros_like(out_a)
    ns.conv3x3_fwd(x, 32, cin, arena.ptr("c"), kc, cout, out_a, 32, lrelu=0.2, s0=0.2, r1=res, r1_coff=0, s1=1.0,
                   wblob_row=arena.ptr("c.row"))
    ops.conv3x3(x, 32, cin, arena.ptr("c"), kc, cout, out_b, 32, lrelu=0.2, s0=0.2, r1=res, r1_coff=0, s1=1.0,
                wblob_row=arena.ptr("c.row"))
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)
    assert torch.all(out_a[..., :32] == 0)
    y = F.conv2d(x[..., 32:128].float().permute(0, 3, 1, 2), wgt.to(torch.bfloat16).float(), bias, padding=1)
    y = 0.2 * F.leaky_relu(y, 0.2) + res.float().permute(0, 3, 1, 2)
    assert rel_l2(out_a[..., 32:].float().permute(0, 3, 1, 2).cpu(), y.cpu()) < 4e-3
    with pytest.raises(RuntimeError, match="xmm_b200"):
        ns.conv3x3_fwd(x, 32, 95, arena.ptr("c"), kc, cout, out_a, 32)  # cin not a multiple of kc: the C ABI's error
    with pytest.raises(RuntimeError):
        ns.conv3x3_fwd(x.float(), 32, cin, arena.ptr("c"), kc, cout, out_a, 32)  # dtype check of the shim


def test_transform_and_adam_ops(dev):
    """normalize / denormalize / image_upsample / adam_step ops vs their closed forms (transforms/normalize.py:66-92,
    imageupsample.py:10-26, torch.optim.Adam)."""
    from xmm_superres_denoise_b200 import torch_ops

    ns = torch_ops.load()
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(0, 40, (2, 1, 33, 21), generator=g, dtype=torch.int32).to(dev)
    out = torch.empty(counts.shape, dtype=torch.float32, device=dev)
    ns.normalize(counts, out, 1.0 / 20000.0, 0.0011, "sqrt")
    want = torch.sqrt(torch.clamp(counts.float() / 20000.0, 0, 0.0011) / 0.0011).clamp(0, 1)
    assert torch.allclose(out, want, atol=3e-7)
    back = torch.empty_like(out)
    ns.denormalize(out, back, torch.tensor([0.0011], device=dev), "sqrt")
    assert torch.allclose(back, torch.clamp(counts.float() / 20000.0, 0, 0.0011), atol=1e-9, rtol=2e-6)
    up = torch.empty(2, 1, 66, 42, dtype=torch.float32, device=dev)
    ns.image_upsample(out, up, 2)
    assert torch.equal(up, F.interpolate(out, scale_factor=2, mode="nearest") / 4)

    p = torch.randn(1000, generator=g).to(dev)
    grad = torch.randn(1000, generator=g).to(dev)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2, 3):
        ref.grad = grad.clone()
        opt.step()
        ns.adam_step(p, grad, m, v, 1e-3, 0.9, 0.999, 1e-8, step)
    assert torch.allclose(p, ref.data, atol=1e-6)


def test_engine_through_registered_ops_is_bit_identical(dev, monkeypatch):
    """ops.conv3x3 routed through the dispatcher (XMM_TORCH_OPS=1) gives the generator the same bits."""
    from oracle import rrdb_oracle as O
    from xmm_superres_denoise_b200 import ops
    from xmm_superres_denoise_b200.models import GeneratorRRDB_SR

    sd = O.init_state_dict("sr", 1, 1, 32, 1, 1, seed=9)
    x = torch.rand(2, 1, 48, 40, generator=torch.Generator().manual_seed(2)).to(dev)
    outs = []
    for flag in (False, True):
        monkeypatch.setattr(ops, "_TORCH_OPS", flag)
        m = GeneratorRRDB_SR(1, 1, 32, 1, num_upsample=1)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        with torch.no_grad():
            outs.append(m(x).clone())
    assert torch.equal(outs[0], outs[1])
