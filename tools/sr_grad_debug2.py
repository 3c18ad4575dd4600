"""Developer tool (GPU): SR single full-size image, network gradient under a fixed upstream field vs oracle autograd,
per tensor.  Env XMM_ROW / XMM_RDB select kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import rrdb_oracle as O  # noqa: E402
from test_gpu_bench_dispatch import _batch, _train_step  # noqa: E402
from helpers import rel_l2  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "sr"
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)
sd = O.init_state_dict(kind, 1, 1, 32, 4, 1, seed=21)
x, t = _batch(4, kind, seed=7)
step = _train_step(kind, {"l1": 0.5, "poisson": 0.5}, sd, dev)
eng = step.engine
xd = x[3:4].to(dev)
out, bufs = eng.forward_train(xd)
hh, ww = out.shape[2], out.shape[3]
yy, xx = torch.meshgrid(torch.arange(hh, dtype=torch.float32), torch.arange(ww, dtype=torch.float32), indexing="ij")
g = (1.0 + 0.5 * torch.sin(yy / 37.0) * torch.cos(xx / 23.0) + 0.2 * torch.sin((xx + 2 * yy) / 7.0))[None, None] / (hh * ww)
import torch.nn.functional as F
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
if kind == "dn":
    pre_o = O._conv(O.trunk_forward(x[3:4], sdg), sdg, "conv_last") + x[3:4]
else:
    fea = F.pixel_shuffle(F.leaky_relu(O._conv(O.trunk_forward(x[3:4], sdg), sdg, "upsampling.0"), 0.01), 2)
    pre_o = O._conv(F.leaky_relu(O._conv(fea, sdg, "HRconv"), 0.2), sdg, "conv_last")
if os.environ.get("COMMON_GATE", "1") == "1":
    bufs["pre"].copy_(pre_o.detach().to(dev))
eng.backward(bufs, eng.generation, xd, g.to(dev), need_x_grad=False)
flat = eng.last_flat_grad.clone().cpu()
o = torch.clamp(pre_o, 0, 1)
(o * g).sum().backward()
print("out rel", rel_l2(out.cpu(), o.detach()), "clamped frac oracle", float(((o <= 0) | (o >= 1)).float().mean()),
      "gate mismatch frac", float((((out.cpu() <= 0) | (out.cpu() >= 1)) != ((o <= 0) | (o >= 1))).float().mean()))
off = 0
rows = []
for n, p in step.model.named_parameters():
    gg = flat[off:off + p.numel()].reshape(p.shape)
    off += p.numel()
    rows.append((rel_l2(gg, sdg[n].grad), n, float(sdg[n].grad.norm())))
want = torch.cat([sdg[n].grad.reshape(-1) for n, _ in step.model.named_parameters()])
print("full", rel_l2(flat, want))
for r, n, nr in rows[:6] + rows[-10:]:
    print(f"  {n:34s} rel {r:.3e} |g| {nr:.3e}")
print("worst:")
for r, n, nr in sorted(rows, reverse=True)[:6]:
    print(f"  {n:34s} rel {r:.3e} |g| {nr:.3e}")
