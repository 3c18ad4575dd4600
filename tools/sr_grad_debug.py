"""Developer tool (GPU): SR single-image full-size gradient vs the oracle's autograd, per parameter tensor, for a given
loss mix.  python tools/sr_grad_debug.py [l1,poisson | l1,poisson,ms_ssim | ms_ssim]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import rrdb_oracle as O  # noqa: E402
from test_gpu_bench_dispatch import _batch, _train_step  # noqa: E402
from helpers import rel_l2  # noqa: E402

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["l1", "poisson", "ms_ssim"]
w_all = {"l1": 0.3, "poisson": 0.3, "ms_ssim": 0.4}
weights = {k: w_all[k] for k in names}
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)
sd = O.init_state_dict("sr", 1, 1, 32, 4, 1, seed=21)
x, t = _batch(4, "sr", seed=7)
step = _train_step("sr", weights, sd, dev)
st, flat = step._fwd_bwd(x[3:4].to(dev), t[3:4].to(dev))
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
out = O.model_forward(x[3:4], sdg, "sr", 1)
loss = O.composite_loss(out, t[3:4], weights, O.sc_dict_for("sqrt"))
loss.backward()
print("loss ours", float(st["total"]), "oracle", float(loss))
off = 0
worst = []
for n, p in step.model.named_parameters():
    g = flat[off:off + p.numel()].cpu().reshape(p.shape)
    off += p.numel()
    worst.append((rel_l2(g, sdg[n].grad), n, float(sdg[n].grad.norm())))
want = torch.cat([sdg[n].grad.reshape(-1) for n, _ in step.model.named_parameters()])
print("full", rel_l2(flat.cpu(), want))
for r, n, nr in sorted(worst, reverse=True)[:8]:
    print(f"  {n:34s} rel {r:.3e} |g| {nr:.3e}")
# output gradient itself
with torch.no_grad():
    pass
