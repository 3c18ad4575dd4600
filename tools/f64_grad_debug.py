"""Developer tool (GPU): per-tensor gradient error of the 64-filter generators vs oracle autograd (nb = 1, 2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import rrdb_oracle as O  # noqa: E402
from oracle.make_golden import LR_MAX  # noqa: E402
from oracle.synthetic import count_batch  # noqa: E402
from helpers import rel_l2  # noqa: E402
from xmm_superres_denoise_b200.models import GeneratorRRDB_DN, GeneratorRRDB_SR  # noqa: E402

dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)
for kind in ("dn", "sr"):
    for nb in (1, 2):
        nf = 64
        sd = O.init_state_dict(kind, 1, 1, nf, nb, 1, seed=25)
        lr, hr, t_lr, t_hr = count_batch(2, seed=5, kind=kind)
        x = O.normalize_image(torch.from_numpy(lr[:, :, 160:208, 168:208].astype(np.float32) / t_lr), LR_MAX, "sqrt")
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        want_out = O.model_forward(x, sdg, kind, 1)
        target = (want_out.detach() * 0.7 + 0.05).clamp(0, 1)
        ((want_out - target).abs().mean() + ((want_out - target) ** 2).mean()).backward()
        m = (GeneratorRRDB_DN(1, 1, nf, nb) if kind == "dn" else GeneratorRRDB_SR(1, 1, nf, nb, num_upsample=1))
        m.load_state_dict(sd)
        m = m.to(dev).train()
        out = torch.clamp(m(x.to(dev)), 0, 1)
        t = target.to(dev)
        ((out - t).abs().mean() + ((out - t) ** 2).mean()).backward()
        per = sorted(((rel_l2(p.grad.cpu(), sdg[n].grad), n, float(sdg[n].grad.norm())) for n, p in m.named_parameters()),
                     reverse=True)
        tot = float(torch.cat([sdg[n].grad.reshape(-1) for n, _ in m.named_parameters()]).norm())
        print(kind, "nb", nb, "total |g|", f"{tot:.3e}", "worst:", [(n, f"{r:.2e}", f"{g:.1e}") for r, n, g in per[:4]])
