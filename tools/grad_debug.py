import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import rrdb_oracle as O
from oracle.make_golden import det_input, counts_like_input
from xmm_superres_denoise_b200.models import GeneratorRRDB_DN, GeneratorRRDB_SR
kind, nf, nb, seed, shape = sys.argv[1], 32, int(sys.argv[2]), int(sys.argv[3]), (2, 1, 32, 32)
dev = torch.device("cuda:0")
sd = O.init_state_dict(kind, 1, 1, nf, nb, 1, seed=seed)
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
x = det_input(shape, seed + 17)
xo = x.clone().requires_grad_(True)
out_o = O.model_forward(xo, sdg, kind, 1)
probe = det_input(tuple(out_o.shape), seed + 29) - float(os.environ.get("PROBE_SHIFT", "0.5"))
(out_o * probe).sum().backward()
m = (GeneratorRRDB_DN(1, 1, nf, nb) if kind == "dn" else GeneratorRRDB_SR(1, 1, nf, nb, num_upsample=1))
m.load_state_dict(sd); m = m.to(dev).train()
xg = x.to(dev).requires_grad_(True)
out = torch.clamp(m(xg), 0, 1)
(out * probe.to(dev)).sum().backward()
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
print("out", rel(out.detach().cpu(), out_o.detach()))
print("grad_x", rel(xg.grad.cpu(), xo.grad))
for n, p in m.named_parameters():
    print(f"{n:32s} rel={rel(p.grad.cpu(), sdg[n].grad):.3e}  |g|={float(sdg[n].grad.norm()):.3e}")
