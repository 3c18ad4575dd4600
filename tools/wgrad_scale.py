"""Developer tool: one main weight-gradient role (3 taps, X 128 channels, N = 160) on XMM_WG_MAX_CTAS CTAs:
does the per-CTA MMA rate depend on how many SMs run (chip-wide limit) or not (per-SM limit)?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xmm_superres_denoise_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
b, h, w, f = 16, 416, 416, 32
x = torch.randn(b, h, w, 5 * f, device=dev).to(torch.bfloat16)
dy = (torch.randn(b, h, w, 5 * f, device=dev) * 0.1).to(torch.bfloat16)
roles = [(3, 3, 0, 2, 0, 160)]
dw = torch.zeros(f, f, 3, 3, device=dev)
dsts = [(dw, f, f, 0, f, 0, 0, 0, 1.0, 0, 0)]
for _ in range(2):
    ops.conv3x3_wgrad(x, dy, roles, dsts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv3x3_wgrad(x, dy, roles, dsts)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
n = int(os.environ.get("XMM_WG_MAX_CTAS", "148"))
mmas = b * h * w / 16 * 3 / n
print(f"ctas={n}: {ms:.3f} ms, {ms * 1e-3 * 1.7e9 / mmas:.1f} cycles/MMA per CTA at 1.7 GHz", flush=True)
