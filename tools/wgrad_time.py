"""Developer tool: time one dense block's weight-gradient launch (batch 16, 416x416, F=32) with the tail as nine
N=32 taps (mode 0) or as a stacked role (mode 1).  XMM_WG_STACKED_COST sets the stacked role's CTA share."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xmm_superres_denoise_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
b, h, w, f = 16, 416, 416, 32
x = torch.randn(b, h, w, 5 * f, device=dev).to(torch.bfloat16)
dy = (torch.randn(b, h, w, 5 * f, device=dev) * 0.1).to(torch.bfloat16)
for stacked in (0, 1):
    roles = [(3 * d, 3, 0, 2, 0, 160) for d in range(3)]
    roles.append((0, 3, 128, 1, 96, 96, 1) if stacked else (0, 9, 128, 1, 128, 32))
    dws = [torch.zeros(f, k * f, 3, 3, device=dev) for k in range(1, 6)]
    dsts = []
    for k in range(1, 6):
        for d in range(3):
            dsts.append((dws[k - 1], f, k * f, 0, min(k * f, 128), d, 0, (k - 1) * f, 1.0, 0, 0))
    dsts.append((dws[4], f, 5 * f, 128, 160, 3, 32 if stacked else 0, 0, 1.0, 0, 0))
    for _ in range(3):
        ops.conv3x3_wgrad(x, dy, roles, dsts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv3x3_wgrad(x, dy, roles, dsts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tf = 2 * 135 * f * f * b * h * w / ms / 1e9
    print(f"wgrad dense block stacked={stacked} cost_scale={os.environ.get('XMM_WG_STACKED_COST', 'default')}: "
          f"{ms:.3f} ms  {tf:.0f} TFLOP/s useful", flush=True)
