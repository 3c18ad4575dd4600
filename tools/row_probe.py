"""Developer tool (GPU): one dense-block / trunk / HRconv layer at a time at the bench's sizes, through a forced kernel
form (tap_mode), timed with CUDA events.  Usage: python tools/row_probe.py CASE [tap_mode] [batch]
CASE: c32 c64 c96 c128 c160 (conv1..4: lrelu) | c5 (cin 160, r1) | c5e (cin 160, r1 + r2) | trunk (cin 32, r1) |
hr (cin 32 at 832x832, lrelu) | dg128 (mask + colsum).  Run each case under `timeout`: a deadlocked kernel never returns."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xmm_superres_denoise_b200 import ops  # noqa: E402
from xmm_superres_denoise_b200.engine import WeightArena, _Blob, _Segment  # noqa: E402


def main():
    case = sys.argv[1]
    tap_mode = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    dev = torch.device("cuda:0")
    h = w = 832 if case == "hr" else 416
    cin = {"c32": 32, "c64": 64, "c96": 96, "c128": 128, "c160": 160, "c5": 160, "c5e": 160, "trunk": 32, "hr": 32,
           "dg128": 128}[case]
    ctot = 32 if case in ("trunk", "hr") else 160
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(batch, h, w, ctot, generator=g, dtype=torch.float32) * 0.5).to(torch.bfloat16).to(dev) if batch * h * w * ctot < 2e9 \
        else torch.zeros(batch, h, w, ctot, dtype=torch.bfloat16, device=dev).normal_(0, 0.5)
    out = torch.zeros(batch, h, w, 32, dtype=torch.bfloat16, device=dev)
    side = torch.zeros(batch, h, w, 32, dtype=torch.bfloat16, device=dev).normal_(0, 0.5)
    wgt = (torch.randn(32, cin, 3, 3, generator=g) * 0.05).to(dev)
    bias = (torch.randn(32, generator=g) * 0.1).to(dev)
    arena = WeightArena()
    arena.add(_Blob("c", 32, 32, cin // 32, [_Segment(wgt, cin, 0, 0, 0, 0, cin, 1.0)], bias))
    arena.add(_Blob("c.row", 32, 32, cin // 32, [_Segment(wgt, cin, 0, 0, 0, 0, cin, 1.0)], bias, tap_order=1))
    arena.ensure(dev)
    kw = dict(tap_mode=tap_mode, wblob_row=arena.ptr("c.row"))
    if case in ("c32", "c64", "c96", "c128", "c160", "hr"):
        kw.update(lrelu=0.2)
    elif case in ("c5", "trunk"):
        kw.update(s0=0.2, r1=side, r1_coff=0, s1=1.0)
    elif case == "c5e":
        kw.update(s0=0.04, r1=side, r1_coff=0, s1=0.2, r2=side, r2_coff=0, s2=1.0)
    elif case == "dg128":
        cs = torch.zeros(32, device=dev)
        kw.update(mask=side, mask_coff=0, mask_slope=0.2, colsum=cs, colsum_scale=1.0)

    def run():
        ops.conv3x3(x, 0, cin, arena.ptr("c"), 32, 32, out, 0, **kw)

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2.0 * 9 * cin * 32 * batch * h * w
    print(f"{case:6s} tap_mode {tap_mode} batch {batch}: {ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s  "
          f"({(batch * h * w * (cin + 32) * 2) / ms / 1e6:.0f} GB/s algorithmic)  checksum {float(out.float().abs().mean()):.5f}", flush=True)


if __name__ == "__main__":
    main()
