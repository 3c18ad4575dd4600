"""Developer tool (GPU): one dense block (rrdb_blocks.py:37-54) at the bench's size through xmm_conv3x3_chain_bf16 in a
given mode, timed with CUDA events.  Usage: python tools/rdb_probe.py MODE [batch] [reps] [hw]
MODE: 2 layer by layer, 3 fused, 259 fused without the x4 store.  Run under `timeout`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xmm_superres_denoise_b200 import ops  # noqa: E402
from xmm_superres_denoise_b200.engine import WeightArena, _Blob, _Segment  # noqa: E402


def main():
    mode = int(sys.argv[1])
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    hw = int(sys.argv[4]) if len(sys.argv) > 4 else 416
    dev = torch.device("cuda:0")
    f = 32
    g = torch.Generator().manual_seed(1)
    ws = [(torch.randn(f, k * f, 3, 3, generator=g) * 0.05).to(dev) for k in range(1, 6)]
    bs = [(torch.randn(f, generator=g) * 0.1).to(dev) for _ in range(5)]
    arena = WeightArena()
    for k in range(1, 6):
        arena.add(_Blob(f"c{k}", f, 32, k, [_Segment(ws[k - 1], k * f, 0, 0, 0, 0, k * f, 1.0)], bs[k - 1]))
        arena.add(_Blob(f"c{k}.row", f, 32, k, [_Segment(ws[k - 1], k * f, 0, 0, 0, 0, k * f, 1.0)], bs[k - 1], tap_order=1))
    arena.ensure(dev)
    buf = torch.zeros(batch, hw, hw, 5 * f, dtype=torch.bfloat16, device=dev)
    buf[..., :f].normal_(0, 0.5)
    nxt = torch.zeros(batch, hw, hw, 5 * f, dtype=torch.bfloat16, device=dev)
    layers = [((buf, 0, k * f, arena.ptr(f"c{k}"), 32, f, buf, k * f), dict(lrelu=0.2, wblob_row=arena.ptr(f"c{k}.row")))
              for k in range(1, 5)]
    layers.append(((buf, 0, 5 * f, arena.ptr("c5"), 32, f, nxt, 0),
                   dict(s0=0.2, r1=buf, r1_coff=0, s1=1.0, wblob_row=arena.ptr("c5.row"))))

    ops.conv3x3_chain(layers, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv3x3_chain(layers, mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2.0 * 9 * 15 * f * f * batch * hw * hw
    print(f"dense block mode {mode} batch {batch} {hw}x{hw}: {ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s  "
          f"checksum {float(nxt[..., :f].float().abs().mean()):.5f} x4 {float(buf[..., 4 * f:].float().abs().mean()):.5f}",
          flush=True)


if __name__ == "__main__":
    main()
