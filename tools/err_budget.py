"""Developer tool (CPU): where does the bf16 path's output error come from?  Re-runs the oracle forward with bf16
rounding switched on per source (weights of a layer group / stored activations of a layer group) and prints the
rel-L2 each source alone contributes for a given config -- the budget behind the F=64 SR tolerance discussion."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rrdb_oracle as O  # noqa: E402
from oracle.make_golden import LR_MAX  # noqa: E402
from oracle.synthetic import count_batch  # noqa: E402


def bf(t):
    return t.to(torch.bfloat16).float()


def group(name):
    if name.startswith("rrdb"):
        return "rdb.conv5" if name.endswith("conv5") else "rdb.conv1-4"
    return name


def run(x, sd, kind, round_w=(), round_a=()):
    orig = O._conv

    def conv(xx, sdd, name):
        g = group(name)
        w = sdd[name + ".weight"]
        sd2 = {name + ".weight": bf(w) if (g in round_w or "all" in round_w) else w, name + ".bias": sdd[name + ".bias"]}
        if g in round_a or "all" in round_a:
            xx = bf(xx)  # the layer reads bf16-stored activations
        return orig(xx, sd2, name)

    O._conv = conv
    try:
        with torch.no_grad():
            return O.model_forward(x, sd, kind, 1)
    finally:
        O._conv = orig


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "sr"
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 7
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 160
    sd = O.init_state_dict(kind, 1, 1, nf, nb, 1, seed=seed)
    lr, _, t_lr, _ = count_batch(2, seed=9, kind=kind)
    x = O.normalize_image(torch.from_numpy(lr[:, :, 100:100 + n, 100:100 + n].astype(np.float32) / t_lr), LR_MAX, "sqrt")
    torch.set_num_threads(os.cpu_count() or 1)
    want = run(x, sd, kind)
    frac_clamped = float(((want <= 0) | (want >= 1)).float().mean())
    print(f"{kind} F={nf} nb={nb}: clamped fraction of the fp32 output = {frac_clamped:.3f}, ||want|| rms = {float(want.pow(2).mean().sqrt()):.4f}")
    groups = ["conv_first", "rdb.conv1-4", "rdb.conv5", "trunk_conv", "upsampling.0", "HRconv", "conv_last"]
    rel = lambda a: float((a - want).norm() / want.norm())
    print("all weights + all activations:", f"{rel(run(x, sd, kind, ('all',), ('all',))):.3e}")
    print("all weights:", f"{rel(run(x, sd, kind, ('all',), ())):.3e}", " all activations:", f"{rel(run(x, sd, kind, (), ('all',))):.3e}")
    for g in groups:
        if kind == "dn" and g in ("upsampling.0", "HRconv"):
            continue
        print(f"  {g:14s} weights {rel(run(x, sd, kind, (g,), ())):.3e}   input activations {rel(run(x, sd, kind, (), (g,))):.3e}")


if __name__ == "__main__":
    main()
