"""Shared helpers for the parity tests (tolerances are the ones BASELINE.json's north_star states)."""
from __future__ import annotations

import numpy as np
import torch

# bf16 compute: per-pixel relative L2 <= 1e-2; loss / gradient relative error <= 1e-2
REL_L2_BF16 = 1e-2
GRAD_REL = 1e-2
PSNR_DELTA_DB = 0.05


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a).double().reshape(-1).cpu()
    b = torch.as_tensor(b).double().reshape(-1).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr_db(a, b, data_range: float = 1.0) -> float:
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * np.log10(data_range ** 2 / mse)


def load_case(golden_dir: str, name: str):
    import os

    g = np.load(os.path.join(golden_dir, f"net_{name}.npz"))
    kind_i, nf, nb, seed, counts = (int(v) for v in g["meta"][:5])
    shape = tuple(int(v) for v in g["meta"][5:])
    return g, ("dn", "sr")[kind_i], nf, nb, seed, bool(counts), shape
