#!/usr/bin/env python
"""BASELINE.json configs[4]: throughput sweep -- RRDB depth / width (num_blocks 8..23, 32..64 filters) and inference
batch 1..256 on one B200, next to the reference's CPU path (oracle port, torch fp32, all host cores) at batch 1.

    python tools/sweep.py [--out profiles/r01_sweep.json] [--no-cpu] [--kind dn|sr]

One JSON document: a row per (filters, blocks, batch) with images/s, ms per batch, algorithmic TFLOP/s and the
fraction of the measured bf16 peak; CPU rows per (filters, blocks)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def flops_per_image(kind: str, nf: int, nb: int, pixels: int = 416 * 416) -> float:
    mac = 9 * 1 * nf + 405 * nb * nf * nf + 9 * nf * nf + (9 * nf if kind == "dn" else 72 * nf * nf + 36 * nf)
    return 2.0 * mac * pixels


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sweep.json"))
    ap.add_argument("--kind", default="dn", choices=["dn", "sr"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--blocks", default="8,16,23")
    ap.add_argument("--filters", default="32,64")
    ap.add_argument("--batches", default="1,4,16,64,256")
    a = ap.parse_args()
    from oracle import rrdb_oracle as O
    from oracle.synthetic import count_batch
    from xmm_superres_denoise_b200 import _lib
    from xmm_superres_denoise_b200.models import GeneratorRRDB_DN, GeneratorRRDB_SR

    _lib.check(_lib.load().xmm_check_device())
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    lr, _, t_lr, _ = count_batch(8, 77, a.kind)
    x8 = O.normalize_image(torch.from_numpy(lr.astype(np.float32) / t_lr), 0.0022336, "sqrt")
    rows, cpu_rows = [], []
    for nf in [int(v) for v in a.filters.split(",")]:
        for nb in [int(v) for v in a.blocks.split(",")]:
            sd = O.init_state_dict(a.kind, 1, 1, nf, nb, 1, seed=5)
            m = GeneratorRRDB_DN(1, 1, nf, nb) if a.kind == "dn" else GeneratorRRDB_SR(1, 1, nf, nb, num_upsample=1)
            m.load_state_dict(sd)
            m = m.to(dev).eval()
            fl = flops_per_image(a.kind, nf, nb)
            for b in [int(v) for v in a.batches.split(",")]:
                need = 3 * b * 416 * 416 * 5 * nf * 2 * (1.0 if a.kind == "dn" else 1.6)
                if need > 150e9:
                    rows.append({"filters": nf, "blocks": nb, "batch": b, "skipped": "activation buffers exceed 150 GB"})
                    continue
                x = x8.repeat((b + 7) // 8, 1, 1, 1)[:b].contiguous().to(dev)
                with torch.no_grad():
                    for _ in range(2):
                        m(x)
                    torch.cuda.synchronize()
                    steps = max(2, min(20, int(2.0 / max(fl * b / 6e14, 1e-4))))
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        m(x)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                tf = fl * b / (ms * 1e-3) / 1e12
                rows.append({"filters": nf, "blocks": nb, "batch": b, "images_per_s": b / (ms * 1e-3), "ms_per_batch": ms,
                             "tflops": tf, "frac_of_sustained_bf16": tf / peaks["bf16_tflops_sustained"],
                             "frac_of_burst_bf16": tf / peaks["bf16_tflops"], "steps": steps})
                print(rows[-1], flush=True)
                del x
            m._engine = None
            del m
            torch.cuda.empty_cache()
            if not a.no_cpu:
                torch.set_num_threads(os.cpu_count() or 1)
                with torch.no_grad():
                    O.model_forward(x8[:1], sd, a.kind, 1)  # warm-up (oneDNN primitive creation, page faults): the round-1
                    t0 = time.perf_counter()                # rows timed this first call and were inconsistent
                    for _ in range(2):
                        O.model_forward(x8[:2], sd, a.kind, 1)
                    dt = (time.perf_counter() - t0) / 4
                cpu_rows.append({"filters": nf, "blocks": nb, "batch": 1, "images_per_s": 1.0 / dt, "s_per_image": dt,
                                 "cores": os.cpu_count(), "kind": "port (oracle, torch %s fp32)" % torch.__version__})
                print(cpu_rows[-1], flush=True)
    doc = {"workload": f"RRDB {a.kind.upper()} inference, 416x416 synthetic count images, 1 x B200", "gpu": rows,
           "cpu_baseline": cpu_rows, "peaks": peaks}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
