#!/usr/bin/env python
"""Inference entry point (reference: README_inference.md / utils/run_inference_on_file.py):

    python inference.py --fits_file a.fits [b.fits ...] --run_config models/XMM-SuperRes_real_data_config.yaml \\
        --checkpoint model.ckpt --out out_dir [--det_mask pn_mask_500_2000_detxy_1x.ds] [--batch_size 16]

Writes ``<name>_input_wcs.fits.gz`` and ``<name>_predict_wcs.fits.gz`` per input file."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--fits_file", nargs="+", required=True)
    ap.add_argument("--run_config", required=True, help="one of the reference's models/*.yaml (or a dict with its keys)")
    ap.add_argument("--checkpoint", default=None, help="Lightning .ckpt or state-dict file (default: random init)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--det_mask", default=None, help="FITS detector mask (res/detector_mask/pn_mask_500_2000_detxy_1x.ds)")
    ap.add_argument("--batch_size", type=int, default=16)
    ap.add_argument("--filters", type=int, default=32)
    ap.add_argument("--residual_blocks", type=int, default=4)
    a = ap.parse_args()
    from xmm_superres_denoise_b200.utils.run_inference_on_file import run_on_file

    res = run_on_file(a.fits_file, a.checkpoint, a.out, a.run_config, det_mask_file=a.det_mask,
                      batch_size=a.batch_size, model_config={"filters": a.filters, "residual_blocks": a.residual_blocks})
    for r in res:
        print(r["input"], "->", r["predict"])


if __name__ == "__main__":
    main()
