"""Inference on FITS files (reference: xmm_superres_denoise/utils/run_inference_on_file.py:52-200 and
README_inference.md): real XMM-Newton EPIC-pn detector-coordinate count images in, WCS-aligned ``*_input_wcs`` /
``*_predict_wcs`` FITS files out, using the keys of the shipped ``models/*.yaml`` configs (``lr_res``, ``hr_res``,
``data_scaling``, ``lr_max``, ``hr_max``, ``hr_exp``, ``det_mask``).

Differences by design: files are processed in BATCHES -- raw int32 counts go to the GPU, the fused feed kernel does
mask / pad / rate / normalise, the tensor-core generator runs once per batch, and both de-normalised images come
back in one copy each; ``denormalize_*`` accepts the scalar maxima (reference issue I5); the counts -> rate division
by EXPOSURE that the reference leaves implicit (I4) is explicit.  No CPU fallback.
"""
from __future__ import annotations

import os
import warnings
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ..data import load_and_combine_simulations
from ..models import GeneratorRRDB_DN, GeneratorRRDB_SR
from ..transforms import Normalize
from . import fits_io
from .filehandling import read_yaml, write_xmm_file_to_fits_wcs


def load_fits(path) -> Dict[str, object]:
    data, header = fits_io.read_primary(str(path))
    exposure = float(header.get("EXPOSURE", header.get("ONTIME", 0.0)) or 0.0)
    return {"img": data, "header": header, "exp": exposure, "file_name": os.path.basename(str(path))}


def build_generator(dataset_config: dict, model_config: Optional[dict] = None, checkpoint=None,
                    device: Optional[torch.device] = None, seed: int = 0) -> torch.nn.Module:
    """GeneratorRRDB_SR when hr_res > lr_res else GeneratorRRDB_DN; weights from a Lightning ``.ckpt``
    (``state_dict`` with the ``model.`` prefix of models/model.py) or a plain state dict."""
    model_config = model_config or {}
    nf = int(model_config.get("filters", 32))
    nb = int(model_config.get("residual_blocks", 4))
    scale = int(dataset_config["hr_res"]) // int(dataset_config["lr_res"])
    if scale > 1:
        ups = {2: 1, 4: 2}.get(scale)
        if ups is None:
            raise ValueError(f"unsupported super-resolution factor {scale}")
        gen = GeneratorRRDB_SR(1, 1, nf, nb, num_upsample=ups)
    else:
        gen = GeneratorRRDB_DN(1, 1, nf, nb)
    if checkpoint is not None:
        sd = torch.load(str(checkpoint), map_location="cpu", weights_only=False)
        sd = sd.get("state_dict", sd)
        sd = {k[len("model."):] if k.startswith("model.") else k: v for k, v in sd.items()}
        gen.load_state_dict({k: v for k, v in sd.items() if k in gen.state_dict()}, strict=True)
    else:
        warnings.warn("no checkpoint given: running with randomly initialised weights (the reference ships none)")
    return gen.to(device or torch.device("cuda", torch.cuda.current_device())).eval()


def infer_files(fits_files: Sequence, dataset_config: dict, out_path, generator: torch.nn.Module, *,
                det_mask: Optional[np.ndarray] = None, batch_size: int = 16) -> List[Dict[str, str]]:
    """Runs the generator on every file; returns [{"input": path, "predict": path}, ...]."""
    if not torch.cuda.is_available():
        raise RuntimeError("xmm_superres_denoise_b200 inference runs on CUDA (sm_100a) only; there is no CPU path")
    dev = next(generator.parameters()).device
    lr_res = int(dataset_config.get("dataset_lr_res", dataset_config["lr_res"]))
    norm = Normalize(lr_max=float(dataset_config["lr_max"]), hr_max=float(dataset_config["hr_max"]),
                     stretch_mode=str(dataset_config.get("data_scaling", "linear")))
    mask_dev = None
    if det_mask is not None:
        mask_dev = torch.from_numpy(np.ascontiguousarray(det_mask.astype(np.uint8))).to(dev)
    results = []
    files = [Path(f) for f in fits_files]
    for f in files:
        if not f.exists():
            raise FileNotFoundError(f"File {f} not found!")
    for i0 in range(0, len(files), batch_size):
        chunk = [load_fits(f) for f in files[i0:i0 + batch_size]]
        shape = chunk[0]["img"].shape
        for c in chunk:
            if c["img"].shape != shape:
                raise ValueError(f"{c['file_name']}: shape {c['img'].shape} differs from {shape} inside one batch")
            ks = c["exp"] / 1000.0
            if c["exp"] <= 0:
                raise ValueError(f"{c['file_name']}: no EXPOSURE keyword; counts cannot be converted to a rate")
            if ks >= 25.0 or ks <= 15.0:  # run_inference_on_file.py:130-137
                warnings.warn(f"The networks were trained on 20 ks exposure images, the exposure time of "
                              f"{c['file_name']} is {ks:.2f} ks.")
        counts = torch.from_numpy(np.stack([c["img"].astype(np.int32) for c in chunk])).pin_memory()
        expo = torch.tensor([c["exp"] for c in chunk], dtype=torch.float32)
        counts_d = counts.to(dev, non_blocking=True)
        x = load_and_combine_simulations(lr_res, counts_d, det_mask=mask_dev, normalizer=norm, which="lr",
                                         exposure=expo.to(dev))
        with torch.no_grad():
            pred = generator(x)  # models/model.py:48-49
            if not getattr(generator, "output_is_clamped", False):
                pred = torch.clamp(pred, 0.0, 1.0)
        in_denorm = norm.denormalize_lr_image(x).cpu().numpy()
        out_denorm = norm.denormalize_hr_image(pred).cpu().numpy()
        res_mult = out_denorm.shape[-1] // in_denorm.shape[-1]
        hr_exp = float(dataset_config.get("hr_exp", 0.0)) * 1000.0
        for j, c in enumerate(chunk):
            stem = Path(c["file_name"]).name.split(".fits")[0]
            in_name = f"{stem}_input_wcs"
            pred_name = in_name.replace("input", "predict")
            p_in = write_xmm_file_to_fits_wcs(
                in_denorm[j, 0], out_path, c["file_name"], 1, c["exp"],
                "Input image padded and WCS aligned. Needs to be multiplied by exposure.", in_name, dict(c["header"]))
            p_out = write_xmm_file_to_fits_wcs(
                out_denorm[j, 0], out_path, c["file_name"], res_mult, hr_exp if hr_exp > 0 else c["exp"],
                "XMM RRDB model prediction. Needs to be multiplied by exposure.", pred_name, dict(c["header"]))
            results.append({"input": p_in, "predict": p_out})
    return results


def run_on_file(fits_file, checkpoint, out, run_config, det_mask_file=None, batch_size: int = 16,
                model_config: Optional[dict] = None):
    """Entry point with the reference's signature (run_inference_on_file.py:52-99, without the matplotlib plots);
    ``fits_file`` may be one path or a list of paths.  ``run_config``: a models/*.yaml path or its dict."""
    cfg = read_yaml(run_config) if not isinstance(run_config, dict) else run_config
    dataset_config = cfg.get("dataset", cfg)
    model_config = model_config or cfg.get("model")
    files = [fits_file] if isinstance(fits_file, (str, Path)) else list(fits_file)
    det_mask = None
    if dataset_config.get("det_mask") and det_mask_file is not None:
        det_mask = fits_io.read_primary(str(det_mask_file))[0]
    gen = build_generator(dataset_config, model_config, checkpoint)
    return infer_files(files, dataset_config, out, gen, det_mask=det_mask, batch_size=batch_size)
