"""GPU implementation of the reference's ``waternet/data.py`` (numpy in, numpy out).

Same function names, argument meaning and return order as the reference
(``transform`` returns ``(wb, gc, he)``, ``data.py:81-90``); the arithmetic runs in
``libwaternet_b200.so`` (``wn_preprocess_u8``) and is bit-exact with the
reference's numpy/OpenCV results.  No CPU fallback: without a CUDA device these
functions raise.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .engine import get_engine


def _as_batch(im: np.ndarray) -> Tuple[np.ndarray, bool]:
    arr = np.asarray(im)
    if arr.dtype != np.uint8:
        raise TypeError(f"expected a uint8 image, got {arr.dtype}")
    if arr.ndim == 3 and arr.shape[2] == 3:
        return arr[None], True
    if arr.ndim == 4 and arr.shape[3] == 3:
        return arr, False
    raise ValueError(f"expected HWC (or NHWC) RGB uint8, got shape {arr.shape}")


def _run(im, device=None):
    batch, single = _as_batch(im)
    eng = get_engine(device)
    dev_in = torch.from_numpy(np.ascontiguousarray(batch)).to(eng.device)
    res = eng.preprocess(dev_in, tensors=False, images=True)
    out = {k: v.cpu().numpy() for k, v in res.items()}
    if single:
        out = {k: v[0] for k, v in out.items()}
    return out


def white_balance_transform(im_rgb, device=None) -> np.ndarray:
    """Simplest colour balance (reference ``data.py:6-58``): HWC RGB, or a 2-D grayscale image (``data.py:30-36``)."""
    arr = np.asarray(im_rgb)
    if arr.ndim == 2:
        if arr.dtype != np.uint8:
            raise TypeError(f"expected a uint8 image, got {arr.dtype}")
        eng = get_engine(device)
        return eng.white_balance_gray(torch.from_numpy(np.ascontiguousarray(arr[None])).to(eng.device))[0].cpu().numpy()
    return _run(im_rgb, device)["wb_u8"]


def gamma_correction(im, device=None) -> np.ndarray:
    """``uint8(clip(255 * (im/255) ** 0.7))`` (reference ``data.py:61-65``)."""
    return _run(im, device)["gc_u8"]


def histeq(im_rgb, device=None) -> np.ndarray:
    """Lab + CLAHE(0.1, 8x8) on L (reference ``data.py:68-78``)."""
    return _run(im_rgb, device)["he_u8"]


def transform(rgb, device=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``transform(rgb) -> wb, gc, he`` (reference ``data.py:81-90``; note the order)."""
    out = _run(rgb, device)
    return out["wb_u8"], out["gc_u8"], out["he_u8"]
