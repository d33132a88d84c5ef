"""SSIM / PSNR as the reference's training loop uses them (train.py:139-144), without torchmetrics.

``structural_similarity_index_measure(preds, target)`` defaults: 11x11 gaussian window,
sigma 1.5, k1 0.01, k2 0.03, data range inferred as max(preds.max()-preds.min(),
target.max()-target.min()), reflect padding cropped away, mean over the batch.
``peak_signal_noise_ratio(preds, target, data_range=1)``: 10*log10(1/mse) over all elements.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gaussian_window(size: int, sigma: float, channels: int, device, dtype):
    coords = torch.arange(size, device=device, dtype=dtype) - (size - 1) / 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g = g / g.sum()
    k = torch.outer(g, g)
    return k.expand(channels, 1, size, size).contiguous()


def ssim(preds: torch.Tensor, target: torch.Tensor, kernel_size: int = 11, sigma: float = 1.5,
         k1: float = 0.01, k2: float = 0.03) -> torch.Tensor:
    c = preds.shape[1]
    data_range = torch.maximum(preds.max() - preds.min(), target.max() - target.min())
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    pad = (kernel_size - 1) // 2
    win = _gaussian_window(kernel_size, sigma, c, preds.device, preds.dtype)
    p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    stack = torch.cat([p, t, p * p, t * t, p * t])
    out = F.conv2d(stack, win, groups=c)
    n = preds.shape[0]
    mu_p, mu_t, e_pp, e_tt, e_pt = (out[i * n:(i + 1) * n] for i in range(5))
    var_p, var_t, cov = e_pp - mu_p ** 2, e_tt - mu_t ** 2, e_pt - mu_p * mu_t
    s = ((2 * mu_p * mu_t + c1) * (2 * cov + c2)) / ((mu_p ** 2 + mu_t ** 2 + c1) * (var_p + var_t + c2))
    s = s[..., pad:-pad, pad:-pad] if s.shape[-1] > 2 * pad and s.shape[-2] > 2 * pad else s
    return s.reshape(n, -1).mean(-1).mean()


def psnr(preds: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    mse = torch.mean((preds - target) ** 2)
    return 10.0 * torch.log10(data_range ** 2 / mse)
