"""Training / scoring loops of the reference (train.py:26-152, score.py) on the B200 forward.

The model's forward values and all of its gradients come from the CUDA library
(``wn_forward_train`` / ``wn_backward``: tensor-core forward that keeps its activations, data-gradient and
weight-gradient kernels; only ``precision="fp32"`` re-evaluates the torch graph for its backward).  The
VGG19 perceptual model, Adam and the metrics stay PyTorch: they are not on the
north-star path.
"""
from __future__ import annotations

import json
import warnings
from pathlib import Path
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from .metrics import psnr, ssim

TRAIN_METRICS_NAMES = ["mse", "ssim", "psnr", "perceptual_loss", "loss"]
VAL_METRICS_NAMES = ["mse", "ssim", "psnr", "perceptual_loss"]
_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


def next_run_dir(root: Path) -> Path:
    """``root/<n>`` with n = 1 + the largest all-digit subdirectory (train.py:209-221)."""
    root.mkdir(exist_ok=True)
    taken = [int(p.name) for p in root.iterdir() if p.is_dir() and p.name.isdecimal()]
    return root / str(max(taken) + 1 if taken else 0)


class PerceptualModel(nn.Module):
    """VGG19 ``features`` without the final max-pool (train.py:254-263)."""

    def __init__(self, pretrained: bool = True):
        super().__init__()
        import torchvision
        vgg = None
        if pretrained:
            try:
                vgg = torchvision.models.vgg19(weights=torchvision.models.VGG19_Weights.IMAGENET1K_V1)
            except Exception as exc:  # offline: no checkpoint can be fetched
                warnings.warn(f"VGG19 ImageNet weights unavailable ({exc}); using a seeded random init")
        if vgg is None:
            state = torch.random.get_rng_state()
            torch.manual_seed(1234)
            vgg = torchvision.models.vgg19(weights=None)
            torch.random.set_rng_state(state)
        self.model = nn.Sequential(*list(vgg.features.children())[:-1])

    def forward(self, x):
        return self.model(x)


def _normalize(x):
    mean = torch.tensor(_MEAN, device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor(_STD, device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
    return (x - mean) / std


def perceptual_loss(vgg, out, ref):
    """mean((255 * (vgg(norm(out)) - vgg(norm(ref))))^2)  (train.py:110-122)."""
    return torch.mean(torch.square(255 * (vgg(_normalize(out)) - vgg(_normalize(ref)))))


def _to_device(batch, device):
    return [batch[k].to(device, non_blocking=True) for k in ("raw", "wb", "he", "gc", "ref")]


def train_one_epoch(model, loader, optimizer, scheduler, vgg, device, log=None) -> Dict[str, float]:
    model.train()
    totals = {k: 0.0 for k in TRAIN_METRICS_NAMES}
    for idx, batch in enumerate(loader):
        raw, wb, he, gc, ref = _to_device(batch, device)
        out = model(raw, wb, he, gc)
        perc = perceptual_loss(vgg, out, ref)
        mse = torch.mean(torch.square(255 * (out - ref)))
        loss = 0.05 * perc + mse
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        scheduler.step()  # per minibatch, like the reference (train.py:133)
        with torch.no_grad():
            totals["loss"] += loss.item()
            totals["perceptual_loss"] += perc.item()
            totals["mse"] += mse.item()
            totals["ssim"] += ssim(out, ref).item()
            totals["psnr"] += psnr(out, ref, 1.0).item()
        if log is not None and idx and idx % 10 == 0:
            log(f"  batch {idx}/{len(loader)} loss {loss.item():.4g}")
    return {k: v / max(len(loader), 1) for k, v in totals.items()}


def eval_one_epoch(model, loader, vgg, device) -> Dict[str, float]:
    """Validation pass; the perceptual loss is averaged over batches (the reference logs only the
    last batch divided by the batch count, train.py:71-74)."""
    model.eval()
    totals = {k: 0.0 for k in VAL_METRICS_NAMES}
    with torch.no_grad():
        for batch in loader:
            raw, wb, he, gc, ref = _to_device(batch, device)
            out = model(raw, wb, he, gc)
            totals["perceptual_loss"] += perceptual_loss(vgg, out, ref).item()
            totals["mse"] += torch.mean(torch.square(255 * (out - ref))).item()
            totals["ssim"] += ssim(out, ref).item()
            totals["psnr"] += psnr(out, ref, 1.0).item()
    model.train()
    return {k: v / max(len(loader), 1) for k, v in totals.items()}


def save_metrics(savedir: Path, train_hist, val_hist, config: dict) -> None:
    """metrics-train.csv, metrics-val.csv, config.json (train.py:311-348)."""
    savedir.mkdir(parents=True, exist_ok=True)
    for fname, names, hist in (("metrics-train.csv", TRAIN_METRICS_NAMES, train_hist),
                               ("metrics-val.csv", VAL_METRICS_NAMES, val_hist)):
        if hist is None:
            continue
        arr = np.array([[row[n] for n in names] for row in hist], dtype=np.float64).reshape(-1, len(names))
        np.savetxt(savedir / fname, arr, fmt="%f", delimiter=",", comments="", header=",".join(names))
    with open(savedir / "config.json", "w") as f:
        json.dump(config, f, indent=4)
