"""torch-CPU restatement of the reference network forward (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/waternet/net.py``: ``ConfidenceMapGenerator.forward``
(``net.py:45-56``), ``Refiner.forward`` (``net.py:75-80``) and
``WaterNet.forward`` (``net.py:99-108``) as one functional evaluation over a
plain ``{key: tensor}`` state dict with the reference's 34 keys.  Floating-point
kernel => a torch reference is kept (fp32 = what the reference computes on CPU;
fp64 = ground truth used to judge both implementations).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# (name, in_channels, out_channels, kernel) -- net.py:12-42
CMG_LAYERS = [
    ("conv1", 12, 128, 7),
    ("conv2", 128, 128, 5),
    ("conv3", 128, 128, 3),
    ("conv4", 128, 64, 1),
    ("conv5", 64, 64, 7),
    ("conv6", 64, 64, 5),
    ("conv7", 64, 64, 3),
    ("conv8", 64, 3, 3),
]
# net.py:62-70
REFINER_LAYERS = [("conv1", 6, 32, 7), ("conv2", 32, 32, 5), ("conv3", 32, 3, 3)]
REFINERS = ["wb_refiner", "ce_refiner", "gc_refiner"]  # net.py:95-97


def state_dict_spec():
    """[(key, shape)] in the order ``WaterNet().state_dict()`` lists them."""
    spec = []
    for name, cin, cout, k in CMG_LAYERS:
        spec.append((f"cmg.{name}.weight", (cout, cin, k, k)))
        spec.append((f"cmg.{name}.bias", (cout,)))
    for ref in REFINERS:
        for name, cin, cout, k in REFINER_LAYERS:
            spec.append((f"{ref}.{name}.weight", (cout, cin, k, k)))
            spec.append((f"{ref}.{name}.bias", (cout,)))
    return spec


def synthetic_state_dict(seed: int = 0, gain: float = 1.0):
    """Deterministic stand-in weights (no pretrained checkpoint offline).

    U(-b, b) with b = gain / sqrt(fan_in) for weights and biases -- the bound
    torch's default Conv2d init uses -- drawn from numpy's PCG64 so the values do
    not depend on the torch version.  ``gain=3`` gives O(1) outputs that stress
    precision.
    """
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in state_dict_spec():
        if key.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            bound = gain / np.sqrt(fan_in)
            last_bound = 1.0 / np.sqrt(fan_in)
        else:
            bound = last_bound
        sd[key] = torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))
    return sd


def _conv(sd, prefix, x, k):
    w = sd[prefix + ".weight"].to(x.dtype)
    b = sd[prefix + ".bias"].to(x.dtype)
    return F.conv2d(x, w, b, stride=1, padding=k // 2)  # padding="same", odd kernels


def confidence_maps(sd, x, wb, ce, gc):
    """net.py:45-56 -- returns the (N,3,H,W) sigmoid maps (wb, ce, gc order)."""
    out = torch.cat([x, wb, ce, gc], dim=1)
    for name, _, _, k in CMG_LAYERS[:-1]:
        out = F.relu(_conv(sd, f"cmg.{name}", out, k))
    return torch.sigmoid(_conv(sd, "cmg.conv8", out, 3))


def refine(sd, which, x, xbar):
    """net.py:75-80 -- three conv+ReLU (the last conv is followed by ReLU too)."""
    out = torch.cat([x, xbar], dim=1)
    for name, _, _, k in REFINER_LAYERS:
        out = F.relu(_conv(sd, f"{which}.{name}", out, k))
    return out


def waternet_forward(sd, x, wb, ce, gc, dtype=torch.float32, return_parts=False):
    """net.py:99-108.  Inputs (N,3,H,W) in the order (raw, wb, he, gc)."""
    with torch.no_grad():
        x, wb, ce, gc = (t.detach().to("cpu", dtype).contiguous() for t in (x, wb, ce, gc))
        cm = confidence_maps(sd, x, wb, ce, gc)
        r_wb = refine(sd, "wb_refiner", x, wb)
        r_ce = refine(sd, "ce_refiner", x, ce)
        r_gc = refine(sd, "gc_refiner", x, gc)
        out = r_wb * cm[:, 0:1] + r_ce * cm[:, 1:2] + r_gc * cm[:, 2:3]
    if return_parts:
        return out, cm, (r_wb, r_ce, r_gc)
    return out


def synthetic_image(seed: int, h: int, w: int, kind: str = "noise") -> np.ndarray:
    """SURVEY.md section 8d inputs: uniform noise, or a smooth blue-green cast."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    base = rng.random((h // 8 + 3, w // 8 + 3, 3))
    ys = np.linspace(0, base.shape[0] - 1.001, h)
    xs = np.linspace(0, base.shape[1] - 1.001, w)
    y0 = ys.astype(int)
    x0 = xs.astype(int)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    img = (
        base[y0][:, x0] * (1 - fy) * (1 - fx)
        + base[y0][:, x0 + 1] * (1 - fy) * fx
        + base[y0 + 1][:, x0] * fy * (1 - fx)
        + base[y0 + 1][:, x0 + 1] * fy * fx
    )
    img = (img - img.min()) / (img.max() - img.min())
    return (img * np.array([90.0, 200.0, 230.0])).astype(np.uint8)
