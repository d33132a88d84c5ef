"""CLI surface of the reference is kept: flags, run-directory numbering, metric files."""
import json
import subprocess
import sys

import numpy as np
import torch

from conftest import ROOT
from waternet_b200 import training as T
from waternet_b200.metrics import psnr, ssim
from waternet_b200.training_utils import FlipRotate, arr2ten, ten2arr


def _help(script):
    res = subprocess.run([sys.executable, script, "--help"], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    return res.stdout


def test_inference_cli_flags():
    out = _help("inference.py")
    for flag in ("--source", "--weights", "--name", "--show-split"):
        assert flag in out


def test_train_and_score_cli_flags():
    out = _help("train.py")
    for flag in ("--epochs", "--batch-size", "--height", "--width", "--weights", "--seed"):
        assert flag in out
    out = _help("score.py")
    for flag in ("--weights", "--batch-size", "--height", "--width"):
        assert flag in out


def test_run_dir_numbering_and_metric_files(tmp_path):
    root = tmp_path / "training"
    assert T.next_run_dir(root).name == "0"
    (root / "0").mkdir()
    (root / "7").mkdir()
    (root / "notes").mkdir()
    d = T.next_run_dir(root)
    assert d.name == "8"
    hist_t = [{k: float(i) for i, k in enumerate(T.TRAIN_METRICS_NAMES)}] * 2
    hist_v = [{k: float(i) for i, k in enumerate(T.VAL_METRICS_NAMES)}] * 2
    T.save_metrics(d, hist_t, hist_v, {"epochs": 2, "batch_size": 16, "im_height": 112, "im_width": 112, "weights": None})
    assert (d / "metrics-train.csv").read_text().splitlines()[0] == "mse,ssim,psnr,perceptual_loss,loss"
    assert (d / "metrics-val.csv").read_text().splitlines()[0] == "mse,ssim,psnr,perceptual_loss"
    assert json.loads((d / "config.json").read_text())["epochs"] == 2


def test_training_utils_layout_contract():
    rgb = np.random.default_rng(0).integers(0, 256, (5, 7, 3), dtype=np.uint8)
    t = arr2ten(rgb)
    assert t.shape == (3, 5, 7) and t.dtype == torch.float32  # no batch dim added here (training_utils.py:11-24)
    assert np.array_equal(ten2arr(t), rgb)
    batch = arr2ten(rgb[None])
    assert batch.shape == (1, 3, 5, 7) and np.array_equal(ten2arr(batch), rgb[None])
    aug = FlipRotate(seed=0)(image=rgb, mask=rgb.copy())
    assert np.array_equal(aug["image"], aug["mask"]) and aug["image"].size == rgb.size


def test_metrics_basic_properties():
    torch.manual_seed(0)
    a = torch.rand(2, 3, 40, 40)
    assert abs(ssim(a, a).item() - 1.0) < 1e-6
    b = (a + 0.1 * torch.randn_like(a)).clamp(0, 1)
    assert 0 < ssim(a, b).item() < 1
    assert abs(psnr(a, b).item() - (10 * torch.log10(1 / torch.mean((a - b) ** 2))).item()) < 1e-5
