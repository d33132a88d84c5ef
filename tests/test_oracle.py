"""Pin the CPU oracle: golden fixtures (always), live reference / cv2 (build container)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from conftest import golden_files, load_golden
from oracle import forward as ofw
from oracle import preprocess as opre


@pytest.mark.parametrize("path", golden_files("preprocess"), ids=os.path.basename)
def test_preprocess_matches_golden(path):
    g = load_golden(path)
    wb, gc, he = opre.transform(g["rgb"])
    assert np.array_equal(wb, g["wb"])
    assert np.array_equal(gc, g["gc"])
    assert np.array_equal(he, g["he"])


@pytest.mark.parametrize("path", golden_files("forward"), ids=os.path.basename)
def test_forward_matches_golden(path):
    g = load_golden(path)
    sd = ofw.synthetic_state_dict(int(g["weight_seed"]), float(g["gain"]))
    ins = [[], [], [], []]
    for rgb in g["rgb"]:
        wb, gc, he = opre.transform(rgb)
        for slot, arr in zip(ins, (rgb, wb, he, gc)):
            slot.append(torch.from_numpy(opre.arr2ten(arr).copy()))
    x, wb, he, gc = (torch.cat(s) for s in ins)
    out = ofw.waternet_forward(sd, x, wb, he, gc).numpy()
    ref = g["out"]
    # same arithmetic (oneDNN fp32 conv) => agreement far inside the 1e-3 bar
    assert np.max(np.abs(out - ref)) <= 1e-5 * np.max(np.abs(ref))
    assert np.array_equal(opre.ten2arr(ref), g["post"])
    # fp64 ground truth agrees with the fp32 reference output
    out64 = ofw.waternet_forward(sd, x, wb, he, gc, dtype=torch.float64).numpy()
    assert np.max(np.abs(out64 - ref)) <= 1e-5 * np.max(np.abs(ref))


def test_state_dict_spec_matches_reference_keys():
    spec = ofw.state_dict_spec()
    assert len(spec) == 34
    assert sum(int(np.prod(s)) for _, s in spec) == 1_090_668
    assert spec[0] == ("cmg.conv1.weight", (128, 12, 7, 7))
    assert spec[-1] == ("gc_refiner.conv3.bias", (3,))


def test_arr2ten_ten2arr_contract():
    rgb = ofw.synthetic_image(3, 9, 7, "noise")
    ten = opre.arr2ten(rgb)
    assert ten.shape == (1, 3, 9, 7) and ten.dtype == np.float32
    assert np.array_equal(ten[0, 1], rgb[..., 1].astype(np.float32) / np.float32(255))
    assert np.array_equal(opre.ten2arr(ten), rgb[None])  # u/255*255 truncates back to u
    weird = np.array([[[[-0.5, 0.9999, 1.7, 0.5]]]], dtype=np.float32)
    assert opre.ten2arr(weird).reshape(-1).tolist() == [0, 254, 255, 127]


# ---- live checks against the real reference / OpenCV (build container only) ----


def _load_ref_data(reference_dir):
    spec = importlib.util.spec_from_file_location("_ref_data", os.path.join(reference_dir, "waternet", "data.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("shape", [(112, 112), (113, 117), (112, 117), (115, 112), (48, 200), (270, 480)])
@pytest.mark.parametrize("kind", ["noise", "smooth"])
def test_preprocess_matches_live_reference(reference_dir, shape, kind):
    pytest.importorskip("cv2")
    ref = _load_ref_data(reference_dir)
    rgb = ofw.synthetic_image(hash((shape, kind)) % 1000, shape[0], shape[1], kind)
    wb_r, gc_r, he_r = ref.transform(rgb)
    wb, gc, he = opre.transform(rgb)
    assert np.array_equal(wb, wb_r) and np.array_equal(gc, gc_r) and np.array_equal(he, he_r)


def test_lab_conversions_match_cv2_on_a_colour_lattice():
    cv2 = pytest.importorskip("cv2")
    # every 3rd level of each channel plus the extremes: 86^3 colours (exhaustive 2^24 verified offline)
    lv = np.unique(np.concatenate([np.arange(0, 256, 3), [254, 255]])).astype(np.uint8)
    cols = np.stack(np.meshgrid(lv, lv, lv, indexing="ij"), -1).reshape(1, -1, 3)
    assert np.array_equal(opre.rgb2lab_u8(cols), cv2.cvtColor(cols, cv2.COLOR_RGB2LAB))
    assert np.array_equal(opre.lab2rgb_u8(cols), cv2.cvtColor(cols, cv2.COLOR_LAB2RGB))


@pytest.mark.parametrize("shape", [(8, 8), (16, 9), (7, 5), (64, 64), (65, 64), (100, 37)])
def test_clahe_matches_cv2(shape):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for clip in (0.1, 2.0, 40.0):
        plane = rng.integers(0, 256, shape, dtype=np.uint8)
        want = cv2.createCLAHE(clipLimit=clip, tileGridSize=(8, 8)).apply(plane)
        assert np.array_equal(opre.clahe_apply(plane, clip, 8), want)


def test_resize_restatement_matches_cv2():
    """oracle.preprocess.resize_linear_u8 == cv2.resize(img, (w, h)) (default INTER_LINEAR) bit for bit: the
    training dataset's resize (training_utils.py:94-103) is third-party arithmetic like CLAHE."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    cases = [(300, 400, 112, 112), (224, 224, 112, 112), (112, 112, 112, 112), (57, 91, 112, 112), (113, 225, 112, 112),
             (641, 480, 320, 240), (1, 1, 8, 8), (2, 3, 112, 112), (700, 900, 256, 256), (225, 224, 112, 112)]
    for sh, sw, dh, dw in cases:
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(opre.resize_linear_u8(src, (dw, dh)), cv2.resize(src, (dw, dh))), (sh, sw, dh, dw)


def test_grayscale_white_balance_matches_live_reference(reference_dir):
    """The 2-D branch of white_balance_transform (data.py:30-36), including its uint8 truncation of the quantiles."""
    spec = importlib.util.spec_from_file_location("_ref_data_gray", os.path.join(reference_dir, "waternet", "data.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    for shape in [(40, 56), (7, 9), (33, 17), (112, 112)]:
        for k in range(2):
            g = rng.integers(0, 256, shape, dtype=np.uint8) if k == 0 else (rng.random(shape) * 90 + 40).astype(np.uint8)
            assert np.array_equal(mod.white_balance_transform(g.copy()), opre.white_balance_transform(g)), (shape, k)
