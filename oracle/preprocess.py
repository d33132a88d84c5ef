"""numpy restatement of the reference preprocess (TEST INFRASTRUCTURE ONLY).

Follows ``/root/reference/waternet/data.py`` function by function.  Where the
reference delegates to OpenCV (``data.py:69,71-72,76``: ``cv2.cvtColor`` RGB2LAB /
LAB2RGB on 8-bit images and ``cv2.createCLAHE(0.1,(8,8)).apply``) the published
OpenCV 4.x algorithm is restated here (opencv ``imgproc/src/color_lab.cpp``
``RGB2Lab_b`` / ``Lab2RGBinteger`` and ``imgproc/src/clahe.cpp``); the reference
does not pin an OpenCV version (``requirements.txt:4`` is a bare ``opencv``), the
de-facto pin is opencv-python-headless 4.13.0.92 of the build image, against
which ``tests/test_oracle.py`` checks this file (exhaustively for the colour
conversions) and from which ``tests/golden`` was generated.

Everything is uint8 HWC in / uint8 HWC out, like the reference.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# white balance  (data.py:6-58)
# --------------------------------------------------------------------------


def white_balance_transform(im_rgb: np.ndarray) -> np.ndarray:
    """"Simplest colour balance", RGB branch of ``data.py:14-27,37-58``.

    Per channel: saturation level 0.005 * (largest channel sum / this channel's
    sum) (``:15-23``), ``np.quantile`` of the float64 channel at
    [sat, 1-sat] (``:39-41``), clip to the two quantiles (``:42-45``), stretch
    ``(v - min) * 255 / (max - min)`` (``:46-48``) and a truncating uint8 cast
    (``:58``).
    """
    if im_rgb.ndim == 2:
        # grayscale branch, data.py:30-36: fixed saturation levels, one channel.  Unlike the RGB branch the flat
        # array stays uint8 (`np.reshape`, :36), so the clipping assignments `temp[temp < q] = q` (:43-44) store the
        # TRUNCATED quantiles: the stretch runs between floor(lo) and floor(hi)
        flat = im_rgb.reshape(-1)
        lo, hi = np.quantile(flat, [0.001, 1 - 0.005])
        chan = flat.astype(np.float64)
        chan[flat < lo] = np.floor(lo)
        chan[flat > hi] = np.floor(hi)
        bottom, top = chan.min(), chan.max()
        return ((chan - bottom) * 255 / (top - bottom)).reshape(im_rgb.shape).astype(np.uint8)
    if im_rgb.ndim != 3 or im_rgb.shape[2] != 3:
        raise ValueError("white_balance_transform expects an HWC RGB (or HW grayscale) uint8 image")
    h, w, _ = im_rgb.shape
    sums = [np.sum(im_rgb[:, :, c], axis=None) for c in range(3)]
    biggest = max(sums)
    ratio = np.array([biggest / s for s in sums])
    sat_lo = 0.005 * ratio
    sat_hi = 0.005 * ratio
    out = np.empty((h, w, 3), dtype=np.float64)
    for c in range(3):
        chan = im_rgb[:, :, c].reshape(-1).astype(np.float64)
        lo, hi = np.quantile(chan, [sat_lo[c], 1 - sat_hi[c]])
        np.clip(chan, lo, hi, out=chan)
        bottom = chan.min()
        top = chan.max()
        out[:, :, c] = ((chan - bottom) * 255 / (top - bottom)).reshape(h, w)
    return out.astype(np.uint8)


# --------------------------------------------------------------------------
# gamma  (data.py:61-65)
# --------------------------------------------------------------------------


def gamma_correction(im: np.ndarray) -> np.ndarray:
    """``uint8(clip(255 * (im/255) ** 0.7, 0, 255))`` in float64 (``data.py:62-64``)."""
    g = np.power(im / 255, 0.7)
    g = np.clip(255 * g, 0, 255)
    return g.astype(np.uint8)


# --------------------------------------------------------------------------
# OpenCV 8-bit RGB <-> Lab (what cv2.cvtColor does at data.py:69 and :76)
# --------------------------------------------------------------------------

_GAMMA_SHIFT = 3
_LAB_SHIFT = 12
_LAB_SHIFT2 = 15
_BASE = 1 << 14
_MIN_AB = -8145
_AB_TAB_SIZE = _BASE * 9 // 4  # 36864

# forward matrix, RGB order, scaled by 1<<12 and divided by the D65 white point
_FWD = np.array([[1777, 1541, 778], [871, 2929, 296], [73, 448, 3575]], dtype=np.int64)
# inverse matrix (XYZ->RGB) times the white point, scaled by 1<<12
_INV = np.array([[12615, -6296, -2223], [-3773, 7684, 185], [217, -836, 4715]], dtype=np.int64)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _build_lab_tables():
    u = np.arange(256, dtype=np.float64) / 255.0
    lin = np.where(u <= 0.04045, u / 12.92, np.power((u + 0.055) / 1.055, 2.4))
    gtab = np.rint(255.0 * (1 << _GAMMA_SHIFT) * lin).astype(np.int64)

    i = np.arange(256 * 3 // 2 * (1 << _GAMMA_SHIFT), dtype=np.float32)
    x = i / np.float32(255.0 * (1 << _GAMMA_SHIFT))
    f = np.where(
        x < np.float32(0.008856),
        x * np.float32(7.787) + np.float32(0.13793103448275862),
        np.cbrt(x.astype(np.float64)).astype(np.float32),  # correctly rounded; SIMD-independent
    ).astype(np.float32)
    ctab = np.rint(np.float32(1 << _LAB_SHIFT2) * f).astype(np.int64)
    # OpenCV fills LabCbrtTab_b with its own cube-root approximation; at these two arguments
    # 32768*f sits on a .5 tie and OpenCV's value is one ulp low.  324 is reachable from 8-bit RGB
    # (cv2 4.13 checked over all 2^24 colours), 2079 is not.
    ctab[324] = 17745
    ctab[2079] = 32975

    lvl = np.arange(256, dtype=np.float32)
    li = lvl * np.float32(100.0) / np.float32(255.0)
    y_lo = li / np.float32(903.3)
    fy_lo = np.float32(7.787) * y_lo + np.float32(16.0) / np.float32(116.0)
    fy_hi = (li + np.float32(16.0)) / np.float32(116.0)
    y_hi = fy_hi * fy_hi * fy_hi
    small = li <= np.float32(8.0)
    ytab = np.rint(np.where(small, y_lo, y_hi).astype(np.float32) * np.float32(_BASE)).astype(np.int64)
    fytab = np.rint(np.where(small, fy_lo, fy_hi).astype(np.float32) * np.float32(_BASE)).astype(np.int64)

    k = np.arange(4096, dtype=np.float64) / 4096.0
    srgb = np.where(k <= 0.0031308, 12.92 * k, 1.055 * np.power(k, 1.0 / 2.4) - 0.055)
    igtab = np.clip(np.rint(255.0 * srgb), 0, 255).astype(np.int64)
    return gtab, ctab, ytab, fytab, igtab


_GTAB, _CTAB, _YTAB, _FYTAB, _IGTAB = _build_lab_tables()


def _ctrunc_div(n, d):
    """C-style (truncate toward zero) integer division of an int64 array."""
    q = np.abs(n) // d
    return np.where(n < 0, -q, q)


def ab_to_xz(t):
    """OpenCV's ``abToXZ_b`` table as arithmetic, for ``t`` = fy +/- a,b term.

    ``t <= 3390 ? t*108/841 - 290 : (t*t/BASE)*t/BASE`` in C integer arithmetic.
    """
    t = np.asarray(t, dtype=np.int64)
    lin = _ctrunc_div(t * 108, 841) - 290
    cub = _ctrunc_div(_ctrunc_div(t * t, _BASE) * t, _BASE)
    return np.where(t <= 3390, lin, cub)


def rgb2lab_u8(rgb: np.ndarray) -> np.ndarray:
    """8-bit ``COLOR_RGB2LAB`` (OpenCV ``RGB2Lab_b``), integer fixed point."""
    r = _GTAB[rgb[..., 0]]
    g = _GTAB[rgb[..., 1]]
    b = _GTAB[rgb[..., 2]]
    fx = _CTAB[_descale(r * _FWD[0, 0] + g * _FWD[0, 1] + b * _FWD[0, 2], _LAB_SHIFT)]
    fy = _CTAB[_descale(r * _FWD[1, 0] + g * _FWD[1, 1] + b * _FWD[1, 2], _LAB_SHIFT)]
    fz = _CTAB[_descale(r * _FWD[2, 0] + g * _FWD[2, 1] + b * _FWD[2, 2], _LAB_SHIFT)]
    l_shift = -((16 * 255 * (1 << _LAB_SHIFT2) + 50) // 100)
    l_scale = (116 * 255 + 50) // 100
    lum = _descale(l_scale * fy + l_shift, _LAB_SHIFT2)
    a = _descale(500 * (fx - fy) + 128 * (1 << _LAB_SHIFT2), _LAB_SHIFT2)
    bb = _descale(200 * (fy - fz) + 128 * (1 << _LAB_SHIFT2), _LAB_SHIFT2)
    out = np.stack([lum, a, bb], axis=-1)
    return np.clip(out, 0, 255).astype(np.uint8)


def lab2rgb_u8(lab: np.ndarray) -> np.ndarray:
    """8-bit ``COLOR_LAB2RGB`` (OpenCV ``Lab2RGBinteger``), integer fixed point."""
    lum = lab[..., 0].astype(np.int64)
    a = lab[..., 1].astype(np.int64)
    b = lab[..., 2].astype(np.int64)
    y = _YTAB[lum]
    ify = _FYTAB[lum]
    adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * _BASE // 500
    bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * _BASE // 200 + 1
    x = ab_to_xz(ify + adiv)
    z = ab_to_xz(ify - bdiv)
    chans = []
    for row in range(3):
        v = _descale(_INV[row, 0] * x + _INV[row, 1] * y + _INV[row, 2] * z, 14)
        chans.append(_IGTAB[np.clip(v, 0, 4095)])
    return np.stack(chans, axis=-1).astype(np.uint8)


# --------------------------------------------------------------------------
# CLAHE (what cv2.createCLAHE(clipLimit, (8,8)).apply does at data.py:71-72)
# --------------------------------------------------------------------------


def _round_half_even_u8(x32: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x32), 0, 255).astype(np.uint8)


def clahe_luts(plane: np.ndarray, clip_limit: float = 0.1, grid: int = 8):
    """Per-tile equalisation LUTs (OpenCV ``CLAHE_CalcLut_Body``).

    Returns ``(luts[grid, grid, 256] uint8, tile_h, tile_w)``.
    """
    h, w = plane.shape
    if h % grid == 0 and w % grid == 0:
        padded = plane
    else:
        # cv::copyMakeBorder(..., 0, ty - h%ty, 0, tx - w%tx, BORDER_REFLECT_101):
        # a full `grid` rows/cols are added to a dimension that is already divisible.
        padded = np.pad(plane, ((0, grid - h % grid), (0, grid - w % grid)), mode="reflect")
    th, tw = padded.shape[0] // grid, padded.shape[1] // grid
    area = th * tw
    clip = max(int(clip_limit * area / 256), 1) if clip_limit > 0 else 0
    scale = np.float32(255) / np.float32(area)
    luts = np.empty((grid, grid, 256), dtype=np.uint8)
    for ty in range(grid):
        for tx in range(grid):
            tile = padded[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            hist = np.bincount(tile.reshape(-1), minlength=256).astype(np.int64)
            if clip > 0:
                excess = int(np.sum(np.maximum(hist - clip, 0)))
                hist = np.minimum(hist, clip)
                batch = excess // 256
                residual = excess - batch * 256
                hist = hist + batch
                if residual != 0:
                    step = max(256 // residual, 1)
                    idx = np.arange(0, 256, step)[:residual]
                    hist[idx] += 1
            cum = np.cumsum(hist).astype(np.float32)
            luts[ty, tx] = _round_half_even_u8(cum * scale)
    return luts, th, tw


def clahe_apply(plane: np.ndarray, clip_limit: float = 0.1, grid: int = 8) -> np.ndarray:
    """CLAHE on one uint8 plane (OpenCV ``CLAHE_Interpolation_Body``), float32 blend."""
    h, w = plane.shape
    luts, th, tw = clahe_luts(plane, clip_limit, grid)
    inv_tw = np.float32(1.0) / np.float32(tw)
    inv_th = np.float32(1.0) / np.float32(th)
    xs = np.arange(w, dtype=np.float32) * inv_tw - np.float32(0.5)
    ys = np.arange(h, dtype=np.float32) * inv_th - np.float32(0.5)
    tx1 = np.floor(xs).astype(np.int64)
    ty1 = np.floor(ys).astype(np.int64)
    xa = (xs - tx1.astype(np.float32)).astype(np.float32)
    ya = (ys - ty1.astype(np.float32)).astype(np.float32)
    xa1 = (np.float32(1.0) - xa).astype(np.float32)
    ya1 = (np.float32(1.0) - ya).astype(np.float32)
    tx2 = np.minimum(tx1 + 1, grid - 1)
    ty2 = np.minimum(ty1 + 1, grid - 1)
    tx1 = np.maximum(tx1, 0)
    ty1 = np.maximum(ty1, 0)
    v = plane.astype(np.int64)
    f32 = np.float32
    l11 = luts[ty1[:, None], tx1[None, :], v].astype(f32)
    l12 = luts[ty1[:, None], tx2[None, :], v].astype(f32)
    l21 = luts[ty2[:, None], tx1[None, :], v].astype(f32)
    l22 = luts[ty2[:, None], tx2[None, :], v].astype(f32)
    top = (l11 * xa1[None, :] + l12 * xa[None, :]).astype(f32)
    bot = (l21 * xa1[None, :] + l22 * xa[None, :]).astype(f32)
    res = (top * ya1[:, None]).astype(f32) + (bot * ya[:, None]).astype(f32)
    return _round_half_even_u8(res.astype(f32))


# --------------------------------------------------------------------------
# histeq / transform  (data.py:68-90)
# --------------------------------------------------------------------------


def histeq(im_rgb: np.ndarray) -> np.ndarray:
    """RGB -> Lab, CLAHE(0.1, 8x8) on L, Lab -> RGB (``data.py:68-78``)."""
    lab = rgb2lab_u8(im_rgb)
    lab[..., 0] = clahe_apply(lab[..., 0], 0.1, 8)
    return lab2rgb_u8(lab)


def transform(rgb: np.ndarray):
    """``(wb, gc, he)`` -- note the order (``data.py:81-90``)."""
    return white_balance_transform(rgb), gamma_correction(rgb), histeq(rgb)


# --------------------------------------------------------------------------
# tensor <-> array contract (hubconf.py:8-34, training_utils.py:11-43)
# --------------------------------------------------------------------------


def arr2ten(arr: np.ndarray) -> np.ndarray:
    """uint8 (N)HWC -> float32 NCHW in [0,1] by true division (``hubconf.py:13-20``).

    Returned as a numpy array; a 3-D input gains a leading batch dim like the
    hubconf / inference.py variants.
    """
    ten = arr.astype(np.float32) / np.float32(255)
    if ten.ndim == 3:
        ten = ten[None]
    return np.transpose(ten, (0, 3, 1, 2))


def ten2arr(ten: np.ndarray) -> np.ndarray:
    """float32 NCHW -> uint8 NHWC: clip to [0,1], *255, truncate (``hubconf.py:29-33``)."""
    arr = np.clip(np.asarray(ten, dtype=np.float32), 0, 1)
    arr = (arr * 255).astype(np.uint8)
    return np.transpose(arr, (0, 2, 3, 1))


# --------------------------------------------------------------------------
# cv2.resize(img, (w, h)) on 8-bit images, default INTER_LINEAR  (training_utils.py:94-103)
# --------------------------------------------------------------------------
_RESIZE_COEF_BITS = 11  # INTER_RESIZE_COEF_BITS


def _resize_coeffs(ssize: int, dsize: int, clamp: bool):
    """Source index and the two 11-bit fixed-point weights of every destination coordinate.

    OpenCV ``imgproc/src/resize.cpp`` (``cv::resize`` -> ``resizeGeneric_``, linear branch):
    ``fx = (float)((dx + 0.5) * scale - 0.5); sx = cvFloor(fx); fx -= sx`` with ``scale = 1 / (dsize / ssize)`` in
    double; along x an index outside ``[0, ssize-1)`` is clamped with weight (1, 0) (``clamp=True``), along y only
    the ROW INDEX is clamped later and the weights keep the fractional part; the weights are
    ``saturate_cast<short>(w * 2048)`` (round half to even).
    """
    scale = 1.0 / (float(dsize) / float(ssize))
    idx = np.empty(dsize, np.int64)
    w0 = np.empty(dsize, np.int64)
    w1 = np.empty(dsize, np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
        idx[d] = s
        w0[d] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
        w1[d] = int(np.rint(f * np.float32(2048)))
    return idx, w0, w1


def resize_linear_u8(src: np.ndarray, dsize) -> np.ndarray:
    """``cv2.resize(src, dsize)`` for uint8 HWC images, ``dsize = (width, height)`` like cv2.

    Restates OpenCV 4.x ``resize.cpp``: horizontal pass ``S[x0]*a0 + S[x1]*a1`` in int32 (``HResizeLinear``),
    vertical pass ``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`` (``VResizeLinear`` for 8-bit).
    An exact 2x reduction in both directions is silently switched to INTER_AREA by cv::resize
    (``(a + b + c + d + 2) >> 2``); equal sizes are a copy.  Third-party arithmetic: the reference does not pin
    OpenCV; checked bit-exact against the build image's cv2 4.13.0 in tests/test_oracle.py.
    """
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = src.shape[:2]
    if (sw, sh) == (dw, dh):
        return src.copy()
    s = src.astype(np.int64)
    if sw == 2 * dw and sh == 2 * dh:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xi, xa0, xa1 = _resize_coeffs(sw, dw, True)
    yi, yb0, yb1 = _resize_coeffs(sh, dh, False)
    xi1 = np.minimum(xi + 1, sw - 1)
    rows = s[:, xi, :] * xa0[None, :, None] + s[:, xi1, :] * xa1[None, :, None]
    r0 = rows[np.clip(yi, 0, sh - 1)]
    r1 = rows[np.clip(yi + 1, 0, sh - 1)]
    out = (((yb0[:, None, None] * (r0 >> 4)) >> 16) + ((yb1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
