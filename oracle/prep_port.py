"""Oracle (test infrastructure): CPU restatement of the reference's image pre-processing.

``letterbox_image`` (src/util.py:349-372) and ``prep_image`` (src/util.py:375-397) call
``cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_CUBIC)`` on a uint8 BGR image.  cv2 is a
third-party dependency of the reference (not vendored; this container: opencv 4.13.0), so its published
algorithm is restated here in numpy and pinned against cv2 itself by ``tests/golden/make_golden.py``:

* ``mode="opencv"``: OpenCV's own 8-bit path (modules/imgproc/src/resize.cpp): per-axis tables
  ``f = float((d + 0.5) * scale - 0.5)``, ``s = floor(f)``, Keys cubic coefficients (A = -0.75) evaluated
  in fp32, quantised to 11-bit fixed point (``cvRound(c * 2048)``); horizontal pass in int32 with the
  source index clamped to the border; vertical pass in fp32 (``beta * 2^-22``, products added from tap 3
  down to tap 0, unfused), rounded half-to-even and saturated.  Bit-identical to ``cv2.resize`` with IPP
  disabled (all 11 reference images x 3 resolutions).
* ``mode="float"``: the same tables with unquantised fp32 coefficients and a float horizontal pass.  The
  stock opencv-python wheel dispatches 8-bit cubic resizes to Intel IPP (closed source); this variant
  differs from it by one grey level on < 0.01 % of the values (and IPP itself differs from OpenCV's own
  path by one level on ~3 %).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def cubic_coefficients(x) -> np.ndarray:
    """OpenCV ``interpolateCubic``: four fp32 weights of the Keys kernel with A = -0.75 (each operation
    rounded to fp32, no fused multiply-add)."""
    x = np.float32(x)
    A, one = np.float32(-0.75), np.float32(1)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], np.float32)


def axis_tables(dst_n: int, src_n: int):
    """Per destination index: first-tap source index minus one (``s``) and the four fp32 weights."""
    scale = 1.0 / (dst_n / src_n)               # resize.cpp: inv_scale = dsize/ssize, scale = 1/inv_scale
    ofs = np.zeros(dst_n, np.int64)
    coef = np.zeros((dst_n, 4), np.float32)
    for d in range(dst_n):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(float(f)))
        ofs[d] = s
        coef[d] = cubic_coefficients(np.float32(f - np.float32(s)))
    return ofs, coef


def resize_cubic(img: np.ndarray, new_w: int, new_h: int, mode: str = "opencv") -> np.ndarray:
    """``cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_CUBIC)`` for uint8 HWC images."""
    assert img.dtype == np.uint8 and img.ndim == 3 and mode in ("opencv", "float")
    h, w, _ = img.shape
    xo, xa = axis_tables(new_w, w)
    yo, yb = axis_tables(new_h, h)
    xi = np.clip(xo[:, None] - 1 + np.arange(4)[None, :], 0, w - 1)          # border: replicate
    yi = np.clip(yo[:, None] - 1 + np.arange(4)[None, :], 0, h - 1)
    if mode == "opencv":
        ia = np.clip(np.rint(xa * np.float32(2048)), -32768, 32767).astype(np.int64)
        ib = np.clip(np.rint(yb * np.float32(2048)), -32768, 32767).astype(np.int64)
        hor = (img.astype(np.int64)[:, xi, :] * ia[None, :, :, None]).sum(2).astype(np.float32)
        b = ib.astype(np.float32) * (np.float32(1.0) / np.float32(2048 * 2048))
        rows = hor[yi]                                                         # [new_h, 4, new_w, 3]
        acc = rows[:, 3] * b[:, 3, None, None]
        for k in (2, 1, 0):
            acc = rows[:, k] * b[:, k, None, None] + acc
    else:
        src = img.astype(np.float32)
        hor = np.zeros((h, new_w, 3), np.float32)
        for k in range(4):
            hor = hor + src[:, xi[:, k], :] * xa[None, :, k, None]
        acc = np.zeros((new_h, new_w, 3), np.float32)
        for k in range(4):
            acc = acc + hor[yi[:, k]] * yb[:, k, None, None]
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def letterbox_geometry(img_w: int, img_h: int, w: int, h: int):
    """src/util.py:360-363, 368-369 -- resized size and its top-left corner on the canvas."""
    new_w = int(img_w * min(w / img_w, h / img_h))
    new_h = int(img_h * min(w / img_w, h / img_h))
    return new_w, new_h, (w - new_w) // 2, (h - new_h) // 2


def letterbox_image(img: np.ndarray, inp_dim, mode: str = "opencv") -> np.ndarray:
    """src/util.py:349-372 -- aspect-preserving cubic resize onto a canvas filled with 128."""
    w, h = inp_dim
    new_w, new_h, left, top = letterbox_geometry(img.shape[1], img.shape[0], w, h)
    canvas = np.full((h, w, 3), 128)                                          # int64 like the reference
    canvas[top:top + new_h, left:left + new_w, :] = resize_cubic(img, new_w, new_h, mode)
    return canvas


def prep_image(img: np.ndarray, inp_dim: int, mode: str = "BGR", resize: str = "opencv") -> torch.Tensor:
    """src/util.py:375-397 -- letterbox, BGR->RGB (unless mode == 'RGB'), HWC->CHW, fp32 / 255, batch axis."""
    assert mode in ("BGR", "RGB")
    canvas = letterbox_image(img, (inp_dim, inp_dim), resize)
    if mode == "RGB":
        chw = canvas.transpose((2, 0, 1)).copy()
    else:
        chw = canvas[:, :, ::-1].transpose((2, 0, 1)).copy()
    return torch.from_numpy(chw).float().div(255.0).unsqueeze(0)
