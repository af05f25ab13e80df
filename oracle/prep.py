"""CPU restatement of the reference's input side.  TEST INFRASTRUCTURE ONLY.

``segmap_manager.py:136-167`` resizes the decoded PIL image to the network's input size with
``Image.BICUBIC`` and ``data_generators.py:176-177`` converts it with ``image.convert('L')`` when the net is
grey.  Both are Pillow calls (a third-party dependency of the reference, unpinned in requirements.txt; 12.2
here); ``resize_bicubic`` / ``rgb_to_l`` restate Pillow's published fixed-point algorithm
(src/libImaging/Resample.c, Convert.c) in NumPy.  PINNED: tests/test_prep.py compares them with Pillow itself
on random images (up- and down-scaling on either axis), and the GPU kernels with both."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc: per output index (first input index, count, int coefficients)."""
    def bic(x):
        a = -0.5
        x = abs(x)
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    bounds, coeffs = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [bic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in k:
            ww += v
        k = [v / ww if ww != 0.0 else v for v in k]
        coeffs.append([int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k])
        bounds.append((xmin, xmax))
    return bounds, coeffs


def resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """(H,W[,C]) uint8 -> (out_h,out_w[,C]) uint8 = ``Image.fromarray(img).resize((out_w, out_h), Image.BICUBIC)``."""
    a = np.asarray(img).astype(np.int64)
    h, w = a.shape[:2]
    if out_w != w:
        b, k = bicubic_coeffs(w, out_w)
        out = np.zeros((h, out_w) + a.shape[2:], np.int64)
        for xx, ((x0, n), kk) in enumerate(zip(b, k)):
            ss = (1 << (PRECISION_BITS - 1)) + sum(a[:, x0 + j] * kk[j] for j in range(n))
            out[:, xx] = np.clip(ss >> PRECISION_BITS, 0, 255)
        a = out
    if out_h != h:
        b, k = bicubic_coeffs(h, out_h)
        out = np.zeros((out_h,) + a.shape[1:], np.int64)
        for yy, ((y0, n), kk) in enumerate(zip(b, k)):
            ss = (1 << (PRECISION_BITS - 1)) + sum(a[y0 + j] * kk[j] for j in range(n))
            out[yy] = np.clip(ss >> PRECISION_BITS, 0, 255)
        a = out
    return a.astype(np.uint8)


def rgb_to_l(img: np.ndarray) -> np.ndarray:
    """``Image.convert('L')`` of an RGB image: ITU-R 601-2 luma in 16-bit fixed point."""
    g = np.asarray(img).astype(np.uint32)
    return ((g[..., 0] * 19595 + g[..., 1] * 38470 + g[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)
