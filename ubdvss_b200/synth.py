"""Synthetic inputs of the shapes BASELINE.json names.

There is no dataset (the reference's is proprietary, README.md:9): images are uint8 grayscale
with a smooth background plus a few rotated bar / checker patches so that thresholded maps
contain real multi-pixel components; stress masks bypass the net for CC parity; training
targets are integer maps in [0, C] built from random filled quads (losses.py:36-39 format).
"""
from __future__ import annotations

import numpy as np


def _smooth(rng, h, w, cell=64):
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.uniform(60, 200, size=(gh, gw)).astype(np.float32)
    ys = np.linspace(0, gh - 1.001, h, dtype=np.float32)
    xs = np.linspace(0, gw - 1.001, w, dtype=np.float32)
    y0 = ys.astype(np.int32); x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)


def synth_images(n, h, w, seed=0, channels=1):
    """(n,h,w,channels) uint8."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, channels), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        img = _smooth(rng, h, w) + rng.normal(0, 4, size=(h, w)).astype(np.float32)
        for _ in range(int(rng.integers(1, 9))):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ph, pw = rng.uniform(h / 32, h / 5), rng.uniform(w / 32, w / 5)
            ang = rng.uniform(0, np.pi)
            u = (xx - cx) * np.cos(ang) + (yy - cy) * np.sin(ang)
            v = -(xx - cx) * np.sin(ang) + (yy - cy) * np.cos(ang)
            inside = (np.abs(u) < pw / 2) & (np.abs(v) < ph / 2)
            period = rng.uniform(3, 9)
            if rng.random() < 0.5:       # 1-D bars
                pat = (np.floor(u / period) % 2)
            else:                        # 2-D checker
                pat = (np.floor(u / period) + np.floor(v / period)) % 2
            img = np.where(inside, 255.0 * pat, img)
        g = np.clip(img, 0, 255).astype(np.uint8)
        out[i] = g[..., None]
    return out


def stress_masks(n, h, w, seed=0):
    """(n,h,w) uint8 in {0,1}: Bernoulli noise at several densities, some dilated/blurred so the
    masks contain holes, nesting, diagonal links and border-touching components."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, h, w), dtype=np.uint8)
    for i in range(n):
        kind = i % 6
        if kind < 3:
            p = (0.3, 0.5, 0.7)[kind]
            m = rng.random((h, w)) < p
        elif kind == 3:                  # blob-like: thresholded box-blurred noise
            f = rng.random((h + 8, w + 8)).astype(np.float32)
            c = np.cumsum(np.cumsum(f, 0), 1)
            s = c[8:, 8:] - c[:-8, 8:] - c[8:, :-8] + c[:-8, :-8]
            m = s[:h, :w] > np.quantile(s, 0.6)
        elif kind == 4:                  # rings with inner blobs (nesting), some touching borders
            m = np.zeros((h, w), bool)
            yy, xx = np.mgrid[0:h, 0:w]
            for _ in range(max(2, h * w // 4096)):
                cy, cx = rng.integers(0, h), rng.integers(0, w)
                r = rng.integers(3, max(4, min(h, w) // 6))
                d2 = (yy - cy) ** 2 + (xx - cx) ** 2
                m |= (d2 <= r * r) & (d2 >= (r - 2) ** 2)
                if r > 6:
                    m |= d2 <= (r // 3) ** 2
        else:                            # sparse noise, dilated 3x3
            s = rng.random((h, w)) < 0.04
            m = np.zeros((h, w), bool)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    m |= np.roll(np.roll(s, dy, 0), dx, 1)
        out[i] = m
    return out


def synth_targets(n, h, w, n_classes=0, seed=0):
    """(n,h,w,1) int32 in [0, max(n_classes,1)]: 0 = background, i>0 = class i-1 (losses.py:36-39)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, h, w, 1), dtype=np.int32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        for _ in range(int(rng.integers(1, 5))):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ph, pw = rng.uniform(h / 16, h / 3), rng.uniform(w / 16, w / 3)
            ang = rng.uniform(0, np.pi)
            u = (xx - cx) * np.cos(ang) + (yy - cy) * np.sin(ang)
            v = -(xx - cx) * np.sin(ang) + (yy - cy) * np.cos(ang)
            inside = (np.abs(u) < pw / 2) & (np.abs(v) < ph / 2)
            cls = int(rng.integers(1, max(n_classes, 1) + 1))
            out[i, ..., 0][inside] = cls
    return out


def _calibration_forward(ws, x, fml=True):
    """Pre-activations of every layer of net.py:286-313 for weight calibration (NumPy, small image only)."""
    def pad(a, t, b, l, r):
        return np.pad(a, ((t, b), (l, r), (0, 0)))

    def depthwise(a, k, stride, pads):
        ap = pad(a, *pads)
        ho, wo = (ap.shape[0] - 3) // stride + 1, (ap.shape[1] - 3) // stride + 1
        out = np.zeros((ho, wo, a.shape[2]), np.float32)
        for i in range(3):
            for j in range(3):
                out += ap[i:i + stride * (ho - 1) + 1:stride, j:j + stride * (wo - 1) + 1:stride] * k[i, j, :, 0]
        return out

    def conv(a, k, d):
        h, w = a.shape[:2]
        ap = pad(a, d, d, d, d)
        out = np.zeros((h, w, k.shape[3]), np.float32)
        for i in range(3):
            for j in range(3):
                out += ap[i * d:i * d + h, j * d:j * d + w] @ k[i, j]
        return out

    s2 = (1, 0, 1, 0) if fml else (0, 1, 0, 1)
    for li, (stride, pads) in enumerate(((2, s2), (1, (1, 1, 1, 1)), (2, s2))):
        z = depthwise(x, ws[3 * li], stride, pads) @ ws[3 * li + 1][0, 0] + ws[3 * li + 2]
        x = yield li, z
    for li, d in enumerate((1, 2, 4, 8, 16, 1)):
        z = conv(x, ws[9 + 2 * li], d) + ws[10 + 2 * li]
        x = yield 3 + li, z
    yield 9, x @ ws[21][0, 0] + ws[22]


def _calibrate(ws, grey, seed):
    """Layer-sequential unit-variance scaling (LSUV, Mishkin & Matas 2016) on one synthetic image: random Glorot
    kernels make this net's logit map spatially constant (its random-sign sums average the content away), so every
    layer's kernel is rescaled per output channel to unit spatial variance of its pre-activation and its bias
    centres a random quantile (20-80 %) at zero.  The detection logit ends with std 2 and its 0.9 quantile at 0
    (about 10 % positive pixels at pixel_threshold 0.5), class logits with std 2, mean 0."""
    rng = np.random.default_rng(seed + 7919)
    img = synth_images(1, 384, 384, seed=seed + 31, channels=1 if grey else 3)[0].astype(np.float32)
    x = (img - 127.5) / 127.5
    gen = _calibration_forward(ws, x)
    li, z = next(gen)
    while True:
        m = z.reshape(-1, z.shape[-1])
        if li < 9:
            ki, bi = (3 * li + 1, 3 * li + 2) if li < 3 else (9 + 2 * (li - 3), 10 + 2 * (li - 3))
            sd = m.std(0) + 1e-6
            ws[ki] = (ws[ki] / sd).astype(np.float32)
            zn = (m - ws[bi]) / sd
            q = np.array([np.quantile(zn[:, c], rng.uniform(0.2, 0.8)) for c in range(zn.shape[1])], np.float32)
            ws[bi] = (-q).astype(np.float32)
            li, z = gen.send(np.maximum(zn - q, 0).reshape(z.shape).astype(np.float32))
        else:
            sd = m.std(0) + 1e-6
            ws[21] = (ws[21] * (2.0 / sd)).astype(np.float32)
            zn = (m - ws[22]) * (2.0 / sd)
            off = zn.mean(0)
            off[0] = np.quantile(zn[:, 0], 0.9)
            ws[22] = (-off).astype(np.float32)
            break
    return ws


def synth_weights(n_classes=0, seed=1234, grey=True, bias_std=0.1, calibrated=False):
    """Random-init weights of the architecture (no checkpoints ship with the reference): Glorot-uniform
    kernels as Keras initialises them (net.py:226) and N(0, bias_std) biases so that the bias paths are
    exercised; the 23 arrays in ``get_weights()`` order.  ``calibrated``: rescale them layer by layer
    (``_calibrate``) so that the logit map responds to the image content - the workload of bench.py and of the
    full-size parity tests."""
    from .engine import weight_shapes
    rng = np.random.default_rng(seed)
    out = []
    for shape in weight_shapes(grey, n_classes):
        if len(shape) == 1:
            out.append((rng.standard_normal(shape) * bias_std).astype(np.float32))
        else:
            kh, kw, cin, cout = shape
            limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
            out.append(rng.uniform(-limit, limit, size=shape).astype(np.float32))
    if calibrated:
        out = _calibrate(out, grey, seed)
    return out
