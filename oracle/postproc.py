"""CPU restatement of the reference's threshold + segmap post-processing.  TEST INFRASTRUCTURE ONLY.

Two layers:

1. ``*_cv2`` functions restate the reference literally (its own OpenCV calls):
   ``model_runner.py:37-38,121-124`` (logit threshold), ``utils.py:51-60``
   (``get_contours_and_boxes``), ``segmap_manager.py:54-69`` (``postprocess``),
   ``utils.py:135-138`` (``np_softmax``).  Only mechanical adaptation: OpenCV-4 returns a
   2-tuple from ``findContours`` (the reference pins ``opencv-python<4``, ``requirements.txt:5``).
   PINNED against the reference executed in the build container and against the fixtures
   in ``tests/golden/`` (``tools/make_golden.py``).

2. ``ccl_spec`` is the cv2-free statement of the same result that the CUDA kernels are
   compared with bit-exactly: label background with 4-connectivity,
   ``outer`` = background reachable from outside the image; label ``~outer`` with
   8-connectivity => the RETR_EXTERNAL components with holes filled (= the filled masks
   ``cv2.drawContours(..., -1)`` draws, ``segmap_manager.py:64``); ``contourArea`` =
   #Q4 + #Q3/2 over 2x2 bit-quads of the filled component.  ``tests/test_oracle_postproc.py``
   proves (1) == (2) on hand cases and on random masks.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- threshold

def logit_threshold(pixel_threshold: float = 0.5) -> np.float64:
    """model_runner.py:37-38."""
    eps = 1e-9
    return -np.log(1 / np.clip(pixel_threshold, eps, 1 - eps) - 1)


def threshold_mask(det_logits, logit_thr) -> np.ndarray:
    """model_runner.py:124: ``np.where(det > thr, 1, 0)`` -> int64, strict '>'.

    The reference ran on NumPy < 1.24 (``np.bool`` in segmap_manager.py:65), where a float32
    array compared with a float64 scalar is compared in float32 (value-based casting), i.e. the
    scalar is rounded to float32 first.  NumPy 2 would compare in float64; the two differ only
    for a logit exactly equal to float32(thr), so the restatement fixes the float32 rule."""
    det = np.asarray(det_logits, dtype=np.float32)
    return np.where(det > np.float32(logit_thr), 1, 0)


def np_softmax(logits, axis=-1):
    """utils.py:135-138."""
    x = logits - np.max(logits, axis=axis, keepdims=True)
    x = np.exp(x)
    return x / np.sum(x, axis=axis, keepdims=True)


# ----------------------------------------------------------------------------- literal (cv2)

def get_contours_and_boxes_cv2(seg_map, min_area=10):
    """utils.py:51-60 (OpenCV-4 return signature)."""
    import cv2
    cnts = cv2.findContours(np.array(seg_map, dtype=np.uint8),
                            mode=cv2.RETR_EXTERNAL, method=cv2.CHAIN_APPROX_SIMPLE)[-2]
    cnts = list(filter(lambda cnt: cv2.contourArea(cnt) > min_area, cnts))
    rects = [cv2.minAreaRect(cnt) for cnt in cnts]
    boxes = [cv2.boxPoints(rect).reshape((8,)) for rect in rects]
    assert len(boxes) == len(cnts)
    return cnts, boxes


def postprocess_cv2(seg_map, seg_map_class_logits=None, scale=1, min_area_threshold=5):
    """segmap_manager.py:54-69 -> list of (bbox int[8], class_id | None)."""
    import cv2
    contours, boxes = get_contours_and_boxes_cv2(seg_map, min_area=min_area_threshold)
    boxes = [np.round(box * scale).astype(int) for box in boxes]
    if seg_map_class_logits is None:
        return [(b, None) for b in boxes]
    probs = np_softmax(seg_map_class_logits, axis=-1)
    out = []
    for bbox, cnt in zip(boxes, contours):
        mask = np.zeros(np.shape(seg_map)[:2], dtype=np.uint8)
        cv2.drawContours(mask, [cnt], -1, 1, -1)
        class_probs = probs[mask.astype(bool)].mean(axis=0)
        out.append((bbox, int(np.argmax(class_probs))))
    return out


def predict_postproc_cv2(pred_logits, pixel_threshold=0.5, classification=False, scale=4, min_area=5):
    """model_runner.py:119-134 after ``model.predict``: returns (det_mask, cls_logits, found)."""
    thr = logit_threshold(pixel_threshold)
    det = threshold_mask(pred_logits[..., :1], thr)
    cls = pred_logits[..., 1:]
    found = [postprocess_cv2(det[i], cls[i] if classification else None, scale, min_area)
             for i in range(pred_logits.shape[0])]
    return det, cls, found


# ----------------------------------------------------------------------------- cv2-free spec

_S4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)
_S8 = np.ones((3, 3), dtype=bool)


def filled_regions(mask2d) -> np.ndarray:
    """bool (h,w): foreground plus every background pixel NOT 4-connected to the outside."""
    from scipy import ndimage
    fg = np.asarray(mask2d).astype(bool)
    bg = np.pad(~fg, 1, constant_values=True)            # virtual outside frame
    lab, _ = ndimage.label(bg, structure=_S4)
    outer = (lab == lab[0, 0])[1:-1, 1:-1]
    return ~outer


def ccl_spec(mask, cls_logits=None):
    """mask (h,w[,1]) any int/bool dtype -> (labels int32 (h,w), components list).

    labels: -1 where the pixel belongs to no component (outer background), else the component
    id = raster index (y*w+x) of the component's first pixel in raster order.
    components (sorted by id): dict(label, xmin, ymin, xmax, ymax, n_pixels (foreground px),
    n_filled (px incl. holes), area_x2 (= 2*cv2.contourArea), cls_prob_sum float64[C] or None).
    """
    from scipy import ndimage
    m = np.asarray(mask)
    if m.ndim == 3:
        m = m[..., 0]
    fg = m.astype(np.uint8) != 0                           # utils.py:52 casts to uint8
    h, w = fg.shape
    filled = filled_regions(fg)
    lab, n = ndimage.label(filled, structure=_S8)
    labels = np.full((h, w), -1, dtype=np.int32)
    comps = []
    if n == 0:
        return labels, comps
    probs = None
    if cls_logits is not None and np.shape(cls_logits)[-1] > 0:
        probs = np_softmax(np.asarray(cls_logits, dtype=np.float32), axis=-1)
    flat = lab.ravel()
    idx = np.flatnonzero(flat)
    first = np.full(n + 1, h * w, dtype=np.int64)
    np.minimum.at(first, flat[idx], idx)
    labels.ravel()[idx] = first[flat[idx]].astype(np.int32)
    # bit-quads over the zero-padded filled label image: any two pixels of a 2x2 window are
    # 8-adjacent, so a window never mixes components -> count per window, attribute to its label
    lp = np.pad(lab, 1)
    q = np.stack([lp[:-1, :-1], lp[:-1, 1:], lp[1:, :-1], lp[1:, 1:]])
    qcnt = (q > 0).sum(0)
    qlab = q.max(0)
    q4 = np.bincount(qlab[qcnt == 4], minlength=n + 1)
    q3 = np.bincount(qlab[qcnt == 3], minlength=n + 1)
    objs = ndimage.find_objects(lab)
    for k in range(1, n + 1):
        ys, xs = objs[k - 1]
        sel = lab[ys, xs] == k
        c = dict(label=int(first[k]), xmin=int(xs.start), ymin=int(ys.start),
                 xmax=int(xs.stop - 1), ymax=int(ys.stop - 1),
                 n_pixels=int((sel & fg[ys, xs]).sum()), n_filled=int(sel.sum()),
                 area_x2=int(2 * q4[k] + q3[k]), cls_prob_sum=None)
        if probs is not None:
            c["cls_prob_sum"] = probs[ys, xs][sel].astype(np.float64).sum(0)
        comps.append(c)
    comps.sort(key=lambda c: c["label"])
    return labels, comps


def keep_component(comp, min_area) -> bool:
    """utils.py:55: ``cv2.contourArea(cnt) > min_area`` in integer form."""
    return comp["area_x2"] > 2 * min_area


def hull_points(labels, label):
    """Integer (x,y) convex hull (counter-clockwise in image coords, no collinear points) of one
    component -- what ``cv2.minAreaRect`` reduces its contour to (utils.py:56)."""
    ys, xs = np.nonzero(labels == label)
    pts = sorted(set(zip(xs.tolist(), ys.tolist())))
    if len(pts) <= 2:
        return np.array(pts, dtype=np.int32).reshape(-1, 2)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lo, up = [], []
    for p in pts:
        while len(lo) >= 2 and cross(lo[-2], lo[-1], p) <= 0:
            lo.pop()
        lo.append(p)
    for p in reversed(pts):
        while len(up) >= 2 and cross(up[-2], up[-1], p) <= 0:
            up.pop()
        up.append(p)
    return np.array(lo[:-1] + up[:-1], dtype=np.int32)


def is_equal_area_tie(box_a, box_b, rel=2e-6) -> bool:
    """True when two rounded rotated boxes are different rectangles of the same area: the only residue between
    the library and cv2.minAreaRect (a float32 tie between calipers positions whose winner depends on where
    cv2's convexHull starts, which in turn depends on duplicated points of the traced contour; observed once
    in 15,062 kept components)."""
    def area_perimeter(b):
        b = np.asarray(b, np.float64).reshape(4, 2)            # boxPoints order: consecutive corners are adjacent
        u, v = np.linalg.norm(b[1] - b[0]), np.linalg.norm(b[2] - b[1])
        return u * v, 2 * (u + v)
    (a, pa), (b, pb) = area_perimeter(box_a), area_perimeter(box_b)
    # corners are rounded to integers (x4 units): each side moves by < 1, the area by < perimeter / 2 + 1
    return abs(a - b) <= rel * max(a, b) + 0.5 * max(pa, pb) + 1.0


def boxes_equivalent(box_a, box_b, tol=0) -> bool:
    """Rotated boxes compared as corner SETS (corner order is OpenCV-version dependent, P3)."""
    a = sorted(map(tuple, np.asarray(box_a).reshape(4, 2).tolist()))
    b = sorted(map(tuple, np.asarray(box_b).reshape(4, 2).tolist()))
    return all(abs(p[0] - q[0]) <= tol and abs(p[1] - q[1]) <= tol for p, q in zip(a, b))
