"""Mirrors of the reference helpers on the hot path (semantic_segmentation/utils.py:51-60, 67-76,
135-138), backed by the CUDA connected-component kernels instead of OpenCV."""
from __future__ import annotations

import numpy as np

_default_engine = None


def default_engine():
    """Process-wide handle on cuda:0 for the model-free entry points (postprocess)."""
    global _default_engine
    if _default_engine is None:
        from .engine import Engine
        _default_engine = Engine(device=0)
    return _default_engine


def get_contours_and_boxes(seg_map, min_area=10):
    """utils.py:51-60.  Returns ``(cnts, boxes)``: ``boxes`` are the (8,) float32 corner arrays of
    ``cv2.boxPoints(cv2.minAreaRect(contour))`` of every external component whose contour area is
    ``> min_area``, bottom-up as OpenCV lists them.  ``cnts`` are the component records (label, bbox,
    pixel counts, 2*area) -- the GPU path never traces contour polygons."""
    m = np.asarray(seg_map)
    if m.ndim == 3:
        m = m[..., 0]
    m = np.array(m, dtype=np.uint8)                        # utils.py:52 cast
    _, comps, _ = default_engine().postprocess(m[None], None, min_area_x2=int(np.floor(2 * min_area)))
    boxes = [np.array(c["box"], dtype=np.float32) for c in comps]
    return list(comps), boxes


def np_softmax(logits, axis=-1):
    """utils.py:135-138."""
    x = logits - np.max(logits, axis=axis, keepdims=True)
    x = np.exp(x)
    return x / np.sum(x, axis=axis, keepdims=True)


def rescale_bbox(bbox, xscale, yscale):
    """utils.py:67-69."""
    scale = np.array([xscale, yscale] * (len(bbox) // 2))
    return (bbox * scale).astype(int)


def rescale_bboxes(bboxes, xscale, yscale):
    """utils.py:72-76."""
    if not bboxes:
        return bboxes
    scale = np.array([xscale, yscale] * (len(bboxes[0]) // 2))
    return (bboxes * scale).astype(int)
