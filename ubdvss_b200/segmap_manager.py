"""``SegmapManager.postprocess`` of the reference (segmap_manager.py:42-69) on the GPU."""
from __future__ import annotations

import numpy as np

from .data_markup import ClassifiedObjectMarkup, ObjectMarkup


def min_area_x2(min_area_threshold) -> int:
    """``contourArea > t`` (utils.py:55) in the kernels' integer form ``2*area > floor(2t)``
    (2*area is an integer, so the two are equivalent for every real t >= 0)."""
    return int(np.floor(2 * float(min_area_threshold)))


def markups_from_components(comps, scale, classification):
    """segmap_manager.py:55-69: ``np.round(box * scale).astype(int)`` -> markup objects."""
    out = []
    for c in comps:
        bbox = np.round(np.asarray(c["box"], dtype=np.float32) * scale).astype(int)
        out.append(ClassifiedObjectMarkup(bbox, int(c["class_id"])) if classification else ObjectMarkup(bbox))
    return out


class SegmapManager:
    @staticmethod
    def postprocess(seg_map, seg_map_class_logits=None, scale=1, min_area_threshold=5, engine=None):
        """Same arguments and result as the reference: ``seg_map`` (h,w[,1]) 0/1 map,
        ``seg_map_class_logits`` (h,w,C) or None -> list of ObjectMarkup / ClassifiedObjectMarkup."""
        from .utils import default_engine
        eng = engine or default_engine()
        m = np.asarray(seg_map)
        if m.ndim == 3:
            m = m[..., 0]
        m = np.array(m, dtype=np.uint8)
        cls = None
        if seg_map_class_logits is not None:
            cls = np.asarray(seg_map_class_logits, dtype=np.float32)[None]
        _, comps, _ = eng.postprocess(m[None], cls, min_area_x2=min_area_x2(min_area_threshold))
        return markups_from_components(comps, scale, seg_map_class_logits is not None)
