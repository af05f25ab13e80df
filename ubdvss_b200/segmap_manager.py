"""``SegmapManager.postprocess`` of the reference (segmap_manager.py:42-69) on the GPU, and the caller-side
size rule that decides what the network is fed (segmap_manager.py:136-173)."""
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

    @staticmethod
    def network_input_size(w, h, net_config, max_side=None):
        """(new_w, new_h) the reference resizes a w x h image to before the network
        (segmap_manager.py:145-165): both sides become multiples of ``get_side_multiple()`` (nearest
        multiple, Python's round-half-to-even, at least one); an image whose longer side exceeds
        ``max_side`` (default ``get_max_side()``) gets that side set to ``max_side`` itself and the other
        one scaled in proportion.  3840x2160 with multiple 64 -> 3840x2176 (config C)."""
        mult = net_config.get_side_multiple()
        limit = net_config.get_max_side() if max_side is None else max_side

        def snap(v):
            return max(1, round(v / mult)) * mult
        longer = max(w, h)
        if longer <= limit:
            return snap(w), snap(h)
        shrink = limit / longer
        return (limit, snap(h * shrink)) if w > h else (snap(w * shrink), limit)

    @staticmethod
    def _rescale_image_and_markup(image, markup, net_config, max_side=None):
        """segmap_manager.py:136-173: PIL image (and its markup, if any) at the network's input size;
        bicubic resampling as the reference; markup corners are scaled, not rounded."""
        from PIL import Image
        w, h = image.size
        new_w, new_h = SegmapManager.network_input_size(w, h, net_config, max_side)
        resized = image.resize(size=(new_w, new_h), resample=Image.BICUBIC)
        if not markup:
            return resized, markup
        sx, sy = new_w / w, new_h / h
        scaled = []
        for m in markup:
            pts = np.array(m.bbox, dtype=np.float64).reshape(-1, 2) * np.array([[sx, sy]])
            scaled.append(m.create_same_markup(pts.reshape(-1)))
        return resized, scaled

    @staticmethod
    def prepare_batch(images, net_config, max_side=None, engine=None):
        """The image half of ``prepare_image_and_target`` + ``BatchGenerator._prepare_image`` for a batch of decoded
        images of ONE size (segmap_manager.py:136-167, data_generators.py:176-177) on the GPU: the size rule, the bicubic
        resize and the grey conversion.  -> ``(batch (N,h,w,1|3) uint8, (xscale, yscale))`` with the scales
        ``MetaInfo`` carries (original / rescaled size, data_generators.py:186-190)."""
        from .utils import default_engine
        x = np.asarray(images)
        n, h, w, c = x.shape
        new_w, new_h = SegmapManager.network_input_size(w, h, net_config, max_side)
        out = (engine or default_engine()).prepare_images(x, new_h, new_w, to_grey=net_config.is_grey())
        return out, (w / new_w, h / new_h)


def group_batches_by_size(items, batch_size, yield_incomplete_batches=True):
    """Size-grouped batching of ``BatchGenerator.generate`` (data_generators.py:135-160): the prepared images are
    sorted by size, grouped by shape and cut into batches of ``batch_size``; an incomplete batch ends its group and is
    skipped unless ``yield_incomplete_batches``.  Yields lists of arrays of one shape."""
    import itertools
    data = sorted(items, key=lambda a: a.size)
    for _, group in itertools.groupby(data, key=lambda a: a.shape):
        group = list(group)
        i = 0
        while len(group) - i > 0:
            if not yield_incomplete_batches and len(group) - i < batch_size:
                break
            yield group[i:i + batch_size]
            i += batch_size
