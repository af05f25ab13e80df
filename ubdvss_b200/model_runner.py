"""The inference operator with the reference's signature (model_runner.py:31-38, 105-148)."""
from __future__ import annotations

import numpy as np

from . import utils
from .segmap_manager import SegmapManager, markups_from_components, min_area_x2


class ModelRunner:
    def __init__(self, net_config, pixel_threshold=0.5):
        """model_runner.py:31-38: ``pixel_probability > pixel_threshold`` is positive."""
        self._net_config = net_config
        eps = 1e-9
        self._logit_threshold = - np.log(1 / np.clip(pixel_threshold, eps, 1 - eps) - 1)

    def predict(self, model, images, rescale=False, meta_infos=None, preprocessing=None):
        """model_runner.py:105-138.  Returns ``(detection mask (N,h,w,1) int, class logits
        (N,h,w,C) float32, found_objects)``.  With a ``B200Model`` the whole chain (net, threshold,
        components, boxes, class vote) is one ``ubd_segment`` call; any other object exposing
        ``predict`` is run as the reference does and post-processed on the GPU.
        ``preprocessing`` ("none" | "mobilenet_like") may be given for uint8 images to fold
        ``NetConfig.get_preprocessing_fn`` into the first layer."""
        assert not rescale or (meta_infos is not None and len(images) == len(meta_infos))
        cfg = self._net_config
        classification = cfg.is_classification_supported()
        scale = cfg.get_scale()
        ma2 = min_area_x2(cfg.get_min_pixels_for_detection())
        thr32 = np.float32(self._logit_threshold)      # NumPy-1 float32 comparison of the reference
        if hasattr(model, "segment"):
            mask, logits, comps, counts = model.segment(images, thr32, ma2, preprocessing=preprocessing)
            detection = mask[..., None].astype(np.int64)
            classification_logits = logits[..., 1:]
        else:
            predicted = np.asarray(model.predict(images), dtype=np.float32)
            classification_logits = predicted[..., 1:]
            detection = np.where(predicted[..., :1] > thr32, 1, 0)
            eng = utils.default_engine()
            _, comps, counts = eng.postprocess(detection[..., 0].astype(np.uint8),
                                               classification_logits if classification else None, ma2)
        found_objects, o = [], 0
        for c in counts:
            found_objects.append(markups_from_components(comps[o:o + c], scale, classification))
            o += c
        if rescale:
            found_objects = self.rescale(found_objects, meta_infos)
        return detection, classification_logits, found_objects

    def predict_stream(self, model, batches, rescale=False, preprocessing=None):
        """The per-batch loop of ``ModelRunner.run`` (model_runner.py:40-103: ``next(generator)`` -> ``predict`` ->
        consume) with up to three batches in flight: while batch k runs on the GPU, batches k+1, k+2 are being copied and batch
        k-1's boxes are finished on the host.  ``batches`` yields ``images`` or ``(images, meta_infos)``; yields
        exactly what ``predict`` returns for each batch, in order, with identical values."""
        cfg = self._net_config
        classification = cfg.is_classification_supported()
        scale = cfg.get_scale()
        ma2 = min_area_x2(cfg.get_min_pixels_for_detection())
        thr32 = np.float32(self._logit_threshold)

        def finish(entry):
            ticket, metas = entry
            mask, logits, comps, counts = model.segment_wait(ticket)
            found, o = [], 0
            for c in counts:
                found.append(markups_from_components(comps[o:o + c], scale, classification))
                o += c
            if rescale:
                found = self.rescale(found, metas)
            return mask[..., None].astype(np.int64), logits[..., 1:], found

        pending = []
        for item in batches:
            images, metas = item if isinstance(item, tuple) else (item, None)
            assert not rescale or (metas is not None and len(images) == len(metas))
            pending.append((model.segment_submit(images, thr32, ma2, preprocessing=preprocessing), metas))
            if len(pending) == 3:
                yield finish(pending.pop(0))
        while pending:
            yield finish(pending.pop(0))

    @staticmethod
    def rescale(found_objects, meta_infos):
        """model_runner.py:140-148."""
        assert len(found_objects) == len(meta_infos)
        return [[obj.create_same_markup(utils.rescale_bbox(obj.bbox, xscale=mi.xscale, yscale=mi.yscale))
                 for obj in objs] for objs, mi in zip(found_objects, meta_infos)]


class ResultSaver:
    """The on-disk result format of the reference (model_runner.py:154-228); only the CSV writer is part of
    the path's boundary, visualisations are out of scope."""

    @staticmethod
    def save_markup_csv(filename, markups):
        """One line per object: the eight corner coordinates truncated to int, an empty quoted field and,
        for classified objects, the type id (model_runner.py:214-228)."""
        from .data_markup import ClassifiedObjectMarkup
        lines = []
        for m in markups:
            fields = [str(int(v)) for v in m.bbox] + ['""']
            if isinstance(m, ClassifiedObjectMarkup):
                fields.append(str(m.object_type))
            lines.append(",".join(fields) + "\n")
        with open(filename, "w") as fh:
            fh.write("".join(lines))
