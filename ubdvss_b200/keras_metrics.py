"""Training metrics with the reference's names (semantic_segmentation/keras_metrics.py:116-193).

In the reference these are graph functions handed to ``model.compile(metrics=...)``; Keras evaluates them on
every batch and ``train_on_batch`` returns ``[loss, *metrics]`` (keras_callbacks.py:30-39 logs them).  Here
they are tokens: ``B200Model.compile`` records their order and ``train_on_batch`` / ``test_on_batch`` fill the
values from the pixel statistics ``ubd_metric_counts`` gathers on the GPU (confusion counts of the detection
channel, class hits over object pixels) and from the loss components the loss kernels return.

The loss-component metrics (``pixel_positive_loss`` & co., losses.py:128-192) are wrapped by
``@_prepare_detection_args`` in the reference, so Keras logs the very components of the loss that is being
minimised; the values reported here are those components as the loss kernels return them."""
from __future__ import annotations

import numpy as np

from . import losses


class _Metric:
    """Callable token: ``metric(counts, parts)`` -> float; ``__name__`` is what Keras would log."""

    def __init__(self, name, fn, doc):
        self.__name__ = name
        self.__doc__ = doc
        self._fn = fn

    def __call__(self, counts, parts):
        return float(self._fn(counts, parts))


def _f32(v):
    return np.float32(v)


def _acc(c, _):
    tp, tn, fp, fn = (_f32(v) for v in c[:4])
    return (tp + tn) / max(_f32(1), tp + tn + fp + fn)


def _precision(c, _):
    tp, fp = _f32(c[0]), _f32(c[2])
    return tp / max(_f32(1), tp + fp)


def _recall(c, _):
    tp, fn = _f32(c[0]), _f32(c[3])
    return tp / max(_f32(1), tp + fn)


def _f1(c, p):
    pr, rc = _precision(c, p), _recall(c, p)
    return _f32(2) * pr * rc / (pr + rc) if pr + rc != 0 else _f32(0)


def _cls_acc(c, _):
    return _f32(c[4]) / max(_f32(1), _f32(c[5]))


detection_pixel_acc = _Metric("detection_pixel_acc", _acc, "keras_metrics.py:117-125")
detection_pixel_precision = _Metric("detection_pixel_precision", _precision, "keras_metrics.py:128-136")
detection_pixel_recall = _Metric("detection_pixel_recall", _recall, "keras_metrics.py:139-147")
detection_pixel_f1 = _Metric("detection_pixel_f1", _f1, "keras_metrics.py:150-158")
classification_pixel_acc = _Metric("classification_pixel_acc", _cls_acc, "keras_metrics.py:161-174")


def get_all_metrics(classification_mode=False):
    """keras_metrics.py:177-193: the four detection metrics, the class accuracy in classification mode, then
    the loss components of ``losses.get_losses``."""
    out = [detection_pixel_acc, detection_pixel_precision, detection_pixel_recall, detection_pixel_f1]
    if classification_mode:
        out.append(classification_pixel_acc)
    return out + losses.get_losses(classification_mode)
