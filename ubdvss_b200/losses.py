"""Loss selectors with the reference's names (semantic_segmentation/losses.py:13-24, 33-62).

In the reference these are Keras-backend graph functions handed to ``model.compile``.  Here the
loss is computed (with its gradient) by CUDA kernels inside ``ubd_train_step`` / ``ubd_loss``; the
functions below are the tokens ``B200Model.compile`` recognises, and when CALLED with arrays they
evaluate the loss on the GPU through ``ubd_loss`` (same argument meaning as the reference:
``y_true`` (bs,h,w,1) ints in [0, n_classes], ``y_pred`` (bs,h,w,1+n_classes) logits)."""
from __future__ import annotations

import numpy as np

L_POSITIVE_WEIGHT = 15.             # losses.py:13
L_NEGATIVE_WEIGHT = 1.              # losses.py:14
L_HARD_NEGATIVE_WEIGHT = 5.         # losses.py:15
L_DETECTION_WEIGHT = 1.             # losses.py:16
L_CLASSIFICATION_WEIGHT = 1.        # losses.py:17

PART_NAMES = ("loss", "positive_loss", "negative_loss", "hard_negative_loss", "classification_loss", "k")


def _evaluate(y_true, y_pred, classification):
    from .engine import Engine
    y_pred = np.asarray(y_pred, dtype=np.float32)
    n_classes = y_pred.shape[-1] - 1
    if not classification and n_classes:
        y_pred = y_pred[..., :1]
        n_classes = 0
    eng = Engine(n_classes=n_classes)
    try:
        parts, _ = eng.loss(y_pred, y_true)
    finally:
        eng.close()
    return float(parts[0])


def detection_loss(y_true, y_pred):
    """losses.py:33-44."""
    return _evaluate(y_true, y_pred, False)


def detection_and_classification_loss(y_true, y_pred):
    """losses.py:47-62."""
    return _evaluate(y_true, y_pred, True)


def get_loss(classification_mode=False):
    """losses.py:20-24."""
    return detection_and_classification_loss if classification_mode else detection_loss


class _LossComponent:
    """Metric token for one component of the loss (losses.py:139-208); value from the loss kernels' parts
    ``[loss, positive, negative, hard_negative, classification, k]``."""

    def __init__(self, name, fn, doc):
        self.__name__ = name
        self.__doc__ = doc
        self._fn = fn

    def __call__(self, counts, parts):
        return float(self._fn(parts))


_detection_part = _LossComponent(
    "detection_loss", lambda p: L_POSITIVE_WEIGHT * p[1] + L_NEGATIVE_WEIGHT * p[2] + L_HARD_NEGATIVE_WEIGHT * p[3], "losses.py:33-44")
pixel_positive_loss = _LossComponent("pixel_positive_loss", lambda p: p[1], "losses.py:139-149")
pixel_negative_loss = _LossComponent("pixel_negative_loss", lambda p: p[2], "losses.py:153-167")
pixel_hard_negative_loss = _LossComponent("pixel_hard_negative_loss", lambda p: p[3], "losses.py:170-191")
classification_loss = _LossComponent("classification_loss", lambda p: p[4], "losses.py:65-83")


def get_losses(classification_mode=False):
    """losses.py:194-208: the loss components in the order the reference logs them."""
    out = [_detection_part, pixel_positive_loss, pixel_negative_loss, pixel_hard_negative_loss]
    if classification_mode:
        out.append(classification_loss)
    return out
