"""CPU restatement of the reference's in-graph training metrics (test infrastructure only).

Follows semantic_segmentation/keras_metrics.py: `_acc` (34-37), `confusion_matrix` (40-62), `precision` /
`recall` / `f1` (65-108), `_get_detection_labels` (111-114), `detection_pixel_*` (117-158),
`classification_pixel_acc` (161-174) and the order of `get_all_metrics` (177-193).  Parity unpinned: the
reference evaluates them inside TensorFlow-1.x, which cannot run here; the arithmetic is elementary."""
import numpy as np


def confusion(y_true, y_pred):
    """tp, tn, fp, fn of `y_pred[..., 0] > 0` against `y_true > 0` (keras_metrics.py:40-62, 111-114)."""
    t = np.asarray(y_true).reshape(-1) > 0
    d = np.asarray(y_pred)[..., 0].reshape(-1) > 0
    return int((t & d).sum()), int((~t & ~d).sum()), int((~t & d).sum()), int((t & ~d).sum())


def detection_metrics(y_true, y_pred):
    """acc, precision, recall, f1 in float32 as the graph computes them."""
    tp, tn, fp, fn = (np.float32(v) for v in confusion(y_true, y_pred))
    one = np.float32(1)
    acc = (tp + tn) / max(one, tp + tn + fp + fn)
    precision = tp / max(one, tp + fp)
    recall = tp / max(one, tp + fn)
    f1 = np.float32(2) * precision * recall / (precision + recall) if precision + recall != 0 else np.float32(0)
    return float(acc), float(precision), float(recall), float(f1)


def classification_acc(y_true, y_pred):
    """keras_metrics.py:161-174: arg-max class vs `y_true - 1` over the pixels with `y_true > 0`."""
    y = np.asarray(y_true).reshape(-1)
    cls = np.asarray(y_pred)[..., 1:].reshape(y.shape[0], -1)
    m = y > 0
    pred = cls.argmax(-1)                                   # first maximum, as tf.argmax
    correct = np.float32(((pred == (y - 1)) & m).sum())
    return float(correct / max(np.float32(1), np.float32(m.sum())))
