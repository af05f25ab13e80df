"""Result types of the boundary: what ``ModelRunner.predict`` / ``SegmapManager.postprocess`` hand back.

Duck-typed against the reference (semantic_segmentation/data_markup.py:9-36): callers read ``.bbox`` (the four
corners of the rotated box, x0, y0, ..., x3, y3), ``.object_type`` on classified objects, and re-wrap a changed
box with ``create_same_markup`` (model_runner.py:140-148)."""
from __future__ import annotations

import numpy as np


class ObjectMarkup:
    """One detected object."""
    __slots__ = ("bbox",)

    def __init__(self, bbox):
        self.bbox = bbox

    def create_same_markup(self, new_bbox):
        """Same kind of object around another box (used when boxes are rescaled to the source image)."""
        return type(self)(new_bbox)

    def corners(self):
        """The box as a (4, 2) array of (x, y) corners."""
        return np.asarray(self.bbox).reshape(4, 2)

    def __repr__(self):
        return f"{type(self).__name__}({[int(v) for v in np.asarray(self.bbox).ravel()]})"


class ClassifiedObjectMarkup(ObjectMarkup):
    """Detected object with its type id; the id is forced to a plain ``int`` (data_markup.py:33)."""
    __slots__ = ("object_type",)

    def __init__(self, bbox, object_type):
        ObjectMarkup.__init__(self, bbox)
        self.object_type = int(object_type)

    def create_same_markup(self, new_bbox):
        return ClassifiedObjectMarkup(new_bbox, self.object_type)

    def __repr__(self):
        return f"{ObjectMarkup.__repr__(self)[:-1]}, type={self.object_type})"
