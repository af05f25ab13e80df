"""Result types of the boundary (mirrors semantic_segmentation/data_markup.py:9-36)."""


class ObjectMarkup:
    """One detected object: ``bbox`` = 8 ints, the 4 corners (x, y) of its rotated box."""
    __slots__ = ["bbox"]

    def __init__(self, bbox):
        self.bbox = bbox

    def create_same_markup(self, new_bbox):
        return ObjectMarkup(new_bbox)


class ClassifiedObjectMarkup(ObjectMarkup):
    """Object with its type id (an ``int``, as data_markup.py:33 forces)."""
    __slots__ = ["object_type"]

    def __init__(self, bbox, object_type):
        super().__init__(bbox)
        self.object_type = int(object_type)

    def create_same_markup(self, new_bbox):
        return ClassifiedObjectMarkup(new_bbox, self.object_type)
