"""Callers either side of the path against goldens produced by running the reference
(tools/make_golden.py -> tests/golden/host_golden.json): input size rule, markup rescaling, CSV format."""
import json
import os

import numpy as np

from ubdvss_b200.data_markup import ClassifiedObjectMarkup, ObjectMarkup
from ubdvss_b200.model_runner import ModelRunner, ResultSaver
from ubdvss_b200.segmap_manager import SegmapManager

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "host_golden.json")))


class _Cfg:
    def __init__(self, mult, max_side):
        self.m, self.s = mult, max_side

    def get_side_multiple(self):
        return self.m

    def get_max_side(self):
        return self.s


def test_network_input_size_rule():
    for g in GOLD["sizes"]:
        got = SegmapManager.network_input_size(g["w"], g["h"], _Cfg(g["side_multiple"], g["max_side"]), g["override"])
        assert got == (g["new_w"], g["new_h"]), (g, got)
    # 2160 / 64 = 33.75 -> 34 * 64
    assert SegmapManager.network_input_size(3840, 2160, _Cfg(64, 4096)) == (3840, 2176)


def test_rescale_image_and_markup():
    from PIL import Image
    for g in GOLD["sizes"][:6]:
        img = Image.new("L", (g["w"], g["h"]))
        markup = [ObjectMarkup(np.array([10.0, 12.0, 50.0, 12.0, 50.0, 40.0, 10.0, 40.0]))]
        rimg, rmark = SegmapManager._rescale_image_and_markup(img, markup, _Cfg(g["side_multiple"], g["max_side"]), g["override"])
        assert rimg.size == (g["new_w"], g["new_h"])
        np.testing.assert_allclose(rmark[0].bbox, g["markup"], rtol=0, atol=1e-12)
    img = Image.new("L", (100, 60))
    rimg, rmark = SegmapManager._rescale_image_and_markup(img, None, _Cfg(64, 512))
    assert rimg.size == (128, 64) and rmark is None


def _objects():
    return [ClassifiedObjectMarkup(np.array(o["bbox"]), o["type"]) if o["type"] is not None else ObjectMarkup(np.array(o["bbox"]))
            for o in GOLD["csv_objects"]]


def test_csv_writer_format(tmp_path):
    objs = _objects()
    for g in GOLD["csv"]:
        f = tmp_path / "m.csv"
        ResultSaver.save_markup_csv(str(f), [objs[i] for i in g["select"]])
        assert f.read_text() == g["text"]


def test_rescale_markups():
    objs = _objects()

    class Meta:
        def __init__(self, s):
            self.xscale, self.yscale = s
    found = [[objs[0], objs[1]], [objs[2]]]
    res = ModelRunner.rescale(found, [Meta(s) for s in GOLD["rescale"]["scales"]])
    assert [[[int(v) for v in o.bbox] for o in f] for f in res] == GOLD["rescale"]["boxes"]
    assert [[getattr(o, "object_type", None) for o in f] for f in res] == GOLD["rescale"]["types"]
