#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE in the build container.

Runs only where ``/root/reference`` exists (it does not travel to the GPU box; the fixtures do).
What can be imported from the reference here: ``semantic_segmentation.model_runner.ModelRunner``
(``predict``: model_runner.py:105-138), ``segmap_manager.SegmapManager.postprocess``
(segmap_manager.py:42-69) and ``utils.get_contours_and_boxes`` (utils.py:51-60).  Their absent
third-party imports that this path never calls (shapely, imgaug) are stubbed, and
``cv2.findContours`` is wrapped to return the OpenCV-3 3-tuple the reference unpacks
(utils.py:52) -- the installed OpenCV is 4.13, the reference pins <4.0.  The network itself
(net.py -> Keras/TF) and the loss (losses.py -> TF) cannot be executed here, so the goldens
start from given logits: they pin threshold + contours + boxes + class vote.

Usage:  python tools/make_golden.py        (rewrites tests/golden/postproc_golden.npz)
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def import_reference():
    for name in ["shapely", "shapely.geometry", "shapely.affinity", "imgaug", "imgaug.augmenters"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    g = sys.modules["shapely.geometry"]
    g.Polygon = g.MultiPoint = g.Point = object
    sys.modules["shapely"].affinity = sys.modules["shapely.affinity"]
    sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
    sys.modules["imgaug"].seed = lambda *a, **k: None
    import cv2
    if not getattr(cv2.findContours, "_ubd_wrapped", False):
        orig = cv2.findContours

        def find_contours_cv3(*a, **k):
            cnts, hier = orig(*a, **k)
            return None, cnts, hier
        find_contours_cv3._ubd_wrapped = True
        cv2.findContours = find_contours_cv3
    sys.path.insert(0, "/root/reference")
    from semantic_segmentation.model_runner import ModelRunner
    from semantic_segmentation.segmap_manager import SegmapManager
    from semantic_segmentation import utils
    return ModelRunner, SegmapManager, utils


class _Cfg:
    def __init__(self, classification, min_area):
        self._c, self._m = classification, min_area

    def is_classification_supported(self):
        return self._c

    def get_scale(self):
        return 4

    def get_min_pixels_for_detection(self):
        return self._m


class _FakeModel:
    def __init__(self, logits):
        self._l = logits

    def predict(self, images):
        return self._l


def hand_masks():
    """The hand cases of SURVEY.md 8c(3), each 24x24."""
    cases = {}

    def blank():
        return np.zeros((24, 24), np.uint8)
    m = blank(); m[3, 3] = 1; m[4, 4] = 1; cases["diagonal_pair"] = m
    m = blank(); m[2:14, 2:14] = 1; m[4:12, 4:12] = 0; m[7:9, 7:9] = 1; cases["ring_inner_blob"] = m
    m = blank(); m[2:5, 2:5] = 1; m[10:14, 10:14] = 1; cases["3x3_rejected_4x4_kept"] = m
    m = blank(); m[0:6, 0:7] = 1; m[18:24, 15:24] = 1; cases["touching_borders"] = m
    m = blank(); m[5:11, 5:6] = 1; m[5:6, 5:15] = 1; m[10:11, 5:15] = 1; m[5:11, 14:15] = 1
    cases["thin_ring"] = m
    m = blank()
    for i in range(12):
        m[4 + i, 3 + i:3 + i + 5] = 1
    cases["sheared_bar"] = m
    m = blank(); m[2:10, 2:10] = 1; m[10:18, 10:18] = 1; cases["blocks_touching_diagonally"] = m
    m = blank(); m[1:23, 1:23] = 1; m[3:21, 3:21] = 0; m[5:19, 5:19] = 1; m[7:17, 7:17] = 0; m[10:14, 10:14] = 1
    cases["double_nesting"] = m
    m = blank(); cases["empty"] = m
    m = np.ones((24, 24), np.uint8); cases["full"] = m
    m = blank(); m[1:9, 1:9] = 1; m[0, 0] = 0; m[1, 1] = 0; m[2:8, 2:8] = 0; m[1, 1] = 0
    m[1, 2] = 1; m[2, 1] = 1; cases["hole_diagonal_to_outside"] = m
    return cases


def main():
    from ubdvss_b200 import synth
    ModelRunner, SegmapManager, utils = import_reference()
    rng = np.random.default_rng(2024)
    out = {}

    # ---- A: SegmapManager.postprocess on masks (detection only and with class logits)
    names, masks = [], []
    for k, m in hand_masks().items():
        names.append(k); masks.append(m)
    for i, m in enumerate(synth.stress_masks(12, 24, 24, seed=11)):
        names.append(f"stress24_{i}"); masks.append(m)
    masks24 = np.stack(masks)
    big = synth.stress_masks(6, 64, 96, seed=12)
    n_cls = 3
    for tag, arr in (("m24", masks24), ("m64x96", big)):
        cls_logits = rng.normal(0, 2, size=arr.shape + (n_cls,)).astype(np.float16).astype(np.float32)
        boxes, counts, classes = [], [], []
        boxes_c = []
        for i in range(arr.shape[0]):
            seg = arr[i][..., None].astype(np.int64)          # what ModelRunner hands over
            objs = SegmapManager.postprocess(seg, None, scale=4, min_area_threshold=5)
            objs_c = SegmapManager.postprocess(seg, cls_logits[i], scale=4, min_area_threshold=5)
            assert len(objs) == len(objs_c)
            counts.append(len(objs))
            boxes += [np.asarray(o.bbox, np.int32) for o in objs]
            boxes_c += [np.asarray(o.bbox, np.int32) for o in objs_c]
            classes += [o.object_type for o in objs_c]
        assert all((a == b).all() for a, b in zip(boxes, boxes_c))
        out[f"{tag}_masks"] = arr
        out[f"{tag}_cls_logits"] = cls_logits.astype(np.float16)   # exactly representable
        out[f"{tag}_counts"] = np.asarray(counts, np.int32)
        out[f"{tag}_boxes"] = np.stack(boxes).astype(np.int32) if boxes else np.zeros((0, 8), np.int32)
        out[f"{tag}_classes"] = np.asarray(classes, np.int32)
    out["m24_names"] = np.asarray(names)

    # ---- B: utils.get_contours_and_boxes raw float boxes + contourArea of every contour
    import cv2
    areas_x2, fboxes, fcounts = [], [], []
    for i in range(big.shape[0]):
        cnts, bxs = utils.get_contours_and_boxes(big[i], min_area=5)
        fcounts.append(len(cnts))
        areas_x2 += [int(round(2 * cv2.contourArea(c))) for c in cnts]
        fboxes += [np.asarray(b, np.float32) for b in bxs]
    out["m64x96_kept_area_x2"] = np.asarray(areas_x2, np.int64)
    out["m64x96_float_boxes"] = np.stack(fboxes).astype(np.float32)
    out["m64x96_float_counts"] = np.asarray(fcounts, np.int32)

    # ---- C: ModelRunner.predict from given logits (threshold + postprocess), two thresholds
    logits = rng.normal(-0.5, 1.5, size=(4, 32, 48, 1 + n_cls)).astype(np.float32)
    # smooth channel 0 so that the mask has blobs, not salt and pepper
    z = logits[..., 0]
    z = (z + np.roll(z, 1, 1) + np.roll(z, 1, 2) + np.roll(z, -1, 1) + np.roll(z, -1, 2)
         + np.roll(np.roll(z, 1, 1), 1, 2) + np.roll(np.roll(z, -1, 1), -1, 2)) / 2.0
    logits[..., 0] = z
    logits = logits.astype(np.float16).astype(np.float32)
    out["predict_logits"] = logits.astype(np.float16)          # exactly representable
    for thr in (0.5, 0.7):
        for classification in (False, True):
            runner = ModelRunner(_Cfg(classification, 5), pixel_threshold=thr)
            t32 = np.float32(runner._logit_threshold)
            assert not np.any(logits[..., 0] == t32), "fixture must not sit on the threshold"
            det, cls, found = runner.predict(_FakeModel(logits), None)
            key = f"predict_t{int(thr * 100)}_{'cls' if classification else 'det'}"
            out[f"{key}_mask"] = det.astype(np.uint8)
            out[f"{key}_counts"] = np.asarray([len(f) for f in found], np.int32)
            bx = [np.asarray(o.bbox, np.int32) for f in found for o in f]
            out[f"{key}_boxes"] = np.stack(bx) if bx else np.zeros((0, 8), np.int32)
            if classification:
                out[f"{key}_classes"] = np.asarray([o.object_type for f in found for o in f], np.int32)
        out[f"logit_threshold_{int(thr * 100)}"] = np.float64(runner._logit_threshold)

    path = os.path.join(ROOT, "tests", "golden", "postproc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: np.shape(v) for k, v in out.items()})
    host_side_golden(ModelRunner, SegmapManager, utils)


def host_side_golden(ModelRunner, SegmapManager, utils):
    """The callers either side of the path: network input size rule
    (segmap_manager.py:136-173), markup rescaling (model_runner.py:140-148, utils.py:67-69) and the CSV
    writer (ResultSaver.save_markup_csv, model_runner.py:214-228), executed from the reference and stored as JSON."""
    import json
    import tempfile
    from PIL import Image
    from semantic_segmentation.data_markup import ObjectMarkup, ClassifiedObjectMarkup
    from semantic_segmentation.model_runner import ResultSaver

    class Cfg:
        def __init__(self, mult, max_side):
            self.m, self.s = mult, max_side

        def get_side_multiple(self):
            return self.m

        def get_max_side(self):
            return self.s

    gold = {"sizes": [], "csv": [], "rescale": []}
    for (w, h, mult, max_side, override) in [(3840, 2160, 64, 4096, None), (3840, 2160, 64, 512, None), (1000, 700, 64, 512, None),
                                             (700, 1000, 64, 512, None), (640, 480, 64, 1024, None), (33, 20, 64, 512, None),
                                             (95, 97, 64, 512, None), (96, 160, 64, 512, None), (1024, 1024, 64, 512, 1024),
                                             (1500, 1100, 32, 512, 768), (800, 2400, 64, 1024, None), (224, 288, 64, 512, None)]:
        img = Image.new("L", (w, h))
        markup = [ObjectMarkup(np.array([10.0, 12.0, 50.0, 12.0, 50.0, 40.0, 10.0, 40.0]))]
        rimg, rmark = SegmapManager._rescale_image_and_markup(img, markup, Cfg(mult, max_side), max_side=override)
        gold["sizes"].append({"w": w, "h": h, "side_multiple": mult, "max_side": max_side, "override": override,
                              "new_w": rimg.size[0], "new_h": rimg.size[1],
                              "markup": [float(v) for v in rmark[0].bbox]})
    objs = [ObjectMarkup(np.array([1, 2, 30, 4, 33, 44, 5, 40])), ClassifiedObjectMarkup(np.array([7, 8, 9, 10, 11, 12, 13, 14]), 3),
            ObjectMarkup(np.array([100.9, 2.2, 300.5, 4.5, 330.1, 440.7, 50.0, 400.0]))]
    for sel in ([0], [1], [0, 1, 2], []):
        with tempfile.NamedTemporaryFile("r", suffix=".csv") as f:
            ResultSaver.save_markup_csv(f.name, [objs[i] for i in sel])
            gold["csv"].append({"select": sel, "text": open(f.name).read()})
    gold["csv_objects"] = [{"bbox": [float(v) for v in o.bbox], "type": getattr(o, "object_type", None)} for o in objs]

    class Meta:
        def __init__(self, xs, ys):
            self.xscale, self.yscale = xs, ys
    found = [[objs[0], objs[1]], [objs[2]]]
    metas = [Meta(1.5, 0.75), Meta(3840 / 3840, 2160 / 2176)]
    res = ModelRunner.rescale(found, metas)
    gold["rescale"] = {"scales": [[m.xscale, m.yscale] for m in metas],
                       "boxes": [[[int(v) for v in o.bbox] for o in f] for f in res],
                       "types": [[getattr(o, "object_type", None) for o in f] for f in res]}
    path = os.path.join(ROOT, "tests", "golden", "host_golden.json")
    json.dump(gold, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
