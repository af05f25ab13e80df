"""-m gpu: the reference-facing operator (ModelRunner.predict / SegmapManager.postprocess /
get_contours_and_boxes) end to end vs the oracle and the reference-made fixtures."""
import numpy as np
import pytest

from oracle import net as onet
from oracle import postproc as pp
from ubdvss_b200 import synth

pytestmark = pytest.mark.gpu


class _FakeModel:
    def __init__(self, logits):
        self._l = logits

    def predict(self, images):
        return self._l


def _cfg(n_classes=0, min_area=5):
    from ubdvss_b200.net import NetConfig
    cfg = NetConfig(min_pixels_for_detection=min_area)
    if n_classes:
        cfg.set_class_names([f"type{i}" for i in range(n_classes)])
    return cfg


@pytest.mark.parametrize("thr", [50, 70])
@pytest.mark.parametrize("mode", ["det", "cls"])
def test_model_runner_predict_matches_reference_golden(golden, thr, mode):
    from ubdvss_b200.model_runner import ModelRunner
    logits = golden["predict_logits"].astype(np.float32)
    runner = ModelRunner(_cfg(3 if mode == "cls" else 0), pixel_threshold=thr / 100)
    assert float(runner._logit_threshold) == float(golden[f"logit_threshold_{thr}"])
    det, cls, found = runner.predict(_FakeModel(logits), None)
    key = f"predict_t{thr}_{mode}"
    assert det.shape == logits.shape[:3] + (1,) and np.array_equal(det.astype(np.uint8), golden[f"{key}_mask"])
    assert np.array_equal(cls, logits[..., 1:])
    assert [len(f) for f in found] == list(golden[f"{key}_counts"])
    boxes = [o.bbox for f in found for o in f]
    bad = sum(not pp.boxes_equivalent(b, g) for b, g in zip(boxes, golden[f"{key}_boxes"]))
    assert bad == 0
    if mode == "cls":
        assert [o.object_type for f in found for o in f] == list(golden[f"{key}_classes"])


@pytest.mark.parametrize("n_classes", [0, 4])
def test_end_to_end_against_oracle(n_classes):
    """net + threshold + CC + boxes through B200Model/ModelRunner vs torch-CPU oracle + cv2."""
    from ubdvss_b200.model_runner import ModelRunner
    from ubdvss_b200.net import B200Model
    w = onet.init_weights(n_classes, seed=1234)
    cfg = _cfg(n_classes)
    model = B200Model(cfg, weights=w)
    x = synth.synth_images(4, 256, 320, seed=21)
    xf = onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32)
    ref_logits = onet.forward_torch(w, xf)
    # pick the threshold so that ~10 % of the pixels are positive (random-init logits)
    p_thr = float(1 / (1 + np.exp(-np.quantile(ref_logits[..., 0], 0.9))))
    runner = ModelRunner(cfg, pixel_threshold=p_thr)
    det, cls, found = runner.predict(model, x, preprocessing="mobilenet_like")
    # probability maps within 1e-3; labels/boxes bit-exact ON THE IDENTICAL MASK the GPU produced
    logits = model.predict(x, preprocessing="mobilenet_like")
    sig = lambda z: 1 / (1 + np.exp(-z.astype(np.float64)))
    assert np.abs(sig(logits[..., 0]) - sig(ref_logits[..., 0])).max() <= 1e-3
    assert np.array_equal(det, pp.threshold_mask(logits[..., :1], runner._logit_threshold))
    assert np.mean(det != pp.threshold_mask(ref_logits[..., :1], runner._logit_threshold)) < 1e-3
    assert det.dtype == np.int64 and 0.02 < det.mean() < 0.3
    for i in range(x.shape[0]):
        ref = pp.postprocess_cv2(det[i], logits[i, ..., 1:] if n_classes else None, scale=4, min_area_threshold=5)
        assert len(found[i]) == len(ref)
        for o, (b, c) in zip(found[i], ref):
            assert pp.boxes_equivalent(o.bbox, b, tol=0) or pp.is_equal_area_tie(o.bbox, b)
            if n_classes:
                assert o.object_type == c
    # float input = already preprocessed (Keras semantics): same result
    det2, _, _ = runner.predict(model, xf)
    assert np.mean(det2 != det) < 1e-4


def test_postprocess_and_contours_api(golden):
    from ubdvss_b200.segmap_manager import SegmapManager
    from ubdvss_b200.utils import get_contours_and_boxes
    masks = golden["m64x96_masks"]
    fb = golden["m64x96_float_boxes"]
    o = 0
    for i in range(masks.shape[0]):
        objs = SegmapManager.postprocess(masks[i][..., None].astype(np.int64), None, scale=4, min_area_threshold=5)
        cnts, boxes = get_contours_and_boxes(masks[i], min_area=5)
        n = int(golden["m64x96_float_counts"][i])
        assert len(objs) == len(boxes) == n
        for b, g in zip(boxes, fb[o:o + n]):
            assert b.dtype == np.float32 and b.shape == (8,)
            assert pp.boxes_equivalent(np.round(b * 4), np.round(g * 4), tol=0)
        assert [int(c["area_x2"]) for c in cnts] == list(golden["m64x96_kept_area_x2"][o:o + n])
        o += n


def test_rescale():
    from ubdvss_b200.data_markup import ClassifiedObjectMarkup, ObjectMarkup
    from ubdvss_b200.model_runner import ModelRunner

    class MI:
        xscale, yscale = 2.0, 0.5
    found = [[ObjectMarkup(np.arange(8)), ClassifiedObjectMarkup(np.arange(8) + 1, 3)]]
    out = ModelRunner.rescale(found, [MI()])
    assert list(out[0][0].bbox) == [0, 0, 4, 1, 8, 2, 12, 3] and out[0][1].object_type == 3


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_chunking_does_not_change_results(precision):
    """The batch is swept in chunks (dilated layers) and sub-chunks (stem); any split must give the same masks,
    logits and components as one sweep: images are independent (model_runner.py:127-134)."""
    from ubdvss_b200 import _lib
    from ubdvss_b200.engine import Engine
    w = onet.init_weights(0, seed=7)
    x = synth.synth_images(11, 128, 320, seed=4)
    outs = []
    for chunk, stem_chunk in ((0, 0), (4, 3), (5, 1)):
        eng = Engine(precision=precision)
        eng.set_weights(w)
        if chunk:
            eng.set_option("chunk", chunk)
            eng.set_option("stem_chunk", stem_chunk)
        lg = eng.forward(x[:2], _lib.PREPROC_MOBILENET)
        thr = float(np.quantile(lg[..., 0], 0.85))
        outs.append((thr,) + tuple(eng.segment(x, outs[0][0] if outs else thr, 10, _lib.PREPROC_MOBILENET)))
    mask0, logits0, _, comps0, counts0 = outs[0][1:]
    for o in outs[1:]:
        mask, logits, _, comps, counts = o[1:]
        assert np.array_equal(mask, mask0) and np.array_equal(logits, logits0)
        assert np.array_equal(counts, counts0) and np.array_equal(comps["box"], comps0["box"])


def test_full_size_properties():
    """BASELINE configs[1] size (64 x 1024 x 1024, tf32), checked through size-independent properties: the mask is
    the strict threshold of the returned logits, components of the fused call equal those of `postprocess` on that
    mask, and permuting the batch permutes the results (images are independent, model_runner.py:127-134)."""
    from ubdvss_b200 import _lib
    from ubdvss_b200.engine import Engine
    eng = Engine(precision="tf32")
    eng.set_weights(onet.init_weights(0, seed=1234))
    base = synth.synth_images(8, 1024, 1024, seed=1)
    x = np.ascontiguousarray(np.concatenate([base] * 8, 0))
    thr = float(np.quantile(eng.forward(x[:2], _lib.PREPROC_MOBILENET)[..., 0], 0.9))
    mask, logits, _, comps, counts = eng.segment(x, thr, 10, _lib.PREPROC_MOBILENET)
    assert np.array_equal(mask, (logits[..., 0] > np.float32(thr)).astype(np.uint8))
    _, comps2, counts2 = eng.postprocess(mask, None, 10)
    assert np.array_equal(counts, counts2) and np.array_equal(comps["box"], comps2["box"])
    assert np.array_equal(comps["area_x2"], comps2["area_x2"])
    # the batch is 8 distinct images repeated 8 times: every repeat must reproduce the first block exactly
    for r in range(1, 8):
        assert np.array_equal(mask[8 * r:8 * r + 8], mask[:8]) and np.array_equal(logits[8 * r:8 * r + 8], logits[:8])
    perm = np.random.default_rng(0).permutation(64)
    mask_p, logits_p, _, comps_p, counts_p = eng.segment(np.ascontiguousarray(x[perm]), thr, 10, _lib.PREPROC_MOBILENET)
    assert np.array_equal(mask_p, mask[perm]) and np.array_equal(logits_p, logits[perm]) and np.array_equal(counts_p, counts[perm])


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_pipelined_submit_wait_equals_predict(precision):
    """ubd_segment_submit / ubd_segment_wait (three batches in flight, ModelRunner.predict_stream) return exactly what
    the synchronous ModelRunner.predict returns, batch by batch, for batches of different sizes and shapes."""
    from ubdvss_b200 import _lib
    from ubdvss_b200.model_runner import ModelRunner
    from ubdvss_b200.net import B200Model
    n_classes = 3
    cfg = _cfg(n_classes)
    w = onet.init_weights(n_classes, seed=77)
    model = B200Model(cfg, weights=w, precision=precision)
    runner = ModelRunner(cfg, pixel_threshold=0.45)
    batches = [synth.synth_images(n, h, ww, seed=30 + i) for i, (n, h, ww) in
               enumerate([(3, 128, 192), (5, 128, 192), (1, 256, 64), (4, 64, 64), (2, 192, 320)])]
    want = [runner.predict(model, b, preprocessing="mobilenet_like") for b in batches]
    got = list(runner.predict_stream(model, iter(batches), preprocessing="mobilenet_like"))
    assert len(got) == len(want)
    for (d0, c0, f0), (d1, c1, f1) in zip(want, got):
        assert np.array_equal(d0, d1) and np.array_equal(c0, c1)
        assert [[(tuple(o.bbox), getattr(o, "object_type", None)) for o in f] for f in f0] == \
               [[(tuple(o.bbox), getattr(o, "object_type", None)) for o in f] for f in f1]
    # call-order errors: a fourth submit, a synchronous call while batches are in flight, collecting out of order
    t0 = model.segment_submit(batches[0], np.float32(0.0), 10, preprocessing="mobilenet_like")
    t1 = model.segment_submit(batches[1], np.float32(0.0), 10, preprocessing="mobilenet_like")
    t2 = model.segment_submit(batches[4], np.float32(0.0), 10, preprocessing="mobilenet_like")
    for bad in (lambda: model.segment_submit(batches[2], np.float32(0.0), 10), lambda: runner.predict(model, batches[0]),
                lambda: model.segment_wait(t1), lambda: model.segment_wait(t2)):
        with pytest.raises(_lib.UbdError) as err:
            bad()
        assert err.value.code in (-6,)
    model.segment_wait(t0)
    model.segment_wait(t1)
    model.segment_wait(t2)
    assert model.predict(batches[3]).shape == (4, 16, 16, 1 + n_classes)
