"""Pins the post-processing oracle: (1) literal cv2 restatement == fixtures made by the
reference's own code (tools/make_golden.py); (2) cv2-free spec (what the CUDA kernels are
compared with) == literal restatement on hand cases and random masks."""
import numpy as np
import pytest

from oracle import postproc as pp
from ubdvss_b200 import synth

cv2 = pytest.importorskip("cv2")


def _split(arr, counts):
    out, o = [], 0
    for c in counts:
        out.append(arr[o:o + c]); o += c
    return out


@pytest.mark.parametrize("tag", ["m24", "m64x96"])
def test_literal_restatement_matches_reference_goldens(golden, tag):
    masks = golden[f"{tag}_masks"]
    cls = golden[f"{tag}_cls_logits"].astype(np.float32)
    boxes = _split(golden[f"{tag}_boxes"], golden[f"{tag}_counts"])
    classes = _split(golden[f"{tag}_classes"], golden[f"{tag}_counts"])
    for i in range(masks.shape[0]):
        got = pp.postprocess_cv2(masks[i][..., None].astype(np.int64), cls[i], scale=4, min_area_threshold=5)
        assert len(got) == len(boxes[i])
        for (b, c), gb, gc in zip(got, boxes[i], classes[i]):
            assert np.array_equal(b, gb) and c == gc


@pytest.mark.parametrize("thr", [50, 70])
def test_predict_restatement_matches_reference_goldens(golden, thr):
    logits = golden["predict_logits"].astype(np.float32)
    assert float(pp.logit_threshold(thr / 100)) == float(golden[f"logit_threshold_{thr}"])
    for mode in ("det", "cls"):
        key = f"predict_t{thr}_{mode}"
        det, cls, found = pp.predict_postproc_cv2(logits, thr / 100, classification=(mode == "cls"))
        assert det.dtype == np.int64 and np.array_equal(det.astype(np.uint8), golden[f"{key}_mask"])
        assert [len(f) for f in found] == list(golden[f"{key}_counts"])
        bx = [b for f in found for b, _ in f]
        assert np.array_equal(np.stack(bx), golden[f"{key}_boxes"])
        if mode == "cls":
            assert [c for f in found for _, c in f] == list(golden[f"{key}_classes"])


def test_threshold_semantics():
    assert pp.logit_threshold(0.5) == 0 and np.signbit(pp.logit_threshold(0.5))     # -0.0
    z = np.array([[-1e-30, 0.0, -0.0, 1e-30]], np.float32)
    assert pp.threshold_mask(z, pp.logit_threshold(0.5)).tolist() == [[0, 0, 0, 1]]   # strict >
    t = pp.logit_threshold(0.7)
    t32 = np.float32(t)
    z = np.array([np.nextafter(t32, np.float32(-9)), t32, np.nextafter(t32, np.float32(9))], np.float32)
    assert pp.threshold_mask(z, t).tolist() == [0, 0, 1]
    assert pp.threshold_mask(z, t).dtype == np.int64


def _cv2_components(mask):
    """Every RETR_EXTERNAL contour (no area filter): (filled mask, 2*area, boundingRect)."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    cnts = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[-2]
    out = []
    for c in cnts:
        f = np.zeros(m.shape, np.uint8)
        cv2.drawContours(f, [c], -1, 1, -1)
        out.append((f.astype(bool), int(round(2 * cv2.contourArea(c))), cv2.boundingRect(c)))
    return out


def _check_spec_equals_cv2(mask):
    labels, comps = pp.ccl_spec(mask)
    ref = _cv2_components(mask)
    assert len(comps) == len(ref)
    by_first = {}
    for f, a2, br in ref:
        by_first[int(np.flatnonzero(f.ravel())[0])] = (f, a2, br)
    for c in comps:
        f, a2, (bx, by, bw, bh) = by_first[c["label"]]
        assert np.array_equal(labels == c["label"], f)
        assert c["area_x2"] == a2
        assert (c["xmin"], c["ymin"], c["xmax"], c["ymax"]) == (bx, by, bx + bw - 1, by + bh - 1)
        assert c["n_filled"] == int(f.sum())
        assert c["n_pixels"] == int((f & (np.asarray(mask) != 0)).sum())
    assert np.array_equal(labels >= 0, pp.filled_regions(mask))


def test_spec_hand_cases(golden):
    names = list(golden["m24_names"])
    masks = golden["m24_masks"]
    for n, m in zip(names, masks):
        _check_spec_equals_cv2(m)
    get = lambda n: pp.ccl_spec(masks[names.index(n)])[1]
    assert len(get("diagonal_pair")) == 1                       # 8-connectivity
    assert len(get("ring_inner_blob")) == 1                     # nested blob dropped
    c = get("3x3_rejected_4x4_kept")
    assert [x["area_x2"] for x in c] == [8, 18]                 # (k-1)^2 polygon area
    assert [pp.keep_component(x, 5) for x in c] == [False, True]
    assert len(get("double_nesting")) == 1 and len(get("empty")) == 0
    assert get("full")[0]["area_x2"] == 2 * 23 * 23
    assert len(get("blocks_touching_diagonally")) == 1


@pytest.mark.parametrize("seed", range(4))
def test_spec_random_masks(seed):
    for m in synth.stress_masks(6, 40, 56, seed=100 + seed):
        _check_spec_equals_cv2(m)


def test_spec_property_hypothesis():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp

    @settings(max_examples=150, deadline=None)
    @given(hnp.arrays(np.uint8, st.tuples(st.integers(1, 14), st.integers(1, 14)), elements=st.integers(0, 1)))
    def run(m):
        _check_spec_equals_cv2(m)
    run()


def test_spec_class_vote_matches_literal(golden):
    masks = golden["m64x96_masks"]
    cls = golden["m64x96_cls_logits"].astype(np.float32)
    classes = _split(golden["m64x96_classes"], golden["m64x96_counts"])
    for i in range(masks.shape[0]):
        _, comps = pp.ccl_spec(masks[i], cls[i])
        kept = [c for c in comps if pp.keep_component(c, 5)]
        votes = sorted((c["label"], int(np.argmax(c["cls_prob_sum"]))) for c in kept)
        # the reference lists contours bottom-up (reverse raster order of the first pixel)
        assert [v for _, v in votes][::-1] == list(classes[i])


def test_hull_points_reduce_to_same_min_area_rect(golden):
    """cv2.minAreaRect depends only on the convex hull: feeding the integer hull of the spec's
    component gives the reference's rounded box except on equal-area ties."""
    masks = golden["m64x96_masks"]
    boxes = _split(golden["m64x96_boxes"], golden["m64x96_counts"])
    n = bad = 0
    for i in range(masks.shape[0]):
        labels, comps = pp.ccl_spec(masks[i])
        kept = sorted((c for c in comps if pp.keep_component(c, 5)), key=lambda c: -c["label"])
        assert len(kept) == len(boxes[i])
        for c, gb in zip(kept, boxes[i]):
            hull = pp.hull_points(labels, c["label"]).astype(np.int32)
            box = np.round(cv2.boxPoints(cv2.minAreaRect(hull)).reshape(8) * 4).astype(int)
            n += 1
            bad += not pp.boxes_equivalent(box, gb, tol=0)
    assert n > 50 and bad == 0
