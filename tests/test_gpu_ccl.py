"""-m gpu: CUDA connected components (ubd_postprocess) vs the oracle's cv2-free spec (bit-exact on
labels and per-component integers) and vs the fixtures produced by the reference's own code."""
import numpy as np
import pytest

from oracle import postproc as pp
from ubdvss_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ubdvss_b200.engine import Engine
    e = Engine()
    e.set_option("max_comps", 1 << 16)
    return e


def _split(arr, counts):
    out, o = [], 0
    for c in counts:
        out.append(arr[o:o + c]); o += c
    return out


def _check_against_spec(eng, masks, cls=None, min_area=5):
    labels, comps, counts = eng.postprocess(masks, cls, min_area_x2=2 * min_area, want_labels=True)
    per_img = _split(comps, counts)
    for i in range(masks.shape[0]):
        ref_labels, ref = pp.ccl_spec(masks[i], None if cls is None else cls[i])
        assert np.array_equal(labels[i], ref_labels)
        kept = sorted((c for c in ref if pp.keep_component(c, min_area)), key=lambda c: -c["label"])
        assert len(kept) == counts[i]
        for g, r in zip(per_img[i], kept):
            assert g["image"] == i
            for k in ("label", "xmin", "ymin", "xmax", "ymax", "n_pixels", "n_filled", "area_x2"):
                assert int(g[k]) == r[k], (i, k, g, r)
            if cls is not None:
                s = np.sort(r["cls_prob_sum"])[::-1]
                if s[0] - s[1] > 1e-4 * s[0]:          # unambiguous vote
                    assert int(g["class_id"]) == int(np.argmax(r["cls_prob_sum"]))
            else:
                assert int(g["class_id"]) == -1
    return comps, counts


def test_hand_cases_and_golden_boxes(eng, golden):
    for tag in ("m24", "m64x96"):
        masks = golden[f"{tag}_masks"]
        cls = golden[f"{tag}_cls_logits"].astype(np.float32)
        comps, counts = _check_against_spec(eng, masks, cls)
        assert list(counts) == list(golden[f"{tag}_counts"])
        gb = golden[f"{tag}_boxes"]
        gc = golden[f"{tag}_classes"]
        bad = 0
        for c, b, k in zip(comps, gb, gc):
            box = np.round(c["box"] * 4).astype(int)
            bad += not pp.boxes_equivalent(box, b)
            assert int(c["class_id"]) == int(k)
        assert bad == 0, bad                          # P3: every golden box is reproduced bit-exactly


@pytest.mark.parametrize("shape", [(12, 40, 56), (6, 256, 256), (2, 544, 960), (3, 17, 33), (2, 1, 70), (2, 50, 1)])
def test_stress_masks_bit_exact(eng, shape):
    masks = synth.stress_masks(*shape, seed=sum(shape))
    _check_against_spec(eng, masks)


@pytest.mark.parametrize("shape", [(12, 40, 56), (6, 256, 256), (2, 544, 960), (3, 17, 33), (2, 1, 70), (2, 50, 1)])
def test_gpu_rectangles_equal_host_rectangles(shape):
    """ccl_boxes_kernel (hull + float32 rotating calipers, one warp per component) against the host path
    (hull candidates -> ubd_min_area_box): every box bit-identical, degenerate components (1 and 2 hull points,
    collinear pixels) included (min_area -1 keeps everything)."""
    from ubdvss_b200.engine import Engine
    e = Engine()
    masks = synth.stress_masks(*shape, seed=5 + sum(shape))
    _, a, ca = e.postprocess(masks, None, -1)
    e.set_option("gpu_boxes", 0)
    _, b, cb = e.postprocess(masks, None, -1)
    assert np.array_equal(ca, cb) and len(a) == len(b) > 0
    assert np.array_equal(a["label"], b["label"])
    assert np.array_equal(a["box"].view(np.uint32), b["box"].view(np.uint32))


@pytest.mark.parametrize("shape", [(12, 40, 56), (6, 256, 256), (3, 17, 33), (2, 1, 70), (2, 50, 1), (1, 128, 512)])
@pytest.mark.parametrize("variant", [1, 2])
def test_whole_image_kernel_bit_exact(shape, variant):
    """Option fused_ccl 1 / 2: maps of up to 65,536 px are labelled by one CTA per image in shared memory (pixel-level
    ccl_image_kernel / run-length ccl_rle_kernel); same labels, records, class votes and boxes as the tiled path."""
    from ubdvss_b200.engine import Engine
    e = Engine()
    e.set_option("fused_ccl", 0)
    masks = synth.stress_masks(*shape, seed=sum(shape))
    comps_t, counts_t = _check_against_spec(e, masks)
    e.set_option("fused_ccl", variant)
    comps_f, counts_f = _check_against_spec(e, masks)
    assert np.array_equal(counts_t, counts_f) and np.array_equal(comps_t, comps_f)
    cls = np.random.default_rng(1).normal(0, 2, size=masks.shape + (5,)).astype(np.float32)
    e.set_option("fused_ccl", 0)
    _, ct, nt = e.postprocess(masks, cls, 10)
    e.set_option("fused_ccl", variant)
    _, cf, nf = e.postprocess(masks, cls, 10)
    assert np.array_equal(nt, nf) and np.array_equal(ct, cf)
    assert np.array_equal(counts_t, counts_f) and np.array_equal(comps_t, comps_f)


def test_degenerate_masks(eng):
    z = np.zeros((2, 32, 32), np.uint8)
    _, comps, counts = eng.postprocess(z, None, 10)
    assert len(comps) == 0 and list(counts) == [0, 0]
    o = np.ones((1, 32, 64), np.uint8)
    labels, comps, counts = eng.postprocess(o, None, 10, want_labels=True)
    assert list(counts) == [1] and int(comps[0]["area_x2"]) == 2 * 31 * 63 and (labels == 0).all()
    assert (int(comps[0]["xmin"]), int(comps[0]["ymin"]), int(comps[0]["xmax"]), int(comps[0]["ymax"])) == (0, 0, 63, 31)
    # non 0/1 foreground values (utils.py:52 casts to uint8 and treats non-zero as foreground)
    m = np.zeros((1, 16, 16), np.uint8); m[0, 2:9, 3:12] = 7
    _check_against_spec(eng, m)


def test_component_slots_grow_and_capacity_is_reported():
    """The reference's cv2 path takes any number of contours (utils.py:52): the per-image slot table grows inside the
    library and only the CC stage is redone; the caller's kept-component capacity is still an error."""
    from ubdvss_b200 import _lib
    from ubdvss_b200.engine import Engine
    e = Engine()
    e.set_option("max_comps", 4)
    m = np.zeros((2, 32, 32), np.uint8); m[0, ::2, ::2] = 1; m[1, 4:9, 4:9] = 1      # 256 isolated pixels | one block
    _, comps, counts = e.postprocess(m, None, -1)
    assert list(counts) == [256, 1]
    labels, comps2 = pp.ccl_spec(m[0])
    assert sorted(int(c["label"]) for c in comps[:256]) == sorted(int(c["label"]) for c in comps2)
    with pytest.raises(_lib.UbdError) as err:
        e.postprocess(m, None, -1, max_comps=16)
    assert err.value.code == -4 and "capacity" in str(err.value)


def test_full_size_properties(eng):
    """BASELINE full size (64 maps of 256x256): size-independent properties instead of the oracle:
    every kept label is the raster-first pixel of its component, pixel counts add up, relabelling
    the filled mask is idempotent, and a vertical flip preserves the multiset of areas."""
    masks = synth.stress_masks(64, 256, 256, seed=77)
    labels, comps, counts = eng.postprocess(masks, None, -1, want_labels=True)   # keep every component
    filled = (labels >= 0).astype(np.uint8)
    assert (filled >= masks).all()
    o = 0
    for i in range(64):
        cs = comps[o:o + counts[i]]; o += counts[i]
        for c in cs[:: max(1, len(cs) // 16)]:
            l = int(c["label"])
            assert labels[i].ravel()[l] == l and (labels[i].ravel()[:l] != l).all()
            assert int(c["n_filled"]) == int((labels[i] == l).sum())
    labels2, comps2, counts2 = eng.postprocess(filled, None, -1, want_labels=True)
    assert np.array_equal(labels2, labels)
    _, comps_f, counts_f = eng.postprocess(masks[:, ::-1, :].copy(), None, -1)
    assert sorted(comps["area_x2"].tolist()) == sorted(comps_f["area_x2"].tolist())
