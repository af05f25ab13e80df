"""CPU-side checks of the drop-in boundary: the library loads, exports exactly what include/ubd.h
declares, fails loudly without a GPU, and its pure-host box finishing matches OpenCV."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ubdvss_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "ubd.h")).read()
    declared = set(re.findall(r"\b(ubd_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert C.sizeof(_lib.Component) == 72 and _lib.COMPONENT_DTYPE.itemsize == 72


def test_fails_loudly_without_gpu(lib):
    if lib.ubd_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.ubd_create(0, 1, 1, 0, 0, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.ubd_last_error(None)
    from ubdvss_b200.engine import Engine
    with pytest.raises(_lib.UbdError):
        Engine()


def test_create_argument_errors(lib):
    h = C.c_void_p()
    assert lib.ubd_create(0, 1, 1, 99, 0, C.byref(h)) == -1
    assert lib.ubd_create(0, 1, 1, 0, 7, C.byref(h)) == -1
    assert lib.ubd_create(0, 1, 1, 0, 0, None) == -1
    assert lib.ubd_destroy(None) == -1


def _corner_set(box, scale=4):
    return sorted(map(tuple, np.round(np.asarray(box, np.float32).reshape(4, 2) * scale).astype(int).tolist()))


def test_min_area_box_matches_opencv(lib):
    """utils.py:56-57 on the host: rounded boxes equal cv2's as corner sets for every contour the reference keeps
    (contourArea > min_area, utils.py:54); below that threshold the only differences allowed are exact
    equal-area ties (cv2's hull start depends on duplicated contour points there)."""
    cv2 = pytest.importorskip("cv2")
    from ubdvss_b200 import synth
    from ubdvss_b200.engine import min_area_box
    n = bad = 0
    for m in synth.stress_masks(12, 96, 128, seed=3):
        for c in cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[-2]:
            pts = c.reshape(-1, 2)
            ref = cv2.boxPoints(cv2.minAreaRect(pts))
            got = min_area_box(pts)
            n += 1
            if _corner_set(got) != _corner_set(ref):
                assert cv2.contourArea(c) <= 5, "a kept component's box differs from cv2"
                bad += 1
                # a mismatch must be an equal-area alternative, not a wrong rectangle
                def area(b):
                    b = np.asarray(b, np.float64).reshape(4, 2)
                    return np.linalg.norm(b[1] - b[0]) * np.linalg.norm(b[2] - b[1])
                assert abs(area(got) - area(ref)) <= 1e-3 * max(1.0, area(ref))
    assert n > 1500 and bad <= n // 200, (n, bad)


def test_min_area_box_degenerate(lib):
    from ubdvss_b200.engine import min_area_box
    cv2 = pytest.importorskip("cv2")
    for pts in ([[3, 4]], [[1, 1], [5, 1]], [[0, 0], [2, 2], [4, 4]], [[2, 2], [2, 2]]):
        p = np.asarray(pts, np.int32)
        ref = cv2.boxPoints(cv2.minAreaRect(p))
        assert _corner_set(min_area_box(p)) == _corner_set(ref)
