"""Keras HDF5 weight files without h5py (net.py:418-427, 443-494): the writer's layout, the reader on a
file produced by the real HDF5 library, and the round trip through the committed fixture."""
import os

import numpy as np
import pytest

from ubdvss_b200 import hdf5, synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "model_weights_c3.h5")

_NAMES = ([f"separable_conv2d_{i}/{p}:0" for i in (1, 2, 3) for p in ("depthwise_kernel", "pointwise_kernel", "bias")]
          + [f"conv2d_{i}/{p}:0" for i in range(1, 8) for p in ("kernel", "bias")])


def _layers(weights, fml=True):
    names = ["input_1"] + (["zero_padding2d_1"] if fml else []) + ["separable_conv2d_1", "separable_conv2d_2"] + \
            (["zero_padding2d_2"] if fml else []) + ["separable_conv2d_3"] + [f"conv2d_{i}" for i in range(1, 8)]
    by = {}
    for n, w in zip(_NAMES, weights):
        by.setdefault(n.split("/")[0], []).append((n, w))
    return [(n, by.get(n, [])) for n in names]


def test_round_trip_weights_only_and_full_model(tmp_path):
    w = synth.synth_weights(3, seed=5)
    for full in (False, True):
        p = tmp_path / f"m{int(full)}.h5"
        hdf5.write_keras_weights(p, _layers(w), full_model=full, model_config='{"class_name": "Model"}')
        got, names = hdf5.read_keras_weights(p)
        assert names == _NAMES
        assert all(np.array_equal(a, b) and a.dtype == np.float32 for a, b in zip(got, w))
        f = hdf5.File(p)
        g = f["model_weights"] if full else f
        assert [n.decode() for n in g.attrs["layer_names"]][:3] == ["input_1", "zero_padding2d_1", "separable_conv2d_1"]
        assert g.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.2.4"
        assert g["conv2d_3/conv2d_3/kernel:0"].shape == (3, 3, 24, 24)
        assert g["input_1"].keys() == [] and len(g["input_1"].attrs["weight_names"]) == 0
        if full:
            assert f.attrs["model_config"] == b'{"class_name": "Model"}'


def test_file_layout_is_superblock0_symbol_tables(tmp_path):
    """What libhdf5 writes for h5py's default ``libver='earliest'``: signature, superblock 0 with 8-byte offsets,
    a root symbol table entry pointing at a version-1 object header, TREE / HEAP / SNOD nodes."""
    p = tmp_path / "w.h5"
    hdf5.write_keras_weights(p, _layers(synth.synth_weights(0, seed=1)))
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    assert int.from_bytes(b[40:48], "little") == len(b)                     # end-of-file address
    root = int.from_bytes(b[64:72], "little")
    assert b[root] == 1 and b.count(b"TREE") >= 15 and b.count(b"SNOD") >= 15 and b.count(b"HEAP") >= 15
    assert len(b) < 132_000 + 64_000                                        # 131,848 bytes of weights + structure


def test_golden_fixture_matches_synthetic_weights():
    got, names = hdf5.read_keras_weights(GOLDEN)
    want = synth.synth_weights(3, seed=1234)
    assert names == _NAMES and len(got) == 23
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_reader_on_a_file_written_by_libhdf5():
    """SciPy ships a MATLAB 7.3 file = HDF5 written by the real library behind a 512-byte user block: groups,
    attributes and datasets must parse (this is the only libhdf5-written file in the image)."""
    sio = pytest.importorskip("scipy.io")
    p = os.path.join(os.path.dirname(sio.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("SciPy test data not installed")
    f = hdf5.File(p)
    assert f.base == 512 and f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert isinstance(d, hdf5.Dataset) and d.shape == (9, 1) and d.dtype == np.float64
    assert d.attrs["MATLAB_class"] == b"double"
    assert np.allclose(d.read()[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_not_hdf5_and_wrong_architecture(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 64)
    with pytest.raises(hdf5.HDF5Error):
        hdf5.File(p)
