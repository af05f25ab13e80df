"""-m gpu: NetManager / B200Model persistence with the reference's file names and call order (net.py:255-494,
train.py:99-105,190, predict.py:67-72), on Keras HDF5 files (no h5py)."""
import os

import numpy as np
import pytest

from ubdvss_b200 import hdf5, synth

pytestmark = pytest.mark.gpu


def _cfg(n_classes=0):
    from ubdvss_b200.net import NetConfig
    cfg = NetConfig(min_pixels_for_detection=5)
    if n_classes:
        cfg.set_class_names([f"type{i}" for i in range(n_classes)])
    return cfg


def test_train_script_flow_save_and_predict_script_flow_load(tmp_path):
    from ubdvss_b200.net import NetConfig, NetManager
    log_dir = str(tmp_path / "run1")
    os.makedirs(log_dir)
    # train.py:99-105
    nm = NetManager(log_dir=log_dir, net_config=_cfg(3))
    nm.build_model()
    nm.save_config()
    model = nm.get_keras_model()
    w = synth.synth_weights(3, seed=21, calibrated=True)
    model.set_weights(w)
    nm.save_model(7)                                   # ModelCheckpoint-style snapshot (net.py:418-420)
    nm.save_inference()                                # train.py:190
    for f in ("model007.h5", "model.h5", "model_weights.h5", "inference_model.h5", "config.pkl"):
        assert os.path.exists(os.path.join(log_dir, f)), f
    assert open(os.path.join(log_dir, "model.h5"), "rb").read(8) == hdf5.SIGNATURE
    arrays, names = hdf5.read_keras_weights(os.path.join(log_dir, "inference_model.h5"))
    assert names[0] == "separable_conv2d_1/depthwise_kernel:0" and all(np.array_equal(a, b) for a, b in zip(arrays, w))
    # predict.py:67-72
    nm2 = NetManager(log_dir)
    cfg2 = nm2.load_model(None)
    cfg2 = NetConfig.from_others(cfg2, max_image_side=1024, min_pixels_for_detection=7)
    assert cfg2.get_max_side() == 1024 and cfg2.get_min_pixels_for_detection() == 7 and cfg2.get_n_classes() == 3
    m2 = nm2.get_keras_model()
    assert all(np.array_equal(a, b) for a, b in zip(m2.get_weights(), w))
    x = synth.synth_images(2, 64, 128, seed=3)
    assert np.array_equal(m2.predict(x), model.predict(x))
    # warm start from another log dir (train.py:100-101)
    nm3 = NetManager(log_dir=str(tmp_path / "run2"), net_config=_cfg(3))
    merged = nm3.load_another_model(another_log_dir=log_dir)
    assert merged.get_n_classes() == 3 and all(np.array_equal(a, b) for a, b in zip(nm3.get_keras_model().get_weights(), w))


def test_load_errors(tmp_path):
    from ubdvss_b200.net import NetManager
    nm = NetManager(str(tmp_path), net_config=_cfg())
    with pytest.raises(FileNotFoundError):
        nm.load_model()
    # a weight file of another architecture (class head) is refused with the array-count message
    nm_c = NetManager(str(tmp_path / "c"), net_config=_cfg(2))
    os.makedirs(str(tmp_path / "c"))
    nm_c.build_model()
    nm_c.get_keras_model().save_weights(str(tmp_path / "w2.h5"))
    nm.build_model()
    from ubdvss_b200 import _lib
    with pytest.raises((_lib.UbdError, ValueError)):
        nm.get_keras_model().load_weights(str(tmp_path / "w2.h5"))
