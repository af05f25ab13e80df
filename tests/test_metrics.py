"""Training metrics: the CPU restatement of keras_metrics.py on hand cases, the token lists in
the reference's order, and (-m gpu) the GPU pixel statistics against the restatement."""
import numpy as np
import pytest

from oracle import metrics as om


def test_oracle_hand_cases():
    y = np.array([[[0], [1]], [[2], [0]]])[None]                       # (1, 2, 2, 1): two object pixels
    p = np.zeros((1, 2, 2, 3), np.float32)
    p[0, :, :, 0] = [[-1.0, 2.0], [-0.5, 3.0]]                         # detection: F T / F T  -> tp 1, tn 1, fp 1, fn 1
    p[0, 0, 1, 1:] = [5.0, 1.0]                                        # class 0 == y - 1 = 0: hit
    p[0, 1, 0, 1:] = [4.0, 4.0]                                        # tie -> class 0, label 1: miss
    assert om.confusion(y, p) == (1, 1, 1, 1)
    acc, pr, rc, f1 = om.detection_metrics(y, p)
    assert (acc, pr, rc, f1) == (0.5, 0.5, 0.5, 0.5)
    assert om.classification_acc(y, p) == 0.5
    # no positives anywhere: every ratio is guarded by max(1, .), f1 falls back to 0
    z = np.zeros((1, 2, 2, 1), np.int64)
    q = np.full((1, 2, 2, 1), -1.0, np.float32)
    assert om.detection_metrics(z, q) == (1.0, 0.0, 0.0, 0.0)


def test_token_lists_follow_the_reference_order():
    from ubdvss_b200 import keras_metrics, losses
    names = [m.__name__ for m in keras_metrics.get_all_metrics(False)]
    assert names == ["detection_pixel_acc", "detection_pixel_precision", "detection_pixel_recall", "detection_pixel_f1",
                     "detection_loss", "pixel_positive_loss", "pixel_negative_loss", "pixel_hard_negative_loss"]
    names = [m.__name__ for m in keras_metrics.get_all_metrics(True)]
    assert names[4] == "classification_pixel_acc" and names[-1] == "classification_loss" and len(names) == 10
    assert [m.__name__ for m in losses.get_losses(True)][-1] == "classification_loss"
    counts = np.array([3, 10, 1, 2, 4, 5])
    parts = np.array([9.0, 0.2, 0.3, 0.4, 0.5, 6.0], np.float32)
    vals = {m.__name__: m(counts, parts) for m in keras_metrics.get_all_metrics(True)}
    assert vals["detection_pixel_precision"] == pytest.approx(0.75) and vals["detection_pixel_recall"] == pytest.approx(0.6)
    assert vals["classification_pixel_acc"] == pytest.approx(0.8)
    assert vals["detection_loss"] == pytest.approx(15 * 0.2 + 0.3 + 5 * 0.4) and vals["pixel_negative_loss"] == pytest.approx(0.3)


@pytest.mark.gpu
@pytest.mark.parametrize("n_classes", [0, 4])
def test_gpu_pixel_statistics_match_restatement(n_classes):
    from ubdvss_b200 import synth
    from ubdvss_b200.engine import Engine
    rng = np.random.default_rng(3 + n_classes)
    y = synth.synth_targets(3, 40, 56, n_classes, seed=5)
    p = rng.normal(0, 2, size=(3, 40, 56, 1 + n_classes)).astype(np.float32)
    if n_classes:
        p[0, :5, :5, 1:] = 1.25                                       # ties: first maximum wins
    eng = Engine(n_classes=n_classes)
    eng.loss(p, y)
    c = eng.metric_counts()
    tp, tn, fp, fn = om.confusion(y, p)
    assert tuple(int(v) for v in c[:4]) == (tp, tn, fp, fn)
    if n_classes:
        m = y.reshape(-1) > 0
        assert int(c[5]) == int(m.sum())
        assert c[4] / max(1, c[5]) == pytest.approx(om.classification_acc(y, p), abs=1e-7)


@pytest.mark.gpu
def test_train_on_batch_returns_loss_and_metrics():
    from oracle import net as onet
    from ubdvss_b200 import keras_metrics, losses, synth
    from ubdvss_b200.net import Adam, B200Model, NetConfig
    model = B200Model(NetConfig(), weights=onet.init_weights(0, seed=2))
    model.compile(Adam(1e-3), loss=losses.get_loss(False), metrics=keras_metrics.get_all_metrics(False))
    assert model.metrics_names[0] == "loss" and len(model.metrics_names) == 9
    x = synth.synth_images(2, 64, 96, seed=1)
    y = synth.synth_targets(2, 16, 24, 0, seed=1)
    out = model.train_on_batch(x, y, preprocessing="mobilenet_like")
    assert len(out) == 9 and all(np.isfinite(out))
    logs = dict(zip(model.metrics_names, out))
    assert logs["detection_loss"] == pytest.approx(logs["loss"], rel=1e-5)
    assert 0.0 <= logs["detection_pixel_acc"] <= 1.0
    ev = model.test_on_batch((x.astype(np.float32) - 127.5) / 127.5, y)
    assert len(ev) == 9
