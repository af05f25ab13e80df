"""Checks of the loss / optimizer oracle: autograd restatement vs the
hand-derived gradient, plus the edge cases the reference's formulae imply."""
import numpy as np
import pytest

from oracle import loss as L
from ubdvss_b200 import synth


def _case(n=2, h=16, w=24, n_classes=0, seed=0, scale=3.0):
    rng = np.random.default_rng(seed)
    y_true = synth.synth_targets(n, h, w, n_classes, seed)
    y_pred = (rng.normal(0, scale, size=(n, h, w, 1 + n_classes))).astype(np.float32)
    return y_true, y_pred


@pytest.mark.parametrize("n_classes", [0, 4])
def test_autograd_matches_hand_derived(n_classes):
    y_true, y_pred = _case(n_classes=n_classes, seed=3)
    la, pa, ga = L.loss_and_grad(y_true, y_pred, n_classes > 0)
    ln, pn, gn = L.loss_and_grad_numpy(y_true, y_pred, n_classes > 0)
    assert abs(la - ln) <= 1e-5 * max(1, abs(ln))
    assert np.abs(ga - gn).max() <= 1e-6 + 1e-4 * np.abs(gn).max()
    for k in ("positive", "negative", "hard_negative"):
        assert abs(pa[k] - pn[k]) <= 1e-5 * max(1, abs(pn[k]))


def test_all_negative_batch_clamps():
    """losses.py:99,110: n_pos clamps to 1 => k = 1, hard = max negative CE, positive term 0."""
    _, y_pred = _case(seed=5)
    y_true = np.zeros(y_pred.shape[:3] + (1,), np.int32)
    loss, parts, grad = L.loss_and_grad(y_true, y_pred, False)
    assert parts["k"] == 1 and parts["positive"] == 0
    ce = np.logaddexp(0, y_pred[..., 0].astype(np.float64))
    # the fp32 clip + re-logit of K.binary_crossentropy costs ~1e-4 at |z| ~ 9 (real behaviour)
    assert abs(parts["hard_negative"] - ce.max()) < 1e-3
    assert abs(parts["negative"] - ce.mean()) < 1e-4


def test_all_positive_batch():
    _, y_pred = _case(seed=6)
    y_true = np.ones(y_pred.shape[:3] + (1,), np.int32)
    loss, parts, grad = L.loss_and_grad(y_true, y_pred, False)
    assert parts["k"] == 1 and parts["negative"] == 0 and parts["hard_negative"] == 0
    assert np.isfinite(loss)


def test_saturated_logits_have_zero_gradient():
    """K.binary_crossentropy clips p to [1e-7, 1-1e-7]: |z| > ~16.6 gets no gradient."""
    y_true = np.zeros((1, 2, 4, 1), np.int32); y_true[0, 0, :2] = 1
    z = np.array([[[-20, 20, 0.3, -0.2], [-17, 17, 18, -30]]], np.float32)[..., None]
    _, _, g = L.loss_and_grad(y_true, z, False)
    _, _, gn = L.loss_and_grad_numpy(y_true, z, False)
    sat = np.abs(z) >= 17
    assert np.all(g[sat] == 0) and np.all(gn[sat] == 0)
    assert np.all(g[~sat] != 0)


def test_topk_ties_prefer_lower_index():
    v = np.array([0.5, 0.7, 0.5, 0.7, 0.5], np.float32)
    assert list(L.topk_indices(v, 3)) == [1, 3, 0]
    y_true = np.zeros((1, 1, 6, 1), np.int32); y_true[0, 0, :2] = 1       # k = 2
    z = np.array([0.0, 0.0, 1.0, 1.0, 1.0, -1.0], np.float32).reshape(1, 1, 6, 1)
    _, parts, g = L.loss_and_grad(y_true, z, False)
    assert parts["k"] == 2
    # pixels 2 and 3 (lower indices among the three ties) carry the hard-negative share
    assert g[0, 0, 2, 0] == g[0, 0, 3, 0] > g[0, 0, 4, 0] > 0


def test_adam_first_step_is_lr_sign():
    p = [np.array([1.0, -2.0, 3.0], np.float32)]
    g = [np.array([0.5, -0.25, 0.0], np.float32)]
    m = [np.zeros(3, np.float32)]; v = [np.zeros(3, np.float32)]
    L.adam_step(p, g, m, v, t=1, lr=1e-3)
    assert np.allclose(p[0], [1 - 1e-3, -2 + 1e-3, 3.0], atol=2e-7)
    assert np.allclose(m[0], 0.1 * g[0]) and np.allclose(v[0], 1e-3 * g[0] ** 2)


def test_train_step_gradients_finite_difference():
    """Whole-graph gradient oracle vs central differences in float64 on a tiny case."""
    from oracle import net
    w = [a.astype(np.float64) for a in net.init_weights(0, seed=2)]
    x = synth.synth_images(1, 64, 64, seed=1).astype(np.float64) / 127.5 - 1
    y = synth.synth_targets(1, 16, 16, 0, seed=1)
    loss, parts, grads, _ = L.train_step_torch(w, x, y, False, dtype="float64")
    rng = np.random.default_rng(0)
    for wi in (0, 4, 9, 17, 21, 22):
        idx = tuple(rng.integers(0, s) for s in w[wi].shape)
        eps = 1e-5
        wp = [a.copy() for a in w]; wp[wi][idx] += eps
        wm = [a.copy() for a in w]; wm[wi][idx] -= eps
        lp = L.train_step_torch(wp, x, y, False, dtype="float64")[0]
        lm = L.train_step_torch(wm, x, y, False, dtype="float64")[0]
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - grads[wi][idx]) <= 1e-5 + 1e-4 * abs(fd)
