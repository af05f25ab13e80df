"""-m gpu: loss, backward and optimizer of the training step (config E) vs the CPU oracle."""
import numpy as np
import pytest

from oracle import loss as L
from oracle import net as onet
from ubdvss_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _engine(**kw):
    from ubdvss_b200.engine import Engine
    return Engine(**kw)


def _loss_case(n=2, h=24, w=40, n_classes=0, seed=0, scale=3.0):
    rng = np.random.default_rng(seed)
    y_true = synth.synth_targets(n, h, w, n_classes, seed)
    y_pred = rng.normal(0, scale, size=(n, h, w, 1 + n_classes)).astype(np.float32)
    return y_true, y_pred


def _check_loss(y_true, y_pred, n_classes):
    eng = _engine(n_classes=n_classes)
    parts, dl = eng.loss(y_pred, y_true)
    ref_loss, ref_parts, ref_grad = L.loss_and_grad(y_true, y_pred, n_classes > 0)
    assert abs(parts[0] - ref_loss) <= 2e-5 * max(1.0, abs(ref_loss)), (parts, ref_loss, ref_parts)
    assert abs(parts[1] - ref_parts["positive"]) <= 2e-5 * max(1.0, ref_parts["positive"])
    assert abs(parts[2] - ref_parts["negative"]) <= 2e-5 * max(1.0, ref_parts["negative"])
    assert abs(parts[3] - ref_parts["hard_negative"]) <= 2e-5 * max(1.0, ref_parts["hard_negative"])
    assert int(parts[5]) == ref_parts["k"]
    if n_classes:
        assert abs(parts[4] - ref_parts["classification"]) <= 2e-5 * max(1.0, ref_parts["classification"])
    assert np.abs(dl - ref_grad).max() <= 1e-7 + 2e-4 * np.abs(ref_grad).max()
    return parts, dl, ref_grad


@pytest.mark.parametrize("n_classes", [0, 4, 26])
@pytest.mark.parametrize("seed", [0, 1])
def test_loss_and_gradient_match_oracle(n_classes, seed):
    y_true, y_pred = _loss_case(n_classes=n_classes, seed=seed)
    _check_loss(y_true, y_pred, n_classes)


def test_loss_edge_cases():
    y_true, y_pred = _loss_case(seed=5)
    _check_loss(np.zeros_like(y_true), y_pred, 0)                    # all negative: n_pos clamps, k = 1
    _check_loss(np.ones_like(y_true), y_pred, 0)                     # all positive: hard = 0
    z = np.array([[[-20, 20, 0.3, -0.2], [-17, 17, 18, -30]]], np.float32)[..., None]
    yt = np.zeros((1, 2, 4, 1), np.int32); yt[0, 0, :2] = 1
    _, dl, _ = _check_loss(yt, z, 0)
    assert np.all(dl[np.abs(z) >= 17] == 0)                          # clipped BCE kills the gradient
    # ties at the k-th value: the lower flat indices carry the hard-negative share
    yt = np.zeros((1, 1, 6, 1), np.int32); yt[0, 0, :2] = 1
    z = np.array([0.0, 0.0, 1.0, 1.0, 1.0, -1.0], np.float32).reshape(1, 1, 6, 1)
    _, dl, ref = _check_loss(yt, z, 0)
    assert dl[0, 0, 2, 0] == dl[0, 0, 3, 0] > dl[0, 0, 4, 0] > 0


def test_loss_full_size_config_e():
    """32 x 128x128 maps (config E per GPU): 524,288 pixels through radix select + ordered ties."""
    y_true, y_pred = _loss_case(n=32, h=128, w=128, n_classes=0, seed=9, scale=2.0)
    _check_loss(y_true, y_pred, 0)


def _grad_check(n_classes, grey, fml, dtype_u8, shape=(2, 64, 96), seed=3):
    w = onet.init_weights(n_classes, seed=seed, grey=grey)
    eng = _engine(grey=grey, fml_compatible=fml, n_classes=n_classes)
    eng.set_weights(w)
    n, H, W = shape
    x = synth.synth_images(n, H, W, seed=seed, channels=1 if grey else 3)
    y = synth.synth_targets(n, H // 4, W // 4, n_classes, seed=seed)
    xf = onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32)
    if dtype_u8:
        parts = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    else:
        parts = eng.train_step(xf, y, _lib.PREPROC_NONE)
    loss, ref_parts, ref_grads, _ = L.train_step_torch(w, xf, y, n_classes > 0, fml_compatible=fml)
    assert abs(parts[0] - loss) <= 1e-4 * max(1.0, abs(loss)), (parts, loss)
    grads = eng.get_grads()
    for i, (g, r) in enumerate(zip(grads, ref_grads)):
        assert g.shape == r.shape
        tol = 1e-6 + 2e-3 * np.abs(r).max()
        assert np.abs(g - r).max() <= tol, (i, np.abs(g - r).max(), np.abs(r).max())
    return eng, w, grads


@pytest.mark.parametrize("n_classes,grey,fml,u8", [(0, True, True, True), (0, True, True, False), (4, True, True, True),
                                                  (0, True, False, True), (3, False, True, True)])
def test_backward_matches_autograd_oracle(n_classes, grey, fml, u8):
    _grad_check(n_classes, grey, fml, u8)


def test_backward_ragged_shape():
    _grad_check(0, True, True, True, shape=(3, 48, 80), seed=8)


def test_adam_step_matches_keras_formula():
    eng, w, grads = _grad_check(0, True, True, True)
    params = [a.copy() for a in w]
    m = [np.zeros_like(a) for a in w]; v = [np.zeros_like(a) for a in w]
    L.adam_step(params, grads, m, v, t=1, lr=1e-3)
    eng.adam_step(lr=1e-3)
    for a, b in zip(eng.get_weights(), params):
        assert np.abs(a - b).max() <= 1e-7 + 1e-5 * np.abs(b).max()
    # second step on the updated weights, gradients scaled as after a 2-rank sum all-reduce
    x = synth.synth_images(2, 64, 96, seed=3)
    y = synth.synth_targets(2, 16, 24, 0, seed=3)
    eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    g2 = eng.get_grads()
    L.adam_step(params, [0.5 * g for g in g2], m, v, t=2, lr=1e-3)
    eng.adam_step(lr=1e-3, grad_scale=0.5)
    for a, b in zip(eng.get_weights(), params):
        assert np.abs(a - b).max() <= 1e-6 + 1e-4 * np.abs(b).max()


def test_adam_before_train_step_is_a_state_error():
    eng = _engine()
    eng.set_weights(onet.init_weights(0, seed=1))
    with pytest.raises(_lib.UbdError) as e:
        eng.adam_step()
    assert e.value.code == -6


def test_keras_shaped_training_loop_reduces_loss():
    from ubdvss_b200 import losses
    from ubdvss_b200.net import Adam, B200Model, NetConfig
    cfg = NetConfig()
    model = B200Model(cfg, seed=0)
    model.compile(Adam(5e-3), loss=losses.get_loss(False))
    x = synth.synth_images(4, 128, 128, seed=2)
    y = synth.synth_targets(4, 32, 32, 0, seed=2)

    def gen():
        while True:
            yield x.astype(np.float32) / 127.5 - 1, y
    first = model.train_on_batch(x.astype(np.float32) / 127.5 - 1, y)
    assert model.metrics_names[:4] == ["loss", "positive_loss", "negative_loss", "hard_negative_loss"]
    model.fit_generator(gen(), steps_per_epoch=20, epochs=3, verbose=0, workers=0)
    assert model.history["loss"][-1] < 0.9 * first[0]
    assert model.history["loss"][-1] < model.history["loss"][0]
    with pytest.raises(RuntimeError):
        B200Model(cfg, seed=0).train_on_batch(x, y)


def test_full_size_config_e_step_is_finite_and_reproducible():
    """Config E: 32 x 512x512 per GPU.  The oracle is too slow here; check finiteness, determinism and
    the bias-gradient identity dL/db_head = sum of dL/dlogits."""
    w = onet.init_weights(0, seed=1234)
    eng = _engine()
    eng.set_weights(w)
    x = np.concatenate([synth.synth_images(8, 512, 512, seed=4)] * 4)
    y = np.concatenate([synth.synth_targets(8, 128, 128, 0, seed=4)] * 4)
    p1 = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    g1 = eng.get_grads()
    p2 = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    g2 = eng.get_grads()
    assert np.isfinite(p1).all() and np.array_equal(p1, p2)
    for a, b in zip(g1, g2):
        assert np.isfinite(a).all() and np.array_equal(a, b)
    logits = eng.forward(x, _lib.PREPROC_MOBILENET)
    parts, dl = eng.loss(logits, y)
    assert abs(parts[0] - p1[0]) <= 1e-5 * abs(p1[0])
    assert abs(g1[22][0] - dl.sum(dtype=np.float64)) <= 1e-4 * max(1.0, np.abs(dl).sum())


def test_native_nccl_exchange_world_of_one():
    """ubd_comm_unique_id / ubd_comm_init / ubd_allreduce_grads on a one-rank communicator: the sum over ranks is
    the identity and the step that follows is the single-GPU step.  (Two or more ranks: tools/dist_train_check.py
    under torchrun; the bench's train_step runs through this path for N > 1.)"""
    from ubdvss_b200.engine import Engine
    w = onet.init_weights(0, seed=3)
    x = synth.synth_images(2, 64, 96, seed=2)
    y = synth.synth_targets(2, 16, 24, 0, seed=2)
    eng = _engine()
    eng.set_weights(w)
    eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    g0 = [g.copy() for g in eng.get_grads()]
    eng.comm_init(Engine.comm_unique_id(), 0, 1)
    eng.allreduce_grads()
    g1 = eng.get_grads()
    assert all(np.array_equal(a, b) for a, b in zip(g0, g1))
    eng.adam_step()
    assert any(not np.array_equal(a, b) for a, b in zip(w, eng.get_weights()))


# ---------------------------------------------------------------------------------------------------------------
# tensor-core training path (a tf32 handle): dilated layers forward / data gradient / weight gradient on tcgen05
# ---------------------------------------------------------------------------------------------------------------

def _wgrad_ref(x, g, d):
    """dK[ky][kx][ic][oc] = sum_px x[y + (ky-1)d][x + (kx-1)d][ic] g[y][x][oc] in float64 (zero padding), dB = sum_px g."""
    n, h, w, _ = x.shape
    xp = np.zeros((n, h + 2 * d, w + 2 * d, 24), np.float64)
    xp[:, d:d + h, d:d + w] = x
    dk = np.zeros((3, 3, 24, 24), np.float64)
    for ky in range(3):
        for kx in range(3):
            dk[ky, kx] = np.einsum("nyxi,nyxo->io", xp[:, ky * d:ky * d + h, kx * d:kx * d + w], g.astype(np.float64))
    return dk, g.astype(np.float64).sum((0, 1, 2))


@pytest.mark.parametrize("d", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("shape", [(2, 32, 128), (1, 8, 8), (3, 20, 36), (1, 40, 260), (5, 64, 64), (1, 3, 12)])
def test_wgrad_tcgen05_matches_numpy(d, shape):
    """ubd_wgrad.cuh: the K = pixels GEMM with both operands MN-major straight from the staged map rows.  Inputs on the
    tf32 grid, so only the fp32 accumulation order differs from the float64 reference; ragged widths (not a multiple of
    the 8-pixel K step), several strips per row (w > 128), maps smaller than the dilation, more CTAs than rows."""
    rng = np.random.default_rng(d * 100 + shape[2])
    x = onet.round_tf32(np.maximum(rng.normal(0, 1, size=shape + (24,)), 0).astype(np.float32))
    g = onet.round_tf32((rng.normal(0, 1, size=shape + (24,)) * (rng.random(shape + (24,)) < 0.7)).astype(np.float32))
    eng = _engine(precision="tf32")
    eng.set_weights(onet.init_weights(0, seed=1))
    dk, db = eng.debug_wgrad(x, g, d)
    rk, rb = _wgrad_ref(x, g, d)
    scale = np.sqrt(shape[0] * shape[1] * shape[2])
    assert np.abs(dk - rk).max() <= 1e-5 * scale + 1e-6 * np.abs(rk).max(), (np.abs(dk - rk).max(), np.abs(rk).max())
    assert np.abs(db - rb).max() <= 1e-5 * scale + 1e-6 * np.abs(rb).max()
    dk2, db2 = eng.debug_wgrad(x, g, d)
    assert np.array_equal(dk, dk2) and np.array_equal(db, db2)          # fixed summation order


def test_tensor_core_training_step_rgb_tf_padding():
    """The tf32 step with a 3-channel input and TF 'same' stride-2 padding (fml_compatible = False): the stem variants the
    grey / FML default does not exercise."""
    w = onet.init_weights(3, seed=5, grey=False)
    x = synth.synth_images(2, 64, 128, seed=6, channels=3)
    y = synth.synth_targets(2, 16, 32, 3, seed=6)
    xf = onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32)
    loss, _, ref_grads, _ = L.train_step_torch(w, xf, y, True, fml_compatible=False)
    eng = _engine(grey=False, fml_compatible=False, n_classes=3, precision="tf32")
    eng.set_weights(w)
    parts = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    assert abs(parts[0] - loss) <= 5e-3 * max(1.0, abs(loss)), (parts, loss)
    for i, (g, r) in enumerate(zip(eng.get_grads(), ref_grads)):
        assert np.linalg.norm((g - r).ravel()) <= 1e-6 + 3e-2 * np.linalg.norm(r.ravel()), i


@pytest.mark.parametrize("n_classes,shape", [(0, (2, 64, 96)), (4, (3, 48, 80)), (0, (2, 128, 640)), (0, (1, 80, 1040))])
def test_tensor_core_training_step_matches_oracle(n_classes, shape):
    """The tf32 handle's training step against the autograd oracle (float32 torch-CPU) and against the library's exact
    FP32 path: tf32 operands (10-bit significands) in the six dilated layers' forward, dgrad and wgrad."""
    w = onet.init_weights(n_classes, seed=3)
    n, H, W = shape
    x = synth.synth_images(n, H, W, seed=3)
    y = synth.synth_targets(n, H // 4, W // 4, n_classes, seed=3)
    xf = onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32)
    loss, _, ref_grads, _ = L.train_step_torch(w, xf, y, n_classes > 0)
    eng = _engine(n_classes=n_classes, precision="tf32")
    eng.set_weights(w)
    parts = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    grads = eng.get_grads()
    assert abs(parts[0] - loss) <= 5e-3 * max(1.0, abs(loss)), (parts, loss)
    for i, (g, r) in enumerate(zip(grads, ref_grads)):
        err = np.abs(g - r).max()
        assert err <= 1e-6 + 5e-2 * np.abs(r).max(), (i, err, np.abs(r).max())
        assert np.linalg.norm((g - r).ravel()) <= 1e-6 + 3e-2 * np.linalg.norm(r.ravel()), i
    # option train_tc 0 = the exact FP32 path on the same handle
    eng.set_option("train_tc", 0)
    p32 = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    assert abs(p32[0] - loss) <= 1e-4 * max(1.0, abs(loss))
    for g, r in zip(eng.get_grads(), ref_grads):
        assert np.abs(g - r).max() <= 1e-6 + 2e-3 * np.abs(r).max()


def test_tensor_core_training_full_size_reproducible_and_learns():
    """Config E (32 x 512x512) on the tf32 handle: finite, bit-reproducible, and Adam steps reduce the loss."""
    w = synth.synth_weights(0, seed=1234, calibrated=True)
    eng = _engine(precision="tf32")
    eng.set_weights(w)
    x = np.concatenate([synth.synth_images(8, 512, 512, seed=4)] * 4)
    y = np.concatenate([synth.synth_targets(8, 128, 128, 0, seed=4)] * 4)
    p1 = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    g1 = eng.get_grads()
    p2 = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    g2 = eng.get_grads()
    assert np.isfinite(p1).all() and np.array_equal(p1, p2)
    for a, b in zip(g1, g2):
        assert np.isfinite(a).all() and np.array_equal(a, b)
    e32 = _engine()
    e32.set_weights(w)
    p32 = e32.train_step(x, y, _lib.PREPROC_MOBILENET)
    assert abs(p32[0] - p1[0]) <= 5e-3 * abs(p32[0])
    for a, b in zip(g1, e32.get_grads()):
        assert np.linalg.norm((a - b).ravel()) <= 2e-2 * np.linalg.norm(b.ravel()) + 1e-7
    losses = [p1[0]]
    for _ in range(5):
        eng.adam_step(lr=2e-3)
        losses.append(eng.train_step(x, y, _lib.PREPROC_MOBILENET)[0])
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_train_update_equals_step_plus_adam(precision):
    """ubd_train_update (step, exchange, Adam queued back to back, one host synchronisation) leaves exactly the weights,
    loss parts and Adam state that ubd_train_step + ubd_adam_step leave, over three consecutive steps."""
    w = onet.init_weights(2, seed=4)
    x = synth.synth_images(2, 64, 96, seed=5)
    y = synth.synth_targets(2, 16, 24, 2, seed=5)
    a = _engine(n_classes=2, precision=precision); a.set_weights(w)
    b = _engine(n_classes=2, precision=precision); b.set_weights(w)
    for _ in range(3):
        pa = a.train_step(x, y, _lib.PREPROC_MOBILENET)
        a.adam_step(lr=2e-3)
        pb = b.train_update(x, y, _lib.PREPROC_MOBILENET, lr=2e-3)
        assert np.array_equal(pa, pb)
    for u, v in zip(a.get_weights(), b.get_weights()):
        assert np.array_equal(u, v)
    assert any(not np.array_equal(u, v) for u, v in zip(w, b.get_weights()))
