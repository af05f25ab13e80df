"""Known-answer and cross-implementation checks of the forward oracle (SURVEY.md 8c(1,2,5))."""
import numpy as np
import pytest

from oracle import net


def _zero_weights(n_classes=0):
    return [np.zeros(s, np.float32) for _, s in net.weight_spec(n_classes)]


def test_weight_spec_matches_keras_layout():
    spec = net.weight_spec(0)
    assert len(spec) == 23
    assert sum(int(np.prod(s)) for _, s in spec) == 32962
    spec = net.weight_spec(26)
    assert sum(int(np.prod(s)) for _, s in spec) == 32962 + 25 * 26
    assert spec[0] == ("separable_conv2d_1/depthwise_kernel", (3, 3, 1, 1))
    assert spec[9] == ("conv2d_1/kernel", (3, 3, 24, 24))
    assert spec[21] == ("conv2d_7/kernel", (1, 1, 24, 27))


@pytest.mark.parametrize("fml", [True, False])
def test_two_restatements_agree(fml):
    w = net.init_weights(n_classes=3, seed=7)
    x = np.random.default_rng(0).uniform(-1, 1, size=(2, 64, 128, 1)).astype(np.float32)
    a = net.forward_numpy(w, x, fml_compatible=fml)
    b = net.forward_torch(w, x, fml_compatible=fml)
    assert a.shape == b.shape == (2, 16, 32, 4)
    assert np.abs(a - b).max() <= 1e-5
    c = net.forward_numpy(w, x, fml_compatible=fml, dtype=np.float64)
    assert np.abs(a - c).max() <= 1e-4


def test_fml_padding_is_top_left():
    """net.py:229-232: the stride-2 layers see one zero row/col at the TOP/LEFT, then 'valid'."""
    w = _zero_weights()
    w[0][1, 1, 0, 0] = 1.0          # depthwise centre tap only
    w[1][0, 0, 0, 0] = 1.0          # pointwise: channel 0 = identity
    x = np.arange(64 * 64, dtype=np.float32).reshape(1, 64, 64, 1)
    _, acts = net.forward_numpy(w, x, fml_compatible=True, return_all=True)
    # output (y,x) centre tap reads padded (2y+1, 2x+1) = original (2y, 2x)
    assert np.array_equal(acts[0][0, :, :, 0], x[0, 0::2, 0::2, 0])
    _, acts = net.forward_numpy(w, x, fml_compatible=False, return_all=True)
    # TF 'same' for even sizes pads bottom/right: centre tap reads original (2y+1, 2x+1)
    assert np.array_equal(acts[0][0, :, :, 0], x[0, 1::2, 1::2, 0])


@pytest.mark.parametrize("layer,d", list(enumerate(net.DILATIONS)))
def test_delta_input_reads_kernel_taps_at_dilation(layer, d):
    """Cross-correlation orientation + dilation: a delta at p in the layer input produces
    k[i,j] at p - (i-1,j-1)*d (Keras Conv2D does not flip)."""
    rng = np.random.default_rng(layer)
    k = rng.normal(size=(3, 3, 24, 24)).astype(np.float32)
    inp = np.zeros((1, 48, 48, 24), np.float32)
    inp[0, 24, 24, 5] = 1.0
    out = net._conv3x3_np(inp, k, d)
    for i in range(3):
        for j in range(3):
            y, x = 24 - (i - 1) * d, 24 - (j - 1) * d
            assert np.allclose(out[0, y, x, :], k[i, j, 5, :])
    assert np.count_nonzero(np.abs(out).sum(-1)) == 9


def test_identity_kernels_pass_through():
    """IdentityInitializer (net.py:31-41) in L4..L9: maps pass through (inputs >= 0 after ReLU)."""
    w = net.init_weights(seed=3)
    for li in range(6):
        k = np.zeros((3, 3, 24, 24), np.float32)
        for c in range(24):
            k[1, 1, c, c] = 1
        w[9 + 2 * li] = k
        w[10 + 2 * li] = np.zeros(24, np.float32)
    x = np.random.default_rng(1).uniform(-1, 1, size=(1, 64, 64, 1)).astype(np.float32)
    _, acts = net.forward_numpy(w, x, return_all=True)
    assert np.array_equal(acts[2], acts[8])


def test_constant_input_border_pins_zero_padding():
    """All-ones kernel on a constant map counts the in-image taps: 9 inside, 6 on edges, 4 in
    corners, at every dilation (zero padding of each layer's own input)."""
    inp = np.ones((1, 40, 40, 24), np.float32)
    k = np.zeros((3, 3, 24, 24), np.float32)
    k[:, :, 0, 0] = 1
    for d in (1, 2, 4, 8, 16):
        out = net._conv3x3_np(inp, k, d)[0, :, :, 0]
        assert out[20, 20] == 9 and out[0, 20] == 6 and out[0, 0] == 4
        assert out[d - 1, 20] == 6 and out[d, 20] == 9


def test_preprocess():
    x = np.array([0, 127.5, 255.0])
    assert np.allclose(net.preprocess(x, "mobilenet_like"), [-1, 0, 1])
    assert net.preprocess(x, "none") is x
