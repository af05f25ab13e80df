"""-m gpu: the tcgen05 (kind::tf32) implicit-GEMM kernel vs the FP32 CUDA-core kernel and the
oracle, layer by layer (every dilation, ragged sizes) -- the bring-up and regression test of the
shared-memory descriptor / phase-decomposition logic."""
import numpy as np
import pytest

from oracle import net as onet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ubdvss_b200.engine import Engine
    e = Engine()
    e.set_weights(onet.init_weights(0, seed=1234))
    return e


def _ref_layer(w, x, layer, tf32):
    k, b = w[9 + 2 * layer], w[10 + 2 * layer]
    if tf32:
        k = onet.round_tf32(k); x = onet.round_tf32(x)
    y = onet._conv3x3_np(x.astype(np.float64), k.astype(np.float64), onet.DILATIONS[layer]) + b
    return np.maximum(y, 0)


@pytest.mark.parametrize("layer", range(6))
@pytest.mark.parametrize("shape", [(2, 64, 128), (1, 32, 256), (3, 20, 36), (1, 136, 240), (1, 4, 4), (12, 64, 64), (3, 160, 528)])
def test_tf32_layer_matches_fp32_and_oracle(eng, layer, shape):
    rng = np.random.default_rng(layer * 10 + shape[1])
    # post-ReLU-like input on the tf32 grid, as the producing kernel's epilogue (cvt.rna) leaves it
    x = onet.round_tf32(np.maximum(rng.normal(0, 1, size=shape + (24,)), 0).astype(np.float32))
    w = onet.init_weights(0, seed=1234)
    got32 = eng.debug_dilated_layer(x, layer, "fp32")
    ref = _ref_layer(w, x, layer, tf32=False)
    assert np.abs(got32 - ref).max() <= 1e-4
    got = eng.debug_dilated_layer(x, layer, "tf32")
    ref_t = _ref_layer(w, x, layer, tf32=True)
    # against the tf32-rounded oracle only fp32 accumulation order remains
    assert np.abs(got - ref_t).max() <= 2e-4, np.abs(got - ref_t).max()
    assert np.abs(got - ref).max() <= 2e-2


def test_tf32_delta_taps(eng):
    """A delta input pins every tap's (dy,dx) address offset at every dilation."""
    w = onet.init_weights(0, seed=1234)
    for layer in range(6):
        d = onet.DILATIONS[layer]
        x = np.zeros((1, 64, 160, 24), np.float32)
        x[0, 32, 70, 5] = 1.0
        x[0, 3, 150, 17] = 2.0
        got = eng.debug_dilated_layer(x, layer, "tf32")
        ref = _ref_layer(w, x, layer, tf32=True)
        assert np.abs(got - ref).max() <= 1e-5, (layer, d)


def _round_bf16(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u + np.uint32(0x7FFF) + ((u >> 16) & 1)) & np.uint32(0xFFFF0000))       # round to nearest even
    return r.view(np.float32)


@pytest.mark.parametrize("layer", range(6))
@pytest.mark.parametrize("shape", [(2, 64, 128), (1, 32, 256), (3, 20, 36), (1, 136, 240), (12, 64, 64), (3, 160, 528)])
def test_bf16_layer_matches_oracle(eng, layer, shape):
    """kind::f16 (bf16 operands, fp32 accumulate): against the oracle evaluated on bf16-rounded inputs and
    weights only the fp32 accumulation order remains."""
    rng = np.random.default_rng(layer * 10 + shape[1])
    x = _round_bf16(np.maximum(rng.normal(0, 1, size=shape + (24,)), 0).astype(np.float32))
    w = onet.init_weights(0, seed=1234)
    got = eng.debug_dilated_layer(x, layer, "bf16")
    k, b = _round_bf16(w[9 + 2 * layer]), w[10 + 2 * layer]
    ref = np.maximum(onet._conv3x3_np(x.astype(np.float64), k.astype(np.float64), onet.DILATIONS[layer]) + b, 0)
    assert np.abs(got - ref).max() <= 2e-4, np.abs(got - ref).max()


@pytest.mark.parametrize("layer", range(6))
@pytest.mark.parametrize("shape", [(2, 64, 128), (3, 20, 36), (1, 136, 240), (3, 160, 528)])
def test_f16_layer_matches_oracle(eng, layer, shape):
    """kind::f16 with IEEE-half operands (precision "f16": the significand of tf32 in a 16-bit container), fp32 accumulate:
    against the oracle on half-rounded inputs and weights only the accumulation order remains.  Interleaved with bf16
    calls on the same handle: the 16-bit weight images are rebuilt when the container changes."""
    rng = np.random.default_rng(layer * 10 + shape[1])
    x = np.maximum(rng.normal(0, 1, size=shape + (24,)), 0).astype(np.float16).astype(np.float32)
    w = onet.init_weights(0, seed=1234)
    got = eng.debug_dilated_layer(x, layer, "f16")
    k, b = w[9 + 2 * layer].astype(np.float16).astype(np.float32), w[10 + 2 * layer]
    ref = np.maximum(onet._conv3x3_np(x.astype(np.float64), k.astype(np.float64), onet.DILATIONS[layer]) + b, 0)
    assert np.abs(got - ref).max() <= 2e-4, np.abs(got - ref).max()
    if layer == 2:
        xb = _round_bf16(x)
        gb = eng.debug_dilated_layer(xb, layer, "bf16")
        kb = _round_bf16(w[9 + 2 * layer])
        refb = np.maximum(onet._conv3x3_np(xb.astype(np.float64), kb.astype(np.float64), onet.DILATIONS[layer]) + b, 0)
        assert np.abs(gb - refb).max() <= 2e-4
        assert np.array_equal(eng.debug_dilated_layer(x, layer, "f16"), got)


def test_tf32_operands_are_truncated(eng):
    """kind::tf32 ignores the low 13 mantissa bits of its operands: garbage there must not change the
    result.  (The stem's L1 producers rely on it: they round by adding half a tf32 ulp without masking.)"""
    rng = np.random.default_rng(7)
    x = onet.round_tf32(np.maximum(rng.normal(0, 1, size=(2, 32, 128, 24)), 0).astype(np.float32))
    noisy = (x.view(np.uint32) | rng.integers(0, 1 << 13, size=x.shape, dtype=np.uint32)).view(np.float32)
    for layer in (0, 3):
        a = eng.debug_dilated_layer(x, layer, "tf32")
        b = eng.debug_dilated_layer(noisy, layer, "tf32")
        assert np.array_equal(a, b)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_first_generation_kernel_still_matches(eng, precision):
    """Option tc_variant 0 (ubd_tc.cuh, one output row per accumulator) and the default column-rotating
    kernel compute the same layer: same operands, only the fp32 accumulation order differs."""
    rng = np.random.default_rng(11)
    x = np.maximum(rng.normal(0, 1, size=(3, 40, 272, 24)), 0).astype(np.float32)
    x = onet.round_tf32(x) if precision == "tf32" else _round_bf16(x)
    for layer in (1, 4):
        new = eng.debug_dilated_layer(x, layer, precision)
        eng.set_option("tc_variant", 0)
        try:
            old = eng.debug_dilated_layer(x, layer, precision)
        finally:
            eng.set_option("tc_variant", 1)
        assert np.abs(new - old).max() <= 2e-4
