"""Input side (segmap_manager.py:136-167, data_generators.py:135-177): the oracle's restatement of
Pillow's bicubic resize / luma conversion against Pillow itself (CPU), the GPU kernels against both (-m gpu), and the
size-grouped batching of BatchGenerator.generate."""
import numpy as np
import pytest

from oracle import prep as oprep

SHAPES = [(37, 53, 64, 64, 3), (100, 80, 64, 128, 1), (216, 384, 256, 384, 3), (300, 500, 128, 192, 3),
          (64, 64, 64, 64, 3), (50, 333, 64, 320, 1), (135, 240, 136, 240, 3)]


def _pil(img, oh, ow, grey):
    from PIL import Image
    im = Image.fromarray(img[..., 0] if img.shape[-1] == 1 else img)
    im = im.resize((ow, oh), Image.BICUBIC)
    if grey:
        im = im.convert("L")
    a = np.asarray(im)
    return a[..., None] if a.ndim == 2 else a


@pytest.mark.parametrize("shape", SHAPES)
def test_oracle_equals_pillow(shape):
    pytest.importorskip("PIL")
    h, w, oh, ow, c = shape
    img = np.random.default_rng(sum(shape)).integers(0, 256, (h, w, c), dtype=np.uint8)
    got = oprep.resize_bicubic(img if c == 3 else img[..., 0], oh, ow)
    assert np.array_equal(got if c == 3 else got[..., None], _pil(img, oh, ow, False))
    if c == 3:
        assert np.array_equal(oprep.rgb_to_l(got)[..., None], _pil(img, oh, ow, True))


def test_size_grouped_batches():
    """BatchGenerator.generate (data_generators.py:135-160): items sorted by size, grouped by shape, cut into batches;
    incomplete batches dropped unless asked for."""
    from ubdvss_b200.segmap_manager import group_batches_by_size
    shapes = [(64, 128), (64, 64), (64, 128), (128, 64), (64, 128), (64, 64), (64, 64)]
    items = [np.full(s + (1,), i, np.uint8) for i, s in enumerate(shapes)]
    full = list(group_batches_by_size(items, 2, yield_incomplete_batches=True))
    assert sorted(sorted(int(b[0, 0, 0]) for b in g) for g in full) == sorted([[1, 5], [6], [0, 2], [4], [3]])
    assert all(len({b.shape for b in g}) == 1 for g in full)
    only_full = list(group_batches_by_size(items, 2, yield_incomplete_batches=False))
    assert sorted(sorted(int(b[0, 0, 0]) for b in g) for g in only_full) == [[0, 2], [1, 5]]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + [(540, 960, 544, 960, 3)])
def test_gpu_prepare_equals_pillow(shape):
    pytest.importorskip("PIL")
    from ubdvss_b200.engine import Engine
    h, w, oh, ow, c = shape
    imgs = np.random.default_rng(sum(shape)).integers(0, 256, (3, h, w, c), dtype=np.uint8)
    e = Engine()
    for grey in (False, True):
        got = e.prepare_images(imgs, oh, ow, to_grey=grey)
        want = np.stack([_pil(imgs[i], oh, ow, grey) for i in range(3)])
        assert got.shape == want.shape and got.dtype == np.uint8
        assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_prepare_batch_follows_the_size_rule():
    """SegmapManager.prepare_batch: the reference's size rule (segmap_manager.py:145-165) + resize + grey for a list of
    decoded images of one size; equals the reference's per-image PIL path."""
    pytest.importorskip("PIL")
    from ubdvss_b200.net import NetConfig
    from ubdvss_b200.segmap_manager import SegmapManager
    cfg = NetConfig(max_image_side=256)
    imgs = np.random.default_rng(3).integers(0, 256, (2, 270, 480, 3), dtype=np.uint8)
    batch, scales = SegmapManager.prepare_batch(imgs, cfg)
    assert batch.shape == (2, 128, 256, 1) and scales == (480 / 256, 270 / 128)
    want = np.stack([_pil(imgs[i], 128, 256, True) for i in range(2)])
    assert np.array_equal(batch, want)
