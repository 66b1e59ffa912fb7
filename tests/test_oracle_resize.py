"""Pin the integer restatement of cv2's uint8 resize (oracle/resize.py) against cv2 itself."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import resize as R


@pytest.mark.parametrize("side", [2, 9, 37, 48, 100, 113, 224, 320, 448, 512, 700])
@pytest.mark.parametrize("dsize", [224, 512])
def test_restatement_is_bit_exact_vs_cv2(side, dsize):
    rng = np.random.default_rng(side * 1000 + dsize)
    for ch in (1, 3):
        a = rng.integers(0, 256, (side, side, ch), dtype=np.uint8)
        a = a[..., 0] if ch == 1 else a
        assert np.array_equal(R.lanczos4_u8(a, dsize), cv2.resize(a, (dsize, dsize), interpolation=cv2.INTER_LANCZOS4))
        assert np.array_equal(R.linear_u8(a, dsize), cv2.resize(a, (dsize, dsize), interpolation=cv2.INTER_LINEAR))


def test_binary_mask_ringing_saturates():
    m = np.zeros((90, 90), np.uint8)
    m[20:70, 30:80] = 255
    assert np.array_equal(R.lanczos4_u8(m, 512), cv2.resize(m, (512, 512), interpolation=cv2.INTER_LANCZOS4))


def test_crop_batch_reference_expression():
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    mask = (rng.integers(0, 2, (200, 300)) * 255).astype(np.uint8)
    boxes = [[10, 20, 110, 120], [150, 50, 250, 150]]
    out = R.crop_batch_reference(frame, mask, boxes, size=64, interp=R.LANCZOS4)
    assert out.shape == (2, 3, 64, 64) and out.dtype == np.float32
    img, mk = R.crop_u8_reference(frame, mask, boxes, 64, R.LANCZOS4)
    lut = R.normalise_lut()
    want = lut[mk[..., None].astype(int), img.astype(int)].transpose(0, 3, 1, 2)
    assert np.array_equal(out, want)
    nomask = R.crop_batch_reference(frame, None, boxes, size=64, interp=R.BILINEAR)
    img2, _ = R.crop_u8_reference(frame, None, boxes, 64, R.BILINEAR)
    assert np.array_equal(nomask, lut[255][img2.astype(int)].transpose(0, 3, 1, 2))
