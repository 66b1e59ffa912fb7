"""The streaming bilinear ROI kernel folds cv2's vertical coefficient pair into a constant (roi_stream.cuh, r3_emit2): that
needs b0 + b1 == 2048 for every destination row.  cv2 rounds (1 - fy) * 2048 and fy * 2048 separately
(oracle/resize.py), so this is a property of the coordinates, checked here over every crop size the
kernel is launched with (S % 32 == 0, S <= 512) and every source size up to 16384."""
import numpy as np
import pytest

from oracle import resize as oresize


@pytest.mark.parametrize("S", list(range(32, 513, 32)))
def test_vertical_bilinear_pairs_sum_to_2048(S):
    d = np.arange(S, dtype=np.float64)[None, :]
    src = np.arange(1, 16385, dtype=np.float64)[:, None]
    for scale in (1.0 / (float(S) / src), src / float(S)):        # the kernel's expression (cv2's) and the oracle's
        fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
        fr = (fx - np.floor(fx)).astype(np.float32)
        b0 = np.rint((np.float32(1.0) - fr).astype(np.float32) * np.float32(2048))
        b1 = np.rint(fr * np.float32(2048))
        assert np.all(b0 + b1 == 2048)


def test_formula_is_the_oracles():
    """The vectorised expression above is the oracle's tap routine (which the golden tests pin to cv2)."""
    for S, src in ((224, 37), (224, 300), (512, 100), (96, 1079), (224, 1), (512, 4000)):
        _, coef = oresize.linear_taps(src, S, vertical=True)
        assert coef.shape == (S, 2) and np.all(coef.sum(axis=1) == 2048)
