"""Oracle: integer box logic of the FloPE crop path (test infrastructure only).

Restates sunflower/utils/mvg.py:324-362 of the reference.
"""
import numpy as np


def squarify_bb(bb):
    """sunflower/utils/mvg.py:324-343 - grow the short side of an xyxy box to a square.

    Arithmetic is kept in the reference's exact form (true division followed by
    int() truncation toward zero) so float and negative inputs behave the same.
    """
    xmin, ymin, xmax, ymax = bb
    xr = xmax - xmin
    yr = ymax - ymin
    diff = abs(xr - yr)
    if diff % 2 == 0:
        lo = diff / 2
        hi = diff / 2
    else:
        lo = (diff + 1) / 2
        hi = (diff - 1) / 2
    if xr > yr:
        ymin -= lo
        ymax += hi
    elif xr < yr:
        xmin -= lo
        xmax += hi
    return [int(xmin), int(ymin), int(xmax), int(ymax)]


def bb_in_frame(bb, img_shape):
    """sunflower/utils/mvg.py:345-351 - a box leaving the frame is dropped, never clamped."""
    h, w = img_shape[0], img_shape[1]
    xmin, ymin, xmax, ymax = bb
    return not (xmin < 0 or ymin < 0 or xmax > w or ymax > h)


def filter_very_large_bb(bb):
    """sunflower/utils/mvg.py:354-362 - drop boxes whose area exceeds 5x the median area."""
    bb = np.array(bb)
    area = (bb[:, 2] - bb[:, 0]) * (bb[:, 3] - bb[:, 1])
    return bb[np.logical_not(area > 5 * np.median(area))]


def squarify_filter(boxes, img_shape):
    """The loop at sunflower/predictor/fast_pose_predictor.py:69-82.

    Returns (square boxes int64 (M,4), keep mask bool (N,)).
    """
    sq, keep = [], []
    for bb in boxes:
        s = squarify_bb(bb)
        ok = bb_in_frame(s, img_shape)
        keep.append(ok)
        if ok:
            sq.append(s)
    return np.array(sq, dtype=np.int64).reshape(-1, 4), np.array(keep, dtype=bool)
