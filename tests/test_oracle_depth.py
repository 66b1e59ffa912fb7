"""The depth-branch oracle (oracle/depth.py) against the reference's own get_depth_value / get_points3d outputs
(tests/golden/depth.npz, made by tests/golden/make_golden.py from /root/reference) and against the real cv2."""
import os

import numpy as np
import pytest

from oracle import depth as od


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "depth.npz"))


@pytest.mark.parametrize("tag", ["pose", "fast"])
def test_get_depth_value_matches_reference(g, tag):
    depth_m = g["raw"].astype(np.float32) / float(g[f"div_{tag}"])
    val, rel = od.get_depth_value(g["boxes"], depth_m, g["mask"], near_plane=0.1, far_plane=float(g[f"far_{tag}"]))
    assert np.array_equal(rel, g[f"rel_{tag}"])
    assert np.array_equal(np.asarray(val, np.float64), g[f"val_{tag}"])          # same numpy ops: bit-identical
    xyz = od.get_points3d(od.box_centres(g["boxes"]), np.asarray(val, np.float64), g["K"])
    assert np.array_equal(xyz, g[f"xyz_{tag}"])


def test_erosion_matches_reference_and_cv2(g):
    d = g["raw"].astype(np.float32) / 10000.0
    seg = np.logical_and(g["mask"] > 128, np.logical_and(d > 0.1, d < 2.5))
    assert np.array_equal(od.shrink_mask(seg, 10), g["eroded_pose"])
    assert np.array_equal(od.erode_ellipse_numpy(seg, 10), g["eroded_pose"])       # cv2-free restatement
    rng = np.random.default_rng(1)
    for k in (3, 5, 10, 11):
        m = rng.random((70, 90)) > 0.08
        assert np.array_equal(od.erode_ellipse_numpy(m, k), od.shrink_mask(m, k)), k


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 10, 11, 15, 16, 21])
def test_ellipse_spans_match_cv2(k):
    import cv2
    el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    for i, (j1, j2) in enumerate(od.ellipse_spans(k)):
        want = np.zeros(k, np.uint8)
        want[j1:j2] = 1
        assert np.array_equal(el[i], want), (k, i)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_yolo_postprocessing_oracle_matches_reference(golden_dir, tag):
    """oracle/detector_post.py against the outputs of the reference's own get_bbox_mask (tests/golden/yolo_post.npz)."""
    import torch
    from oracle import detector_post as op
    y = np.load(os.path.join(golden_dir, "yolo_post.npz"))
    H, W = y[f"mask_{tag}"].shape
    bbox, mask = op.bbox_mask_from_results(torch.from_numpy(y[f"masks_{tag}"].astype(np.float32)),
                                           torch.from_numpy(y[f"boxes_{tag}"]), H, W)
    assert bbox.dtype == np.int16 and np.array_equal(bbox, y[f"bbox_{tag}"])
    assert np.array_equal(mask, y[f"mask_{tag}"])
