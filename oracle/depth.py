"""Oracle: the depth / translation branch of get_flower_poses (test infrastructure only).

Restates, on numpy + the real cv2:
  get_depth_value  sunflower/utils/image_manipulation.py:39-96   (near/far filter, mask AND, 10x10 elliptical
                   erosion, per-box masked mean in millimetres, >= 50 px reliability)
  shrink_mask      sunflower/utils/image_manipulation.py:21-36   (cv2.erode with MORPH_ELLIPSE)
  get_points3d     sunflower/utils/mvg.py:387-408                (ray length -> xyz through K^-1)
as called from sunflower/predictor/pose_predictor.py:118-135 (depth/10000, near 0.1, far 2.5) and
fast_pose_predictor.py:90-105 (depth/1000).

Pinned: tests/golden/depth.npz holds the outputs of the reference's own get_depth_value / get_points3d
(imported from /root/reference with matplotlib / plotly stubbed) on seeded inputs; tests/test_oracle_golden.py
checks this restatement against them.  ellipse_spans() restates OpenCV's getStructuringElement(MORPH_ELLIPSE)
(third party, opencv-python 4.10.0.84 pinned by the reference, 4.13.0 here) and is checked against the real cv2.
"""
import numpy as np


def ellipse_spans(k):
    """Per kernel row i: the half-open column span [j1, j2) of ones of cv2.getStructuringElement(MORPH_ELLIPSE, (k,k))."""
    r, c = k // 2, k // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    spans = []
    for i in range(k):
        dy = i - r
        j1 = j2 = 0
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))     # cv::saturate_cast<int> = round half to even
            j1, j2 = max(c - dx, 0), min(c + dx + 1, k)
        spans.append((j1, j2))
    return spans


def shrink_mask(mask, kernel_size=3):
    """image_manipulation.py:21-36."""
    import cv2
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (kernel_size, kernel_size))
    return cv2.erode(mask.astype(np.uint8), kernel, iterations=1) > 0


def erode_ellipse_numpy(valid, k):
    """The same erosion without cv2: anchor (k//2, k//2), out-of-image pixels never erode (cv2's default border)."""
    H, W = valid.shape
    a = k // 2
    pad = np.ones((H + k, W + k), bool)
    pad[a:a + H, a:a + W] = valid
    out = np.ones((H, W), bool)
    for i, (j1, j2) in enumerate(ellipse_spans(k)):
        for j in range(j1, j2):
            out &= pad[i:i + H, j:j + W]
    return out


def get_depth_value(bbox, depth, seg_mask, scale=None, near_plane=0.1, far_plane=3.0):
    """image_manipulation.py:39-96 without the visualisation branch; `depth` is not modified.
    -> (depth values in metres (N,), reliable (N,) bool)"""
    depth = np.array(depth, dtype=np.float32, copy=True)
    if scale:
        depth *= scale
    good_depth = np.logical_and(depth > near_plane, depth < far_plane)
    seg = np.logical_and(seg_mask > 128, good_depth)
    seg = shrink_mask(seg, 10)
    depth *= 1000
    vals, rel = [], []
    for bb in bbox:
        wmin, hmin, wmax, hmax = bb
        good = depth[hmin:hmax, wmin:wmax][seg[hmin:hmax, wmin:wmax]]
        rel.append(good.shape[0] >= 50)
        vals.append(0 if good.shape[0] == 0 else np.mean(good))
    return np.array(vals) / 1000, np.array(rel)


def get_points3d(uv, Zray, K):
    """mvg.py:387-408."""
    N = uv.shape[0]
    uv1 = np.hstack((uv, np.ones(N).reshape(-1, 1)))
    xnyn1 = (np.linalg.inv(K) @ uv1.T).T
    Z = Zray / np.linalg.norm(xnyn1, axis=1)
    return xnyn1 * Z.reshape(-1, 1)


def box_centres(boxes):
    """pose_predictor.py:98-100: u = (xmax+xmin)/2, v = (ymax+ymin)/2 of the detector boxes."""
    b = np.asarray(boxes, dtype=np.float64)
    return np.stack([(b[:, 2] + b[:, 0]) / 2, (b[:, 3] + b[:, 1]) / 2], 1)
