"""Host-side box logic and yaw nullification, mirroring sunflower/utils/mvg.py.

Same names, argument meaning and return types as the reference:
  squarify_bb            sunflower/utils/mvg.py:324-343
  bb_in_frame            sunflower/utils/mvg.py:345-351
  filter_very_large_bb   sunflower/utils/mvg.py:354-362
  nullify_yaw_batch      sunflower/utils/mvg.py:240-251
  get_points3d           sunflower/utils/mvg.py:387-408   (host, float64: four flops per flower)
  pose_cam_to_world      sunflower/utils/mvg.py:416-422
  rot_average            sunflower/utils/mvg.py:365-384   (pairwise SLERP, used by the aggregator)
``squarify_filter_batch`` is the vectorised form used by the predictors; it calls
the C-ABI ``flope_squarify_filter`` (integer, bit-exact with the scalar pair).
"""
import numpy as np


def squarify_bb(bb):
    """Grow the short side of an xyxy box to the long side, centred; an odd difference puts the extra pixel on the
    min side.  Returns four Python ints (truncation toward zero), like the reference."""
    x0, y0, x1, y1 = bb
    w, h = x1 - x0, y1 - y0
    d = abs(w - h)
    before = (d + 1) / 2 if d % 2 else d / 2        # taken from the min coordinate
    after = d - before                               # added to the max coordinate
    if w > h:
        y0, y1 = y0 - before, y1 + after
    elif h > w:
        x0, x1 = x0 - before, x1 + after
    return [int(x0), int(y0), int(x1), int(y1)]


def bb_in_frame(bb, img_shape):
    """True when the box lies inside the (h, w, ...) image; touching the right / bottom edge is inside (slices are exclusive)."""
    x0, y0, x1, y1 = bb
    return not (x0 < 0 or y0 < 0 or x1 > img_shape[1] or y1 > img_shape[0])


def filter_very_large_bb(bb_dino):
    """Drop the boxes whose area exceeds five times the median area (GroundingDINO's whole-plant boxes)."""
    boxes = np.array(bb_dino)
    area = np.prod(boxes[:, 2:4] - boxes[:, 0:2], axis=1)
    return boxes[area <= 5 * np.median(area)]


def squarify_filter_batch(boxes, img_shape):
    """(N,4) integer xyxy -> (square boxes (M,4) int32, keep (N,) bool) through the C-ABI."""
    from . import _lib
    boxes = np.ascontiguousarray(np.asarray(boxes).reshape(-1, 4), dtype=np.int32)
    return _lib.squarify_filter(boxes, int(img_shape[0]), int(img_shape[1]))


def nullify_yaw_batch(rotmat):
    """(N,3,3) rotations (numpy) -> float64 (N,3,3) with the 'zyx' yaw angle zeroed.

    Same contract as the reference; computed by the fused pose-head kernel in fp64
    (R' = Rx(gamma) Ry(beta), the closed form of SciPy's as_euler/from_euler round trip).
    """
    import torch
    from .conversion import nullify_yaw_batch_cuda
    r = torch.as_tensor(np.asarray(rotmat, dtype=np.float32)).cuda()
    return nullify_yaw_batch_cuda(r).cpu().numpy()


def get_points3d(uv, Zray, K):
    """(N,2) pixel coordinates, (N,) ray lengths in metres, (3,3) intrinsics -> (N,3) camera-frame points: the pixel's
    viewing ray K^-1 [u v 1]^T scaled to the measured length."""
    px = np.concatenate([np.asarray(uv, dtype=np.float64), np.ones((len(uv), 1))], axis=1)
    rays = (np.linalg.inv(K) @ px.T).T
    return rays * (np.asarray(Zray, dtype=np.float64) / np.linalg.norm(rays, axis=1))[:, None]


def pose_cam_to_world(obj_pose, cam_pose):
    """(N,4,4) object poses in the camera frame, (4,4) camera pose -> (N,4,4) in the world frame."""
    return cam_pose @ obj_pose


def rot_average(quat1, quat2, weight1, weight2):
    """Row-wise weighted mean of two sets of xyzw quaternions on the geodesic between them: the rotation a fraction
    w2 / (w1 + w2) of the way from quat1 to quat2.  All rows at once through rotation vectors."""
    from scipy.spatial.transform import Rotation as R
    t = np.asarray(weight2, dtype=np.float64) / (np.asarray(weight1, dtype=np.float64) + np.asarray(weight2, dtype=np.float64))
    start = R.from_quat(np.asarray(quat1))
    step = (start.inv() * R.from_quat(np.asarray(quat2))).as_rotvec()
    return (start * R.from_rotvec(step * t[:, None])).as_quat().reshape(-1, 4)
