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
    xmin, ymin, xmax, ymax = bb
    xrange_ = xmax - xmin
    yrange_ = ymax - ymin
    diff = abs(xrange_ - yrange_)
    if diff % 2 == 0:
        decrease_min = increase_max = diff / 2
    else:
        decrease_min = (diff + 1) / 2
        increase_max = (diff - 1) / 2
    if xrange_ > yrange_:
        ymin -= decrease_min
        ymax += increase_max
    elif xrange_ < yrange_:
        xmin -= decrease_min
        xmax += increase_max
    return [int(xmin), int(ymin), int(xmax), int(ymax)]


def bb_in_frame(bb, img_shape):
    h, w = img_shape[0], img_shape[1]
    xmin, ymin, xmax, ymax = bb
    if xmin < 0 or ymin < 0 or xmax > w or ymax > h:
        return False
    return True


def filter_very_large_bb(bb_dino):
    bb_dino = np.array(bb_dino)
    x_range = bb_dino[:, 2] - bb_dino[:, 0]
    y_range = bb_dino[:, 3] - bb_dino[:, 1]
    area = x_range * y_range
    large_area = area > 5 * np.median(area)
    return bb_dino[np.logical_not(large_area)]


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
    """(N,2) pixel coordinates, (N,) ray lengths in metres, (3,3) intrinsics -> (N,3) camera-frame points."""
    uv = np.asarray(uv, dtype=np.float64)
    N = uv.shape[0]
    uv1 = np.hstack((uv, np.ones(N).reshape(-1, 1)))
    xnyn1 = (np.linalg.inv(K) @ uv1.T).T
    xnyn1_norm = np.linalg.norm(xnyn1, axis=1)
    Z = np.asarray(Zray) / xnyn1_norm
    return xnyn1 * Z.reshape(-1, 1)


def pose_cam_to_world(obj_pose, cam_pose):
    """(N,4,4) object poses in the camera frame, (4,4) camera pose -> (N,4,4) in the world frame."""
    return cam_pose @ obj_pose


def rot_average(quat1, quat2, weight1, weight2):
    """Row-wise weighted average of two sets of xyzw quaternions by spherical interpolation at t = w2 / (w1 + w2)."""
    from scipy.spatial.transform import Rotation as R, Slerp
    avg_quat = []
    for q1, q2, w1, w2 in zip(quat1, quat2, weight1, weight2):
        slerp = Slerp([0, 1], R.concatenate([R.from_quat(q1), R.from_quat(q2)]))
        avg_quat.append(slerp([w2 / (w1 + w2)]).as_quat()[0])
    return np.array(avg_quat)
