"""Detection files and multi-frame aggregation, mirroring scripts/flower_pose_aggregrator.py.

  read_detection_txt      the reader of the 15-column ``detection/*.txt`` rows that predictor.write_detection_txt /
                          scripts/test_posenet.py:150-161 write (scripts/flower_pose_aggregrator.py:183,204-206)
  frame_measurements      one frame's detections -> world-frame translations + quaternions
                          (scripts/flower_pose_aggregrator.py:189-232); the depth values come from the device-side
                          get_depth_value mirror
  Env3D                   nearest-neighbour re-identification with score-weighted running averages
                          (scripts/flower_pose_aggregrator.py:23-135)
Host code in float64 like the reference: a few dozen flowers per frame, sequentially dependent between frames - there is
nothing here for a GPU to do beyond the depth lookup.
"""
import pickle

import numpy as np

from .mvg import get_points3d, pose_cam_to_world, rot_average


def read_detection_txt(path):
    """-> (bbox (N,4) int16, uv (N,2) float64, rotmat (N,9) float64)   [flower_pose_aggregrator.py:183,204-206]"""
    det = np.loadtxt(path).reshape(-1, 15)
    return det[:, :4].astype(np.int16), det[:, 4:6], det[:, 6:]


def get_pose_mat(trans_rot):
    """(N,12) = 3 translation + 9 rotation entries -> (N,4,4)   [sunflower/utils/conversion.py:61-76]"""
    trans_rot = np.asarray(trans_rot, dtype=np.float64).reshape(-1, 12)
    pose = np.tile(np.eye(4), (trans_rot.shape[0], 1, 1))
    pose[:, :3, 3] = trans_rot[:, :3]
    pose[:, :3, :3] = trans_rot[:, 3:].reshape(-1, 3, 3)
    return pose


def frame_measurements(det_path_or_arrays, depth, seg_mask, cam_pose_4x4, K, scale=None, near_plane=0.1, far_plane=2.5,
                       device=None):
    """One frame -> (trans (M,3), quat (M,4) xyzw) in the world frame, or None when no detection has reliable depth."""
    from scipy.spatial.transform import Rotation as sciR
    from .image_manipulation import get_depth_value
    if isinstance(det_path_or_arrays, (str, bytes)) or hasattr(det_path_or_arrays, "__fspath__"):
        bbox, uv, rotmat = read_detection_txt(det_path_or_arrays)
    else:
        bbox, uv, rotmat = det_path_or_arrays
    depth_vals, good, _ = get_depth_value(bbox, depth, seg_mask, scale, near_plane, far_plane, False, device=device)
    depth_vals, uv, rotmat = depth_vals[good], np.asarray(uv)[good], np.asarray(rotmat)[good]
    if depth_vals.shape[0] == 0:
        return None
    points3d_cam = get_points3d(uv, depth_vals, K)
    pose_mat = pose_cam_to_world(get_pose_mat(np.hstack((points3d_cam, rotmat))), cam_pose_4x4)
    return pose_mat[:, :3, 3], sciR.from_matrix(pose_mat[:, :3, :3]).as_quat()


class Env3D:
    """Re-identification of flowers across frames by nearest neighbour in the world frame, with running means.

    A measurement closer than ``th`` (mm) to a known flower is merged into it - position by the weighted mean, orientation
    by rot_average, weights = observation count so far vs 1 - and the flower's score grows by one; the rest start new
    flowers with score 1.  get_final_data keeps the flowers seen more than ``score_th`` times.  When several measurements
    of one frame pick the same flower the LAST one is merged and the score still grows by one (the reference's numpy
    fancy assignment does exactly that), which is reproduced here by de-duplicating before one masked update.
    ``all_new_trans`` / ``all_new_quat`` log, per frame that matched anything, the merged measurements in flower order
    (zeros elsewhere) - the arrays scripts/align_measurements.py reads back."""

    def __init__(self, th=40, score_th=200):
        self.th = th / 1000
        self.score_th = score_th
        self.num = 0
        self.trans = self.quat = self.score = None
        self.all_new_trans, self.all_new_quat = [], []

    def _append(self, tvec, qvec):
        self.trans = np.vstack((self.trans, tvec))
        self.quat = np.vstack((self.quat, qvec))
        self.score = np.concatenate((self.score, np.ones(tvec.shape[0])))

    def add_measurement(self, tvec, qvec):
        """tvec (N,3) world-frame positions, qvec (N,4) xyzw orientations of one frame."""
        from scipy.spatial.distance import cdist
        if self.trans is None:
            self.trans, self.quat, self.score = tvec, qvec, np.ones(tvec.shape[0])
            self.all_new_trans.append(tvec)
            self.all_new_quat.append(qvec)
            return
        dist = cdist(tvec, self.trans, metric='euclidean')
        nearest = dist.argmin(axis=1)
        known = dist[np.arange(tvec.shape[0]), nearest] < self.th
        if known.any():
            # one measurement per flower: the last claimant
            rows = np.flatnonzero(known)
            flowers, first_from_end = np.unique(nearest[rows][::-1], return_index=True)
            rows = rows[::-1][first_from_end]
            seen = self.score[flowers]
            w_old, w_new = seen / (seen + 1.0), 1.0 / (seen + 1.0)
            self.trans[flowers] = self.trans[flowers] * w_old[:, None] + tvec[rows] * w_new[:, None]
            self.quat[flowers] = rot_average(self.quat[flowers], qvec[rows], w_old, w_new)
            self.score[flowers] += 1
            log_t, log_q = np.zeros_like(self.trans), np.zeros_like(self.quat)
            log_t[flowers], log_q[flowers] = tvec[rows], qvec[rows]
            self.all_new_trans.append(log_t)
            self.all_new_quat.append(log_q)
        self._append(tvec[~known], qvec[~known])

    def get_final_data(self):
        often = self.score > self.score_th
        return self.trans[often], self.quat[often]

    def save_filtered_data(self, path='filtered_data.pkl'):
        with open(path, 'wb') as fp:
            pickle.dump({'trans': self.trans, 'quat': self.quat, 'score': self.score}, fp)

    def save_measurements(self, path='meas.pkl'):
        with open(path, 'wb') as fp:
            pickle.dump({'trans': self.all_new_trans, 'quat': self.all_new_quat}, fp)
