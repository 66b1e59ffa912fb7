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
    def __init__(self, th=40, score_th=200):
        """th: distance in mm below which two measurements are the same flower; score_th: observations needed."""
        self.th = th / 1000
        self.score_th = score_th
        self.num = 0
        self.trans = None
        self.quat = None
        self.score = None
        self.all_new_trans = []
        self.all_new_quat = []

    def add_measurement(self, tvec, qvec):
        """tvec (N,3), qvec (N,4): match to the known flowers, average matched ones, append the rest."""
        from scipy.spatial.distance import cdist
        if self.trans is None:
            self.trans = tvec
            self.quat = qvec
            self.score = np.ones(tvec.shape[0])
            self.all_new_trans.append(tvec)
            self.all_new_quat.append(qvec)
            return
        distance_matrix = cdist(tvec, self.trans, metric='euclidean')
        min_idx = np.argmin(distance_matrix, axis=1)
        min_vals = np.min(distance_matrix, axis=1)
        good_match = min_vals < self.th
        min_idx_good = min_idx[good_match]
        tvec_good = tvec[good_match]
        qvec_good = qvec[good_match]
        state_score = self.score[min_idx_good]
        meas_score = np.ones(state_score.shape[0])
        normalizer = state_score + meas_score
        state_weight = state_score / normalizer
        meas_weight = meas_score / normalizer
        if min_idx_good.shape[0] == 0:
            self.trans = np.vstack((self.trans, tvec))
            self.quat = np.vstack((self.quat, qvec))
            self.score = np.concatenate((self.score, np.ones(tvec.shape[0])))
        else:
            self.trans[min_idx_good] = self.trans[min_idx_good] * state_weight.reshape(-1, 1) + tvec_good * meas_weight.reshape(-1, 1)
            self.quat[min_idx_good] = rot_average(self.quat[min_idx_good], qvec_good, state_weight, meas_weight)
            new_trans = np.zeros_like(self.trans)
            new_trans[min_idx_good] = tvec_good
            self.all_new_trans.append(new_trans)
            new_quat = np.zeros_like(self.quat)
            new_quat[min_idx_good] = qvec_good
            self.all_new_quat.append(new_quat)
            self.score[min_idx_good] += 1
            unmatched = np.logical_not(good_match)
            self.trans = np.vstack((self.trans, tvec[unmatched]))
            self.quat = np.vstack((self.quat, qvec[unmatched]))
            self.score = np.concatenate((self.score, np.ones(int(unmatched.sum()))))

    def get_final_data(self):
        score_filter = self.score > self.score_th
        return self.trans[score_filter], self.quat[score_filter]

    def save_filtered_data(self, path='filtered_data.pkl'):
        with open(path, 'wb') as fp:
            pickle.dump({'trans': self.trans, 'quat': self.quat, 'score': self.score}, fp)

    def save_measurements(self, path='meas.pkl'):
        with open(path, 'wb') as fp:
            pickle.dump({'trans': self.all_new_trans, 'quat': self.all_new_quat}, fp)
