"""Tracking consumer of the pose path, mirroring sunflower/predictor/flower_model.py.

  get_kalman_filter        flower_model.py:18-26   (7-state constant model: F = H = P = I, Q = 0.001 I, R = 0.1 I)
  FlowerModel.add_data     flower_model.py:219-255 (camera pose -> 4x4, get_flower_poses, camera -> world, rotation ->
                           quaternion, measurement rows [x y z qx qy qz qw])
  FlowerModel.assign_meas_to_state
                           flower_model.py:146-216 (nearest-neighbour association under dist_th, Kalman predict + update
                           of matched states, quaternion re-normalisation, unmatched measurements become new states)
Host-side float64 like the reference (a few dozen 7x7 filters per frame); the poses come from the GPU predictor.
The reference takes its Kalman filter from ``filterpy`` (1.4.5 pinned, absent here): ``KalmanFilter`` below restates
filterpy's published predict / update equations (Joseph-form covariance update) - PARITY UNPINNED for that arithmetic;
the association / update-order / normalisation logic around it is pinned against the reference class itself.
The live matplotlib plots of the reference (``get_plots``) are not reproduced.
"""
import numpy as np

from .mvg import pose_cam_to_world


class KalmanFilter:
    """The subset of filterpy.kalman.KalmanFilter the reference uses: x, F, H, P, Q, R, predict(), update(z)."""

    def __init__(self, dim_x, dim_z):
        self.dim_x, self.dim_z = dim_x, dim_z
        self.x = np.zeros(dim_x)
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.R = np.eye(dim_z)

    def predict(self):
        self.x = self.F @ self.x
        self.P = self.F @ self.P @ self.F.T + self.Q

    def update(self, z):
        z = np.asarray(z, dtype=np.float64).reshape(self.dim_z)
        y = z - self.H @ self.x
        PHT = self.P @ self.H.T
        S = self.H @ PHT + self.R
        K = PHT @ np.linalg.inv(S)
        self.x = self.x + K @ y
        I_KH = np.eye(self.dim_x) - K @ self.H
        self.P = I_KH @ self.P @ I_KH.T + K @ self.R @ K.T


def get_kalman_filter(initial_value):
    kf = KalmanFilter(dim_x=7, dim_z=7)
    kf.x = np.array(initial_value)
    kf.F = np.eye(7)
    kf.H = np.eye(7)
    kf.P = np.eye(7)
    kf.Q = np.eye(7) * 0.001
    kf.R = np.eye(7) * 0.1
    return kf


class FlowerModel:
    def __init__(self, dist_th=50, intrin_path=None, get_plots=False, *, pose_predictor=None):
        """dist_th in mm.  ``pose_predictor`` is a flope_b200.predictor.PosePredictor (the reference builds one from
        hard-coded weight paths); it is only needed by add_data."""
        self.get_plots = False                     # live plotting is outside the path
        self.state = None
        self.scores = None
        self.kfs = []
        self.th = dist_th / 1000
        self.intrin_path = intrin_path
        self.pose_predictor = pose_predictor

    def assign_meas_to_state(self, meas):
        from scipy.spatial.distance import cdist
        if self.state is None:
            self.state = meas
            self.scores = np.ones(meas.shape[0])
            for each_meas in meas:
                self.kfs.append(get_kalman_filter(each_meas))
            return
        distance_matrix = cdist(meas[:, :3], self.state[:, :3], metric='euclidean')
        min_idx = np.argmin(distance_matrix, axis=1)
        good_matches = np.min(distance_matrix, axis=1) < self.th
        for i in range(meas.shape[0]):
            measurement = meas[i]
            if good_matches[i]:
                kf = self.kfs[min_idx[i]]
                kf.predict()
                kf.update(measurement)
                kf.x[3:] /= np.linalg.norm(kf.x[3:])
                self.scores[min_idx[i]] += 1
            else:
                self.state = np.vstack((self.state, measurement.reshape(1, 7)))
                self.scores = np.hstack((self.scores, np.array([1])))
                self.kfs.append(get_kalman_filter(measurement))

    def add_data(self, rgb, depth, cam_pose, ignore=False):
        """rgb (H,W,3), depth (H,W), cam_pose (7,) = [translation, xyzw quaternion] ->
        (flower poses in the camera frame (N,4,4) float64, in the world frame (N,4,4) float32), or (None, None)."""
        from scipy.spatial.transform import Rotation as sciR
        cam_pose = np.asarray(cam_pose, dtype=np.float64)
        cam_posemat = np.eye(4)
        cam_posemat[:3, :3] = sciR.from_quat(cam_pose[3:]).as_matrix()
        cam_posemat[:3, 3] = cam_pose[:3]
        flower_pose_cam = self.pose_predictor.get_flower_poses(rgb, depth)
        if flower_pose_cam is None:
            return None, None
        flower_pose = pose_cam_to_world(flower_pose_cam, cam_posemat)
        flower_quat = sciR.from_matrix(flower_pose[:, :3, :3]).as_quat()
        meas = np.hstack((flower_pose[:, :3, 3], flower_quat))
        if ignore:                                  # sic: the reference only updates the tracker when `ignore` is set
            self.assign_meas_to_state(meas)
        return flower_pose_cam, flower_pose.astype(np.float32)

    def get_state(self):
        return self.state

    def get_filtered_state(self):
        """Current filter means (M,7); the reference leaves them inside self.kfs."""
        return np.array([kf.x for kf in self.kfs]) if self.kfs else np.zeros((0, 7))
