"""Tracking consumer of the pose path, mirroring sunflower/predictor/flower_model.py.

  get_kalman_filter        flower_model.py:18-26   (7-state constant model: F = H = P = I, Q = 0.001 I, R = 0.1 I)
  FlowerModel.add_data     flower_model.py:219-255 (camera pose -> 4x4, get_flower_poses, camera -> world, rotation ->
                           quaternion, measurement rows [x y z qx qy qz qw])
  FlowerModel.assign_meas_to_state
                           flower_model.py:146-216 (nearest-neighbour association under dist_th, Kalman predict + update
                           of matched states, quaternion re-normalisation, unmatched measurements become new states)
Host-side float64 like the reference; the per-measurement Python loop of the reference is replaced by stacked (batched)
7x7 filter updates over all matched flowers of a frame.  The poses come from the GPU predictor.
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


class _FilterBank:
    """All flowers' 7-state Kalman filters as stacked arrays: x (M,7), P (M,7,7), shared F, H, Q, R (get_kalman_filter's
    model).  step(idx, z) runs predict + update + quaternion re-normalisation for the filters idx (distinct) at once."""

    def __init__(self):
        proto = get_kalman_filter(np.zeros(7))
        self.F, self.H, self.Q, self.R = proto.F, proto.H, proto.Q, proto.R
        self.P0 = proto.P
        self.x = np.zeros((0, 7))
        self.P = np.zeros((0, 7, 7))

    def __len__(self):
        return self.x.shape[0]

    def add(self, meas):
        meas = np.asarray(meas, dtype=np.float64).reshape(-1, 7)
        self.x = np.concatenate([self.x, meas])
        self.P = np.concatenate([self.P, np.tile(self.P0, (meas.shape[0], 1, 1))])

    def step(self, idx, z):
        F, H, Q, R = self.F, self.H, self.Q, self.R
        x = self.x[idx] @ F.T
        P = F @ self.P[idx] @ F.T + Q
        y = z - x @ H.T
        PHT = P @ H.T
        K = PHT @ np.linalg.inv(H @ PHT + R)
        x = x + np.einsum('nij,nj->ni', K, y)
        I_KH = np.eye(7) - K @ H
        P = I_KH @ P @ np.transpose(I_KH, (0, 2, 1)) + K @ R @ np.transpose(K, (0, 2, 1))
        x[:, 3:] /= np.linalg.norm(x[:, 3:], axis=1, keepdims=True)
        self.x[idx], self.P[idx] = x, P


class _FilterView:
    """kfs[i] of the reference (an object with .x and .P) as a view into the bank."""

    def __init__(self, bank, i):
        self._b, self._i = bank, i

    x = property(lambda self: self._b.x[self._i])
    P = property(lambda self: self._b.P[self._i])


class FlowerModel:
    def __init__(self, dist_th=50, intrin_path=None, get_plots=False, *, pose_predictor=None):
        """dist_th in mm.  ``pose_predictor`` is a flope_b200.predictor.PosePredictor (the reference builds one from
        hard-coded weight paths); it is only needed by add_data."""
        self.get_plots = False                     # live plotting is outside the path
        self.state = None
        self.scores = None
        self._bank = _FilterBank()
        self.th = dist_th / 1000
        self.intrin_path = intrin_path
        self.pose_predictor = pose_predictor

    @property
    def kfs(self):
        return [_FilterView(self._bank, i) for i in range(len(self._bank))]

    def assign_meas_to_state(self, meas):
        """One frame of measurements (N,7) = [x y z qx qy qz qw].  Each measurement is associated with the nearest flower
        of the state list AS IT WAS when the frame arrived (distance on the translation part, first-seen positions, like
        the reference); under dist_th it drives that flower's filter (predict, update, renormalise the quaternion) and
        raises its score, otherwise it founds a new flower.  Measurements that share a flower are applied in their order:
        the batch is split into rounds in which every flower occurs at most once, each round is one stacked update."""
        from scipy.spatial.distance import cdist
        meas = np.asarray(meas, dtype=np.float64)
        if self.state is None:
            self.state = meas
            self.scores = np.ones(meas.shape[0])
            self._bank.add(meas)
            return
        dist = cdist(meas[:, :3], self.state[:, :3], metric='euclidean')
        nearest = dist.argmin(axis=1)
        known = dist[np.arange(meas.shape[0]), nearest] < self.th
        rows, flowers = np.flatnonzero(known), nearest[known]
        if rows.size:
            # occurrence number of each row among the rows of its flower: 0 for the first claimant, 1 for the second ...
            order = np.argsort(flowers, kind='stable')
            sorted_f = flowers[order]
            start = np.r_[0, np.flatnonzero(np.diff(sorted_f)) + 1]
            occ = np.empty(rows.size, dtype=np.int64)
            occ[order] = np.arange(rows.size) - np.repeat(start, np.diff(np.r_[start, rows.size]))
            for k in range(int(occ.max()) + 1):
                sel = occ == k
                self._bank.step(flowers[sel], meas[rows[sel]])
            np.add.at(self.scores, flowers, 1)
        fresh = meas[~known]
        if fresh.shape[0]:
            self.state = np.vstack((self.state, fresh))
            self.scores = np.hstack((self.scores, np.ones(fresh.shape[0])))
            self._bank.add(fresh)

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
        return self._bank.x.copy()
