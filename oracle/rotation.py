"""Oracle: 9-vector -> SO(3) -> yaw-nullified rotation -> Rt (test infrastructure only).

* ``special_procrustes`` restates ``roma.special_procrustes`` (roma==1.5.1,
  environment.yml:214), the third-party call behind
  sunflower/utils/conversion.py:54-58.  roma is absent from /root/reference and
  from this image and the reference has no test or golden vector for it:
  PARITY UNPINNED for this function (restated from its documented behaviour:
  the rotation nearest to M in Frobenius norm, R = U diag(1,1,det(U V^T)) V^T).
* ``nullify_yaw_batch`` restates sunflower/utils/mvg.py:240-251 with the same
  SciPy calls as sunflower/utils/conversion.py:45-51.
* ``nullify_yaw_closed_form`` is the algebraic identity R' = Rx(g) Ry(b) used by
  the fused CUDA head; tests pin it against the SciPy path.
"""
import numpy as np
import torch
from scipy.spatial.transform import Rotation as sciR


def special_procrustes(m):
    """(…,3,3) torch tensor -> nearest rotation matrices (det = +1)."""
    m = torch.as_tensor(m)
    u, _, vh = torch.linalg.svd(m)
    d = torch.det(u @ vh)
    s = torch.ones(m.shape[:-1], dtype=m.dtype)
    s[..., 2] = torch.where(d < 0, -torch.ones_like(d), torch.ones_like(d))
    return (u * s[..., None, :]) @ vh


def procrustes_to_rotmat(inp):
    """sunflower/utils/conversion.py:54-58 - row-major reshape of the 9 outputs, then Procrustes."""
    return special_procrustes(torch.as_tensor(inp).reshape(-1, 3, 3))


def R2E(R):   # conversion.py:45-47
    return sciR.from_matrix(R).as_euler('zyx', degrees=True)


def E2R(E):   # conversion.py:49-51
    return sciR.from_euler('zyx', E, degrees=True).as_matrix()


def nullify_yaw_batch(rotmat):
    """sunflower/utils/mvg.py:240-251 - returns float64 (N,3,3)."""
    e = R2E(np.asarray(rotmat))
    e[:, 0] = 0.0
    return E2R(e)


def nullify_yaw_closed_form(R):
    """R' = Rx(gamma) Ry(beta), beta = atan2(R02, hypot(R00,R01)), gamma = atan2(-R12, R22)."""
    R = np.asarray(R, dtype=np.float64)
    beta = np.arctan2(R[:, 0, 2], np.hypot(R[:, 0, 0], R[:, 0, 1]))
    gamma = np.arctan2(-R[:, 1, 2], R[:, 2, 2])
    cb, sb, cg, sg = np.cos(beta), np.sin(beta), np.cos(gamma), np.sin(gamma)
    z = np.zeros_like(cb)
    return np.stack([np.stack([cb, z, sb], -1),
                     np.stack([sg * sb, cg, -sg * cb], -1),
                     np.stack([-cg * sb, sg, cg * cb], -1)], -2)


def assemble_rt(rot, xyz=None):
    """pose_predictor.py:172-174 - (N,4,4) float64 with rotation and translation."""
    n = rot.shape[0]
    rt = np.repeat(np.eye(4)[None], n, axis=0)
    rt[:, :3, :3] = rot
    if xyz is not None:
        rt[:, :3, 3] = xyz
    return rt


def geodesic_deg(Ra, Rb):
    """Angle of Ra^T Rb in degrees, per sample (chordal form: well conditioned near 0)."""
    Ra = np.asarray(Ra, dtype=np.float64)
    Rb = np.asarray(Rb, dtype=np.float64)
    chord = np.linalg.norm((Ra - Rb).reshape(-1, 9), axis=1)          # = 2*sqrt(2)*sin(theta/2)
    return np.degrees(2.0 * np.arcsin(np.clip(chord / (2.0 * np.sqrt(2.0)), 0.0, 1.0)))
