"""Oracle: the composed pose path (test infrastructure only).

frame + mask + detector boxes -> squarify/filter -> crop batch -> PoseNet ->
Procrustes -> yaw nullification -> Rt, following
sunflower/predictor/fast_pose_predictor.py:66-82,108-156; with ``depth`` and ``K`` also the
depth / translation branch :90-105 (get_depth_value -> reliability filter -> get_points3d).
"""
import numpy as np
import torch

from . import boxes as obox
from . import depth as odepth
from . import posenet as onet
from . import resize as ores
from . import rotation as orot


def run(model, frame, mask, det_boxes, size=512, interp=ores.LANCZOS4, nullify_yaw=True, xyz=None, depth=None, K=None,
        depth_scale=1000.0, near_plane=0.1, far_plane=2.5):
    """Returns dict(sq_boxes, keep, r9, rot, rot_yaw, Rt) or None when no box survives."""
    sq, keep = obox.squarify_filter(det_boxes, frame.shape)
    if sq.shape[0] == 0:
        return None
    reliable = None
    if depth is not None and K is not None:
        good = np.asarray(det_boxes)[keep].astype(np.int16)                  # fast_pose_predictor.py:81
        val, reliable = odepth.get_depth_value(good, depth.astype(np.float32) / depth_scale, mask,
                                               near_plane=near_plane, far_plane=far_plane)
        sq = sq[reliable]
        if sq.shape[0] == 0:
            return None
        xyz = odepth.get_points3d(odepth.box_centres(good)[reliable], val[reliable], K)
    batch = ores.crop_batch_reference(frame, mask, sq, size=size, interp=interp)
    r9 = onet.forward_fp32(model, torch.from_numpy(batch))
    rot = orot.procrustes_to_rotmat(r9).numpy()
    rot_out = orot.nullify_yaw_batch(rot) if nullify_yaw else rot.astype(np.float64)
    return dict(sq_boxes=sq, keep=keep, reliable=reliable, crops=batch, r9=r9.numpy(), rot=rot, rot_yaw=rot_out,
                Rt=orot.assemble_rt(rot_out, xyz))
