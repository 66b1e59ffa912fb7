"""Oracle: the composed pose path (test infrastructure only).

frame + mask + detector boxes -> squarify/filter -> crop batch -> PoseNet ->
Procrustes -> yaw nullification -> Rt, following
sunflower/predictor/fast_pose_predictor.py:66-82,108-156 (the depth/translation
branch :90-105 is out of scope; callers pass ``xyz`` or get zeros).
"""
import numpy as np
import torch

from . import boxes as obox
from . import posenet as onet
from . import resize as ores
from . import rotation as orot


def run(model, frame, mask, det_boxes, size=512, interp=ores.LANCZOS4, nullify_yaw=True, xyz=None):
    """Returns dict(sq_boxes, keep, r9, rot, rot_yaw, Rt) or None when no box survives."""
    sq, keep = obox.squarify_filter(det_boxes, frame.shape)
    if sq.shape[0] == 0:
        return None
    batch = ores.crop_batch_reference(frame, mask, sq, size=size, interp=interp)
    r9 = onet.forward_fp32(model, torch.from_numpy(batch))
    rot = orot.procrustes_to_rotmat(r9).numpy()
    rot_out = orot.nullify_yaw_batch(rot) if nullify_yaw else rot.astype(np.float64)
    return dict(sq_boxes=sq, keep=keep, crops=batch, r9=r9.numpy(), rot=rot, rot_yaw=rot_out,
                Rt=orot.assemble_rt(rot_out, xyz))
