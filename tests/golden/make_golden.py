"""Generate the golden fixtures in this directory from the REAL reference code.

Run once in the authoring container (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference with two shims that the
environment forces:
  * ``sys.modules['roma']`` stub providing ``special_procrustes`` via torch SVD
    (roma is not installed; this makes ``sunflower.utils.conversion`` importable);
  * ``torchvision.models.resnet18`` patched to ``weights=None`` (no network).
PoseNet is run in eval() under no_grad().

Nothing here is copied from the reference: the fixtures are its *outputs* on
seeded inputs.  The GPU box has no /root/reference, so tests only read the .npz.
"""
import hashlib
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def _roma_stub():
    m = types.ModuleType("roma")

    def special_procrustes(M):
        U, _, Vh = torch.linalg.svd(M)
        d = torch.det(U @ Vh)
        S = torch.ones(M.shape[:-1], dtype=M.dtype)
        S[..., 2] = torch.where(d < 0, -torch.ones_like(d), torch.ones_like(d))
        return (U * S[..., None, :]) @ Vh

    m.special_procrustes = special_procrustes
    return m


def main():
    sys.modules["roma"] = _roma_stub()
    import torchvision.models as tvm
    from sunflower.utils import mvg as ref_mvg
    from sunflower.utils import conversion as ref_conv
    import sunflower.models.posenet as ref_posenet

    rng = np.random.default_rng(20260101)

    # ---- boxes: squarify_bb / bb_in_frame / filter_very_large_bb (mvg.py:324-362) ----
    H, W = 1080, 1920
    n = 4000
    w = rng.integers(1, 700, n)
    h = rng.integers(1, 700, n)
    x0 = rng.integers(-40, W, n)
    y0 = rng.integers(-40, H, n)
    boxes = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.int16)        # YOLO path dtype (fast_pose_predictor.py:56)
    # force the interesting edges in: equal sides, diff 1, xmax == W, ymax == H, touching 0
    boxes[:8] = [[0, 0, 10, 10], [0, 0, 11, 10], [0, 5, 10, 6], [W - 50, 10, W, 61], [10, H - 50, 61, H],
                 [5, 0, 6, 100], [W - 101, 0, W, 100], [0, H - 100, 101, H]]
    sq = np.array([ref_mvg.squarify_bb(b) for b in boxes], dtype=np.int64)
    keep = np.array([ref_mvg.bb_in_frame(s, (H, W, 3)) for s in sq], dtype=bool)
    small = boxes[:64].astype(np.int64)
    small[5] = [0, 0, 1500, 1000]
    flt = ref_mvg.filter_very_large_bb(small)
    np.savez_compressed(os.path.join(HERE, "boxes.npz"), boxes=boxes, frame_hw=np.array([H, W]),
                        squarified=sq, keep=keep, vlb_in=small, vlb_out=flt)

    # ---- yaw nullification (mvg.py:240-251) on random rotations ----
    from scipy.spatial.transform import Rotation as sciR
    Rr = sciR.random(256, random_state=5).as_matrix().astype(np.float32)
    yaw = ref_mvg.nullify_yaw_batch(Rr.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "yaw.npz"), R=Rr, R_yaw_nullified=yaw)

    # ---- PoseResNet (posenet.py:5-34), seed 0, eval, fp32 CPU ----
    real_resnet18 = tvm.resnet18
    with mock.patch.object(tvm, "resnet18", lambda *a, **k: real_resnet18(weights=None)):
        torch.manual_seed(0)
        net = ref_posenet.PoseResNet().eval()
    sd = net.state_dict()
    digest = hashlib.sha256()
    for k in sd:
        digest.update(k.encode())
        digest.update(sd[k].numpy().tobytes())
    from flope_b200 import synth
    out = {}
    with torch.no_grad():
        for size, nb in ((224, 8), (512, 2)):
            x = synth.mixed_crops(nb, size)
            r9 = net(x)
            rot = ref_conv.procrustes_to_rotmat(r9)          # real reshape + stubbed roma
            out[f"r9_{size}"] = r9.numpy()
            out[f"rot_{size}"] = rot.numpy()
            out[f"in_sha_{size}"] = np.frombuffer(hashlib.sha256(x.numpy().tobytes()).digest(), np.uint8)
    out["state_sha"] = np.frombuffer(digest.digest(), np.uint8)
    out["state_keys"] = np.array(list(sd.keys()))
    out["state_shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
    out["probe_weights"] = np.concatenate([sd["base.conv1.weight"].flatten()[:16].numpy(),
                                           sd["base.fc.0.weight"].flatten()[:16].numpy(),
                                           sd["fc_rot.bias"].numpy()])
    np.savez_compressed(os.path.join(HERE, "posenet_seed0.npz"), **out)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
