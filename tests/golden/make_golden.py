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
    # ---- depth branch: get_depth_value (image_manipulation.py:39-96) + get_points3d (mvg.py:387-408) ----
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "matplotlib.colors", "plotly",
              "plotly.graph_objects", "plotly.express", "plotly.subplots", "plyfile", "icecream", "mpl_toolkits",
              "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(m, mock.MagicMock())          # plotting / IO modules image_manipulation.py drags in
    from sunflower.utils import image_manipulation as ref_im
    drng = np.random.default_rng(77)
    dH, dW = 360, 640
    yy, xx = np.mgrid[0:dH, 0:dW]
    raw = (3000 + 2500 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + drng.integers(-40, 40, (dH, dW))).astype(np.uint16)
    raw[drng.random((dH, dW)) < 0.02] = 0                                   # sensor holes
    raw[:, 600:] = 40000                                                    # beyond the far plane
    dmask = np.zeros((dH, dW), np.uint8)
    for cx, cy, rad in ((120, 100, 60), (330, 200, 90), (520, 90, 45), (60, 300, 30), (600, 330, 25), (250, 40, 9)):
        dmask[(xx - cx) ** 2 + (yy - cy) ** 2 <= rad * rad] = 255
    dmask[drng.random((dH, dW)) < 0.003] = 0                                # pinholes: erosion grows them
    dboxes = np.array([[60, 40, 180, 160], [240, 110, 420, 290], [475, 45, 565, 135], [30, 270, 90, 330],
                       [575, 305, 625, 355], [241, 31, 259, 49], [0, 0, 40, 40], [300, 170, 360, 230],
                       [0, 0, 640, 360], [610, 0, 640, 30]], np.int16)
    Kmat = np.array([[430.0, 0, 318.5], [0, 431.5, 181.2], [0, 0, 1]])
    dout = {"raw": raw, "mask": dmask, "boxes": dboxes, "K": Kmat}
    for tag, div, far in (("pose", 10000.0, 2.5), ("fast", 1000.0, 3.0)):
        depth_m = raw.astype(np.float32) / div                              # pose_predictor.py:118 / fast_pose_predictor.py:90
        val, rel, _ = ref_im.get_depth_value(dboxes, depth_m.copy(), dmask, near_plane=0.1, far_plane=far)
        uv = np.stack([(dboxes[:, 2].astype(np.float64) + dboxes[:, 0]) / 2, (dboxes[:, 3].astype(np.float64) + dboxes[:, 1]) / 2], 1)
        dout[f"val_{tag}"] = np.asarray(val, np.float64)
        dout[f"rel_{tag}"] = rel
        dout[f"xyz_{tag}"] = ref_mvg.get_points3d(uv, np.asarray(val, np.float64), Kmat)
        dout[f"div_{tag}"] = np.array(div)
        dout[f"far_{tag}"] = np.array(far)
    seg = np.logical_and(dmask > 128, np.logical_and(raw.astype(np.float32) / 10000.0 > 0.1, raw.astype(np.float32) / 10000.0 < 2.5))
    dout["eroded_pose"] = ref_im.shrink_mask(seg, 10)
    np.savez_compressed(os.path.join(HERE, "depth.npz"), **dout)
    # ---- YOLO-seg post-processing: the real FastPosePredictor.get_bbox_mask (fast_pose_predictor.py:44-57) on an
    #      instance whose detector is a stub (ultralytics and the other heavy imports are absent here) ----
    for m in ("ultralytics", "hydra", "omegaconf", "filterpy", "filterpy.kalman", "shapely", "shapely.geometry", "tyro"):
        sys.modules.setdefault(m, mock.MagicMock())
    from sunflower.predictor.fast_pose_predictor import FastPosePredictor as RefFast
    yrng = np.random.default_rng(31)
    yout = {}
    for tag, (h, w, H, W, n) in {"a": (384, 640, 1080, 1920, 5), "b": (160, 256, 360, 640, 3), "c": (96, 128, 97, 131, 2),
                                 "d": (384, 640, 192, 320, 4)}.items():
        yy, xx = np.mgrid[0:h, 0:w]
        masks = np.zeros((n, h, w), np.float32)
        for k in range(n):
            cx, cy, rad = yrng.integers(0, w), yrng.integers(0, h), yrng.integers(5, h // 3)
            masks[k][(xx - cx) ** 2 + (yy - cy) ** 2 <= rad * rad] = 1.0
        boxes = (yrng.random((n, 4)) * [W, H, W, H]).astype(np.float32)
        boxes[0] = [0.0, 0.99, W - 0.01, H - 0.5]
        res = types.SimpleNamespace(masks=types.SimpleNamespace(data=torch.from_numpy(masks)),
                                    boxes=types.SimpleNamespace(xyxy=torch.from_numpy(boxes)))
        inst = object.__new__(RefFast)
        inst.yolo = lambda img, _r=res: [_r]
        bbox, mask = inst.get_bbox_mask(np.zeros((H, W, 3), np.uint8))
        yout[f"masks_{tag}"] = masks.astype(np.uint8)
        yout[f"boxes_{tag}"] = boxes
        yout[f"bbox_{tag}"] = bbox
        yout[f"mask_{tag}"] = mask
    np.savez_compressed(os.path.join(HERE, "yolo_post.npz"), **yout)
    # ---- aggregation: the reference's own Env3D (scripts/flower_pose_aggregrator.py:23-135) and rot_average
    #      (mvg.py:365-384) on a seeded sequence of noisy re-observations of 12 flowers ----
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_aggregator", "/root/reference/scripts/flower_pose_aggregrator.py")
    ref_agg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_agg)
    arng = np.random.default_rng(55)
    true_t = arng.uniform(-0.5, 0.5, (12, 3))
    true_q = sciR.random(12, random_state=8).as_quat()
    env = ref_agg.Env3D(th=40, score_th=3)
    frames_t, frames_q = [], []
    for f in range(14):
        seen = arng.random(12) < 0.7
        seen[arng.integers(0, 12)] = True
        t = true_t[seen] + arng.normal(0, 0.004, (int(seen.sum()), 3))
        q = (sciR.from_quat(true_q[seen]) * sciR.from_rotvec(arng.normal(0, 0.05, (int(seen.sum()), 3)))).as_quat()
        if f == 5:                                          # a frame of only-new flowers far away: the no-match branch
            t, q = t + 3.0, q
        order = arng.permutation(t.shape[0])
        frames_t.append(t[order]); frames_q.append(q[order])
        env.add_measurement(frames_t[-1].copy(), frames_q[-1].copy())
    fin_t, fin_q = env.get_final_data()
    aout = {"n_frames": np.array(14), "trans": env.trans, "quat": env.quat, "score": env.score, "final_trans": fin_t,
            "final_quat": fin_q}
    for f in range(14):
        aout[f"t{f}"] = frames_t[f]; aout[f"q{f}"] = frames_q[f]
    q1, q2 = sciR.random(20, random_state=1).as_quat(), sciR.random(20, random_state=2).as_quat()
    w1, w2 = arng.uniform(0.1, 5, 20), arng.uniform(0.1, 5, 20)
    aout.update(ra_q1=q1, ra_q2=q2, ra_w1=w1, ra_w2=w2, ra_out=ref_mvg.rot_average(q1, q2, w1, w2))
    det = np.hstack([arng.integers(0, 600, (7, 4)).astype(np.float64), arng.uniform(0, 600, (7, 2)), arng.normal(0, 1, (7, 9))])
    aout["det_rows"] = det
    np.savez_compressed(os.path.join(HERE, "aggregate.npz"), **aout)
    # ---- tracking: the reference's own FlowerModel.assign_meas_to_state (flower_model.py:146-216) with filterpy's
    #      KalmanFilter replaced by the restatement in flope_b200/flower_model.py (filterpy is absent here: the filter
    #      arithmetic itself stays unpinned, the association / update-order / normalisation logic is the reference's) ----
    from flope_b200 import flower_model as our_fm
    fk = types.ModuleType("filterpy.kalman")
    fk.KalmanFilter = our_fm.KalmanFilter
    sys.modules["filterpy"] = types.ModuleType("filterpy")
    sys.modules["filterpy.kalman"] = fk
    sys.modules.pop("sunflower.predictor.flower_model", None)
    sys.modules["matplotlib.ticker"] = mock.MagicMock()
    sys.modules["icecream"] = types.SimpleNamespace(ic=lambda *a, **k: None)
    from sunflower.predictor import flower_model as ref_fm
    ref_fm.ic = lambda *a, **k: None
    fm = object.__new__(ref_fm.FlowerModel)
    fm.get_plots, fm.state, fm.scores, fm.kfs, fm.th = False, None, None, [], 50 / 1000
    trng = np.random.default_rng(91)
    t_true = trng.uniform(-0.4, 0.4, (9, 3))
    q_true = sciR.random(9, random_state=12).as_quat()
    tout = {"n_frames": np.array(12)}
    for f in range(12):
        seen = trng.random(9) < 0.75
        seen[trng.integers(0, 9)] = True
        t = t_true[seen] + trng.normal(0, 0.006, (int(seen.sum()), 3))
        q = (sciR.from_quat(q_true[seen]) * sciR.from_rotvec(trng.normal(0, 0.04, (int(seen.sum()), 3)))).as_quat()
        meas = np.hstack([t, q])[trng.permutation(int(seen.sum()))]
        tout[f"m{f}"] = meas
        fm.assign_meas_to_state(meas.copy())
    tout.update(state=fm.state, scores=fm.scores, kf_x=np.array([k.x for k in fm.kfs]), kf_P=np.array([k.P for k in fm.kfs]))
    np.savez_compressed(os.path.join(HERE, "tracking.npz"), **tout)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
