"""GPU parity of the depth branch (flope_depth_values / get_depth_value mirror / predictors with intrinsics) against
the reference's own outputs (tests/golden/depth.npz) and the oracle.  Validity, erosion and pixel counts are exact;
the per-box mean is fp64-accumulated on the device where numpy sums float32 pairwise: tolerance 2e-6 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import depth as od

pytestmark = pytest.mark.gpu
RTOL = 2e-6


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "depth.npz"))


@pytest.mark.parametrize("tag", ["pose", "fast"])
@pytest.mark.parametrize("as_u16", [True, False])
def test_depth_values_match_reference_golden(cuda_lib, g, tag, as_u16):
    div, far = float(g[f"div_{tag}"]), float(g[f"far_{tag}"])
    mask = torch.from_numpy(g["mask"]).cuda()
    boxes = torch.from_numpy(g["boxes"].astype(np.int32)).cuda()
    if as_u16:
        d = torch.from_numpy(g["raw"]).cuda()
        val, cnt, eroded = cuda_lib.depth_values(d, mask, boxes, 0.1, far, depth_div=div)
    else:
        d = torch.from_numpy(g["raw"].astype(np.float32) / div).cuda()
        val, cnt, eroded = cuda_lib.depth_values(d, mask, boxes, 0.1, far)
    torch.cuda.synchronize()
    assert np.array_equal(cnt.cpu().numpy() >= 50, g[f"rel_{tag}"])
    np.testing.assert_allclose(val.cpu().numpy(), g[f"val_{tag}"], rtol=RTOL, atol=0)
    if tag == "pose":
        assert np.array_equal(eroded.cpu().numpy() > 0, g["eroded_pose"])
    # pixel counts against the oracle's erosion
    dm = g["raw"].astype(np.float32) / div
    seg = od.shrink_mask(np.logical_and(g["mask"] > 128, np.logical_and(dm > 0.1, dm < far)), 10)
    want_cnt = [int(seg[b[1]:b[3], b[0]:b[2]].sum()) for b in g["boxes"]]
    assert cnt.cpu().numpy().tolist() == want_cnt


def test_get_depth_value_mirror_and_erosion_sizes(cuda_lib, g):
    from flope_b200 import image_manipulation as im
    from flope_b200 import mvg
    depth_m = g["raw"].astype(np.float32) / 10000.0
    keep = depth_m.copy()
    val, rel, vis = im.get_depth_value(g["boxes"], depth_m, g["mask"], near_plane=0.1, far_plane=2.5)
    assert vis is None and val.dtype == np.float64 and rel.dtype == bool
    assert np.array_equal(depth_m, keep)                         # unlike the reference, the argument is left alone
    assert np.array_equal(rel, g["rel_pose"])
    np.testing.assert_allclose(val, g["val_pose"], rtol=RTOL)
    xyz = mvg.get_points3d(od.box_centres(g["boxes"]), val, g["K"])
    np.testing.assert_allclose(xyz, g["xyz_pose"], rtol=RTOL, atol=1e-12)
    rng = np.random.default_rng(4)
    for k in (1, 2, 3, 5, 10, 11, 31):
        m = rng.random((97, 131)) > 0.05
        assert np.array_equal(im.shrink_mask(m, k), od.shrink_mask(m, k)), k
    with pytest.raises(cuda_lib.FlopeError):
        im.get_depth_value(g["boxes"], depth_m, g["mask"], vis=True)
    with pytest.raises(cuda_lib.FlopeError):
        im.shrink_mask(np.ones((8, 8), bool), 33)


def test_empty_and_ragged_boxes(cuda_lib, g):
    d = torch.from_numpy(g["raw"]).cuda()
    mask = torch.from_numpy(g["mask"]).cuda()
    val, cnt, _ = cuda_lib.depth_values(d, mask, torch.zeros((0, 4), dtype=torch.int32, device="cuda"), 0.1, 2.5, depth_div=10000.0)
    assert val.shape == (0,) and cnt.shape == (0,)
    # degenerate / partly out-of-frame boxes are clipped to the frame like numpy slicing clips the upper bounds
    boxes = torch.tensor([[100, 80, 100, 200], [500, 300, 900, 700], [630, 350, 640, 360]], dtype=torch.int32, device="cuda")
    val, cnt, eroded = cuda_lib.depth_values(d, mask, boxes, 0.1, 2.5, depth_div=10000.0)
    e = eroded.cpu().numpy() > 0
    assert cnt.cpu().numpy().tolist() == [0, int(e[300:360, 500:640].sum()), int(e[350:360, 630:640].sum())]
    assert float(val[0]) == 0.0


@pytest.mark.parametrize("cls_name,scale", [("PosePredictor", 10000.0), ("FastPosePredictor", 1000.0)])
def test_predictors_with_intrinsics_fill_translation(cuda_lib, g, cls_name, scale):
    """Full drop-in call: detector boxes + uint16 depth + K -> (N,4,4) with rotations AND translations, unreliable boxes
    dropped, against the oracle pipeline (fast_pose_predictor.py:66-156 / pose_predictor.py:83-186)."""
    from flope_b200 import predictor as P, synth
    from flope_b200.posenet import PoseResNet
    from oracle import pipeline as opipe, posenet as onet, resize as ores, rotation as orot
    net = onet.build(synth.WEIGHT_SEED)
    m = PoseResNet(device="cuda:0", max_batch=16, crop_hw=224)
    m.load_state_dict(net.state_dict())
    rng = np.random.default_rng(9)
    frame = rng.integers(0, 256, g["mask"].shape + (3,), dtype=np.uint8)
    raw = g["raw"] if scale == 10000.0 else (g["raw"] // 10).astype(np.uint16)
    det = g["boxes"][[0, 1, 2, 3, 5, 7]].astype(np.int16)        # box 5: fewer than 50 valid pixels -> dropped
    cls = getattr(P, cls_name)
    pred = cls("cuda:0", detector=lambda rgb: (det, g["mask"]), posenet=m, crop_hw=224, interp=ores.BILINEAR, K=g["K"])
    Rt = pred.get_flower_poses(frame, raw)
    from oracle import boxes as obox
    det_in = obox.filter_very_large_bb(det) if cls_name == "PosePredictor" else det      # pose_predictor.py:83
    want = opipe.run(net, frame, g["mask"], det_in, size=224, interp=ores.BILINEAR, depth=raw, K=g["K"], depth_scale=scale,
                     far_plane=2.5)
    assert want["reliable"].sum() < len(det_in)
    assert Rt.shape == want["Rt"].shape and Rt.dtype == np.float64
    np.testing.assert_allclose(Rt[:, :3, 3], want["Rt"][:, :3, 3], rtol=RTOL, atol=1e-12)
    assert np.all(Rt[:, :3, 3][:, 2] > 0.05)
    assert orot.geodesic_deg(Rt[:, :3, :3], want["Rt"][:, :3, :3]).mean() <= 0.5
    # all boxes unreliable -> None, like the reference (pose_predictor.py:129-130)
    none = cls("cuda:0", detector=lambda rgb: (det, np.zeros_like(g["mask"])), posenet=m, crop_hw=224, interp=ores.BILINEAR, K=g["K"])
    assert none.get_flower_poses(frame, raw) is None


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_yolo_mask_postprocessing_bit_exact(cuda_lib, golden_dir, tag):
    """flope_yolo_mask against the reference's own get_bbox_mask outputs: up- and down-scaling, odd sizes."""
    y = np.load(os.path.join(golden_dir, "yolo_post.npz"))
    H, W = y[f"mask_{tag}"].shape
    masks = torch.from_numpy(y[f"masks_{tag}"].astype(np.float32)).cuda()
    got = cuda_lib.yolo_mask(masks, H, W)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), y[f"mask_{tag}"])
    empty = cuda_lib.yolo_mask(torch.zeros((0,) + tuple(masks.shape[1:]), device="cuda"), H, W)
    assert int(empty.max()) == 0


def test_fast_predictor_with_raw_yolo_results(cuda_lib, golden_dir, g):
    """FastPosePredictor(yolo=...) does get_bbox_mask on the device and feeds the mask tensor straight into the depth
    and ROI kernels; the public get_bbox_mask keeps the reference's numpy contract."""
    import types
    from flope_b200 import predictor as P, synth
    from flope_b200.posenet import PoseResNet
    from oracle import detector_post as op, pipeline as opipe, posenet as onet, resize as ores, rotation as orot
    y = np.load(os.path.join(golden_dir, "yolo_post.npz"))
    masks = torch.from_numpy(y["masks_b"].astype(np.float32))
    boxes = torch.tensor([[60.4, 40.9, 180.2, 160.7], [240.0, 110.5, 420.9, 290.1], [475.3, 45.0, 565.8, 135.9]])
    res = [types.SimpleNamespace(masks=types.SimpleNamespace(data=masks.cuda()), boxes=types.SimpleNamespace(xyxy=boxes.cuda()))]
    net = onet.build(synth.WEIGHT_SEED)
    m = PoseResNet(device="cuda:0", max_batch=8, crop_hw=224)
    m.load_state_dict(net.state_dict())
    pred = P.FastPosePredictor("cuda:0", yolo=lambda img: res, posenet=m, crop_hw=224, interp=ores.BILINEAR, K=g["K"])
    frame = np.random.default_rng(2).integers(0, 256, (360, 640, 3), dtype=np.uint8)
    bbox, mask = pred.get_bbox_mask(frame)
    want_bbox, want_mask = op.bbox_mask_from_results(masks, boxes, 360, 640)
    assert bbox.dtype == np.int16 and np.array_equal(bbox, want_bbox) and np.array_equal(mask, want_mask)
    raw = (g["raw"] // 10).astype(np.uint16)
    Rt = pred.get_flower_poses(frame, raw)
    want = opipe.run(net, frame, want_mask, want_bbox, size=224, interp=ores.BILINEAR, depth=raw, K=g["K"], depth_scale=1000.0)
    if want is None:
        assert Rt is None
    else:
        assert Rt.shape == want["Rt"].shape
        np.testing.assert_allclose(Rt[:, :3, 3], want["Rt"][:, :3, 3], rtol=RTOL, atol=1e-12)
        assert orot.geodesic_deg(Rt[:, :3, :3], want["Rt"][:, :3, :3]).mean() <= 0.5


def test_frame_measurements_for_the_aggregator(cuda_lib, g, tmp_path):
    """aggregate.frame_measurements (detection file + depth + mask + camera pose -> world-frame measurements,
    scripts/flower_pose_aggregrator.py:189-232) against the same composition on the oracle's depth branch."""
    from scipy.spatial.transform import Rotation as sciR
    from flope_b200 import aggregate as agg
    from flope_b200.predictor import write_detection_txt
    rng = np.random.default_rng(3)
    boxes = g["boxes"][[0, 1, 2, 3, 5, 7]]
    rots = sciR.random(len(boxes), random_state=4).as_matrix()
    p = tmp_path / "f.txt"
    write_detection_txt(str(p), boxes, rots)
    depth_m = g["raw"].astype(np.float32) / 10000.0
    cam = np.eye(4); cam[:3, :3] = sciR.random(1, random_state=6).as_matrix()[0]; cam[:3, 3] = rng.normal(0, 1, 3)
    got = agg.frame_measurements(str(p), depth_m, g["mask"], cam, g["K"], None, 0.1, 2.5)
    bbox, uv, rotmat = agg.read_detection_txt(str(p))
    val, rel = od.get_depth_value(bbox, depth_m, g["mask"], near_plane=0.1, far_plane=2.5)
    xyz = od.get_points3d(uv[rel], np.asarray(val)[rel], g["K"])
    want_t = (cam[:3, :3] @ xyz.T).T + cam[:3, 3]
    want_R = cam[:3, :3] @ rotmat[rel].reshape(-1, 3, 3)
    assert got is not None and got[0].shape == want_t.shape and rel.sum() < len(boxes)
    np.testing.assert_allclose(got[0], want_t, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(sciR.from_quat(got[1]).as_matrix(), want_R, atol=1e-6)
    assert agg.frame_measurements(str(p), depth_m, np.zeros_like(g["mask"]), cam, g["K"]) is None


def test_flower_model_add_data_end_to_end(cuda_lib, g):
    """FlowerModel.add_data (flower_model.py:219-255) on the GPU predictor: camera-frame poses, world-frame poses and
    the tracker state after two frames of the same scene."""
    from scipy.spatial.transform import Rotation as sciR
    from flope_b200 import predictor as P, synth
    from flope_b200.flower_model import FlowerModel
    from flope_b200.posenet import PoseResNet
    from oracle import posenet as onet, resize as ores
    net = onet.build(synth.WEIGHT_SEED)
    m = PoseResNet(device="cuda:0", max_batch=8, crop_hw=224)
    m.load_state_dict(net.state_dict())
    det = g["boxes"][[0, 1, 2, 3, 7]].astype(np.int16)
    pp = P.PosePredictor("cuda:0", detector=lambda rgb: (det, g["mask"]), posenet=m, crop_hw=224, interp=ores.BILINEAR, K=g["K"])
    fm = FlowerModel(dist_th=50, pose_predictor=pp)
    frame = np.random.default_rng(12).integers(0, 256, (360, 640, 3), dtype=np.uint8)
    cam = np.concatenate([[0.1, -0.2, 0.3], sciR.from_euler("xyz", [10, 20, 30], degrees=True).as_quat()])
    cam_pose, world_pose = fm.add_data(frame, g["raw"], cam, ignore=True)
    assert cam_pose.dtype == np.float64 and world_pose.dtype == np.float32 and cam_pose.shape == world_pose.shape
    Rc = sciR.from_quat(cam[3:]).as_matrix()
    np.testing.assert_allclose(world_pose[:, :3, 3], (Rc @ cam_pose[:, :3, 3].T).T + cam[:3], atol=1e-6)
    n = cam_pose.shape[0]
    assert fm.get_state().shape == (n, 7) and np.array_equal(fm.scores, np.ones(n))
    fm.add_data(frame, g["raw"], cam, ignore=True)                 # same scene again: every flower re-identified
    assert fm.get_state().shape == (n, 7) and np.array_equal(fm.scores, 2 * np.ones(n))
    np.testing.assert_allclose(fm.get_filtered_state()[:, :3], fm.get_state()[:, :3], atol=1e-9)
    empty = FlowerModel(pose_predictor=P.PosePredictor("cuda:0", detector=lambda rgb: (np.zeros((0, 4), np.int16), g["mask"]),
                                                       posenet=m, crop_hw=224, K=g["K"]))
    assert empty.add_data(frame, g["raw"], cam) == (None, None)
