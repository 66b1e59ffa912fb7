"""GPU drop-in test: FastPosePredictor / PosePredictor with an injected detector vs the oracle pipeline."""
import numpy as np
import pytest
import torch

from flope_b200 import synth
from oracle import pipeline as opipe
from oracle import posenet as onet
from oracle import resize as ores
from oracle import rotation as orot

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net():
    return onet.build(synth.WEIGHT_SEED)


def _scene(seed=2, n_boxes=6):
    frames, masks, det = synth.frames_and_boxes(1, n_boxes, H=360, W=640, seed=seed, smooth=True)
    det = det[0].copy()
    det[0] = [600, 10, 640, 200]       # squarified box leaves the frame -> dropped by bb_in_frame
    return frames[0], masks[0], det


@pytest.mark.parametrize("crop_hw,interp", [(512, ores.LANCZOS4), (224, ores.BILINEAR)])
def test_fast_pose_predictor_matches_oracle(net, crop_hw, interp):
    from flope_b200.posenet import PoseResNet
    from flope_b200.predictor import FastPosePredictor
    frame, mask, det = _scene()
    m = PoseResNet(device="cuda:0", max_batch=8, crop_hw=crop_hw)
    m.load_state_dict(net.state_dict())
    pred = FastPosePredictor("cuda:0", detector=lambda rgb: (det.astype(np.int16), mask), posenet=m,
                             crop_hw=crop_hw, interp=interp)
    Rt = pred.get_flower_poses(frame, np.zeros(frame.shape[:2], np.uint16))
    want = opipe.run(net, frame, mask, det, size=crop_hw, interp=interp)
    assert Rt.dtype == np.float64 and Rt.shape == want["Rt"].shape == (5, 4, 4)
    assert np.array_equal(Rt[:, 3], np.tile([0, 0, 0, 1.0], (5, 1)))
    g = orot.geodesic_deg(Rt[:, :3, :3], want["Rt"][:, :3, :3])
    print("predictor geodesic mean %.4f max %.4f" % (g.mean(), g.max()))
    assert g.mean() <= 0.5
    # raw (not yaw-nullified) rotations, the scripts/test_posenet.py output
    raw = pred.poses_from_boxes(frame, mask, want["sq_boxes"].astype(np.int32), nullify_yaw=False)
    assert raw.dtype == np.float32
    assert orot.geodesic_deg(raw, want["rot"]).mean() <= 0.5


def test_predictor_returns_none_like_reference(net):
    from flope_b200.posenet import PoseResNet
    from flope_b200.predictor import FastPosePredictor, PosePredictor
    frame, mask, det = _scene()
    m = PoseResNet(device="cuda:0", max_batch=8, crop_hw=224)
    m.load_state_dict(net.state_dict())
    none_boxes = FastPosePredictor("cuda:0", detector=lambda rgb: (np.zeros((0, 4), np.int16), mask), posenet=m,
                                   crop_hw=224)
    assert none_boxes.get_flower_poses(frame, None) is None
    all_out = FastPosePredictor("cuda:0", detector=lambda rgb: (np.array([[600, 10, 640, 200]], np.int16), mask),
                                posenet=m, crop_hw=224)
    assert all_out.get_flower_poses(frame, None) is None
    # PosePredictor applies filter_very_large_bb first (pose_predictor.py:83)
    small = np.array([[20 + 70 * i, 30 + 40 * i, 70 + 70 * i, 80 + 40 * i] for i in range(5)], np.int64)
    big = np.concatenate([small, [[0, 0, 350, 350]]])            # 122500 px^2 > 5 x median (2500 px^2) -> dropped
    pp = PosePredictor("cuda:0", detector=lambda rgb: (big, mask), posenet=m, crop_hw=224, interp=ores.BILINEAR)
    Rt = pp.get_flower_poses(frame, np.zeros(frame.shape[:2], np.uint16))
    assert Rt.shape == (5, 4, 4)
    fp = FastPosePredictor("cuda:0", detector=lambda rgb: (big, mask), posenet=m, crop_hw=224, interp=ores.BILINEAR)
    assert fp.get_flower_poses(frame, np.zeros(frame.shape[:2], np.uint16)).shape == (6, 4, 4)


def test_detection_txt_format(tmp_path, net):
    from flope_b200.predictor import write_detection_txt
    det = np.array([[10, 20, 110, 121], [5, 6, 50, 60]], np.int16)
    rot = np.stack([np.eye(3, dtype=np.float32)] * 2)
    p = tmp_path / "det.txt"
    write_detection_txt(str(p), det, rot)
    rows = np.loadtxt(str(p))
    assert rows.shape == (2, 15)
    assert np.allclose(rows[0, :6], [10, 20, 110, 121, 60, 70.5])
    assert open(p).read().split()[0] == "10.0000000"


def test_bulk_inference_over_an_image_folder(tmp_path, net):
    """flope_b200.bulk_infer: images + detector boxes (+ masks) on disk -> detection/*.txt, the PoseNet half of
    scripts/test_posenet.py:62-161, against the oracle pipeline (raw Procrustes rotations, '%.7f' rows)."""
    import cv2
    from flope_b200 import bulk_infer
    from flope_b200.aggregate import read_detection_txt
    frames, masks, det = synth.frames_and_boxes(3, 5, H=360, W=640, seed=31, smooth=True)
    det[2, 0] = [600, 10, 640, 200]                      # squarified box leaves the frame: dropped
    for d in ("rgb", "boxes", "masks"):
        (tmp_path / d).mkdir()
    for i in range(3):
        cv2.imwrite(str(tmp_path / "rgb" / f"{i:06d}.png"), frames[i])
        cv2.imwrite(str(tmp_path / "masks" / f"{i:06d}.png"), masks[i])
        np.savetxt(str(tmp_path / "boxes" / f"{i:06d}.txt"), det[i].astype(np.float64) if i else np.zeros((0, 4)))
    n_img, n_fl = bulk_infer.run(str(tmp_path / "rgb"), str(tmp_path / "boxes"), str(tmp_path / "det"), str(tmp_path / "masks"),
                                 state_dict=net.state_dict(), crop=224, interp="linear", frames_per_batch=2)
    assert n_img == 3
    assert (tmp_path / "det" / "000000.txt").read_text() == ""                  # no boxes: empty file, like the reference
    for i in (1, 2):
        want = opipe.run(net, frames[i], masks[i], det[i], size=224, interp=ores.BILINEAR, nullify_yaw=False)
        bbox, uv, rot = read_detection_txt(str(tmp_path / "det" / f"{i:06d}.txt"))
        assert np.array_equal(bbox, det[i][want["keep"]].astype(np.int16))
        assert orot.geodesic_deg(rot.reshape(-1, 3, 3), want["rot"]).mean() <= 0.5
    assert n_fl == 5 + 4


@pytest.mark.parametrize("crop_hw,interp", [(224, ores.BILINEAR), (224, ores.LANCZOS4)])
def test_packed_box_upload_equals_full_frame_upload(net, crop_hw, interp):
    """A few flowers in a 1080p frame: the predictor uploads the packed box regions (flope_pack_boxes) instead of the
    frame.  Same kernels on the same pixels: the poses must be bit-identical to the full-frame path, with and without mask."""
    from flope_b200.posenet import PoseResNet
    from flope_b200.predictor import FastPosePredictor
    frames, masks, det = synth.frames_and_boxes(1, 6, seed=41, with_mask=True)
    m = PoseResNet(device="cuda:0", max_batch=8, crop_hw=crop_hw)
    m.load_state_dict(net.state_dict())
    for mask in (masks[0], None):
        pred = FastPosePredictor("cuda:0", detector=lambda rgb: (det[0].astype(np.int16), mask), posenet=m, crop_hw=crop_hw,
                                 interp=interp)
        assert pred._packed_upload(frames[0], mask, np.array([[0, 0, 100, 100]], np.int32)) is not None
        a = pred.get_flower_poses(frames[0], None)
        pred.PACK_BOXES = False
        b = pred.get_flower_poses(frames[0], None)
        assert a.shape == (6, 4, 4) and np.array_equal(a, b)
    many = np.tile(np.array([[0, 0, 600, 600]], np.int32), (8, 1))          # boxes larger than half the frame in total: full-frame path
    pred.PACK_BOXES = True
    assert pred._packed_upload(frames[0], None, many) is None
