"""Drop-in predictors mirroring sunflower/predictor/{fast_pose_predictor,pose_predictor}.py.

Same class names, constructor keywords and ``get_flower_poses(rgb, depth) -> (N,4,4) float64
| None`` contract.  What changes is everything between the detector and the returned poses:
box squarify/filter (C ABI host function), the crop batch (fused ROI kernel on the uint8
frame - the frame crosses PCIe once instead of 3 MiB of float32 per crop), PoseNet (tcgen05
backbone), Procrustes and yaw nullification (fused head kernel).

Out of scope and therefore injected: the detector/segmenter
(ultralytics YOLO-seg, GroundingDINO + SAM - third-party models whose weights are not
available offline): ``detector(rgb) -> (boxes (N,4) int, mask (H,W) uint8 ndarray or CUDA tensor)`` is a
constructor keyword.  ``FastPosePredictor(yolo=callable)`` instead takes the raw YOLO-seg model call
(``yolo(image)[0].masks.data`` / ``.boxes.xyxy`` like ultralytics) and does the reference's
``get_bbox_mask`` post-processing on the GPU (``flope_yolo_mask``): instance masks
are merged and resized to the frame with cv2-exact bilinear arithmetic without leaving the device.

The depth / translation branch (pose_predictor.py:118-135, fast_pose_predictor.py:90-105:
``get_depth_value`` -> drop unreliable boxes -> ``get_points3d``) runs on the GPU
(``flope_depth_values``) whenever camera intrinsics are known
(``intrin_path`` or ``K=``); the raw uint16 depth frame crosses PCIe as is and is scaled on the device.
``depth_fn(good_boxes, depth_m, mask) -> (xyz (N,3), reliable (N,) bool)`` overrides it; without
intrinsics and without ``depth_fn`` every box is kept and translations are zero.
"""
import numpy as np
import torch

from . import _lib
from .image_manipulation import RELIABLE_MIN_PIXELS
from .mvg import filter_very_large_bb, get_points3d
from .posenet import PoseResNet


def _load_intrinsics(intrin_path):
    """sunflower/utils/io.py:92-98."""
    if intrin_path is None:
        return None, None, None
    import yaml
    with open(intrin_path, "r") as f:
        d = yaml.safe_load(f)
    K = np.array([[d['fx'], 0, d['cx']], [0, d['fy'], d['cy']], [0, 0, 1]])
    return K, d['h'], d['w']


class _PredictorBase:
    CROP = 512                      # the reference's crop side (pose_predictor.py:145)
    INTERP = _lib.INTERP_LANCZOS4   # cv2.INTER_LANCZOS4 (pose_predictor.py:145-146)
    DEPTH_SCALE = 1000.0
    NEAR_PLANE, FAR_PLANE = 0.1, 2.5

    def _init_common(self, device, posenet_path, intrin_path, debug, detector, depth_fn, posenet, max_batch,
                     crop_hw, interp, K=None):
        self.device = torch.device(device if device != 'cuda' else 'cuda:0')
        self.debug = debug
        self.detector = detector
        self.depth_fn = depth_fn
        self.crop_hw = crop_hw or self.CROP
        self.interp = self.INTERP if interp is None else interp
        if posenet is not None:
            self.posenet = posenet
        else:
            self.posenet = PoseResNet(device=str(self.device), max_batch=max_batch, crop_hw=self.crop_hw)
            if posenet_path is not None:
                self.posenet.load_state_dict(torch.load(posenet_path, weights_only=True, map_location="cpu"))
        self.K, self.height, self.width = _load_intrinsics(intrin_path)
        if K is not None:
            self.K = np.asarray(K, dtype=np.float64)

    def _filter_boxes(self, boxes):
        return boxes

    def get_flower_poses(self, rgb, depth):
        if self.detector is None:
            raise _lib.FlopeError("no detector injected: pass detector=callable(rgb)->(boxes, mask)")
        boxes, mask = self.detector(rgb)
        if torch.is_tensor(mask) and self.depth_fn is not None:
            mask = mask.cpu().numpy()               # an injected depth function gets the reference's numpy mask
        boxes = np.asarray(boxes)
        if boxes.shape[0] == 0:
            return None
        boxes = self._filter_boxes(boxes)
        H, W = rgb.shape[0], rgb.shape[1]
        # squarify_bb + bb_in_frame over all boxes (fast_pose_predictor.py:69-82)
        boxes_i32 = np.ascontiguousarray(boxes.reshape(-1, 4), dtype=np.int32)
        sq_bb, keep = _lib.squarify_filter(boxes_i32, H, W)
        good_bb = boxes_i32[keep].astype(np.int16)
        if good_bb.shape[0] == 0:
            return None
        xyz = None
        mask_dev = None
        if self.depth_fn is not None:
            xyz, reliable = self.depth_fn(good_bb, depth.astype(np.float32) / self.DEPTH_SCALE, mask)
            sq_bb = sq_bb[reliable]
            xyz = np.asarray(xyz)[reliable] if len(xyz) == len(reliable) else np.asarray(xyz)
            if sq_bb.shape[0] == 0:
                return None
        elif self.K is not None and depth is not None:
            # pose_predictor.py:118-135 on the device: depth values, reliability filter, lift the box centres to 3-D
            u = (good_bb[:, 2].astype(np.float64) + good_bb[:, 0]) / 2          # pose_predictor.py:98-99
            v = (good_bb[:, 3].astype(np.float64) + good_bb[:, 1]) / 2
            depth = np.asarray(depth)
            with torch.cuda.device(self.device):
                mask_dev = mask.to(self.device) if torch.is_tensor(mask) else torch.from_numpy(np.ascontiguousarray(mask)).to(self.device)
                if depth.dtype == np.uint16:
                    d_dev, div = torch.from_numpy(np.ascontiguousarray(depth)).to(self.device), self.DEPTH_SCALE
                else:
                    d_dev, div = torch.from_numpy(depth.astype(np.float32) / np.float32(self.DEPTH_SCALE)).to(self.device), None
                b_dev = torch.from_numpy(good_bb.astype(np.int32)).to(self.device)
                val, cnt, _ = _lib.depth_values(d_dev, mask_dev, b_dev, self.NEAR_PLANE, self.FAR_PLANE, depth_div=div)
                reliable = cnt.cpu().numpy() >= RELIABLE_MIN_PIXELS
                depth_val = val.cpu().numpy()[reliable]
            sq_bb = sq_bb[reliable]
            if sq_bb.shape[0] == 0:
                return None
            xyz = get_points3d(np.stack([u, v], 1)[reliable], depth_val, self.K)
        if mask_dev is None and torch.is_tensor(mask):
            mask_dev = mask.to(self.device)
        rot = self.poses_from_boxes(rgb, mask if mask_dev is None else mask_dev[None], sq_bb)
        Rt = np.repeat(np.eye(4)[None], rot.shape[0], axis=0)      # fast_pose_predictor.py:142-144
        Rt[:, :3, :3] = rot
        if xyz is not None:
            Rt[:, :3, 3] = xyz
        return Rt

    PACK_BOXES = True               # upload the box regions instead of the frame when they are a fraction of it

    def _packed_upload(self, rgb, mask, sq_bb):
        """A handful of flowers per frame (scripts/live_pose.py): the boxes cover a fraction of the frame, and a
        pageable 6 MB frame upload costs more than the whole device pipeline.  Pack the box regions into fixed-size
        slots of one pinned arena on the host (flope_pack_boxes: row memcpys), upload the arena with ONE copy and run
        the same ROI kernel on it - every slot is a small "frame" with its box at the origin, so the crops are bit-identical.
        Returns (frames (n,sh,sw,3), masks (n,sh,sw) or None, boxes5 (n,5)) on the device, or None to take the full-frame path."""
        if not self.PACK_BOXES or torch.is_tensor(mask) or not isinstance(rgb, np.ndarray) or rgb.dtype != np.uint8 or rgb.ndim != 3:
            return None
        n = sq_bb.shape[0]
        sides = np.maximum(sq_bb[:, 2] - sq_bb[:, 0], sq_bb[:, 3] - sq_bb[:, 1])
        sh = int(sides.max())
        sw = (sh + 15) & ~15
        H, W = rgb.shape[:2]
        if 2 * n * sh * sw > H * W:                     # many or large boxes: the frame is the smaller upload
            return None
        has_mask = mask is not None
        if has_mask:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
        fb, mb = n * sh * sw * 3, (n * sh * sw if has_mask else 0)
        total = fb + mb + n * 20
        arena = getattr(self, "_arena", None)
        if arena is None or arena[0].numel() < total:
            cap = max(total, 1 << 20)
            arena = (torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.empty(cap, dtype=torch.uint8, device=self.device))
            self._arena = arena
        host, dev = arena
        hn = host.numpy()
        bx = np.ascontiguousarray(sq_bb, dtype=np.int32)
        _lib.pack_boxes(np.ascontiguousarray(rgb), bx, sh, sw, hn[:fb])
        if has_mask:
            _lib.pack_boxes(mask, bx, sh, sw, hn[fb:fb + mb])
        b5 = np.zeros((n, 5), np.int32)
        b5[:, 0] = np.arange(n)
        b5[:, 3] = bx[:, 2] - bx[:, 0]
        b5[:, 4] = bx[:, 3] - bx[:, 1]
        hn[fb + mb:total] = np.frombuffer(b5.tobytes(), np.uint8)
        with torch.cuda.device(self.device):
            dev[:total].copy_(host[:total], non_blocking=True)
        frames = dev[:fb].view(n, sh, sw, 3)
        masks = dev[fb:fb + mb].view(n, sh, sw) if has_mask else None
        return frames, masks, dev[fb + mb:total].view(torch.int32).view(n, 5)

    def poses_from_boxes(self, rgb, mask, sq_bb, nullify_yaw=True):
        """uint8 frame (H,W,3) + mask (H,W) or None + square in-frame boxes (N,4) -> rotations (N,3,3).

        float64 yaw-nullified (the predictor's output) or float32 raw Procrustes rotations
        (what scripts/test_posenet.py:144-161 writes) when nullify_yaw=False.
        """
        eng = self.posenet.engine
        sq_bb = np.asarray(sq_bb)
        if (sq_bb[:, 2] <= sq_bb[:, 0]).any() or (sq_bb[:, 3] <= sq_bb[:, 1]).any():
            # the reference hands such a slice to cv2.resize, which raises (pose_predictor.py:145)
            raise ValueError("empty box (xmax <= xmin or ymax <= ymin) after squarify")
        packed = self._packed_upload(rgb, mask, sq_bb)
        if packed is not None:
            frame, msk, b5 = packed
            with torch.cuda.device(self.device):
                _, R, Ry = eng.infer_frames(frame, msk, b5, self.interp, want_R=not nullify_yaw, want_yaw=nullify_yaw)
                return (Ry if nullify_yaw else R).cpu().numpy()
        with torch.cuda.device(self.device):
            frame = torch.from_numpy(np.ascontiguousarray(rgb, dtype=np.uint8)).to(self.device, non_blocking=True)[None]
            if mask is None:
                msk = None
            elif torch.is_tensor(mask):
                msk = mask.to(self.device)
                if msk.dtype == torch.bool:
                    msk = msk.to(torch.uint8) * 255
                elif msk.dtype != torch.uint8:
                    msk = msk.to(torch.uint8)
                msk = msk.reshape(1, *msk.shape[-2:]).contiguous()
            else:
                msk = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.uint8)).to(self.device)[None]
            b5 = np.zeros((sq_bb.shape[0], 5), np.int32)
            b5[:, 1:] = sq_bb
            _, R, Ry = eng.infer_frames(frame, msk, b5, self.interp, want_R=not nullify_yaw, want_yaw=nullify_yaw)
            return (Ry if nullify_yaw else R).cpu().numpy()


class FastPosePredictor(_PredictorBase):
    """sunflower/predictor/fast_pose_predictor.py:19-156 (YOLO-seg front end, depth in mm)."""
    DEPTH_SCALE = 1000.0            # fast_pose_predictor.py:90
    NEAR_PLANE, FAR_PLANE = 0.1, 2.5    # fast_pose_predictor.py:91-94

    def __init__(self, device: str, yolo_path: str = None, posenet_path: str = None, intrin_path: str = None,
                 debug: bool = False, *, detector=None, depth_fn=None, posenet=None, max_batch=64, crop_hw=None,
                 interp=None, K=None, yolo=None):
        self.yolo_path = yolo_path
        self.yolo = yolo
        if detector is None and yolo is not None:
            detector = self._bbox_mask_device
        self._init_common(device, posenet_path, intrin_path, debug, detector, depth_fn, posenet, max_batch, crop_hw,
                          interp, K)

    def _bbox_mask_device(self, image):
        """fast_pose_predictor.py:44-57 with the mask kept on the GPU: (bbox (n,4) int16 ndarray, mask (H,W) uint8 CUDA tensor)."""
        H, W, _ = image.shape
        results = self.yolo(image)
        masks = results[0].masks.data
        mask = _lib.yolo_mask(torch.as_tensor(masks).to(self.device), H, W)
        bbox = torch.as_tensor(results[0].boxes.xyxy).cpu().numpy().astype(np.int16)
        return bbox, mask

    def get_bbox_mask(self, image):
        """Same contract as the reference method: (bbox (n,4) int16, mask (H,W) uint8), both numpy."""
        bbox, mask = self._bbox_mask_device(image)
        return bbox, mask.cpu().numpy()


class PosePredictor(_PredictorBase):
    """sunflower/predictor/pose_predictor.py:40-186 (GroundingDINO + SAM front end, depth in 0.1 mm)."""
    DEPTH_SCALE = 10000.0           # pose_predictor.py:118
    NEAR_PLANE, FAR_PLANE = 0.1, 2.5    # pose_predictor.py:119-122

    def __init__(self, device: str, posenet_path: str = None, intrin_path: str = None, debug: bool = False, *,
                 detector=None, depth_fn=None, posenet=None, max_batch=64, crop_hw=None, interp=None, K=None):
        self._init_common(device, posenet_path, intrin_path, debug, detector, depth_fn, posenet, max_batch, crop_hw,
                          interp, K)

    def _filter_boxes(self, boxes):
        return filter_very_large_bb(boxes)      # pose_predictor.py:83


def write_detection_txt(path, det_boxes, rot):
    """scripts/test_posenet.py:150-161 - one row per flower: xmin ymin xmax ymax u v r00..r22, '%.7f'."""
    rows = []
    for bb, R in zip(np.asarray(det_boxes), np.asarray(rot)):
        xmin, ymin, xmax, ymax = bb
        rows.append(list(map(float, bb)) + [float((xmin + xmax) / 2), float((ymin + ymax) / 2)] + R.flatten().tolist())
    np.savetxt(path, np.array(rows), fmt='%.7f')
