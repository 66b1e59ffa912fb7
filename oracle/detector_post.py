"""Oracle: YOLO-seg post-processing (test infrastructure only).

Restates the body of FastPosePredictor.get_bbox_mask after the detector call
(sunflower/predictor/fast_pose_predictor.py:48-57; duplicate at scripts/generate_metrics_utils.py:114-127) on torch-CPU
and the real cv2.  Pinned: tests/golden/yolo_post.npz holds the outputs of the reference's own method (run on a
FastPosePredictor instance whose ``yolo`` is a stub returning seeded masks / boxes; ultralytics itself is absent).
"""
import numpy as np
import torch


def bbox_mask_from_results(masks, boxes_xyxy, H, W):
    """masks (n,h,w) float tensor, boxes (n,4) float tensor -> (bbox (n,4) int16, mask (H,W) uint8)."""
    import cv2
    mask = torch.sum(masks, axis=0)
    mask = torch.clip(mask, 0, 1) * 255
    mask = mask.cpu().numpy().astype(np.uint8)
    mask = cv2.resize(mask, (W, H))
    bbox = boxes_xyxy.cpu().numpy().astype(np.int16)
    return bbox, mask
