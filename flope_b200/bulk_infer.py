"""Offline bulk inference over an image folder: the PoseNet half of scripts/test_posenet.py:62-161.

    python -m flope_b200.bulk_infer --images rgb/ --boxes boxes/ [--masks masks/] --out detection/ \
        --weights posenet_e183.pth [--device cuda:0] [--crop 512] [--interp lanczos4|linear] [--frames-per-batch 16]

The reference script runs GroundingDINO + SAM on every image, then PoseNet, and writes one ``detection/<name>.txt`` per
image (15 columns, '%.7f': xmin ymin xmax ymax u v r00..r22 - the RAW Procrustes rotation, no yaw nullification).  The
detectors are outside this repository's scope, so their outputs are inputs here: ``boxes/<name>.txt``
holds the detector's xyxy rows (what the script keeps in ``bb_dino``), ``masks/<name>.png`` the segmentation mask the
script writes (optional; without it no background removal).  Everything after that is the GPU path: squarify + in-frame
filter, ROI crop of the uint8 frame (cv2-exact Lanczos4 -> 512 by default, like the script), PoseNet, Procrustes.
Images of equal size are batched: several frames go through ONE flope_infer_frames call.
An image without surviving boxes gets an empty file, like the reference (test_posenet.py:117-122).
"""
import argparse
import os
import sys

import numpy as np


def _stem(name):
    return os.path.splitext(os.path.basename(name))[0]


def run(images, boxes_dir, out_dir, masks_dir=None, weights=None, state_dict=None, device="cuda:0", crop=512,
        interp="lanczos4", frames_per_batch=16, log=None):
    """Returns the number of (images, flowers) written.  ``state_dict`` (a mapping) takes precedence over ``weights`` (a path)."""
    import cv2
    import torch
    from . import _lib
    from .posenet import PoseResNet
    from .predictor import write_detection_txt

    os.makedirs(out_dir, exist_ok=True)
    names = sorted(f for f in os.listdir(images) if f.lower().endswith((".png", ".jpg", ".jpeg", ".bmp")))
    dev = torch.device(device)
    model = PoseResNet(device=str(dev), max_batch=64, crop_hw=crop)
    if state_dict is None:
        if weights is None:
            raise ValueError("bulk_infer needs PoseNet weights (weights= path or state_dict=)")
        state_dict = torch.load(weights, weights_only=True, map_location="cpu")
    model.load_state_dict(state_dict)
    eng = model.engine
    mode = _lib.INTERP_LANCZOS4 if interp == "lanczos4" else _lib.INTERP_LINEAR
    n_img = n_flowers = 0

    def flush(batch):
        nonlocal n_img, n_flowers
        if not batch:
            return
        frames = torch.from_numpy(np.stack([b[1] for b in batch])).to(dev)
        masks = torch.from_numpy(np.stack([b[2] for b in batch])).to(dev) if batch[0][2] is not None else None
        rows = [np.concatenate([np.full((len(b[3]), 1), i, np.int32), b[3]], 1) for i, b in enumerate(batch) if len(b[3])]
        R = None
        if rows:
            with torch.cuda.device(dev):
                _, R, _ = eng.infer_frames(frames, masks, np.concatenate(rows), mode, want_R=True, want_yaw=False)
            R = R.cpu().numpy()
        k = 0
        for name, _, _, sq, kept in batch:
            path = os.path.join(out_dir, _stem(name) + ".txt")
            if len(sq) == 0:
                np.savetxt(path, np.array([]), fmt='%.7f')                      # test_posenet.py:120
            else:
                write_detection_txt(path, kept, R[k:k + len(sq)])
                k += len(sq)
                n_flowers += len(sq)
            n_img += 1
        if log:
            log(f"{n_img}/{len(names)} images, {n_flowers} flowers")

    batch, shape = [], None
    for name in names:
        img = cv2.imread(os.path.join(images, name), cv2.IMREAD_COLOR)          # BGR, like img_cv in the script
        if img is None:
            continue
        bpath = os.path.join(boxes_dir, _stem(name) + ".txt")
        det = np.loadtxt(bpath).reshape(-1, 4) if os.path.exists(bpath) and os.path.getsize(bpath) else np.zeros((0, 4))
        mask = None
        if masks_dir is not None:
            mask = cv2.imread(os.path.join(masks_dir, _stem(name) + ".png"), cv2.IMREAD_GRAYSCALE)
            if mask is None or mask.shape != img.shape[:2]:
                raise ValueError(f"mask for {name} missing or of the wrong size")
        det_i = np.ascontiguousarray(det.astype(np.int64).astype(np.int32))      # int() truncation of the script's boxes
        sq, keep = _lib.squarify_filter(det_i, img.shape[0], img.shape[1])
        if (img.shape, mask is None) != shape or len(batch) >= frames_per_batch:
            flush(batch)
            batch, shape = [], (img.shape, mask is None)
        batch.append((name, img, mask, sq, det[keep]))
    flush(batch)
    return n_img, n_flowers


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--images", required=True)
    ap.add_argument("--boxes", required=True)
    ap.add_argument("--masks")
    ap.add_argument("--out", required=True)
    ap.add_argument("--weights", required=True)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--crop", type=int, default=512)
    ap.add_argument("--interp", choices=["lanczos4", "linear"], default="lanczos4")
    ap.add_argument("--frames-per-batch", type=int, default=16)
    a = ap.parse_args(argv)
    n_img, n_fl = run(a.images, a.boxes, a.out, a.masks, a.weights, None, a.device, a.crop, a.interp, a.frames_per_batch,
                      log=lambda s: print(s, file=sys.stderr))
    print(f"{n_img} images, {n_fl} flowers -> {a.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
