"""Seeded synthetic inputs for the pose path (SURVEY.md section 8d fixes these seeds).

Used by bench.py, the tests and __graft_entry__.smoke(); pure numpy / torch-CPU.
"""
import numpy as np
import torch

WEIGHT_SEED = 0
CROP_SEED = 100
FRAME_SEED = 7


def uniform_crops(n, size=224, seed=CROP_SEED):
    """(n,3,size,size) float32 U[0,1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((n, 3, size, size), generator=g, dtype=torch.float32)


def blob_crops(n, size=224, seed=CROP_SEED + 1):
    """Structured crops: 3-6 random Gaussian blobs per channel on black, clipped to [0,1]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    out = np.zeros((n, 3, size, size), np.float32)
    for i in range(n):
        for _ in range(int(rng.integers(3, 7))):
            cx, cy = rng.uniform(0, size, 2)
            sig = rng.uniform(size / 20, size / 4)
            amp = rng.uniform(0.2, 1.0, 3).astype(np.float32)
            g = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sig * sig))
            out[i] += amp[:, None, None] * g[None]
    return torch.from_numpy(np.clip(out, 0.0, 1.0))


def mixed_crops(n, size=224, seed=CROP_SEED):
    """Half uniform noise, half blobs (the parity set of SURVEY.md H3)."""
    a = uniform_crops((n + 1) // 2, size, seed)
    b = blob_crops(n // 2, size, seed + 1)
    return torch.cat([a, b])[:n]


def frames_and_boxes(n_frames=64, boxes_per_frame=32, H=1080, W=1920, seed=FRAME_SEED, with_mask=True,
                     smooth=False):
    """uint8 frames (n,H,W,3), masks (n,H,W) in {0,255} (or None) and detector boxes.

    Boxes: w,h ~ U{48..320}, xmin ~ U{0..W-w}, ymin ~ U{0..H-h}, int32 xyxy, resampled until
    exactly ``boxes_per_frame`` of them survive squarify_bb + bb_in_frame, so the frame config
    stays at n_frames*boxes_per_frame crops.  Returns (frames, masks, det_boxes (n,k,4) int32).
    """
    from .mvg import squarify_bb, bb_in_frame
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (n_frames, H, W, 3), dtype=np.uint8)
    if smooth:
        f = frames.astype(np.float32)
        for ax in (1, 2):
            f = (f + np.roll(f, 1, ax) + np.roll(f, -1, ax) + np.roll(f, 2, ax)) / 4
        frames = np.clip(f, 0, 255).astype(np.uint8)
    masks = None
    if with_mask:
        masks = np.zeros((n_frames, H, W), np.uint8)
    det = np.zeros((n_frames, boxes_per_frame, 4), np.int32)
    for f in range(n_frames):
        k = 0
        while k < boxes_per_frame:
            w, h = rng.integers(48, 321, 2)
            x0 = int(rng.integers(0, W - w + 1))
            y0 = int(rng.integers(0, H - h + 1))
            bb = [x0, y0, x0 + int(w), y0 + int(h)]
            if not bb_in_frame(squarify_bb(bb), (H, W, 3)):
                continue
            det[f, k] = bb
            if with_mask:  # an ellipse inscribed in the detector box, like a flower segment
                yy, xx = np.ogrid[y0:y0 + h, x0:x0 + w]
                e = ((xx - (x0 + w / 2)) / (w / 2)) ** 2 + ((yy - (y0 + h / 2)) / (h / 2)) ** 2 <= 1.0
                masks[f, y0:y0 + h, x0:x0 + w][e] = 255
            k += 1
    return frames, masks, det


def random_state_dict(seed=WEIGHT_SEED):
    """Random-init PoseResNet weights with the reference's key schema (sunflower/models/posenet.py:5-34).

    Builds the same torch modules in the same order as the reference constructor (torchvision
    ResNet-18 trunk with weights=None, then base.fc = Linear(512,2048)+ReLU, then fc_rot =
    Linear(2048,9)), so a given seed yields the same tensors as the reference class built under
    that seed.  Used for init only - nothing here computes a forward pass.
    """
    import torch.nn as nn
    import torchvision.models as tvm
    torch.manual_seed(seed)
    base = tvm.resnet18(weights=None)
    base.avgpool = nn.AdaptiveAvgPool2d(1)
    base.fc = nn.Sequential(nn.Linear(512, 2048), nn.ReLU())
    fc_rot = nn.Linear(2048, 9)
    sd = {"base." + k: v.detach().clone() for k, v in base.state_dict().items()}
    sd.update({"fc_rot." + k: v.detach().clone() for k, v in fc_rot.state_dict().items()})
    return sd
