"""Depth branch of the predictors, mirroring sunflower/utils/image_manipulation.py.

  get_depth_value   sunflower/utils/image_manipulation.py:39-96  - same arguments and return triple
  shrink_mask       sunflower/utils/image_manipulation.py:21-36

Both run on the GPU through the C ABI (``flope_depth_values``: validity mask, cv2-exact elliptical erosion, per-box
masked mean); there is no CPU path.  Differences from the reference as written: ``depth`` is not modified in place
(the reference scales its argument by 1000), the values come back as float64 always (the reference's dtype depends
on whether any box was empty), and the 50x50 visualisation crops (``vis=True``) are not produced.
"""
import numpy as np

from . import _lib

RELIABLE_MIN_PIXELS = 50            # image_manipulation.py:77


def _device(device):
    import torch
    if not torch.cuda.is_available():
        raise _lib.FlopeError("flope_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda:0" if device in (None, "cuda") else device)


def get_depth_value(bbox, depth, seg_mask, scale=None, near_plane=0.1, far_plane=3.0, vis=False, *, device=None,
                    depth_div=None):
    """bbox (N,4) int xmin,ymin,xmax,ymax; depth (H,W) float metres (or uint16 sensor units with ``depth_div`` -
    the predictors' ``depth.astype(np.float32) / 10000`` then happens on the device); seg_mask (H,W) uint8.
    -> (depth values in metres (N,) float64, reliable (N,) bool, None)"""
    import torch
    if vis:
        raise _lib.FlopeError("vis=True (50x50 depth visualisation crops) is outside the pose path and not implemented")
    dev = _device(device)
    if torch.is_tensor(depth):
        d = depth.to(dev)
    else:
        depth = np.asarray(depth)
        if depth.dtype == np.uint16 and depth_div is not None:
            d = torch.from_numpy(np.ascontiguousarray(depth)).to(dev)
        else:
            d = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32)).to(dev)
    if scale:
        if d.dtype != torch.float32:
            raise _lib.FlopeError("scale applies to float depth")
        d = d * np.float32(scale)                       # image_manipulation.py:64 (float32 multiply)
    m = seg_mask.to(dev) if torch.is_tensor(seg_mask) else torch.from_numpy(np.ascontiguousarray(seg_mask, dtype=np.uint8)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(np.asarray(bbox).reshape(-1, 4), dtype=np.int32)).to(dev)
    val, cnt, _ = _lib.depth_values(d, m, b, near_plane, far_plane, depth_div=depth_div)
    cnt = cnt.cpu().numpy()
    return val.cpu().numpy(), cnt >= RELIABLE_MIN_PIXELS, None


def shrink_mask(mask, kernel_size=3, *, device=None):
    """Boolean (H,W) mask -> eroded boolean mask (cv2.erode with MORPH_ELLIPSE, kernel_size <= 31)."""
    import torch
    dev = _device(device)
    m = torch.from_numpy(np.ascontiguousarray(np.asarray(mask).astype(np.uint8) * 255)).to(dev)
    d = torch.ones(m.shape, dtype=torch.float32, device=dev)
    _, _, eroded = _lib.depth_values(d, m, torch.zeros((0, 4), dtype=torch.int32, device=dev), 0.0, 2.0,
                                     erode_k=kernel_size)
    return eroded.cpu().numpy() > 0
