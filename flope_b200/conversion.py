"""Mirror of sunflower/utils/conversion.py:54-58 on the flope_b200 pose-head kernel."""
import torch

from . import _lib

_engines = {}


def _engine_for(device):
    """Head-only calls need no weights; one tiny engine per device is kept for them."""
    idx = device.index or 0
    if idx not in _engines:
        _engines[idx] = _lib.Engine(idx, max_batch=1, crop_hw=32)
    return _engines[idx]


def procrustes_to_rotmat(inp: torch.Tensor) -> torch.Tensor:
    """(…,9) float32 cuda tensor -> (B,3,3) rotation matrices (special orthogonal Procrustes)."""
    r9 = inp.reshape(-1, 9).to(torch.float32).contiguous()
    if not r9.is_cuda:
        raise _lib.FlopeError("procrustes_to_rotmat runs on the GPU only (no CPU fallback); pass a cuda tensor")
    with torch.cuda.device(r9.device):
        R, _ = _engine_for(r9.device).pose_head(r9, want_yaw=False)
    return R


def nullify_yaw_batch_cuda(R: torch.Tensor) -> torch.Tensor:
    """(B,3,3) float32 cuda rotations -> (B,3,3) float64 yaw-nullified (sunflower/utils/mvg.py:240-251)."""
    R = R.reshape(-1, 9).to(torch.float32).contiguous()
    with torch.cuda.device(R.device):
        return _engine_for(R.device).nullify_yaw(R)
