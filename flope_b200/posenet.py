"""PoseResNet drop-in (mirrors sunflower/models/posenet.py:5-34) running on the flope_b200 engine.

Same constructor signature, same ``state_dict`` key schema (the reference's 124 entries load
unchanged through ``load_state_dict``), same call contract: ``(B,3,S,S)`` float32 in [0,1]
on the engine's device -> ``(B,9)`` float32.  Inference only (eval semantics: BatchNorm uses
running statistics, dropout is the identity - SURVEY.md section 0, D4).
"""
import torch

from . import _lib


class PoseResNet:
    def __init__(self, backbone_out_dim=2048, dropout=0.5, device="cuda:0", max_batch=256, crop_hw=224):
        if backbone_out_dim != 2048:
            raise ValueError("the flope_b200 engine implements the reference configuration backbone_out_dim=2048")
        self.dropout = dropout          # kept for signature compatibility; identity at inference
        self.device = torch.device(device)
        self.training = False
        self._engine = _lib.Engine(self.device.index or 0, max_batch, crop_hw)
        self._state = None

    # --- torch.nn.Module-shaped surface used by the reference's callers ---
    def to(self, device):
        if torch.device(device) != self.device:
            raise _lib.FlopeError("a flope_b200 engine is bound to the device it was created on")
        return self

    def eval(self):
        return self

    def load_state_dict(self, state_dict, strict=True):
        self._engine.load_state_dict(state_dict)
        self._state = {k: v for k, v in state_dict.items()}
        return self

    def state_dict(self):
        return self._state

    @property
    def engine(self):
        return self._engine

    def forward(self, x):
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected (B,3,S,S) float32")
        if x.shape[2] != self._engine.crop_hw or x.shape[3] != self._engine.crop_hw:
            raise ValueError(f"engine was created for {self._engine.crop_hw}x{self._engine.crop_hw} crops")
        x = x.to(self.device).contiguous()
        with torch.cuda.device(self.device):
            return self._engine.posenet_forward(x)

    __call__ = forward
