"""One warm-up + N PoseNet steps at the bench configuration, for ncu captures (run under gpurun)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flope_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--frames", type=int, default=0, help="also run the ROI kernel on this many 1080p frames x 32 boxes")
a = ap.parse_args()
eng = _lib.Engine(0, max_batch=a.batch, crop_hw=a.size)
for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):   # e.g. FLOPE_SET=fuse_pool=1,use_graph=0
    eng.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
eng.load_state_dict(synth.random_state_dict(0))
x = torch.rand((a.batch, 3, a.size, a.size), device="cuda")
out = torch.empty((a.batch, 9), device="cuda")
for _ in range(1 + a.steps):
    eng.posenet_forward(x, out=out)
    eng.pose_head(out)
torch.cuda.synchronize()
if a.frames:
    import numpy as np
    frames, masks, det = synth.frames_and_boxes(a.frames, 32, with_mask=True)
    b5 = []
    for f in range(a.frames):
        sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
        b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
    b5 = torch.from_numpy(np.concatenate(b5)[: a.batch]).cuda()
    fr, mk = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda()
    for _ in range(2):
        eng.roi_crop(fr, mk, b5, a.size, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
        eng.roi_crop(fr, None, b5, a.size, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
    torch.cuda.synchronize()
print("done")
