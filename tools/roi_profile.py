import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flope_b200 import _lib, synth
n_frames = 16
frames, masks, det = synth.frames_and_boxes(n_frames, 32, with_mask=True)
b5 = []
for f in range(n_frames):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = np.concatenate(b5)
fr, mk, bx = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(b5).cuda()
eng = _lib.Engine(0, max_batch=len(b5), crop_hw=224)
for _ in range(2):
    eng.roi_crop(fr, mk, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
    eng.roi_crop(fr, None, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
torch.cuda.synchronize()
print("done")
