"""Launches for the ncu capture of the ROI kernels: BASELINE configs[2] (64 x 1080p frames, 2048 crops, bilinear -> 224 into
the stem input, with and without mask) and the reference's crop mode (Lanczos4 -> 512, fp32 NCHW, mask, 256 crops).
    ncu --set full --clock-control none --import-source on -k regex:roi3 -s 6 -c 3 -o gpurun_out/r2_roi python tools/roi_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flope_b200 import _lib, synth
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames, masks, det = synth.frames_and_boxes(n_frames, 32, with_mask=True)
b5 = []
for f in range(n_frames):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = np.concatenate(b5)
fr, mk, bx = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(b5).cuda()
eng = _lib.Engine(0, max_batch=len(b5), crop_hw=224)
out512 = torch.empty((256, 3, 512, 512), device="cuda")
for _ in range(3):          # launches 0..5 warm up, 6..8 are captured
    eng.roi_crop(fr, mk, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
    eng.roi_crop(fr, None, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
    eng.roi_crop(fr, mk, bx[:256], 512, _lib.INTERP_LANCZOS4, out=out512)
torch.cuda.synchronize()
print("done", len(b5))
