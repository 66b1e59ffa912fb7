"""Tiny forward through every kernel family (fused and unfused stem, pair and single-CTA convs, both ROI kernels,
pose head) for compute-sanitizer runs:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

sd = synth.random_state_dict(0)
for S, B in ((64, 5), (288, 2)):                      # 288: crop side above 252 -> separate stem + max-pool kernels
    for pair in (1, 0):
        eng = _lib.Engine(0, max_batch=B, crop_hw=S)
        eng.debug_set("pair", pair)
        eng.load_state_dict(sd)
        x = torch.rand((B, 3, S, S), device="cuda")
        r9 = eng.posenet_forward(x)
        R, Ry = eng.pose_head(r9)
        frames, masks, det = synth.frames_and_boxes(1, B, H=360, W=640, seed=5)
        sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[0]), 360, 640)
        b5 = torch.from_numpy(np.concatenate([np.zeros((len(sq), 1), np.int32), sq], 1)[:B]).cuda()
        fr, mk = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda()
        for interp in (_lib.INTERP_LINEAR, _lib.INTERP_LANCZOS4):
            eng.roi_crop(fr, mk, b5, S, interp)
            eng.roi_crop(fr, None, b5, S, interp, out_fmt=_lib.OUT_ENGINE)
        torch.cuda.synchronize()
        print(f"S={S} pair={pair}: ok, |r9| = {float(r9.abs().sum()):.4f}", flush=True)
        eng.close()
