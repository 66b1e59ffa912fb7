"""Launch every kernel of the library once at bench-like sizes (after a warm-up pass) for one ncu capture:
ROI bilinear / Lanczos4 (streaming kernels in both layouts with and without mask, the Lanczos4 table pre-kernel, the generic
kernels a frame width that is not a multiple of 16 falls back to), ingest, the backbone (fused stem + trunk launch at 256
crops, stem + max-pool + per-stage shapes at 288, the latency-tile trunk launch with split-K at 8 crops), avgpool, fc, pose
head, depth branch, YOLO-seg mask post-processing.
    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r2_all python tools/all_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

B = 256
sd = synth.random_state_dict(0)
frames, masks, det = synth.frames_and_boxes(8, 32, with_mask=True)
rows = []
for f in range(8):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    rows.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = torch.from_numpy(np.concatenate(rows)).cuda()
fr, mk = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda()
eng = _lib.Engine(0, max_batch=B, crop_hw=224)
eng.debug_set("use_graph", 0)
eng.load_state_dict(sd)
eng288 = _lib.Engine(0, max_batch=32, crop_hw=288)
eng288.debug_set("use_graph", 0)
eng288.load_state_dict(sd)
eng512 = _lib.Engine(0, max_batch=8, crop_hw=512)
eng8 = _lib.Engine(0, max_batch=8, crop_hw=224)
eng8.debug_set("use_graph", 0)
eng8.load_state_dict(sd)
x8 = torch.rand((8, 3, 224, 224), device="cuda")
fr_odd = fr[:, :, :1912].contiguous()                   # W % 16 != 0: the generic ROI kernels
mk_odd = mk[:, :, :1912].contiguous()
b5_odd = b5[(b5[:, 3] <= 1912)][:64].contiguous()
out224 = torch.empty((64, 3, 224, 224), device="cuda")
x = torch.rand((B, 3, 224, 224), device="cuda")
x288 = torch.rand((32, 3, 288, 288), device="cuda")
out512 = torch.empty((64, 3, 512, 512), device="cuda")
depth = torch.from_numpy((np.random.default_rng(0).integers(2000, 20000, (1080, 1920))).astype(np.uint16)).cuda()
inst = (torch.rand((12, 384, 640), device="cuda") > 0.7).float()
boxes4 = b5[:32, 1:].contiguous()
for it in range(2):                                    # pass 0 warms up, pass 1 is the one to look at
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    eng.roi_crop(fr, mk, b5, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
    eng.roi_crop(fr, None, b5, 224, _lib.INTERP_LINEAR)
    eng.roi_crop(fr, mk, b5, 224, _lib.INTERP_LANCZOS4, out_fmt=_lib.OUT_ENGINE)
    eng512.roi_crop(fr, mk, b5[:64], 512, _lib.INTERP_LANCZOS4, out=out512)
    eng512.roi_crop(fr, None, b5[:64], 512, _lib.INTERP_LANCZOS4, out=out512)
    eng512.roi_crop(fr, mk, b5[:64], 512, _lib.INTERP_LINEAR, out=out512)
    eng512.roi_crop(fr, None, b5[:64], 512, _lib.INTERP_LINEAR, out=out512)
    eng.roi_crop(fr_odd, mk_odd, b5_odd, 224, _lib.INTERP_LINEAR, out=out224)
    eng.roi_crop(fr_odd, None, b5_odd, 224, _lib.INTERP_LANCZOS4, out=out224)
    eng8.posenet_forward(x8)
    r9 = eng.posenet_forward(x)
    eng.pose_head(r9)
    eng288.posenet_forward(x288)
    _lib.depth_values(depth, mk[0], boxes4, 0.1, 2.5, depth_div=10000.0)
    _lib.yolo_mask(inst, 1080, 1920)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
