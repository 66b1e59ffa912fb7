"""A/B of one debug key of the streaming ROI kernels on BASELINE configs[2] (2048 crops, bilinear -> 224, stem layout).
    python tools/roi_ab.py roi_item_rows 56 28 14        (NF=8 NB=32 in the environment: 8 frames x 32 boxes instead of 64 x 32)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flope_b200 import _lib, synth
key, values = sys.argv[1], [int(v) for v in sys.argv[2:]]
NF, NB = int(os.environ.get("NF", 64)), int(os.environ.get("NB", 32))
frames, masks, det = synth.frames_and_boxes(NF, NB, with_mask=True)
b5 = []
for f in range(NF):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = np.concatenate(b5)
fr, mk, bx = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(b5).cuda()
eng = _lib.Engine(0, max_batch=len(b5), crop_hw=224)
if key == "roi_item_rows":
    eng.debug_set("roi_item_auto", 0)


def t(m, reps=25):
    for _ in range(3):
        eng.roi_crop(fr, m, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.roi_crop(fr, m, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


for rnd in range(3):
    for v in values:
        eng.debug_set(key, v)
        print(f"{key}={v}: mask {t(mk):7.1f} us   no mask {t(None):7.1f} us", flush=True)
