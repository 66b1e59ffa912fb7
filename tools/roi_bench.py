"""ROI kernel throughput on the frame configuration (BASELINE configs[2]): 64 x 1080p frames, 32 boxes each."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames, masks, det = synth.frames_and_boxes(n_frames, 32, with_mask=True)
b5 = []
for f in range(n_frames):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = np.concatenate(b5)
n = len(b5)
side = (b5[:, 3] - b5[:, 1]).astype(np.int64)
fr, mk, bx = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(b5).cuda()
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()                       # evict L2 (256 MB > 126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


eng = _lib.Engine(0, max_batch=n, crop_hw=224)
for strip, mask_on in ((14, True), (14, False), (8, True), (8, False), (4, True), (28, True)):
    eng.debug_set("roi_strip", strip)
    print(f"{strip}-row strips", end=": ")
    m = mk if mask_on else None
    ms = timeit(lambda: eng.roi_crop(fr, m, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE))
    byts = float((3 * side ** 2 + (side ** 2 if mask_on else 0) + 301056 + 20).sum())
    print(f"bilinear->224 bf16 engine fmt mask={mask_on}: {n} crops {ms*1e3:.1f} us  {n/ms*1e3:.0f} crops/s  "
          f"{byts/ms/1e6:.0f} GB/s = {byts/ms/1e6/peak*100:.1f}% of measured HBM peak ({byts/n:.0f} B/crop)")
eng.close()
nb = min(n, 512)
eng = _lib.Engine(0, max_batch=8, crop_hw=512)
out = torch.empty((nb, 3, 512, 512), device="cuda")
for interp, name in ((_lib.INTERP_LANCZOS4, "lanczos4"), (_lib.INTERP_LINEAR, "bilinear")):
    ms = timeit(lambda: eng.roi_crop(fr, mk, bx[:nb], 512, interp, out=out), reps=5)
    s = side[:nb]
    byts = float((4 * s ** 2 + 3145728 + 20).sum())
    print(f"{name}->512 f32 NCHW (reference layout) mask=True: {nb} crops {ms*1e3:.1f} us  {nb/ms*1e3:.0f} crops/s  "
          f"{byts/ms/1e6:.0f} GB/s = {byts/ms/1e6/peak*100:.1f}% of measured HBM peak")
