"""Streaming ROI kernel on BASELINE configs[2] (64 x 1080p frames, 2048 crops): sweep of the launch parameters.

usage: python tools/roi_sweep.py "rows,stage_kb,stages,ctas_per_sm,dynamic;..." [n_frames] [lanczos]
No L2 flush: the inputs (530 MB) and outputs (822 MB) of one launch are several times the L2."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

cfgs = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1].split(";")]
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lanczos = len(sys.argv) > 3
frames, masks, det = synth.frames_and_boxes(n_frames, 32, with_mask=True)
b5 = []
for f in range(n_frames):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = np.concatenate(b5)
if lanczos:
    b5 = b5[:256]
n = len(b5)
side = (b5[:, 3] - b5[:, 1]).astype(np.int64)
fr, mk, bx = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(b5).cuda()
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
S = 512 if lanczos else 224
eng = _lib.Engine(0, max_batch=8 if lanczos else n, crop_hw=S)
out = torch.empty((n, 3, 512, 512), device="cuda") if lanczos else None
obytes = 3145728 if lanczos else 301056


def run(m):
    if lanczos:
        eng.roi_crop(fr, m, bx, 512, _lib.INTERP_LANCZOS4, out=out)
    else:
        eng.roi_crop(fr, m, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)


for rows, kb, stages, per_sm, dyn in cfgs:
    eng.debug_set("roi_item_rows8" if lanczos else "roi_item_rows", rows); eng.debug_set("roi_stage_kb", kb); eng.debug_set("roi_stages", stages)
    eng.debug_set("roi_ctas_per_sm", per_sm); eng.debug_set("roi_dynamic", dyn)
    for mask_on in (True, False):
        m = mk if mask_on else None
        for _ in range(3):
            run(m)
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(m); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        byts = float(((4 if mask_on else 3) * side ** 2 + obytes + 20).sum())
        print(f"rows={rows} kb={kb} stages={stages} per_sm={per_sm} dyn={dyn} mask={int(mask_on)}: {n} crops {ms*1e3:.1f} us "
              f"{n/ms*1e3:.0f} crops/s {byts/ms/1e6:.0f} GB/s = {byts/ms/1e6/peak*100:.1f}% of measured HBM peak", flush=True)
