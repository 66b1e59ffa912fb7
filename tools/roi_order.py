"""How much of the ROI launch is schedule tail: the same 2048 crops (BASELINE configs[2]) in detector order, sorted by
box side (largest first = LPT for the dynamic item claim, smallest first = worst case), and item sizes 28/56/112 rows.
    python tools/roi_order.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flope_b200 import _lib, synth
frames, masks, det = synth.frames_and_boxes(64, 32, with_mask=True)
b5 = []
for f in range(64):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = np.concatenate(b5)
side = b5[:, 3] - b5[:, 1]
fr, mk = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda()
eng = _lib.Engine(0, max_batch=len(b5), crop_hw=224)


def t(bx, m, reps=15):
    for _ in range(3):
        eng.roi_crop(fr, m, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.roi_crop(fr, m, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


orders = {"detector order": np.arange(len(b5)), "largest first": np.argsort(-side, kind="stable"), "smallest first": np.argsort(side, kind="stable"),
          "shuffled": np.random.default_rng(0).permutation(len(b5))}
for rows in (56, 28, 112):
    eng.debug_set("roi_item_rows", rows)
    for name, o in orders.items():
        bx = torch.from_numpy(np.ascontiguousarray(b5[o])).cuda()
        print(f"item rows {rows:3d} {name:15s} mask {t(bx, mk):7.1f} us   no mask {t(bx, None):7.1f} us", flush=True)
