"""ROI kernel throughput on the frame configuration (BASELINE configs[2]): 1080p frames, 32 boxes each.

usage: python tools/roi_bench.py [n_frames] [quick]
Prints, per kernel variant, the time per launch (CUDA events, L2 flushed between launches), crops/s and the
algorithmic HBM bandwidth (source box + mask + output bytes) against the measured peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
quick = len(sys.argv) > 2
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


def report(tag, ms, byts, nn):
    print(f"{tag}: {nn} crops {ms*1e3:.1f} us  {nn/ms*1e3:.0f} crops/s  {byts/ms/1e6:.0f} GB/s = "
          f"{byts/ms/1e6/peak*100:.1f}% of measured HBM peak ({byts/nn:.0f} B/crop)", flush=True)


eng = _lib.Engine(0, max_batch=n, crop_hw=224)
by_m = float((4 * side ** 2 + 301056 + 20).sum())
by_n = float((3 * side ** 2 + 301056 + 20).sum())
scfgs = [(28, 10, 3, 0, 1)] if quick else [(28, 10, 3, 0, 1), (28, 10, 3, 0, 0), (16, 10, 3, 0, 1), (14, 10, 3, 0, 1), (56, 10, 3, 0, 1),
                                           (28, 6, 4, 0, 1), (28, 8, 3, 0, 1), (28, 10, 3, 3, 1), (28, 10, 3, 5, 1), (28, 6, 3, 5, 1), (28, 6, 3, 6, 1)]
for rows, kb, stages, per_sm, dyn in scfgs:
    eng.debug_set("roi_item_rows", rows); eng.debug_set("roi_stage_kb", kb); eng.debug_set("roi_stages", stages)
    eng.debug_set("roi_ctas_per_sm", per_sm); eng.debug_set("roi_dynamic", dyn)
    for mask_on in (True, False):
        m = mk if mask_on else None
        ms = timeit(lambda: eng.roi_crop(fr, m, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE))
        report(f"stream rows={rows} kb={kb} stages={stages} per_sm={per_sm} dyn={dyn} bilinear->224 bf16 engine mask={int(mask_on)}", ms,
               by_m if mask_on else by_n, n)
eng.debug_set("roi_stream", 0)
ms = timeit(lambda: eng.roi_crop(fr, mk, bx, 224, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE))
report("generic  bilinear->224 bf16 engine mask=1", ms, by_m, n)
eng.close()

nb = min(n, 256)
eng = _lib.Engine(0, max_batch=8, crop_hw=512)
out = torch.empty((nb, 3, 512, 512), device="cuda")
s = side[:nb]
byts = float((4 * s ** 2 + 3145728 + 20).sum())
ms = timeit(lambda: eng.roi_crop(fr, mk, bx[:nb], 512, _lib.INTERP_LINEAR, out=out), reps=5)
report("stream bilinear->512 f32 NCHW mask=1", ms, byts, nb)
for rows, kb, stages in ([(128, 10, 3)] if quick else [(128, 10, 3), (64, 10, 3), (128, 16, 3), (128, 6, 4), (86, 10, 3)]):
    eng.debug_set("roi_item_rows8", rows); eng.debug_set("roi_stage_kb", kb); eng.debug_set("roi_stages", stages)
    for mask_on in (True, False):
        ms = timeit(lambda: eng.roi_crop(fr, mk if mask_on else None, bx[:nb], 512, _lib.INTERP_LANCZOS4, out=out), reps=5)
        report(f"stream rows={rows} kb={kb} stages={stages} lanczos4->512 f32 NCHW mask={int(mask_on)}", ms,
               byts if mask_on else float((3 * s ** 2 + 3145728 + 20).sum()), nb)
eng.debug_set("roi_stage_kb", 10); eng.debug_set("roi_stages", 3)
eng.debug_set("roi_stream", 0)
for interp, name in ((_lib.INTERP_LANCZOS4, "lanczos4"), (_lib.INTERP_LINEAR, "bilinear")):
    ms = timeit(lambda: eng.roi_crop(fr, mk, bx[:nb], 512, interp, out=out), reps=5)
    report(f"generic  {name}->512 f32 NCHW mask=1", ms, byts, nb)
eng.close()
