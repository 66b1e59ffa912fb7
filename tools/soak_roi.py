"""Soak test of the streaming ROI kernels: thousands of launches over random box sets (1..300 boxes of every size, with
and without mask, both interpolation modes, fp32 and stem-layout outputs, changing launch geometries), two engines on two
streams with PoseNet steps in between, every result compared bit for bit with the generic kernel's result for the same
boxes.  Exercises the ring / item / self-resetting work-counter protocol under every interleaving the scheduler produces.
usage: python tools/soak_roi.py [seconds]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(5)
H, W = 720, 1280
frames = rng.integers(0, 256, (4, H, W, 3), dtype=np.uint8)
masks = (rng.random((4, H, W)) < 0.5).astype(np.uint8) * 255
masks[:2] = 0
yy, xx = np.ogrid[0:H, 0:W]
masks[0][((xx - W / 2) / (W / 3)) ** 2 + ((yy - H / 2) / (H / 3)) ** 2 <= 1] = 255
masks[1][100:600, 200:1100] = 255
fr, mk = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda()
sd = synth.random_state_dict(0)
engs = [_lib.Engine(0, max_batch=300, crop_hw=224) for _ in range(2)]
ref_eng = _lib.Engine(0, max_batch=300, crop_hw=224)
for e in engs + [ref_eng]:
    e.load_state_dict(sd)
ref_eng.debug_set("roi_stream", 0)
streams = [torch.cuda.Stream() for _ in engs]
geoms = [(56, 14, 3, 0, 1), (7, 1, 2, 1, 0), (28, 4, 4, 0, 1), (128, 16, 3, 2, 1), (13, 6, 8, 3, 1)]
t0, launches, bad = time.time(), 0, 0
it = 0
while time.time() - t0 < budget:
    n = int(rng.choice([1, 2, 7, 33, 150, 300]))
    side = np.where(rng.random(n) < 0.2, rng.integers(1, 12, n), rng.integers(12, H + 1, n))
    x0 = (rng.random(n) * (W - side + 1)).astype(np.int64); y0 = (rng.random(n) * (H - side + 1)).astype(np.int64)
    b5 = torch.from_numpy(np.stack([rng.integers(0, 4, n), x0, y0, x0 + side, y0 + side], 1).astype(np.int32)).cuda()
    interp = int(rng.integers(0, 2))
    m = mk if rng.random() < 0.7 else None
    k = it & 1
    e = engs[k]
    rows, kb, stages, per_sm, dyn = geoms[int(rng.integers(0, len(geoms)))]
    e.debug_set("roi_item_rows8" if interp else "roi_item_rows", rows); e.debug_set("roi_stage_kb", kb); e.debug_set("roi_stages", stages)
    e.debug_set("roi_ctas_per_sm", per_sm); e.debug_set("roi_dynamic", dyn)
    want32 = ref_eng.roi_crop(fr, m, b5, 224, interp)
    ref_eng.roi_crop(fr, m, b5, 224, interp, out_fmt=_lib.OUT_ENGINE)
    want9 = ref_eng.posenet_forward(None, n=n).clone()
    torch.cuda.synchronize()
    with torch.cuda.stream(streams[k]):
        for rep in range(3):
            got32 = e.roi_crop(fr, m, b5, 224, interp)
            e.roi_crop(fr, m, b5, 224, interp, out_fmt=_lib.OUT_ENGINE)
            got9 = e.posenet_forward(None, n=n)
            launches += 2
            if not (torch.equal(got32, want32) and torch.equal(got9, want9)):
                bad += 1
                print("MISMATCH", it, rep, n, interp, m is not None, (rows, kb, stages, per_sm, dyn), flush=True)
    it += 1
torch.cuda.synchronize()
print(f"{launches} streaming ROI launches in {time.time() - t0:.0f} s over {it} random box sets: {bad} mismatches")
sys.exit(1 if bad else 0)
