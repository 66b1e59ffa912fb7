"""BASELINE configs[2]: full-frame pipeline - 1920x1080 uint8 frames, 32 synthetic YOLO boxes per frame, 64 frames per
batch (2048 crops): ROI crop (bilinear 224, masks) -> PoseNet -> Procrustes -> yaw, frames resident and from pinned host
memory (double-buffered H2D).  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
frames, masks, det = synth.frames_and_boxes(n_frames, 32, with_mask=True)
rows = []
for f in range(n_frames):
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
    rows.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
b5 = np.concatenate(rows)
n = len(b5)
eng = _lib.Engine(0, max_batch=n, crop_hw=224)
eng.load_state_dict(synth.random_state_dict(0))
hf, hm, hb = torch.from_numpy(frames).pin_memory(), torch.from_numpy(masks).pin_memory(), torch.from_numpy(b5).pin_memory()
df = [torch.empty_like(hf, device="cuda") for _ in range(2)]
dm = [torch.empty_like(hm, device="cuda") for _ in range(2)]
db = [torch.empty_like(hb, device="cuda") for _ in range(2)]
do = [torch.empty((n, 3, 3), dtype=torch.float64, device="cuda") for _ in range(2)]
ho = [torch.empty((n, 3, 3), dtype=torch.float64).pin_memory() for _ in range(2)]
for b in range(2):
    df[b].copy_(hf); dm[b].copy_(hm); db[b].copy_(hb)
copy_s, comp_s = torch.cuda.Stream(), torch.cuda.Stream()
ev_c = [torch.cuda.Event() for _ in range(2)]
ev_d = [torch.cuda.Event() for _ in range(2)]


def run(steps, copy):
    for i in range(steps):
        b = i & 1
        if copy:
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ev_d[b])
                df[b].copy_(hf, non_blocking=True); dm[b].copy_(hm, non_blocking=True); db[b].copy_(hb, non_blocking=True)
                ev_c[b].record(copy_s)
        with torch.cuda.stream(comp_s):
            if copy:
                comp_s.wait_event(ev_c[b])
            eng.infer_frames(df[b], dm[b], db[b], _lib.INTERP_LINEAR, want_R=False, want_yaw=True, out=do[b])
            ho[b].copy_(do[b], non_blocking=True)
            ev_d[b].record(comp_s)
    copy_s.synchronize(); comp_s.synchronize()


res = {}
for copy in (False, True):
    run(3, copy)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(iters, copy)
    dt = time.perf_counter() - t0
    res["host_frames" if copy else "frames_resident"] = {"crops_per_s": n * iters / dt, "ms_per_batch": dt / iters * 1e3}
print(json.dumps({"config": f"{n_frames} frames 1080p x 32 boxes = {n} crops per batch, bilinear 224 + masks",
                  "h2d_bytes_per_batch": int(hf.numel() + hm.numel() + hb.numel() * 4), **res}))
