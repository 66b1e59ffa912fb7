"""BASELINE configs[4]: live_pose streaming - one 1080p frame, 8 flowers, p50/p99 of
[H2D frame+mask+boxes -> ROI crop -> PoseNet -> Procrustes -> yaw -> D2H rotations] on one B200."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for S, interp, name in ((224, _lib.INTERP_LINEAR, "bilinear-224"), (512, _lib.INTERP_LANCZOS4, "lanczos4-512 (reference mode)")):
    frames, masks, det = synth.frames_and_boxes(1, 8, with_mask=True)
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[0]), 1080, 1920)
    b5 = np.concatenate([np.zeros((len(sq), 1), np.int32), sq], 1)
    eng = _lib.Engine(0, max_batch=8, crop_hw=S)
    for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):
        eng.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
    eng.load_state_dict(synth.random_state_dict(0))
    hf, hm, hb = torch.from_numpy(frames).pin_memory(), torch.from_numpy(masks).pin_memory(), torch.from_numpy(b5).pin_memory()
    df, dm, db = hf.cuda(), hm.cuda(), hb.cuda()
    out_d = torch.empty((len(b5), 3, 3), dtype=torch.float64, device="cuda")
    out_h = torch.empty((len(b5), 3, 3), dtype=torch.float64).pin_memory()

    def once(copy=True):
        if copy:
            df.copy_(hf, non_blocking=True); dm.copy_(hm, non_blocking=True); db.copy_(hb, non_blocking=True)
        eng.infer_frames(df, dm, db, interp, want_R=False, want_yaw=True, out=out_d)
        out_h.copy_(out_d, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(20):
        once()
    res = {}
    for copy in (True, False):
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter(); once(copy); ts.append((time.perf_counter() - t0) * 1e3)
        ts = np.sort(ts)
        res["with_h2d_frame" if copy else "frame_resident"] = {"p50_ms": float(ts[len(ts) // 2]), "p99_ms": float(ts[int(len(ts) * 0.99)])}
    print(json.dumps({"config": f"1 frame 1080p, {len(b5)} flowers, {name}", **res}))
    eng.close()
