"""Staging of step i+1 (float32 ingest, or ROI crop with --roi) overlapped with the backbone of step i:
EnginePool(serial_backbones=True) against one engine on one stream (run under gpurun).
usage: ab_staged.py [B] [S] [--roi]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth
from flope_b200.pipeline import EnginePool

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if args else 256
S = int(args[1]) if len(args) > 1 else 224
roi = "--roi" in sys.argv
dev = torch.device("cuda:0")
sd = synth.random_state_dict(0)
xs = [torch.rand((B, 3, S, S), device=dev) for _ in range(2)]
if roi:
    nf = max(1, B // 32)
    frames, masks, det = synth.frames_and_boxes(nf, 32, with_mask=True)
    b5 = []
    for f in range(nf):
        sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
        b5.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
    b5 = torch.from_numpy(np.concatenate(b5)[:B]).to(dev)
    fr, mk = torch.from_numpy(frames).to(dev), torch.from_numpy(masks).to(dev)
    n = len(b5)
else:
    n = B
for ne in (1, 2, 3):
    pool = EnginePool(dev, n_engines=ne, max_batch=B, crop_hw=S, state_dict=sd, serial_backbones=True)
    for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):
        for e in pool.engines:
            e.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
    outs = [torch.empty((n, 9), device=dev) for _ in range(ne)]

    def step(i):
        if roi:
            stage = lambda e, k: e.roi_crop(fr, mk, b5, S, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
        else:
            stage = lambda e, k: e.ingest_crops(xs[i & 1])
        pool.submit_staged(stage, lambda e, k: e.posenet_forward(None, n=n, out=outs[k]))

    for i in range(6):
        step(i)
    pool.join(); torch.cuda.synchronize()
    ref = outs[0].clone()
    best = 1e9
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(40):
            step(i)
        pool.join()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 40)
    same = all(bool(torch.equal(o, ref)) for o in outs) if roi else True
    print(f"{'roi' if roi else 'ingest'} staged, {ne} engine(s): {best*1e3:8.1f} us/step  {n/best*1e3:9.0f} crops/s  same-bits={same}", flush=True)
    pool.close()
