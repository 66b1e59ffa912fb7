"""A/B timing of the graph-replayed backbone step under debug toggles (run under gpurun).
usage: ab_step.py B S key=val[,key=val] [key=val ...]   -- one timing line per toggle set"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flope_b200 import _lib, synth

B = int(sys.argv[1]); S = int(sys.argv[2])
sd = synth.random_state_dict(0)
x = [torch.rand((B, 3, S, S), device="cuda") for _ in range(2)]
out = torch.empty((B, 9), device="cuda")
ref = None
for spec in sys.argv[3:] or ["pair=1"]:
    eng = _lib.Engine(0, max_batch=B, crop_hw=S)
    for kv in spec.split(","):
        k, v = kv.split("=")
        eng.debug_set(k, int(v))
    eng.load_state_dict(sd)
    for i in range(5):
        eng.posenet_forward(x[i & 1], out=out)
    torch.cuda.synchronize()
    r = eng.posenet_forward(x[0]).clone()
    if ref is None:
        ref = r
    same = bool(torch.equal(ref, r))
    best = 1e9
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            eng.posenet_forward(x[i & 1], out=out)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 20)
    print(f"{spec:30s} {best*1e3:8.1f} us/step  {B/best*1e3:9.0f} crops/s  bit-identical-to-first={same}", flush=True)
    eng.close()
