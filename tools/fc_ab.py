"""A/B of the fc tile shape (debug option fc_small) on the isolated-launch table and the step time (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flope_b200 import _lib, synth

sd = synth.random_state_dict(0)
for B in (256, 8):
    x = torch.rand((B, 3, 224, 224), device="cuda")
    ref = None
    for small in (0, 1):
        e = _lib.Engine(0, max_batch=B, crop_hw=224)
        e.debug_set("fc_small", small)
        e.load_state_dict(sd)
        for _ in range(5):
            out = e.posenet_forward(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                e.posenet_forward(x, out=out)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 20)
        e.profile(1)
        for _ in range(5):
            e.posenet_forward(x, out=out)
        fc = [t for n, t in e.profile_read() if n == "conv:fc"]
        e.profile(False)
        same = True if ref is None else bool(torch.equal(ref, out))
        ref = out.clone() if ref is None else ref
        print(f"B={B} fc_small={small}: step {min(ts)*1e3:.1f} us, fc isolated {sorted(fc)[len(fc)//2]*1e3:.1f} us, same bits as fc_small=0: {same}", flush=True)
        e.close()
