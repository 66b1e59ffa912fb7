"""Per-layer time / TFLOP/s table from one instrumented step (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flope_b200 import _lib, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 224
eng = _lib.Engine(0, max_batch=B, crop_hw=S)
for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):
    eng.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
eng.load_state_dict(synth.random_state_dict(0))
x = torch.rand((B, 3, S, S), device="cuda")
out = torch.empty((B, 9), device="cuda")
for _ in range(3):
    eng.posenet_forward(x, out=out)
torch.cuda.synchronize()
eng.profile(True)
R = 10
for _ in range(R):
    eng.posenet_forward(x, out=out)
prof = eng.profile_read()
eng.profile(False)
by = {}
order = []
for n, t in prof:
    if n not in by:
        order.append(n)
    by[n] = by.get(n, 0) + t / R
mm = {"conv:conv1": 118.0, "conv:fc": 1.05}
for s, c in ((1, 115.6),):
    for b in (0, 1):
        for cv in (1, 2):
            mm[f"conv:layer1.{b}.conv{cv}"] = 115.6
for L in (2, 3, 4):
    mm[f"conv:layer{L}.0.conv1"] = 57.8
    mm[f"conv:layer{L}.0.downsample"] = 6.4
    for nm in ("0.conv2", "1.conv1", "1.conv2"):
        mm[f"conv:layer{L}.{nm}"] = 115.6
scale = (S / 224.0) ** 2
tot = 0
for n in order:
    t = by[n]
    tot += t
    if n in mm:
        fl = mm[n] * 2e6 * B * (scale if n != "conv:fc" else 1)
        print(f"{n:28s} {t*1e3:8.1f} us  {fl / t / 1e9:8.1f} TFLOP/s")
    else:
        print(f"{n:28s} {t*1e3:8.1f} us")
print(f"total {tot*1e3:.1f} us  -> {B / tot * 1e3:.0f} crops/s")
