"""Per-launch times of one 8-crop forward (the streaming configuration): per-layer launches (where the time sits), the
trunk launch with and without split-K.   python tools/small_batch_profile.py [B]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flope_b200 import _lib, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = _lib.Engine(0, max_batch=B, crop_hw=224)
eng.load_state_dict(synth.random_state_dict(0))
x = torch.rand((B, 3, 224, 224), device="cuda")


def prof(tag, reps=30):
    for _ in range(5):
        eng.posenet_forward(x)
    torch.cuda.synchronize()
    eng.profile(True)
    for _ in range(reps):
        eng.posenet_forward(x)
    torch.cuda.synchronize()
    acc = collections.OrderedDict()
    for name, ms in eng.profile_read():
        acc.setdefault(name, []).append(ms)
    eng.profile(False)
    print("==", tag)
    tot = 0.0
    for name, v in acc.items():
        m = float(np.median(v)) * 1e3
        if "layer" in name or tag.startswith("per-layer"):
            print(f"   {name:50s} {m:8.1f} us")
        if not name.startswith("trunk") and not name.startswith("whole"):
            tot += m
    print(f"   sum {tot:.1f} us")


for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):
    eng.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
    eng.load_state_dict(synth.random_state_dict(0))
eng.debug_set("use_graph", 0)
eng.debug_set("chain", 0)
prof("per-layer launches")
eng.debug_set("chain", 1)
eng.debug_set("trunk_splitk", 0)
prof("trunk launch, no split-K")
for ks, stages in ((4, 15),):
    eng.debug_set("trunk_splitk", ks); eng.debug_set("trunk_split_stages", stages)
    prof(f"trunk launch, split-K up to {ks}, stage mask {stages}")
