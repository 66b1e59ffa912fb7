"""Soak test: many back-to-back steps (CUDA graph replay + programmatic dependent launch + CTA-pair kernels) with several
engines in flight, every result compared bit for bit with the first one.  Prints the number of mismatching steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flope_b200 import synth
from flope_b200.pipeline import EnginePool

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ne = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
pool = EnginePool(dev, n_engines=ne, max_batch=B, crop_hw=224, state_dict=synth.random_state_dict(0))
for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):      # e.g. FLOPE_SET=chain_coop=0 (no re-pack needed)
    for e in pool.engines:
        e.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
xs = [synth.mixed_crops(B, 224, seed=100 + i).to(dev) for i in range(2)]
outs = [[torch.empty((B, 9), device=dev) for _ in range(2)] for _ in range(ne)]
ref = []
for i in range(2):
    ref.append(pool.engines[0].posenet_forward(xs[i]).clone())
torch.cuda.synchronize()
bad = torch.zeros((), dtype=torch.int64, device=dev)
t0 = time.time()
for i in range(steps):
    def work(e, k, i=i):
        o = outs[k][i & 1]
        e.posenet_forward(xs[i & 1], out=o)
        bad.add_((o != ref[i & 1]).any().to(torch.int64))      # on the engine's stream, right behind the step
    pool.submit(work)
pool.join()
torch.cuda.synchronize()
print(f"{steps} steps of {B} crops, {ne} engines in flight: {int(bad)} mismatching steps, {time.time() - t0:.1f} s")
sys.exit(1 if int(bad) else 0)
