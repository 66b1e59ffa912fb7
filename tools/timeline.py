"""Where the time of each conv_igemm launch goes: per-CTA phase stamps (flope_debug_timeline, %globaltimer) of one
forward at the bench configuration, summarised per launch (run under gpurun).
usage: timeline.py [B] [S]   (FLOPE_SET=key=v,... applies)

Columns (microseconds, medians over the launch's CTAs - over the leader CTAs of the pairs where the MMA warp's stamps
are involved):
  span      last CTA exit - first CTA entry
  gap       this launch's first entry - the previous launch's last exit (negative: overlapped through PDL)
  skew      spread of the CTA entry times
  prolog    entry -> barriers/TMEM/bias ready
  depwait   griddepcontrol.wait (predecessor grid complete)
  first     -> first operands in shared memory (MMA warp's first full barrier)
  main      -> last MMA issued
  tail      last MMA issued -> last epilogue store issued
  life      CTA entry -> its last epilogue store
  idle      SM-time of the launch not covered by a CTA's life: mean over CTAs of (span - life)
(the stamp after the kernel's final barrier is not used for timing: ptxas schedules its timer read before the barrier.
 trunk_chain_kernel records no first-operand / first-accumulator stamps: nan in `first` and `main`.)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flope_b200 import _lib, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 224
eng = _lib.Engine(0, max_batch=B, crop_hw=S)
for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):
    eng.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
eng.load_state_dict(synth.random_state_dict(0))
x = torch.rand((B, 3, S, S), device="cuda")
for _ in range(3):
    eng.posenet_forward(x)
eng.debug_set("timeline", 1)
eng.debug_set("use_graph", 0)
for _ in range(2):
    eng.posenet_forward(x)
torch.cuda.synchronize()
t = eng.timeline().astype(np.int64)
print(f"{B} crops of {S}x{S}; {len(t)} conv launches")
print(f"{'launch':>6} {'CTAs':>5} {'span':>8} {'gap':>7} {'skew':>6} {'prolog':>7} {'depwait':>8} {'first':>6} {'main':>8} {'tail':>6} {'life':>7} {'idle':>6}")
prev_end = None
for i, L in enumerate(t):
    L = L[L[:, 0] != 0]
    if not len(L):
        continue

    def c(a, b):                         # CTAs that recorded both stamps (the MMA warp's exist in the leader CTAs only)
        rows = (L[:, a] != 0) & (L[:, b] != 0)
        return float(np.median((L[:, b] - L[:, a])[rows])) / 1e3 if rows.any() else float("nan")
    start, end = L[:, 0].min(), max(L[:, 6].max(), L[:, 7].max())
    gap = (start - prev_end) / 1e3 if prev_end is not None else float("nan")
    life = L[:, 6] - L[:, 0]
    print(f"{i:>6} {len(L):>5} {(end-start)/1e3:>8.1f} {gap:>7.1f} {(L[:,0].max()-start)/1e3:>6.1f} {c(0,1):>7.2f} {c(1,2):>8.2f} "
          f"{c(2,3):>6.2f} {c(3,4):>8.1f} {c(4,6):>6.2f} {float(np.median(life))/1e3:>7.1f} {float(((end-start)-life).mean())/1e3:>6.1f}")
    prev_end = end
raw = os.environ.get("TIMELINE_RAW")
if raw is not None:                      # stamps of the first CTAs of one launch, ns relative to the launch's first entry
    L = t[int(raw)]
    L = L[L[:, 0] != 0]
    print((L[:6] - L[:, 0].min()).tolist())
