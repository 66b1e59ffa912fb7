"""Throughput with two steps in flight: two engines (own workspace + CUDA graph each) on two streams, steps alternating
between them, so one step's partial tail waves overlap the other step's kernels (run under gpurun).
usage: ab_overlap.py B S [n_engines ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flope_b200 import _lib, synth

B = int(sys.argv[1]); S = int(sys.argv[2])
sd = synth.random_state_dict(0)
x = [torch.rand((B, 3, S, S), device="cuda") for _ in range(2)]
for ne in [int(v) for v in sys.argv[3:]] or [1, 2, 3]:
    engs = [_lib.Engine(0, max_batch=B, crop_hw=S) for _ in range(ne)]
    for e in engs:
        for kv in filter(None, os.environ.get("FLOPE_SET", "").split(",")):
            e.debug_set(kv.split("=")[0], int(kv.split("=")[1]))
        e.load_state_dict(sd)
    streams = [torch.cuda.Stream() for _ in range(ne)]
    outs = [torch.empty((B, 9), device="cuda") for _ in range(ne)]

    def run(steps):
        for i in range(steps):
            k = i % ne
            with torch.cuda.stream(streams[k]):
                engs[k].posenet_forward(x[i & 1], out=outs[k])

    run(3 * ne); torch.cuda.synchronize()
    ref = engs[0].posenet_forward(x[0]).clone(); torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s_ in streams:
            s_.wait_event(a)
        run(40)
        for s_ in streams:
            torch.cuda.current_stream().wait_stream(s_)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 40)
    with torch.cuda.stream(streams[-1]):
        chk = engs[-1].posenet_forward(x[0]).clone()
    torch.cuda.synchronize()
    print(f"{ne} engine(s)/stream(s): {best*1e3:8.1f} us/step  {B/best*1e3:9.0f} crops/s  same-bits={bool(torch.equal(ref, chk))}", flush=True)
    for e in engs:
        e.close()
