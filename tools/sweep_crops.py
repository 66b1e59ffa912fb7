"""BASELINE configs[3]: N synthetic crops batch-sharded across the ranks (torchrun), micro-batches through the
engine, one final all_gather of the rotations.  Prints crops/s (max over ranks) on rank 0."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from flope_b200 import _lib, shard, synth

ap = argparse.ArgumentParser()
ap.add_argument("--crops", type=int, default=1_000_000)
ap.add_argument("--micro", type=int, default=2048)
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--inflight", type=int, default=1, help="micro-batches in flight per GPU (engines / streams); 1 = one engine with layer1-4 as one launch, the fastest mode")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from flope_b200.pipeline import EnginePool
engines = EnginePool(dev, n_engines=a.inflight, max_batch=a.micro, crop_hw=a.size, state_dict=synth.random_state_dict(0))
pool = torch.rand((2, a.micro, 3, a.size, a.size), device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
r9 = [torch.empty((a.micro, 9), device=dev) for _ in range(len(engines))]


def fn(lo, hi):                       # crops are cycled from a 2-micro-batch device pool (1 M x 602 KB would not fit)
    n = hi - lo
    out = torch.empty((n, 9), dtype=torch.float64, device=dev)
    if n:
        x = pool[(lo // a.micro) & 1, :n]

        def work(eng, k):             # micro-batches alternate between the engines / streams of the pool
            eng.posenet_forward(x, out=r9[k][:n])
            _lib.check(_lib.lib().flope_pose_head(eng._h, _lib._ptr(r9[k]), n, None, _lib._ptr(out), _lib._stream()))
        engines.submit(work)
    return out


shard.run_sharded(fn, min(a.crops, 4 * a.micro * world), a.micro, join=engines.join)          # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = shard.run_sharded(fn, a.crops, a.micro, join=engines.join)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"crops": a.crops, "n_gpus": world, "micro_batch": a.micro, "seconds": float(t) / 1e3,
                      "crops_per_s": a.crops / (float(t) / 1e3), "gathered_rows": int(res.shape[0])}))
if world > 1:
    dist.destroy_process_group()
