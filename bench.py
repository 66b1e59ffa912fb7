#!/usr/bin/env python
"""bench.py - PoseNet crops/s of the flope_b200 pose path on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

Workload of `value` / `e2e` (BASELINE.json configs[1]): PoseNet, batch 256 synthetic 224x224 crops per GPU,
random-init weights (seed 0), bf16 tensor-core compute with fp32 accumulation.  One step = 256 crops through
PoseResNet.forward -> Procrustes -> yaw nullification (sunflower/models/posenet.py:31-34,
sunflower/utils/conversion.py:54-58, sunflower/utils/mvg.py:240-251).
  value : crops/s with the float32 crop batch already resident in HBM (two alternating 154 MB batches, so the inputs
          alone exceed the 126 MB L2 between consecutive steps).  A timed region is EXACTLY K steps between a
          barrier + synchronize on both sides (CUDA events, max over ranks); the region is repeated `timed_regions`
          times and the median region is reported, so that a 20-step x 0.9 ms region is not at the mercy of one hiccup.
  e2e   : the same 256 crops per step through the predictor-level call with HOST buffers (FastPosePredictor path =
          flope_infer_frames): 8 uint8 1080p frames + masks + boxes in ONE pinned arena -> one H2D copy -> ROI crop
          (bilinear 224) -> PoseNet -> pose head -> D2H of the (256,3,3) float64 rotations, every step.
          `h2d_ceiling` is a bare pinned copy of the same bytes on every rank at the same time: the e2e ceiling of the box.
  roofline : the dominant kernel (trunk_chain_kernel: layer1..layer4 in one persistent launch) against the measured bf16
          peak; roofline_roi / roofline_roi_lanczos512: the ROI kernels against the measured HBM peak.
Extra keys carry the other BASELINE configs: frames_pipeline (configs[2]), sweep_1m (configs[3], with the gathered rows
checked against rank 0's own results), latency (configs[4], through flope_infer_frames and through
FastPosePredictor.get_flower_poses), and the baselines: cpu_baseline (+ cpu_legs) and incumbent_gpu (the reference-shaped
torch module under torch eager / cuDNN on the same GPU) - rank 0 at N=1 only.
Multi-GPU: crops sharded by batch across ranks, no data-path collective, one final NCCL all_gather of the results.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CROP = {224: 3.6293e9, 512: 18.952e9}     # 2*MACs, SURVEY.md appendix A
METRIC = "posenet_crops_per_sec"
UNIT = "crops/s"
FH, FW = 1080, 1920


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")      # B200_PROFILING.md fallback


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while a region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid = uuid
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def median(v):
    v = sorted(v)
    return v[len(v) // 2]


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle, the only place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------
def cpu_path_setup(size):
    import torch
    from flope_b200 import synth
    from oracle import posenet as onet
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    try:
        import cv2
        cv2.setNumThreads(cores)
    except Exception:
        pass
    net = onet.build(synth.WEIGHT_SEED)
    return net, cores


def cpu_path_step(net, x):
    """One pass of the reference path on CPU: PoseNet fp32 eval -> Procrustes -> yaw nullification."""
    from oracle import posenet as onet
    from oracle import rotation as orot
    r9 = onet.forward_fp32(net, x, chunk=32)
    rot = orot.procrustes_to_rotmat(r9).numpy()
    return orot.nullify_yaw_batch(rot)


def cpu_baseline(size, batch, budget_s=10.0):
    from flope_b200 import synth
    net, cores = cpu_path_setup(size)
    n = min(batch, 32)
    x = synth.mixed_crops(n, size)
    cpu_path_step(net, x[:4])                                  # warm-up
    t0 = time.perf_counter()
    done = 0
    while True:
        cpu_path_step(net, x)
        done += n
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} crops ({n}-crop batches, {size}x{size}, fp32 torch-CPU eval PoseResNet + Procrustes + "
                      f"yaw nullification) in {dt:.1f} s"}


def cpu_legs(size):
    """The other CPU legs SURVEY.md section 8(d) lists, each a bounded sample of the oracle on the host cores:
    (i) B=1 fp32 latency (configs[0]), (iii) the crop step alone in both modes, (iv) the full pipeline on 1 frame x 32 boxes."""
    import numpy as np
    import torch
    from flope_b200 import synth
    from oracle import boxes as obox
    from oracle import pipeline as opipe
    from oracle import resize as ores
    net, cores = cpu_path_setup(size)
    out = {"cores": cores, "kind": "port"}
    x1 = synth.mixed_crops(1, size)
    for _ in range(2):
        cpu_path_step(net, x1)
    ts = []
    for _ in range(15):
        t0 = time.perf_counter(); cpu_path_step(net, x1); ts.append(time.perf_counter() - t0)
    out["b1_fp32_latency_ms_p50"] = median(ts) * 1e3
    frames, masks, det = synth.frames_and_boxes(1, 32, H=FH, W=FW, seed=synth.FRAME_SEED)
    sq, keep = obox.squarify_filter(det[0], frames[0].shape)
    for name, s, interp, nb in (("crop_lanczos4_512_ms_per_crop", 512, ores.LANCZOS4, 8), ("crop_bilinear_224_ms_per_crop", 224, ores.BILINEAR, 32)):
        ores.crop_batch_reference(frames[0], masks[0], sq[:2], size=s, interp=interp)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); ores.crop_batch_reference(frames[0], masks[0], sq[:nb], size=s, interp=interp)
            ts.append((time.perf_counter() - t0) / nb)
        out[name] = median(ts) * 1e3
    t0 = time.perf_counter()
    r = opipe.run(net, frames[0], masks[0], det[0], size=224, interp=ores.BILINEAR)
    dt = time.perf_counter() - t0
    out["pipeline_1frame_32boxes_bilinear224"] = {"crops": int(r["rot"].shape[0]), "seconds": dt, "crops_per_s": r["rot"].shape[0] / dt}
    t0 = time.perf_counter()
    r = opipe.run(net, frames[0], masks[0], det[0][:4], size=512, interp=ores.LANCZOS4)
    dt = time.perf_counter() - t0
    out["pipeline_1frame_4boxes_lanczos512"] = {"crops": int(r["rot"].shape[0]), "seconds": dt, "crops_per_s": r["rot"].shape[0] / dt,
                                                "note": "the reference's own crop mode (fast_pose_predictor.py:115-116)"}
    return out


def incumbent_gpu(size, batch, dev):
    """BASELINE LEG, not the product: the reference-shaped torch module (oracle.posenet, the restatement of
    sunflower/models/posenet.py:5-34) under torch eager / cuDNN on the same GPU, as self.posenet(image_batch)
    (fast_pose_predictor.py:126) would run it on a B200 - fp32 NCHW as written, and bf16 channels_last."""
    import torch
    from flope_b200 import synth
    from oracle import posenet as onet
    res = {"what": "torch %s eager (cuDNN/cuBLAS) on the reference-shaped PoseResNet, eval, batch %d @ %d, device-resident input; "
                   "baseline only" % (torch.__version__, batch, size)}
    x = synth.mixed_crops(batch, size).to(dev)
    for name, dt, cl in (("fp32_nchw", torch.float32, False), ("bf16_channels_last", torch.bfloat16, True)):
        net = onet.build(synth.WEIGHT_SEED).to(dev).eval()
        xin = x.to(dt)
        if cl:
            net = net.to(memory_format=torch.channels_last).to(dt)
            xin = xin.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            for _ in range(3):
                net(xin)
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); net(xin); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        res[name] = {"ms_per_batch": median(ts), "crops_per_s": batch / (median(ts) / 1e3)}
        del net
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from flope_b200 import synth
    net, cores = cpu_path_setup(args.size)
    probe = synth.mixed_crops(8, args.size)
    cpu_path_step(net, probe[:2])
    t0 = time.perf_counter()
    cpu_path_step(net, probe)
    per_crop = (time.perf_counter() - t0) / 8
    total_steps = args.steps + args.warmup
    n = int(max(1, min(args.batch, 150.0 / total_steps / per_crop)))       # whole run <= ~2.5 min
    x = synth.mixed_crops(n, args.size)
    for _ in range(args.warmup):
        cpu_path_step(net, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_path_step(net, x)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} of the workload's {args.batch} crops per step, {args.steps} steps, {cores} host threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    return {"workload": f"PoseNet bf16 batch {args.batch} synthetic {args.size}x{args.size} crops per GPU "
                        f"(BASELINE.json configs[1]); step = PoseResNet.forward + Procrustes + yaw nullification",
            "batch_per_gpu": args.batch, "crop_hw": args.size, "weights": "random-init seed 0",
            "l2_policy": "two alternating float32 input batches of %.0f MB each (> 126 MB L2), activations %.0f MB"
                         % (args.batch * 3 * args.size * args.size * 4 / 1e6,
                            args.batch * 4.7 * (args.size / 224.0) ** 2),
            "parallelism": f"dp{args.gpus} (crops sharded by batch, weights replicated, one final all_gather); "
                           f"{getattr(args, 'inflight', 1)} independent steps in flight per GPU (one engine + stream each)"}


def bind_to_gpu_numa_node(local_rank):
    """Restrict this process to the CPUs NVML reports as local to its GPU, so that the pinned staging buffers of the
    end-to-end legs are allocated on that socket (with 8 ranks the H2D copies otherwise cross the socket interconnect).
    Returns the number of CPUs bound to, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class FrameArena:
    """frames + masks + boxes of one step in ONE pinned host buffer and one device buffer: a step's upload is a single
    cudaMemcpyAsync.  Views keep the shapes the C ABI wants."""

    def __init__(self, torch, frames_np, masks_np, b5_np, dev, with_mask=True):
        import numpy as np
        fb, mb = frames_np.nbytes, (masks_np.nbytes if with_mask else 0)
        bb = b5_np.nbytes
        self.nbytes = fb + mb + ((bb + 15) & ~15)
        self.host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        h = self.host.numpy()
        h[:fb] = frames_np.reshape(-1)
        if with_mask:
            h[fb:fb + mb] = masks_np.reshape(-1)
        h[fb + mb:fb + mb + bb] = np.frombuffer(b5_np.tobytes(), np.uint8)
        self.frames = self.dev[:fb].view(frames_np.shape)
        self.masks = self.dev[fb:fb + mb].view(masks_np.shape) if with_mask else None
        self.boxes = self.dev[fb + mb:fb + mb + bb].view(torch.int32).view(b5_np.shape)

    def upload(self):
        self.dev.copy_(self.host, non_blocking=True)


def frame_batch(np, _lib, synth, n_fr, per_frame, seed):
    frames_np, masks_np, det = synth.frames_and_boxes(n_fr, per_frame, H=FH, W=FW, seed=seed)
    rows = []
    for f in range(n_fr):
        sq, keep = _lib.squarify_filter(np.ascontiguousarray(det[f]), FH, FW)       # host box logic is part of the call
        rows.append(np.concatenate([np.full((len(sq), 1), f, np.int32), sq], 1))
    return frames_np, masks_np, det, np.ascontiguousarray(np.concatenate(rows))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from flope_b200 import _lib, synth
    from flope_b200.posenet import PoseResNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa = bind_to_gpu_numa_node(local)                 # pinned host buffers are first-touched on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S, K, W = args.batch, args.size, args.steps, args.warmup
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    from flope_b200.pipeline import EnginePool
    sd = synth.random_state_dict(synth.WEIGHT_SEED)
    model = PoseResNet(device=str(dev), max_batch=B, crop_hw=S)
    model.load_state_dict(sd)
    eng = model.engine
    # `value`: args.inflight independent steps in flight (one engine + stream each, flope_b200/pipeline.py).  Default 1:
    # one engine whose layer1..layer4 run as one persistent launch beats two engines with per-layer launches
    pool = EnginePool(dev, n_engines=args.inflight, max_batch=B, crop_hw=S, state_dict=sd)

    xs = [synth.mixed_crops(B, S, seed=synth.CROP_SEED + 10 * rank + i).to(dev) for i in range(2)]
    r9 = torch.empty((B, 9), dtype=torch.float32, device=dev)
    r9s = [torch.empty((B, 9), dtype=torch.float32, device=dev) for _ in range(len(pool))]
    results = torch.empty((K, B, 9), dtype=torch.float64, device=dev)               # yaw-nullified rotations

    def step(i, out_slot):
        def work(e, k):
            e.posenet_forward(xs[i & 1], out=r9s[k])
            _lib.check(_lib.lib().flope_pose_head(e._h, _lib._ptr(r9s[k]), B, None, _lib._ptr(results[out_slot]),
                                                  _lib._stream()))
        pool.submit(work)

    for i in range(max(W, 2 * len(pool))):
        step(i, 0)
    pool.join()
    launches_per_step = 0
    pool.engines[0].posenet_forward(xs[0], out=r9s[0])                                 # the engines the timed steps run on
    launches_per_step += pool.engines[0].last_launches() + 1                           # + the pose-head launch
    torch.cuda.synchronize()
    gathered = torch.empty((world, K, B, 9), dtype=torch.float64, device=dev) if world > 1 else None

    uuid = str(torch.cuda.get_device_properties(local).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    sampler = ClockSampler(uuid)
    sampler.start()
    time.sleep(0.15)

    # ---- `value`: a region = EXACTLY K steps between barrier + synchronize; median of `regions` regions ----
    def timed_region():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(i, i)
        pool.join()
        if world > 1:
            dist.all_gather_into_tensor(gathered, results)          # the one collective: final gather, 72 B/crop
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    t_start = time.time()
    first_ms = timed_region()
    regions = max(1, min(40, int(math.ceil(0.6 / max(first_ms / 1e3, 1e-4)))))        # >= 0.6 s of timed work in total
    region_ms = [first_ms] + [timed_region() for _ in range(regions - 1)]
    t_end = time.time()
    ms = median(region_ms)
    clocks = sampler.summary(t_start, t_end) or {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    value = world * B * K / (ms / 1e3)

    # the same K steps strictly one after another on one stream / one engine (what a single in-order caller sees)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(K):
        eng.posenet_forward(xs[i & 1], out=r9)
        _lib.check(_lib.lib().flope_pose_head(eng._h, _lib._ptr(r9), B, None, _lib._ptr(results[i]), _lib._stream()))
    s1.record()
    torch.cuda.synchronize()
    ms_serial = max_over_ranks(s0.elapsed_time(s1))

    copy_s, comp_s = torch.cuda.Stream(), torch.cuda.Stream()
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]

    def timed_wall(fn, steps, reps=3):
        """median over `reps` of the wall time of fn(steps) between barriers (host-side API call included)."""
        fn(max(W, 3))
        ts = []
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            fn(steps)
            torch.cuda.synchronize()
            ts.append(max_over_ranks(time.perf_counter() - t0))
        return median(ts)

    # ---- end to end, module level: host float32 crops -> H2D -> module -> head -> D2H, double buffered ----
    h_in = [synth.mixed_crops(B, S, seed=synth.CROP_SEED + 10 * rank + i).pin_memory() for i in range(2)]
    d_in = [torch.empty_like(xs[0]) for _ in range(2)]
    h_out = [torch.empty((B, 3, 3), dtype=torch.float64).pin_memory() for _ in range(2)]
    d_out = [torch.empty((B, 3, 3), dtype=torch.float64, device=dev) for _ in range(2)]

    def e2e_module(steps):
        for i in range(steps):
            b = i & 1
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ev_done[b])                       # the buffer's previous consumer has finished
                d_in[b].copy_(h_in[b], non_blocking=True)
                ev_copied[b].record(copy_s)
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(ev_copied[b])
                out9 = model(d_in[b])                                # the drop-in module call
                _lib.check(_lib.lib().flope_pose_head(eng._h, _lib._ptr(out9), B, None, _lib._ptr(d_out[b]), _lib._stream()))
                h_out[b].copy_(d_out[b], non_blocking=True)
                ev_done[b].record(comp_s)
        copy_s.synchronize(); comp_s.synchronize()

    e2e_mod_value = world * B * K / timed_wall(e2e_module, K)
    h2d_mod = B * 3 * S * S * 4
    d2h = B * 9 * 8
    del h_in, d_in

    # ---- end to end, predictor level: host uint8 frames + masks + boxes -> poses ----
    per_frame = 32
    n_fr = max(1, B // per_frame)
    frames_np, masks_np, det, b5_np = frame_batch(np, _lib, synth, n_fr, per_frame, synth.FRAME_SEED + rank)
    nb = b5_np.shape[0]
    ho = [torch.empty((nb, 3, 3), dtype=torch.float64).pin_memory() for _ in range(2)]
    do = [torch.empty((nb, 3, 3), dtype=torch.float64, device=dev) for _ in range(2)]

    def make_e2e(arenas):
        def run(steps):
            for i in range(steps):
                b = i & 1
                a = arenas[b]
                with torch.cuda.stream(copy_s):
                    copy_s.wait_event(ev_done[b])
                    a.upload()                                       # ONE cudaMemcpyAsync: frames + masks + boxes
                    ev_copied[b].record(copy_s)
                with torch.cuda.stream(comp_s):
                    comp_s.wait_event(ev_copied[b])
                    eng.infer_frames(a.frames, a.masks, a.boxes, _lib.INTERP_LINEAR, want_R=False, want_yaw=True, out=do[b])
                    ho[b].copy_(do[b], non_blocking=True)
                    ev_done[b].record(comp_s)
            copy_s.synchronize(); comp_s.synchronize()
        return run

    arenas = [FrameArena(torch, frames_np, masks_np, b5_np, dev, True) for _ in range(2)]
    e2e_value = world * nb * K / timed_wall(make_e2e(arenas), K)
    h2d = arenas[0].nbytes
    d2h_fr = nb * 9 * 8
    arenas_nm = [FrameArena(torch, frames_np, masks_np, b5_np, dev, False) for _ in range(2)]
    e2e_nomask_value = world * nb * K / timed_wall(make_e2e(arenas_nm), K)
    h2d_nm = arenas_nm[0].nbytes

    # the box's H2D ceiling: the same pinned bytes, bare copies, every rank at the same time
    def bare_copies(steps):
        for i in range(steps):
            arenas[i & 1].upload()
        torch.cuda.synchronize()
    t_copy = timed_wall(bare_copies, K)
    h2d_gbs = world * h2d * K / t_copy / 1e9
    ceiling_crops = world * nb * K / t_copy
    h2d_ceiling = {"aggregate_gbs": h2d_gbs, "per_rank_gbs": h2d_gbs / world, "bytes_per_step": h2d,
                   "crops_per_s_if_copy_bound": ceiling_crops, "e2e_frac_of_ceiling": e2e_value / ceiling_crops,
                   "method": "bare pinned cudaMemcpyAsync of the step's arena (frames + masks + boxes), %d copies, all %d ranks "
                             "concurrently, wall clock between barriers" % (K, world)}
    del arenas_nm

    # ---- ROI kernel rooflines (HBM) ----
    def time_kernel(fn, reps, flush=None):
        # these legs follow host-side work (engine creation, weight folding) during which the GPU idles: launch for ~40 ms
        # first so that the clocks are back up (a kernel timed alone is compared with the burst peak)
        t_end = time.perf_counter() + 0.04
        while time.perf_counter() < t_end:
            fn()
            torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            if flush is not None:
                flush.zero_()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(); fn(); eb.record(); torch.cuda.synchronize()
            ts.append(ea.elapsed_time(eb))
        return median(ts)

    def roi_entry(kernel, nbytes, n, ms_, l2):
        gbs = nbytes / (ms_ / 1e3) / 1e9
        return {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                "traffic": None, "algorithmic_bytes_per_launch": nbytes, "crops_per_launch": int(n), "us_per_launch": ms_ * 1e3, "l2": l2}

    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath))
        except Exception:
            traffic = {}
    side = (b5_np[:, 3] - b5_np[:, 1]).astype(np.int64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    a0 = arenas[0]
    a0.upload(); torch.cuda.synchronize()
    small = {}
    for tag, m in (("mask", a0.masks), ("nomask", None)):
        eng.roi_crop(a0.frames, m, a0.boxes, S, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
        t = time_kernel(lambda: eng.roi_crop(a0.frames, m, a0.boxes, S, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE), 9, flush)
        nbytes = float(((4 if m is not None else 3) * side ** 2 + S * S * 3 * 2 + 20).sum())
        small[tag] = roi_entry("roi3_kernel<2,%s,bf16 engine layout> (bilinear -> %d)" % ("mask" if m is not None else "no mask", S),
                               nbytes, nb, t, "flushed (256 MB write) before every timed launch")
    # configs[2]: 64 frames x 32 boxes = 2048 crops per launch; inputs (530 MB) and outputs (822 MB) are several L2s
    nf2 = args.frames
    frames2, masks2, det2, b5_2 = frame_batch(np, _lib, synth, nf2, per_frame, synth.FRAME_SEED + 100 + rank)
    big = FrameArena(torch, frames2, masks2, b5_2, dev, True)
    big.upload(); torch.cuda.synchronize()
    n2 = b5_2.shape[0]
    side2 = (b5_2[:, 3] - b5_2[:, 1]).astype(np.int64)
    eng2 = _lib.Engine(local, max_batch=n2, crop_hw=S)
    eng2.load_state_dict(sd)
    roi_big = {}
    for tag, m in (("mask", big.masks), ("nomask", None)):
        for _ in range(2):
            eng2.roi_crop(big.frames, m, big.boxes, S, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE)
        t = time_kernel(lambda: eng2.roi_crop(big.frames, m, big.boxes, S, _lib.INTERP_LINEAR, out_fmt=_lib.OUT_ENGINE), 15)
        nbytes = float(((4 if m is not None else 3) * side2 ** 2 + S * S * 3 * 2 + 20).sum())
        roi_big[tag] = roi_entry("roi3_kernel<2,%s,bf16 engine layout> (bilinear -> %d)" % ("mask" if m is not None else "no mask", S),
                                 nbytes, n2, t, "not flushed: one launch reads %.0f MB and writes %.0f MB (L2 is 126 MB)"
                                 % (nbytes / 1e6 - n2 * S * S * 6 / 1e6, n2 * S * S * 8 / 1e6))
    roofline_roi = dict(roi_big["mask"])
    roofline_roi["traffic"] = traffic.get("roi_bilinear224_mask_bytes_per_launch_2048")
    roofline_roi["workload"] = "BASELINE configs[2]: %d x 1080p frames, %d boxes each" % (nf2, per_frame)
    roofline_roi["no_mask"] = roi_big["nomask"]
    roofline_roi["batch_%d_mask" % nb] = small["mask"]
    roofline_roi["batch_%d_no_mask" % nb] = small["nomask"]
    roofline_roi["note"] = ("the kernel writes the stem's 8-byte (3 channels + zero pad) bf16 pixels: real DRAM traffic is "
                            "4/3 of the algorithmic output bytes counted here")
    # the reference's own crop mode: Lanczos4 -> 512, fp32 NCHW, with mask (fast_pose_predictor.py:115-116)
    nl = min(nb, 256)
    out512 = torch.empty((nl, 3, 512, 512), dtype=torch.float32, device=dev)
    eng.roi_crop(a0.frames, a0.masks, a0.boxes[:nl], 512, _lib.INTERP_LANCZOS4, out=out512)
    t = time_kernel(lambda: eng.roi_crop(a0.frames, a0.masks, a0.boxes[:nl], 512, _lib.INTERP_LANCZOS4, out=out512), 7)
    nbytes = float((4 * side[:nl] ** 2 + 512 * 512 * 3 * 4 + 20).sum())
    roofline_roi_l = roi_entry("roi3_kernel<8,mask,f32 NCHW> (Lanczos4 -> 512, the reference's crop mode)", nbytes, nl, t,
                               "not flushed: one launch writes %.0f MB" % (nl * 3.145728))
    roofline_roi_l["traffic"] = traffic.get("roi_lanczos512_mask_bytes_per_launch_256")
    t = time_kernel(lambda: eng.roi_crop(a0.frames, a0.masks, a0.boxes[:nl], 512, _lib.INTERP_LINEAR, out=out512), 7)
    roofline_roi_l["bilinear_512_f32_nchw"] = roi_entry("roi3_kernel<2,mask,f32 NCHW> (bilinear -> 512)", nbytes, nl, t, "not flushed")
    del out512

    # ---- frames_pipeline (BASELINE configs[2]): ROI crop + PoseNet + head over 64 frames x 32 boxes per step ----
    do2 = torch.empty((n2, 3, 3), dtype=torch.float64, device=dev)
    ho2 = torch.empty((n2, 3, 3), dtype=torch.float64).pin_memory()

    def frames_resident(steps):
        for _ in range(steps):
            eng2.infer_frames(big.frames, big.masks, big.boxes, _lib.INTERP_LINEAR, want_R=False, want_yaw=True, out=do2)
        torch.cuda.synchronize()

    big2 = FrameArena(torch, frames2, masks2, b5_2, dev, True)
    do2s = [do2, torch.empty_like(do2)]
    ho2s = [ho2, torch.empty((n2, 3, 3), dtype=torch.float64).pin_memory()]

    def frames_e2e(steps):                                   # double buffered like `e2e`: upload of step i+1 under compute of step i
        for i in range(steps):
            b = i & 1
            a = (big, big2)[b]
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ev_done[b])
                a.upload()
                ev_copied[b].record(copy_s)
            with torch.cuda.stream(comp_s):
                comp_s.wait_event(ev_copied[b])
                eng2.infer_frames(a.frames, a.masks, a.boxes, _lib.INTERP_LINEAR, want_R=False, want_yaw=True, out=do2s[b])
                ho2s[b].copy_(do2s[b], non_blocking=True)
                ev_done[b].record(comp_s)
        copy_s.synchronize(); comp_s.synchronize()

    fp_steps = 6
    t_res = timed_wall(frames_resident, fp_steps)
    t_e2e = timed_wall(frames_e2e, fp_steps)
    frames_pipeline = {"workload": "BASELINE configs[2]: %d x 1080p uint8 frames + masks per GPU and step, %d boxes per frame -> "
                                   "ROI crop (bilinear %d) -> PoseNet -> Procrustes -> yaw" % (nf2, per_frame, S),
                       "crops_per_step_per_gpu": int(n2), "steps": fp_steps,
                       "value": world * n2 * fp_steps / t_res, "unit": UNIT, "ms_per_step": t_res / fp_steps * 1e3,
                       "e2e": {"value": world * n2 * fp_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": big.nbytes,
                               "d2h_bytes_per_step": n2 * 72, "note": "one pinned arena upload per step, double buffered"}}
    eng2.close()
    del big, big2, do2, do2s

    # ---- sweep_1m (BASELINE configs[3]): 1M synthetic crops sharded by batch over the ranks, one final gather ----
    sweep = None
    if args.sweep > 0:
        P, MB = 16, B                                            # pool of 16 micro-batches, cycled
        pool_x = [synth.mixed_crops(MB, S, seed=synth.CROP_SEED + 1000 + i).to(dev) for i in range(P)]    # the same pool on every rank
        per_rank = (args.sweep + world - 1) // world
        n_mb = (per_rank + MB - 1) // MB                         # micro-batches per rank (last shard padded)
        out_l = torch.empty((n_mb * MB, 9), dtype=torch.float32, device=dev)
        first_mb = rank * n_mb                                   # global micro-batch index of this rank's first batch
        gath = torch.empty((world * n_mb * MB, 9), dtype=torch.float32, device=dev) if world > 1 else out_l
        for j in range(3):
            eng.posenet_forward(pool_x[j % P], out=out_l[:MB])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(n_mb):
            eng.posenet_forward(pool_x[(first_mb + j) % P], out=out_l[j * MB:(j + 1) * MB])
        if world > 1:
            dist.all_gather_into_tensor(gath, out_l)             # rank-major = original index order
        e1.record()
        barrier()
        t_sweep = max_over_ranks(e0.elapsed_time(e1)) / 1e3
        # every gathered row against this rank's own result for the same pool crop (bitwise): order and content
        ref = torch.cat([eng.posenet_forward(pool_x[j]).clone() for j in range(P)])               # (P*MB, 9)
        g = torch.arange(gath.shape[0], device=dev)
        same = bool(torch.equal(gath, ref[g % (P * MB)]))
        total = world * n_mb * MB
        sweep = {"workload": "BASELINE configs[3]: %d synthetic crops (cycled from a %d-crop device pool), sharded by batch: "
                             "rank r takes micro-batches [r*%d, (r+1)*%d) of %d crops" % (args.sweep, P * MB, n_mb, n_mb, MB),
                 "crops": int(total), "seconds": t_sweep, "value": total / t_sweep, "unit": UNIT, "scaling": "strong",
                 "collective": "one all_gather_into_tensor of (n,9) float32 inside the timed region" if world > 1 else "none (1 rank)",
                 "gathered_rows_equal_single_gpu": same,
                 "check": "all %d gathered rows bitwise equal to rank %d's own forward of the same pool crops, in index order" % (total, rank)}
        del pool_x, out_l, gath, ref

    # ---- roofline of the dominant kernel (trunk_chain_kernel) and of the trunk launches of a step ----
    flop = FLOP_PER_CROP.get(S, 3.6293e9 * (S / 224.0) ** 2) * B
    fc_flop = 2.0 * 512 * 2048 * B
    stem_flop = 2.0 * 147 * 64 * (S // 2) ** 2 * B
    prof_steps = 10
    eng.profile(2)
    for i in range(prof_steps):
        eng.posenet_forward(xs[i & 1], out=r9)
    chain = [t for name, t in eng.profile_read() if name == "trunk"]
    eng.profile(False)
    chain_ms = median(chain)
    eng.profile(1)
    for i in range(prof_steps):
        eng.posenet_forward(xs[i & 1], out=r9)
    prof = eng.profile_read()
    eng.profile(False)
    by = {}
    for name, t in prof:
        by[name] = by.get(name, 0.0) + t / prof_steps
    conv_ms = sum(t for n_, t in by.items() if n_.startswith("conv:") and n_ != "conv:fc")
    chain_launches = sum(1 for n_ in by if n_.startswith("conv:") and n_ != "conv:fc")
    trunk_tf = (flop - fc_flop) / (chain_ms / 1e3) / 1e12
    step_ms = ms / K
    trunk_entry = {"kernel": "the %d trunk launches of one step (conv_igemm_kernel<stem+maxpool> + trunk_chain_kernel), back to back; fc excluded"
                             % chain_launches, "achieved": trunk_tf, "frac": trunk_tf / peaks["tf_burst"],
                   "frac_of_sustained_peak": trunk_tf / peaks["tf_sust"] if peaks["tf_sust"] else None,
                   "launches": chain_launches, "trunk_ms_per_step": chain_ms, "conv_ms_per_step_isolated": conv_ms,
                   "share_of_step": chain_ms / step_ms}
    one = [(n_, t) for n_, t in by.items() if n_.startswith("conv:layer1-4")]
    if one:
        dom_ms = one[0][1]
        dom_flop = flop - fc_flop - stem_flop                     # layer1..layer4 = everything but the stem and the fc
        tf = dom_flop / (dom_ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "trunk_chain_kernel (layer1..layer4: 16 convolutions + residuals in one persistent launch)",
                    "achieved": tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": tf / peaks["tf_burst"],
                    "frac_of_sustained_peak": tf / peaks["tf_sust"] if peaks["tf_sust"] else None,
                    "traffic": traffic.get(f"trunk_chain_bytes_per_launch_b{B}_s{S}"),
                    "algorithmic_flop_per_launch": dom_flop, "ms_per_launch": dom_ms, "launches_per_step": 1,
                    "share_of_step": dom_ms / step_ms,
                    "method": "algorithmic FLOP (SURVEY.md appendix A, un-padded) / CUDA-event time of the launch on the launch stream, "
                              "mean of %d steps, isolated launch" % prof_steps}
    else:
        roofline = {"bound": "tensor", "kernel": trunk_entry["kernel"], "achieved": trunk_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                    "frac": trunk_tf / peaks["tf_burst"], "traffic": None}
    roofline["peak_source"] = peaks["src"] + " (MEASURED_PEAKS.json burst bf16)" if peaks["src"] == "measured" else "fallback 1.59 PFLOP/s"
    roofline["stem_and_trunk"] = trunk_entry
    roofline["whole_step"] = {"achieved": flop / (step_ms / 1e3) / 1e12, "frac": (flop / (step_ms / 1e3) / 1e12) / peaks["tf_burst"],
                              "algorithmic_flop_per_step": flop}
    roofline["per_kernel_ms_isolated"] = {k: round(v, 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:10]}

    sampler.stop()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "timed_regions": regions, "region_ms": {"median": ms, "min": min(region_ms), "max": max(region_ms), "first": first_ms},
            "steps_in_flight": len(pool), "host_cpus_bound": numa,
            "value_single_stream": world * B * K / (ms_serial / 1e3), "ms_per_step_single_stream": ms_serial / K,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_fr,
                    "crops_per_step": int(nb),
                    "api": "predictor path (flope_b200.predictor / flope_infer_frames): %d pinned host uint8 1080p frames + "
                           "masks + %d boxes (one arena, one copy) -> ROI crop (bilinear %d) -> PoseNet -> Procrustes -> yaw -> host "
                           "float64 rotations" % (n_fr, nb, S)},
            "e2e_no_mask": {"value": e2e_nomask_value, "unit": UNIT, "h2d_bytes_per_step": h2d_nm, "d2h_bytes_per_step": d2h_fr,
                            "api": "same call with masks=None: what crosses PCIe when the mask is produced on the device "
                                   "(flope_yolo_mask) or not used"},
            "e2e_module_f32": {"value": e2e_mod_value, "unit": UNIT, "h2d_bytes_per_step": h2d_mod, "d2h_bytes_per_step": d2h,
                               "api": "flope_b200.posenet.PoseResNet.__call__ + flope_pose_head on pinned host float32 crops "
                                      "(the reference's tensor contract: 602 KB per crop over PCIe)"},
            "h2d_ceiling": h2d_ceiling,
            "gpu_launches": launches_per_step * K, "roofline": roofline, "roofline_roi": roofline_roi,
            "roofline_roi_lanczos512": roofline_roi_l, "frames_pipeline": frames_pipeline}
    if sweep is not None:
        line["sweep_1m"] = sweep
    if rank == 0 and world == 1:
        line["latency"] = latency_leg(torch, np, _lib, synth, sd, dev, S)
        if not args.no_cpu:
            line["incumbent_gpu"] = incumbent_gpu(S, B, dev)
            line["cpu_baseline"] = cpu_baseline(S, B)
            line["cpu_legs"] = cpu_legs(S)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def latency_leg(torch, np, _lib, synth, sd, dev, S, iters=1000):
    """BASELINE configs[4] (scripts/live_pose.py:32-41): one 1080p frame, 8 flowers, batch-1 frame latency, host numpy in ->
    host poses out, (a) through the C ABI call flope_infer_frames with pinned staging, (b) through the drop-in
    FastPosePredictor.get_flower_poses(rgb, depth) with an injected detector that returns precomputed boxes + mask."""
    from flope_b200.posenet import PoseResNet
    from flope_b200.predictor import FastPosePredictor
    frames_np, masks_np, det, b5_np = frame_batch(np, _lib, synth, 1, 8, synth.FRAME_SEED + 5)
    net = PoseResNet(device=str(dev), max_batch=8, crop_hw=S)
    net.load_state_dict(sd)
    e = net.engine
    arena = FrameArena(torch, frames_np, masks_np, b5_np, dev, True)
    n = b5_np.shape[0]
    do = torch.empty((n, 3, 3), dtype=torch.float64, device=dev)
    ho = torch.empty((n, 3, 3), dtype=torch.float64).pin_memory()

    def one_abi():
        arena.upload()
        e.infer_frames(arena.frames, arena.masks, arena.boxes, _lib.INTERP_LINEAR, want_R=False, want_yaw=True, out=do)
        ho.copy_(do, non_blocking=True)
        torch.cuda.synchronize()

    pred = FastPosePredictor(str(dev), detector=lambda rgb: (det[0].astype(np.int16), masks_np[0]), posenet=net, crop_hw=S,
                             interp=_lib.INTERP_LINEAR)
    rgb = frames_np[0]

    def one_pred():
        return pred.get_flower_poses(rgb, None)

    def dev_only():
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        e.infer_frames(arena.frames, arena.masks, arena.boxes, _lib.INTERP_LINEAR, want_R=False, want_yaw=True, out=do)
        eb.record(); torch.cuda.synchronize()
        return ea.elapsed_time(eb)

    out = {"workload": "BASELINE configs[4]: 1 x 1080p frame, %d flowers, bilinear %d, %d iterations each" % (n, S, iters)}
    for name, fn in (("infer_frames_host_to_host", one_abi), ("get_flower_poses", one_pred)):
        for _ in range(20):
            fn()
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        out[name] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[int(len(ts) * 0.99)], "mean_ms": sum(ts) / len(ts)}
    ts = sorted(dev_only() for _ in range(200))
    out["device_only_infer_frames"] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[int(len(ts) * 0.99)]}
    out["h2d_bytes_per_frame"] = arena.nbytes
    net.engine.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--frames", type=int, default=64, help="frames per step of the frames_pipeline / ROI roofline legs (BASELINE configs[2])")
    ap.add_argument("--sweep", type=int, default=1000000, help="crops of the sharded sweep leg (BASELINE configs[3]); 0 = skip")
    ap.add_argument("--no-cpu", action="store_true", help="skip the baseline legs (cpu_baseline, cpu_legs, incumbent_gpu)")
    ap.add_argument("--inflight", type=int, default=1, help="independent steps in flight per GPU for `value` (engines/streams); 1 = one engine with layer1-4 as one launch (fastest since round 1's trunk kernel), 2+ = EnginePool with per-layer launches")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
