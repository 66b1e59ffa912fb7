"""Several steps in flight on one GPU: a pool of engines (own workspace and CUDA graph each), one stream per engine.

Two modes.  serial_backbones=True (what bench.py uses): a step is split into a STAGING part (ROI crop or float32
ingest into the engine's stem input - small CTAs that fit on an SM beside a persistent conv CTA) and a RUN part
(backbone + head).  The run parts of consecutive steps are chained by events, so two backbones never compete for the
SMs and every engine keeps its fastest schedule (layer1..layer4 as one launch); only the staging of step i+1 overlaps
step i.  Measured at 256 crops: ROI staging 905 -> 885 us per step, float32 ingest 885 -> 880 (round-1 A/B runs, profiles/r1_notes.md) - the
staging CTAs mostly run in the gaps between conv kernels rather than beside them, also with 128-thread blocks, a
high-priority run stream and the conv kernels' shared-memory carve-out.  serial_backbones=False: fully independent steps, engines fall back to one launch per layer (see below).


The backbone kernels are persistent (one CTA pair per TPC) and every layer ends in a partial wave - at 256 crops
layer3 runs 3.04 waves, layer4 1.73 - so a single in-order stream leaves SMs idle at every layer boundary.  With two
independent steps on two streams the scheduler fills those SMs with the other step's kernels: +6 % crops/s on B200
for 256-crop steps (tools/ab_overlap.py), bit-identical results.  Steps are independent (the path has no exchange
step), so this is the single-GPU analogue of sharding crops across ranks.
"""
import torch

from . import _lib


class EnginePool:
    def __init__(self, device, n_engines=2, max_batch=256, crop_hw=224, state_dict=None, cooperative_chains=False, dynamic_chains=False,
                 serial_backbones=False):
        self.device = torch.device(device)
        idx = self.device.index or 0
        self.engines = [_lib.Engine(idx, max_batch, crop_hw) for _ in range(n_engines)]
        self.serial_backbones = serial_backbones
        self._last_run = None
        if n_engines > 1 and not serial_backbones:
            # The default schedule (layer1..layer4 as one persistent launch whose tiles wait for each other) needs its
            # whole grid resident; the library therefore runs such launches one at a time per device (engine.cu,
            # ResidencyGate), which is safe but gives no backbone overlap.  Engines that should overlap pick a schedule
            # that waits for nothing: one launch per layer (default here), cooperative stage chains (gang-scheduled; 1 %
            # faster, but Nsight Compute cannot profile cooperative cluster launches) or stage chains that claim their
            # work items from an atomic counter (deadlock-free under partial residency, profilable, no faster than
            # per-layer launches - 923 vs 911 vs 904 us/step for dynamic / per-layer / cooperative, two engines).
            sched = _lib.SCHED_DYNAMIC if dynamic_chains else _lib.SCHED_COOPERATIVE if cooperative_chains else _lib.SCHED_PER_LAYER
            for e in self.engines:
                e.set_schedule(sched)
        if state_dict is not None:
            for e in self.engines:
                e.load_state_dict(state_dict)
        with torch.cuda.device(self.device):
            # serial_backbones: the run parts go to high-priority streams, so that when a staging kernel and a conv
            # kernel are both ready the block scheduler places the conv CTAs (one per SM, all of its shared memory)
            # first and fits staging CTAs into what is left, instead of filling the SMs with staging CTAs
            self.streams = [torch.cuda.Stream(priority=-1 if serial_backbones else 0) for _ in range(n_engines)]
            self.stage_streams = [torch.cuda.Stream() for _ in range(n_engines)] if serial_backbones else []
        self._run_done = [None] * n_engines
        self._next = 0

    def __len__(self):
        return len(self.engines)

    def submit(self, fn):
        """Run fn(engine, slot) on the next engine's stream, ordered after everything already queued on the caller's
        current stream (its inputs).  Returns the slot index (0 .. n_engines-1) the step ran on."""
        k = self._next
        self._next = (k + 1) % len(self.engines)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.streams[k].wait_event(ready)
        with torch.cuda.stream(self.streams[k]):
            fn(self.engines[k], k)
        return k

    def submit_staged(self, stage_fn, run_fn):
        """serial_backbones mode: stage_fn(engine, slot) may overlap the previous step's run part; run_fn(engine, slot)
        starts after it.  Both are ordered after the caller's current stream.  Returns the slot."""
        k = self._next
        self._next = (k + 1) % len(self.engines)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        ss, rs = self.stage_streams[k], self.streams[k]
        ss.wait_event(ready)
        if self._run_done[k] is not None:
            ss.wait_event(self._run_done[k])            # this engine's previous step has consumed its stem input
        with torch.cuda.stream(ss):
            stage_fn(self.engines[k], k)
            staged = torch.cuda.Event()
            staged.record(ss)
        rs.wait_event(staged)
        if self._last_run is not None:
            rs.wait_event(self._last_run)               # one backbone at a time on the device
        with torch.cuda.stream(rs):
            run_fn(self.engines[k], k)
            done = torch.cuda.Event()
            done.record(rs)
        self._last_run = self._run_done[k] = done
        return k

    def join(self):
        """Make the caller's current stream wait for every submitted step."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def close(self):
        for e in self.engines:
            e.close()
