"""Several steps in flight on one GPU: a pool of engines (own workspace and CUDA graph each), one stream per engine.

The backbone kernels are persistent (one CTA pair per TPC) and every layer ends in a partial wave - at 256 crops
layer3 runs 3.04 waves, layer4 1.73 - so a single in-order stream leaves SMs idle at every layer boundary.  With two
independent steps on two streams the scheduler fills those SMs with the other step's kernels: +6 % crops/s on B200
for 256-crop steps (tools/ab_overlap.py), bit-identical results.  Steps are independent (the path has no exchange
step), so this is the single-GPU analogue of sharding crops across ranks.
"""
import torch

from . import _lib


class EnginePool:
    def __init__(self, device, n_engines=2, max_batch=256, crop_hw=224, state_dict=None, cooperative_chains=False, dynamic_chains=False):
        self.device = torch.device(device)
        idx = self.device.index or 0
        self.engines = [_lib.Engine(idx, max_batch, crop_hw) for _ in range(n_engines)]
        if n_engines > 1:
            # Stage chains (one persistent launch per ResNet stage whose tiles wait for each other) need every CTA of a
            # chain kernel to become resident.  Two chain kernels from two streams could each hold part of the SMs while
            # waiting for their own unscheduled CTAs, so engines that run concurrently either launch one kernel per
            # layer (default) or launch their chains cooperatively (gang-scheduled grids; 1 % faster than per-layer
            # launches, but Nsight Compute cannot profile cooperative cluster launches: "LaunchFailed"), or let the
            # chains claim their work items from an atomic counter (every claimed item then belongs to a resident CTA
            # and only waits on lower items: deadlock-free under partial residency, profilable, but no faster than
            # per-layer launches - 923 vs 911 vs 904 us/step for dynamic / per-layer / cooperative, two engines).
            for e in self.engines:
                if dynamic_chains:
                    e.debug_set("chain_dynamic", 1)
                elif cooperative_chains:
                    e.debug_set("chain_coop", 1)
                else:
                    e.debug_set("chain", 0)
        if state_dict is not None:
            for e in self.engines:
                e.load_state_dict(state_dict)
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream() for _ in range(n_engines)]
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

    def join(self):
        """Make the caller's current stream wait for every submitted step."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def close(self):
        for e in self.engines:
            e.close()
