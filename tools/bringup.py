"""GPU bring-up: per-layer comparison of the engine against the CPU oracle (run under gpurun)."""
import argparse
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from flope_b200 import _lib, synth
from oracle import posenet as onet, rotation as orot


def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).abs().max()), float((a - b).norm() / (b.norm() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--batch", type=int, default=6)
    ap.add_argument("--swap", type=int, default=-1)
    ap.add_argument("--bench", type=int, default=0)
    a = ap.parse_args()
    print(torch.cuda.get_device_name(0), flush=True)
    net = onet.build(0)
    x = synth.mixed_crops(a.batch, a.size)
    acts = onet.trunk_activations(net, x)
    eng = _lib.Engine(0, max_batch=max(a.batch, 8), crop_hw=a.size)
    eng.load_state_dict(net.state_dict())
    xd = x.cuda()
    for swap in [0]:
        r9 = eng.posenet_forward(xd)
        torch.cuda.synchronize()
        print(f"--- swap_lbo_sbo={swap}", flush=True)
        for name in ["stem", "maxpool", "layer1.0", "layer1.1", "layer2.0", "layer2.1", "layer3.0", "layer3.1",
                     "layer4.0", "layer4.1"]:
            buf, chw = eng.debug_activation(name, a.batch)
            torch.cuda.synchronize()
            got = buf.cpu().reshape(acts[name].shape)
            mx, rl = rel(got, acts[name])
            print(f"{name:10s} max_abs_err {mx:.4e} rel_l2 {rl:.4e}  ref_absmax {float(acts[name].abs().max()):.3f}", flush=True)
        mx, rl = rel(r9.cpu(), acts["r9"])
        print(f"r9         max_abs_err {mx:.4e} rel_l2 {rl:.4e}")
        R, Ry = eng.pose_head(r9)
        torch.cuda.synchronize()
        Rref = orot.procrustes_to_rotmat(acts["r9"]).numpy()
        print("geodesic deg vs fp32 oracle: mean %.4f max %.4f" % (orot.geodesic_deg(R.cpu().numpy(), Rref).mean(),
                                                                 orot.geodesic_deg(R.cpu().numpy(), Rref).max()))
        Rself = orot.procrustes_to_rotmat(r9.cpu()).numpy()
        print("head-only geodesic (same r9): max %.2e" % orot.geodesic_deg(R.cpu().numpy(), Rself).max())
        yref = orot.nullify_yaw_batch(R.cpu().numpy().astype(np.float64))
        print("yaw max abs diff vs scipy: %.2e" % np.abs(Ry.cpu().numpy() - yref).max(), flush=True)
    if a.bench:
        B = a.bench
        eng2 = _lib.Engine(0, max_batch=B, crop_hw=a.size)
        eng2.load_state_dict(net.state_dict())
        xb = torch.rand((B, 3, a.size, a.size), device="cuda")
        out = torch.empty((B, 9), device="cuda")
        for _ in range(3):
            eng2.posenet_forward(xb, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng2.posenet_forward(xb, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 3.6293e9 if a.size == 224 else 18.952e9
        print(f"B={B} size={a.size}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} crops/s  {B * fl / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
