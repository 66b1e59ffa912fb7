"""The C ABI from plain C: tests/c_abi/c_abi_smoke.c is compiled with gcc against include/flope_b200.h and
libflope_b200.so (no Python, torch or C++ on its side), fed the reference state_dict and a crop batch through files, and
must reproduce the ctypes path bit for bit."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

from flope_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_home():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))


def test_plain_c_client_matches_ctypes_path(cuda_lib, tmp_path):
    cuda = _cuda_home()
    exe = tmp_path / "c_abi_smoke"
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.c"), "-o", str(exe),
           "-L", os.path.join(ROOT, "flope_b200"), "-lflope_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "flope_b200"), "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sd = synth.random_state_dict(synth.WEIGHT_SEED)
    with open(tmp_path / "w.bin", "wb") as f:
        items = [(k, v.numpy().astype(np.float32)) for k, v in sd.items() if not k.endswith("num_batches_tracked")]
        f.write(struct.pack("<i", len(items)))
        for k, a in items:
            kb = k.encode()
            f.write(struct.pack("<i", len(kb))); f.write(kb)
            f.write(struct.pack("<i", a.ndim)); f.write(struct.pack(f"<{a.ndim}q", *a.shape))
            f.write(np.ascontiguousarray(a).tobytes())
    n, S = 5, 96
    x = synth.mixed_crops(n, S, seed=3)
    with open(tmp_path / "x.bin", "wb") as f:
        f.write(struct.pack("<ii", n, S)); f.write(x.numpy().tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "w.bin"), str(tmp_path / "x.bin"), str(tmp_path / "o.bin")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(tmp_path / "o.bin", "rb").read()
    r9_c = np.frombuffer(raw[: n * 9 * 4], np.float32).reshape(n, 9)
    yaw_c = np.frombuffer(raw[n * 9 * 4:], np.float64).reshape(n, 3, 3)
    e = cuda_lib.Engine(0, max_batch=n, crop_hw=S)
    try:
        e.load_state_dict(sd)
        r9 = e.posenet_forward(x.cuda())
        _, Ry = e.pose_head(r9)
        torch.cuda.synchronize()
        assert np.array_equal(r9.cpu().numpy(), r9_c)
        assert np.array_equal(Ry.cpu().numpy().reshape(n, 3, 3), yaw_c)
    finally:
        e.close()
