"""CPU check of the staged ROI kernels' thread programs (flope_b200/csrc/roi_crop.cuh) against real cv2.

tests/emu/roi_emu.cu compiles the kernels' per-thread functions for the host and walks the grid
sequentially (memcpy instead of the TMA bulk copies), so the window/permute/DP2A/ring index arithmetic
is verified bit for bit without a GPU.  The GPU tests (tests/test_gpu_roi.py) check the kernels themselves.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import resize as R

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "roi_emu.cu")
LIB = os.path.join(HERE, "emu", "roi_emu.so")
DEP = os.path.join(os.path.dirname(HERE), "flope_b200", "csrc", "roi_crop.cuh")


@pytest.fixture(scope="module")
def emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    if not os.path.exists(LIB) or max(os.path.getmtime(SRC), os.path.getmtime(DEP)) > os.path.getmtime(LIB):
        subprocess.run([nvcc, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-shared", "-o", LIB, SRC], check=True, capture_output=True)
    L = C.CDLL(LIB)
    L.roi_emu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                          C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return L


def make_geom(S, max_batch):
    """Geometry of the engine's stem input (csrc/engine.cu make_geom(16, S/2, S/2, pad 2))."""
    H = W = S // 2
    Hp, Wp = H + 2, W + 2
    base = (8 * Wp + 8 + 7) // 8 * 8
    npos = max_batch * Hp * Wp
    plane = base + (npos + 1023) // 1024 * 1024 + 1024 + (4 * Wp + 8 + 7) // 8 * 8
    return np.array([16, H, W, Hp, Wp, base], np.int32), plane


def bf16_bits(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def run_emu(L, frames, masks, boxes5, S, interp, fmt, strip, cols_cta, n_sub, data_bytes, block, lut=1):
    n_frames, H, W, _ = frames.shape
    n = len(boxes5)
    geom, plane = make_geom(S, n)
    if fmt == 0:
        out = np.full((n, 3, S, S), np.nan, np.float32)
    else:
        out = np.zeros((2 * plane * 8,), np.uint16)
    mc = C.c_int(0)
    rc = L.roi_emu(frames.ctypes.data, n_frames, H, W, masks.ctypes.data if masks is not None else None,
                   boxes5.ctypes.data, n, S, interp, fmt, out.ctypes.data, geom.ctypes.data, plane, strip, cols_cta, n_sub,
                   data_bytes, block, lut, C.byref(mc))
    assert rc == 0, rc
    if fmt == 1:    # unpack the s2d blocked-pixel layout: plane y&1, position (y>>1, x>>1), lane (x&1)*4 + c
        _, _, _, Hp, Wp, base = geom
        planes = out.reshape(2, plane, 8)
        got = np.zeros((n, 3, S, S), np.uint16)
        yy, xx = np.mgrid[0:S, 0:S]
        for i in range(n):
            pos = base + (i * Hp + (yy >> 1)) * Wp + (xx >> 1)
            for c in range(3):
                got[i, c] = planes[yy & 1, pos, (xx & 1) * 4 + c]
            assert (planes[yy & 1, pos, (xx & 1) * 4 + 3] == 0).all()
        out = got
    return out, mc.value


def _scene(rng, H, W, n_frames=1):
    frames = rng.integers(0, 256, (n_frames, H, W, 3), dtype=np.uint8)
    masks = np.zeros((n_frames, H, W), np.uint8)
    yy, xx = np.ogrid[0:H, 0:W]
    masks[:, ((xx - W / 2) / (W / 3)) ** 2 + ((yy - H / 2) / (H / 3)) ** 2 <= 1] = 255
    for f in range(n_frames):
        masks[f, rng.integers(0, H, 300), rng.integers(0, W, 300)] = rng.integers(0, 256, 300)
    return frames, masks


BOXES = np.array([[0, 0, 2, 2], [5, 7, 14, 16], [100, 50, 137, 87], [10, 10, 234, 234], [0, 0, 360, 360],
                  [120, 0, 480, 360], [200, 100, 301, 201], [300, 180, 480, 360], [33, 44, 81, 92], [7, 3, 8, 4],
                  [1, 2, 300, 301], [131, 17, 330, 216]], np.int32)


def _cfg(S, interp, strip=None, data_kb=36):
    """The launch configuration csrc/engine.cu run_roi picks."""
    cols = S if S <= 256 else (S // 2 + 1) // 2 * 2
    if interp == R.LANCZOS4:
        return dict(strip=strip or 128, cols_cta=cols, n_sub=1, block=(cols + 31) // 32 * 32, data_bytes=data_kb * 1024)
    pairs = cols // 2
    n_sub = max(1, min(2, 256 // pairs))
    return dict(strip=strip or 32, cols_cta=cols, n_sub=n_sub, block=(pairs * n_sub + 31) // 32 * 32, data_bytes=data_kb * 1024)


@pytest.mark.parametrize("interp,S", [(R.BILINEAR, 224), (R.BILINEAR, 512), (R.LANCZOS4, 224), (R.LANCZOS4, 512),
                                      (R.BILINEAR, 32), (R.LANCZOS4, 64)])
@pytest.mark.parametrize("with_mask", [True, False])
@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("lut", [1, 0])
def test_emu_bit_exact(emu, interp, S, with_mask, fmt, lut):
    rng = np.random.default_rng(11)
    frames, masks = _scene(rng, 360, 480)
    b5 = np.concatenate([np.zeros((len(BOXES), 1), np.int32), BOXES], 1)
    want = R.crop_batch_reference(frames[0], masks[0] if with_mask else None, BOXES, size=S, interp=interp)
    got, _ = run_emu(emu, frames, masks if with_mask else None, b5, S, interp, fmt, lut=lut, **_cfg(S, interp))
    if fmt == 1:
        want = bf16_bits(want)
    bad = got != want
    assert not bad.any(), f"{bad.sum()} of {bad.size} differ; first at {np.argwhere(bad)[:5].tolist()}"


@pytest.mark.parametrize("interp", [R.BILINEAR, R.LANCZOS4])
def test_emu_chunked_staging(emu, interp):
    """A staging area that only holds a few source rows forces several chunks per strip (whole-frame boxes)."""
    rng = np.random.default_rng(5)
    frames, masks = _scene(rng, 240, 320, n_frames=2)
    boxes5 = np.array([[1, 0, 0, 240, 240], [0, 80, 0, 320, 240], [1, 3, 5, 200, 202], [0, 10, 10, 40, 40]], np.int32)
    cfg = _cfg(224, interp, data_kb=16 if interp == R.LANCZOS4 else 8)
    got, max_chunks = run_emu(emu, frames, masks, boxes5, 224, interp, 0, **cfg)
    assert max_chunks > 1
    for i, (f, *bb) in enumerate(boxes5):
        want = R.crop_batch_reference(frames[f], masks[f], [bb], size=224, interp=interp)[0]
        assert np.array_equal(got[i], want), i


def test_emu_frame_edges_use_byte_copies(emu):
    """Rows whose 16-byte aligned superset would leave the frame buffer are staged byte by byte."""
    rng = np.random.default_rng(9)
    frames, masks = _scene(rng, 64, 68)            # W*3 = 204: rows are not 16-byte aligned
    boxes5 = np.array([[0, 0, 0, 64, 64], [0, 4, 0, 68, 64], [0, 1, 1, 63, 63]], np.int32)
    for interp in (R.BILINEAR, R.LANCZOS4):
        got, _ = run_emu(emu, frames, masks, boxes5, 96, interp, 0, **_cfg(96, interp))
        for i, (f, *bb) in enumerate(boxes5):
            want = R.crop_batch_reference(frames[f], masks[f], [bb], size=96, interp=interp)[0]
            assert np.array_equal(got[i], want), (interp, i)
