"""CPU-side checks of the C-ABI library: it builds, loads, exports every declared symbol,
its host function is bit-exact with the reference goldens, and it fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from flope_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    build.build()
    return _lib.lib()


def test_header_symbols_are_exported(L):
    hdr = open(os.path.join(ROOT, "include", "flope_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(flope_[a-z0-9_]+)\s*\(", hdr))
    assert names, "no declarations found"
    assert names == set(_lib.SYMBOLS)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/flope_b200.h but not exported"
    assert L.flope_version() >= 100


def test_squarify_filter_bit_exact_vs_reference_golden(L, golden_dir):
    g = np.load(os.path.join(golden_dir, "boxes.npz"))
    H, W = (int(v) for v in g["frame_hw"])
    sq, keep = _lib.squarify_filter(np.ascontiguousarray(g["boxes"], dtype=np.int32), H, W)
    assert np.array_equal(keep, g["keep"])
    assert np.array_equal(sq, g["squarified"][g["keep"]].astype(np.int32))
    sq0, keep0 = _lib.squarify_filter(np.zeros((0, 4), np.int32), H, W)
    assert sq0.shape == (0, 4) and keep0.shape == (0,)


def test_python_mirror_matches_reference_golden(golden_dir):
    from flope_b200 import mvg
    g = np.load(os.path.join(golden_dir, "boxes.npz"))
    H, W = g["frame_hw"]
    got = np.array([mvg.squarify_bb(b) for b in g["boxes"]])
    assert np.array_equal(got, g["squarified"])
    assert np.array_equal(np.array([mvg.bb_in_frame(s, (H, W, 3)) for s in got]), g["keep"])
    assert np.array_equal(mvg.filter_very_large_bb(g["vlb_in"]), g["vlb_out"])


def test_errors_are_codes_not_crashes(L):
    import ctypes as C
    assert L.flope_engine_create(None, 0, 1, 224) == -1
    assert b"NULL" in L.flope_last_error()
    h = C.c_void_p()
    assert L.flope_engine_create(C.byref(h), 0, 0, 224) == -1
    assert L.flope_engine_create(C.byref(h), 0, 4, 100) == -1       # crop side must be a multiple of 32
    assert L.flope_squarify_filter(None, 3, 10, 10, None, None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(_lib.FlopeError):
        _lib.Engine(0, 4, 224)
    from flope_b200.conversion import procrustes_to_rotmat
    with pytest.raises(_lib.FlopeError):
        procrustes_to_rotmat(torch.zeros(2, 9))
    from flope_b200 import image_manipulation as im
    from flope_b200.pipeline import EnginePool
    with pytest.raises(_lib.FlopeError):
        im.get_depth_value(np.zeros((1, 4), np.int32), np.ones((8, 8), np.float32), np.zeros((8, 8), np.uint8))
    with pytest.raises(_lib.FlopeError):
        im.shrink_mask(np.ones((8, 8), bool), 3)
    with pytest.raises(_lib.FlopeError):
        EnginePool("cuda:0", n_engines=2, max_batch=4, crop_hw=64)
    with pytest.raises(_lib.FlopeError):
        _lib.yolo_mask(torch.zeros((1, 8, 8)), 16, 16)
    with pytest.raises(_lib.FlopeError):
        _lib.depth_values(torch.zeros((8, 8)), torch.zeros((8, 8), dtype=torch.uint8), torch.zeros((0, 4), dtype=torch.int32), 0.1, 2.5)


def test_new_entry_points_reject_bad_arguments(L):
    """The stateless device entry points validate their arguments before touching CUDA (codes, not crashes)."""
    assert L.flope_depth_values(0, None, 0, 1.0, None, 8, 8, None, 0, 0.1, 2.5, 10, None, None, None, None) == -1
    assert L.flope_yolo_mask(0, None, 1, 8, 8, None, None, 16, 16, None, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "flope_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_squarify_filter_property_vs_reference_arithmetic(L):
    """Random boxes far outside the golden set (negative coordinates, degenerate and inverted boxes, frame edges): the C
    ABI host function must agree with the reference's scalar arithmetic (true division + int() truncation, restated in
    flope_b200/mvg.py and pinned by the golden test above) on every one of them."""
    from hypothesis import given, settings, strategies as st
    from flope_b200 import mvg

    coord = st.integers(min_value=-5000, max_value=5000)

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.tuples(coord, coord, coord, coord), min_size=1, max_size=40), st.integers(1, 4000), st.integers(1, 4000))
    def check(boxes, H, W):
        b = np.array(boxes, dtype=np.int32)
        sq, keep = _lib.squarify_filter(b, H, W)
        want = np.array([mvg.squarify_bb([int(v) for v in bb]) for bb in b], dtype=np.int64).reshape(-1, 4)
        want_keep = np.array([mvg.bb_in_frame(s, (H, W, 3)) for s in want], dtype=bool)
        assert np.array_equal(keep, want_keep)
        assert np.array_equal(sq, want[want_keep].astype(np.int32))

    check()


def test_python_constants_match_the_header():
    """The ctypes layer repeats the header's enumerators; a missing or drifting one is an AttributeError (or worse, a
    wrong schedule) in product code such as pipeline.EnginePool."""
    import re
    from flope_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "flope_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(FLOPE_[A-Z0-9_]+)\s+(-?\d+)\b", hdr)}
    for name in ("SCHED_PERSISTENT", "SCHED_PER_LAYER", "SCHED_DYNAMIC", "SCHED_COOPERATIVE"):
        assert getattr(_lib, name) == defs["FLOPE_" + name], name
    for name, key in (("INTERP_LINEAR", "FLOPE_INTERP_LINEAR"), ("INTERP_LANCZOS4", "FLOPE_INTERP_LANCZOS4"),
                      ("OUT_F32_NCHW", "FLOPE_OUT_F32_NCHW"), ("OUT_ENGINE", "FLOPE_OUT_ENGINE")):
        if key in defs:
            assert getattr(_lib, name) == defs[key], name


def test_every_lib_attribute_the_package_uses_exists():
    """flope_b200/*.py only reaches for _lib names that _lib defines (catches a constant used before it was added)."""
    import re
    from flope_b200 import _lib
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flope_b200")
    missing = []
    for fn in sorted(os.listdir(pkg)):
        if fn.endswith(".py") and fn != "_lib.py":
            for name in set(re.findall(r"\b_lib\.([A-Za-z_][A-Za-z0-9_]*)", open(os.path.join(pkg, fn)).read())):
                if not hasattr(_lib, name):
                    missing.append((fn, name))
    assert not missing, missing
