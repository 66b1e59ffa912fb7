"""GPU parity: fused ROI crop kernel vs the oracle crop loop (real cv2), bit-exact, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import resize as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(cuda_lib):
    e = cuda_lib.Engine(0, max_batch=16, crop_hw=224)
    yield e
    e.close()


def _frame(rng, H=360, W=480, smooth=False):
    f = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    m = np.zeros((H, W), np.uint8)
    yy, xx = np.ogrid[0:H, 0:W]
    m[((xx - W / 2) / (W / 3)) ** 2 + ((yy - H / 2) / (H / 3)) ** 2 <= 1] = 255
    m[rng.integers(0, H, 400), rng.integers(0, W, 400)] = 255
    return f, m


def _boxes():
    # square, in-frame boxes: tiny, odd, equal to the output size, larger than it, touching every edge
    return np.array([[0, 0, 2, 2], [5, 7, 14, 16], [100, 50, 137, 87], [10, 10, 234, 234], [0, 0, 360, 360],
                     [120, 0, 480, 360], [200, 100, 301, 201], [300, 180, 480, 360], [33, 44, 81, 92],
                     [7, 3, 8, 4]], np.int32)


def test_normalise_arithmetic_exhaustive(cuda_lib):
    lut = cuda_lib.debug_normalise_lut(0).cpu().numpy()
    assert np.array_equal(lut, R.normalise_lut())


@pytest.mark.parametrize("interp,size", [(R.LANCZOS4, 512), (R.LANCZOS4, 224), (R.BILINEAR, 224), (R.BILINEAR, 512)])
@pytest.mark.parametrize("with_mask", [True, False])
def test_roi_crop_bit_exact(cuda_lib, eng, interp, size, with_mask):
    rng = np.random.default_rng(11)
    frame, mask = _frame(rng)
    boxes = _boxes()
    want = R.crop_batch_reference(frame, mask if with_mask else None, boxes, size=size, interp=interp)
    fr = torch.from_numpy(frame).cuda()[None]
    mk = torch.from_numpy(mask).cuda()[None] if with_mask else None
    b5 = torch.from_numpy(np.concatenate([np.zeros((len(boxes), 1), np.int32), boxes], 1)).cuda()
    got = eng.roi_crop(fr, mk, b5, size, interp)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    bad = got != want
    assert not bad.any(), f"{bad.sum()} of {bad.size} values differ; max |d| {np.abs(got - want).max()}"


def test_roi_multi_frame_indexing(cuda_lib, eng):
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (3, 120, 160, 3), dtype=np.uint8)
    boxes5 = np.array([[2, 10, 10, 74, 74], [0, 0, 0, 120, 120], [1, 40, 0, 160, 120], [2, 100, 60, 160, 120]], np.int32)
    got = eng.roi_crop(torch.from_numpy(frames).cuda(), None, torch.from_numpy(boxes5).cuda(), 224, R.BILINEAR)
    torch.cuda.synchronize()
    for i, (f, *bb) in enumerate(boxes5):
        want = R.crop_batch_reference(frames[f], None, [bb], size=224, interp=R.BILINEAR)[0]
        assert np.array_equal(got[i].cpu().numpy(), want)


def test_engine_format_equals_f32_path(cuda_lib):
    """Crops written straight into the stem input give bit-identical PoseNet outputs to crops that take
    the float32 NCHW detour (the bf16 rounding of the same float32 values happens in both)."""
    from oracle import posenet as onet
    net = onet.build(0)
    e = cuda_lib.Engine(0, max_batch=8, crop_hw=224)
    e.load_state_dict(net.state_dict())
    rng = np.random.default_rng(5)
    frame, mask = _frame(rng)
    boxes = _boxes()[2:8]
    fr = torch.from_numpy(frame).cuda()[None]
    mk = torch.from_numpy(mask).cuda()[None]
    b5 = torch.from_numpy(np.concatenate([np.zeros((len(boxes), 1), np.int32), boxes], 1)).cuda()
    crops = e.roi_crop(fr, mk, b5, 224, R.LANCZOS4)
    a = e.posenet_forward(crops).clone()
    e.roi_crop(fr, mk, b5, 224, R.LANCZOS4, out_fmt=cuda_lib.OUT_ENGINE)
    b = e.posenet_forward(None, n=len(boxes))
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    e.close()
