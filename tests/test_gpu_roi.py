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


def _b5(boxes, frame=0):
    return torch.from_numpy(np.concatenate([np.full((len(boxes), 1), frame, np.int32), np.asarray(boxes, np.int32)], 1)).cuda()


# (stream, item rows, stage KB, ring stages, CTAs per SM, dynamic claiming): the streaming kernels under several launch
# geometries - the default (items sized to the launch), the large-launch item size, tiny stages (one or two rows per
# stage), short and tall items, a single CTA per SM - and the generic kernel
_ROI_MODES = [(1, 0, 0, 0, 0, 1), (1, 56, 14, 3, 0, 1), (1, 7, 1, 2, 1, 0), (1, 16, 4, 4, 0, 1), (1, 128, 16, 3, 2, 1), (0, 0, 0, 0, 0, 1)]


def _set_roi_mode(e, mode, lanczos):
    stream, rows, kb, stages, per_sm, dyn = mode
    e.debug_set("roi_stream", stream)
    if rows:
        e.debug_set("roi_item_rows8" if lanczos else "roi_item_rows", rows)
        e.debug_set("roi_stage_kb", kb); e.debug_set("roi_stages", stages)
    e.debug_set("roi_ctas_per_sm", per_sm); e.debug_set("roi_dynamic", dyn)


@pytest.mark.parametrize("interp,size", [(R.LANCZOS4, 512), (R.BILINEAR, 224), (R.LANCZOS4, 224), (R.BILINEAR, 512)])
def test_streaming_kernels_match_generic_under_every_geometry(cuda_lib, interp, size):
    """The streaming kernels (TMA row ring, DP2A taps, FMA-pipe vertical pass) under different item / stage / grid
    geometries and the generic one-thread-per-column kernel implement the same integer arithmetic: bitwise equal.
    Repeated launches also exercise the self-resetting work counter."""
    rng = np.random.default_rng(21)
    frame, mask = _frame(rng)
    fr, mk, b5 = torch.from_numpy(frame).cuda()[None], torch.from_numpy(mask).cuda()[None], _b5(_boxes())
    e = cuda_lib.Engine(0, max_batch=16, crop_hw=size)
    outs = []
    for mode in _ROI_MODES:
        _set_roi_mode(e, mode, interp == R.LANCZOS4)
        for m in (mk, None):
            outs.append((m is None, e.roi_crop(fr, m, b5, size, interp).clone()))
    torch.cuda.synchronize()
    for nomask, o in outs[2:]:
        assert torch.equal(o, outs[1 if nomask else 0][1])
    e.close()


def test_engine_format_streaming_equals_generic(cuda_lib):
    """bf16 stem-input crops of the streaming kernels vs the generic kernel: identical PoseNet outputs."""
    from oracle import posenet as onet
    net = onet.build(0)
    e = cuda_lib.Engine(0, max_batch=16, crop_hw=224)
    e.load_state_dict(net.state_dict())
    rng = np.random.default_rng(22)
    frame, mask = _frame(rng)
    fr, mk, b5 = torch.from_numpy(frame).cuda()[None], torch.from_numpy(mask).cuda()[None], _b5(_boxes())
    for interp in (R.BILINEAR, R.LANCZOS4):
        outs = []
        for mode in _ROI_MODES:
            _set_roi_mode(e, mode, interp == R.LANCZOS4)
            e.roi_crop(fr, mk, b5, 224, interp, out_fmt=cuda_lib.OUT_ENGINE)
            outs.append(e.posenet_forward(None, n=len(b5)).clone())
        torch.cuda.synchronize()
        for o in outs[1:]:
            assert torch.equal(outs[0], o), interp
    e.close()


@pytest.mark.parametrize("interp", [R.BILINEAR, R.LANCZOS4])
def test_roi_frame_width_not_multiple_of_16_takes_generic_path(cuda_lib, eng, interp):
    rng = np.random.default_rng(23)
    frame, mask = _frame(rng, H=241, W=322)
    boxes = np.array([[0, 0, 241, 241], [81, 0, 322, 241], [13, 17, 150, 154], [300, 200, 322, 222]], np.int32)
    want = R.crop_batch_reference(frame, mask, boxes, size=224, interp=interp)
    got = eng.roi_crop(torch.from_numpy(frame).cuda()[None], torch.from_numpy(mask).cuda()[None], _b5(boxes), 224, interp)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("interp,size", [(R.BILINEAR, 224), (R.LANCZOS4, 512)])
def test_roi_1080p_last_frames_of_64(cuda_lib, interp, size):
    """BASELINE configs[2] geometry: 64 x 1080p frames (398 MB: byte offsets beyond 2^31), the configs[2] box
    distribution; boxes taken from the last frames, whose-frame boxes and boxes touching the last row included."""
    from flope_b200 import synth
    frames, masks, det = synth.frames_and_boxes(3, 32, seed=13, with_mask=True)       # content of frames 61..63
    big = torch.zeros((64, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
    bigm = torch.zeros((64, 1080, 1920), dtype=torch.uint8, device="cuda")
    big[61:] = torch.from_numpy(frames).cuda()
    bigm[61:] = torch.from_numpy(masks).cuda()
    rows = []
    for f in range(3):
        sq, _ = cuda_lib.squarify_filter(np.ascontiguousarray(det[f]), 1080, 1920)
        sel = sq[:10 if size == 224 else 3]
        extra = np.array([[840, 0, 1920, 1080], [0, 1000, 80, 1080], [1900, 1060, 1920, 1080]], np.int32)
        for bb in np.concatenate([sel, extra]):
            rows.append([61 + f, *bb])
    rows = np.array(rows, np.int32)
    e = cuda_lib.Engine(0, max_batch=8, crop_hw=size)
    got = e.roi_crop(big, bigm, torch.from_numpy(rows).cuda(), size, interp)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    for i, (f, *bb) in enumerate(rows):
        want = R.crop_batch_reference(frames[f - 61], masks[f - 61], [bb], size=size, interp=interp)[0]
        assert np.array_equal(got[i], want), (i, f, bb)
    e.close()


@pytest.mark.parametrize("interp,size,n_boxes", [(R.BILINEAR, 224, 160), (R.LANCZOS4, 224, 96), (R.LANCZOS4, 512, 40), (R.BILINEAR, 512, 40),
                                                  (R.BILINEAR, 96, 64), (R.LANCZOS4, 160, 48)])
def test_roi_random_boxes_and_alignments_bit_exact(cuda_lib, interp, size, n_boxes):
    """Random square boxes of every size and position (all 16-byte phases of the row segments, boxes touching every
    edge, sides from 1 px to the whole frame height, up- and down-scaling by more than 2x) in two frames of different
    widths, the second one addressed through a non-zero frame index: bit-exact against cv2, with and without mask."""
    rng = np.random.default_rng(1000 + size + interp)
    for H, W in ((300, 480), (200, 272)):
        frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
        masks = np.zeros((2, H, W), np.uint8)
        yy, xx = np.ogrid[0:H, 0:W]
        masks[0][((xx - W / 2) / (W / 2.5)) ** 2 + ((yy - H / 2) / (H / 2.2)) ** 2 <= 1] = 255
        masks[1] = rng.integers(0, 2, (H, W), dtype=np.uint8) * 255                       # worst case: noise mask, all partial
        side = np.concatenate([rng.integers(1, 9, n_boxes // 4), rng.integers(9, H + 1, n_boxes - n_boxes // 4)])
        side[-1] = H
        x0 = (rng.random(n_boxes) * (W - side + 1)).astype(np.int64)
        y0 = (rng.random(n_boxes) * (H - side + 1)).astype(np.int64)
        x0[::7] = 0; y0[::5] = 0
        x0[3::11] = (W - side)[3::11]; y0[4::13] = (H - side)[4::13]
        f = rng.integers(0, 2, n_boxes)
        b5 = np.stack([f, x0, y0, x0 + side, y0 + side], 1).astype(np.int32)
        e = cuda_lib.Engine(0, max_batch=8, crop_hw=size)
        try:
            fr, mk = torch.from_numpy(frames).cuda(), torch.from_numpy(masks).cuda()
            for m_dev, with_mask in ((mk, True), (None, False)):
                got = e.roi_crop(fr, m_dev, b5, size, interp)
                torch.cuda.synchronize()
                got = got.cpu().numpy()
                for i, (fi, *bb) in enumerate(b5):
                    want = R.crop_batch_reference(frames[fi], masks[fi] if with_mask else None, [bb], size=size, interp=interp)[0]
                    assert np.array_equal(got[i], want), (H, W, with_mask, i, fi, bb)
        finally:
            e.close()


@pytest.mark.parametrize("interp", [R.BILINEAR, R.LANCZOS4])
def test_engine_format_masked_normalise_exhaustive(cuda_lib, interp):
    """The bf16 stem-input crops use one branch-free expression for every (value, mask) pair.  A 256 x 256 box resized to
    256 x 256 is the identity for both interpolation modes (coefficients 2048 / 0), so with img[y][x] = x and
    mask[y][x] = y the crop holds all 65 536 pairs: every stored bf16 must equal bf16_rn of the reference's float32
    expression (oracle.resize.normalise_lut)."""
    S = 256
    frame = np.broadcast_to(np.arange(S, dtype=np.uint8)[None, :, None], (S, S, 3)).copy()
    mask = np.broadcast_to(np.arange(S, dtype=np.uint8)[:, None], (S, S)).copy()
    e = cuda_lib.Engine(0, max_batch=1, crop_hw=S)
    try:
        b5 = np.array([[0, 0, 0, S, S]], np.int32)
        f32 = e.roi_crop(torch.from_numpy(frame).cuda()[None], torch.from_numpy(mask).cuda()[None], b5, S, interp)
        e.roi_crop(torch.from_numpy(frame).cuda()[None], torch.from_numpy(mask).cuda()[None], b5, S, interp, out_fmt=cuda_lib.OUT_ENGINE)
        buf, chw = e.debug_activation("x0", 1)
        torch.cuda.synchronize()
        lut = R.normalise_lut()                                                # [img][mask] float32
        want = lut[np.arange(S)[None, :], np.arange(S)[:, None]]               # [y][x] = lut[img = x][mask = y]
        assert np.array_equal(f32.cpu().numpy()[0, 0], want)                   # the fp32 path is exact as ever
        x0 = buf.cpu().numpy().reshape(16, S // 2, S // 2)                     # channel k = (y & 1) * 8 + (x & 1) * 4 + c
        want_bf16 = torch.from_numpy(want).to(torch.bfloat16).to(torch.float32).numpy()
        for c in range(3):
            got = np.empty((S, S), np.float32)
            for by in range(2):
                for bx in range(2):
                    got[by::2, bx::2] = x0[by * 8 + bx * 4 + c]
            assert np.array_equal(got, want_bf16), (interp, c, int((got != want_bf16).sum()))
        assert not x0[3::4].any()                                              # the pad lane of every pixel stays zero
    finally:
        e.close()


@pytest.mark.parametrize("interp,size", [(R.BILINEAR, 224), (R.LANCZOS4, 512)])
@pytest.mark.parametrize("stream", [1, 0])
def test_empty_boxes_give_zero_crops_and_do_not_disturb_the_others(cuda_lib, interp, size, stream):
    """Boxes that are already on the device cannot be range-checked by the host.  The reference's cv2.resize raises on an
    empty slice; the kernels define such a crop as zeros (its output slot must never keep a previous batch's pixels) and
    the neighbouring crops stay bit-exact - in the streaming kernels and in the generic one."""
    rng = np.random.default_rng(77)
    frame, mask = _frame(rng)
    boxes = np.array([[100, 50, 137, 87], [60, 60, 60, 90], [10, 10, 234, 234], [70, 80, 40, 50], [200, 100, 301, 201]], np.int32)
    b5 = torch.from_numpy(np.concatenate([np.zeros((5, 1), np.int32), boxes], 1)).cuda()       # device boxes: no host check
    e = cuda_lib.Engine(0, max_batch=8, crop_hw=size)
    try:
        e.debug_set("roi_stream", stream)
        out = torch.full((5, 3, size, size), 7.0, device="cuda")
        fr, mk = torch.from_numpy(frame).cuda()[None], torch.from_numpy(mask).cuda()[None]
        e.roi_crop(fr, mk, b5, size, interp, out=out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert not got[1].any() and not got[3].any()
        want = R.crop_batch_reference(frame, mask, boxes[[0, 2, 4]], size=size, interp=interp)
        assert np.array_equal(got[[0, 2, 4]], want)
        if size == 224:
            # engine format: run a valid batch first, then the batch with empty boxes - their stem-input slots must be zero
            valid = b5.clone(); valid[1, 1:] = torch.tensor([0, 0, 50, 50]); valid[3, 1:] = torch.tensor([5, 5, 90, 90])
            e.roi_crop(fr, mk, valid, size, interp, out_fmt=cuda_lib.OUT_ENGINE)
            e.roi_crop(fr, mk, b5, size, interp, out_fmt=cuda_lib.OUT_ENGINE)
            buf, chw = e.debug_activation("x0", 5)
            torch.cuda.synchronize()
            x0 = buf.cpu().numpy().reshape(5, -1)
            assert not x0[1].any() and not x0[3].any() and x0[0].any() and x0[2].any()
    finally:
        e.close()
