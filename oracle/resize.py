"""Oracle: the crop-batch step of the FloPE pose path (test infrastructure only).

* ``crop_batch_reference`` restates the inline loop the reference repeats four
  times (sunflower/predictor/pose_predictor.py:138-153,
  fast_pose_predictor.py:108-123, scripts/test_posenet.py:124-140,
  scripts/generate_metrics_utils.py:17-35) and calls the real ``cv2.resize``.
* ``lanczos4_u8`` / ``linear_u8`` restate the arithmetic cv2 performs for uint8
  images (third-party: opencv-python, reference pin 4.10.0.84, this image
  4.13.0) so the CUDA kernel has an integer specification to match and so the
  restatement itself can be pinned against cv2 in tests/test_oracle_resize.py.
"""
import numpy as np

try:  # cv2 is part of the image on both the authoring container and the GPU box
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

LANCZOS4 = 1
BILINEAR = 0
COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


# --------------------------------------------------------------------------
# cv2-backed reference loop
# --------------------------------------------------------------------------
def crop_u8_reference(frame, mask, boxes, size=512, interp=LANCZOS4):
    """uint8 halves of the reference loop: returns (img (B,S,S,C) u8, mask (B,S,S) u8)."""
    flag = cv2.INTER_LANCZOS4 if interp == LANCZOS4 else cv2.INTER_LINEAR
    imgs, masks = [], []
    for xmin, ymin, xmax, ymax in boxes:
        img_crop = frame[ymin:ymax, xmin:xmax]
        imgs.append(cv2.resize(img_crop, (size, size), interpolation=flag))
        if mask is not None:
            masks.append(cv2.resize(mask[ymin:ymax, xmin:xmax], (size, size), interpolation=flag))
    imgs = np.stack(imgs) if imgs else np.zeros((0, size, size, frame.shape[2]), np.uint8)
    masks = (np.stack(masks) if masks else np.zeros((0, size, size), np.uint8)) if mask is not None else None
    return imgs, masks


def crop_batch_reference(frame, mask, boxes, size=512, interp=LANCZOS4):
    """pose_predictor.py:138-153 - (B,3,S,S) float32 in [0,1], background removed.

    ``mask=None`` is the benchmark-mode extension (equivalent to an all-255 mask).
    """
    imgs, masks = crop_u8_reference(frame, mask, boxes, size, interp)
    if mask is not None:
        batch = [im * (mk.reshape(size, size, 1) / 255.0) for im, mk in zip(imgs, masks)]
    else:
        batch = [im * (np.full((size, size, 1), 255, np.uint8) / 255.0) for im in imgs]
    batch = np.array(batch).reshape(-1, size, size, frame.shape[2]) / 255.0
    return np.ascontiguousarray(batch.astype(np.float32).transpose(0, 3, 1, 2))


def normalise_lut():
    """float32((double(i) * (double(m)/255.0)) / 255.0) for every (m, i) pair - (256,256) f32.

    This is the whole value range of pose_predictor.py:148,151-152; the CUDA
    kernel's mask/normalise arithmetic is checked exhaustively against it.
    """
    i = np.arange(256, dtype=np.float64)[None, :]
    m = np.arange(256, dtype=np.float64)[:, None]
    return ((i * (m / 255.0)) / 255.0).astype(np.float32)


# --------------------------------------------------------------------------
# integer restatement of cv2's uint8 resize
# --------------------------------------------------------------------------
def _rint_sat16(x):
    return np.clip(np.rint(x), -32768, 32767).astype(np.int32)


def lanczos4_taps(src, dst):
    """Per-axis tap table of cv2 INTER_LANCZOS4 for uint8: (ofs (dst,8) int32, coef (dst,8) int32).

    Follows imgproc/resize.cpp (interpolateLanczos4 + the fixed-point conversion in
    cv::resize): float32 coordinate, double sin/cos of the base phase, float32
    weights normalised in float32, rounded (half-even) to 11-bit fixed point.
    Taps are clamped to the crop, i.e. replicate border of the *crop*.
    """
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)

    s45 = 0.70710678118654752440
    cs = np.array([[1, 0], [-s45, -s45], [0, 1], [s45, -s45], [-1, 0], [s45, s45], [0, -1], [-s45, s45]],
                  dtype=np.float64)
    coef = np.zeros((dst, 8), np.float32)
    f3 = np.float32(3.0)
    for k in range(dst):
        x = fx[k]
        xb = np.float32(x + f3)                                   # (x+3) evaluated in float
        y0 = -np.float64(xb) * np.pi * 0.25
        s0, c0 = np.sin(y0), np.cos(y0)
        ssum = np.float32(0)
        for i in range(8):
            t = np.float32(xb - np.float32(i))                    # float y0_ = (x+3-i)
            if abs(t) >= np.float32(1e-6):
                y = -np.float64(t) * np.pi * 0.25
                v = np.float32((cs[i, 0] * s0 + cs[i, 1] * c0) / (y * y))
            else:
                v = np.float32(1e30)                              # x ~ 0 or ~ 1: a single unit tap
            coef[k, i] = v
            ssum = np.float32(ssum + v)
        inv = np.float32(np.float32(1.0) / ssum)
        coef[k] = (coef[k] * inv).astype(np.float32)
    icoef = _rint_sat16(coef.astype(np.float32) * np.float32(COEF_ONE))
    ofs = np.clip(sx[:, None] - 3 + np.arange(8)[None, :], 0, src - 1).astype(np.int32)
    return ofs, icoef


def lanczos4_u8(src_img, dsize):
    """cv2.resize(src, (dsize,dsize), INTER_LANCZOS4) for uint8, integer-exact."""
    squeeze = src_img.ndim == 2
    a = src_img[..., None] if squeeze else src_img
    sh, sw = a.shape[:2]
    xofs, xc = lanczos4_taps(sw, dsize)
    yofs, yc = lanczos4_taps(sh, dsize)
    a = a.astype(np.int64)
    hbuf = np.zeros((sh, dsize, a.shape[2]), np.int64)
    for j in range(8):
        hbuf += a[:, xofs[:, j], :] * xc[None, :, j, None]
    out = np.zeros((dsize, dsize, a.shape[2]), np.int64)
    for j in range(8):
        out += hbuf[yofs[:, j], :, :] * yc[:, j, None, None]
    out = np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out


def linear_taps(src, dst, vertical=False):
    """Per-axis tap table of cv2 INTER_LINEAR for uint8: (ofs (dst,2) int32, coef (dst,2) int32).

    cv::resize treats the two axes differently at the border: horizontally a
    coordinate left of pixel 0 / right of the last pixel gets weight (1,0) on the
    clamped pixel; vertically the fractional weights are kept and only the two
    row indices are clamped (resizeGeneric_Invoker), which matters for rounding.
    """
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = sx < 0
        fx[lo] = 0
        sx[lo] = 0
        hi = sx >= src - 1
        fx[hi] = 0
        sx[hi] = src - 1
    c = np.stack([np.float32(1.0) - fx, fx], axis=1).astype(np.float32)
    ofs = np.clip(np.stack([sx, sx + 1], axis=1), 0, src - 1).astype(np.int32)
    return ofs, _rint_sat16(c * np.float32(COEF_ONE))


def linear_u8(src_img, dsize):
    """cv2.resize(src, (dsize,dsize), INTER_LINEAR) for uint8, integer-exact."""
    squeeze = src_img.ndim == 2
    a = src_img[..., None] if squeeze else src_img
    sh, sw = a.shape[:2]
    xo, xc = linear_taps(sw, dsize)
    yo, yc = linear_taps(sh, dsize, vertical=True)
    a = a.astype(np.int64)
    h = a[:, xo[:, 0], :] * xc[None, :, 0, None] + a[:, xo[:, 1], :] * xc[None, :, 1, None]
    s0, s1 = h[yo[:, 0]], h[yo[:, 1]]
    b0, b1 = yc[:, 0, None, None], yc[:, 1, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out
