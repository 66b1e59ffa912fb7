"""ctypes binding of libflope_b200.so (include/flope_b200.h) - the only way Python reaches the kernels.

PyTorch is used for device memory and streams only: tensors are passed as raw device
pointers and the current CUDA stream handle.  There is no fallback: if the library is
missing or a call fails, FlopeError is raised.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflope_b200.so")

INTERP_LINEAR, INTERP_LANCZOS4 = 0, 1
OUT_F32_NCHW, OUT_ENGINE = 0, 1
SCHED_PERSISTENT, SCHED_PER_LAYER, SCHED_DYNAMIC, SCHED_COOPERATIVE = 0, 1, 2, 3      # FLOPE_SCHED_* (include/flope_b200.h)

SYMBOLS = [
    "flope_version", "flope_last_error", "flope_engine_create", "flope_engine_destroy",
    "flope_engine_load_weights", "flope_squarify_filter", "flope_roi_crop", "flope_posenet_forward",
    "flope_pose_head", "flope_nullify_yaw", "flope_infer_frames", "flope_engine_last_launches",
    "flope_debug_activation", "flope_debug_normalise_lut", "flope_debug_set", "flope_engine_set_schedule", "flope_pack_boxes", "flope_debug_timeline", "flope_ingest_crops",
    "flope_engine_profile", "flope_engine_profile_read", "flope_depth_values", "flope_yolo_mask",
]


class FlopeError(RuntimeError):
    pass


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int), ("shape", C.c_int64 * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FlopeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.flope_last_error.restype = C.c_char_p
        L.flope_engine_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
        L.flope_engine_destroy.argtypes = [C.c_void_p]
        L.flope_engine_destroy.restype = None
        L.flope_engine_load_weights.argtypes = [C.c_void_p, C.POINTER(TensorDesc), C.c_int]
        L.flope_squarify_filter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.flope_roi_crop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.flope_posenet_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.flope_ingest_crops.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.flope_pose_head.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.flope_nullify_yaw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.flope_infer_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.flope_engine_last_launches.argtypes = [C.c_void_p]
        L.flope_debug_activation.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.flope_debug_activation.restype = C.c_int64
        L.flope_debug_normalise_lut.argtypes = [C.c_void_p, C.c_void_p]
        L.flope_debug_set.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.flope_engine_set_schedule.argtypes = [C.c_void_p, C.c_int]
        L.flope_pack_boxes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.flope_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.flope_engine_profile.argtypes = [C.c_void_p, C.c_int]
        L.flope_engine_profile_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.c_int]
        L.flope_depth_values.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
        L.flope_yolo_mask.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        raise FlopeError(f"flope_b200 error {rc}: {lib().flope_last_error().decode()}")
    return rc


def squarify_filter(boxes_i32, H, W):
    """(N,4) int32 -> (kept square boxes (M,4) int32, keep (N,) bool).  Host function of the C ABI."""
    n = boxes_i32.shape[0]
    sq = np.zeros((n, 4), np.int32)
    keep = np.zeros((n,), np.uint8)
    check(lib().flope_squarify_filter(boxes_i32.ctypes.data, n, H, W, sq.ctypes.data, keep.ctypes.data))
    keep = keep.astype(bool)
    return sq[keep], keep


def pack_boxes(img, boxes_xyxy, slot_h, slot_w, out):
    """Host: copy the box regions of a contiguous uint8 (H,W[,ch]) numpy image into the slots of `out` (n,slot_h,slot_w[,ch])."""
    ch = img.shape[2] if img.ndim == 3 else 1
    check(lib().flope_pack_boxes(img.ctypes.data, img.shape[0], img.shape[1], ch, boxes_xyxy.ctypes.data, boxes_xyxy.shape[0],
                                 int(slot_h), int(slot_w), out.ctypes.data))


def _stream(device=None):
    """The current torch stream of `device` (an index, a torch.device or a CUDA tensor); default: the current device."""
    import torch
    if device is not None and hasattr(device, "device"):
        device = device.device
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _want(t, name, dtype, device, ndim=None, last=None, allow_none=False):
    """Validate an argument that is handed to the C ABI as a raw pointer: CUDA tensor on the engine's device with the
    dtype and layout the kernels assume.  Non-contiguous tensors are made contiguous (a copy), everything else that
    does not match raises FlopeError instead of being reinterpreted as bytes."""
    import torch
    if t is None:
        if allow_none:
            return None
        raise FlopeError(f"{name} is required")
    if not torch.is_tensor(t) or not t.is_cuda:
        raise FlopeError(f"{name} must be a CUDA tensor (there is no CPU fallback)")
    if t.device.index != device:
        raise FlopeError(f"{name} lives on {t.device}, the engine on cuda:{device}")
    if t.dtype != dtype:
        raise FlopeError(f"{name} must be {dtype}, not {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise FlopeError(f"{name} must have {ndim} dimensions, not {tuple(t.shape)}")
    if last is not None and t.shape[-1] != last:
        raise FlopeError(f"{name} must have a last dimension of {last}, not {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def _check_boxes5(boxes5, n_frames, H, W):
    """Host-side range check of (n,5) boxes that are still on the host (numpy): frame index, in-frame, positive area."""
    b = np.asarray(boxes5)
    if b.size == 0:
        return
    if b[:, 0].min() < 0 or b[:, 0].max() >= n_frames:
        raise FlopeError("box frame index out of range")
    if (b[:, 1] < 0).any() or (b[:, 2] < 0).any() or (b[:, 3] > W).any() or (b[:, 4] > H).any():
        raise FlopeError("box outside the frame (squarify_filter drops those)")
    if (b[:, 3] <= b[:, 1]).any() or (b[:, 4] <= b[:, 2]).any():
        raise FlopeError("empty box (xmax <= xmin or ymax <= ymin): the reference's cv2.resize raises on these")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def depth_values(depth, mask, boxes, near_plane, far_plane, depth_div=None, erode_k=10):
    """flope_depth_values on device tensors: depth (H,W) float32 metres, or uint16 raw with depth_div;
    mask (H,W) uint8; boxes (n,4) int32 xyxy  ->  (val (n,) float64 metres, count (n,) int32), both on the device."""
    import torch
    if not depth.is_cuda:
        raise FlopeError("depth_values needs CUDA tensors (there is no CPU fallback)")
    H, W = depth.shape
    n = int(boxes.shape[0])
    mask = _want(mask, "mask", torch.uint8, depth.device.index, ndim=2)
    if tuple(mask.shape) != (H, W):
        raise FlopeError(f"mask {tuple(mask.shape)} does not match depth {(H, W)}")
    boxes = _want(boxes, "boxes", torch.int32, depth.device.index, ndim=2, last=4)
    if depth.dtype == torch.uint16:
        if depth_div is None:
            raise FlopeError("uint16 depth needs depth_div (sensor units per metre)")
        dtype, div = 1, float(depth_div)
    elif depth.dtype == torch.float32:
        dtype, div = 0, 1.0
    else:
        raise FlopeError(f"depth must be float32 or uint16, not {depth.dtype}")
    dev = depth.device
    scratch = torch.empty((H, W), dtype=torch.uint8, device=dev)
    val = torch.empty((n,), dtype=torch.float64, device=dev)
    cnt = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().flope_depth_values(dev.index or 0, _ptr(depth.contiguous()), dtype, div, _ptr(mask.contiguous()), H, W,
                                       _ptr(boxes.contiguous()), n, float(near_plane), float(far_plane), int(erode_k),
                                       _ptr(scratch), _ptr(val), _ptr(cnt), _stream(dev)))
    return val, cnt, scratch


def yolo_mask(masks, H, W):
    """flope_yolo_mask: (n,h,w) float32 CUDA instance masks -> (H,W) uint8 CUDA mask (0/255 union, cv2-exact resize)."""
    import torch
    if not masks.is_cuda:
        raise FlopeError("yolo_mask needs CUDA tensors (there is no CPU fallback)")
    masks = masks.to(torch.float32).contiguous()
    n, h, w = masks.shape
    dev = masks.device
    small = torch.empty((h, w), dtype=torch.uint8, device=dev)
    out = torch.empty((H, W), dtype=torch.uint8, device=dev)
    tables = torch.empty(((W + H) * 8,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().flope_yolo_mask(dev.index or 0, _ptr(masks), n, h, w, _ptr(small), _ptr(out), int(H), int(W),
                                    _ptr(tables), _stream(dev)))
    return out


class Engine:
    """Owns one flope_engine (weights + workspace) on one CUDA device."""

    def __init__(self, device=0, max_batch=256, crop_hw=224):
        import torch
        if not torch.cuda.is_available():
            raise FlopeError("flope_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = int(device)
        self.max_batch = int(max_batch)
        self.crop_hw = int(crop_hw)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().flope_engine_create(C.byref(self._h), self.device, self.max_batch, self.crop_hw))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().flope_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """Takes the reference PoseResNet.state_dict() (torch tensors or numpy arrays)."""
        keep, descs = [], []
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue
            a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append((k.encode(), a))
        arr = (TensorDesc * len(keep))()
        for i, (name, a) in enumerate(keep):
            arr[i].name = name
            arr[i].data = a.ctypes.data
            arr[i].ndim = a.ndim
            for d in range(a.ndim):
                arr[i].shape[d] = a.shape[d]
        check(lib().flope_engine_load_weights(self._h, arr, len(keep)))

    # -- kernels ------------------------------------------------------------------
    def _frame_args(self, frames, masks, boxes5):
        """(frames (n,H,W,3) u8, masks (n,H,W) u8 or None, boxes5 (m,5) i32 CUDA tensor or host numpy array) validated
        for the raw-pointer calls.  Host boxes are range-checked (frame index, in-frame, positive area) and uploaded."""
        import torch
        frames = _want(frames, "frames", torch.uint8, self.device, ndim=4, last=3)
        n_frames, H, W, _ = frames.shape
        masks = _want(masks, "masks", torch.uint8, self.device, ndim=3, allow_none=True)
        if masks is not None and tuple(masks.shape) != (n_frames, H, W):
            raise FlopeError(f"masks {tuple(masks.shape)} do not match frames {tuple(frames.shape)}")
        if not torch.is_tensor(boxes5):
            b = np.ascontiguousarray(boxes5, dtype=np.int32).reshape(-1, 5)
            _check_boxes5(b, n_frames, H, W)
            boxes5 = torch.from_numpy(b).to(frames.device)
        boxes5 = _want(boxes5, "boxes5", torch.int32, self.device, ndim=2, last=5)
        return frames, masks, boxes5

    def roi_crop(self, frames, masks, boxes5, S, interp, out=None, out_fmt=OUT_F32_NCHW):
        import torch
        frames, masks, boxes5 = self._frame_args(frames, masks, boxes5)
        n = boxes5.shape[0]
        n_frames, H, W, _ = frames.shape
        if out_fmt == OUT_F32_NCHW:
            if out is None:
                out = torch.empty((n, 3, S, S), dtype=torch.float32, device=frames.device)
            elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < n * 3 * S * S or out.device != frames.device:
                raise FlopeError("out must be a contiguous float32 CUDA tensor of at least (n,3,S,S) on the engine's device")
        check(lib().flope_roi_crop(self._h, _ptr(frames), n_frames, H, W, frames.stride(0), _ptr(masks), _ptr(boxes5), n, S,
                                   interp, _ptr(out), out_fmt, _stream(self.device)))
        return out

    def ingest_crops(self, x):
        """Stage (n,3,S,S) float32 crops as the engine's stem input; posenet_forward(None, n) consumes them."""
        check(lib().flope_ingest_crops(self._h, _ptr(x), x.shape[0], _stream(self.device)))

    def posenet_forward(self, x, n=None, out=None):
        """x: (n,3,S,S) float32 cuda tensor, or None to consume crops written by roi_crop(OUT_ENGINE)."""
        import torch
        if x is not None:
            x = _want(x, "x", torch.float32, self.device, ndim=4)
            if x.shape[1] != 3 or x.shape[2] != self.crop_hw or x.shape[3] != self.crop_hw:
                raise FlopeError(f"x must be (n,3,{self.crop_hw},{self.crop_hw}), not {tuple(x.shape)}")
            n = x.shape[0]
        if out is None:
            out = torch.empty((n, 9), dtype=torch.float32, device=f"cuda:{self.device}")
        elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < n * 9 or out.device.index != self.device:
            raise FlopeError("out must be a contiguous float32 CUDA tensor of at least (n,9) on the engine's device")
        check(lib().flope_posenet_forward(self._h, _ptr(x), n, _ptr(out), _stream(self.device)))
        return out

    def pose_head(self, r9, want_yaw=True):
        import torch
        n = r9.shape[0]
        R = torch.empty((n, 3, 3), dtype=torch.float32, device=r9.device)
        Ry = torch.empty((n, 3, 3), dtype=torch.float64, device=r9.device) if want_yaw else None
        check(lib().flope_pose_head(self._h, _ptr(r9), n, _ptr(R), _ptr(Ry), _stream(self.device)))
        return R, Ry

    def nullify_yaw(self, R):
        import torch
        n = R.shape[0]
        Ry = torch.empty((n, 3, 3), dtype=torch.float64, device=R.device)
        check(lib().flope_nullify_yaw(self._h, _ptr(R), n, _ptr(Ry), _stream(self.device)))
        return Ry

    def infer_frames(self, frames, masks, boxes5, interp, want_r9=False, want_R=True, want_yaw=True, out=None):
        import torch
        frames, masks, boxes5 = self._frame_args(frames, masks, boxes5)
        n = boxes5.shape[0]
        n_frames, H, W, _ = frames.shape
        dev = frames.device
        if out is not None and (out.dtype != torch.float64 or not out.is_contiguous() or out.numel() < n * 9 or out.device != dev):
            raise FlopeError("out must be a contiguous float64 CUDA tensor of at least (n,3,3) on the engine's device")
        r9 = torch.empty((n, 9), dtype=torch.float32, device=dev) if want_r9 else None
        R = torch.empty((n, 3, 3), dtype=torch.float32, device=dev) if want_R else None
        Ry = (out if out is not None else torch.empty((n, 3, 3), dtype=torch.float64, device=dev)) if want_yaw else None
        check(lib().flope_infer_frames(self._h, _ptr(frames), n_frames, H, W, frames.stride(0), _ptr(masks), _ptr(boxes5), n,
                                       interp, _ptr(r9), _ptr(R), _ptr(Ry), _stream(self.device)))
        return r9, R, Ry

    def last_launches(self):
        return int(lib().flope_engine_last_launches(self._h))

    def profile(self, enable=True):
        """1/True: an event pair per launch; 2: one pair around the trunk's conv chain; 0/False: off."""
        check(lib().flope_engine_profile(self._h, int(enable)))

    def profile_read(self, max_entries=65536):
        """-> list of (kernel name, milliseconds) for every launch since profile(True)."""
        names = C.create_string_buffer(max_entries * 24)
        ms = (C.c_float * max_entries)()
        n = check(lib().flope_engine_profile_read(self._h, names, len(names), ms, max_entries))
        return list(zip(names.value.decode().split("\n")[:n], list(ms[:n])))

    # -- test hooks ---------------------------------------------------------------
    def debug_activation(self, name, n):
        import torch
        # generous upper bound, then trim to the size the library reports
        S = self.crop_hw
        buf = torch.empty((n * 64 * (S // 2) * (S // 2),), dtype=torch.float32, device=f"cuda:{self.device}")
        chw = check(lib().flope_debug_activation(self._h, name.encode(), n, _ptr(buf), _stream(self.device)))
        return buf[: n * chw], chw

    def timeline(self, max_launches=32):
        """Phase stamps of the conv launches of the last forward: uint64 array (launches, 148, 8); needs
        debug_set("timeline", 1).  See flope_debug_timeline in include/flope_b200.h."""
        import numpy as np
        buf = np.zeros((max_launches, 148, 8), np.uint64)
        n = check(lib().flope_debug_timeline(self._h, buf.ctypes.data, max_launches))
        return buf[:n]

    def set_schedule(self, schedule):
        """SCHED_PERSISTENT (default) / SCHED_PER_LAYER / SCHED_DYNAMIC / SCHED_COOPERATIVE: see include/flope_b200.h."""
        check(lib().flope_engine_set_schedule(self._h, int(schedule)))

    def debug_set(self, key, value):
        check(lib().flope_debug_set(self._h, key.encode(), int(value)))


def debug_normalise_lut(device=0):
    import torch
    out = torch.empty((256, 256), dtype=torch.float32, device=f"cuda:{device}")
    check(lib().flope_debug_normalise_lut(_ptr(out), _stream()))
    return out
