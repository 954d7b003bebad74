"""NumPy-facing wrappers of the C ABI, one per OpenCV call the reference makes on the
hot path.  Signatures follow the cv2 functions they replace; all pixel work happens in
libbevk.so on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

INTER_NEAREST, INTER_LINEAR = L.INTER_NEAREST, L.INTER_LINEAR


def _interp(flag: int) -> int:
    if flag not in (INTER_NEAREST, INTER_LINEAR):
        raise L.BevkError(f"interpolation {flag} is not supported (INTER_NEAREST / INTER_LINEAR only)")
    return flag


def fisheye_init_undistort_rectify_map(K, D, P, size, ctx: L.Context | None = None):
    """cv2.fisheye.initUndistortRectifyMap(K, D, eye(3), P, size, CV_16SC2)."""
    return _undistort_map(L.MODEL_FISHEYE, K, D, P, size, ctx)


def init_undistort_rectify_map(K, D, P, size, ctx: L.Context | None = None):
    """cv2.initUndistortRectifyMap(K, D(k1,k2,p1,p2,k3), eye(3), P, size, CV_16SC2)."""
    return _undistort_map(L.MODEL_PINHOLE, K, D, P, size, ctx)


def _undistort_map(model, K, D, P, size, ctx):
    ctx = ctx or L.default_context()
    w, h = int(size[0]), int(size[1])
    d = np.asarray(D, np.float64).reshape(-1)
    m1 = np.empty((h, w, 2), np.int16)
    m2 = np.empty((h, w), np.uint16)
    L.check(ctx.lib.bevk_undistort_map(ctx.h, model, L.dptr(K), L.dptr(d), int(d.size), L.dptr(P), w, h,
                                       L.vptr(m1), L.vptr(m2)))
    return m1, m2


def _out(shape, out):
    """cv2-style optional destination (e.g. a page-locked array from pinned_empty for full-rate D2H)."""
    if out is None:
        return np.empty(shape, np.uint8)
    if out.dtype != np.uint8 or out.shape != tuple(shape) or not out.flags.c_contiguous:
        raise L.BevkError(f"out must be a C-contiguous uint8 array of shape {tuple(shape)}")
    return out


def _cuda_ptr(arr, shape):
    """(device pointer, shape) of a uint8 C-contiguous array exposing ``__cuda_array_interface__``."""
    iface = getattr(arr, "__cuda_array_interface__", None)
    if iface is None:
        raise L.BevkError("expected a CUDA array (an object with __cuda_array_interface__)")
    got = tuple(iface["shape"])
    if iface["typestr"] not in ("|u1", "<u1", "=u1"):
        raise L.BevkError(f"CUDA array must be uint8, got typestr {iface['typestr']}")
    if shape is not None and got != tuple(shape):
        raise L.BevkError(f"CUDA array must have shape {tuple(shape)}, got {got}")
    strides = iface.get("strides")
    if strides is not None:
        dense, step = [], 1
        for n in reversed(got):
            dense.append(step)
            step *= n
        if any(n > 1 and s != d for n, s, d in zip(got, strides, reversed(dense))):
            raise L.BevkError("CUDA array must be C-contiguous")
    ptr = iface["data"][0]
    if not ptr:
        raise L.BevkError("CUDA array has a null data pointer")
    return int(ptr), got


def jpeg_decode(jpegs, width: int, height: int, out=None, ctx: L.Context | None = None):
    """Decode JPEG byte strings on the GPU (nvJPEG) into a uint8 CUDA frame stack [n][height][width][3] (BGR, what
    cv2.imread's layout is).  ``out``: a CUDA array of that shape; default a new torch tensor.  Enqueued on the ctx
    stream; returns ``out``."""
    ctx = ctx or L.default_context()
    n = len(jpegs)
    if out is None:
        import torch
        out = torch.empty((n, height, width, 3), dtype=torch.uint8, device=torch.device("cuda", ctx.device))
    d_out = _cuda_ptr(out, (n, height, width, 3))[0]
    keep = [(C.c_char * len(x)).from_buffer_copy(bytes(x)) for x in jpegs]
    ptrs = (C.c_void_p * n)(*[C.addressof(k) for k in keep])
    sizes = (C.c_uint64 * n)(*[len(x) for x in jpegs])
    L.check(ctx.lib.bevk_jpeg_decode(ctx.h, ptrs, sizes, n, int(width), int(height), C.c_void_p(d_out), width * height * 3))
    return out


def remap(src: np.ndarray, map1: np.ndarray, map2: np.ndarray | None, interpolation: int = INTER_LINEAR,
          ctx: L.Context | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """cv2.remap with CV_16SC2 (+CV_16UC1) maps, BORDER_CONSTANT 0."""
    ctx = ctx or L.default_context()
    img, sw, sh, ss, ch = L.image_view(src)
    m1 = np.ascontiguousarray(map1, np.int16)
    if m1.ndim != 3 or m1.shape[2] != 2:
        raise L.BevkError("map1 must be int16[h][w][2] (CV_16SC2)")
    dh, dw = m1.shape[:2]
    m2 = None if map2 is None else np.ascontiguousarray(map2, np.uint16)
    if m2 is not None and m2.shape != (dh, dw):
        raise L.BevkError("map2 must be uint16[h][w] (CV_16UC1)")
    out = _out((dh, dw) if src.ndim == 2 else (dh, dw, ch), out)
    L.check(ctx.lib.bevk_remap(ctx.h, L.vptr(img), sw, sh, ss, ch, L.vptr(m1), None if m2 is None else L.vptr(m2),
                               dw, dh, L.vptr(out), dw * ch, _interp(interpolation)))
    return out


def warp_perspective(src: np.ndarray, H, dsize, flags: int = INTER_LINEAR, ctx: L.Context | None = None,
                     out: np.ndarray | None = None):
    """cv2.warpPerspective(src, H, dsize, flags) for uint8 images, BORDER_CONSTANT 0."""
    ctx = ctx or L.default_context()
    img, sw, sh, ss, ch = L.image_view(src)
    dw, dh = int(dsize[0]), int(dsize[1])
    out = _out((dh, dw) if src.ndim == 2 else (dh, dw, ch), out)
    L.check(ctx.lib.bevk_warp_perspective(ctx.h, L.vptr(img), sw, sh, ss, ch, L.dptr(H), L.vptr(out), dw, dh,
                                          dw * ch, _interp(flags)))
    return out


def warp_perspective_maps(map1, map2, H, dsize, ctx: L.Context | None = None):
    """cv2.warpPerspective applied to a CV_16SC2 / CV_16UC1 map pair (Camera.get_bev_maps)."""
    ctx = ctx or L.default_context()
    m1 = np.ascontiguousarray(map1, np.int16)
    m2 = np.ascontiguousarray(map2, np.uint16)
    sh, sw = m2.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    o1 = np.empty((dh, dw, 2), np.int16)
    o2 = np.empty((dh, dw), np.uint16)
    L.check(ctx.lib.bevk_warp_maps(ctx.h, L.vptr(m1), L.vptr(m2), sw, sh, L.dptr(H), dw, dh, L.vptr(o1), L.vptr(o2)))
    return o1, o2


def _dense_bgr(img: np.ndarray) -> np.ndarray:
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
        raise L.BevkError("expected a uint8[h][w][3] BGR image")
    return np.ascontiguousarray(img)


def apply_mask(img: np.ndarray, mask: np.ndarray, blend: bool, ctx: L.Context | None = None) -> np.ndarray:
    """Mask.__call__ (blend=False) / BlendMask.__call__ (blend=True)."""
    ctx = ctx or L.default_context()
    a = _dense_bgr(img)
    m = np.ascontiguousarray(mask, np.uint8)
    if m.shape != a.shape[:2]:
        raise L.BevkError("mask and image sizes differ")
    out = np.empty_like(a)
    L.check(ctx.lib.bevk_apply_mask(ctx.h, L.vptr(a), L.vptr(m), a.shape[1], a.shape[0], int(blend), L.vptr(out)))
    return out


def color_balance(image: np.ndarray, ctx: L.Context | None = None) -> np.ndarray:
    ctx = ctx or L.default_context()
    a = _dense_bgr(image)
    out = np.empty_like(a)
    L.check(ctx.lib.bevk_color_balance(ctx.h, L.vptr(a), a.shape[1], a.shape[0], L.vptr(out)))
    return out


def luminance_balance(images, ctx: L.Context | None = None):
    ctx = ctx or L.default_context()
    imgs = [_dense_bgr(i) for i in images]
    if len({i.shape for i in imgs}) != 1:
        raise L.BevkError("luminance_balance: frames must share one size")
    outs = [np.empty_like(i) for i in imgs]
    n = len(imgs)
    ip = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    L.check(ctx.lib.bevk_luminance_balance(ctx.h, ip, n, imgs[0].shape[1], imgs[0].shape[0], op))
    return outs


class Undistorter:
    """Device-resident undistortion map (or fused camera model) + per-frame gather.

    A bevk_ctx has 8 undistorter slots.  Each live Undistorter owns one slot of its ctx; the slot returns to the
    pool on close() / garbage collection, and a 9th live object on one ctx raises instead of silently taking over a
    slot another object still uses."""

    def __init__(self, K, D, P, size, model: str = "fisheye", fused: bool = False, ctx: L.Context | None = None,
                 slot: int | None = None):
        self.ctx = ctx or L.default_context()
        self.slot = None
        live = self.ctx.__dict__.setdefault("_und_slots", set())
        if slot is None:
            free = [i for i in range(8) if i not in live]
            if not free and ctx is None:
                # the shared default context is full (e.g. two BevGenerators' cameras): this object gets its own
                self.ctx = L.Context(self.ctx.device)
                live = self.ctx.__dict__.setdefault("_und_slots", set())
                free = [0]
            if not free:
                raise L.BevkError("all 8 undistorter slots of this context are in use: close() an Undistorter "
                                  "(or let it be collected), or give this one its own Context")
            slot = free[0]
        elif not 0 <= int(slot) < 8:
            raise L.BevkError(f"slot {slot} out of range [0, 8)")
        elif slot in live:
            raise L.BevkError(f"undistorter slot {slot} of this context is owned by a live Undistorter")
        self.w, self.h = int(size[0]), int(size[1])
        d = np.asarray(D, np.float64).reshape(-1)
        m = L.MODEL_FISHEYE if model == "fisheye" else L.MODEL_PINHOLE
        L.check(self.ctx.lib.bevk_undistorter_set(self.ctx.h, slot, m, L.dptr(K), L.dptr(d), int(d.size), L.dptr(P),
                                                  self.w, self.h, int(fused)))
        self.slot = int(slot)
        live.add(self.slot)

    def close(self):
        if getattr(self, "slot", None) is not None:
            self.ctx.__dict__.get("_und_slots", set()).discard(self.slot)
            self.slot = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _live(self):
        if self.slot is None:
            raise L.BevkError("this Undistorter was closed")

    def maps(self):
        self._live()
        m1 = np.empty((self.h, self.w, 2), np.int16)
        m2 = np.empty((self.h, self.w), np.uint16)
        L.check(self.ctx.lib.bevk_undistorter_maps(self.ctx.h, self.slot, L.vptr(m1), L.vptr(m2)))
        return m1, m2

    def __call__(self, src: np.ndarray, interpolation: int = INTER_LINEAR, out: np.ndarray | None = None) -> np.ndarray:
        self._live()
        img, sw, sh, ss, ch = L.image_view(src)
        out = _out((self.h, self.w) if src.ndim == 2 else (self.h, self.w, ch), out)
        L.check(self.ctx.lib.bevk_undistort(self.ctx.h, self.slot, L.vptr(img), sw, sh, ss, ch, L.vptr(out), self.w, self.h,
                                            self.w * ch, _interp(interpolation)))
        return out


class BevEngine:
    """The fused surround-BEV engine (bevk_bev_* entry points)."""

    def __init__(self, n_cam: int, frame_size, bev_size, ctx: L.Context | None = None):
        self.ctx = ctx or L.Context(L.default_context().device)   # own ctx: the plan is per-ctx state
        self.n_cam = n_cam
        self.FW, self.FH = int(frame_size[0]), int(frame_size[1])
        self.BW, self.BH = int(bev_size[0]), int(bev_size[1])
        L.check(self.ctx.lib.bevk_bev_configure(self.ctx.h, n_cam, self.FW, self.FH, self.BW, self.BH))
        self.finalized = False

    def set_camera(self, cam: int, K, D, P, und_size, H):
        d = np.zeros(4)
        dd = np.asarray(D, np.float64).reshape(-1)
        d[:min(4, dd.size)] = dd[:4]
        L.check(self.ctx.lib.bevk_bev_set_camera(self.ctx.h, cam, L.dptr(K), L.dptr(d), L.dptr(P),
                                                 int(und_size[0]), int(und_size[1]), L.dptr(H)))
        self.finalized = False

    def set_maps(self, cam: int, map1, map2):
        m1 = np.ascontiguousarray(map1, np.int16)
        m2 = np.ascontiguousarray(map2, np.uint16)
        if m1.shape != (self.BH, self.BW, 2) or m2.shape != (self.BH, self.BW):
            raise L.BevkError("BEV maps must match the canvas size")
        L.check(self.ctx.lib.bevk_bev_set_maps(self.ctx.h, cam, L.vptr(m1), L.vptr(m2)))
        self.finalized = False

    def get_maps(self, cam: int):
        m1 = np.empty((self.BH, self.BW, 2), np.int16)
        m2 = np.empty((self.BH, self.BW), np.uint16)
        L.check(self.ctx.lib.bevk_bev_get_maps(self.ctx.h, cam, L.vptr(m1), L.vptr(m2)))
        return m1, m2

    def set_interpolation(self, interpolation: int):
        """INTER_LINEAR (the reference) or INTER_NEAREST for the raw2bev gather; before finalize()."""
        L.check(self.ctx.lib.bevk_bev_set_interpolation(self.ctx.h, _interp(interpolation)))
        self.finalized = False

    def set_mask(self, cam: int, mask: np.ndarray):
        m = np.ascontiguousarray(mask, np.uint8)
        if m.shape != (self.BH, self.BW):
            raise L.BevkError("mask must be uint8[bev_h][bev_w]")
        L.check(self.ctx.lib.bevk_bev_set_mask(self.ctx.h, cam, L.vptr(m)))
        self.finalized = False

    def blend_masks(self, polys: np.ndarray, lines: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(polys, np.uint8)
        ln = np.ascontiguousarray(lines, np.int32).reshape(8, 4)
        out = np.empty_like(p)
        L.check(self.ctx.lib.bevk_blend_masks(self.ctx.h, L.vptr(p), L.vptr(ln), self.BW, self.BH, L.vptr(out)))
        return out

    def finalize(self):
        L.check(self.ctx.lib.bevk_bev_finalize(self.ctx.h))
        self.finalized = True

    def plan_info(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        L.check(self.ctx.lib.bevk_bev_plan_info(self.ctx.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"tiles": a.value, "items": b.value, "lut_bytes": c.value}

    def run(self, frame_sets, car: np.ndarray | None = None, balance: bool = False, out: np.ndarray | None = None):
        """frame_sets: list (batch) of lists (n_cam) of uint8[FH][FW][3] arrays.  Returns
        uint8[batch][BH][BW][3]."""
        if not self.finalized:
            self.finalize()
        batch = len(frame_sets)
        if batch < 1:
            raise L.BevkError("run() needs at least one frame-set")
        keep, ptrs, stride = [], (C.c_void_p * (batch * self.n_cam))(), None
        for b, fs in enumerate(frame_sets):
            if len(fs) != self.n_cam:
                raise L.BevkError(f"frame-set {b} has {len(fs)} frames, expected {self.n_cam}")
            for k, f in enumerate(fs):
                f = self._conform(f)
                img, w, h, s, ch = L.image_view(f)
                if stride is None:
                    stride = s
                elif s != stride:
                    img = np.ascontiguousarray(img)
                    if img.strides[0] != stride:
                        raise L.BevkError("all frames of a call must share one row stride")
                keep.append(img)
                ptrs[b * self.n_cam + k] = img.ctypes.data
        if out is None:   # a fresh array per call, as the reference returns -- page-locked and recycled (PinnedPool)
            if getattr(self, "_pool", None) is None:
                self._pool = L.PinnedPool()
            out = self._pool.get((batch, self.BH, self.BW, 3))
        else:
            out = _out((batch, self.BH, self.BW, 3), out)
        carp = None
        if car is not None:
            car = np.ascontiguousarray(car, np.uint8)
            if car.shape != (self.BH, self.BW, 3):
                raise L.BevkError("car must be uint8[bev_h][bev_w][3]")
            carp = L.vptr(car)
        L.check(self.ctx.lib.bevk_bev_run(self.ctx.h, ptrs, stride, batch, carp, L.FLAG_BALANCE if balance else 0,
                                          L.vptr(out)))
        return out

    def run_jpeg(self, jpeg_sets, car: np.ndarray | None = None, balance: bool = False, out: np.ndarray | None = None):
        """jpeg_sets: list (batch) of lists (n_cam) of JPEG byte strings (the files cv2.imread would open).  The streams
        are decoded on the GPU (nvJPEG) straight into the frame stack the fused kernel reads: only compressed bytes cross
        PCIe on the way in.  Pixels are nvJPEG's decode (not libjpeg-turbo's); the BEV path on them is bit-exact."""
        if not self.finalized:
            self.finalize()
        batch = len(jpeg_sets)
        if batch < 1:
            raise L.BevkError("run_jpeg() needs at least one frame-set")
        flat = []
        for b, fs in enumerate(jpeg_sets):
            if len(fs) != self.n_cam:
                raise L.BevkError(f"frame-set {b} has {len(fs)} streams, expected {self.n_cam}")
            flat += [bytes(x) if not isinstance(x, (bytes, bytearray)) else x for x in fs]
        keep = [(C.c_char * len(x)).from_buffer_copy(x) if isinstance(x, bytes) else (C.c_char * len(x)).from_buffer(x) for x in flat]
        ptrs = (C.c_void_p * len(flat))(*[C.addressof(k) for k in keep])
        sizes = (C.c_uint64 * len(flat))(*[len(x) for x in flat])
        if out is None:
            if getattr(self, "_pool", None) is None:
                self._pool = L.PinnedPool()
            out = self._pool.get((batch, self.BH, self.BW, 3))
        else:
            out = _out((batch, self.BH, self.BW, 3), out)
        carp = None
        if car is not None:
            car = np.ascontiguousarray(car, np.uint8)
            if car.shape != (self.BH, self.BW, 3):
                raise L.BevkError("car must be uint8[bev_h][bev_w][3]")
            carp = L.vptr(car)
        L.check(self.ctx.lib.bevk_bev_run_jpeg(self.ctx.h, ptrs, sizes, batch, carp, L.FLAG_BALANCE if balance else 0, L.vptr(out)))
        return out

    def host_copy_bytes(self, balance: bool = False):
        """(host->device, device->host) bytes per frame-set that run() moves over PCIe."""
        if not self.finalized:
            self.finalize()
        a, b = C.c_int64(), C.c_int64()
        L.check(self.ctx.lib.bevk_bev_host_copy_bytes(self.ctx.h, L.FLAG_BALANCE if balance else 0, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_h2d_bytes(self) -> int:
        """Host->device bytes moved by the last run() call."""
        return int(self.ctx.lib.bevk_bev_last_h2d_bytes(self.ctx.h))

    def _conform(self, f: np.ndarray) -> np.ndarray:
        """The reference never validates frame sizes (cv2.remap samples whatever it is given, zero outside).  The engine's
        LUT is compiled for FW x FH: a SMALLER frame is embedded into an FW x FH zero canvas, which samples identically
        (cv2's BORDER_CONSTANT 0).  A LARGER frame is cropped, which differs from the reference wherever the LUT points
        beyond FW x FH (the reference would sample the extra pixels there, this engine reads zeros)."""
        if f.ndim != 3 or f.shape[2] != 3 or f.dtype != np.uint8:
            raise L.BevkError("frames must be uint8[h][w][3] (BGR)")
        if f.shape[0] == self.FH and f.shape[1] == self.FW:
            return f
        g = np.zeros((self.FH, self.FW, 3), np.uint8)
        h, w = min(self.FH, f.shape[0]), min(self.FW, f.shape[1])
        g[:h, :w] = f[:h, :w]
        return g

    # device-resident entry points (raw device pointers, e.g. torch tensors' data_ptr())
    def run_device(self, d_srcs_ptr: int, batch: int, d_out_ptr: int, d_car_ptr: int = 0, balance: bool = False):
        if not self.finalized:
            self.finalize()
        L.check(self.ctx.lib.bevk_bev_run_device(self.ctx.h, C.c_void_p(d_srcs_ptr), batch, C.c_void_p(d_car_ptr or None),
                                                 L.FLAG_BALANCE if balance else 0, C.c_void_p(d_out_ptr)))

    def run_stack(self, d_frames_ptr: int, frame_stride: int, batch: int, d_out_ptr: int, d_car_ptr: int = 0, balance: bool = False):
        """Frame stack on the device (frame i at d_frames_ptr + i * frame_stride, i = set * n_cam + camera): the
        TMA-staged kernel when base and stride are 16-byte aligned.  Only enqueues on the ctx stream."""
        if not self.finalized:
            self.finalize()
        L.check(self.ctx.lib.bevk_bev_run_stack(self.ctx.h, C.c_void_p(d_frames_ptr), int(frame_stride), batch,
                                                C.c_void_p(d_car_ptr or None), L.FLAG_BALANCE if balance else 0, C.c_void_p(d_out_ptr)))

    def run_stack_cams(self, d_frames_ptr: int, frame_stride: int, batch: int, cam_lo: int, cam_hi: int, d_out_ptr: int):
        if not self.finalized:
            self.finalize()
        L.check(self.ctx.lib.bevk_bev_run_stack_cams(self.ctx.h, C.c_void_p(d_frames_ptr), int(frame_stride), batch, cam_lo,
                                                     cam_hi, C.c_void_p(d_out_ptr)))

    def last_path(self) -> str:
        """Which fused kernel the last call launched: 'tma' (k_bev_tma) or 'gather' (k_bev)."""
        return {1: "gather", 2: "tma"}.get(int(self.ctx.lib.bevk_bev_last_path(self.ctx.h)), "none")

    def tma_plan_info(self):
        v = [C.c_int64() for _ in range(5)]
        L.check(self.ctx.lib.bevk_bev_tma_plan_info(self.ctx.h, *[C.byref(x) for x in v]))
        return dict(zip(("items", "shapes", "box_bytes", "tma_entries", "gather_entries"), (x.value for x in v)))

    def run_cuda(self, frames, car=None, balance: bool = False, out=None, stream: int | None = None):
        """Frames that already live on the GPU (decoder output, torch / CuPy arrays): no PCIe in the call.

        frames: one uint8 CUDA array [batch][n_cam][FH][FW][3], or a list (batch) of lists (n_cam) of uint8
        CUDA arrays [FH][FW][3] -- anything exposing ``__cuda_array_interface__``, C-contiguous, on this
        engine's device.  car / out likewise ([BH][BW][3] / [batch][BH][BW][3]); without ``out`` a torch
        tensor is allocated.  ``stream``: raw CUDA stream handle the work is enqueued on for THIS call; default torch's
        current stream on the engine's device (so that torch kernels that produced ``frames``, this render and whatever
        consumes ``out`` are ordered), or the ctx's own stream when torch is not in use.  The ctx goes back to its previous
        stream afterwards.  The call does not synchronise.  Returns ``out``."""
        if not self.finalized:
            self.finalize()
        frame_shape = (self.FH, self.FW, 3)
        if hasattr(frames, "__cuda_array_interface__"):
            base, shape = _cuda_ptr(frames, None)
            if len(shape) != 5 or tuple(shape[1:]) != (self.n_cam,) + frame_shape:
                raise L.BevkError(f"frames must be uint8[batch][{self.n_cam}][{self.FH}][{self.FW}][3], got {tuple(shape)}")
            batch, fb = shape[0], self.FH * self.FW * 3
            ptrs = [base + i * fb for i in range(batch * self.n_cam)]
        else:
            batch, ptrs = len(frames), []
            for b, fs in enumerate(frames):
                if len(fs) != self.n_cam:
                    raise L.BevkError(f"frame-set {b} has {len(fs)} frames, expected {self.n_cam}")
                ptrs += [_cuda_ptr(f, frame_shape)[0] for f in fs]
        if batch < 1:
            raise L.BevkError("batch must be >= 1")
        if out is None:
            import torch
            out = torch.empty((batch, self.BH, self.BW, 3), dtype=torch.uint8, device=torch.device("cuda", self.ctx.device))
        d_out = _cuda_ptr(out, (batch, self.BH, self.BW, 3))[0]
        d_car = _cuda_ptr(car, (self.BH, self.BW, 3))[0] if car is not None else None
        if stream is None:
            from .sharding import _torch_current_stream
            stream = _torch_current_stream(self.ctx.device)
        table = (C.c_void_p * len(ptrs))(*ptrs)
        with self.ctx.on_stream(stream):
            L.check(self.ctx.lib.bevk_bev_run_frames(self.ctx.h, table, batch, C.c_void_p(d_car), L.FLAG_BALANCE if balance else 0,
                                                     C.c_void_p(d_out)))
        return out

    def run_device_cams(self, d_srcs_ptr: int, batch: int, cam_lo: int, cam_hi: int, d_out_ptr: int):
        if not self.finalized:
            self.finalize()
        L.check(self.ctx.lib.bevk_bev_run_device_cams(self.ctx.h, C.c_void_p(d_srcs_ptr), batch, cam_lo, cam_hi,
                                                      C.c_void_p(d_out_ptr)))

    def sat_sum_device(self, part_ptrs, nbytes: int, d_out_ptr: int, d_car_ptr: int = 0):
        arr = (C.c_void_p * len(part_ptrs))(*part_ptrs)
        L.check(self.ctx.lib.bevk_sat_sum_device(self.ctx.h, arr, len(part_ptrs), nbytes, C.c_void_p(d_car_ptr or None),
                                                 C.c_void_p(d_out_ptr)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        L.check(self.ctx.lib.bevk_last_kernel_ms(self.ctx.h, C.byref(ms)))
        return ms.value
