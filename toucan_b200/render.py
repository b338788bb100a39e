"""ctypes binding of libtoucan_render.so (include/toucan_render.h): the C++ host --
ImageEffectHost + TimelineWrapper + ImageGraph -- for the Python test and bench harness.

Plumbing only.  Frames are numpy arrays / torch tensors of shape (h, w, 4); the render
itself is the C++ graph evaluation enqueuing the sm_100a kernels.  No fallback.
"""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

from . import capi

RENDER_LIB = os.path.join(capi.LIB_DIR, "libtoucan_render.so")
PLUGIN_DIR = os.path.join(capi.LIB_DIR, "plugins")

_NP_OF = {capi.U8: np.uint8, capi.U16: np.uint16, capi.F16: np.float16, capi.F32: np.float32}
_T_OF = {np.dtype(np.uint8): capi.U8, np.dtype(np.uint16): capi.U16, np.dtype(np.float16): capi.F16, np.dtype(np.float32): capi.F32}

_lib = None
_VP = ctypes.c_void_p
_IP = ctypes.POINTER(ctypes.c_int)
_DP = ctypes.POINTER(ctypes.c_double)
_SIGS = {
    "tr_bind_device": [ctypes.c_int],
    "tr_sync": [],
    "tr_bind_stream": [ctypes.c_int, _VP],
    "tr_host_create": [ctypes.c_char_p, ctypes.POINTER(_VP)],
    "tr_host_plugin_count": [_VP],
    "tr_timeline_open": [ctypes.c_char_p, ctypes.POINTER(_VP)],
    "tr_timeline_open_json": [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(_VP)],
    "tr_timeline_media_path": [_VP, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t],
    "tr_timeline_set_media": [_VP, ctypes.c_char_p, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int],
    "tr_timeline_prepare": [_VP],
    "tr_timeline_set_source_cache": [_VP, ctypes.c_size_t],
    "tr_timeline_upload_stats": [_VP, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)],
    "tr_timeline_drop_source_cache": [_VP],
    "tr_timeline_info": [_VP, _DP, _DP, _IP, _IP, _IP, _IP],
    "tr_timeline_describe": [_VP, _VP, ctypes.c_double, ctypes.c_char_p, ctypes.c_size_t],
    "tr_timeline_render": [_VP, _VP, ctypes.c_double, _VP, ctypes.c_size_t, _IP, _IP, _IP],
    "tr_decode_png": [ctypes.c_char_p, _VP, ctypes.c_size_t, _IP, _IP, _IP],
    "tr_decode_jpeg": [ctypes.c_char_p, _VP, ctypes.c_size_t, _IP, _IP, _IP],
    "tr_decode_exr": [ctypes.c_char_p, _VP, ctypes.c_size_t, _IP, _IP, _IP],
    "tr_decode_tiff": [ctypes.c_char_p, _VP, ctypes.c_size_t, _IP, _IP, _IP],
    "tr_write_exr": [ctypes.c_char_p, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int],
    "tr_timeline_render_resident": [_VP, _VP, ctypes.c_double, ctypes.POINTER(_VP), _IP, _IP, _IP],
}
FRAME_CALLBACK = ctypes.CFUNCTYPE(None, _VP, ctypes.c_int, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int)
_SIGS["tr_render_range"] = [_VP, ctypes.c_char_p, _IP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, FRAME_CALLBACK, _VP]
_SIGS["tr_render_range_raw"] = [_VP, ctypes.c_char_p, _IP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                FRAME_CALLBACK, _VP]
_SIGS["tr_render_range_y4m"] = _SIGS["tr_render_range_raw"]
_SIGS["tr_y4m_header"] = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
EXPORTS = sorted(set(_SIGS) | {"tr_frame_block", "tr_fit", "tr_last_error", "tr_host_destroy", "tr_host_plugin_identifier", "tr_timeline_close", "tr_set_fusion", "tr_set_interleave", "tr_set_node_timers", "tr_node_timers_report",
                                  "tr_set_unassociated_alpha", "tr_get_unassociated_alpha"})


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        capi.lib()  # libtoucan_cuda.so first (RTLD_GLOBAL)
        if not os.path.exists(RENDER_LIB):
            raise capi.TcError(f"{RENDER_LIB} is missing: run `python -m toucan_b200.build`")
        _lib = ctypes.CDLL(RENDER_LIB, mode=ctypes.RTLD_GLOBAL)
        for name, sig in _SIGS.items():
            fn = getattr(_lib, name)
            fn.argtypes = sig
            fn.restype = ctypes.c_int
        _lib.tr_last_error.restype = ctypes.c_char_p
        _lib.tr_host_destroy.argtypes = [_VP]
        _lib.tr_host_destroy.restype = None
        _lib.tr_timeline_close.argtypes = [_VP]
        _lib.tr_timeline_close.restype = None
        _lib.tr_host_plugin_identifier.argtypes = [_VP, ctypes.c_int]
        _lib.tr_host_plugin_identifier.restype = ctypes.c_char_p
        _lib.tr_set_fusion.argtypes = [ctypes.c_int]
        _lib.tr_set_fusion.restype = None
        _lib.tr_set_interleave.argtypes = [ctypes.c_int]
        _lib.tr_set_interleave.restype = None
        _lib.tr_set_node_timers.argtypes = [ctypes.c_int]
        _lib.tr_set_node_timers.restype = None
        _lib.tr_node_timers_report.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        _lib.tr_node_timers_report.restype = ctypes.c_size_t
    return _lib


def _check(status, what):
    if status != 0:
        raise capi.TcError(f"{what} failed: {lib().tr_last_error().decode()}")


def bind_device(device: int = 0):
    _check(lib().tr_bind_device(device), "tr_bind_device")


def bind_stream(device: int, stream_ptr: int):
    """Render on a stream the caller owns (e.g. torch's current stream, for CUDA-event timing)."""
    _check(lib().tr_bind_stream(device, _VP(stream_ptr)), "tr_bind_stream")


def sync():
    _check(lib().tr_sync(), "tr_sync")


def set_fusion(enabled: bool):
    lib().tr_set_fusion(1 if enabled else 0)


def set_node_timers(enabled: bool):
    """Per-node device timers (toucan::NodeTimers)."""
    lib().tr_set_node_timers(1 if enabled else 0)


def node_timers_report() -> str:
    buf = ctypes.create_string_buffer(1 << 16)
    lib().tr_node_timers_report(buf, len(buf))
    return buf.value.decode()


def set_interleave(enabled: bool):
    """Round-robin (default) or contiguous-block assignment of frames to the scheduler's workers."""
    lib().tr_set_interleave(1 if enabled else 0)


def y4m_header(width: int, height: int, rate: float, y4m_format: str) -> bytes:
    """The y4m stream header toucan-render writes before the first frame (App.cpp:302-338)."""
    buf = ctypes.create_string_buffer(256)
    _check(lib().tr_y4m_header(int(width), int(height), float(rate), y4m_format.encode(), buf, 256), "tr_y4m_header")
    return buf.value


def frame_block(first: int, last: int, index: int, count: int):
    """toucan::frameBlock -- the scheduler's contiguous block of frames for worker `index` of `count`."""
    lo, hi = ctypes.c_int(), ctypes.c_int()
    f = lib().tr_frame_block
    f.argtypes = [ctypes.c_int] * 4 + [_IP, _IP]
    f.restype = None
    f(first, last, index, count, ctypes.byref(lo), ctypes.byref(hi))
    return lo.value, hi.value


def fit(a, b):
    """toucan::fit -- (min x, min y, width, height) of the box a (bw, bh) image is fitted into inside an (aw, ah) frame."""
    out = (ctypes.c_int * 4)()
    f = lib().tr_fit
    f.argtypes = [ctypes.c_int] * 4 + [ctypes.c_int * 4]
    f.restype = None
    f(int(a[0]), int(a[1]), int(b[0]), int(b[1]), out)
    return tuple(out)


def set_unassociated_alpha(keep: bool) -> None:
    """OIIO's "oiio:UnassociatedAlpha" reader hint for the built-in PNG / TIFF readers (default off = associate on read)."""
    f = lib().tr_set_unassociated_alpha
    f.argtypes = [ctypes.c_int]
    f.restype = None
    f(1 if keep else 0)


def decode_png(path: str) -> np.ndarray:
    """The C++ host's built-in PNG reader (RGBA; alpha associated like OIIO's reader unless set_unassociated_alpha(True))."""
    w, h, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(lib().tr_decode_png(path.encode(), None, 0, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_png")
    out = np.empty((h.value, w.value, 4), _NP_OF[t.value])
    _check(lib().tr_decode_png(path.encode(), out.ctypes.data, out.nbytes, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_png")
    return out


def decode_jpeg(path: str) -> np.ndarray:
    """The C++ host's built-in JPEG reader (RGBA u8, A = 255; libjpeg's default decode, bit for bit)."""
    w, h, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(lib().tr_decode_jpeg(path.encode(), None, 0, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_jpeg")
    out = np.empty((h.value, w.value, 4), _NP_OF[t.value])
    _check(lib().tr_decode_jpeg(path.encode(), out.ctypes.data, out.nbytes, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_jpeg")
    return out


def decode_exr(path: str) -> np.ndarray:
    """The C++ host's built-in OpenEXR reader (RGBA half or float)."""
    w, h, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(lib().tr_decode_exr(path.encode(), None, 0, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_exr")
    out = np.empty((h.value, w.value, 4), _NP_OF[t.value])
    _check(lib().tr_decode_exr(path.encode(), out.ctypes.data, out.nbytes, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_exr")
    return out


def decode_tiff(path: str) -> np.ndarray:
    """The C++ host's built-in TIFF reader (RGBA u8 / u16 / float32)."""
    w, h, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(lib().tr_decode_tiff(path.encode(), None, 0, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_tiff")
    out = np.empty((h.value, w.value, 4), _NP_OF[t.value])
    _check(lib().tr_decode_tiff(path.encode(), out.ctypes.data, out.nbytes, ctypes.byref(w), ctypes.byref(h), ctypes.byref(t)), "tr_decode_tiff")
    return out


def write_exr(path: str, frame: np.ndarray, zip: bool = True):
    """Write an (h, w, 4) float16 / float32 array as a scanline OpenEXR file (ZIP-compressed or uncompressed)."""
    assert frame.ndim == 3 and frame.shape[2] == 4 and frame.flags.c_contiguous and frame.dtype in (np.float16, np.float32)
    _check(lib().tr_write_exr(path.encode(), frame.ctypes.data, frame.shape[1], frame.shape[0], _T_OF[frame.dtype], 1 if zip else 0), "tr_write_exr")


class Host:
    """toucan::ImageEffectHost over the in-tree *.ofx modules."""

    def __init__(self, plugin_dir: str = PLUGIN_DIR):
        self._h = _VP()
        _check(lib().tr_host_create(plugin_dir.encode(), ctypes.byref(self._h)), "tr_host_create")

    def identifiers(self):
        n = lib().tr_host_plugin_count(self._h)
        return [lib().tr_host_plugin_identifier(self._h, i).decode() for i in range(n)]

    def close(self):
        if self._h:
            lib().tr_host_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Timeline:
    """toucan::TimelineWrapper + toucan::ImageGraph + the timeline's media store."""

    def __init__(self, path: str | None = None, json_doc=None, directory: str = ""):
        self._t = _VP()
        self._keep = []  # registered frames are borrowed by the C++ side
        if path is not None:
            _check(lib().tr_timeline_open(path.encode(), ctypes.byref(self._t)), "tr_timeline_open")
        else:
            text = json_doc if isinstance(json_doc, str) else json.dumps(json_doc)
            _check(lib().tr_timeline_open_json(text.encode(), directory.encode(), ctypes.byref(self._t)), "tr_timeline_open_json")

    def media_path(self, url: str) -> str:
        buf = ctypes.create_string_buffer(4096)
        _check(lib().tr_timeline_media_path(self._t, url.encode(), buf, len(buf)), "tr_timeline_media_path")
        return buf.value.decode()

    def set_media(self, path: str, frame):
        """frame: numpy (h, w, 4) host array, or a torch CUDA tensor (kept resident)."""
        if isinstance(frame, np.ndarray):
            assert frame.ndim == 3 and frame.shape[2] == 4 and frame.flags.c_contiguous
            ptr, dev, t = frame.ctypes.data, 0, _T_OF[frame.dtype]
        else:
            import torch

            from .ops import _DT

            assert frame.dim() == 3 and frame.shape[2] == 4 and frame.is_contiguous()
            ptr, dev, t = frame.data_ptr(), (1 if frame.is_cuda else 0), _DT[frame.dtype]
            del torch
        self._keep.append(frame)
        _check(lib().tr_timeline_set_media(self._t, path.encode(), ptr, dev, frame.shape[1], frame.shape[0], t), "tr_timeline_set_media")

    def set_source_cache(self, nbytes: int):
        """Keep up to `nbytes` of uploaded host frames resident per GPU (0 = upload on every frame)."""
        _check(lib().tr_timeline_set_source_cache(self._t, int(nbytes)), "tr_timeline_set_source_cache")

    def upload_stats(self):
        """(uploads, bytes) of source frames copied host -> device so far."""
        n, b = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        _check(lib().tr_timeline_upload_stats(self._t, ctypes.byref(n), ctypes.byref(b)), "tr_timeline_upload_stats")
        return n.value, b.value

    def drop_source_cache(self):
        _check(lib().tr_timeline_drop_source_cache(self._t), "tr_timeline_drop_source_cache")

    def prepare(self):
        _check(lib().tr_timeline_prepare(self._t), "tr_timeline_prepare")

    def info(self):
        s, r = ctypes.c_double(), ctypes.c_double()
        n, w, h, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(-1)
        _check(lib().tr_timeline_info(self._t, ctypes.byref(s), ctypes.byref(r), ctypes.byref(n), ctypes.byref(w), ctypes.byref(h),
                                      ctypes.byref(t)), "tr_timeline_info")
        return {"start": s.value, "rate": r.value, "frames": n.value, "width": w.value, "height": h.value, "type": t.value}

    def describe(self, host: Host, time_value: float) -> str:
        buf = ctypes.create_string_buffer(1 << 20)
        _check(lib().tr_timeline_describe(self._t, host._h, float(time_value), buf, len(buf)), "tr_timeline_describe")
        return buf.value.decode()

    def render(self, host: Host, time_value: float, max_bytes: int | None = None) -> np.ndarray:
        """graph->exec(host, time); node->exec(); download -- returns (h, w, 4) numpy."""
        i = self.info()
        cap = max_bytes or max(16, i["width"] * i["height"] * 16)
        raw = np.empty(cap, np.uint8)
        w, h, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(lib().tr_timeline_render(self._t, host._h, float(time_value), raw.ctypes.data, cap, ctypes.byref(w), ctypes.byref(h),
                                        ctypes.byref(t)), "tr_timeline_render")
        if w.value <= 0 or h.value <= 0:
            return np.zeros((0, 0, 4), np.uint8)
        dt = np.dtype(_NP_OF[t.value])
        n = w.value * h.value * 4
        return raw[: n * dt.itemsize].view(dt).reshape(h.value, w.value, 4).copy()

    def render_resident(self, host: Host, time_value: float):
        """Render and leave the frame in HBM: returns (device_ptr, w, h, tc_type).  Valid
        until the next call; enqueued on the bound thread's stream (sync() before use)."""
        p = _VP()
        w, h, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _check(lib().tr_timeline_render_resident(self._t, host._h, float(time_value), ctypes.byref(p), ctypes.byref(w), ctypes.byref(h),
                                                 ctypes.byref(t)), "tr_timeline_render_resident")
        return p.value, w.value, h.value, t.value

    def render_range(self, first: int, last: int, writer, devices=(0,), frames_in_flight: int = 2, plugin_dir: str = PLUGIN_DIR,
                     raw_format: str | None = None, y4m_format: str | None = None):
        """toucan::FrameScheduler: frames [first, last) on `devices`, writer(frame, ndarray) called in
        order on this thread; the array aliases pinned memory valid only during the call.  raw_format:
        device-side output conversion ("rgb24", "rgba", "rgb48", "rgba64", "rgbaf16", "rgbf32", "rgbaf32").
        y4m_format ("422", "444", "444alpha", "444p16"): device-side conversion to the planar YUV bytes of
        `toucan-render -y4m`; the writer then receives a flat uint8 array (split it with ops.y4m_planes)."""
        def cb(_user, frame, pixels, w, h, t):
            if t & 0x200 and w > 0 and h > 0 and pixels:
                n = int(capi.lib().tc_y4m_frame_bytes(t & 0xf, w, h))
                writer(frame, np.frombuffer((ctypes.c_char * n).from_address(pixels), dtype=np.uint8))
            elif w > 0 and h > 0 and pixels:
                nch = 3 if t & 0x100 else 4
                dt = np.dtype(_NP_OF[t & 0xff])
                buf = (ctypes.c_char * (w * h * nch * dt.itemsize)).from_address(pixels)
                writer(frame, np.frombuffer(buf, dtype=dt).reshape(h, w, nch))
            else:
                writer(frame, np.zeros((0, 0, 4), np.uint8))

        c_cb = FRAME_CALLBACK(cb)
        devs = (ctypes.c_int * len(devices))(*devices)
        if y4m_format:
            _check(lib().tr_render_range_y4m(self._t, plugin_dir.encode(), devs, len(devices), int(first), int(last), int(frames_in_flight),
                                             y4m_format.encode(), c_cb, None), "tr_render_range_y4m")
            return
        _check(lib().tr_render_range_raw(self._t, plugin_dir.encode(), devs, len(devices), int(first), int(last), int(frames_in_flight),
                                         raw_format.encode() if raw_format else None, c_cb, None), "tr_render_range")

    def close(self):
        if self._t:
            lib().tr_timeline_close(self._t)
            self._t = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
