"""ctypes binding of libtoucan_cuda.so (include/toucan_cuda.h).

Thin plumbing only: device buffers are torch CUDA tensors (or raw pointers), every op is
the hand-written kernel behind the C ABI.  There is NO fallback: if the library is
missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
CUDA_LIB = os.path.join(LIB_DIR, "libtoucan_cuda.so")

U8, U16, F16, F32 = 0, 1, 2, 3
Y4M_FORMATS = {"422": 0, "444": 1, "444alpha": 2, "444p16": 3}  # bin/toucan-render/App.cpp:37-43
TYPE_NAMES = {U8: "u8", U16: "u16", F16: "half", F32: "float"}


class TcError(RuntimeError):
    pass


class tc_image(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("w", ctypes.c_int), ("h", ctypes.c_int), ("type", ctypes.c_int)]


class tc_chain_op(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("value", ctypes.c_float), ("flag", ctypes.c_int), ("name_a", ctypes.c_char * 96),
                ("name_b", ctypes.c_char * 96), ("color", ctypes.c_float * 4), ("w", ctypes.c_int), ("h", ctypes.c_int)]


CHAIN_KINDS = {"invert": 0, "pow": 1, "saturate": 2, "color_map": 3, "premult": 4, "unpremult": 5, "color_convert": 6,
               "over_fill": 8, "flip": 9, "flop": 10, "blur": 11, "unsharp": 12}


class tc_stack_layer(ctypes.Structure):
    _fields_ = [("src_a", ctypes.c_void_p), ("src_b", ctypes.c_void_p), ("radius_a", ctypes.c_float),
                ("radius_b", ctypes.c_float), ("value", ctypes.c_double)]


_lib = None

_IMG = ctypes.POINTER(tc_image)
_F4 = ctypes.POINTER(ctypes.c_float)
_S = ctypes.c_void_p
_SIGS = {
    "tc_fill": [_IMG, _F4, _S],
    "tc_checkers": [_IMG, ctypes.c_int, ctypes.c_int, _F4, _F4, _S],
    "tc_gradient": [_IMG, _F4, _F4, ctypes.c_int, _S],
    "tc_noise": [_IMG, ctypes.c_char_p, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, _S],
    "tc_blur": [_IMG, _IMG, ctypes.c_float, _S],
    "tc_color_map": [_IMG, _IMG, ctypes.c_char_p, _S],
    "tc_invert": [_IMG, _IMG, _S],
    "tc_pow": [_IMG, _IMG, ctypes.c_float, _S],
    "tc_saturate": [_IMG, _IMG, ctypes.c_float, _S],
    "tc_unsharp_mask": [_IMG, _IMG, ctypes.c_char_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, _S],
    "tc_color_convert": [_IMG, _IMG, ctypes.c_char_p, ctypes.c_char_p, _S],
    "tc_color_convert2": [_IMG, _IMG, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, _S],
    "tc_premult": [_IMG, _IMG, _S],
    "tc_unpremult": [_IMG, _IMG, _S],
    "tc_crop": [_IMG, _IMG, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _S],
    "tc_flip": [_IMG, _IMG, _S],
    "tc_flop": [_IMG, _IMG, _S],
    "tc_resize": [_IMG, _IMG, ctypes.c_char_p, ctypes.c_float, _S],
    "tc_rotate": [_IMG, _IMG, ctypes.c_float, ctypes.c_char_p, ctypes.c_float, _S],
    "tc_dissolve": [_IMG, _IMG, _IMG, ctypes.c_double, _S],
    "tc_wipe": [_IMG, _IMG, _IMG, ctypes.c_double, ctypes.c_int, _S],
    "tc_over": [_IMG, _IMG, _IMG, ctypes.c_int, _S],
    "tc_over_fill": [_IMG, _IMG, _F4, ctypes.c_int, ctypes.c_int, _S],
    "tc_paste_on_black": [_IMG, ctypes.c_int, ctypes.c_int, _IMG, _S],
    "tc_convert": [_IMG, _IMG, _S],
    "tc_pack_raw": [ctypes.c_void_p, _IMG, ctypes.c_int, ctypes.c_int, _S],
    "tc_rgb_to_y4m": [ctypes.c_void_p, _IMG, ctypes.c_int, _S],
    "tc_pointwise_chain": [_IMG, _IMG, ctypes.POINTER(tc_chain_op), ctypes.c_int, _S],
    "tc_draw_box": [_IMG, _IMG, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F4, ctypes.c_int, _S],
    "tc_draw_line": [_IMG, _IMG, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F4, ctypes.c_int, _S],
    "tc_stack_f32": [_IMG, _F4, ctypes.POINTER(tc_stack_layer), ctypes.c_int, _S],
    "tc_set_device": [ctypes.c_int],
    "tc_device_count": [ctypes.POINTER(ctypes.c_int)],
    "tc_stream_sync": [_S],
    "tc_malloc": [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, _S],
    "tc_free": [ctypes.c_void_p, _S],
    "tc_free_sync": [ctypes.c_void_p],
    "tc_upload": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, _S],
    "tc_download": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, _S],
    "tc_host_alloc": [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t],
    "tc_host_free": [ctypes.c_void_p],
    "tc_host_register": [ctypes.c_void_p, ctypes.c_size_t],
    "tc_host_unregister": [ctypes.c_void_p],
    "tc_type_merge": [ctypes.c_int, ctypes.c_int],
    "tc_event_create": [ctypes.POINTER(ctypes.c_void_p)],
    "tc_event_create_timing": [ctypes.POINTER(ctypes.c_void_p)],
    "tc_event_elapsed_ms": [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)],
    "tc_event_destroy": [ctypes.c_void_p],
    "tc_event_record": [ctypes.c_void_p, _S],
    "tc_event_sync": [ctypes.c_void_p],
    "tc_stream_wait_event": [_S, ctypes.c_void_p],
    "tc_color_space_known": [ctypes.c_char_p],
}

# every symbol include/toucan_cuda.h declares (checked by tests/test_abi.py)
EXPORTS = sorted(set(_SIGS) | {"tc_version", "tc_last_error", "tc_stream_create", "tc_stream_destroy", "tc_type_size",
                               "tc_image_bytes", "tc_pool_trim", "tc_memset", "tc_launch_count", "tc_y4m_frame_bytes", "tc_alloc_stats"})


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(CUDA_LIB):
            raise TcError(f"{CUDA_LIB} is missing: run `python -m toucan_b200.build` (there is no CPU fallback)")
        _lib = ctypes.CDLL(CUDA_LIB, mode=ctypes.RTLD_GLOBAL)
        for name, sig in _SIGS.items():
            fn = getattr(_lib, name)
            fn.argtypes = sig
            fn.restype = ctypes.c_int
        _lib.tc_version.restype = ctypes.c_char_p
        _lib.tc_last_error.restype = ctypes.c_char_p
        _lib.tc_launch_count.restype = ctypes.c_ulonglong
        _lib.tc_type_size.restype = ctypes.c_size_t
        _lib.tc_image_bytes.restype = ctypes.c_size_t
        _lib.tc_y4m_frame_bytes.restype = ctypes.c_size_t
        _lib.tc_y4m_frame_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        raise TcError(f"{what} failed ({status}): {lib().tc_last_error().decode()}")


def f4(c):
    return (ctypes.c_float * 4)(*[float(v) for v in c])


def alloc_stats():
    """(calls, seconds) spent in the stream-ordered allocator so far."""
    c, ns = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    lib().tc_alloc_stats(ctypes.byref(c), ctypes.byref(ns))
    return c.value, ns.value * 1e-9


def launch_count() -> int:
    return int(lib().tc_launch_count())
