"""Per-op Python front-end over the C ABI, on torch CUDA tensors of shape (h, w, 4).

torch is used for device memory and the current stream only; every function below is
one call into libtoucan_cuda.so (hand-written sm_100a kernels).  No fallback paths.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import capi
from .capi import F16, F32, U8, U16, check, f4, lib, tc_image

_DT = {torch.uint8: U8, torch.uint16: U16, torch.float16: F16, torch.float32: F32}
TORCH_OF = {U8: torch.uint8, U16: torch.uint16, F16: torch.float16, F32: torch.float32}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def img(t: torch.Tensor) -> tc_image:
    if not t.is_cuda:
        raise capi.TcError("toucan_b200 ops need CUDA tensors: there is no CPU fallback")
    assert t.dim() == 3 and t.shape[2] == 4 and t.is_contiguous(), (t.shape, t.is_contiguous())
    return tc_image(t.data_ptr(), t.shape[1], t.shape[0], _DT[t.dtype])


def tcode(t: torch.Tensor) -> int:
    return _DT[t.dtype]


def new(w, h, t, device="cuda"):
    dt = TORCH_OF[t] if isinstance(t, int) else t
    return torch.empty((h, w, 4), dtype=dt, device=device)


def _B(x):
    return ctypes.byref(x)


# generators -----------------------------------------------------------------------------
def fill(w, h, color, t=U8):
    out = new(w, h, t)
    check(lib().tc_fill(_B(img(out)), f4(color), _stream()), "tc_fill")
    return out


def checkers(w, h, cw, ch, c1, c2, t=U8):
    out = new(w, h, t)
    check(lib().tc_checkers(_B(img(out)), int(cw), int(ch), f4(c1), f4(c2), _stream()), "tc_checkers")
    return out


def gradient(w, h, c1, c2, vertical, t=U8):
    out = new(w, h, t)
    check(lib().tc_gradient(_B(img(out)), f4(c1), f4(c2), int(bool(vertical)), _stream()), "tc_gradient")
    return out


def noise(w, h, ntype, a, b, mono, seed, t=U8):
    out = new(w, h, t)
    check(lib().tc_noise(_B(img(out)), ntype.encode(), float(a), float(b), int(bool(mono)), int(seed), _stream()), "tc_noise")
    return out


# filters ------------------------------------------------------------------------------------
def _unary(fn, name, src, *args, size=None, out_type=None):
    h, w = src.shape[:2]
    if size is not None:
        w, h = size
    out = new(w, h, src.dtype if out_type is None else out_type)
    check(fn(_B(img(out)), _B(img(src)), *args, _stream()), name)
    return out


def invert(src):
    return _unary(lib().tc_invert, "tc_invert", src)


def pow_(src, value):
    return _unary(lib().tc_pow, "tc_pow", src, float(value))


def saturate(src, value):
    return _unary(lib().tc_saturate, "tc_saturate", src, float(value))


def color_map(src, name):
    return _unary(lib().tc_color_map, "tc_color_map", src, name.encode())


def premult(src):
    return _unary(lib().tc_premult, "tc_premult", src)


def unpremult(src):
    return _unary(lib().tc_unpremult, "tc_unpremult", src)


def blur(src, radius, out=None):
    if out is None:
        return _unary(lib().tc_blur, "tc_blur", src, float(radius))
    check(lib().tc_blur(_B(img(out)), _B(img(src)), float(radius), _stream()), "tc_blur")
    return out


def unsharp_mask(src, kernel, width, contrast, threshold):
    return _unary(lib().tc_unsharp_mask, "tc_unsharp_mask", src, kernel.encode(), float(width), float(contrast), float(threshold))


def colorconvert(src, fromspace, tospace, unpremult=False):
    return _unary(lib().tc_color_convert2, "tc_color_convert2", src, fromspace.encode(), tospace.encode(), int(bool(unpremult)))


# transforms -----------------------------------------------------------------------------------
def crop(src, pos, size):
    return _unary(lib().tc_crop, "tc_crop", src, int(pos[0]), int(pos[1]), int(size[0]), int(size[1]), size=size)


def flip(src):
    return _unary(lib().tc_flip, "tc_flip", src)


def flop(src):
    return _unary(lib().tc_flop, "tc_flop", src)


def resize(src, size, filter_name="", filter_width=0.0, out_type=None):
    return _unary(lib().tc_resize, "tc_resize", src, filter_name.encode(), float(filter_width), size=size, out_type=out_type)


def rotate(src, angle_degrees, filter_name="", filter_width=0.0):
    import numpy as np

    ang = float(np.float32(float(angle_degrees) / 360.0 * 2.0 * math.pi))
    return _unary(lib().tc_rotate, "tc_rotate", src, ang, filter_name.encode(), float(filter_width))


# transitions / comp ---------------------------------------------------------------------------
def dissolve(frm, to, value, out=None):
    if out is None:
        out = new(frm.shape[1], frm.shape[0], frm.dtype)
    check(lib().tc_dissolve(_B(img(out)), _B(img(frm)), _B(img(to)), float(value), _stream()), "tc_dissolve")
    return out


def wipe(frm, to, value, vertical):
    out = new(frm.shape[1], frm.shape[0], frm.dtype)
    check(lib().tc_wipe(_B(img(out)), _B(img(frm)), _B(img(to)), float(value), int(bool(vertical)), _stream()), "tc_wipe")
    return out


def over(fg, bg, premult_fg=False, out=None):
    if out is None:
        out = new(bg.shape[1], bg.shape[0], lib().tc_type_merge(tcode(fg), tcode(bg)))
    check(lib().tc_over(_B(img(out)), _B(img(fg)), _B(img(bg)), int(bool(premult_fg)), _stream()), "tc_over")
    return out


def over_fill(fg, color, bg_type=U8, premult_fg=False):
    out_t = capi.lib().tc_type_merge(tcode(fg), bg_type)
    return _unary(lib().tc_over_fill, "tc_over_fill", fg, f4(color), bg_type, 1 if premult_fg else 0, out_type=TORCH_OF[out_t])


def paste_on_black(src, size, ox, oy, t):
    out = new(size[0], size[1], t)
    check(lib().tc_paste_on_black(_B(img(out)), int(ox), int(oy), _B(img(src)), _stream()), "tc_paste_on_black")
    return out


def convert(src, t, size=None):
    h, w = src.shape[:2]
    if size is not None:
        w, h = size
    out = new(w, h, t)
    check(lib().tc_convert(_B(img(out)), _B(img(src)), _stream()), "tc_convert")
    return out


def pointwise_chain(src, steps):
    """tc_pointwise_chain: steps = [("invert",), ("pow", 2.0), ("saturate", 0.5), ("color_map", "plasma"), ("premult",),
    ("unpremult",), ("color_convert", from, to, unpremult)], first step first."""
    arr = (capi.tc_chain_op * len(steps))()
    for i, st in enumerate(steps):
        arr[i].kind = capi.CHAIN_KINDS[st[0]]
        if st[0] in ("pow", "saturate"):
            arr[i].value = float(st[1])
        elif st[0] == "color_map":
            arr[i].name_a = st[1].encode()
        elif st[0] == "color_convert":
            arr[i].name_a, arr[i].name_b = st[1].encode(), st[2].encode()
            arr[i].flag = 1 if (len(st) > 3 and st[3]) else 0
        elif st[0] == "blur":  # ("blur", radius): only as the first step
            arr[i].value = float(st[1])
        elif st[0] == "unsharp":  # ("unsharp", kernel, width, contrast, threshold): only as the first step
            arr[i].name_a = st[1].encode()
            arr[i].value = float(st[2])
            arr[i].color[0], arr[i].color[1] = float(st[3]), float(st[4])
        elif st[0] == "over_fill":  # ("over_fill", colour, premult): CompNode over the Fill generator's frame
            for c in range(4):
                arr[i].color[c] = float(st[1][c])
            arr[i].flag = 1 if st[2] else 0
            arr[i].w, arr[i].h = int(src.shape[1]), int(src.shape[0])
    return _unary(lib().tc_pointwise_chain, "tc_pointwise_chain", src, arr, len(steps))


def draw_box(src, pos1, pos2, color, fill=True):
    return _unary(lib().tc_draw_box, "tc_draw_box", src, int(pos1[0]), int(pos1[1]), int(pos2[0]), int(pos2[1]), f4(color), 1 if fill else 0)


def draw_line(src, pos1, pos2, color, skip_first_point=False):
    return _unary(lib().tc_draw_line, "tc_draw_line", src, int(pos1[0]), int(pos1[1]), int(pos2[0]), int(pos2[1]), f4(color),
                  1 if skip_first_point else 0)


def y4m_planes(buf, fmt, w, h):
    """Split the bytes of one y4m frame (host or device uint8 tensor / numpy array) into its planes."""
    cw = (w + 1) // 2 if fmt == "422" else w
    if fmt == "444p16":
        buf = buf.view(torch.uint16) if isinstance(buf, torch.Tensor) else buf.view("<u2")
    shapes = [(h, w), (h, cw), (h, cw)] + ([(h, w)] if fmt == "444alpha" else [])
    out, o = [], 0
    for sh in shapes:
        n = sh[0] * sh[1]
        out.append(buf[o:o + n].reshape(sh))
        o += n
    return out


def rgb_to_y4m(src, fmt):
    """tc_rgb_to_y4m: the frame's y4m bytes (App.cpp:340-485) as a flat uint8 device tensor."""
    h, w = src.shape[:2]
    f = capi.Y4M_FORMATS[fmt]
    n = int(lib().tc_y4m_frame_bytes(f, w, h))
    out = torch.empty((n,), dtype=torch.uint8, device=src.device)
    check(lib().tc_rgb_to_y4m(ctypes.c_void_p(out.data_ptr()), _B(img(src)), f, _stream()), "tc_rgb_to_y4m")
    return out


def stack_f32(w, h, background, layers, out=None):
    """layers: list of (src_a, src_b|None, radius_a, radius_b, value), bottom track first."""
    if out is None:
        out = new(w, h, F32)
    arr = (capi.tc_stack_layer * len(layers))()
    for i, (a, b, ra, rb, v) in enumerate(layers):
        arr[i].src_a = a.data_ptr()
        arr[i].src_b = b.data_ptr() if b is not None else None
        arr[i].radius_a = float(ra)
        arr[i].radius_b = float(rb)
        arr[i].value = float(v)
    check(lib().tc_stack_f32(_B(img(out)), f4(background), arr, len(layers), _stream()), "tc_stack_f32")
    return out
